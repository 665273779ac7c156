"""Equations, fluxes, initial conditions, boundary conditions and source terms as named singletons.

In the reference these are Trixi.jl types/functions passed through `DGSEMGPU(...)` and
`SemidiscretizationHyperbolicGPU(...)` and inlined into CUDA.jl kernels (call sites: reference
src/solvers/dg_3d_kernel.jl:93-95,226-234,1166,1190-1191,1327-1343,1836). Across a C ABI they become enums
(include/trixib200.h); anything outside the enumerated set raises instead of silently falling back.
Initial conditions also carry a vectorised numpy evaluation used by `compute_coefficients` on the host
(the shim-side alternative the reference itself notes at src/solvers/solvers.jl:49-51).
"""
import numpy as np

from . import _lib


class _Named:
    def __init__(self, name, code=None):
        self.name, self.code = name, code

    def __repr__(self):
        return self.name


# ---------------------------------------------------------------------------------------------- fluxes
class Flux(_Named):
    pass


flux_central = Flux("flux_central", _lib.FLUX["flux_central"])
flux_lax_friedrichs = Flux("flux_lax_friedrichs", _lib.FLUX["flux_lax_friedrichs"])
flux_hll = Flux("flux_hll", _lib.FLUX["flux_hll"])
flux_ranocha = Flux("flux_ranocha", _lib.FLUX["flux_ranocha"])
flux_shima_etal = Flux("flux_shima_etal", _lib.FLUX["flux_shima_etal"])
flux_hindenlang_gassner = Flux("flux_hindenlang_gassner", _lib.FLUX["flux_hindenlang_gassner"])
flux_hlle = Flux("flux_hlle", _lib.FLUX["flux_hlle"])
flux_nonconservative_powell = _Named("flux_nonconservative_powell")
max_abs_speed_naive = _Named("max_abs_speed_naive")
max_abs_speed = _Named("max_abs_speed")
min_max_speed_naive = _Named("min_max_speed_naive")
min_max_speed_davis = _Named("min_max_speed_davis")
min_max_speed_einfeldt = _Named("min_max_speed_einfeldt")


def FluxLaxFriedrichs(max_abs_speed_fn=max_abs_speed):
    """`FluxLaxFriedrichs(max_abs_speed_naive)` is the Trixi <= 0.12 `flux_lax_friedrichs`."""
    if max_abs_speed_fn is max_abs_speed_naive:
        return Flux("FluxLaxFriedrichs(max_abs_speed_naive)", _lib.FLUX["flux_lax_friedrichs_naive"])
    if max_abs_speed_fn is max_abs_speed:
        return flux_lax_friedrichs
    raise NotImplementedError(f"FluxLaxFriedrichs({max_abs_speed_fn}) is not an enumerated libtrixib200 flux")


def FluxHLL(min_max_speed=min_max_speed_davis):
    if min_max_speed is min_max_speed_naive:
        return Flux("FluxHLL(min_max_speed_naive)", _lib.FLUX["flux_hll_naive"])
    if min_max_speed is min_max_speed_davis:
        return flux_hll
    if min_max_speed is min_max_speed_einfeldt:
        return flux_hlle
    raise NotImplementedError(f"FluxHLL({min_max_speed}) is not an enumerated libtrixib200 flux")


def split_flux(f):
    """surface_flux / volume_flux may be `flux` or `(flux, flux_nonconservative_powell)`."""
    if isinstance(f, (tuple, list)):
        if len(f) != 2 or f[1] is not flux_nonconservative_powell:
            raise NotImplementedError("only (flux, flux_nonconservative_powell) tuples are enumerated")
        return f[0], True
    return f, False


# ---------------------------------------------------------------------------------------------- equations
class AbstractEquations:
    kind = None
    ndim = None
    nvars = None
    gamma = 1.4
    advection_velocity = (0.0, 0.0, 0.0)
    c_h = 0.0

    def have_constant_speed(self):
        return False


class _Advection(AbstractEquations):
    kind = _lib.EQ_ADVECTION
    nvars = 1

    def __init__(self, a):
        a = tuple(np.atleast_1d(np.asarray(a, dtype=np.float64)).tolist())
        if len(a) != self.ndim:
            raise ValueError("advection velocity has the wrong dimension")
        self.advection_velocity = a + (0.0,) * (3 - len(a))

    def have_constant_speed(self):
        return True


class LinearScalarAdvectionEquation1D(_Advection):
    ndim = 1


class LinearScalarAdvectionEquation2D(_Advection):
    ndim = 2


class LinearScalarAdvectionEquation3D(_Advection):
    ndim = 3


class _Euler(AbstractEquations):
    kind = _lib.EQ_EULER

    def __init__(self, gamma):
        self.gamma = float(gamma)
        self.nvars = self.ndim + 2

    def prim2cons(self, q):
        nd = self.ndim
        u = np.empty_like(q)
        u[0] = q[0]
        ke = 0.0
        for d in range(nd):
            u[1 + d] = q[0] * q[1 + d]
            ke = ke + u[1 + d] * q[1 + d]
        u[nd + 1] = q[nd + 1] / (self.gamma - 1) + 0.5 * ke
        return u


class CompressibleEulerEquations1D(_Euler):
    ndim = 1


class CompressibleEulerEquations2D(_Euler):
    ndim = 2


class CompressibleEulerEquations3D(_Euler):
    ndim = 3


class IdealGlmMhdEquations3D(AbstractEquations):
    kind = _lib.EQ_MHD
    ndim = 3
    nvars = 9

    def __init__(self, gamma, initial_c_h=float("nan")):
        self.gamma = float(gamma)
        self.c_h = float(initial_c_h)   # Trixi default is NaN until GlmSpeedCallback sets it

    def prim2cons(self, q):
        u = np.empty_like(q)
        u[0] = q[0]
        for d in range(3):
            u[1 + d] = q[0] * q[1 + d]
        u[5:9] = q[5:9]
        u[4] = (q[4] / (self.gamma - 1) + 0.5 * (u[1] * q[1] + u[2] * q[2] + u[3] * q[3])
                + 0.5 * (q[5] ** 2 + q[6] ** 2 + q[7] ** 2) + 0.5 * q[8] ** 2)
        return u


# ---------------------------------------------------------------------------------------------- initial conditions
class InitialCondition(_Named):
    """Callable `ic(x, t, equations)` with x of shape [ndim, ...] -> conservative variables [nvars, ...]."""

    def __init__(self, name, code, fn):
        super().__init__(name, code)
        self._fn = fn

    def __call__(self, x, t, equations):
        return self._fn(np.asarray(x, dtype=np.float64), float(t), equations)


def _ic_constant(x, t, eq):
    """Trixi `initial_condition_constant`: a CONSERVATIVE state (Euler (1, 0.1[, -0.2[, 0.7]], 10); GLM-MHD
    (1, 0.1, -0.2, -0.5, 50, 3, -1.2, 0.5, 0); advection 2)."""
    shape = x.shape[1:]
    if eq.kind == _lib.EQ_ADVECTION:
        return np.full((1,) + shape, 2.0)
    if eq.kind == _lib.EQ_EULER:
        vals = [1.0] + [0.1, -0.2, 0.7][: eq.ndim] + [10.0]
    else:
        vals = [1.0, 0.1, -0.2, -0.5, 50.0, 3.0, -1.2, 0.5, 0.0]
    u = np.empty((eq.nvars,) + shape)
    for v, val in enumerate(vals):
        u[v] = val
    return u


def _ic_convergence_test(x, t, eq):
    nd = eq.ndim
    if eq.kind == _lib.EQ_ADVECTION:
        s = 0.0
        for d in range(nd):
            s = s + (x[d] - eq.advection_velocity[d] * t)
        c, A, L = 1.0, 0.5, 2.0
        omega = 2 * np.pi * (1 / L)
        return (c + A * np.sin(omega * s))[None]
    if eq.kind == _lib.EQ_EULER:
        c, A, L = 2.0, 0.1, 2.0
        omega = 2 * np.pi * (1 / L)
        s = -t
        for d in range(nd):
            s = s + x[d]
        ini = c + A * np.sin(omega * s)
        u = np.empty((nd + 2,) + x.shape[1:])
        u[: nd + 1] = ini
        u[nd + 1] = ini * ini
        return u
    # Alfven wave
    omega, r, e = 2.0 * np.pi, 2.0, 0.2
    nx, ny = 1 / np.sqrt(r * r + 1.0), r / np.sqrt(r * r + 1.0)
    sqr = 1.0
    Va = omega / (ny * sqr)
    phi = omega / ny * (nx * (x[0] - 0.5 * r) + ny * (x[1] - 0.5 * r)) - Va * t
    q = np.empty((9,) + x.shape[1:])
    q[0] = 1.0
    q[1] = -e * ny * np.cos(phi) / q[0]
    q[2] = e * nx * np.cos(phi) / q[0]
    q[3] = e * np.sin(phi) / q[0]
    q[4] = 1.0
    q[5] = nx - q[0] * q[1] * sqr
    q[6] = ny - q[0] * q[2] * sqr
    q[7] = -q[0] * q[3] * sqr
    q[8] = 0.0
    return eq.prim2cons(q)


def _ic_weak_blast_wave(x, t, eq):
    nd = eq.ndim
    if eq.kind == _lib.EQ_ADVECTION:
        raise NotImplementedError("initial_condition_weak_blast_wave needs Euler or MHD")
    r2 = 0.0
    for d in range(nd):
        r2 = r2 + x[d] * x[d]
    r = np.sqrt(r2)
    out = r > 0.5
    nv = eq.nvars
    q = np.empty((nv,) + x.shape[1:])
    q[0] = np.where(out, 1.0, 1.1691)
    if nd == 1:
        q[1] = np.where(out, 0.0, 0.1882 * np.where(x[0] > 0, 1.0, -1.0))
    elif nd == 2:
        phi = np.arctan2(x[1], x[0])
        q[1] = np.where(out, 0.0, 0.1882 * np.cos(phi))
        q[2] = np.where(out, 0.0, 0.1882 * np.sin(phi))
    else:
        phi = np.arctan2(x[1], x[0])
        with np.errstate(invalid="ignore", divide="ignore"):
            theta = np.where(r == 0.0, 0.0, np.arccos(x[2] / np.where(r == 0.0, 1.0, r)))
        q[1] = np.where(out, 0.0, 0.1882 * np.cos(phi) * np.sin(theta))
        q[2] = np.where(out, 0.0, 0.1882 * np.sin(phi) * np.sin(theta))
        q[3] = np.where(out, 0.0, 0.1882 * np.cos(theta))
    if eq.kind == _lib.EQ_EULER:
        q[nd + 1] = np.where(out, 1.0, 1.245)
    else:
        q[4] = np.where(out, 1.0, 1.245)
        q[5], q[6], q[7], q[8] = 1.0, 1.0, 1.0, 0.0
    return eq.prim2cons(q)


def _ic_density_wave(x, t, eq):
    """Trixi `initial_condition_density_wave` (1D: v = 0.1, 2D: v = (0.1, 0.2)): rho = 1 + 0.98 sinpi(2 (sum x - t sum v)),
    p = 20. Trixi has no 3D method: the 3D case continues the pattern with v3 = 0.3 (smooth synthetic timing state)."""
    if eq.kind != _lib.EQ_EULER:
        raise NotImplementedError("initial_condition_density_wave needs compressible Euler")
    nd = eq.ndim
    v = (0.1, 0.2, 0.3)[:nd]
    s = 0.0
    for d in range(nd):
        s = s + x[d]
    q = np.empty((nd + 2,) + x.shape[1:])
    q[0] = 1 + 0.98 * np.sin(np.pi * (2 * (s - t * sum(v))))
    for d in range(nd):
        q[1 + d] = v[d]
    q[nd + 1] = 20.0
    return eq.prim2cons(q)


initial_condition_constant = InitialCondition("initial_condition_constant", _lib.IC["constant"], _ic_constant)
initial_condition_convergence_test = InitialCondition("initial_condition_convergence_test",
                                                      _lib.IC["convergence_test"], _ic_convergence_test)
initial_condition_weak_blast_wave = InitialCondition("initial_condition_weak_blast_wave",
                                                     _lib.IC["weak_blast_wave"], _ic_weak_blast_wave)
initial_condition_density_wave = InitialCondition("initial_condition_density_wave", _lib.IC["density_wave"],
                                                  _ic_density_wave)

# ---------------------------------------------------------------------------------------------- BCs / sources
boundary_condition_periodic = _Named("boundary_condition_periodic", _lib.BC_PERIODIC)


class BoundaryConditionDirichlet(_Named):
    def __init__(self, boundary_value_function):
        if not isinstance(boundary_value_function, InitialCondition):
            raise NotImplementedError("BoundaryConditionDirichlet needs an enumerated initial condition "
                                      "(a C ABI cannot take closures)")
        super().__init__(f"BoundaryConditionDirichlet({boundary_value_function.name})", _lib.BC_DIRICHLET_IC)
        self.boundary_value_function = boundary_value_function


# `boundary_condition_slip_wall` (compressible Euler): enumerated, evaluated on the device
boundary_condition_slip_wall = _Named("boundary_condition_slip_wall", _lib.BC_SLIP_WALL)

source_terms_convergence_test = _Named("source_terms_convergence_test", _lib.SRC["convergence_test"])

# indicator variables
density = _Named("density", _lib.IND["density"])
pressure = _Named("pressure", _lib.IND["pressure"])
density_pressure = _Named("density_pressure", _lib.IND["density_pressure"])
