"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

Parity status: "parity unpinned" (no Julia/Trixi.jl available; see oracle/ORACLE_ASSUMPTIONS.md).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (trixicuda.jl_b200/) must never import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

# enums (mirror oracle/equations.hpp, oracle/dg.hpp)
EQ = {"advection": 0, "euler": 1, "mhd": 2}
FLUX = {"flux_central": 0, "flux_lax_friedrichs": 1, "flux_lax_friedrichs_naive": 2, "flux_hll": 3,
        "flux_hll_naive": 4, "flux_ranocha": 5, "flux_shima_etal": 6, "flux_hindenlang_gassner": 7,
        "flux_hlle": 8}
VI = {"weak_form": 0, "flux_differencing": 1, "shock_capturing_hg": 2}
IC = {"constant": 0, "convergence_test": 1, "weak_blast_wave": 2, "density_wave": 3}
SRC = {"none": 0, "convergence_test": 1}
IND = {"density": 0, "pressure": 1, "density_pressure": 2}
BC = {"periodic": 0, "dirichlet_ic": 1, "slip_wall": 2}


class OrcConfig(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("eq_kind", C.c_int), ("polydeg", C.c_int),
        ("volume_integral", C.c_int), ("volume_flux", C.c_int), ("volume_flux_fv", C.c_int),
        ("surface_flux", C.c_int), ("nonconservative", C.c_int),
        ("alpha_smooth", C.c_int), ("indicator_variable", C.c_int), ("initial_condition", C.c_int),
        ("source", C.c_int),
        ("bc", C.c_int * 6),
        ("initial_refinement_level", C.c_int),
        ("periodicity", C.c_int * 3),
        ("n_patches", C.c_int),
        ("gamma", C.c_double), ("advection_velocity", C.c_double * 3), ("c_h", C.c_double),
        ("alpha_max", C.c_double), ("alpha_min", C.c_double),
        ("coordinates_min", C.c_double * 3), ("coordinates_max", C.c_double * 3),
        ("patch_lo", (C.c_double * 3) * 4), ("patch_hi", (C.c_double * 3) * 4),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in ("oracle_capi.cpp", "dg.hpp", "tree.hpp", "equations.hpp", "basis.hpp")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        for f in (L.orc_size, L.orc_len_f64, L.orc_len_i64):
            f.restype = C.c_longlong
            f.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_f64.restype = C.c_longlong
        L.orc_get_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]
        L.orc_get_i64.restype = C.c_longlong
        L.orc_get_i64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]
        L.orc_compute_coefficients.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.orc_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_stage.restype = C.c_int
        L.orc_stage.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_max_dt.restype = C.c_double
        L.orc_max_dt.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_error_norms.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_entropy_rate.restype = C.c_double
        L.orc_entropy_rate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_solve_ck2n54.restype = C.c_longlong
        L.orc_solve_ck2n54.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_longlong]
        L.orc_two_point_flux.restype = C.c_int
        L.orc_two_point_flux.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_double, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_noncons_powell.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_time_rhs.restype = C.c_double
        L.orc_time_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One semidiscretization (mesh + equations + solver) evaluated by the CPU oracle."""

    def __init__(self, ndim=3, equations="euler", polydeg=3, volume_integral="weak_form",
                 volume_flux="flux_central", volume_flux_fv="flux_lax_friedrichs",
                 surface_flux="flux_lax_friedrichs", nonconservative=False,
                 alpha_max=0.5, alpha_min=0.001, alpha_smooth=True, indicator_variable="density_pressure",
                 initial_condition="convergence_test", source="none", bc=("periodic",) * 6,
                 gamma=1.4, advection_velocity=(1.0, 1.0, 1.0), c_h=1.0,
                 coordinates_min=(-1.0, -1.0, -1.0), coordinates_max=(1.0, 1.0, 1.0),
                 initial_refinement_level=2, periodicity=(True, True, True), refinement_patches=()):
        cfg = OrcConfig()
        cfg.ndim, cfg.eq_kind, cfg.polydeg = ndim, EQ[equations], polydeg
        cfg.volume_integral = VI[volume_integral]
        cfg.volume_flux, cfg.volume_flux_fv = FLUX[volume_flux], FLUX[volume_flux_fv]
        cfg.surface_flux, cfg.nonconservative = FLUX[surface_flux], int(nonconservative)
        cfg.alpha_smooth, cfg.indicator_variable = int(alpha_smooth), IND[indicator_variable]
        cfg.initial_condition, cfg.source = IC[initial_condition], SRC[source]
        for i in range(6):
            cfg.bc[i] = BC[bc[i]] if i < len(bc) else 0
        cfg.initial_refinement_level = initial_refinement_level
        per = tuple(periodicity) if hasattr(periodicity, "__len__") else (periodicity,) * 3
        for d in range(3):
            cfg.periodicity[d] = int(per[d]) if d < len(per) else 1
            cfg.advection_velocity[d] = advection_velocity[d] if d < len(advection_velocity) else 0.0
            cfg.coordinates_min[d] = coordinates_min[d] if d < ndim else 0.0
            cfg.coordinates_max[d] = coordinates_max[d] if d < ndim else 0.0
        cfg.n_patches = len(refinement_patches)
        for p, (lo, hi) in enumerate(refinement_patches):
            for d in range(ndim):
                cfg.patch_lo[p][d], cfg.patch_hi[p][d] = lo[d], hi[d]
        cfg.gamma, cfg.c_h, cfg.alpha_max, cfg.alpha_min = gamma, c_h, alpha_max, alpha_min
        self._L = lib()
        self._h = self._L.orc_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("oracle: " + self._L.orc_last_error().decode())
        self.ndim = ndim
        self.nvars = self.size("nvars")
        self.nnodes = self.size("nnodes")
        self.nelements = self.size("nelements")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    def size(self, name):
        return int(self._L.orc_size(self._h, name.encode()))

    def f64(self, name):
        n = self._L.orc_len_f64(self._h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self._L.orc_get_f64(self._h, name.encode(), _ptr(out), n)
        return out

    def i64(self, name):
        n = self._L.orc_len_i64(self._h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int64)
        self._L.orc_get_i64(self._h, name.encode(), _ptr(out), n)
        return out

    def new_u(self):
        return np.zeros(self.size("nunknowns"), dtype=np.float64)

    def compute_coefficients(self, t=0.0):
        u = self.new_u()
        self._L.orc_compute_coefficients(self._h, t, _ptr(u))
        return u

    def rhs(self, u, t=0.0):
        du = self.new_u()
        self._L.orc_rhs(self._h, _ptr(du), _ptr(u), t)
        return du

    def stage(self, name, du, u, t=0.0):
        if self._L.orc_stage(self._h, name.encode(), _ptr(du), _ptr(u), t) != 0:
            raise KeyError(name)

    def max_dt(self, u):
        return float(self._L.orc_max_dt(self._h, _ptr(u)))

    def error_norms(self, u, t):
        l2 = np.zeros(self.nvars)
        linf = np.zeros(self.nvars)
        self._L.orc_error_norms(self._h, _ptr(u), t, _ptr(l2), _ptr(linf))
        return l2, linf

    def integrate(self, u):
        out = np.zeros(self.nvars)
        self._L.orc_integrate(self._h, _ptr(u), _ptr(out))
        return out

    def entropy_rate(self, du, u):
        return float(self._L.orc_entropy_rate(self._h, _ptr(du), _ptr(u)))

    def indicator(self, u):
        du = self.new_u()
        self.stage("calc_indicator", du, u)
        return self.f64("alpha")

    def solve(self, u, t0, t1, cfl=1.0, dt=0.0, max_steps=10**9):
        u = u.copy()
        steps = self._L.orc_solve_ck2n54(self._h, _ptr(u), t0, t1, cfl, dt, max_steps)
        return u, int(steps)

    def time_rhs(self, u, warm=1, reps=3, t=0.0):
        du = self.new_u()
        return float(self._L.orc_time_rhs(self._h, _ptr(du), _ptr(u), t, warm, reps))


def two_point_flux(equations, ndim, flux, ul, ur, orientation, gamma=1.4, adv=(1.0, 1.0, 1.0), c_h=1.0):
    ul = np.ascontiguousarray(ul, dtype=np.float64)
    ur = np.ascontiguousarray(ur, dtype=np.float64)
    f = np.zeros_like(ul)
    a = np.asarray(list(adv) + [0.0] * (3 - len(adv)), dtype=np.float64)
    rc = lib().orc_two_point_flux(EQ[equations], ndim, gamma, _ptr(a), c_h, FLUX[flux], _ptr(ul), _ptr(ur),
                                  orientation, _ptr(f))
    if rc != 0:
        raise KeyError(flux)
    return f


def noncons_powell(ul, ur, orientation, gamma=5 / 3, c_h=1.0):
    ul = np.ascontiguousarray(ul, dtype=np.float64)
    ur = np.ascontiguousarray(ur, dtype=np.float64)
    f = np.zeros_like(ul)
    lib().orc_noncons_powell(gamma, c_h, _ptr(ul), _ptr(ur), orientation, _ptr(f))
    return f


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())
