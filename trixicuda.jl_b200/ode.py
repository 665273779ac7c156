"""Minimal time integration around `rhs_gpu_`: CarpenterKennedy2N54, StepsizeCallback, AnalysisCallback.

In the Julia drop-in these come from OrdinaryDiffEq / Trixi unchanged (reference examples/euler_ec_3d.jl:25-55);
this module only lets the Python mirror (and the parity tests) run the same call sequence:
`solve(ode, CarpenterKennedy2N54(williamson_condition=False), dt=..., callback=CallbackSet(...))`.
Error norms follow reference src/callbacks_step/analysis_dg_3d.jl:45-89 (host copy, 2p analyzer).
"""
import numpy as np

from .basis import SolutionAnalyzer

_A = (0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
      -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0)
_B = (1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
      3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0)
_C = (0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
      2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0)


class CarpenterKennedy2N54:
    def __init__(self, williamson_condition=False):
        self.williamson_condition = williamson_condition


class StepsizeCallback:
    """`StepsizeCallback(cfl=...)`: dt = cfl * max_dt(u) before the first and after every step."""

    def __init__(self, cfl=1.0):
        self.cfl = float(cfl)


def calculate_dt(u_ode, t, cfl_number, semi):
    """Trixi `calculate_dt(u_ode, t, cfl_number, semi)` (callbacks_step/stepsize.jl), the body of StepsizeCallback:
    the argument tuple below is what reaches the reference's / the shim's `max_dt` method."""
    from .semidiscretization import mesh_equations_solver_cache, wrap_array, max_dt
    mesh, equations, solver, cache = mesh_equations_solver_cache(semi)
    u = wrap_array(u_ode, mesh, equations, solver, cache)
    return cfl_number * max_dt(u, t, mesh, equations.have_constant_speed(), equations, solver, cache)


class AnalysisCallback:
    """`AnalysisCallback(semi, interval=...)`. on_device=True evaluates the error norms with the library's reduction
    kernels (trixib200_calc_error_norms; enumerated initial conditions, any number of ranks) instead of copying u to
    the host as the reference does (reference src/callbacks_step/analysis_dg_3d.jl:45-89)."""

    def __init__(self, semi, interval=0, on_device=False):
        self.semi, self.interval, self.on_device = semi, int(interval), bool(on_device)
        self.analyzer = SolutionAnalyzer(semi.solver.basis)
        self.history = []

    def __call__(self, u, t):
        if self.on_device:
            l2, linf = self.semi.calc_error_norms(u, t, self.analyzer)
        else:
            l2, linf = calc_error_norms(u, t, self.analyzer, self.semi)
        self.history.append((t, l2, linf))
        return l2, linf


class CallbackSet:
    def __init__(self, *callbacks):
        self.callbacks = callbacks

    def find(self, cls):
        for c in self.callbacks:
            if isinstance(c, cls):
                return c
        return None


class Solution:
    def __init__(self, u, t, nsteps):
        self.u, self.t, self.nsteps = [u], [t], nsteps


def _apply_dimensionwise(V, a, ndim):
    """a: [ncomp, n, (n, (n,)) E] (Fortran-like index order) -> interpolate every node axis with V [m, n]."""
    for d in range(ndim):
        a = np.moveaxis(np.tensordot(V, a, axes=([1], [1 + d])), 0, 1 + d)
    return a


def calc_error_norms(u_ode, t, analyzer, semi):
    """L2 / Linf errors against the initial condition at time t on the 2p analysis nodes (single rank)."""
    if semi.nranks != 1:
        raise NotImplementedError("calc_error_norms is implemented for a single rank")
    nd, n, nv, E = semi.mesh.ndim, semi.nnodes, semi.nvars, semi.nelements
    u = u_ode.detach().cpu().numpy() if hasattr(u_ode, "detach") else np.asarray(u_ode)
    u = u.reshape((nv,) + (n,) * nd + (E,), order="F")
    x = semi.cache_cpu.elements.node_coordinates
    ua = _apply_dimensionwise(analyzer.vandermonde, u, nd)
    xa = _apply_dimensionwise(analyzer.vandermonde, x, nd)
    exact = semi.initial_condition(xa, t, semi.equations)
    diff = exact - ua
    w = analyzer.weights
    wt = w
    for _ in range(nd - 1):
        wt = np.multiply.outer(wt, w)
    vol_jac = (1.0 / semi.cache_cpu.elements.inverse_jacobian) ** nd
    weights = wt[..., None] * vol_jac
    l2 = np.sqrt((diff ** 2 * weights[None]).reshape(nv, -1).sum(axis=1) / semi.mesh.length_level_0 ** nd)
    linf = np.abs(diff).reshape(nv, -1).max(axis=1)
    return l2, linf


def solve(ode, alg=None, dt=1.0, callback=None, maxiters=10 ** 9, fused_stages=False, **_ignored):
    """Low-storage RK loop: `tmp = A_s tmp + dt f(u, t + c_s dt); u += B_s tmp`.

    fused_stages=True runs every stage as ONE library call (`trixib200_rk2n_stage`: rhs! and the 2N update in the
    same kernel on the line-owner path), ping-ponging between two state vectors."""
    semi = ode.p
    u = ode.u0.clone()
    du = semi.new_vector()
    tmp = semi.new_vector().zero_()
    u_alt = semi.new_vector() if fused_stages else None
    t, t_end = float(ode.tspan[0]), float(ode.tspan[1])
    cb = callback if isinstance(callback, CallbackSet) else CallbackSet(*( [callback] if callback else [] ))
    step_cb, ana_cb = cb.find(StepsizeCallback), cb.find(AnalysisCallback)
    from .solution_file import SaveSolutionCallback
    save_cb = cb.find(SaveSolutionCallback)
    nsteps = 0
    if ana_cb is not None:
        ana_cb(u, t)
    if save_cb is not None:
        save_cb(u, t, 0.0, 0, semi)
    while t < t_end and nsteps < maxiters:
        if step_cb is not None:
            dt = calculate_dt(u, t, step_cb.cfl, semi)
        if t + dt > t_end or abs(t + dt - t_end) < 100 * 2.2e-16 * max(1.0, abs(t_end)):
            dt = t_end - t
        if fused_stages:
            res = semi.rk2n_step_ck54(u, u_alt, tmp, t, dt)
            if res is u_alt:
                u, u_alt = u_alt, u
        else:
            for s in range(5):
                ode.f(du, u, semi, t + _C[s] * dt)
                semi.rk2n_update(u, tmp, du, _A[s], _B[s], dt)
        t += dt
        nsteps += 1
        if ana_cb is not None and ana_cb.interval > 0 and nsteps % ana_cb.interval == 0:
            ana_cb(u, t)
        if save_cb is not None:
            save_cb(u, t, dt, nsteps, semi)
    if ana_cb is not None:
        ana_cb(u, t)
    if save_cb is not None:
        save_cb(u, t, dt, nsteps, semi, finished=True)
    sol = Solution(u, t, nsteps)
    return sol
