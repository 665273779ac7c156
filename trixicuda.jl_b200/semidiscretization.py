"""`SemidiscretizationHyperbolicGPU`, `semidiscretizeGPU`, `rhs_gpu_` (Julia: `rhs_gpu!`) and `max_dt`.

Host mirror of the reference's drop-in API for the rhs! path:
  SemidiscretizationHyperbolicGPU  reference src/semidiscretization/semidiscretization_hyperbolic.jl:5-87
  semidiscretizeGPU                reference src/solvers/solvers.jl:43-56
  rhs_gpu!(du_ode, u_ode, semi, t) reference src/solvers/solvers.jl:18-31
  max_dt                           reference src/callbacks_step/stepsize_dg_3d.jl:1-45
  wrap_array                       reference src/solvers/dg.jl:14-21
Device vectors are torch CUDA tensors (PyTorch is only the owner of device memory and streams); all compute
goes through libtrixib200's C ABI. No CPU fallback exists: without the library or a CUDA device this raises.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .equations import (InitialCondition, BoundaryConditionDirichlet, boundary_condition_periodic,
                        boundary_condition_slip_wall)
from .solver import solver_enums
from .treemesh import TreeMesh, init_containers


def _torch():
    import torch
    return torch


class CacheB200:
    """What `mesh_equations_solver_cache` hands to the callbacks (Julia: `CacheB200`, julia/TrixiB200.jl): the library
    handle plus Trixi's CPU containers; attribute access falls through to the CPU containers (`cache.elements`, ...).
    Replaces the reference's NamedTuple of CuArray containers (reference src/solvers/cache.jl:130-212)."""

    def __init__(self, semi):
        self._semi = semi
        self.cpu = semi.cache_cpu

    @property
    def handle(self):
        return self._semi._h

    def __getattr__(self, name):
        return getattr(self.cpu, name)


class SemidiscretizationHyperbolicGPU:
    def __init__(self, mesh: TreeMesh, equations, initial_condition, solver, source_terms=None,
                 boundary_conditions=boundary_condition_periodic, staged_only=False, no_warp_kernel=False, no_line_kernel=False,
                 device=None,
                 rank=0, nranks=1, comm_id=None, node_coordinates="auto"):
        if mesh.ndim != equations.ndim:
            raise ValueError("mesh and equations have different dimensions")
        self.mesh, self.equations, self.initial_condition, self.solver = mesh, equations, initial_condition, solver
        self.source_terms, self.boundary_conditions = source_terms, boundary_conditions
        self.rank, self.nranks = int(rank), int(nranks)
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("SemidiscretizationHyperbolicGPU needs a CUDA device (libtrixib200 has no CPU fallback)")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.cache_cpu = init_containers(mesh, solver.basis.nodes)   # what Trixi's init_* give the Julia shim
        self.cache_gpu = CacheB200(self)
        self._create(staged_only, comm_id, node_coordinates, no_warp_kernel, no_line_kernel)

    # ------------------------------------------------------------------ handle construction
    def _config(self, staged_only):
        eq, solver = self.equations, self.solver
        cfg = _lib.Config()
        cfg.ndim, cfg.polydeg, cfg.equations = eq.ndim, solver.polydeg, eq.kind
        vi, vflux, fvflux, sflux, nc, ind = solver_enums(solver)
        cfg.volume_integral, cfg.volume_flux, cfg.volume_flux_fv, cfg.surface_flux = vi, vflux, fvflux, sflux
        cfg.nonconservative = nc
        if ind is not None:
            cfg.indicator_variable, cfg.alpha_smooth = ind.variable.code, int(ind.alpha_smooth)
            cfg.alpha_max, cfg.alpha_min = ind.alpha_max, ind.alpha_min
        bcs = self.boundary_conditions
        if not isinstance(bcs, (tuple, list, dict)):
            bcs = (bcs,) * (2 * eq.ndim)
        elif isinstance(bcs, dict):
            bcs = tuple(bcs[k] for k in ("x_neg", "x_pos", "y_neg", "y_pos", "z_neg", "z_pos")[: 2 * eq.ndim])
        ic_for_bc = None
        for i, bc in enumerate(bcs):
            if bc is boundary_condition_periodic:
                cfg.boundary_conditions[i] = _lib.BC_PERIODIC
            elif isinstance(bc, BoundaryConditionDirichlet):
                cfg.boundary_conditions[i] = _lib.BC_DIRICHLET_IC
                ic_for_bc = bc.boundary_value_function
            elif bc is boundary_condition_slip_wall:
                cfg.boundary_conditions[i] = _lib.BC_SLIP_WALL
            else:
                raise NotImplementedError(f"boundary condition {bc!r} is not enumerated in libtrixib200")
        ic = self.initial_condition
        if ic_for_bc is not None and isinstance(ic, InitialCondition) and ic_for_bc is not ic:
            raise NotImplementedError("BoundaryConditionDirichlet must use the semidiscretization's initial condition")
        ic_enum = ic_for_bc if ic_for_bc is not None else ic
        cfg.initial_condition = ic_enum.code if isinstance(ic_enum, InitialCondition) else -1   # TRIXIB200_IC_NONE
        if self.source_terms is None:
            cfg.source_terms = _lib.SRC["none"]
        elif getattr(self.source_terms, "code", None) is not None:
            cfg.source_terms = self.source_terms.code
        else:
            raise NotImplementedError(f"source terms {self.source_terms!r} are not enumerated in libtrixib200")
        cfg.device, cfg.rank, cfg.nranks = self.device_index, self.rank, self.nranks
        cfg.flags = _lib.FLAG_STAGED_ONLY if staged_only else 0
        cfg.gamma, cfg.c_h = eq.gamma, eq.c_h
        for d in range(3):
            cfg.advection_velocity[d] = eq.advection_velocity[d]
        return cfg

    def _host_structs(self, node_coordinates):
        solver, c = self.solver, self.cache_cpu
        b, m = solver.basis, solver.mortar
        keep = []   # keep numpy buffers alive while C reads them

        def f64(a, colmajor=False):
            a = _lib.colmajor(a) if colmajor else np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return _lib.fptr(a)

        def i64(a, fortran=False):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel(order="F" if fortran else "C"))
            keep.append(a)
            return _lib.fptr(a)

        bh = _lib.BasisHost()
        bh.nnodes = b.nnodes
        bh.nodes, bh.weights, bh.inverse_weights = f64(b.nodes), f64(b.weights), f64(b.inverse_weights)
        bh.derivative_dhat, bh.derivative_split = f64(b.derivative_dhat, True), f64(b.derivative_split, True)
        bh.boundary_interpolation = f64(b.boundary_interpolation, True)
        bh.inverse_vandermonde_legendre = f64(b.inverse_vandermonde_legendre, True)
        bh.forward_upper, bh.forward_lower = f64(m.forward_upper, True), f64(m.forward_lower, True)
        bh.reverse_upper, bh.reverse_lower = f64(m.reverse_upper, True), f64(m.reverse_lower, True)

        mh = _lib.MeshHost()
        mh.nelements = c.elements.inverse_jacobian.shape[0]
        mh.ninterfaces = c.interfaces.orientations.shape[0]
        mh.nboundaries = c.boundaries.neighbor_ids.shape[0]
        mh.nmortars = c.mortars.orientations.shape[0]
        mh.inverse_jacobian = f64(c.elements.inverse_jacobian)
        want_nc = node_coordinates is True or (node_coordinates == "auto" and mh.nelements <= (1 << 16))
        mh.node_coordinates = f64(c.elements.node_coordinates.ravel(order="F")) if want_nc else None
        mh.cell_centers = f64(self.mesh.cell_centers()[:, : self.mesh.ndim].ravel())
        mh.interfaces_neighbor_ids = i64(c.interfaces.neighbor_ids, True)
        mh.interfaces_orientations = i64(c.interfaces.orientations)
        mh.boundaries_neighbor_ids = i64(c.boundaries.neighbor_ids)
        mh.boundaries_orientations = i64(c.boundaries.orientations)
        mh.boundaries_neighbor_sides = i64(c.boundaries.neighbor_sides)
        mh.boundaries_node_coordinates = f64(c.boundaries.node_coordinates.ravel(order="F"))
        mh.n_boundaries_per_direction = i64(c.boundaries.n_boundaries_per_direction)
        mh.mortars_neighbor_ids = i64(c.mortars.neighbor_ids, True)
        mh.mortars_large_sides = i64(c.mortars.large_sides)
        mh.mortars_orientations = i64(c.mortars.orientations)
        return bh, mh, keep

    def _create(self, staged_only, comm_id, node_coordinates, no_warp_kernel=False, no_line_kernel=False):
        L = _lib.lib()
        cfg = self._config(staged_only)
        if no_warp_kernel:
            cfg.flags |= _lib.FLAG_NO_WARP_KERNEL
        if no_line_kernel:
            cfg.flags |= _lib.FLAG_NO_LINE_KERNEL
        bh, mh, keep = self._host_structs(node_coordinates)
        h = C.c_void_p()
        _lib.check(L.trixib200_create(C.byref(cfg), C.byref(bh), C.byref(mh), C.byref(h)))
        self._L, self._h, self._cfg = L, h, cfg
        del keep
        if self.nranks > 1:
            if comm_id is None:
                raise ValueError("nranks > 1 needs the 128-byte NCCL unique id (see trixib200.distributed)")
            _lib.check(L.trixib200_comm_init(h, comm_id))
        self.nvars, self.nnodes = self.size("nvars"), self.size("nnodes")
        self.nelements, self.first_element = self.size("nelements"), self.size("first_element")
        self.nelements_global = self.size("nelements_global")
        self.fused = bool(self.size("fused"))
        self.warp3d = bool(self.size("warp3d"))
        self.line3d = bool(self.size("line3d"))
        self._stream = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.trixib200_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ sizes / arrays
    def size(self, name):
        return int(self._L.trixib200_size(self._h, name.encode()))

    def ndofs(self):
        return self.size("ndofs")

    def ndofsglobal(self):
        return self.nelements_global * self.nnodes ** self.mesh.ndim

    def nunknowns(self):
        return self.size("nunknowns")

    def _sync_stream(self):
        """Adopt torch's current stream so library kernels are ordered with torch work on the tensors."""
        s = _torch().cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            _lib.check(self._L.trixib200_set_stream(self._h, s))
            self._stream = s

    def new_vector(self):
        return _torch().empty(self.nunknowns(), dtype=_torch().float64, device=self.device)

    def local_slice(self, u_global):
        per = self.nvars * self.nnodes ** self.mesh.ndim
        return u_global[per * self.first_element: per * (self.first_element + self.nelements)]

    # ------------------------------------------------------------------ initial condition
    def compute_coefficients(self, t=0.0, func=None):
        """Host evaluation on the global mesh (Trixi `compute_coefficients`), flat Trixi-layout vector."""
        func = self.initial_condition if func is None else func
        x = self.cache_cpu.elements.node_coordinates
        if self.mesh.ndim == 1:
            x = x.copy()
            x[:, 0, :] = np.nextafter(x[:, 0, :], np.inf)
            x[:, -1, :] = np.nextafter(x[:, -1, :], -np.inf)
        u = np.asarray(func(x, t, self.equations), dtype=np.float64)
        return np.ascontiguousarray(u.ravel(order="F"))

    def compute_coefficients_gpu(self, t=0.0, on_device=False):
        torch = _torch()
        if on_device:
            if not isinstance(self.initial_condition, InitialCondition):
                raise NotImplementedError("device-side initial conditions must be enumerated")
            u = self.new_vector()
            self._sync_stream()
            _lib.check(self._L.trixib200_fill_initial_condition(self._h, u.data_ptr(), float(t)))
            return u
        u_host = self.local_slice(self.compute_coefficients(t))
        return torch.from_numpy(np.ascontiguousarray(u_host)).to(self.device)

    # ------------------------------------------------------------------ hot path
    def rhs(self, du, u, t):
        self._sync_stream()
        _lib.check(self._L.trixib200_rhs(self._h, du.data_ptr(), u.data_ptr(), float(t)))

    def rhs_host(self, du_host, u_host, t):
        """`rhs!(du_ode, u_ode, semi, t)` on HOST vectors (numpy float64 or CPU torch tensors, ideally pinned):
        upload, rhs!, download inside the library (trixib200_rhs_host)."""
        self._sync_stream()
        pu = u_host.data_ptr() if hasattr(u_host, "data_ptr") else u_host.ctypes.data
        pdu = du_host.data_ptr() if hasattr(du_host, "data_ptr") else du_host.ctypes.data
        _lib.check(self._L.trixib200_rhs_host(self._h, pdu, pu, float(t)))

    def max_dt(self, u, t=0.0):
        self._sync_stream()
        out = C.c_double()
        _lib.check(self._L.trixib200_max_dt(self._h, u.data_ptr(), float(t), C.byref(out)))
        return out.value

    def stage(self, name, du, u, t=0.0):
        self._sync_stream()
        _lib.check(self._L.trixib200_stage(self._h, name.encode(), du.data_ptr(), u.data_ptr(), float(t)))

    def cache(self, name):
        n = self._L.trixib200_cache_len(self._h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        _lib.check(self._L.trixib200_cache_get(self._h, name.encode(), _lib.fptr(out), n))
        return out

    def rk2n_update(self, u, tmp, du, a, b, dt):
        self._sync_stream()
        _lib.check(self._L.trixib200_rk2n_update(self._h, u.data_ptr(), tmp.data_ptr(), du.data_ptr(), a, b, dt))

    def rk2n_stage(self, u_out, u_in, tmp, t, a, b, dt):
        """rhs! fused with one 2N Runge-Kutta stage: tmp = a tmp + dt rhs(u_in, t); u_out = u_in + b tmp
        (trixib200_rk2n_stage; one launch on the line-owner kernel path, du never materialised)."""
        self._sync_stream()
        _lib.check(self._L.trixib200_rk2n_stage(self._h, u_out.data_ptr(), u_in.data_ptr(), tmp.data_ptr(), float(t),
                                                float(a), float(b), float(dt)))

    def rk2n_step_ck54(self, u, u_alt, tmp, t, dt):
        """One CarpenterKennedy2N54 step; returns the vector that holds the result (u_alt after five stages)."""
        self._sync_stream()
        flag = C.c_int(0)
        _lib.check(self._L.trixib200_rk2n_step_ck54(self._h, u.data_ptr(), u_alt.data_ptr(), tmp.data_ptr(), float(t),
                                                    float(dt), C.byref(flag)))
        return u_alt if flag.value else u

    def calc_error_norms(self, u, t, analyzer):
        """`calc_error_norms(cons2cons, u, t, analyzer, ...)` on the device (trixib200_calc_error_norms): L2 / Linf
        errors against the enumerated initial condition at time t on the analyzer's nodes, reduced over all ranks."""
        self._sync_stream()
        V = np.ascontiguousarray(analyzer.vandermonde, dtype=np.float64)       # [n_analysis, nnodes] row-major
        w = np.ascontiguousarray(analyzer.weights, dtype=np.float64)
        l2, linf = np.empty(self.nvars), np.empty(self.nvars)
        vol = float(self.mesh.length_level_0) ** self.mesh.ndim
        _lib.check(self._L.trixib200_calc_error_norms(self._h, u.data_ptr(), float(t), int(V.shape[0]), _lib.fptr(V),
                                                      _lib.fptr(w), vol, _lib.fptr(l2), _lib.fptr(linf)))
        return l2, linf

    def integrate(self, u, normalize=True):
        """`integrate(cons2cons, u, ...)` on the device: domain integrals of the conserved variables."""
        self._sync_stream()
        out = np.empty(self.nvars)
        vol = float(self.mesh.length_level_0) ** self.mesh.ndim
        _lib.check(self._L.trixib200_integrate(self._h, u.data_ptr(), int(bool(normalize)), vol, _lib.fptr(out)))
        return out

    def launch_count(self):
        return int(self._L.trixib200_launch_count(self._h))

    def time_rhs(self, du, u, t, reps):
        self._sync_stream()
        ms = C.c_float()
        _lib.check(self._L.trixib200_time_rhs(self._h, du.data_ptr(), u.data_ptr(), float(t), int(reps), C.byref(ms)))
        return ms.value


@dataclass
class ODEProblem:
    """`ODEProblem{true, FullSpecialize}(rhs_gpu!, u0_ode, tspan, semi)` (reference src/solvers/solvers.jl:53-55)."""
    f: object
    u0: object
    tspan: tuple
    p: SemidiscretizationHyperbolicGPU


def rhs_gpu_(du_ode, u_ode, semi, t):
    """Julia `rhs_gpu!(du_ode, u_ode, semi, t)`: in place, returns None."""
    semi.rhs(du_ode, u_ode, t)
    return None


def semidiscretizeGPU(semi, tspan, on_device_ic=False):
    u0 = semi.compute_coefficients_gpu(tspan[0], on_device=on_device_ic)
    return ODEProblem(rhs_gpu_, u0, tuple(tspan), semi)


def mesh_equations_solver_cache(semi):
    """reference src/semidiscretization/semidiscretization_hyperbolic.jl:91-95: the GPU cache goes to the callbacks."""
    return semi.mesh, semi.equations, semi.solver, semi.cache_gpu


def wrap_array(u_ode, mesh_or_semi, equations=None, dg=None, cache=None):
    """`wrap_array(u_ode, mesh, equations, dg, cache)` = `reshape(u_ode, nvars, N.., nelements)` (reference
    src/solvers/dg.jl:14-21); torch is row-major so the view is [element, (k, j,) i, v]. `wrap_array(u_ode, semi)` is
    accepted as a shorthand."""
    semi = mesh_or_semi if cache is None else cache._semi
    n, nd = semi.nnodes, semi.mesh.ndim
    return u_ode.view((semi.nelements,) + (n,) * nd + (semi.nvars,))


def _need_cache(cache, what):
    if not isinstance(cache, CacheB200):
        raise TypeError(f"{what}: the cache must be the CacheB200 returned by mesh_equations_solver_cache(semi) "
                        f"(got {type(cache).__name__}); there is no CPU method behind this name")
    return cache._semi


def max_dt(u, t, mesh, constant_speed, equations, dg, cache):
    """The reference's method signature (src/callbacks_step/stepsize_dg_3d.jl:1-45), exactly as Trixi's
    StepsizeCallback calls it: `max_dt(u, t, mesh, have_constant_speed(equations), equations, solver, cache)` with
    (mesh, equations, solver, cache) = mesh_equations_solver_cache(semi) and u = wrap_array(u_ode, ...)."""
    semi = _need_cache(cache, "max_dt")
    if mesh is not semi.mesh or equations is not semi.equations or dg is not semi.solver:
        raise ValueError("max_dt: mesh / equations / solver do not belong to this cache")
    if bool(constant_speed) != bool(equations.have_constant_speed()):
        raise ValueError("max_dt: constant_speed must be have_constant_speed(equations)")
    return semi.max_dt(u, t)


def calc_error_norms(func, u_ode, t, analyzer, semi, cache_analysis=None):
    """reference src/semidiscretization/semidiscretization_hyperbolic.jl:97-105 -> analysis_dg_3d.jl:45-89; `func` must
    be `cons2cons` (the device kernel reduces the conserved variables against the enumerated initial condition)."""
    if func is not cons2cons:
        raise NotImplementedError("calc_error_norms on the device is enumerated for cons2cons")
    mesh, equations, solver, cache = mesh_equations_solver_cache(semi)
    u = wrap_array(u_ode, mesh, equations, solver, cache)
    return _need_cache(cache, "calc_error_norms").calc_error_norms(u, t, analyzer)


def integrate(func, u, mesh, equations, dg, cache, normalize=True):
    """reference src/callbacks_step/analysis_dg_3d.jl:35-43 for `func = cons2cons`."""
    if func is not cons2cons:
        raise NotImplementedError("integrate on the device is enumerated for cons2cons")
    return _need_cache(cache, "integrate").integrate(u, normalize=normalize)


def cons2cons(u, equations):
    return u
