# TrixiB200.jl -- the Julia side of the drop-in: TrixiCUDA.jl's exported API (DGSEMGPU, SemidiscretizationHyperbolicGPU,
# semidiscretizeGPU; reference src/TrixiCUDA.jl:74-77) kept as is, with `rhs_gpu!` and `max_dt` going through `ccall`
# into libtrixib200.so (include/trixib200.h) instead of launching CUDA.jl-generated kernels.
#
# STATUS: written against the reference sources, NOT executed -- neither Julia nor Trixi.jl / CUDA.jl exist in the
# build image or on the GPU boxes (DESIGN.md "Host language"). The identical call sequence -- including the argument
# tuples Trixi's StepsizeCallback / AnalysisCallback pass (`mesh_equations_solver_cache` -> `wrap_array` -> `max_dt` /
# `calc_error_norms` / `integrate` dispatching on the cache type) -- is exercised through the Python mirror
# (trixicuda.jl_b200/semidiscretization.py) by tests/test_gpu_parity.py::test_callback_argument_tuples. Every function
# cites the reference definition it replaces.
module TrixiB200

using Trixi
using Trixi: AbstractSemidiscretization, DG, TreeMesh, PerformanceCounter, nvariables, nnodes, nelements, ndofs,
             create_cache, digest_boundary_conditions, check_periodicity_mesh_boundary_conditions,
             compute_coefficients, local_leaf_cells, init_elements, init_interfaces, init_boundaries, init_mortars,
             polydeg, have_nonconservative_terms, BoundaryConditionPeriodic, BoundaryConditionDirichlet,
             VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing, VolumeIntegralShockCapturingHG,
             IndicatorHennemannGassner, LobattoLegendreBasis, LobattoLegendreMortarL2
using Trixi: True, False, AbstractMesh, cons2cons, total_volume, analyze
import Trixi: wrap_array, wrap_array_native, max_dt, calc_error_norms, integrate, integrate_via_indices,
              analyze_integrals, mesh_equations_solver_cache      # extended exactly like reference src/TrixiCUDA.jl:56-62
using SciMLBase: ODEProblem, FullSpecialize
using StaticArrays: SVector
import CUDA                                     # `CUDA.deviceid`, `CUDA.device`, `CUDA.stream`
using CUDA: CuArray, CuVector, AbstractGPUArray   # storage only: no CUDA.jl kernel runs on the rhs! path

export DGSEMGPU, SemidiscretizationHyperbolicGPU, semidiscretizeGPU

const LIB = get(ENV, "TRIXIB200_LIB", "libtrixib200.so")

# ------------------------------------------------------------------------------------------------ C structs
# mirror include/trixib200.h field for field
struct Config
    ndim::Int32; polydeg::Int32; equations::Int32; volume_integral::Int32
    volume_flux::Int32; volume_flux_fv::Int32; surface_flux::Int32; nonconservative::Int32
    indicator_variable::Int32; alpha_smooth::Int32
    boundary_conditions::NTuple{6, Int32}
    initial_condition::Int32; source_terms::Int32
    device::Int32; rank::Int32; nranks::Int32; flags::Int32
    alpha_max::Float64; alpha_min::Float64; gamma::Float64
    advection_velocity::NTuple{3, Float64}
    c_h::Float64
end
struct BasisHost
    nnodes::Int32
    nodes::Ptr{Float64}; weights::Ptr{Float64}; inverse_weights::Ptr{Float64}
    derivative_dhat::Ptr{Float64}; derivative_split::Ptr{Float64}; boundary_interpolation::Ptr{Float64}
    inverse_vandermonde_legendre::Ptr{Float64}
    forward_upper::Ptr{Float64}; forward_lower::Ptr{Float64}; reverse_upper::Ptr{Float64}; reverse_lower::Ptr{Float64}
end
struct MeshHost
    nelements::Int64; ninterfaces::Int64; nboundaries::Int64; nmortars::Int64
    inverse_jacobian::Ptr{Float64}; node_coordinates::Ptr{Float64}; cell_centers::Ptr{Float64}
    interfaces_neighbor_ids::Ptr{Int64}; interfaces_orientations::Ptr{Int64}
    boundaries_neighbor_ids::Ptr{Int64}; boundaries_orientations::Ptr{Int64}; boundaries_neighbor_sides::Ptr{Int64}
    boundaries_node_coordinates::Ptr{Float64}; n_boundaries_per_direction::Ptr{Int64}
    mortars_neighbor_ids::Ptr{Int64}; mortars_large_sides::Ptr{Int64}; mortars_orientations::Ptr{Int64}
end

check(rc) = rc == 0 ? nothing :
            error("libtrixib200 error $rc: " * unsafe_string(ccall((:trixib200_last_error, LIB), Cstring, ())))

# ------------------------------------------------------------------------------------------------ enumerations
# A C ABI cannot take Julia closures: Trixi singletons map to the enums of include/trixib200.h, anything else is
# an error (never a fallback).
equations_id(::LinearScalarAdvectionEquation1D) = Int32(0)
equations_id(::LinearScalarAdvectionEquation2D) = Int32(0)
equations_id(::LinearScalarAdvectionEquation3D) = Int32(0)
equations_id(::CompressibleEulerEquations1D) = Int32(1)
equations_id(::CompressibleEulerEquations2D) = Int32(1)
equations_id(::CompressibleEulerEquations3D) = Int32(1)
equations_id(::IdealGlmMhdEquations3D) = Int32(2)
equations_id(eq) = error("libtrixib200: equations $(typeof(eq)) are not enumerated")

flux_id(::typeof(flux_central)) = Int32(0)
flux_id(f::FluxLaxFriedrichs) = f.dissipation.max_abs_speed === max_abs_speed_naive ? Int32(2) : Int32(1)
flux_id(f::FluxHLL) = f.min_max_speed === min_max_speed_naive ? Int32(4) : Int32(3)   # flux_hlle -> 8, see below
flux_id(::typeof(flux_ranocha)) = Int32(5)
flux_id(::typeof(flux_shima_etal)) = Int32(6)
flux_id(::typeof(flux_hindenlang_gassner)) = Int32(7)
flux_id(f::Tuple) = (f[2] === flux_nonconservative_powell ||
                     error("libtrixib200: only flux_nonconservative_powell is enumerated"); flux_id(f[1]))
flux_id(f) = f === flux_hlle ? Int32(8) : error("libtrixib200: flux $(f) is not enumerated")

volume_integral_ids(::VolumeIntegralWeakForm) = (Int32(0), Int32(0), Int32(0), nothing)
volume_integral_ids(vi::VolumeIntegralFluxDifferencing) = (Int32(1), flux_id(vi.volume_flux), Int32(0), nothing)
volume_integral_ids(vi::VolumeIntegralShockCapturingHG) =
    (Int32(2), flux_id(vi.volume_flux_dg), flux_id(vi.volume_flux_fv), vi.indicator)

indicator_variable_id(f) = f === density ? Int32(0) : f === pressure ? Int32(1) :
                           f === density_pressure ? Int32(2) : error("libtrixib200: indicator variable not enumerated")
# TRIXIB200_IC_NONE (-1) for anything else: the library then refuses Dirichlet(ic) boundaries, the device IC fill
# and the device error norms instead of substituting another state; rhs! itself never needs the IC.
initial_condition_id(f) = f === initial_condition_constant ? Int32(0) :
                          f === initial_condition_convergence_test ? Int32(1) :
                          f === initial_condition_weak_blast_wave ? Int32(2) :
                          f === initial_condition_density_wave ? Int32(3) : Int32(-1)
# a Dirichlet boundary evaluates the enumerated IC on the device, so its boundary_value_function must BE the
# semidiscretization's initial condition
function check_dirichlet(bc::BoundaryConditionDirichlet, initial_condition)
    bc.boundary_value_function === initial_condition ||
        error("libtrixib200: BoundaryConditionDirichlet must use the semidiscretization's initial condition")
    initial_condition_id(initial_condition) >= 0 ||
        error("libtrixib200: BoundaryConditionDirichlet needs an enumerated initial condition")
    return Int32(1)
end
source_terms_id(::Nothing) = Int32(0)
source_terms_id(f) = f === source_terms_convergence_test ? Int32(1) :
                     error("libtrixib200: source terms $(f) are not enumerated")

# ------------------------------------------------------------------------------------------------ solver
# reference src/solvers/dgsem_gpu.jl:43-53. The reference needs GPU copies of the basis inside the solver
# (LobattoLegendreBasisGPU); here the basis stays Trixi's CPU basis and its operators are handed to C once.
function DGSEMGPU(; RealT = Float64, polydeg::Integer, surface_flux = flux_central,
                  surface_integral = SurfaceIntegralWeakForm(surface_flux),
                  volume_integral = VolumeIntegralWeakForm())
    RealT === Float64 || error("libtrixib200 computes in Float64")
    return DGSEM(; RealT, polydeg, surface_flux, surface_integral, volume_integral)
end

# ------------------------------------------------------------------------------------------------ semidiscretization
# reference src/semidiscretization/semidiscretization_hyperbolic.jl:5-87
# What `mesh_equations_solver_cache` hands to Trixi's callbacks in place of the reference's NamedTuple of CuArray
# containers (reference src/solvers/cache.jl:130-212): the library handle plus Trixi's own CPU containers. Property
# access falls through to the CPU cache, so Trixi code that only inspects containers (`nelements(dg, cache)`,
# `cache.elements.inverse_jacobian`, `create_cache_analysis`) keeps working on host arrays, while the methods that touch
# `u` dispatch on `cache::CacheB200` below and go to the device.
struct CacheB200{CacheCPU}
    handle::Ptr{Cvoid}             # trixib200_handle*
    cpu::CacheCPU
end
Base.getproperty(c::CacheB200, s::Symbol) =
    (s === :handle || s === :cpu) ? getfield(c, s) : getproperty(getfield(c, :cpu), s)
Base.propertynames(c::CacheB200) = (:handle, :cpu, propertynames(getfield(c, :cpu))...)

mutable struct SemidiscretizationHyperbolicGPU{Mesh, Equations, InitialCondition, BoundaryConditions, SourceTerms,
                                               Solver, CacheCPU} <: AbstractSemidiscretization
    mesh::Mesh
    equations::Equations
    initial_condition::InitialCondition
    boundary_conditions::BoundaryConditions
    source_terms::SourceTerms
    solver::Solver
    cache_gpu::CacheB200{CacheCPU}  # same field names as the reference (semidiscretization_hyperbolic.jl:5-22)
    cache_cpu::CacheCPU
    performance_counter::PerformanceCounter
end
handle(semi::SemidiscretizationHyperbolicGPU) = semi.cache_gpu.handle

function SemidiscretizationHyperbolicGPU(mesh::TreeMesh, equations, initial_condition, solver;
                                         source_terms = nothing,
                                         boundary_conditions = boundary_condition_periodic,
                                         RealT = real(solver), uEltype = RealT,
                                         rank = 0, nranks = 1)
    @assert ndims(mesh) == ndims(equations)
    # Trixi's CPU containers, exactly as the reference obtains them (src/solvers/cache.jl:130-158)
    cache_cpu = create_cache(mesh, equations, solver, RealT, uEltype)
    _bcs = digest_boundary_conditions(boundary_conditions, mesh, solver, cache_cpu)
    check_periodicity_mesh_boundary_conditions(mesh, _bcs)
    (; elements, interfaces, boundaries, mortars) = cache_cpu
    basis, mortar = solver.basis, solver.mortar
    nd = ndims(mesh)

    vi, vflux, fvflux, indicator = volume_integral_ids(solver.volume_integral)
    sflux = flux_id(solver.surface_integral.surface_flux)
    bc_ids = ntuple(6) do i
        i > 2nd && return Int32(0)
        bc = _bcs isa BoundaryConditionPeriodic ? _bcs : _bcs[i]
        bc isa BoundaryConditionPeriodic ? Int32(0) :
        bc isa BoundaryConditionDirichlet ? check_dirichlet(bc, initial_condition) :
        bc === Trixi.boundary_condition_slip_wall ? Int32(2) :
        error("libtrixib200: boundary condition not enumerated")
    end
    adv = equations isa Trixi.AbstractLinearScalarAdvectionEquation ?
          ntuple(i -> i <= nd ? Float64(equations.advection_velocity[i]) : 0.0, 3) : (0.0, 0.0, 0.0)
    cfg = Config(nd, polydeg(solver), equations_id(equations), vi, vflux, fvflux, sflux,
                 Int32(have_nonconservative_terms(equations) == Trixi.True()),
                 indicator === nothing ? Int32(2) : indicator_variable_id(indicator.variable),
                 indicator === nothing ? Int32(0) : Int32(indicator.alpha_smooth),
                 bc_ids, initial_condition_id(initial_condition), source_terms_id(source_terms),
                 Int32(CUDA.deviceid(CUDA.device())), Int32(rank), Int32(nranks), Int32(0),
                 indicator === nothing ? 0.0 : indicator.alpha_max, indicator === nothing ? 0.0 : indicator.alpha_min,
                 hasproperty(equations, :gamma) ? equations.gamma : 0.0, adv,
                 hasproperty(equations, :c_h) ? equations.c_h : 0.0)

    # column-major Julia arrays are passed as they are (the header documents Trixi's layouts)
    f64(a) = pointer(a isa Array{Float64} ? a : (a = Array{Float64}(a)))
    keep = Any[]
    p(a) = (b = Array(a); push!(keep, b); pointer(b))
    bh = BasisHost(nnodes(basis), p(basis.nodes), p(basis.weights), p(basis.inverse_weights),
                   p(basis.derivative_dhat), p(basis.derivative_split), p(basis.boundary_interpolation),
                   p(basis.inverse_vandermonde_legendre), p(mortar.forward_upper), p(mortar.forward_lower),
                   p(mortar.reverse_upper), p(mortar.reverse_lower))
    i64(a) = (b = Array{Int64}(a); push!(keep, b); pointer(b))
    mh = MeshHost(nelements(elements), Trixi.ninterfaces(interfaces), Trixi.nboundaries(boundaries),
                  Trixi.nmortars(mortars),
                  p(elements.inverse_jacobian), p(elements.node_coordinates), Ptr{Float64}(0),
                  i64(interfaces.neighbor_ids), i64(interfaces.orientations),
                  i64(boundaries.neighbor_ids), i64(boundaries.orientations), i64(boundaries.neighbor_sides),
                  p(boundaries.node_coordinates), i64(boundaries.n_boundaries_per_direction),
                  i64(mortars.neighbor_ids), i64(mortars.large_sides), i64(mortars.orientations))
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:trixib200_create, LIB), Cint,
                                  (Ref{Config}, Ref{BasisHost}, Ref{MeshHost}, Ref{Ptr{Cvoid}}), cfg, bh, mh, handle))
    semi = SemidiscretizationHyperbolicGPU(mesh, equations, initial_condition, _bcs, source_terms, solver,
                                           CacheB200(handle[], cache_cpu), cache_cpu, PerformanceCounter())
    finalizer(s -> ccall((:trixib200_destroy, LIB), Cint, (Ptr{Cvoid},), s.cache_gpu.handle), semi)
    return semi
end

@inline Base.ndims(semi::SemidiscretizationHyperbolicGPU) = ndims(semi.mesh)
# reference src/semidiscretization/semidiscretization_hyperbolic.jl:91-95: the GPU cache goes to the callbacks
@inline function mesh_equations_solver_cache(semi::SemidiscretizationHyperbolicGPU)
    (; mesh, equations, solver, cache_gpu) = semi
    return mesh, equations, solver, cache_gpu
end

# reference src/solvers/dg.jl:14-21,24-33: Trixi's `unsafe_wrap(Array, pointer(u_ode), ...)` cannot take a CuVector
@inline function wrap_array(u_ode::AbstractGPUArray, mesh::AbstractMesh, equations, dg::DG, cache)
    reshape(u_ode, nvariables(equations), ntuple(_ -> nnodes(dg), ndims(mesh))..., nelements(dg, cache))
end
@inline wrap_array_native(u_ode::AbstractGPUArray, mesh::AbstractMesh, equations, dg::DG, cache) =
    wrap_array(u_ode, mesh, equations, dg, cache)

devptr(a::AbstractGPUArray{Float64}) = reinterpret(Ptr{Float64}, pointer(a))
set_stream!(h::Ptr{Cvoid}) = check(ccall((:trixib200_set_stream, LIB), Cint, (Ptr{Cvoid}, Int64), h,
                                         reinterpret(Int64, CUDA.stream().handle)))

# ------------------------------------------------------------------------------------------------ rhs!
# reference src/solvers/solvers.jl:18-31 -> src/solvers/dg_3d.jl:895-925 (ten CUDA.jl kernel stages): one ccall.
# `wrap_array` is a reshape, so the flat vectors go through unchanged (reference src/solvers/dg.jl:14-21).
function rhs_gpu!(du_ode::CuVector{Float64}, u_ode::CuVector{Float64}, semi::SemidiscretizationHyperbolicGPU, t)
    # library work is ordered on CUDA.jl's task-local stream, like the broadcasts OrdinaryDiffEq issues around it
    set_stream!(handle(semi))
    check(ccall((:trixib200_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), handle(semi),
                devptr(du_ode), devptr(u_ode), t))
    return nothing
end
# Trixi's AnalysisCallback evaluates `rhs!(du_ode, u_ode, semi, t)` itself (entropy time derivative): same entry point
Trixi.rhs!(du_ode::CuVector{Float64}, u_ode::CuVector{Float64}, semi::SemidiscretizationHyperbolicGPU, t) =
    rhs_gpu!(du_ode, u_ode, semi, t)
# host vectors (Trixi's CPU signature): upload, rhs!, download inside the library
function rhs_gpu!(du_ode::Vector{Float64}, u_ode::Vector{Float64}, semi::SemidiscretizationHyperbolicGPU, t)
    check(ccall((:trixib200_rhs_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), handle(semi),
                du_ode, u_ode, t))
    return nothing
end

# reference src/solvers/solvers.jl:43-56; the initial state is computed by Trixi on the host and uploaded, the
# alternative the reference itself notes at solvers.jl:49-51
function semidiscretizeGPU(semi::SemidiscretizationHyperbolicGPU, tspan)
    u0_ode = CuArray(compute_coefficients(first(tspan), semi))
    return ODEProblem{true, FullSpecialize}(rhs_gpu!, u0_ode, tspan, semi)
end

# ------------------------------------------------------------------------------------------------ StepsizeCallback
# Trixi's `StepsizeCallback` calls (callbacks_step/stepsize.jl, `calculate_dt`):
#     mesh, equations, solver, cache = mesh_equations_solver_cache(semi)
#     u  = wrap_array(u_ode, mesh, equations, solver, cache)
#     dt = cfl * max_dt(u, t, mesh, have_constant_speed(equations), equations, solver, cache)
# The methods below have the reference's signatures (src/callbacks_step/stepsize_dg_{1,2,3}d.jl: `u::CuArray`, `True` /
# `False`) with `cache::CacheB200`; the reference copies u to the host and loops serially, here one double comes back
# from a device reduction (all ranks: ncclAllReduce(max) inside the library).
for ND in 1:3, CS in (:True, :False)
    @eval function max_dt(u::CuArray, t, mesh::TreeMesh{$ND}, constant_speed::$CS, equations, dg::DG,
                          cache::CacheB200)
        out = Ref{Float64}(0.0)
        set_stream!(cache.handle)
        check(ccall((:trixib200_max_dt, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ref{Float64}), cache.handle,
                    devptr(u), Float64(t), out))
        return out[]
    end
end

# ------------------------------------------------------------------------------------------------ AnalysisCallback
# reference src/semidiscretization/semidiscretization_hyperbolic.jl:97-105 and src/callbacks_step/analysis_dg_{1,2,3}d.jl
function calc_error_norms(func, u_ode, t, analyzer, semi::SemidiscretizationHyperbolicGPU, cache_analysis)
    (; mesh, equations, initial_condition, solver, cache_gpu) = semi
    u = wrap_array(u_ode, mesh, equations, solver, cache_gpu)
    calc_error_norms(func, u, t, analyzer, mesh, equations, initial_condition, solver, cache_gpu, cache_analysis)
end
# L2 / Linf errors of the conserved variables against an enumerated initial condition: on the device
# (trixib200_calc_error_norms interpolates u and the node coordinates to the analyzer's nodes like Trixi's
# multiply_dimensionwise!, evaluates the IC there and reduces deterministically). Everything else takes the
# reference's route (src/callbacks_step/analysis_dg_3d.jl:52-54: host copy, Trixi's CPU method on the CPU containers).
function calc_error_norms(func, u::CuArray, t, analyzer, mesh::TreeMesh, equations, initial_condition, dg::DG,
                          cache::CacheB200, cache_analysis)
    if func === cons2cons && initial_condition_id(initial_condition) >= 0
        nv = nvariables(equations)
        l2, linf = zeros(nv), zeros(nv)
        V = permutedims(Array{Float64}(analyzer.vandermonde))     # C wants [n_analysis][nnodes] row-major
        w = Array{Float64}(analyzer.weights)
        set_stream!(cache.handle)
        check(ccall((:trixib200_calc_error_norms, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
                     Ptr{Float64}), cache.handle, devptr(u), Float64(t), Int32(length(w)), V, w,
                    Float64(total_volume(mesh)), l2, linf))
        return SVector{nv}(l2), SVector{nv}(linf)
    end
    return calc_error_norms(func, Array(u), t, analyzer, mesh, equations, initial_condition, dg, cache.cpu,
                            cache_analysis)
end
# domain integrals: the conserved variables on the device, any other functional through a host copy as in the
# reference (src/callbacks_step/analysis_dg_3d.jl:1-43)
function integrate(func::Func, u::CuArray, mesh::TreeMesh, equations, dg::DG, cache::CacheB200;
                   normalize = true) where {Func}
    if func === cons2cons
        out = zeros(nvariables(equations))
        set_stream!(cache.handle)
        check(ccall((:trixib200_integrate, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{Float64}),
                    cache.handle, devptr(u), Int32(normalize), Float64(total_volume(mesh)), out))
        return SVector{length(out)}(out)
    end
    return integrate(func, Array(u), mesh, equations, dg, cache.cpu; normalize)
end
integrate_via_indices(func::Func, u::CuArray, mesh::TreeMesh, equations, dg::DG, cache::CacheB200, args...;
                      normalize = true) where {Func} =
    integrate_via_indices(func, Array(u), mesh, equations, dg, cache.cpu, args...; normalize)
# reference src/callbacks_step/analysis.jl:1-36: the analysis integrals are arbitrary Julia functionals of (du, u):
# ONE host copy for all of them (the reference copies per quantity), then Trixi's own `analyze`
function analyze_integrals(analysis_integrals::NTuple{N, Any}, io, du::CuArray, u::CuArray, t,
                           semi::SemidiscretizationHyperbolicGPU) where {N}
    analyze_integrals(analysis_integrals, io, Array(du), Array(u), t, semi)
end
analyze_integrals(::Tuple{}, io, du::CuArray, u::CuArray, t, semi::SemidiscretizationHyperbolicGPU) = nothing

# ------------------------------------------------------------------------------------------------ fused RK stages
# Optional fast path (SURVEY.md section 8(f) row 1): OrdinaryDiffEq's `perform_step!` for `CarpenterKennedy2N54` is
# `rhs_gpu!` followed by two broadcasts per stage; `trixib200_rk2n_stage` does rhs! and the 2N update in one kernel
# and never writes du. `step_ck2n54!` advances (u, t) by dt with five fused stages, ping-ponging between u and u_alt,
# and returns the vector that holds the new state. A `DiscreteCallback`-free driver can call it in a plain loop with
# Trixi's StepsizeCallback logic (`dt = cfl * max_dt(u, ...)`).
function rk2n_stage!(u_out::CuVector{Float64}, u_in::CuVector{Float64}, tmp::CuVector{Float64},
                     semi::SemidiscretizationHyperbolicGPU, t, a, b, dt)
    set_stream!(handle(semi))
    check(ccall((:trixib200_rk2n_stage, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Float64),
                handle(semi), devptr(u_out), devptr(u_in), devptr(tmp), t, a, b, dt))
    return nothing
end
function step_ck2n54!(u::CuVector{Float64}, u_alt::CuVector{Float64}, tmp::CuVector{Float64},
                      semi::SemidiscretizationHyperbolicGPU, t, dt)
    in_alt = Ref{Cint}(0)
    set_stream!(handle(semi))
    check(ccall((:trixib200_rk2n_step_ck54, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ref{Cint}), handle(semi),
                devptr(u), devptr(u_alt), devptr(tmp), t, dt, in_alt))
    return in_alt[] == 1 ? u_alt : u
end

end # module
