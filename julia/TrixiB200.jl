# TrixiB200.jl -- the Julia side of the drop-in: TrixiCUDA.jl's exported API (DGSEMGPU, SemidiscretizationHyperbolicGPU,
# semidiscretizeGPU; reference src/TrixiCUDA.jl:74-77) kept as is, with `rhs_gpu!` and `max_dt` going through `ccall`
# into libtrixib200.so (include/trixib200.h) instead of launching CUDA.jl-generated kernels.
#
# STATUS: written against the reference sources, NOT executed -- neither Julia nor Trixi.jl / CUDA.jl exist in the
# build image or on the GPU boxes (DESIGN.md "Host language"). The identical call sequence is exercised through the
# Python mirror (trixicuda.jl_b200/semidiscretization.py) by tests/test_gpu_parity.py. Every function cites the
# reference definition it replaces.
module TrixiB200

using Trixi
using Trixi: AbstractSemidiscretization, DG, TreeMesh, PerformanceCounter, nvariables, nnodes, nelements, ndofs,
             create_cache, digest_boundary_conditions, check_periodicity_mesh_boundary_conditions,
             compute_coefficients, local_leaf_cells, init_elements, init_interfaces, init_boundaries, init_mortars,
             polydeg, have_nonconservative_terms, BoundaryConditionPeriodic, BoundaryConditionDirichlet,
             VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing, VolumeIntegralShockCapturingHG,
             IndicatorHennemannGassner, LobattoLegendreBasis, LobattoLegendreMortarL2
using SciMLBase: ODEProblem, FullSpecialize
using CUDA: CuArray, CuVector, device, stream    # storage only: no CUDA.jl kernel runs on the rhs! path

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
initial_condition_id(f) = f === initial_condition_constant ? Int32(0) :
                          f === initial_condition_convergence_test ? Int32(1) :
                          f === initial_condition_weak_blast_wave ? Int32(2) :
                          f === initial_condition_density_wave ? Int32(3) : Int32(-1)
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
mutable struct SemidiscretizationHyperbolicGPU{Mesh, Equations, InitialCondition, BoundaryConditions, SourceTerms,
                                               Solver, CacheCPU} <: AbstractSemidiscretization
    mesh::Mesh
    equations::Equations
    initial_condition::InitialCondition
    boundary_conditions::BoundaryConditions
    source_terms::SourceTerms
    solver::Solver
    cache_gpu::Ptr{Cvoid}          # trixib200_handle*: replaces the NamedTuple of CuArray containers
    cache_cpu::CacheCPU
    performance_counter::PerformanceCounter
end

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
        bc isa BoundaryConditionDirichlet ? Int32(1) :
        bc === Trixi.boundary_condition_slip_wall ? Int32(2) :
        error("libtrixib200: boundary condition not enumerated")
    end
    adv = equations isa Trixi.AbstractLinearScalarAdvectionEquation ?
          ntuple(i -> i <= nd ? Float64(equations.advection_velocity[i]) : 0.0, 3) : (0.0, 0.0, 0.0)
    cfg = Config(nd, polydeg(solver), equations_id(equations), vi, vflux, fvflux, sflux,
                 Int32(have_nonconservative_terms(equations) == Trixi.True()),
                 indicator === nothing ? Int32(2) : indicator_variable_id(indicator.variable),
                 indicator === nothing ? Int32(0) : Int32(indicator.alpha_smooth),
                 bc_ids, max(initial_condition_id(initial_condition), Int32(0)), source_terms_id(source_terms),
                 Int32(CUDA.deviceid(device())), Int32(rank), Int32(nranks), Int32(0),
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
                                           handle[], cache_cpu, PerformanceCounter())
    finalizer(s -> ccall((:trixib200_destroy, LIB), Cint, (Ptr{Cvoid},), s.cache_gpu), semi)
    return semi
end

@inline Base.ndims(semi::SemidiscretizationHyperbolicGPU) = ndims(semi.mesh)
@inline Trixi.mesh_equations_solver_cache(semi::SemidiscretizationHyperbolicGPU) =
    (semi.mesh, semi.equations, semi.solver, semi.cache_cpu)

# ------------------------------------------------------------------------------------------------ rhs!
# reference src/solvers/solvers.jl:18-31 -> src/solvers/dg_3d.jl:895-925 (ten CUDA.jl kernel stages): one ccall.
# `wrap_array` is a reshape, so the flat vectors go through unchanged (reference src/solvers/dg.jl:14-21).
function rhs_gpu!(du_ode::CuVector{Float64}, u_ode::CuVector{Float64}, semi::SemidiscretizationHyperbolicGPU, t)
    # library work is ordered on CUDA.jl's task-local stream, like the broadcasts OrdinaryDiffEq issues around it
    check(ccall((:trixib200_set_stream, LIB), Cint, (Ptr{Cvoid}, Int64), semi.cache_gpu,
                reinterpret(Int64, stream().handle)))
    check(ccall((:trixib200_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), semi.cache_gpu,
                reinterpret(Ptr{Float64}, pointer(du_ode)), reinterpret(Ptr{Float64}, pointer(u_ode)), t))
    return nothing
end
# host vectors (Trixi's CPU signature): upload, rhs!, download inside the library
function rhs_gpu!(du_ode::Vector{Float64}, u_ode::Vector{Float64}, semi::SemidiscretizationHyperbolicGPU, t)
    check(ccall((:trixib200_rhs_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), semi.cache_gpu,
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
# reference src/callbacks_step/stepsize_dg_3d.jl:20-45 (full device -> host copy of u + serial loop): device
# reduction, one double comes back
function Trixi.max_dt(u::CuArray{Float64}, t, mesh::TreeMesh, constant_speed, equations, dg::DG,
                      semi::SemidiscretizationHyperbolicGPU)
    out = Ref{Float64}(0.0)
    check(ccall((:trixib200_max_dt, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ref{Float64}), semi.cache_gpu,
                reinterpret(Ptr{Float64}, pointer(u)), t, out))
    return out[]
end

# ------------------------------------------------------------------------------------------------ fused RK stages
# Optional fast path (SURVEY.md section 8(f) row 1): OrdinaryDiffEq's `perform_step!` for `CarpenterKennedy2N54` is
# `rhs_gpu!` followed by two broadcasts per stage; `trixib200_rk2n_stage` does rhs! and the 2N update in one kernel
# and never writes du. `step_ck2n54!` advances (u, t) by dt with five fused stages, ping-ponging between u and u_alt,
# and returns the vector that holds the new state. A `DiscreteCallback`-free driver can call it in a plain loop with
# Trixi's StepsizeCallback logic (`dt = cfl * max_dt(u, ...)`).
function rk2n_stage!(u_out::CuVector{Float64}, u_in::CuVector{Float64}, tmp::CuVector{Float64},
                     semi::SemidiscretizationHyperbolicGPU, t, a, b, dt)
    check(ccall((:trixib200_set_stream, LIB), Cint, (Ptr{Cvoid}, Int64), semi.cache_gpu,
                reinterpret(Int64, stream().handle)))
    check(ccall((:trixib200_rk2n_stage, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Float64),
                semi.cache_gpu, reinterpret(Ptr{Float64}, pointer(u_out)), reinterpret(Ptr{Float64}, pointer(u_in)),
                reinterpret(Ptr{Float64}, pointer(tmp)), t, a, b, dt))
    return nothing
end
function step_ck2n54!(u::CuVector{Float64}, u_alt::CuVector{Float64}, tmp::CuVector{Float64},
                      semi::SemidiscretizationHyperbolicGPU, t, dt)
    in_alt = Ref{Cint}(0)
    check(ccall((:trixib200_set_stream, LIB), Cint, (Ptr{Cvoid}, Int64), semi.cache_gpu,
                reinterpret(Int64, stream().handle)))
    check(ccall((:trixib200_rk2n_step_ck54, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ref{Cint}), semi.cache_gpu,
                reinterpret(Ptr{Float64}, pointer(u)), reinterpret(Ptr{Float64}, pointer(u_alt)),
                reinterpret(Ptr{Float64}, pointer(tmp)), t, dt, in_alt))
    return in_alt[] == 1 ? u_alt : u
end

end # module
