"""trixib200: B200-native TreeMesh DGSEM rhs! behind the TrixiCUDA.jl API.

The package directory is `trixicuda.jl_b200/`; import it as `trixib200` (see trixib200.py at the repo root).
Exports mirror the reference's four public names (reference src/TrixiCUDA.jl:74-77) plus the Trixi.jl names a
reference example script uses (reference examples/euler_ec_3d.jl).
"""
from . import _lib
from ._lib import TrixiB200Error
from .basis import LobattoLegendreBasisGPU, MortarL2GPU, SolutionAnalyzer
from .treemesh import TreeMesh, init_containers
from .equations import (
    LinearScalarAdvectionEquation1D, LinearScalarAdvectionEquation2D, LinearScalarAdvectionEquation3D,
    CompressibleEulerEquations1D, CompressibleEulerEquations2D, CompressibleEulerEquations3D,
    IdealGlmMhdEquations3D,
    flux_central, flux_lax_friedrichs, flux_hll, flux_ranocha, flux_shima_etal, flux_hindenlang_gassner,
    flux_hlle, flux_nonconservative_powell, FluxLaxFriedrichs, FluxHLL, max_abs_speed_naive, max_abs_speed,
    min_max_speed_naive, min_max_speed_davis, min_max_speed_einfeldt,
    initial_condition_constant, initial_condition_convergence_test, initial_condition_weak_blast_wave,
    initial_condition_density_wave, InitialCondition,
    boundary_condition_periodic, boundary_condition_slip_wall, BoundaryConditionDirichlet,
    source_terms_convergence_test,
    density, pressure, density_pressure,
)
from .solver import (DGSEMGPU, SurfaceIntegralWeakForm, VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing,
                     VolumeIntegralShockCapturingHG, IndicatorHennemannGassner)
from .semidiscretization import (SemidiscretizationHyperbolicGPU, semidiscretizeGPU, rhs_gpu_, wrap_array, max_dt,
                                 ODEProblem, CacheB200, mesh_equations_solver_cache, cons2cons, integrate)
from . import semidiscretization as _semi_mod
from .ode import (CarpenterKennedy2N54, StepsizeCallback, AnalysisCallback, CallbackSet, solve, calc_error_norms,
                  calculate_dt)
from .solution_file import (SaveSolutionCallback, save_solution_file, save_mesh_file, load_solution_file, cons2prim,
                            varnames)
calc_error_norms_gpu = _semi_mod.calc_error_norms

__all__ = [n for n in dir() if not n.startswith("_")]
