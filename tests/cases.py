"""Shared problem definitions: each case builds the SAME semidiscretization on the CPU oracle and on the
product (trixib200), mirroring how every reference test builds a CPU `DGSEM` and a GPU `DGSEMGPU` solver
on the same TreeMesh (reference test/tree_dgsem_3d/euler_ec.jl:14-26)."""
import numpy as np

import oracle as O

BOX3 = (dict(type="box", coordinates_min=(-0.5, -0.5, -0.5), coordinates_max=(0.5, 0.5, 0.5)),)
BOX2 = (dict(type="box", coordinates_min=(0.0, -1.0), coordinates_max=(1.0, 1.0)),)
# a patch whose mortars are cut by the Morton-range partition at 2, 4 and 8 ranks (7 / 11 / 15 of its 24 mortars)
BOX3_OFF = (dict(type="box", coordinates_min=(-0.5, -0.5, -1.0), coordinates_max=(0.5, 0.5, 0.0)),)


def case(ndim, equations, level=2, polydeg=3, vi="weak_form", volume_flux="flux_central",
         volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs", noncons=False,
         ic="convergence_test", source="none", bc="periodic", periodic=True, patches=(), gamma=1.4,
         adv=(0.2, -0.7, 0.9), c_h=1.0, cmin=-1.0, cmax=1.0, alpha_max=0.5, alpha_min=0.001, alpha_smooth=True,
         variable="density_pressure"):
    return dict(ndim=ndim, equations=equations, level=level, polydeg=polydeg, vi=vi, volume_flux=volume_flux,
                volume_flux_fv=volume_flux_fv, surface_flux=surface_flux, noncons=noncons, ic=ic, source=source,
                bc=bc, periodic=periodic, patches=tuple(patches), gamma=gamma, adv=tuple(adv[:ndim]), c_h=c_h,
                cmin=(cmin,) * ndim, cmax=(cmax,) * ndim, alpha_max=alpha_max, alpha_min=alpha_min,
                alpha_smooth=alpha_smooth, variable=variable)


CASES = {
    # BASELINE.json configs (reduced levels where the full size is only a bench workload)
    "c1_advection_1d": case(1, "advection", level=4, adv=(1.0,)),
    "c2_euler_ec_2d": case(2, "euler", level=4, vi="flux_differencing", volume_flux="flux_ranocha",
                           surface_flux="flux_ranocha", ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "c3_euler_sc_3d": case(3, "euler", level=3, vi="shock_capturing_hg", volume_flux="flux_ranocha",
                           volume_flux_fv="flux_ranocha", surface_flux="flux_ranocha", ic="weak_blast_wave",
                           cmin=-2.0, cmax=2.0),
    "c4_mhd_alfven_mortar_3d": case(3, "mhd", level=2, vi="flux_differencing", volume_flux="flux_hindenlang_gassner",
                                    surface_flux="flux_hlle", noncons=True, gamma=5 / 3, patches=BOX3, c_h=1.3),
    "c5_euler_ec_3d": case(3, "euler", level=3, vi="flux_differencing", volume_flux="flux_ranocha",
                           surface_flux="flux_ranocha", ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    # the rest of the reference's 3D/2D/1D test matrix (test/tree_dgsem_*d/*.jl) restricted to enumerated physics
    "advection_basic_2d": case(2, "advection", level=3),
    "advection_basic_3d": case(3, "advection", level=2),
    "advection_mortar_2d": case(2, "advection", level=2, patches=BOX2),
    "advection_mortar_3d": case(3, "advection", level=2, patches=BOX3),
    "euler_ec_1d": case(1, "euler", level=4, vi="flux_differencing", volume_flux="flux_ranocha",
                        surface_flux="flux_ranocha", ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "euler_shima_3d": case(3, "euler", level=2, vi="flux_differencing", volume_flux="flux_shima_etal",
                           surface_flux="flux_lax_friedrichs", ic="density_wave"),
    "euler_source_terms_3d": case(3, "euler", level=2, source="convergence_test", cmin=0.0, cmax=2.0),
    "euler_source_terms_hll_2d": case(2, "euler", level=3, source="convergence_test", surface_flux="flux_hll",
                                      cmin=0.0, cmax=2.0),
    "euler_nonperiodic_2d": case(2, "euler", level=3, source="convergence_test", bc="dirichlet_ic", periodic=False,
                                 cmin=0.0, cmax=2.0),
    "euler_nonperiodic_3d": case(3, "euler", level=2, source="convergence_test", bc="dirichlet_ic", periodic=False,
                                 cmin=0.0, cmax=2.0, surface_flux="flux_lax_friedrichs_naive"),
    "euler_mortar_3d": case(3, "euler", level=2, source="convergence_test", patches=BOX3, cmin=0.0, cmax=2.0),
    "euler_ec_mortar_2d": case(2, "euler", level=3, vi="flux_differencing", volume_flux="flux_ranocha",
                               surface_flux="flux_ranocha", ic="weak_blast_wave", patches=BOX2),
    "euler_shock_1d": case(1, "euler", level=5, vi="shock_capturing_hg", volume_flux="flux_shima_etal",
                           volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs",
                           ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "euler_shock_2d": case(2, "euler", level=4, vi="shock_capturing_hg", volume_flux="flux_shima_etal",
                           volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs",
                           ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "euler_shock_mortar_3d": case(3, "euler", level=2, vi="shock_capturing_hg", volume_flux="flux_ranocha",
                                  volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs",
                                  ic="weak_blast_wave", patches=BOX3),
    "mhd_ec_3d": case(3, "mhd", level=2, vi="flux_differencing", volume_flux="flux_hindenlang_gassner",
                      surface_flux="flux_hindenlang_gassner", noncons=True, ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "mhd_alfven_wave_3d": case(3, "mhd", level=2, vi="flux_differencing", volume_flux="flux_hindenlang_gassner",
                               surface_flux="flux_lax_friedrichs", noncons=True, gamma=5 / 3),
    "mhd_shock_3d": case(3, "mhd", level=2, vi="shock_capturing_hg", volume_flux="flux_hindenlang_gassner",
                         volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs", noncons=True,
                         ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    # 3D Euler flux differencing with faces whose flux is given (mortars, Dirichlet boundaries): the line-owner
    # kernel consumes surface_flux_values on those faces
    "euler_ec_mortar_3d": case(3, "euler", level=2, vi="flux_differencing", volume_flux="flux_ranocha",
                               surface_flux="flux_ranocha", ic="weak_blast_wave", patches=BOX3),
    "euler_fd_nonperiodic_3d": case(3, "euler", level=2, vi="flux_differencing", volume_flux="flux_ranocha",
                                    surface_flux="flux_lax_friedrichs", source="convergence_test", bc="dirichlet_ic",
                                    periodic=False, cmin=0.0, cmax=2.0),
    # closed boxes with boundary_condition_slip_wall: line-owner kernel with given wall
    # fluxes, and the 2D weak form
    "euler_slip_wall_3d": case(3, "euler", level=2, vi="flux_differencing", volume_flux="flux_ranocha",
                               surface_flux="flux_lax_friedrichs", ic="weak_blast_wave", bc="slip_wall", periodic=False,
                               cmin=-2.0, cmax=2.0),
    "euler_slip_wall_2d": case(2, "euler", level=3, surface_flux="flux_hll", ic="weak_blast_wave", bc="slip_wall",
                               periodic=False, cmin=-2.0, cmax=2.0),
    # mortars that cross partition cuts (multi-GPU: replicated on every rank that owns one of their elements)
    "euler_ec_mortar_off_3d": case(3, "euler", level=2, vi="flux_differencing", volume_flux="flux_ranocha",
                                   surface_flux="flux_ranocha", ic="weak_blast_wave", patches=BOX3_OFF),
    "mhd_alfven_mortar_off_3d": case(3, "mhd", level=2, vi="flux_differencing", volume_flux="flux_hindenlang_gassner",
                                     surface_flux="flux_hlle", noncons=True, gamma=5 / 3, patches=BOX3_OFF, c_h=1.3),
    "euler_shock_mortar_off_3d": case(3, "euler", level=2, vi="shock_capturing_hg", volume_flux="flux_ranocha",
                                      volume_flux_fv="flux_lax_friedrichs", surface_flux="flux_lax_friedrichs",
                                      ic="weak_blast_wave", patches=BOX3_OFF),
    # other polynomial degrees go through the staged kernels
    "euler_ec_3d_p2": case(3, "euler", level=2, polydeg=2, vi="flux_differencing", volume_flux="flux_ranocha",
                           surface_flux="flux_ranocha", ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "euler_ec_2d_p5": case(2, "euler", level=3, polydeg=5, vi="flux_differencing", volume_flux="flux_ranocha",
                           surface_flux="flux_ranocha", ic="weak_blast_wave", cmin=-2.0, cmax=2.0),
    "advection_mortar_3d_p4": case(3, "advection", level=2, polydeg=4, patches=BOX3),
}


def make_oracle(c, level=None):
    nd = c["ndim"]
    return O.Oracle(
        ndim=nd, equations=c["equations"], polydeg=c["polydeg"], volume_integral=c["vi"],
        volume_flux=c["volume_flux"], volume_flux_fv=c["volume_flux_fv"], surface_flux=c["surface_flux"],
        nonconservative=c["noncons"], alpha_max=c["alpha_max"], alpha_min=c["alpha_min"],
        alpha_smooth=c["alpha_smooth"], indicator_variable=c["variable"], initial_condition=c["ic"],
        source=c["source"], bc=(c["bc"],) * 6, gamma=c["gamma"], advection_velocity=c["adv"] + (0.0,) * (3 - nd),
        c_h=c["c_h"], coordinates_min=c["cmin"], coordinates_max=c["cmax"],
        initial_refinement_level=c["level"] if level is None else level, periodicity=(c["periodic"],) * 3,
        refinement_patches=[(p["coordinates_min"], p["coordinates_max"]) for p in c["patches"]])


def make_semi(c, level=None, staged_only=False, **kw):
    """Build the product-side semidiscretization through the reference-shaped API."""
    import trixib200 as T
    nd = c["ndim"]
    if c["equations"] == "advection":
        eq = (T.LinearScalarAdvectionEquation1D, T.LinearScalarAdvectionEquation2D,
              T.LinearScalarAdvectionEquation3D)[nd - 1](c["adv"])
    elif c["equations"] == "euler":
        eq = (T.CompressibleEulerEquations1D, T.CompressibleEulerEquations2D,
              T.CompressibleEulerEquations3D)[nd - 1](c["gamma"])
    else:
        eq = T.IdealGlmMhdEquations3D(c["gamma"], initial_c_h=c["c_h"])
    fluxes = {"flux_central": T.flux_central, "flux_lax_friedrichs": T.flux_lax_friedrichs,
              "flux_lax_friedrichs_naive": T.FluxLaxFriedrichs(T.max_abs_speed_naive), "flux_hll": T.flux_hll,
              "flux_hll_naive": T.FluxHLL(T.min_max_speed_naive), "flux_ranocha": T.flux_ranocha,
              "flux_shima_etal": T.flux_shima_etal, "flux_hindenlang_gassner": T.flux_hindenlang_gassner,
              "flux_hlle": T.flux_hlle}

    def fl(name):
        return (fluxes[name], T.flux_nonconservative_powell) if c["noncons"] else fluxes[name]

    basis = T.LobattoLegendreBasisGPU(c["polydeg"])
    if c["vi"] == "weak_form":
        vi = T.VolumeIntegralWeakForm()
    elif c["vi"] == "flux_differencing":
        vi = T.VolumeIntegralFluxDifferencing(fl(c["volume_flux"]))
    else:
        ind = T.IndicatorHennemannGassner(eq, basis, alpha_max=c["alpha_max"], alpha_min=c["alpha_min"],
                                          alpha_smooth=c["alpha_smooth"],
                                          variable={"density": T.density, "pressure": T.pressure,
                                                    "density_pressure": T.density_pressure}[c["variable"]])
        vi = T.VolumeIntegralShockCapturingHG(ind, volume_flux_dg=fl(c["volume_flux"]),
                                              volume_flux_fv=fl(c["volume_flux_fv"]))
    solver = T.DGSEMGPU(polydeg=c["polydeg"], surface_flux=fl(c["surface_flux"]), volume_integral=vi, basis=basis)
    mesh = T.TreeMesh(c["cmin"], c["cmax"], initial_refinement_level=c["level"] if level is None else level,
                      refinement_patches=c["patches"], periodicity=c["periodic"], n_cells_max=10 ** 8)
    ic = {"constant": T.initial_condition_constant, "convergence_test": T.initial_condition_convergence_test,
          "weak_blast_wave": T.initial_condition_weak_blast_wave,
          "density_wave": T.initial_condition_density_wave}[c["ic"]]
    src = T.source_terms_convergence_test if c["source"] == "convergence_test" else None
    bc = {"periodic": T.boundary_condition_periodic, "slip_wall": T.boundary_condition_slip_wall}.get(c["bc"])
    if bc is None:
        bc = T.BoundaryConditionDirichlet(ic)
    return T.SemidiscretizationHyperbolicGPU(mesh, eq, ic, solver, source_terms=src, boundary_conditions=bc,
                                             staged_only=staged_only, **kw)


def rel_max_err(a, b):
    """max|a - b| / max|b|  -- the parity metric of SURVEY.md section 8(d)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def nan_rule_equal(gpu, cpu, tol=0.0):
    """Reference `@test_approx` NaN rule (test/test_macros.jl:41-71): NaN (CPU, unused slot) vs 0 (GPU) passes."""
    gpu, cpu = np.asarray(gpu), np.asarray(cpu)
    nan = np.isnan(cpu)
    if np.any(gpu[nan] != 0.0):
        return False
    den = np.abs(cpu[~nan]).max() if (~nan).any() else 1.0
    return bool(np.all(np.abs(gpu[~nan] - cpu[~nan]) <= tol * max(den, 1e-300)))
