"""`DGSEMGPU` and the volume-integral / indicator types it is configured with.

Mirrors reference src/solvers/dgsem_gpu.jl:6-53 (`DGSEMGPU(; RealT, polydeg, surface_flux, surface_integral,
volume_integral)`), and the Trixi types the reference accepts there: `VolumeIntegralWeakForm`,
`VolumeIntegralFluxDifferencing`, `VolumeIntegralShockCapturingHG` with `IndicatorHennemannGassner`
(reference src/solvers/indicators.jl:7-42, examples/euler_shockcapturing_3d.jl:19-30).
"""
from dataclasses import dataclass

import numpy as np

from . import _lib
from .basis import LobattoLegendreBasisGPU, MortarL2GPU
from .equations import flux_central, density_pressure, split_flux


class SurfaceIntegralWeakForm:
    def __init__(self, surface_flux=flux_central):
        self.surface_flux = surface_flux


class VolumeIntegralWeakForm:
    kind = _lib.VI_WEAK_FORM


class VolumeIntegralFluxDifferencing:
    kind = _lib.VI_FLUX_DIFFERENCING

    def __init__(self, volume_flux):
        self.volume_flux = volume_flux


class IndicatorHennemannGassner:
    """`IndicatorHennemannGassner(equations, basis; alpha_max, alpha_min, alpha_smooth, variable)`.
    The reference evaluates it on the CPU after copying u to the host (dg_3d.jl:187-188); libtrixib200
    evaluates it on the device."""

    def __init__(self, equations, basis, alpha_max=0.5, alpha_min=0.001, alpha_smooth=True,
                 variable=density_pressure):
        self.alpha_max, self.alpha_min = float(alpha_max), float(alpha_min)
        self.alpha_smooth, self.variable = bool(alpha_smooth), variable


class VolumeIntegralShockCapturingHG:
    kind = _lib.VI_SHOCK_CAPTURING_HG

    def __init__(self, indicator, volume_flux_dg, volume_flux_fv):
        self.indicator, self.volume_flux_dg, self.volume_flux_fv = indicator, volume_flux_dg, volume_flux_fv


@dataclass
class DG:
    """The plain `DG` struct the reference's `DGSEMGPU` returns (basis, mortar, surface/volume integral)."""
    basis: LobattoLegendreBasisGPU
    mortar: MortarL2GPU
    surface_integral: SurfaceIntegralWeakForm
    volume_integral: object

    @property
    def polydeg(self):
        return self.basis.polydeg


def DGSEMGPU(polydeg, surface_flux=flux_central, surface_integral=None, volume_integral=None, RealT=np.float64,
             basis=None):
    basis = basis if basis is not None else LobattoLegendreBasisGPU(polydeg, RealT)
    if surface_integral is None:
        surface_integral = SurfaceIntegralWeakForm(surface_flux)
    if volume_integral is None:
        volume_integral = VolumeIntegralWeakForm()
    return DG(basis, MortarL2GPU(basis), surface_integral, volume_integral)


def solver_enums(solver):
    """(volume_integral, volume_flux, volume_flux_fv, surface_flux, nonconservative, indicator) as C enums."""
    sflux, s_nc = split_flux(solver.surface_integral.surface_flux)
    vi = solver.volume_integral
    vflux = fvflux = sflux
    v_nc = f_nc = s_nc
    ind = None
    if vi.kind == _lib.VI_FLUX_DIFFERENCING:
        vflux, v_nc = split_flux(vi.volume_flux)
        fvflux, f_nc = vflux, v_nc
    elif vi.kind == _lib.VI_SHOCK_CAPTURING_HG:
        vflux, v_nc = split_flux(vi.volume_flux_dg)
        fvflux, f_nc = split_flux(vi.volume_flux_fv)
        ind = vi.indicator
    if vi.kind != _lib.VI_WEAK_FORM and not (v_nc == s_nc == f_nc):
        raise NotImplementedError("volume and surface fluxes must agree on the nonconservative term")
    for f in (sflux, vflux, fvflux):
        if getattr(f, "code", None) is None:
            raise NotImplementedError(f"{f!r} is not an enumerated libtrixib200 flux")
    return vi.kind, vflux.code, fvflux.code, sflux.code, int(s_nc), ind
