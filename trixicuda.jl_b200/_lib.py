"""ctypes binding of libtrixib200.so (the C ABI in include/trixib200.h) and its in-tree build.

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRIXIB200_LIB") or os.path.join(_HERE, "libtrixib200.so")   # override: A/B builds
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["runtime.cu", "kernels_analysis.cuh", "kernels_staged.cuh", "kernels_fused.cuh", "kernels_warp3d.cuh", "kernels_line3d.cuh", "kernels_line6.cuh", "kernels_line6_phase.inc", "equations.cuh",
           "device.cuh"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC"]

# enums (include/trixib200.h)
EQ_ADVECTION, EQ_EULER, EQ_MHD = 0, 1, 2
FLUX = {"flux_central": 0, "flux_lax_friedrichs": 1, "flux_lax_friedrichs_naive": 2, "flux_hll": 3,
        "flux_hll_naive": 4, "flux_ranocha": 5, "flux_shima_etal": 6, "flux_hindenlang_gassner": 7,
        "flux_hlle": 8}
VI_WEAK_FORM, VI_FLUX_DIFFERENCING, VI_SHOCK_CAPTURING_HG = 0, 1, 2
IND = {"density": 0, "pressure": 1, "density_pressure": 2}
BC_PERIODIC, BC_DIRICHLET_IC, BC_SLIP_WALL = 0, 1, 2
IC = {"constant": 0, "convergence_test": 1, "weak_blast_wave": 2, "density_wave": 3}
SRC = {"none": 0, "convergence_test": 1}
FLAG_STAGED_ONLY = 1
FLAG_NO_WARP_KERNEL = 2
FLAG_NO_LINE_KERNEL = 4


class Config(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("polydeg", C.c_int32), ("equations", C.c_int32), ("volume_integral", C.c_int32),
        ("volume_flux", C.c_int32), ("volume_flux_fv", C.c_int32), ("surface_flux", C.c_int32),
        ("nonconservative", C.c_int32), ("indicator_variable", C.c_int32), ("alpha_smooth", C.c_int32),
        ("boundary_conditions", C.c_int32 * 6), ("initial_condition", C.c_int32), ("source_terms", C.c_int32),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("flags", C.c_int32),
        ("alpha_max", C.c_double), ("alpha_min", C.c_double), ("gamma", C.c_double),
        ("advection_velocity", C.c_double * 3), ("c_h", C.c_double),
    ]


class BasisHost(C.Structure):
    _fields_ = [("nnodes", C.c_int32)] + [(n, C.c_void_p) for n in (
        "nodes", "weights", "inverse_weights", "derivative_dhat", "derivative_split", "boundary_interpolation",
        "inverse_vandermonde_legendre", "forward_upper", "forward_lower", "reverse_upper", "reverse_lower")]


class MeshHost(C.Structure):
    _fields_ = [("nelements", C.c_int64), ("ninterfaces", C.c_int64), ("nboundaries", C.c_int64),
                ("nmortars", C.c_int64)] + [(n, C.c_void_p) for n in (
        "inverse_jacobian", "node_coordinates", "cell_centers", "interfaces_neighbor_ids",
        "interfaces_orientations", "boundaries_neighbor_ids", "boundaries_orientations",
        "boundaries_neighbor_sides", "boundaries_node_coordinates", "n_boundaries_per_direction",
        "mortars_neighbor_ids", "mortars_large_sides", "mortars_orientations")]


EXPORTS = [
    "trixib200_last_error", "trixib200_version", "trixib200_create", "trixib200_destroy", "trixib200_size",
    "trixib200_rhs", "trixib200_max_dt", "trixib200_stage", "trixib200_cache_len", "trixib200_cache_get",
    "trixib200_alloc", "trixib200_free", "trixib200_upload", "trixib200_download", "trixib200_sync",
    "trixib200_stream", "trixib200_fill_initial_condition", "trixib200_rk2n_update", "trixib200_rk2n_stage", "trixib200_rk2n_step_ck54", "trixib200_calc_error_norms", "trixib200_integrate",
    "trixib200_time_rhs",
    "trixib200_launch_count", "trixib200_comm_unique_id", "trixib200_comm_init", "trixib200_set_stream",
    "trixib200_rhs_host", "trixib200_host_register", "trixib200_host_unregister",
    "trixib200_plan_create", "trixib200_plan_destroy", "trixib200_plan_len", "trixib200_plan_get",
]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(_HERE, "..", "include", "trixib200.h")
    return any(os.path.getmtime(p) > t for p in [os.path.join(CSRC, s) for s in SOURCES] + [hdr])


def build(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU; the .so stays in-tree so it travels to the GPU box."""
    if not (force or needs_build()):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH,
                                                                          os.path.join(CSRC, "runtime.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load libtrixib200.so; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(libtrixib200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.trixib200_last_error.restype = C.c_char_p
    L.trixib200_version.restype = C.c_int
    L.trixib200_create.restype = C.c_int
    L.trixib200_create.argtypes = [C.POINTER(Config), C.POINTER(BasisHost), C.POINTER(MeshHost),
                                   C.POINTER(C.c_void_p)]
    L.trixib200_destroy.argtypes = [C.c_void_p]
    L.trixib200_size.restype = C.c_int64
    L.trixib200_size.argtypes = [C.c_void_p, C.c_char_p]
    L.trixib200_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.trixib200_rhs_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.trixib200_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.trixib200_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
    L.trixib200_max_dt.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    L.trixib200_stage.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_double]
    L.trixib200_cache_len.restype = C.c_int64
    L.trixib200_cache_len.argtypes = [C.c_void_p, C.c_char_p]
    L.trixib200_cache_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.trixib200_alloc.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    L.trixib200_free.argtypes = [C.c_void_p, C.c_void_p]
    L.trixib200_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.trixib200_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.trixib200_sync.argtypes = [C.c_void_p]
    L.trixib200_stream.restype = C.c_int64
    L.trixib200_stream.argtypes = [C.c_void_p]
    L.trixib200_fill_initial_condition.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.trixib200_rk2n_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                        C.c_double]
    L.trixib200_rk2n_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                       C.c_double, C.c_double]
    L.trixib200_rk2n_step_ck54.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                           C.POINTER(C.c_int)]
    L.trixib200_calc_error_norms.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_void_p, C.c_void_p]
    L.trixib200_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    L.trixib200_time_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                     C.POINTER(C.c_float)]
    L.trixib200_launch_count.restype = C.c_int64
    L.trixib200_launch_count.argtypes = [C.c_void_p]
    L.trixib200_set_stream.argtypes = [C.c_void_p, C.c_int64]
    L.trixib200_plan_create.argtypes = [C.POINTER(Config), C.POINTER(MeshHost), C.POINTER(C.c_void_p)]
    L.trixib200_plan_destroy.argtypes = [C.c_void_p]
    L.trixib200_plan_len.restype = C.c_int64
    L.trixib200_plan_len.argtypes = [C.c_void_p, C.c_char_p]
    L.trixib200_plan_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.trixib200_comm_unique_id.argtypes = [C.c_char_p]
    L.trixib200_comm_init.argtypes = [C.c_void_p, C.c_char_p]
    _lib = L
    return L


class TrixiB200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise TrixiB200Error(f"libtrixib200 error {rc}: {lib().trixib200_last_error().decode()}")


def fptr(a):
    """Pointer to a numpy array's data (column-major data must be passed as a Fortran-contiguous array)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def colmajor(m):
    """(row, col) numpy matrix -> flat column-major float64 buffer (Julia layout)."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float64).ravel(order="F"))
