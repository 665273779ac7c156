/* libtrixib200 -- C ABI of the B200-native TreeMesh DGSEM rhs! hot path.
 *
 * Drop-in boundary for TrixiCUDA.jl: the Julia shim (julia/TrixiB200.jl) keeps the
 * reference's exported API (DGSEMGPU, SemidiscretizationHyperbolicGPU, semidiscretizeGPU; reference
 * src/TrixiCUDA.jl:74-77) and `ccall`s these entry points instead of launching CUDA.jl kernels.
 * Every function returns 0 on success and a negative TRIXIB200_E* code on failure; the message is available
 * from trixib200_last_error(). Nothing throws or aborts across this boundary. There is NO CPU fallback:
 * if no CUDA device / kernel image is usable, create() fails.
 *
 * Array conventions are Trixi's: column-major, variable-fastest `u[v, i, j, k, element]`
 * (reference src/solvers/dg.jl:14-21 `wrap_array`), Int64 1-based ids in the containers
 * (reference src/solvers/containers_3d.jl:10,59-60,101-105,165-167).
 */
#ifndef TRIXIB200_H
#define TRIXIB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trixib200_handle trixib200_handle;

enum { TRIXIB200_OK = 0, TRIXIB200_EINVAL = -1, TRIXIB200_EUNSUPPORTED = -2, TRIXIB200_ECUDA = -3,
       TRIXIB200_ENOMEM = -4, TRIXIB200_ECOMM = -5 };

/* equations (reference: passed through as Trixi types) */
enum { TRIXIB200_EQ_ADVECTION = 0,      /* LinearScalarAdvectionEquation{1,2,3}D */
       TRIXIB200_EQ_EULER = 1,          /* CompressibleEulerEquations{1,2,3}D */
       TRIXIB200_EQ_MHD = 2 };          /* IdealGlmMhdEquations3D */

/* two-point / surface fluxes (Trixi callables invoked at reference src/solvers/dg_3d_kernel.jl:226-234,1166) */
enum { TRIXIB200_FLUX_CENTRAL = 0, TRIXIB200_FLUX_LAX_FRIEDRICHS = 1, TRIXIB200_FLUX_LAX_FRIEDRICHS_NAIVE = 2,
       TRIXIB200_FLUX_HLL = 3, TRIXIB200_FLUX_HLL_NAIVE = 4, TRIXIB200_FLUX_RANOCHA = 5,
       TRIXIB200_FLUX_SHIMA_ETAL = 6, TRIXIB200_FLUX_HINDENLANG_GASSNER = 7, TRIXIB200_FLUX_HLLE = 8 };

/* volume integrals (reference src/solvers/dg_3d.jl:18,57,106,175,259) */
enum { TRIXIB200_VI_WEAK_FORM = 0, TRIXIB200_VI_FLUX_DIFFERENCING = 1, TRIXIB200_VI_SHOCK_CAPTURING_HG = 2 };

enum { TRIXIB200_IND_DENSITY = 0, TRIXIB200_IND_PRESSURE = 1, TRIXIB200_IND_DENSITY_PRESSURE = 2 };
enum { TRIXIB200_BC_PERIODIC = 0, TRIXIB200_BC_DIRICHLET_IC = 1,   /* BoundaryConditionDirichlet(initial_condition) */
       TRIXIB200_BC_SLIP_WALL = 2 };   /* boundary_condition_slip_wall, compressible Euler only */
enum { TRIXIB200_IC_NONE = -1,           /* the caller's initial condition is not one of the enumerated ones */
       TRIXIB200_IC_CONSTANT = 0, TRIXIB200_IC_CONVERGENCE_TEST = 1, TRIXIB200_IC_WEAK_BLAST_WAVE = 2,
       TRIXIB200_IC_DENSITY_WAVE = 3 };
enum { TRIXIB200_SRC_NONE = 0, TRIXIB200_SRC_CONVERGENCE_TEST = 1 };

/* flags */
enum { TRIXIB200_FLAG_STAGED_ONLY = 1,     /* force the staged (materialising) kernels; debugging / stage parity */
       TRIXIB200_FLAG_NO_WARP_KERNEL = 2,  /* keep the thread-per-node fused kernel where the warp-per-element one applies */
       TRIXIB200_FLAG_NO_LINE_KERNEL = 4 }; /* keep the warp-per-element kernel where the line-owner one applies */

typedef struct trixib200_config {
  int32_t ndim;                 /* 1, 2, 3 */
  int32_t polydeg;              /* DGSEMGPU(polydeg = ...), reference src/solvers/dgsem_gpu.jl:43-53 */
  int32_t equations;
  int32_t volume_integral;
  int32_t volume_flux;          /* volume_flux (FD) or volume_flux_dg (SC) */
  int32_t volume_flux_fv;       /* SC only */
  int32_t surface_flux;
  int32_t nonconservative;      /* 1: (flux, flux_nonconservative_powell) tuples (GLM-MHD) */
  int32_t indicator_variable;   /* IndicatorHennemannGassner(variable = ...) */
  int32_t alpha_smooth;
  int32_t boundary_conditions[6]; /* per direction -x,+x,-y,+y,-z,+z */
  int32_t initial_condition;    /* enumerated IC: used by BoundaryConditionDirichlet(ic) and fill_initial_condition */
  int32_t source_terms;
  int32_t device;               /* CUDA device ordinal of this process */
  int32_t rank, nranks;         /* Morton-curve partition: this handle owns one contiguous range of elements */
  int32_t flags;
  double alpha_max, alpha_min;
  double gamma;
  double advection_velocity[3];
  double c_h;                   /* IdealGlmMhdEquations3D.c_h */
} trixib200_config;

/* LobattoLegendreBasisGPU + MortarL2GPU operators (reference src/solvers/basis_lobatto_legendre.jl:25-47,
 * 155-173), column-major nnodes x nnodes unless noted; mortar operators may be NULL when nmortars == 0 */
typedef struct trixib200_basis_host {
  int32_t nnodes;
  const double* nodes;
  const double* weights;
  const double* inverse_weights;
  const double* derivative_dhat;
  const double* derivative_split;
  const double* boundary_interpolation;        /* nnodes x 2 */
  const double* inverse_vandermonde_legendre;
  const double* forward_upper;
  const double* forward_lower;
  const double* reverse_upper;
  const double* reverse_lower;
} trixib200_basis_host;

/* Containers exactly as Trixi's init_elements/init_interfaces/init_boundaries/init_mortars produce them for
 * the WHOLE mesh (reference src/solvers/cache.jl:130-158; layouts src/solvers/containers_3d.jl). All
 * pointers are host memory, copied during create() and may be freed afterwards. */
typedef struct trixib200_mesh_host {
  int64_t nelements, ninterfaces, nboundaries, nmortars;
  const double* inverse_jacobian;              /* [E] */
  const double* node_coordinates;              /* [ndim, N.., E]; may be NULL (then cell_centers is used) */
  const double* cell_centers;                  /* [ndim, E]; may be NULL if node_coordinates given */
  const int64_t* interfaces_neighbor_ids;      /* [2, I] */
  const int64_t* interfaces_orientations;      /* [I] */
  const int64_t* boundaries_neighbor_ids;      /* [B] */
  const int64_t* boundaries_orientations;      /* [B] */
  const int64_t* boundaries_neighbor_sides;    /* [B] */
  const double* boundaries_node_coordinates;   /* [ndim, N.., B] */
  const int64_t* n_boundaries_per_direction;   /* [2*ndim] */
  const int64_t* mortars_neighbor_ids;         /* [2^(ndim-1)+1, M] */
  const int64_t* mortars_large_sides;          /* [M] */
  const int64_t* mortars_orientations;         /* [M] */
} trixib200_mesh_host;

const char* trixib200_last_error(void);
int trixib200_version(void);

/* replaces: SemidiscretizationHyperbolicGPU(...) -> create_cache_gpu (reference
 * src/semidiscretization/semidiscretization_hyperbolic.jl:60-87, src/solvers/cache.jl:130-212) */
int trixib200_create(const trixib200_config* cfg, const trixib200_basis_host* basis,
                     const trixib200_mesh_host* mesh, trixib200_handle** out);
int trixib200_destroy(trixib200_handle* h);

/* sizes: "nelements" (local), "nelements_global", "first_element" (0-based global index of the first local
 * element), "nvars", "nnodes", "ndofs" (local, per field), "nunknowns" (local length of u), "ninterfaces",
 * "nboundaries", "nmortars", "nhalo_faces", "fused" (1 if a fused kernel is active), "warp3d" (1 if the
 * warp-per-element 3D flux-differencing kernel is available), "line3d" (1 if the line-owner 3D Euler
 * flux-differencing kernel is the one in use) */
int64_t trixib200_size(const trixib200_handle* h, const char* name);

/* replaces: rhs_gpu!(du_ode, u_ode, semi, t) (reference src/solvers/solvers.jl:18-31 -> src/solvers/dg_3d.jl:895-925).
 * du, u: DEVICE pointers to the local part of the flat vectors (length nunknowns). Asynchronous on the
 * handle's stream; du is fully overwritten. */
int trixib200_rhs(trixib200_handle* h, double* du, const double* u, double t);

/* Same operation on HOST vectors (length nunknowns): host->device copy of u, rhs!, device->host copy of du on
 * the handle's streams, synchronous. This is the entry point for callers whose state lives in host memory
 * (Trixi's CPU `rhs!(du_ode, u_ode, semi, t)` signature); the transfers are pipelined over element chunks so
 * that upload, kernels and download overlap. Host buffers should be page-locked (trixib200_host_register). */
int trixib200_rhs_host(trixib200_handle* h, double* du_host, const double* u_host, double t);
/* page-lock / unlock a caller-owned host buffer of n doubles (cudaHostRegister) */
int trixib200_host_register(trixib200_handle* h, double* host, int64_t n);
int trixib200_host_unregister(trixib200_handle* h, double* host);

/* replaces: max_dt(u, t, mesh, constant_speed, equations, dg, cache) (reference
 * src/callbacks_step/stepsize_dg_3d.jl:1-45). Device reduction (+ allreduce-max over ranks); synchronises. */
int trixib200_max_dt(trixib200_handle* h, const double* u, double t, double* out_host);

/* Per-stage entry points in the reference's stage order (reference src/solvers/dg_3d.jl:895-925), operating on
 * the handle's materialised containers in Trixi's layouts -- what the reference's per-stage tests compare
 * (reference test/tree_dgsem_3d/euler_ec.jl:57-121). Stage names: "reset_du", "calc_volume_integral",
 * "prolong2interfaces", "calc_interface_flux", "prolong2boundaries", "calc_boundary_flux", "prolong2mortars",
 * "calc_mortar_flux", "calc_surface_integral", "apply_jacobian", "calc_sources", "calc_indicator". */
int trixib200_stage(trixib200_handle* h, const char* stage, double* du, const double* u, double t);
/* copy a cache array to the host: "interfaces.u", "boundaries.u", "surface_flux_values", "alpha",
 * "mortars.u_upper_left|u_upper_right|u_lower_left|u_lower_right" (3D), "mortars.u_upper|u_lower" (2D) */
int64_t trixib200_cache_len(const trixib200_handle* h, const char* name);
int trixib200_cache_get(trixib200_handle* h, const char* name, double* out_host, int64_t n);

/* device memory + transfers for callers that own no CUDA runtime of their own (the C driver, the Julia shim's
 * vector type). upload/download are host<->device copies of n doubles on the handle's stream, synchronous. */
int trixib200_alloc(trixib200_handle* h, int64_t n, double** out_dev);
int trixib200_free(trixib200_handle* h, double* dev);
int trixib200_upload(trixib200_handle* h, double* dst_dev, const double* src_host, int64_t n);
int trixib200_download(trixib200_handle* h, double* dst_host, const double* src_dev, int64_t n);
int trixib200_sync(trixib200_handle* h);
/* cudaStream_t of the handle as an integer, and adoption of a caller-owned stream (PyTorch's / CUDA.jl's
 * current stream) so that library work is stream-ordered with the caller's own kernels */
int64_t trixib200_stream(const trixib200_handle* h);
int trixib200_set_stream(trixib200_handle* h, int64_t cuda_stream);

/* enumerated initial conditions evaluated on the device (benchmark / C driver only; the Julia shim uses Trixi's
 * compute_coefficients on the host and uploads, cf. reference src/solvers/solvers.jl:49-51) */
int trixib200_fill_initial_condition(trixib200_handle* h, double* u, double t);

/* 2N low-storage Runge-Kutta stage update (CarpenterKennedy2N54): tmp = a*tmp + dt*du; u += b*tmp */
int trixib200_rk2n_update(trixib200_handle* h, double* u, double* tmp, const double* du, double a, double b,
                          double dt);

/* rhs! fused with one 2N low-storage Runge-Kutta stage (the caller is OrdinaryDiffEq's
 * `perform_step!` for `CarpenterKennedy2N54`, outside the reference, whose stage is `rhs_gpu!` (reference
 * src/solvers/solvers.jl:18-31) followed by two broadcasts):
 *     tmp = a * tmp + dt * rhs(u_in, t);   u_out = u_in + b * tmp
 * On the line-owner kernel path this is ONE launch and du is never written to memory (160 instead of 280 B/DOF of HBM
 * traffic per stage); other kernel families run rhs! into a library-owned scratch vector followed by one update
 * kernel. u_out must not alias u_in. tmp is not read when a == 0 (first stage). Device pointers, asynchronous. */
int trixib200_rk2n_stage(trixib200_handle* h, double* u_out, const double* u_in, double* tmp, double t, double a,
                         double b, double dt);
/* one full CarpenterKennedy2N54 step t -> t + dt (five fused stages ping-ponging between u and u_alt); the result
 * is in u_alt iff *result_in_alt == 1 (five stages: it is) */
int trixib200_rk2n_step_ck54(trixib200_handle* h, double* u, double* u_alt, double* tmp, double t, double dt,
                             int* result_in_alt);

/* Device-side AnalysisCallback pieces. replaces: calc_error_norms(cons2cons, u, t,
 * analyzer, mesh, equations, initial_condition, dg, cache, cache_analysis) and integrate(cons2cons, u, ...)
 * (reference src/callbacks_step/analysis_dg_3d.jl:45-89 and :1-42, analysis_dg_2d.jl, analysis_dg_1d.jl), which copy u
 * and node_coordinates to the host and loop serially. `vandermonde` is the SolutionAnalyzer's interpolation matrix,
 * ROW-major [n_analysis][nnodes], `weights` its quadrature weights [n_analysis] (host pointers); the initial condition
 * is the handle's enumerated one evaluated at time t; total_volume = total_volume(mesh). Outputs are host arrays of
 * nvars doubles: l2 = sqrt(sum diff^2 w J / total_volume), linf = max |diff| -- reduced over all ranks. Synchronous. */
int trixib200_calc_error_norms(trixib200_handle* h, const double* u, double t, int32_t n_analysis,
                               const double* vandermonde, const double* weights, double total_volume,
                               double* l2_out, double* linf_out);
/* integral of the conserved variables over the domain, divided by total_volume if normalize != 0 */
int trixib200_integrate(trixib200_handle* h, const double* u, int32_t normalize, double total_volume, double* out);

/* timing helpers: run rhs `reps` times back to back and return the elapsed device time in milliseconds,
 * measured with CUDA events on the handle's stream */
int trixib200_time_rhs(trixib200_handle* h, double* du, const double* u, double t, int reps, float* ms_out);
/* number of kernels launched by this handle since create() */
int64_t trixib200_launch_count(const trixib200_handle* h);

/* multi-GPU (one process per GPU): rank 0 obtains an id, the host broadcasts it, every rank calls comm_init.
 * Halo-face traces are exchanged with ncclSend/ncclRecv inside trixib200_rhs. */
int trixib200_comm_unique_id(char* id128);
int trixib200_comm_init(trixib200_handle* h, const char* id128);

/* Host-only view of the partition plan create() would build for cfg->rank (no CUDA needed): local interface
 * list, halo send plan, face neighbour table. Arrays by name, widened to int64: "if_left", "if_right",
 * "if_dim", "if_global", "face_nbr", "send_elem", "send_dir", "send_global_iface", "peers", "peer_count",
 * "elems_interior", "elems_halo", "bd_elem", "bd_global", "mo_ids"; scalars via plan_len: "first_element",
 * "nelements". Halo sides / faces are encoded as -2 - slot, boundary/mortar faces as -1. */
int trixib200_plan_create(const trixib200_config* cfg, const trixib200_mesh_host* mesh, void** out_plan);
int trixib200_plan_destroy(void* plan);
int64_t trixib200_plan_len(void* plan, const char* name);
int trixib200_plan_get(void* plan, const char* name, int64_t* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* TRIXIB200_H */
