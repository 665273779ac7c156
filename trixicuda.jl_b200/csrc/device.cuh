// Device-side view of one handle (passed to kernels by value) and small index helpers.
#pragma once
#include <stdint.h>
#include "equations.cuh"

namespace tb {

constexpr int MAXN = 8;  // polydeg <= 7

// Operators of the LGL basis / L2 mortars, column-major with leading dimension N: M(i,j) = m[i + N*j]
// (reference src/solvers/basis_lobatto_legendre.jl:25-47,155-173).
struct Ops {
  double nodes[MAXN], weights[MAXN], inv_w[MAXN];
  double Dhat[MAXN * MAXN], Dsplit[MAXN * MAXN], invV[MAXN * MAXN];
  double fwd_u[MAXN * MAXN], fwd_l[MAXN * MAXN], rev_u[MAXN * MAXN], rev_l[MAXN * MAXN];
  double factor_1, factor_2;  // boundary_interpolation[1,1], boundary_interpolation[N,2]
};

// Neighbour encoding for element faces / interface sides: >= 0 local element, NB_SFV: flux comes from the
// materialised surface_flux_values (boundary / mortar face), <= NB_HALO0: halo slot (-2 - code).
constexpr int NB_SFV = -1;
constexpr int NB_HALO0 = -2;
__host__ __device__ inline bool nb_is_halo(int c) { return c <= NB_HALO0; }
__host__ __device__ inline int nb_halo_slot(int c) { return NB_HALO0 - c; }
__host__ __device__ inline int nb_from_halo_slot(int s) { return NB_HALO0 - s; }

struct Dev {
  int ndim, N, nn, nf, nv;
  int64_t E, I, B, M;          // local counts
  int64_t nhalo_recv, nhalo_send;
  const Ops* ops;
  const double* inv_jac;       // [E]
  const double* node_coords;   // [ndim, nn, E] or null
  const double* centers;       // [ndim, E] or null
  // interfaces (local list): sides are local element ids or halo codes; dim = orientation-1
  const int* if_left; const int* if_right; const int* if_dim;
  double* interfaces_u;        // [2, nv, nf, I]
  double* sfv;                 // surface_flux_values [nv, nf, 2*ndim, E]
  // boundaries
  const int* bd_elem; const int* bd_dim; const int* bd_side; const int* bd_dir;
  const double* bd_coords;     // [ndim, nf, B]
  double* boundaries_u;        // [2, nv, nf, B]
  // mortars: ids [nsmall+1, M] local element ids (rows as Trixi: 3D lower_left, lower_right, upper_left,
  // upper_right, large; 2D lower, upper, large)
  const int* mo_ids; const int* mo_side; const int* mo_dim;
  double* mortar_u[4];         // storage order 3D: upper_left, upper_right, lower_left, lower_right; 2D: upper, lower
  double* fstar_p[4]; double* fstar_s[4];  // [nv, nf, M]
  // shock capturing
  double* alpha; double* alpha_tmp;  // [E]
  // per-element face neighbours for the fused path: [E, 2*ndim]
  const int* face_nbr;
  // halo
  const int* send_elem; const int* send_dir;  // [nhalo_send]
  double* halo_send; const double* halo_recv; // [nv, nf, nhalo]
  double* halo_alpha_send; double* halo_alpha_recv;   // [nhalo]: indicator value of the element behind every exchanged face
  // physics
  EqPrm prm;
  int volume_integral, vol_flux, fv_flux, surf_flux, noncons, ic, src, ind_var, alpha_smooth;
  int bc[6];
  double alpha_max, alpha_min;
};

// node index (0-based, i + N j + N^2 k) of face node f = a + N b on the face of dim d at position `fixed`
template <int ND> TB_D int face_node(int N, int d, int fixed, int f) {
  if (ND == 1) return fixed;
  if (ND == 2) return d == 0 ? fixed + N * f : f + N * fixed;
  int a = f % N, b = f / N;
  if (d == 0) return fixed + N * a + N * N * b;
  if (d == 1) return a + N * fixed + N * N * b;
  return a + N * b + N * N * fixed;
}
TB_D int ipow_stride(int N, int d) { return d == 0 ? 1 : (d == 1 ? N : N * N); }

// mortar storage slot q -> neighbor_ids row
template <int ND> TB_D int mortar_small_row(int q) {
  if (ND == 3) return q == 0 ? 2 : (q == 1 ? 3 : (q == 2 ? 0 : 1));
  return q == 0 ? 1 : 0;
}

// cudaFuncSetAttribute is per device / context: remember per kernel WHICH devices have been configured (a second
// handle on another GPU of the same process otherwise launches without the shared-memory opt-in).
struct DeviceOnce {
  unsigned long long mask = 0;
  bool need() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
  void undo() {
    int dev = 0;
    cudaGetDevice(&dev);
    mask &= ~(1ull << (dev & 63));
  }
};

}  // namespace tb
