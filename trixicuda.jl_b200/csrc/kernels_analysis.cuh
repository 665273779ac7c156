// Device-side AnalysisCallback pieces: calc_error_norms and integrate for the conserved
// variables. The reference copies u and node_coordinates to the host and loops serially over the elements
// (reference src/callbacks_step/analysis_dg_3d.jl:45-89 and :1-33, 2D/1D analogues; its own FIXMEs say so); here a
// fixed grid of CTAs walks the elements, every CTA keeps per-thread partial sums and writes ONE partial per variable,
// and the host adds the partials in CTA order -- deterministic, no atomics.
//
// The interpolation to the analysis nodes is Trixi's multiply_dimensionwise! (x, then y, then z; inner sum over the
// source node index in increasing order with fused multiply-adds), applied to u AND to the node coordinates exactly as
// calc_error_norms does, so the coordinates that enter the initial-condition function carry the same rounding as on the
// CPU path (matters for discontinuous initial conditions such as the weak blast wave).
#pragma once
#include "device.cuh"
#include "kernels_staged.cuh"

namespace tb {

constexpr int AN_THREADS = 256;
constexpr int AN_MAXV = 9;   // GLM-MHD

// doubles of shared memory: two ping-pong buffers of the largest intermediate, for nv + ndim components
inline size_t analysis_smem_doubles(int nd, int N, int NA, int nv) {
  size_t big = 1;
  for (int q = 0; q < nd; ++q) big *= (size_t)NA;
  return 2 * big * (size_t)(nv + nd) + (size_t)NA * N + NA + (size_t)AN_THREADS * 2;
}

// one dimension of multiply_dimensionwise!: out[c, .., i_d, ..] = sum_l V[i_d, l] * in[c, .., l, ..]
TB_D void an_stage(const double* in, double* out, const double* V, int ncomp, int N, int NA, const int* cur, int dim) {
  int nxt[3] = {cur[0], cur[1], cur[2]};
  nxt[dim] = NA;
  const int total = ncomp * nxt[0] * nxt[1] * nxt[2];
  for (int o = threadIdx.x; o < total; o += blockDim.x) {
    const int cc = o % ncomp;
    int r = o / ncomp;
    int id[3];
    id[0] = r % nxt[0]; r /= nxt[0];
    id[1] = r % nxt[1];
    id[2] = r / nxt[1];
    double s = 0;
    for (int l = 0; l < N; ++l) {
      int is[3] = {id[0], id[1], id[2]};
      is[dim] = l;
      s = fma(V[id[dim] * N + l], in[cc + ncomp * (is[0] + cur[0] * (is[1] + cur[1] * is[2]))], s);
    }
    out[o] = s;
  }
}

// interpolate one element's nodal field (component-fastest) to the analysis nodes; returns the buffer holding the result
TB_D double* an_interpolate(double* a, double* b, const double* V, int ncomp, int nd, int N, int NA) {
  int cur[3] = {nd > 0 ? N : 1, nd > 1 ? N : 1, nd > 2 ? N : 1};
  for (int q = 0; q < nd; ++q) {
    __syncthreads();
    an_stage(a, b, V, ncomp, N, NA, cur, q);
    cur[q] = NA;
    double* t = a; a = b; b = t;
  }
  __syncthreads();
  return a;
}

// block-wide reduction in thread order (deterministic): result valid in thread 0
template <bool MAX>
TB_D double an_block_reduce(double x, double* scratch) {
  __syncthreads();
  scratch[threadIdx.x] = x;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
    r = scratch[0];
    for (int i = 1; i < (int)blockDim.x; ++i) r = MAX ? fmax(r, scratch[i]) : r + scratch[i];
  }
  return r;
}

// part[(2 * block + 0) * nv + v] = sum over this CTA's elements of diff^2 * w * J^nd;  [.. + 1 ..] = max |diff|
template <class Eq>
__global__ void __launch_bounds__(AN_THREADS)
k_error_norms(Dev d, const double* __restrict__ u, double t, int NA, const double* __restrict__ Vg,
              const double* __restrict__ wg, double* __restrict__ part) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  extern __shared__ double an_smem[];
  const int N = d.N, nn = d.nn;
  int na = 1;
  for (int q = 0; q < ND; ++q) na *= NA;
  double* V = an_smem;                        // [NA][N]
  double* wa = V + NA * N;                    // [NA]
  double* scratch = wa + NA;                  // [2 * AN_THREADS]
  double* bu0 = scratch + 2 * AN_THREADS;     // u ping
  double* bu1 = bu0 + (size_t)na * NV;        // u pong
  double* bx0 = bu1 + (size_t)na * NV;        // x ping
  double* bx1 = bx0 + (size_t)na * ND;        // x pong
  for (int i = threadIdx.x; i < NA * N; i += blockDim.x) V[i] = Vg[i];
  for (int i = threadIdx.x; i < NA; i += blockDim.x) wa[i] = wg[i];
  double l2[NV], li[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { l2[v] = 0; li[v] = 0; }
  for (int64_t e = blockIdx.x; e < d.E; e += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < NV * nn; i += blockDim.x) bu0[i] = u[(size_t)NV * nn * e + i];
    const double jac = 1.0 / d.inv_jac[e];
    for (int i = threadIdx.x; i < ND * nn; i += blockDim.x) {
      const int q = i % ND, n = i / ND;
      const int idx[3] = {n % N, (n / N) % N, n / (N * N)};
      bx0[i] = d.node_coords ? d.node_coords[(size_t)ND * nn * e + i]
                             : __dadd_rn(d.centers[q + (size_t)ND * e], __dmul_rn(jac, d.ops->nodes[idx[q]]));
    }
    const double* ua = an_interpolate(bu0, bu1, V, NV, ND, N, NA);
    const double* xa = an_interpolate(bx0, bx1, V, ND, ND, N, NA);
    double vj = jac;
    for (int q = 1; q < ND; ++q) vj *= jac;
    for (int q = threadIdx.x; q < na; q += blockDim.x) {
      const int idx[3] = {q % NA, (q / NA) % NA, q / (NA * NA)};
      double w = 1;
      for (int c = 0; c < ND; ++c) w *= wa[idx[c]];
      double x[3] = {0, 0, 0}, ue[NV];
      for (int c = 0; c < ND; ++c) x[c] = xa[c + ND * q];
      Eq::initial_condition(d.ic, x, t, d.prm, ue);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double diff = ue[v] - ua[v + NV * q];
        l2[v] += diff * diff * (w * vj);
        li[v] = fmax(li[v], fabs(diff));
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double s = an_block_reduce<false>(l2[v], scratch);
    const double m = an_block_reduce<true>(li[v], scratch);
    if (threadIdx.x == 0) {
      part[(2 * (size_t)blockIdx.x + 0) * NV + v] = s;
      part[(2 * (size_t)blockIdx.x + 1) * NV + v] = m;
    }
  }
}

// part[block * nv + v] = sum over this CTA's elements of J^nd * w_n * u[v, n]   (integrate(cons2cons, ...))
template <class Eq>
__global__ void __launch_bounds__(AN_THREADS)
k_integrate(Dev d, const double* __restrict__ u, double* __restrict__ part) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  __shared__ double scratch[AN_THREADS];
  const int N = d.N, nn = d.nn;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0;
  for (int64_t e = blockIdx.x; e < d.E; e += gridDim.x) {
    const double jac = 1.0 / d.inv_jac[e];
    double vj = jac;
    for (int q = 1; q < ND; ++q) vj *= jac;
    for (int n = threadIdx.x; n < nn; n += blockDim.x) {
      const int idx[3] = {n % N, (n / N) % N, n / (N * N)};
      double w = vj;
      for (int q = 0; q < ND; ++q) w *= d.ops->weights[idx[q]];
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = fma(w, u[(size_t)NV * (n + (size_t)nn * e) + v], acc[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double s = an_block_reduce<false>(acc[v], scratch);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * NV + v] = s;
  }
}

}  // namespace tb
