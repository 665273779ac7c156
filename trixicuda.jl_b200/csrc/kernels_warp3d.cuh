// Warp-per-element fused rhs! kernel for 3D flux differencing at polydeg 3 (the north-star path:
// 3D compressible Euler, entropy-conserving flux_ranocha volume + surface flux).
//
// One warp owns one element at a time (persistent loop over a strided element list); each lane owns the two
// nodes (i, j, k) and (i, j, k+2). Consequences:
//  * no CTA-wide barriers at all: every exchange is inside the warp (__syncwarp only);
//  * two independent flux evaluations per lane per slot -> ILP for the dependent DFMA chains;
//  * the z matching r1 = {02, 13} pairs a lane's own two nodes: no exchange for that slot;
//  * 96 face nodes / 32 lanes = 3 surface fluxes per lane: together with 9 volume pair evaluations that is
//    exactly 6 flux evaluations per node with no idle lanes (symmetric pairs evaluated once).
// Memory: the element's 2560-byte block and its six neighbour face traces arrive by cp.async (LDGSTS) into
// shared memory; the next element's block is prefetched while the current one is being computed. du is
// transposed back to AoS in shared memory and stored with 16-byte coalesced writes. Nothing else touches HBM:
// no interfaces.u, no surface_flux_values (reference src/solvers/dg_3d.jl:895-925 materialises both).
#pragma once
#include "device.cuh"

namespace tb {

constexpr int W3_WARPS = 4;  // warps (elements in flight) per CTA

template <class Eq> struct Warp3Cfg {
  static constexpr int NV = Eq::NV, NN = 64, NFN = 96;
  // doubles per warp: staging (AoS block) | q SoA | exchange / du staging | neighbour traces -> surface fluxes
  static constexpr int PER_WARP = NV * NN * 3 + NV * NFN;
  static constexpr size_t SMEM = (size_t)(PER_WARP * W3_WARPS + 32) * sizeof(double);
};

TB_D void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
TB_D void cp_async8(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
TB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> TB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class Eq, int VFLUX, int SFLUX>
__global__ void __launch_bounds__(32 * W3_WARPS, 4)
k_warp3d(Dev d, double* __restrict__ du, const double* __restrict__ u, double t, const int* __restrict__ elems,
         int64_t count) {
  using C = Warp3Cfg<Eq>;
  constexpr int NV = C::NV, NN = 64, NFN = 96, N = 4;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sD = smem;                                  // Dsplit [16] + inverse-weight / factors
  double* stg = smem + 32 + (size_t)warp * C::PER_WARP;  // [NN*NV] AoS staging of u
  double* sq = stg + NV * NN;                         // [NV][NN]
  double* sx = sq + NV * NN;                          // [NV][NN] exchange, later [NN][NV] du staging
  double* sn = sx + NV * NN;                          // [NFN][NV] neighbour traces, then surface fluxes
  const Ops& op = *d.ops;
  if (threadIdx.x < 16) sD[threadIdx.x] = op.Dsplit[threadIdx.x];
  if (threadIdx.x == 16) sD[16] = op.factor_1;
  if (threadIdx.x == 17) sD[17] = op.factor_2;
  __syncthreads();
  const double factor_1 = sD[16], factor_2 = sD[17];
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int i = lane & 3, j = (lane >> 2) & 3, kk = lane >> 4;
  const int nA = lane, nB = lane + 32, kA = kk, kB = kk + 2;
  const int64_t wid = (int64_t)blockIdx.x * W3_WARPS + warp, nw = (int64_t)gridDim.x * W3_WARPS;

  auto elem_of = [&](int64_t s) -> int64_t { return elems ? (int64_t)elems[s] : s; };
  auto issue_block = [&](int64_t e) {
    const double* ue = u + (size_t)NV * NN * e;
#pragma unroll
    for (int m = 0; m < (NV * NN) / 64; ++m) cp_async16(stg + 2 * (lane + 32 * m), ue + 2 * (lane + 32 * m));
    cp_async_commit();
  };
  // three face nodes per lane: faces (2t + kk), face node f = lane & 15
  auto issue_traces = [&](int64_t e, int* codes) {
#pragma unroll
    for (int tt = 0; tt < 3; ++tt) {
      const int face = 2 * tt + kk, f = lane & 15;
      const int code = d.face_nbr[(size_t)e * 6 + face];
      codes[tt] = code;
      const double* p;
      if (code >= 0) p = u + ((size_t)code * NN + face_node<3>(N, tt, (face & 1) ? 0 : N - 1, f)) * NV;
      else if (code == NB_SFV) p = d.sfv + (size_t)NV * (f + (size_t)16 * (face + (size_t)6 * e));
      else p = d.halo_recv + ((size_t)nb_halo_slot(code) * 16 + f) * NV;
      double* dst = sn + (face * 16 + f) * NV;
#pragma unroll
      for (int v = 0; v < NV; ++v) cp_async8(dst + v, p + v);
    }
    cp_async_commit();
  };

  int codes[3] = {NB_SFV, NB_SFV, NB_SFV};
  int64_t s = wid;
  if (s < count) {
    const int64_t e0 = elem_of(s);
    issue_block(e0);
    issue_traces(e0, codes);
  }
  for (; s < count; s += nw) {
    const int64_t e = elem_of(s);
    const int64_t s_next = s + nw;
    const bool has_next = s_next < count;
    const int64_t e_next = has_next ? elem_of(s_next) : 0;
    // ---- own block has landed (at most the trace group is still in flight)
    cp_async_wait<1>();
    __syncwarp();
    double qA[NV], qB[NV], accA[NV], accB[NV];
    {
      double ua[NV], ub[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) { ua[v] = stg[NV * nA + v]; ub[v] = stg[NV * nB + v]; accA[v] = 0; accB[v] = 0; }
      Eq::to_qf(ua, d.prm, qA);
      Eq::to_qf(ub, d.prm, qB);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) { sq[v * NN + nA] = qA[v]; sq[v * NN + nB] = qB[v]; }
    __syncwarp();
    // staging is free again: prefetch the next element's block behind the volume work
    if (has_next) issue_block(e_next); else cp_async_commit();

    // one exchange slot: both nodes of the lane evaluate the pair (self, self + dpc) and pick up the pair
    // (self, self + dpr) evaluated by that partner. dpc/dpr are node-index offsets, wc/wr the Dsplit weights.
    auto slot = [&](int o, int dpcA, int dpcB, int dprA, int dprB, double wcA, double wcB, double wrA, double wrB) {
      double qp[NV], qr[NV], fa[NV], fb[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) { qp[v] = sq[v * NN + nA + dpcA]; qr[v] = sq[v * NN + nB + dpcB]; }
      Eq::two_point_qf(vflux, qA, qp, o, d.prm, fa);
      Eq::two_point_qf(vflux, qB, qr, o, d.prm, fb);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        sx[v * NN + nA] = fa[v]; sx[v * NN + nB] = fb[v];
        accA[v] = fma(wcA, fa[v], accA[v]); accB[v] = fma(wcB, fb[v], accB[v]);
      }
      __syncwarp();
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        accA[v] = fma(wrA, sx[v * NN + nA + dprA], accA[v]);
        accB[v] = fma(wrB, sx[v * NN + nB + dprB], accB[v]);
      }
      __syncwarp();
    };
    {  // x: matchings r0 + r2
      const int pc = (i & 1) ? (i ^ 3) : (i ^ 1), pr = (i & 1) ? (i ^ 1) : (i ^ 3);
      const double wc = sD[i + 4 * pc], wr = sD[i + 4 * pr];
      slot(1, pc - i, pc - i, pr - i, pr - i, wc, wc, wr, wr);
    }
    {  // y: matchings r0 + r2
      const int pc = (j & 1) ? (j ^ 3) : (j ^ 1), pr = (j & 1) ? (j ^ 1) : (j ^ 3);
      const double wc = sD[j + 4 * pc], wr = sD[j + 4 * pr];
      slot(2, 4 * (pc - j), 4 * (pc - j), 4 * (pr - j), 4 * (pr - j), wc, wc, wr, wr);
    }
    {  // z: matchings r0 + r2 (node A has k = kk, node B has k = kk + 2)
      const int pcA = kk ? (kA ^ 3) : (kA ^ 1), prA = kk ? (kA ^ 1) : (kA ^ 3);
      const int pcB = kk ? (kB ^ 3) : (kB ^ 1), prB = kk ? (kB ^ 1) : (kB ^ 3);
      slot(3, 16 * (pcA - kA), 16 * (pcB - kB), 16 * (prA - kA), 16 * (prB - kB), sD[kA + 4 * pcA], sD[kB + 4 * pcB],
           sD[kA + 4 * prA], sD[kB + 4 * prB]);
    }
    {  // matchings r1 of x and y share a slot: selector bit1(i) ^ bit1(j)
      const bool bx = (((i >> 1) ^ (j >> 1)) & 1) == 0;
      const int o = bx ? 1 : 2;
      const int dpc = bx ? ((i ^ 2) - i) : 4 * ((j ^ 2) - j), dpr = bx ? 4 * ((j ^ 2) - j) : ((i ^ 2) - i);
      const double wc = bx ? sD[i + 4 * (i ^ 2)] : sD[j + 4 * (j ^ 2)], wr = bx ? sD[j + 4 * (j ^ 2)] : sD[i + 4 * (i ^ 2)];
      slot(o, dpc, dpc, dpr, dpr, wc, wc, wr, wr);
    }
    {  // z matching r1 pairs the lane's own two nodes
      double f[NV];
      Eq::two_point_qf(vflux, qA, qB, 3, d.prm, f);
      const double wA = sD[kA + 4 * kB], wB = sD[kB + 4 * kA];
#pragma unroll
      for (int v = 0; v < NV; ++v) { accA[v] = fma(wA, f[v], accA[v]); accB[v] = fma(wB, f[v], accB[v]); }
    }

    // ---- surface fluxes: traces have landed (only the prefetch of the next block may still be in flight)
    cp_async_wait<1>();
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < 3; ++tt) {
      const int face = 2 * tt + kk, f = lane & 15, side = face & 1;
      double* slotp = sn + (face * 16 + f) * NV;
      if (codes[tt] != NB_SFV) {
        const int own = face_node<3>(N, tt, side ? N - 1 : 0, f);
        double nb[NV], qo[NV], qn[NV], qa[NV], qb[NV], fl[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { nb[v] = slotp[v]; qo[v] = sq[v * NN + own]; }
        Eq::to_qf(nb, d.prm, qn);
#pragma unroll
        for (int v = 0; v < NV; ++v) { qa[v] = side ? qo[v] : qn[v]; qb[v] = side ? qn[v] : qo[v]; }
        Eq::two_point_qf(sflux, qa, qb, tt + 1, d.prm, fl);
#pragma unroll
        for (int v = 0; v < NV; ++v) slotp[v] = fl[v];
      }
    }
    __syncwarp();
    // ---- surface integral + Jacobian (+ sources), reference dg_3d_kernel.jl:1773-1844
    const double inv_jac = d.inv_jac[e];
    auto finish = [&](double* acc, int n, int k) {
      const int idx[3] = {i, j, k};
      const int fx = j + 4 * k, fy = i + 4 * k, fz = i + 4 * j;
      const int ff[3] = {fx, fy, fz};
#pragma unroll
      for (int dd = 0; dd < 3; ++dd) {
        if (idx[dd] == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] = fma(-factor_1, sn[((2 * dd) * 16 + ff[dd]) * NV + v], acc[v]);
        }
        if (idx[dd] == N - 1) {
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] = fma(factor_2, sn[((2 * dd + 1) * 16 + ff[dd]) * NV + v], acc[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] *= -inv_jac;
      if (d.src != TRIXIB200_SRC_NONE) {
        double x[3], sv[NV];
        if (d.node_coords) {
#pragma unroll
          for (int q = 0; q < 3; ++q) x[q] = d.node_coords[q + (size_t)3 * (n + (size_t)NN * e)];
        } else {
          const double jac = 1.0 / inv_jac;
#pragma unroll
          for (int q = 0; q < 3; ++q) x[q] = __dadd_rn(d.centers[q + (size_t)3 * e], __dmul_rn(jac, op.nodes[idx[q]]));
        }
        Eq::source(d.src, acc, x, t, d.prm, sv);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += sv[v];
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) sx[NV * n + v] = acc[v];
    };
    finish(accA, nA, kA);
    finish(accB, nB, kB);
    __syncwarp();
    // traces / surface fluxes are consumed: prefetch the next element's traces, then store du
    if (has_next) issue_traces(e_next, codes); else cp_async_commit();
    double2* due = reinterpret_cast<double2*>(du + (size_t)NV * NN * e);
    const double2* sx2 = reinterpret_cast<const double2*>(sx);
#pragma unroll
    for (int m = 0; m < (NV * NN) / 64; ++m) due[lane + 32 * m] = sx2[lane + 32 * m];
    __syncwarp();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
inline bool warp3d_available(const trixib200_config& c) {
  return c.ndim == 3 && c.polydeg == 3 && c.volume_integral == TRIXIB200_VI_FLUX_DIFFERENCING && !c.nonconservative &&
         (c.equations == TRIXIB200_EQ_EULER || c.equations == TRIXIB200_EQ_ADVECTION);
}

template <class Eq, int VFLUX, int SFLUX>
static int warp3d_launch_t(const Dev& d, double* du, const double* u, double t, const int* elems, int64_t count,
                           cudaStream_t stream, int sm_count) {
  using C = Warp3Cfg<Eq>;
  auto kern = k_warp3d<Eq, VFLUX, SFLUX>;
  static DeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
      { configured.undo(); return TRIXIB200_ECUDA; }
  }
  if (count <= 0) return 0;
  int64_t want = (count + W3_WARPS - 1) / W3_WARPS;
  unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * 4);
  kern<<<blocks, 32 * W3_WARPS, C::SMEM, stream>>>(d, du, u, t, elems, count);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

static int warp3d_launch(const trixib200_config& c, const Dev& d, double* du, const double* u, double t,
                         const int* elems, int64_t count, cudaStream_t s, int sm_count) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  if (c.equations == TRIXIB200_EQ_EULER) {
    if (c.volume_flux == R && c.surface_flux == R)
      return warp3d_launch_t<EqEuler<3>, R, R>(d, du, u, t, elems, count, s, sm_count);
    return warp3d_launch_t<EqEuler<3>, -1, -1>(d, du, u, t, elems, count, s, sm_count);
  }
  return warp3d_launch_t<EqAdvection<3>, -1, -1>(d, du, u, t, elems, count, s, sm_count);
}

}  // namespace tb
