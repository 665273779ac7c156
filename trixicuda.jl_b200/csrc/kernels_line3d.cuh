// Line-owner fused rhs! kernel for 3D compressible Euler flux differencing at polydeg 3 (north-star path:
// entropy-conserving flux_ranocha volume + surface flux; other symmetric volume fluxes take the generic branch).
//
// Why a second design: ncu on the warp-per-element kernel (profiles/r1_ncu_warp3d_v1_l6.json) showed the shared-
// memory pipe at 73 % (766 wavefronts per element) with the FP64 pipe at 44 %: exchanging every pair flux between
// lanes costs three 8-byte shared accesses per flux value. Here a half-warp owns an element and each lane owns a
// whole 4-node LINE of it, in three phases (x-lines, y-lines, z-lines):
//  * all 6 symmetric node pairs of a line are evaluated in registers -- no flux exchange at all;
//  * each line ends on two faces of the element, so every lane also evaluates exactly 2 surface fluxes per phase
//    (6 volume + 2 surface evaluations per lane and phase, no idle lanes, 8 independent dependency chains);
//  * between phases only the 5 accumulators per node are handed over through shared memory, in a swizzled SoA
//    layout p(i,j,k) = 16k + 4((j+k)&3) + ((i+k)&3) that is bank-conflict-free for x-, y- and z-line access;
//  * neighbour face traces arrive as 16-byte cp.async chunks of the contiguous runs of the neighbour's block
//    (z-face: one 640 B run, y-face: four 160 B runs, x-face: sixteen 48 B windows), prefetched one element pair
//    ahead, as is the element block itself; du leaves through a shared AoS tile with 16-byte coalesced stores.
// HBM traffic is unchanged (u read once + face re-reads from L2, du written once; nothing else is materialised,
// unlike reference src/solvers/dg_3d.jl:895-925). Shared traffic drops to ~320 wavefronts per element and the FP64
// instruction count by ~20 % (Taylor branch of ln_mean without a second reciprocal, beta = rho/p per node).
#pragma once
#include <type_traits>
#include "device.cuh"
#include "kernels_warp3d.cuh"  // cp_async helpers

namespace tb {

constexpr int L3_WARPS = 4;            // warps per CTA; each warp works on two elements at a time
constexpr int L3_NQ = 6;               // per-node working set: rho, v1, v2, v3, p, beta = rho / p
constexpr int L3_STG = 2 * 320;        // AoS staging of two element blocks (doubles)
constexpr int L3_SQ = 2 * L3_NQ * 64;  // swizzled SoA q of two elements; reused as the AoS tile of du
constexpr int L3_SACC = 2 * 5 * 64;    // swizzled SoA accumulators
constexpr int L3_TRX = 2 * 16 * 6, L3_TRY = 2 * 16 * 5, L3_TRZ = 2 * 16 * 5;   // per element
constexpr int L3_TR = L3_TRX + L3_TRY + L3_TRZ;                                  // 512 doubles per element
constexpr int L3_PER_WARP = L3_STG + L3_SQ + L3_SACC + 2 * L3_TR;
constexpr size_t L3_SMEM = (size_t)L3_PER_WARP * L3_WARPS * sizeof(double);

// by-value operator block: lands in the constant bank, so the weights are immediate operands of the DFMAs
struct LineOps {
  double ds[16];   // Dsplit, column-major: ds[a + 4 b] = Dsplit[a, b]
  double factor_1, factor_2;
};

TB_D int l3_swz(int i, int j, int k) { return 16 * k + 4 * ((j + k) & 3) + ((i + k) & 3); }

// ---- Euler working variables and the Ranocha flux on them
TB_D void l3_to_q(const double* u, const EqPrm& p, double* q) {
  const double r = rcp_fast(u[0]);
  double ke = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) { q[1 + d] = u[1 + d] * r; ke = fma(u[1 + d], q[1 + d], ke); }
  q[0] = u[0];
  q[4] = (p.gamma - 1) * fma(-0.5, ke, u[4]);
  q[5] = u[0] * rcp_fast(q[4]);
}

// flux_ranocha from the two logarithmic means (Trixi flux_ranocha), orientation O = 1, 2, 3
template <int O>
TB_D void l3_ranocha_from_means(const double* ql, const double* qr, double rho_mean, double inv_rho_p_mean,
                                const EqPrm& p, double* f) {
  const double sv1 = ql[1] + qr[1], sv2 = ql[2] + qr[2], sv3 = ql[3] + qr[3];
  const double vv = fma(ql[3], qr[3], fma(ql[2], qr[2], ql[1] * qr[1]));
  const double psum = ql[4] + qr[4];
  const double pv = fma(ql[4], qr[O], qr[4] * ql[O]);
  const double svo = (O == 1) ? sv1 : (O == 2 ? sv2 : sv3);
  const double f1 = rho_mean * (0.5 * svo);
  const double hf = 0.5 * f1;
  f[0] = f1;
  f[1] = (O == 1) ? fma(0.5, psum, hf * sv1) : hf * sv1;
  f[2] = (O == 2) ? fma(0.5, psum, hf * sv2) : hf * sv2;
  f[3] = (O == 3) ? fma(0.5, psum, hf * sv3) : hf * sv3;
  f[4] = fma(f1, fma(inv_rho_p_mean, p.inv_gm1, 0.5 * vv), 0.5 * pv);
}

// Smooth-branch means: ln_mean(rho_l, rho_r) = s/2 * (1 - f2/3 - 4 f2^2/45 - 44 f2^3/945), f2 = ((x-y)/(x+y))^2 < 1e-4
// (the series of Trixi's (x+y)/(2 + f2(2/3 + f2(2/5 + 2 f2/7))), truncation 3e-18), where one Newton step on the
// reciprocal suffices because it only enters through f2; 1/ln_mean(beta_l, beta_r) = Trixi's p_l p_r *
// inv_ln_mean(rho_l p_r, rho_r p_l). Returns true if either mean needs the logarithmic branch.
TB_D bool l3_means_taylor(const double* ql, const double* qr, double& rho_mean, double& inv_rho_p_mean) {
  const double s = ql[0] + qr[0];
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  r = fma(r, fma(-s, r, 1.0), r);
  const double uu = (ql[0] - qr[0]) * r, f2 = uu * uu;
  rho_mean = s * fma(f2, fma(f2, fma(f2, -22.0 / 945, -2.0 / 45), -1.0 / 6), 0.5);
  const double sb = ql[5] + qr[5], rb = rcp_fast(sb);
  const double ub = (ql[5] - qr[5]) * rb, g2 = ub * ub;
  inv_rho_p_mean = rb * fma(g2, fma(g2, fma(g2, 2.0 / 7, 2.0 / 5), 2.0 / 3), 2.0);
  return !(f2 < 1.0e-4) || !(g2 < 1.0e-4);
}
// exact branches, out of line (rare: strong jumps)
__device__ __noinline__ void l3_means_exact(double rl, double rr, double bl, double br, double* out) {
  {
    const double s = rl + rr, uu = (rl - rr) / s, f2 = uu * uu;
    out[0] = (f2 < 1.0e-4) ? s / (2 + f2 * (2.0 / 3 + f2 * (2.0 / 5 + f2 * (2.0 / 7)))) : (rr - rl) / log(rr / rl);
  }
  {
    const double s = bl + br, uu = (bl - br) / s, f2 = uu * uu;
    out[1] = (f2 < 1.0e-4) ? (2 + f2 * (2.0 / 3 + f2 * (2.0 / 5 + f2 * (2.0 / 7)))) / s : log(br / bl) / (br - bl);
  }
}

// two-point flux of one pair; returns the "needs exact means" flag (always false on the generic branch)
template <int O, int KIND>
TB_D bool l3_flux(int kind_rt, const double* ql, const double* qr, const EqPrm& p, double* f) {
  if (KIND == TRIXIB200_FLUX_RANOCHA) {
    double rm, im;
    const bool slow = l3_means_taylor(ql, qr, rm, im);
    l3_ranocha_from_means<O>(ql, qr, rm, im, p, f);
    return slow;
  } else {
    EqEuler<3>::two_point_qf(kind_rt, ql, qr, O, p, f);
    return false;
  }
}
// cold fix-up: f_exact - f_taylor of a flagged pair. Everything crosses the call by value so that the hot path's
// register arrays never have their address taken (no local-memory copies).
struct L3Vec5 { double v[5]; };
struct L3Q { double v[6]; };
template <int O>
__device__ __noinline__ L3Vec5 l3_flux_correction(L3Q a, L3Q b, double gamma, double inv_gm1) {
  EqPrm p;
  p.gamma = gamma; p.inv_gm1 = inv_gm1;
  double rm, im, ex[2], ft[5], fe[5];
  l3_means_taylor(a.v, b.v, rm, im);
  l3_ranocha_from_means<O>(a.v, b.v, rm, im, p, ft);
  l3_means_exact(a.v[0], b.v[0], a.v[5], b.v[5], ex);
  l3_ranocha_from_means<O>(a.v, b.v, ex[0], ex[1], p, fe);
  L3Vec5 df;
#pragma unroll
  for (int v = 0; v < 5; ++v) df.v[v] = fe[v] - ft[v];
  return df;
}
TB_D L3Q l3_pack(const double* q) {
  L3Q r;
#pragma unroll
  for (int v = 0; v < 6; ++v) r.v[v] = q[v];
  return r;
}

template <int VFLUX, int SFLUX>
__global__ void __launch_bounds__(32 * L3_WARPS, 2)
k_line3d(Dev d, LineOps ops, double* __restrict__ du, const double* __restrict__ u, double t,
         const int* __restrict__ elems, int64_t count) {
  using Eq = EqEuler<3>;
  constexpr int NV = 5, NQ = L3_NQ, NN = 64;
  extern __shared__ __align__(16) double smem_l3[];
  double* smem = smem_l3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  double* wbase = smem + (size_t)warp * L3_PER_WARP;
  double* stg = wbase + half * 320;                   // this element's AoS block
  double* sq = wbase + L3_STG + half * (NQ * NN);      // [NQ][64] swizzled
  double* outt = wbase + L3_STG + half * 320;          // AoS tile of du, aliases the sq region of the warp
  double* sacc = wbase + L3_STG + L3_SQ + half * (NV * NN);
  double* tr = wbase + L3_STG + L3_SQ + L3_SACC + half * L3_TR;
  double* trf[6] = {tr, tr + 96, tr + L3_TRX, tr + L3_TRX + 80, tr + L3_TRX + L3_TRY, tr + L3_TRX + L3_TRY + 80};
  const EqPrm prm = d.prm;
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int64_t npairs = (count + 1) >> 1;
  const int64_t wid = (int64_t)blockIdx.x * L3_WARPS + warp, nw = (int64_t)gridDim.x * L3_WARPS;

  // element of this half-warp in pair `pr` (the odd tail duplicates the last element; its store is masked)
  auto elem_of = [&](int64_t pr, bool& valid) -> int64_t {
    int64_t s = 2 * pr + half;
    valid = s < count;
    if (!valid) s = count - 1;
    return elems ? (int64_t)elems[s] : s;
  };
  auto issue_block = [&](int64_t e) {
    const double* ue = u + (size_t)NV * NN * e;
#pragma unroll
    for (int m = 0; m < 10; ++m) cp_async16(stg + 2 * (l16 + 16 * m), ue + 2 * (l16 + 16 * m));
  };
  // face traces of direction DIR (faces 2 DIR, 2 DIR + 1) of element e; returns the two neighbour codes
  auto issue_traces = [&](auto dir_tag, int64_t e, const int c0, const int c1) {
    constexpr int dir = decltype(dir_tag)::value;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int face = 2 * dir + sd;
      const int code = sd == 0 ? c0 : c1;
      double* dst = trf[face];
      if (code >= 0) {
        const double* nb = u + (size_t)NV * NN * code;
        if (dir == 0) {
          // neighbour nodes i = 3 (low face) or i = 0 (high face): 48-byte aligned windows, 3 chunks each
#pragma unroll
          for (int c = l16; c < 48; c += 16) {
            const int f = c / 3, w = c - 3 * f;
            cp_async16(dst + 6 * f + 2 * w, nb + 20 * f + (sd == 0 ? 14 : 0) + 2 * w);
          }
        } else if (dir == 1) {
          // neighbour nodes j = 3 / j = 0: four runs of 160 B (k = 0..3)
#pragma unroll
          for (int c = l16; c < 48; c += 16) {
            if (c < 40) {
              const int k = c / 10, w = c - 10 * k;
              cp_async16(dst + 20 * k + 2 * w, nb + 80 * k + (sd == 0 ? 60 : 0) + 2 * w);
            }
          }
        } else {
#pragma unroll
          for (int c = l16; c < 48; c += 16)
            if (c < 40) cp_async16(dst + 2 * c, nb + (sd == 0 ? 240 : 0) + 2 * c);
        }
      } else {
        // ready-made flux (boundary / mortar face) or halo trace: dense [f][v], 640 B
        const double* src = (code == NB_SFV) ? d.sfv + (size_t)NV * 16 * (face + (size_t)6 * e)
                                             : d.halo_recv + (size_t)nb_halo_slot(code) * 16 * NV;
#pragma unroll
        for (int c = l16; c < 48; c += 16)
          if (c < 40) cp_async16(dst + 2 * c, src + 2 * c);
      }
    }
  };

  using D0 = std::integral_constant<int, 0>;
  using D1 = std::integral_constant<int, 1>;
  using D2 = std::integral_constant<int, 2>;
  auto load_codes = [&](int64_t el, int* c) {
    const int2* p = reinterpret_cast<const int2*>(d.face_nbr + (size_t)el * 6);   // 24 B per element, 8 B aligned
    const int2 a = p[0], b = p[1], cc = p[2];
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = cc.x; c[5] = cc.y;
  };
  int code[6] = {NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV};
  int64_t pr = wid;
  bool valid = false;
  int64_t e = 0;
  if (pr < npairs) {
    e = elem_of(pr, valid);
    load_codes(e, code);
    issue_block(e); cp_async_commit();
    issue_traces(D0{}, e, code[0], code[1]); cp_async_commit();
    issue_traces(D1{}, e, code[2], code[3]); cp_async_commit();
    issue_traces(D2{}, e, code[4], code[5]); cp_async_commit();
  }

  // one phase: the lane owns the line `l16` of direction DIR (nodes m = 0..3 along DIR)
  auto phase = [&](auto dir_tag, double (&acc)[4][NV], const int c_lo, const int c_hi) {
    constexpr int DIR = decltype(dir_tag)::value;
    constexpr int O = DIR + 1;
    const int a = l16 & 3, b = l16 >> 2;
    double q[4][NQ];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int pos = (DIR == 0) ? l3_swz(m, a, b) : (DIR == 1 ? l3_swz(a, m, b) : l3_swz(a, b, m));
#pragma unroll
      for (int v = 0; v < NQ; ++v) q[m][v] = sq[v * NN + pos];
    }
    unsigned slow = 0;
    // ---- volume: all 6 pairs of the line (reference dg_3d_kernel.jl:188-257 evaluates 12 per node)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
#pragma unroll
      for (int y = x + 1; y < 4; ++y) {
        double f[NV];
        const bool s = l3_flux<O, VFLUX>(vflux, q[x], q[y], prm, f);
        slow |= (unsigned)s << (x * 4 + y);
        const double wxy = ops.ds[x + 4 * y], wyx = ops.ds[y + 4 * x];
#pragma unroll
        for (int v = 0; v < NV; ++v) { acc[x][v] = fma(wxy, f[v], acc[x][v]); acc[y][v] = fma(wyx, f[v], acc[y][v]); }
      }
    }
    if (VFLUX == TRIXIB200_FLUX_RANOCHA && slow != 0) {
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int y = x + 1; y < 4; ++y) {
          if (slow & (1u << (x * 4 + y))) {
            const L3Vec5 df = l3_flux_correction<O>(l3_pack(q[x]), l3_pack(q[y]), prm.gamma, prm.inv_gm1);
            const double wxy = ops.ds[x + 4 * y], wyx = ops.ds[y + 4 * x];
#pragma unroll
            for (int v = 0; v < NV; ++v) { acc[x][v] = fma(wxy, df.v[v], acc[x][v]); acc[y][v] = fma(wyx, df.v[v], acc[y][v]); }
          }
        }
      }
    }
    // ---- surface: the line ends on faces 2 DIR (node 0) and 2 DIR + 1 (node 3); traces must have landed
    cp_async_wait<3>();
    __syncwarp();
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int cd = sd == 0 ? c_lo : c_hi;
      const double* src = trf[2 * DIR + sd];
      int stride = NV, off = 0;
      if (DIR == 0 && cd >= 0) { stride = 6; off = sd == 0 ? 1 : 0; }
      double nbv[NV], fl[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) nbv[v] = src[l16 * stride + off + v];
      if (cd == NB_SFV) {
#pragma unroll
        for (int v = 0; v < NV; ++v) fl[v] = nbv[v];
      } else {
        double qn[NQ];
        l3_to_q(nbv, prm, qn);
        // (ll, rr) = (neighbour, own) on the low face, (own, neighbour) on the high face: both elements of an
        // interface evaluate the identical expression
        const double* qa = sd == 0 ? qn : q[3];
        const double* qb = sd == 0 ? q[0] : qn;
        const bool s = l3_flux<O, SFLUX>(sflux, qa, qb, prm, fl);
        if (SFLUX == TRIXIB200_FLUX_RANOCHA && s) {
          const L3Vec5 df = l3_flux_correction<O>(l3_pack(qa), l3_pack(qb), prm.gamma, prm.inv_gm1);
#pragma unroll
          for (int v = 0; v < NV; ++v) fl[v] += df.v[v];
        }
      }
      // surface integral (reference dg_3d_kernel.jl:1787-1794)
      if (sd == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[0][v] = fma(-ops.factor_1, fl[v], acc[0][v]);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[3][v] = fma(ops.factor_2, fl[v], acc[3][v]);
      }
    }
  };

  for (; pr < npairs; pr += nw) {
    const int64_t pr_next = pr + nw;
    const bool has_next = pr_next < npairs;
    bool valid_next = false;
    const int64_t e_next = has_next ? elem_of(pr_next, valid_next) : 0;
    int code_next[6] = {NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV};
    if (has_next) load_codes(e_next, code_next);   // issued early, consumed when the trace copies are issued

    // ---- block has landed: cons -> (rho, v, p, beta), z-line ownership (conflict-free AoS reads)
    cp_async_wait<3>();
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double un[NV], qn[NQ];
#pragma unroll
      for (int v = 0; v < NV; ++v) un[v] = stg[NV * (l16 + 16 * m) + v];
      l3_to_q(un, prm, qn);
      const int pos = l3_swz(l16 & 3, l16 >> 2, m);
#pragma unroll
      for (int v = 0; v < NQ; ++v) sq[v * NN + pos] = qn[v];
    }
    __syncwarp();
    if (has_next) issue_block(e_next);
    cp_async_commit();

    double acc[4][NV];
    // ---- x lines
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[m][v] = 0;
    phase(D0{}, acc, code[0], code[1]);
    {
      const int a = l16 & 3, b = l16 >> 2;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int pos = l3_swz(m, a, b);
#pragma unroll
        for (int v = 0; v < NV; ++v) sacc[v * NN + pos] = acc[m][v];
      }
    }
    __syncwarp();
    if (has_next) issue_traces(D0{}, e_next, code_next[0], code_next[1]);
    cp_async_commit();
    // ---- y lines
    {
      const int a = l16 & 3, b = l16 >> 2;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int pos = l3_swz(a, m, b);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[m][v] = sacc[v * NN + pos];
      }
      phase(D1{}, acc, code[2], code[3]);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int pos = l3_swz(a, m, b);
#pragma unroll
        for (int v = 0; v < NV; ++v) sacc[v * NN + pos] = acc[m][v];
      }
    }
    __syncwarp();
    if (has_next) issue_traces(D1{}, e_next, code_next[2], code_next[3]);
    cp_async_commit();
    // ---- z lines, then Jacobian, sources, output
    {
      const int a = l16 & 3, b = l16 >> 2;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int pos = l3_swz(a, b, m);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[m][v] = sacc[v * NN + pos];
      }
      phase(D2{}, acc, code[4], code[5]);   // its __syncwarp also ends every lane's reads of sq
      const double inv_jac = d.inv_jac[e];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[m][v] *= -inv_jac;
        const int n = l16 + 16 * m;
        if (d.src != TRIXIB200_SRC_NONE) {
          double x[3], un[NV], sv[NV];
          if (d.node_coords) {
#pragma unroll
            for (int c = 0; c < 3; ++c) x[c] = d.node_coords[c + (size_t)3 * (n + (size_t)NN * e)];
          } else {
            const double jac = 1.0 / inv_jac;
            const int idx[3] = {a, b, m};
#pragma unroll
            for (int c = 0; c < 3; ++c)
              x[c] = __dadd_rn(d.centers[c + (size_t)3 * e], __dmul_rn(jac, d.ops->nodes[idx[c]]));
          }
#pragma unroll
          for (int v = 0; v < NV; ++v) un[v] = u[(size_t)NV * (n + (size_t)NN * e) + v];
          Eq::source(d.src, un, x, t, prm, sv);
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[m][v] += sv[v];
        }
      }
      __syncwarp();   // all lanes are past their sq reads (surface part of the z phase used q[0], q[3] in registers)
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int v = 0; v < NV; ++v) outt[NV * (l16 + 16 * m) + v] = acc[m][v];
    }
    __syncwarp();
    if (has_next) issue_traces(D2{}, e_next, code_next[4], code_next[5]);
    cp_async_commit();
    if (valid) {
      double2* due = reinterpret_cast<double2*>(du + (size_t)NV * NN * e);
      const double2* o2 = reinterpret_cast<const double2*>(outt);
#pragma unroll
      for (int m = 0; m < 10; ++m) due[l16 + 16 * m] = o2[l16 + 16 * m];
    }
    __syncwarp();
    e = e_next; valid = valid_next;
#pragma unroll
    for (int c = 0; c < 6; ++c) code[c] = code_next[c];
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
inline bool line3d_available(const trixib200_config& c) {
  return c.ndim == 3 && c.polydeg == 3 && c.volume_integral == TRIXIB200_VI_FLUX_DIFFERENCING && !c.nonconservative &&
         c.equations == TRIXIB200_EQ_EULER;
}

template <int VFLUX, int SFLUX>
static int line3d_launch_t(const Dev& d, const LineOps& ops, double* du, const double* u, double t, const int* elems,
                           int64_t count, cudaStream_t stream, int sm_count) {
  auto kern = k_line3d<VFLUX, SFLUX>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L3_SMEM) != cudaSuccess)
      return TRIXIB200_ECUDA;
    configured = true;
  }
  if (count <= 0) return 0;
  const int64_t npairs = (count + 1) / 2;
  const int64_t want = (npairs + L3_WARPS - 1) / L3_WARPS;
  const unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * 2);
  kern<<<blocks, 32 * L3_WARPS, L3_SMEM, stream>>>(d, ops, du, u, t, elems, count);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

static int line3d_launch(const trixib200_config& c, const Dev& d, const LineOps& ops, double* du, const double* u,
                         double t, const int* elems, int64_t count, cudaStream_t s, int sm_count) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  if (c.volume_flux == R && c.surface_flux == R) return line3d_launch_t<R, R>(d, ops, du, u, t, elems, count, s, sm_count);
  if (c.volume_flux == R) return line3d_launch_t<R, -1>(d, ops, du, u, t, elems, count, s, sm_count);
  return line3d_launch_t<-1, -1>(d, ops, du, u, t, elems, count, s, sm_count);
}

}  // namespace tb
