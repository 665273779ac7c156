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

constexpr int L3_NQ = 5;               // per-node working set: rho, v1/2, v2/2, v3/2, p
constexpr int L3_STG = 2 * 320;        // AoS staging of two element blocks (prefetch target)
constexpr int L3_SQ = 2 * L3_NQ * 64;  // swizzled SoA q of two elements; later the AoS tile of du
constexpr int L3_SACC = 2 * 5 * 64;    // swizzled SoA accumulators handed from phase to phase
constexpr int L3_TRX = 2 * 16 * 6, L3_TRY = 2 * 16 * 5, L3_TRZ = 2 * 16 * 5;   // per element
constexpr int L3_TR = L3_TRX + L3_TRY + L3_TRZ;                                  // 512 doubles per element
constexpr int L3_PER_WARP = L3_STG + L3_SQ + L3_SACC + 2 * L3_TR;                // 23 KiB
constexpr size_t l3_smem(int warps) { return (size_t)L3_PER_WARP * warps * sizeof(double); }

// by-value operator block: lands in the constant bank, so the weights are immediate operands of the DFMAs
struct LineOps {
  double ds[16];   // Dsplit, column-major: ds[a + 4 b] = Dsplit[a, b]
  double factor_1, factor_2;
};

TB_D int l3_swz(int i, int j, int k) { return 16 * k + 4 * ((j + k) & 3) + ((i + k) & 3); }

// full-precision reciprocal: MUFU.RCP64H seed (relative error 2^-23) + one cubic step r0 (1 + e + e^2)
TB_D double l3_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// f2 >= 1e-4 decided on the high word (both branches of the means are accurate near the threshold; NaN -> true)
TB_D bool l3_is_rough(double f2) { return __double2hiint(f2) >= 0x3F1A36E2; }

// ---- Euler working variables q = (rho, v/2, p) and the Ranocha flux on them
TB_D void l3_to_q(const double* u, double gm1, double* q) {
  const double hr = 0.5 * l3_rcp(u[0]);
  double ke = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) { q[1 + d] = u[1 + d] * hr; ke = fma(u[1 + d], q[1 + d], ke); }
  q[0] = u[0];
  q[4] = gm1 * (u[4] - ke);
}

// flux_ranocha from the two logarithmic means (Trixi flux_ranocha), orientation O = 1, 2, 3
template <int O>
TB_D void l3_ranocha_from_means(const double* ql, const double* qr, double rho_mean, double inv_rho_p_mean,
                                double inv_gm1, double* f) {
  const double pl = ql[4], pr = qr[4];
  const double a1 = ql[1] + qr[1], a2 = ql[2] + qr[2], a3 = ql[3] + qr[3];   // arithmetic mean velocities
  const double hh = fma(ql[3], qr[3], fma(ql[2], qr[2], ql[1] * qr[1]));    // v_l . v_r / 4
  const double psum = pl + pr;
  const double pv = fma(pl, qr[O], pr * ql[O]);                             // (p_l v_r + p_r v_l) / 2
  const double ao = (O == 1) ? a1 : (O == 2 ? a2 : a3);
  const double f1 = rho_mean * ao;
  f[0] = f1;
  f[1] = (O == 1) ? fma(0.5, psum, f1 * a1) : f1 * a1;
  f[2] = (O == 2) ? fma(0.5, psum, f1 * a2) : f1 * a2;
  f[3] = (O == 3) ? fma(0.5, psum, f1 * a3) : f1 * a3;
  f[4] = fma(f1, fma(inv_rho_p_mean, inv_gm1, 2.0 * hh), pv);
}

// Smooth-branch means. ln_mean(rho_l, rho_r) = s/2 (1 - f2/3 - 4 f2^2/45 - 44 f2^3/945), f2 = ((x-y)/(x+y))^2 < 1e-4:
// the series of Trixi's (x+y)/(2 + f2(2/3 + f2(2/5 + 2 f2/7))), truncation 3e-18; one Newton step on the reciprocal
// suffices there because it only enters through f2. The second mean is Trixi's p_l p_r inv_ln_mean(rho_l p_r,
// rho_r p_l) with inv_ln_mean(x, y) = (2 + g2(2/3 + g2(2/5 + 2 g2/7)))/(x + y).
// Returns true if either mean needs its logarithmic branch.
TB_D bool l3_means_taylor(const double* ql, const double* qr, double& rho_mean, double& inv_rho_p_mean) {
  const double s = ql[0] + qr[0];
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  r = fma(r, fma(-s, r, 1.0), r);
  const double uu = (ql[0] - qr[0]) * r, f2 = uu * uu;
  rho_mean = s * fma(f2, fma(f2, fma(f2, -22.0 / 945, -2.0 / 45), -1.0 / 6), 0.5);
  const double x = ql[0] * qr[4], y = qr[0] * ql[4];
  const double rt = l3_rcp(x + y);
  const double ut = (x - y) * rt, g2 = ut * ut;
  inv_rho_p_mean = ((ql[4] * qr[4]) * rt) * fma(g2, fma(g2, fma(g2, 2.0 / 7, 2.0 / 5), 2.0 / 3), 2.0);
  return l3_is_rough(f2) || l3_is_rough(g2);
}
// exact means (Trixi's branches), used by the cold fix-up only
TB_D void l3_means_exact(double rl, double rr, double pl, double pr, double& rho_mean, double& inv_rho_p_mean) {
  rho_mean = ln_mean(rl, rr);
  inv_rho_p_mean = pl * pr * inv_ln_mean(rl * pr, rr * pl);
}

// two-point flux of one pair; returns the "needs exact means" flag (always false on the generic branch)
template <int O, int KIND>
TB_D bool l3_flux(int kind_rt, const double* ql, const double* qr, const EqPrm& p, double* f) {
  if (KIND == TRIXIB200_FLUX_RANOCHA) {
    double rm, im;
    const bool rough = l3_means_taylor(ql, qr, rm, im);
    l3_ranocha_from_means<O>(ql, qr, rm, im, p.inv_gm1, f);
    return rough;
  } else {
    // generic kinds work on (rho, v, p)
    double a[5] = {ql[0], 2 * ql[1], 2 * ql[2], 2 * ql[3], ql[4]}, b[5] = {qr[0], 2 * qr[1], 2 * qr[2], 2 * qr[3], qr[4]};
    EqEuler<3>::two_point_qf(kind_rt, a, b, O, p, f);
    return false;
  }
}

// ---- cold fix-ups. Everything crosses these calls by value so that the hot path's register arrays never have
// their address taken.
struct L3Vec5 { double v[5]; };
struct L3Nodes { double q[6][5]; };   // virtual nodes: low neighbour, line nodes 0..3, high neighbour
struct L3Acc { double a[4][5]; };
// the 8 pairs of a line: 6 volume pairs of the line nodes (virtual nodes 1..4) and the two surface pairs
// (low neighbour, node 0) and (node 3, high neighbour), each ordered (ll, rr)
__device__ constexpr int L3_PA[8] = {1, 1, 1, 2, 2, 3, 0, 4};
__device__ constexpr int L3_PB[8] = {2, 3, 4, 3, 4, 4, 1, 5};

// f_exact - f_taylor of one pair (orientation 1: the phase rotates the velocity components)
TB_D void l3_pair_correction(const double* ql, const double* qr, double inv_gm1, double* df) {
  double rm, im, re, ie, ft[5], fe[5];
  l3_means_taylor(ql, qr, rm, im);
  l3_ranocha_from_means<1>(ql, qr, rm, im, inv_gm1, ft);
  l3_means_exact(ql[0], qr[0], ql[4], qr[4], re, ie);
  l3_ranocha_from_means<1>(ql, qr, re, ie, inv_gm1, fe);
#pragma unroll
  for (int v = 0; v < 5; ++v) df[v] = fe[v] - ft[v];
}
// all flagged pairs of one line: returns the accumulator corrections
__device__ __noinline__ L3Acc l3_line_correction(L3Nodes nd, unsigned mask, LineOps ops, double inv_gm1) {
  L3Acc r;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int v = 0; v < 5; ++v) r.a[m][v] = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (mask & (1u << k)) {
      double df[5];
      l3_pair_correction(nd.q[L3_PA[k]], nd.q[L3_PB[k]], inv_gm1, df);
      if (k < 6) {
        const int x = L3_PA[k] - 1, y = L3_PB[k] - 1;
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          r.a[x][v] = fma(ops.ds[x + 4 * y], df[v], r.a[x][v]);
          r.a[y][v] = fma(ops.ds[y + 4 * x], df[v], r.a[y][v]);
        }
      } else if (k == 6) {
#pragma unroll
        for (int v = 0; v < 5; ++v) r.a[0][v] = fma(-ops.factor_1, df[v], r.a[0][v]);
      } else {
#pragma unroll
        for (int v = 0; v < 5; ++v) r.a[3][v] = fma(ops.factor_2, df[v], r.a[3][v]);
      }
    }
  return r;
}
// source terms of one node (reference dg_3d_kernel.jl:1821-1844), out of line: not on the benchmark path
__device__ __noinline__ L3Vec5 l3_source(const Dev* dp, int64_t e, int n, int i, int j, int k, double inv_jac, double t,
                                         const double* __restrict__ u) {
  const Dev& d = *dp;
  double x[3], un[5];
  L3Vec5 s;
  if (d.node_coords) {
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = d.node_coords[c + (size_t)3 * (n + (size_t)64 * e)];
  } else {
    const double jac = 1.0 / inv_jac;
    const int idx[3] = {i, j, k};
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __dadd_rn(d.centers[c + (size_t)3 * e], __dmul_rn(jac, d.ops->nodes[idx[c]]));
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) un[v] = u[(size_t)5 * (n + (size_t)64 * e) + v];
  EqEuler<3>::source(d.src, un, x, t, d.prm, s.v);
  return s;
}

// WARPS warps per CTA, 2 CTAs per SM = 8 warps/SM at up to 255 registers. Measured on B200 (level 7): 10 warps/SM
// at 200 registers and 12 at 168 (accumulator tile aliased onto the staging buffer to fit) spill and run 1.4-1.6x
// slower -- this kernel wants registers (8 independent flux chains per lane), not occupancy.
//
// The three phases run through ONE copy of the flux code (`dir` is a run-time value): phase `dir` reads the velocity
// / momentum rows of the shared tiles rotated so that slot 0 is the component normal to the lines, which makes every
// flux an orientation-1 flux, and only ~20 address computations per phase depend on `dir`. Unrolled phases give a
// 40-60 KB loop body, beyond the 32 KB L1.5 instruction cache: 15-16 % "no instruction" stalls
// (profiles/r1_ncu_line3d_v2_l6.json, r1_ncu_line3d_v4_l6.json). The hot path is also branch-free where ptxas would
// otherwise duplicate the flux code per path (given fluxes of boundary / mortar faces override the computed ones by
// selects; the last iteration prefetches its own element again).
template <int VFLUX, int SFLUX, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 2)
k_line3d(const __grid_constant__ Dev d, const __grid_constant__ LineOps ops, double* __restrict__ du,
         const double* __restrict__ u, double t, const int* __restrict__ elems, int64_t count) {
  constexpr int NV = 5, NQ = L3_NQ, NN = 64;
  constexpr bool FAST = (VFLUX == TRIXIB200_FLUX_RANOCHA && SFLUX == TRIXIB200_FLUX_RANOCHA);
  extern __shared__ __align__(16) double smem_l3[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int la = l16 & 3, lb = l16 >> 2;
  double* wbase = smem_l3 + (size_t)warp * L3_PER_WARP;
  double* stg = wbase + half * 320;                 // this element's AoS block (prefetch target)
  double* sq = wbase + L3_STG + half * (NQ * NN);   // [NQ][64] swizzled ...
  double* outt = sq;                                // ... and, after the z phase, the AoS tile of du
  double* sacc = wbase + L3_STG + L3_SQ + half * (NV * NN);   // swizzled accumulators [NV][64]
  double* tr = wbase + L3_STG + L3_SQ + L3_SACC + half * L3_TR;
  const EqPrm prm = d.prm;
  const double gm1 = prm.gamma - 1;
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int64_t npairs = (count + 1) >> 1;
  const int64_t wid = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;

  // element of this half-warp in pair `pr` (the odd tail duplicates the last element; its store is masked)
  auto elem_of = [&](int64_t pr, bool& valid) -> int64_t {
    int64_t s = 2 * pr + half;
    valid = s < count;
    if (!valid) s = count - 1;
    return elems ? (int64_t)elems[s] : s;
  };
  auto load_codes = [&](int64_t el, int* c) {
    const int2* p = reinterpret_cast<const int2*>(d.face_nbr + (size_t)el * 6);   // 24 B per element, 8 B aligned
    const int2 a = p[0], b = p[1], cc = p[2];
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = cc.x; c[5] = cc.y;
  };
  auto issue_block = [&](int64_t e) {
    const double* ue = u + (size_t)NV * NN * e;
#pragma unroll
    for (int m = 0; m < 10; ++m) cp_async16(stg + 2 * (l16 + 16 * m), ue + 2 * (l16 + 16 * m));
  };
  // offset (doubles) of a face's trace tile inside `tr`: x faces 2 x 96, then y and z faces 4 x 80
  auto face_off = [](int dir, int sd) { return dir == 0 ? 96 * sd : 32 + 160 * dir + 80 * sd; };
  // face traces of direction dir (faces 2 dir, 2 dir + 1) of element e with neighbour codes c0, c1, as 16-byte chunks
  // of the contiguous runs of the source: x faces of a local neighbour are 16 windows of 48 B (nodes i = 3 / i = 0),
  // y faces four runs of 160 B (j = 3 / j = 0), z faces one run of 640 B (k = 3 / k = 0); ready-made fluxes
  // (boundary / mortar faces) and halo traces are dense [f][v] runs of 640 B
  auto issue_traces = [&](auto dir_tag, int64_t e, const int c0, const int c1) {
    constexpr int dir = decltype(dir_tag)::value;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int code = sd == 0 ? c0 : c1;
      double* dst = tr + face_off(dir, sd);
      const double* src;
      if (code >= 0) src = u + (size_t)NV * NN * code;
      else if (code == NB_SFV) src = d.sfv + (size_t)NV * 16 * (2 * dir + sd + (size_t)6 * e);
      else src = d.halo_recv + (size_t)nb_halo_slot(code) * 16 * NV;
      const bool local = code >= 0;
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int c = l16 + 16 * it;
        int so = 2 * c, dof = 2 * c;
        bool on = c < 40;
        if (local) {
          if (dir == 0) {
            const int f = c / 3, w = c - 3 * f;
            so = 20 * f + (sd == 0 ? 14 : 0) + 2 * w; dof = 6 * f + 2 * w; on = true;
          } else if (dir == 1) {
            const int k = c / 10;
            so = 60 * k + (sd == 0 ? 60 : 0) + 2 * c;       // 80 k + 2 (c - 10 k)
          } else {
            so = (sd == 0 ? 240 : 0) + 2 * c;
          }
        }
        if (on) cp_async16(dst + dof, src + so);
      }
    }
  };

  using D0 = std::integral_constant<int, 0>;
  using D1 = std::integral_constant<int, 1>;
  using D2 = std::integral_constant<int, 2>;
  // swizzled tile positions of this lane's four line nodes per direction, one byte each
  unsigned pk0 = 0, pk1 = 0, pk2 = 0;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    pk0 |= (unsigned)l3_swz(m, la, lb) << (8 * m);
    pk1 |= (unsigned)l3_swz(la, m, lb) << (8 * m);
    pk2 |= (unsigned)l3_swz(la, lb, m) << (8 * m);
  }
  int code[6] = {NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV, NB_SFV};
  int64_t pr = wid;
  bool valid = false;
  int64_t e = 0;
  // cp.async groups are committed in the order  block, x-traces, y-traces, z-traces  (of the NEXT pair), each right
  // after its buffer was consumed, so whenever one of them is needed exactly three younger groups may be pending.
  if (pr < npairs) {
    e = elem_of(pr, valid);
    load_codes(e, code);
    issue_block(e); cp_async_commit();
    issue_traces(D0{}, e, code[0], code[1]); cp_async_commit();
    issue_traces(D1{}, e, code[2], code[3]); cp_async_commit();
    issue_traces(D2{}, e, code[4], code[5]); cp_async_commit();
  }

  for (; pr < npairs; pr += nw) {
    // the last iteration prefetches its own element again instead of branching around the copies
    const int64_t pr_next = (pr + nw < npairs) ? pr + nw : pr;
    bool valid_next = false;
    const int64_t e_next = elem_of(pr_next, valid_next);
    int code_next[6];
    load_codes(e_next, code_next);          // issued early, consumed when the trace copies are issued
    const double inv_jac = d.inv_jac[e];

    // ---- block has landed: cons -> q, z-line ownership (conflict-free AoS reads)
    cp_async_wait<3>();
    __syncwarp();
    {
      double un[4][NV], qn[4][NQ];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int v = 0; v < NV; ++v) un[m][v] = stg[NV * (l16 + 16 * m) + v];
#pragma unroll
      for (int m = 0; m < 4; ++m) l3_to_q(un[m], gm1, qn[m]);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int pos = l3_swz(la, lb, m);
#pragma unroll
        for (int v = 0; v < NQ; ++v) sq[v * NN + pos] = qn[m][v];
      }
    }
    __syncwarp();
    issue_block(e_next);
    cp_async_commit();

    double acc[4][NV];
#pragma unroll 1
    for (int dir = 0; dir < 3; ++dir) {
      // rows of the velocity / momentum components in slot order (slot 0 = normal component)
      const int c0 = 1 + dir, c1 = (dir == 2) ? 1 : dir + 2, c2 = (dir == 0) ? 3 : dir;
      const int r0 = c0 * NN, r1 = c1 * NN, r2 = c2 * NN;
      const int c_lo = dir == 0 ? code[0] : (dir == 1 ? code[2] : code[4]);
      const int c_hi = dir == 0 ? code[1] : (dir == 1 ? code[3] : code[5]);
      const unsigned pk = dir == 0 ? pk0 : (dir == 1 ? pk1 : pk2);
      const int pos[4] = {(int)(pk & 0xff), (int)((pk >> 8) & 0xff), (int)((pk >> 16) & 0xff), (int)(pk >> 24)};
      // ---- this direction's traces have landed (three younger groups may still be in flight)
      cp_async_wait<3>();
      __syncwarp();
      double Q[6][NQ];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        Q[1 + m][0] = sq[pos[m]];
        Q[1 + m][1] = sq[r0 + pos[m]];
        Q[1 + m][2] = sq[r1 + pos[m]];
        Q[1 + m][3] = sq[r2 + pos[m]];
        Q[1 + m][4] = sq[4 * NN + pos[m]];
      }
      double nbv[2][NV];
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        const int cd = sd == 0 ? c_lo : c_hi;
        const bool win = dir == 0 && cd >= 0;      // 48-byte windows: stride 6, data at +1 on the low face
        const double* src = tr + face_off(dir, sd) + l16 * (win ? 6 : 5) + ((win && sd == 0) ? 1 : 0);
        nbv[sd][0] = src[0]; nbv[sd][1] = src[c0]; nbv[sd][2] = src[c1]; nbv[sd][3] = src[c2]; nbv[sd][4] = src[4];
      }
      if (dir == 0) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[m][v] = 0;
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          acc[m][0] = sacc[pos[m]];
          acc[m][1] = sacc[r0 + pos[m]];
          acc[m][2] = sacc[r1 + pos[m]];
          acc[m][3] = sacc[r2 + pos[m]];
          acc[m][4] = sacc[4 * NN + pos[m]];
        }
      }
      l3_to_q(nbv[0], gm1, Q[0]);
      l3_to_q(nbv[1], gm1, Q[5]);
      const bool sfv_lo = c_lo == NB_SFV, sfv_hi = c_hi == NB_SFV;

      // ---- 6 volume + 2 surface fluxes, all orientation 1 in the rotated frame, evaluated stage by stage across
      // the 8 pairs so that consecutive instructions are independent (reference dg_3d_kernel.jl:188-257 evaluates
      // 12 volume fluxes per node, and the interface fluxes in two more kernels)
      double F[8][NV];
      unsigned rough = 0;
      if (FAST) {
        double rm[8], im[8];
        {
          double s[8], r[8], dd[8], tt[8], rt[8], xy[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double* a = Q[L3_PA[k]]; const double* b = Q[L3_PB[k]];
            s[k] = a[0] + b[0];
            dd[k] = a[0] - b[0];
            const double x = a[0] * b[4], y = b[0] * a[4];
            tt[k] = x + y;
            xy[k] = x - y;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(s[k]));
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rt[k]) : "d"(tt[k]));
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            r[k] = fma(r[k], fma(-s[k], r[k], 1.0), r[k]);
            const double e1 = fma(-tt[k], rt[k], 1.0);
            rt[k] = fma(rt[k], fma(e1, e1, e1), rt[k]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double uu = dd[k] * r[k], f2 = uu * uu;
            const double ut = xy[k] * rt[k], g2 = ut * ut;
            rm[k] = s[k] * fma(f2, fma(f2, fma(f2, -22.0 / 945, -2.0 / 45), -1.0 / 6), 0.5);
            im[k] = ((Q[L3_PA[k]][4] * Q[L3_PB[k]][4]) * rt[k]) * fma(g2, fma(g2, fma(g2, 2.0 / 7, 2.0 / 5), 2.0 / 3), 2.0);
            if (l3_is_rough(f2) || l3_is_rough(g2)) rough |= 1u << k;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          l3_ranocha_from_means<1>(Q[L3_PA[k]], Q[L3_PB[k]], rm[k], im[k], prm.inv_gm1, F[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          l3_flux<1, -1>(k < 6 ? vflux : sflux, Q[L3_PA[k]], Q[L3_PB[k]], prm, F[k]);
      }
      // faces whose flux is given (boundary / mortar): the trace IS the flux
      rough &= ~((sfv_lo ? 1u << 6 : 0u) | (sfv_hi ? 1u << 7 : 0u));
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        F[6][v] = sfv_lo ? nbv[0][v] : F[6][v];
        F[7][v] = sfv_hi ? nbv[1][v] : F[7][v];
      }
      // ---- accumulate: flux differencing weights, then the surface integral (reference dg_3d_kernel.jl:1787-1794)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int x = L3_PA[k] - 1, y = L3_PB[k] - 1;
        const double wxy = ops.ds[x + 4 * y], wyx = ops.ds[y + 4 * x];
#pragma unroll
        for (int v = 0; v < NV; ++v) { acc[x][v] = fma(wxy, F[k][v], acc[x][v]); acc[y][v] = fma(wyx, F[k][v], acc[y][v]); }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        acc[0][v] = fma(-ops.factor_1, F[6][v], acc[0][v]);
        acc[3][v] = fma(ops.factor_2, F[7][v], acc[3][v]);
      }
      if (FAST && rough != 0) {
        L3Nodes nd;
#pragma unroll
        for (int m = 0; m < 6; ++m)
#pragma unroll
          for (int v = 0; v < NQ; ++v) nd.q[m][v] = Q[m][v];
        const L3Acc r = l3_line_correction(nd, rough, ops, prm.inv_gm1);
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[m][v] += r.a[m][v];
      }
      if (dir < 2) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          sacc[pos[m]] = acc[m][0];
          sacc[r0 + pos[m]] = acc[m][1];
          sacc[r1 + pos[m]] = acc[m][2];
          sacc[r2 + pos[m]] = acc[m][3];
          sacc[4 * NN + pos[m]] = acc[m][4];
        }
      }
      __syncwarp();   // traces consumed, accumulators visible; after the z phase: every lane is done with sq
      if (dir == 0) issue_traces(D0{}, e_next, code_next[0], code_next[1]);
      else if (dir == 1) issue_traces(D1{}, e_next, code_next[2], code_next[3]);
      else issue_traces(D2{}, e_next, code_next[4], code_next[5]);
      cp_async_commit();
    }
    // ---- Jacobian, sources, output. After the z phase the slots hold the components (z, x, y).
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[m][v] *= -inv_jac;
      if (d.src != TRIXIB200_SRC_NONE) {
        const L3Vec5 sv = l3_source(&d, e, l16 + 16 * m, la, lb, m, inv_jac, t, u);
        acc[m][0] += sv.v[0]; acc[m][1] += sv.v[3]; acc[m][2] += sv.v[1]; acc[m][3] += sv.v[2]; acc[m][4] += sv.v[4];
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double* o = outt + NV * (l16 + 16 * m);
      o[0] = acc[m][0]; o[3] = acc[m][1]; o[1] = acc[m][2]; o[2] = acc[m][3]; o[4] = acc[m][4];
    }
    __syncwarp();
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

template <int VFLUX, int SFLUX, int WARPS>
static int line3d_launch_t(const Dev& d, const LineOps& ops, double* du, const double* u, double t, const int* elems,
                           int64_t count, cudaStream_t stream, int sm_count) {
  auto kern = k_line3d<VFLUX, SFLUX, WARPS>;
  static DeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3_smem(WARPS)) != cudaSuccess)
      { configured.undo(); return TRIXIB200_ECUDA; }
  }
  if (count <= 0) return 0;
  const int64_t npairs = (count + 1) / 2;
  const int64_t want = (npairs + WARPS - 1) / WARPS;
  const unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * 2);
  kern<<<blocks, 32 * WARPS, l3_smem(WARPS), stream>>>(d, ops, du, u, t, elems, count);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

static int line3d_launch(const trixib200_config& c, const Dev& d, const LineOps& ops, double* du, const double* u,
                         double t, const int* elems, int64_t count, cudaStream_t s, int sm_count) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  if (c.volume_flux == R && c.surface_flux == R) return line3d_launch_t<R, R, 4>(d, ops, du, u, t, elems, count, s, sm_count);
  if (c.volume_flux == R) return line3d_launch_t<R, -1, 4>(d, ops, du, u, t, elems, count, s, sm_count);
  return line3d_launch_t<-1, -1, 4>(d, ops, du, u, t, elems, count, s, sm_count);
}

}  // namespace tb
