// Shared pieces of the line-owner kernels for 3D compressible Euler flux differencing at polydeg 3: LineOps, the
// per-node working set (rho, v/2, p), flux_ranocha from the two logarithmic means with the Taylor-branch series, the
// out-of-line corrections for pairs that need the logarithmic branch, source terms of one node, `line3d_available`.
// The kernel that used to live here (k_line3d, generation 5 of the design: 4.85 ms per level-7 rhs!, A/B record in
// profiles/r1_ncu_line3d_v{1..5}_l6.json and r1_bench_n1_v5.json) was superseded by k_line6 (kernels_line6.cuh,
// 4.66 ms) and removed from the build in round 2.
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

// ---------------------------------------------------------------------------------------------- host side
inline bool line3d_available(const trixib200_config& c) {
  return c.ndim == 3 && c.polydeg == 3 && c.volume_integral == TRIXIB200_VI_FLUX_DIFFERENCING && !c.nonconservative &&
         c.equations == TRIXIB200_EQ_EULER;
}

}  // namespace tb
