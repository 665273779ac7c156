// Line-owner fused rhs! kernel, generation 6: the same ownership as k_line3d (a half-warp owns an element, a lane owns
// a 4-node line per phase; kernels_line3d.cuh) rebuilt around occupancy. ncu on k_line3d (profiles/
// r1_ncu_line3d_v5_l6.json) showed the FP64 pipe 50 % busy with 2 warps per scheduler: one warp's instruction stream
// needs ~5200 issue cycles per element pair of which ~2700 are FP64 issue, and two warps do not cover each other's
// shared-memory / MUFU / dependency waits. This kernel runs 3 warps per scheduler (12 per SM, 3 CTAs x 4 warps):
//  * <= 168 registers: the 8 pair fluxes of a line are evaluated as two batches of 4 -- batch A = {(lo,0), (0,1),
//    (0,2), (0,3)} completes node 0, batch B = {(1,2), (1,3), (2,3), (3,hi)} completes nodes 1..3 -- and the running
//    sums of earlier phases are added at the END of a phase (read-modify-write of the shared accumulator tile), so at
//    most 4 flux chains + 5 nodes + 3 partial nodes are live at a time;
//  * 9 KiB of shared memory per element instead of 11.5: phases run z, y, x. The z-line owner of the first phase is
//    also the conflict-free reader of the AoS block (stride-5 doubles), so the block is converted to the swizzled SoA
//    tile IN PLACE and its q never has to be re-read for the first phase; the x-line owner of the last phase holds 20
//    CONTIGUOUS doubles of du (nodes (0..3, j, k) x 5 variables), which go straight to global memory as 16-byte stores:
//    no output tile. The next element's block is prefetched into the q tile as soon as the last phase has read it;
//  * the swizzle is an XOR, p(i,j,k) = 16 k + 4 (j ^ k) + (i ^ k): conflict-free for x-, y- and z-line access like the
//    additive one, and the four positions of a line are P ^ (m * stride), stride = 1, 4, 21 -- three LOP3 per phase.
// Everything else (cp.async trace windows, rotated component slots so that one copy of the flux code serves all three
// directions, Taylor-branch means with the logarithmic branch as an out-of-line correction) is as in k_line3d.
// Reference stages covered: src/solvers/dg_3d_kernel.jl:188-257 (volume_flux_integral_kernel!), :1121-1152
// (prolong_interfaces_kernel!), :1155-1225 (surface_flux_kernel!, interface_flux_kernel!), :1773-1799
// (surface_integral_kernel!), :1802-1818 (jacobian_kernel!), :1821-1844 (source_terms_kernel!).
#pragma once
#include "kernels_line3d.cuh"

namespace tb {

constexpr int L6_TILE = 320;                          // 5 x 64 doubles
constexpr int L6_TR = L3_TR;                          // 512: x faces 2 x 96, y and z faces 4 x 80
constexpr int L6_EL = 2 * L6_TILE + L6_TR;            // 1152 doubles = 9 KiB per element
constexpr size_t l6_smem(int warps) { return (size_t)2 * L6_EL * warps * sizeof(double); }

struct L6Pair { double a[5], b[5]; };
// f_exact - f_taylor of one flagged pair (cold)
__device__ __noinline__ L3Vec5 l6_pair_correction(L6Pair p, double inv_gm1) {
  L3Vec5 r;
  l3_pair_correction(p.a, p.b, inv_gm1, r.v);
  return r;
}

// Four two-point fluxes (orientation 1 in the rotated frame) between the virtual line nodes (A_k, B_k) of Q, staged
// across the pairs so that consecutive instructions are independent. Returns the "needs the logarithmic branch" mask.
template <bool FAST, int A0, int B0, int A1, int B1, int A2, int B2, int A3, int B3>
TB_D unsigned l6_flux4(const double (&Q)[6][5], int kind_a, int kind_b, int kind_c, int kind_d, const EqPrm& prm,
                       double (&F)[4][5]) {
  constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
  unsigned rough = 0;
  if (FAST) {
    double s[4], r[4], dd[4], tt[4], rt[4], xy[4], rm[4], im[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double* a = Q[PA[k]]; const double* b = Q[PB[k]];
      s[k] = a[0] + b[0];
      dd[k] = a[0] - b[0];
      const double x = a[0] * b[4], y = b[0] * a[4];
      tt[k] = x + y;
      xy[k] = x - y;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(s[k]));
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rt[k]) : "d"(tt[k]));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[k] = fma(r[k], fma(-s[k], r[k], 1.0), r[k]);
      const double e1 = fma(-tt[k], rt[k], 1.0);
      rt[k] = fma(rt[k], fma(e1, e1, e1), rt[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double uu = dd[k] * r[k], f2 = uu * uu;
      const double ut = xy[k] * rt[k], g2 = ut * ut;
      rm[k] = s[k] * fma(f2, fma(f2, fma(f2, -22.0 / 945, -2.0 / 45), -1.0 / 6), 0.5);
      im[k] = ((Q[PA[k]][4] * Q[PB[k]][4]) * rt[k]) * fma(g2, fma(g2, fma(g2, 2.0 / 7, 2.0 / 5), 2.0 / 3), 2.0);
      if (l3_is_rough(f2) || l3_is_rough(g2)) rough |= 1u << k;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) l3_ranocha_from_means<1>(Q[PA[k]], Q[PB[k]], rm[k], im[k], prm.inv_gm1, F[k]);
  } else {
    const int kind[4] = {kind_a, kind_b, kind_c, kind_d};
#pragma unroll
    for (int k = 0; k < 4; ++k) l3_flux<1, -1>(kind[k], Q[PA[k]], Q[PB[k]], prm, F[k]);
  }
  return rough;
}

template <int VFLUX, int SFLUX, bool SFV, int WARPS, int CTAS>
__global__ void __launch_bounds__(32 * WARPS, CTAS)
k_line6(const __grid_constant__ Dev d, const __grid_constant__ LineOps ops, double* __restrict__ du,
        const double* __restrict__ u, double t, const int* __restrict__ elems, int64_t count) {
  constexpr int NV = 5, NN = 64;
  constexpr bool FAST = (VFLUX == TRIXIB200_FLUX_RANOCHA && SFLUX == TRIXIB200_FLUX_RANOCHA);
  extern __shared__ __align__(16) double smem_l6[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int la = l16 & 3, lb = l16 >> 2;
  double* ebase = smem_l6 + (size_t)(2 * warp + half) * L6_EL;
  double* sq = ebase;                    // AoS landing zone of the element block, then the swizzled SoA q [5][64]
  double* sacc = ebase + L6_TILE;        // swizzled SoA running sums [5][64], handed from phase to phase
  double* tr = ebase + 2 * L6_TILE;      // face traces of the six neighbours
  const EqPrm prm = d.prm;
  const double gm1 = prm.gamma - 1;
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int64_t npairs = (count + 1) >> 1;
  const int64_t wid = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;

  auto elem_of = [&](int64_t pr, bool& valid) -> int {
    int64_t s = 2 * pr + half;
    valid = s < count;
    if (!valid) s = count - 1;
    return elems ? elems[s] : (int)s;
  };
  // neighbour codes of the two faces of direction dir (24 B per element, 8 B aligned)
  auto load_codes = [&](int el, int dir) -> int2 {
    return reinterpret_cast<const int2*>(d.face_nbr + (size_t)el * 6)[dir];
  };
  // what the phases need to know about the faces of the CURRENT element: 4 bits per direction
  // (low face local, low face given flux, high face local, high face given flux)
  auto face_bits = [](int2 c, int dir) -> unsigned {
    return ((c.x >= 0 ? 1u : 0u) | (c.x == NB_SFV ? 2u : 0u) | (c.y >= 0 ? 4u : 0u) | (c.y == NB_SFV ? 8u : 0u)) << (4 * dir);
  };
  auto issue_block = [&](int e) {
    const double* ue = u + (size_t)NV * NN * e;
#pragma unroll
    for (int m = 0; m < 10; ++m) cp_async16(sq + 2 * (l16 + 16 * m), ue + 2 * (l16 + 16 * m));
  };
  auto face_off = [](int dir, int sd) { return dir == 0 ? 96 * sd : 32 + 160 * dir + 80 * sd; };
  // trace windows exactly as in k_line3d (x faces 16 windows of 48 B, y faces four runs of 160 B, z faces one run of
  // 640 B of the neighbour's block; given fluxes and halo traces are dense [f][v] runs of 640 B)
  auto issue_traces = [&](auto dir_tag, int e, const int c0, const int c1) {
    constexpr int dir = decltype(dir_tag)::value;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      const int code = sd == 0 ? c0 : c1;
      double* dst = tr + face_off(dir, sd);
      const double* src;
      if (code >= 0) src = u + (size_t)NV * NN * code;
      else if (SFV && code == NB_SFV) src = d.sfv + (size_t)NV * 16 * (2 * dir + sd + (size_t)6 * e);
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
            so = 60 * k + (sd == 0 ? 60 : 0) + 2 * c;
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

  // tile positions of this lane's line nodes: P ^ (m * stride) with stride 1 (x), 4 (y), 21 (z)
  const int Px = 16 * lb + 4 * (la ^ lb) + lb, Py = 20 * lb + (la ^ lb), Pz = l16;

  unsigned fbits = 0;
  int64_t pr = wid;
  bool valid = false;
  int e = 0;
  // cp.async groups are committed in the order  z-traces, y-traces, block, x-traces  (of the NEXT pair): at the top
  // of an iteration only the x-traces may be pending (wait_group 1), at the x phase the z- and y-traces of the next
  // pair (wait_group 2)
  if (pr < npairs) {
    e = elem_of(pr, valid);
    const int2 cx = load_codes(e, 0), cy = load_codes(e, 1), cz = load_codes(e, 2);
    fbits = face_bits(cx, 0) | face_bits(cy, 1) | face_bits(cz, 2);
    issue_traces(D2{}, e, cz.x, cz.y); cp_async_commit();
    issue_traces(D1{}, e, cy.x, cy.y); cp_async_commit();
    issue_block(e); cp_async_commit();
    issue_traces(D0{}, e, cx.x, cx.y); cp_async_commit();
  }

  for (; pr < npairs; pr += nw) {
    const int64_t pr_next = (pr + nw < npairs) ? pr + nw : pr;   // the last iteration prefetches its own element again
    bool valid_next = false;
    const int e_next = elem_of(pr_next, valid_next);
    unsigned fbits_next = 0;
    const double inv_jac = d.inv_jac[e];

    // ---- the block has landed: cons -> q in place. Lane (la, lb) reads the nodes (la, lb, m) = its z-line.
    cp_async_wait<1>();
    __syncwarp();
    double Q[6][NV];   // virtual line nodes: low neighbour, nodes 0..3, high neighbour; slots (rho, vn/2, vt1/2, vt2/2, p)
    {
      double un[4][NV];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int v = 0; v < NV; ++v) un[m][v] = sq[NV * (l16 + 16 * m) + v];
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double q[NV];
        l3_to_q(un[m], gm1, q);
        const int pos = Pz ^ (21 * m);
#pragma unroll
        for (int v = 0; v < NV; ++v) sq[v * NN + pos] = q[v];
        // z phase first: slot 1 = v3/2, then (v1/2, v2/2)
        Q[1 + m][0] = q[0]; Q[1 + m][1] = q[3]; Q[1 + m][2] = q[1]; Q[1 + m][3] = q[2]; Q[1 + m][4] = q[4];
      }
    }
    __syncwarp();

#pragma unroll 1
    for (int step = 0; step < 3; ++step) {
      const int dir = 2 - step;
      // rows of the velocity / momentum components in slot order (slot 1 = normal component)
      const int c0 = 1 + dir, c1 = (dir == 2) ? 1 : dir + 2, c2 = (dir == 0) ? 3 : dir;
      const int r0 = c0 * NN, r1 = c1 * NN, r2 = c2 * NN;
      const int2 cn = load_codes(e_next, dir);   // issued early, consumed when the trace copies are issued
      const unsigned fb = fbits >> (4 * dir);
      const int P = dir == 0 ? Px : (dir == 1 ? Py : Pz);
      const int stp = dir == 0 ? 1 : (dir == 1 ? 4 : 21);
      const int pos[4] = {P, P ^ stp, P ^ (2 * stp), P ^ (3 * stp)};
      if (step == 2) {
        cp_async_wait<2>();
        __syncwarp();
      }
      if (step > 0) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          Q[1 + m][0] = sq[pos[m]];
          Q[1 + m][1] = sq[r0 + pos[m]];
          Q[1 + m][2] = sq[r1 + pos[m]];
          Q[1 + m][3] = sq[r2 + pos[m]];
          Q[1 + m][4] = sq[4 * NN + pos[m]];
        }
      }
      double nbv[2][NV];
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        const bool win = dir == 0 && (fb & (sd == 0 ? 1u : 4u));      // 48-byte windows: stride 6, data at +1 on the low face
        const double* src = tr + face_off(dir, sd) + l16 * (win ? 6 : 5) + ((win && sd == 0) ? 1 : 0);
        nbv[sd][0] = src[0]; nbv[sd][1] = src[c0]; nbv[sd][2] = src[c1]; nbv[sd][3] = src[c2]; nbv[sd][4] = src[4];
      }
      if (step == 2) {
        // every lane has read its q: the tile becomes the landing zone of the next block
        __syncwarp();
        issue_block(e_next);
        cp_async_commit();
      }
      l3_to_q(nbv[0], gm1, Q[0]);
      l3_to_q(nbv[1], gm1, Q[5]);
      const bool sfv_lo = SFV && (fb & 2u), sfv_hi = SFV && (fb & 8u);

      // a finished node: step 0 starts the running sums, step 1 adds to them, step 2 adds, scales and keeps the result
      auto hand_over = [&](int m, double* a) {
        double* p = sacc + pos[m];
        if (step == 0) {
          p[0] = a[0]; p[r0] = a[1]; p[r1] = a[2]; p[r2] = a[3]; p[4 * NN] = a[4];
        } else if (step == 1) {
          p[0] += a[0]; p[r0] += a[1]; p[r1] += a[2]; p[r2] += a[3]; p[4 * NN] += a[4];
        } else {
          // x phase: slots are (x, y, z) = the natural component order
          a[0] = (a[0] + p[0]) * -inv_jac;
          a[1] = (a[1] + p[1 * NN]) * -inv_jac;
          a[2] = (a[2] + p[2 * NN]) * -inv_jac;
          a[3] = (a[3] + p[3 * NN]) * -inv_jac;
          a[4] = (a[4] + p[4 * NN]) * -inv_jac;
          if (d.src != TRIXIB200_SRC_NONE) {
            const L3Vec5 sv = l3_source(&d, e, m + 4 * l16, m, la, lb, inv_jac, t, u);
#pragma unroll
            for (int v = 0; v < NV; ++v) a[v] += sv.v[v];
          }
        }
      };

      // ---- batch A: (lo,0), (0,1), (0,2), (0,3)  (reference dg_3d_kernel.jl:188-257 evaluates 12 volume fluxes per
      // node, and the interface fluxes in two more kernels)
      double acc0[NV], acc1[NV], acc2[NV], acc3[NV];
      {
        double F[4][NV];
        unsigned rough = l6_flux4<FAST, 0, 1, 1, 2, 1, 3, 1, 4>(Q, sflux, vflux, vflux, vflux, prm, F);
        if (SFV) {
          if (sfv_lo) rough &= ~1u;
#pragma unroll
          for (int v = 0; v < NV; ++v) F[0][v] = sfv_lo ? nbv[0][v] : F[0][v];
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc0[v] = fma(ops.ds[0 + 4 * 3], F[3][v], fma(ops.ds[0 + 4 * 2], F[2][v], fma(ops.ds[0 + 4 * 1], F[1][v], -ops.factor_1 * F[0][v])));
          acc1[v] = ops.ds[1 + 4 * 0] * F[1][v];
          acc2[v] = ops.ds[2 + 4 * 0] * F[2][v];
          acc3[v] = ops.ds[3 + 4 * 0] * F[3][v];
        }
        if (FAST && rough != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (rough & (1u << k)) {
              L6Pair pp;
#pragma unroll
              for (int v = 0; v < NV; ++v) { pp.a[v] = Q[k == 0 ? 0 : 1][v]; pp.b[v] = Q[k == 0 ? 1 : 1 + k][v]; }
              const L3Vec5 df = l6_pair_correction(pp, prm.inv_gm1);
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                if (k == 0) acc0[v] = fma(-ops.factor_1, df.v[v], acc0[v]);
                else acc0[v] = fma(ops.ds[0 + 4 * k], df.v[v], acc0[v]);
                if (k == 1) acc1[v] = fma(ops.ds[1], df.v[v], acc1[v]);
                if (k == 2) acc2[v] = fma(ops.ds[2], df.v[v], acc2[v]);
                if (k == 3) acc3[v] = fma(ops.ds[3], df.v[v], acc3[v]);
              }
            }
        }
      }
      hand_over(0, acc0);
      if (step == 2 && valid) {
        double2* o = reinterpret_cast<double2*>(du + (size_t)NV * NN * e + 20 * l16);
        o[0] = make_double2(acc0[0], acc0[1]);
        o[1] = make_double2(acc0[2], acc0[3]);
      }
      // ---- batch B: (1,2), (1,3), (2,3), (3,hi)
      {
        double F[4][NV];
        unsigned rough = l6_flux4<FAST, 2, 3, 2, 4, 3, 4, 4, 5>(Q, vflux, vflux, vflux, sflux, prm, F);
        if (SFV) {
          if (sfv_hi) rough &= ~8u;
#pragma unroll
          for (int v = 0; v < NV; ++v) F[3][v] = sfv_hi ? nbv[1][v] : F[3][v];
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc1[v] = fma(ops.ds[1 + 4 * 3], F[1][v], fma(ops.ds[1 + 4 * 2], F[0][v], acc1[v]));
          acc2[v] = fma(ops.ds[2 + 4 * 3], F[2][v], fma(ops.ds[2 + 4 * 1], F[0][v], acc2[v]));
          acc3[v] = fma(ops.factor_2, F[3][v], fma(ops.ds[3 + 4 * 2], F[2][v], fma(ops.ds[3 + 4 * 1], F[1][v], acc3[v])));
        }
        if (FAST && rough != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (rough & (1u << k)) {
              constexpr int XA[4] = {2, 2, 3, 4}, XB[4] = {3, 4, 4, 5};
              L6Pair pp;
#pragma unroll
              for (int v = 0; v < NV; ++v) { pp.a[v] = Q[XA[k]][v]; pp.b[v] = Q[XB[k]][v]; }
              const L3Vec5 df = l6_pair_correction(pp, prm.inv_gm1);
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                if (k == 0) { acc1[v] = fma(ops.ds[1 + 4 * 2], df.v[v], acc1[v]); acc2[v] = fma(ops.ds[2 + 4 * 1], df.v[v], acc2[v]); }
                if (k == 1) { acc1[v] = fma(ops.ds[1 + 4 * 3], df.v[v], acc1[v]); acc3[v] = fma(ops.ds[3 + 4 * 1], df.v[v], acc3[v]); }
                if (k == 2) { acc2[v] = fma(ops.ds[2 + 4 * 3], df.v[v], acc2[v]); acc3[v] = fma(ops.ds[3 + 4 * 2], df.v[v], acc3[v]); }
                if (k == 3) acc3[v] = fma(ops.factor_2, df.v[v], acc3[v]);
              }
            }
        }
      }
      hand_over(1, acc1);
      hand_over(2, acc2);
      hand_over(3, acc3);
      if (step == 2 && valid) {
        double2* o = reinterpret_cast<double2*>(du + (size_t)NV * NN * e + 20 * l16);
        o[2] = make_double2(acc0[4], acc1[0]);
        o[3] = make_double2(acc1[1], acc1[2]);
        o[4] = make_double2(acc1[3], acc1[4]);
        o[5] = make_double2(acc2[0], acc2[1]);
        o[6] = make_double2(acc2[2], acc2[3]);
        o[7] = make_double2(acc2[4], acc3[0]);
        o[8] = make_double2(acc3[1], acc3[2]);
        o[9] = make_double2(acc3[3], acc3[4]);
      }
      __syncwarp();   // traces consumed, running sums visible
      if (dir == 0) issue_traces(D0{}, e_next, cn.x, cn.y);
      else if (dir == 1) issue_traces(D1{}, e_next, cn.x, cn.y);
      else issue_traces(D2{}, e_next, cn.x, cn.y);
      cp_async_commit();
      fbits_next |= face_bits(cn, dir);
    }
    e = e_next; valid = valid_next; fbits = fbits_next;
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------- host side
template <int VFLUX, int SFLUX, bool SFV, int CTAS = 3>
static int line6_launch_t(const Dev& d, const LineOps& ops, double* du, const double* u, double t, const int* elems,
                          int64_t count, cudaStream_t stream, int sm_count) {
  constexpr int WARPS = 4;
  auto kern = k_line6<VFLUX, SFLUX, SFV, WARPS, CTAS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l6_smem(WARPS)) != cudaSuccess)
      return TRIXIB200_ECUDA;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  if (count <= 0) return 0;
  const int64_t npairs = (count + 1) / 2;
  const int64_t want = (npairs + WARPS - 1) / WARPS;
  const unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)sm_count * CTAS);
  kern<<<blocks, 32 * WARPS, l6_smem(WARPS), stream>>>(d, ops, du, u, t, elems, count);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

static int line6_launch(const trixib200_config& c, const Dev& d, const LineOps& ops, double* du, const double* u,
                        double t, const int* elems, int64_t count, cudaStream_t s, int sm_count) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  const bool sfv = d.B > 0 || d.M > 0;   // faces whose flux is given in surface_flux_values (boundaries, mortars)
  // TRIXIB200_LINE_CTAS=2: the same code at 2 CTAs (8 warps) per SM and up to 255 registers (A/B measurements)
  static const bool two = getenv("TRIXIB200_LINE_CTAS") && atoi(getenv("TRIXIB200_LINE_CTAS")) == 2;
  if (two && !sfv && c.volume_flux == R && c.surface_flux == R)
    return line6_launch_t<R, R, false, 2>(d, ops, du, u, t, elems, count, s, sm_count);
  if (c.volume_flux == R && c.surface_flux == R)
    return sfv ? line6_launch_t<R, R, true>(d, ops, du, u, t, elems, count, s, sm_count)
               : line6_launch_t<R, R, false>(d, ops, du, u, t, elems, count, s, sm_count);
  return line6_launch_t<-1, -1, true>(d, ops, du, u, t, elems, count, s, sm_count);
}

}  // namespace tb
