// Fused rhs! kernel for polydeg 3 (N = 4) in 2D/3D: ONE launch does volume integral (weak form / flux
// differencing / shock-capturing blend) + conforming-interface fluxes + surface integral + Jacobian + sources.
// `u` is read once per element (plus neighbour face traces, normally L2 hits), `du` is written once; the
// reference's interfaces.u / surface_flux_arr / surface_flux_values intermediates are never materialised
// (reference src/solvers/dg_3d.jl:895-925 runs 6+ launches and 3 read-modify-write sweeps of du).
//
// Thread mapping: one thread per node, 256 threads per CTA (4 elements in 3D, 16 in 2D).
//  * An element's nv*N^d block is one contiguous run of u (AoS at the boundary) -> coalesced load, converted
//    once per node to primitive variables kept SoA in shared memory.
//  * Flux differencing evaluates every symmetric node pair ONCE: the 6 pairs of a 4-node line are the perfect
//    matchings r0 = {01,23}, r1 = {02,13}, r2 = {03,12}. Slot d (per direction) packs r0+r2: lanes with even
//    line coordinate compute their r0 pair, odd ones their r2 pair, and each lane receives the other through a
//    double-buffered shared exchange tile. The r1 matchings of x and y share one slot (selector
//    bit1(ix)^bit1(iy)); in 3D the z r1 matching occupies warp 0 of the element while warp 1 computes the 32
//    z-face surface fluxes in the same slot. Result: 4.5 volume + 1.5 surface flux evaluations per node with
//    no idle lanes (the reference evaluates 12 + 0.75 and re-reads u from global memory each time).
//  * Interface fluxes are evaluated by both neighbours from identical (ll, rr) inputs, so they agree bitwise.
//  * Boundary / mortar faces read surface_flux_values written by the staged kernels; halo faces read the
//    trace received from the peer rank.
#pragma once
#include "device.cuh"

namespace tb {

constexpr int FUSED_THREADS = 256;

template <class Eq, int VI> struct FusedCfg {
  static constexpr int ND = Eq::NDIM, NV = Eq::NV, N = 4;
  static constexpr int NN = (ND == 3) ? 64 : 16, NF = NN / 4, NFACES = 2 * ND, NFN = NFACES * NF;
  static constexpr int EPB = FUSED_THREADS / NN;
  // shared doubles per element: q [NV][NN], exchange [2][NV][NN] (WF: flux tiles [ND][NV][NN]), sf [NV][NFN]
  static constexpr int XCH = (VI == TRIXIB200_VI_WEAK_FORM) ? ((ND > 2) ? 3 : 2) : 2;
  static constexpr int PER_ELEM = NV * NN + XCH * NV * NN + NV * NFN;
  static constexpr size_t SMEM = (size_t)PER_ELEM * EPB * sizeof(double);
};

template <class Eq, int VI, int VFLUX, int SFLUX, int FFLUX, bool NONCONS>
__global__ void __launch_bounds__(FUSED_THREADS, 3)
k_fused(Dev d, double* __restrict__ du, const double* __restrict__ u, double t, const int* __restrict__ elems,
        int64_t count) {
  using C = FusedCfg<Eq, VI>;
  constexpr int ND = C::ND, NV = C::NV, N = 4, NN = C::NN, NF = C::NF, NFACES = C::NFACES, NFN = C::NFN;
  extern __shared__ double smem[];
  const int tid = threadIdx.x, le = tid / NN, n = tid % NN;
  const int64_t slot = (int64_t)blockIdx.x * C::EPB + le;
  const bool active = slot < count;
  const int64_t e = active ? (elems ? (int64_t)elems[slot] : slot) : 0;
  double* sq = smem + (size_t)le * C::PER_ELEM;  // [NV][NN]
  double* sx = sq + NV * NN;                     // exchange / staging
  double* sf = sx + C::XCH * NV * NN;            // [NV][NFN] surface fluxes
  const Ops& op = *d.ops;
  const int vflux = (VFLUX >= 0) ? VFLUX : d.vol_flux;
  const int sflux = (SFLUX >= 0) ? SFLUX : d.surf_flux;
  const int fflux = (FFLUX >= 0) ? FFLUX : d.fv_flux;
  const int ix = n & 3, iy = (n >> 2) & 3, iz = (ND == 3) ? (n >> 4) : 0;
  const int idx[3] = {ix, iy, iz};

  // ---- coalesced load of the element block (AoS), staged through shared memory
  const double* ue = u + (size_t)NV * NN * e;
  if (active) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sx[n + k * NN] = ue[n + k * NN];
  }
  // ---- neighbour face data, issued early so the latency hides behind the volume work.
  // S1: all threads, faces 0..3 (3D) / all 4 faces (2D); S2 (3D only): warp 1 of the element, faces 4,5.
  double nb1[NV], nb2[NV];
  int code1 = NB_SFV, code2 = NB_SFV;
  const int face1 = n / NF, f1 = n % NF;
  const int face2 = 4 + ((n - 32) >> 4), f2 = n & 15;
  auto load_face = [&](int face, int f, int& code, double* nb) {
    code = d.face_nbr[(size_t)e * NFACES + face];
    const double* p;
    if (code >= 0) {
      int nbn = face_node<ND>(N, face >> 1, (face & 1) ? 0 : N - 1, f);
      p = u + ((size_t)code * NN + nbn) * NV;
    } else if (code == NB_SFV) {
      p = d.sfv + (size_t)NV * (f + (size_t)NF * (face + (size_t)NFACES * e));
    } else {
      p = d.halo_recv + ((size_t)nb_halo_slot(code) * NF + f) * NV;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) nb[v] = p[v];
  };
  if (active) {
    load_face(face1, f1, code1, nb1);
    if (ND == 3 && n >= 32) load_face(face2, f2, code2, nb2);
  }
  __syncthreads();
  double un[NV], q[NV], acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { un[v] = active ? sx[NV * n + v] : 1.0; acc[v] = 0; }
  Eq::to_qf(un, d.prm, q);
#pragma unroll
  for (int v = 0; v < NV; ++v) sq[v * NN + n] = q[v];
  __syncthreads();

  // surface flux at one face node: own trace from sq, neighbour trace from nb (or a ready-made flux)
  auto surface_flux = [&](int face, int f, int code, const double* nb) {
    double fl[NV];
    if (code == NB_SFV) {
#pragma unroll
      for (int v = 0; v < NV; ++v) fl[v] = nb[v];
    } else {
      const int dim = face >> 1, side = face & 1;
      const int own = face_node<ND>(N, dim, side ? N - 1 : 0, f);
      double qo[NV], qn[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) qo[v] = sq[v * NN + own];
      Eq::to_qf(nb, d.prm, qn);
      // order (ll, rr) by selects, not by a branch: both neighbours then evaluate bitwise the same flux
      double qa[NV], qb[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) { qa[v] = side ? qo[v] : qn[v]; qb[v] = side ? qn[v] : qo[v]; }
      Eq::two_point_qf(sflux, qa, qb, dim + 1, d.prm, fl);
      if (NONCONS) {
        double g[NV];
        Eq::noncons_q(qo, qn, dim + 1, d.prm, g);
#pragma unroll
        for (int v = 0; v < NV; ++v) fl[v] += 0.5 * g[v];
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) sf[v * NFN + face * NF + f] = fl[v];
  };

  if (VI == TRIXIB200_VI_WEAK_FORM) {
    // F_d(u_node) tiles, then du = sum_l Dhat[i,l] F1[l,j,k] + ... (reference dg_3d_kernel.jl:39-61)
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      double f[NV];
      Eq::flux(un, dd + 1, d.prm, f);
#pragma unroll
      for (int v = 0; v < NV; ++v) sx[(dd * NV + v) * NN + n] = f[v];
    }
    __syncthreads();
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const int st = (dd == 0) ? 1 : (dd == 1 ? 4 : 16);
      const int base = n - idx[dd] * st;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        double w = op.Dhat[idx[dd] + N * l];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += w * sx[(dd * NV + v) * NN + base + l * st];
      }
    }
    surface_flux(face1, f1, code1, nb1);
    if (ND == 3 && n >= 32) surface_flux(face2, f2, code2, nb2);
  } else {
    double alpha = 0.0;
    bool blend = false;
    if (VI == TRIXIB200_VI_SHOCK_CAPTURING_HG) {
      alpha = active ? d.alpha[e] : 0.0;
      blend = !(fabs(alpha) <= 1.8189894035458565e-12);  // reference dg_3d.jl:189
    }
    const double scale = blend ? 1.0 - alpha : 1.0;

    // one symmetric-pair slot: this lane evaluates the flux of (self, partner pc along dim dc) if `compute`,
    // publishes it, then picks up the flux of (self, partner pr along dim dr) evaluated by that partner.
    auto pair_slot = [&](int buf, bool compute, int dc, int pc, bool receive, int dr, int pr) {
      double* xb = sx + buf * NV * NN;
      if (compute) {
        const int st = (dc == 0) ? 1 : (dc == 1 ? 4 : 16);
        const int ic = (dc == 0) ? ix : (dc == 1 ? iy : iz);
        const int np = n + (pc - ic) * st;
        double qp[NV], f[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) qp[v] = sq[v * NN + np];
        // symmetric two-point flux: argument order is immaterial up to rounding, and both nodes of the pair use
        // this one value (a per-lane order branch would make the warp evaluate the flux twice)
        Eq::two_point_qf(vflux, q, qp, dc + 1, d.prm, f);
        const double w = scale * op.Dsplit[ic + N * pc];
#pragma unroll
        for (int v = 0; v < NV; ++v) { xb[v * NN + n] = f[v]; acc[v] += w * f[v]; }
      }
      __syncthreads();
      if (receive) {
        const int st = (dr == 0) ? 1 : (dr == 1 ? 4 : 16);
        const int ir = (dr == 0) ? ix : (dr == 1 ? iy : iz);
        const int np = n + (pr - ir) * st;
        const double w = scale * op.Dsplit[ir + N * pr];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += w * xb[v * NN + np];
      }
    };
    // slots 0..ND-1: matchings r0 + r2 of direction dd
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const int c = idx[dd];
      const bool even = (c & 1) == 0;
      pair_slot(dd & 1, true, dd, even ? (c ^ 1) : (c ^ 3), true, dd, even ? (c ^ 3) : (c ^ 1));
    }
    // slot ND: matchings r1 of x and y
    {
      const bool bx = (((ix >> 1) ^ (iy >> 1)) & 1) == 0;
      pair_slot(ND & 1, true, bx ? 0 : 1, bx ? (ix ^ 2) : (iy ^ 2), true, bx ? 1 : 0, bx ? (iy ^ 2) : (ix ^ 2));
    }
    if (ND == 3) {
      // slot 4: matching r1 of z on warp 0 of the element; warp 1 computes the z-face surface fluxes meanwhile
      const bool lowz = (iz & 2) == 0;
      if (!lowz) surface_flux(face2, f2, code2, nb2);
      pair_slot(0, lowz, 2, iz ^ 2, !lowz, 2, iz ^ 2);
    }
    if (NONCONS) {
      // remaining nonsymmetric volume terms: 0.5 * sum_l Dsplit[i,l] * nc(u_i, u_l) (reference dg_3d_kernel.jl:306-336)
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int st = (dd == 0) ? 1 : (dd == 1 ? 4 : 16);
        const int base = n - idx[dd] * st;
#pragma unroll
        for (int l = 0; l < N; ++l) {
          double ql[NV], g[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) ql[v] = sq[v * NN + base + l * st];
          Eq::noncons_q(q, ql, dd + 1, d.prm, g);
          const double w = scale * 0.5 * op.Dsplit[idx[dd] + N * l];
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] += w * g[v];
        }
      }
    }
    if (VI == TRIXIB200_VI_SHOCK_CAPTURING_HG) {
      // FV sub-cell fluxes f*(u_i, u_{i+1}) per direction, only where the element blends
      // (reference dg_3d_kernel.jl:460-487,576-581); syncs are unconditional, work is predicated.
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int st = (dd == 0) ? 1 : (dd == 1 ? 4 : 16);
        const int c = idx[dd];
        double* xb = sx + ((dd + 1) & 1) * NV * NN;  // alternates with the last pair slot (buffer 0)
        double fl[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) fl[v] = 0;
        if (blend && c < N - 1) {
          double qp[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) qp[v] = sq[v * NN + n + st];
          Eq::two_point_qf(fflux, q, qp, dd + 1, d.prm, fl);
#pragma unroll
          for (int v = 0; v < NV; ++v) xb[v * NN + n] = fl[v];
          if (NONCONS) {
            double g[NV];
            Eq::noncons_q(q, qp, dd + 1, d.prm, g);
#pragma unroll
            for (int v = 0; v < NV; ++v) fl[v] += 0.5 * g[v];
          }
        }
        __syncthreads();
        if (blend) {
          double fr[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) fr[v] = 0;
          if (c > 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v) fr[v] = xb[v * NN + n - st];
            if (NONCONS) {
              double qm[NV], g[NV];
#pragma unroll
              for (int v = 0; v < NV; ++v) qm[v] = sq[v * NN + n - st];
              Eq::noncons_q(q, qm, dd + 1, d.prm, g);
#pragma unroll
              for (int v = 0; v < NV; ++v) fr[v] += 0.5 * g[v];
            }
          }
          const double iw = op.inv_w[c];
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] += alpha * (iw * (fl[v] - fr[v]));
        }
      }
    }
    // remaining surface fluxes (faces 0..3 in 3D, all faces in 2D)
    surface_flux(face1, f1, code1, nb1);
  }
  __syncthreads();

  // ---- surface integral + Jacobian + sources (reference dg_3d_kernel.jl:1773-1844)
#pragma unroll
  for (int dd = 0; dd < ND; ++dd) {
    int f;
    if (ND == 2) f = idx[1 - dd];
    else f = (dd == 0) ? iy + 4 * iz : (dd == 1 ? ix + 4 * iz : ix + 4 * iy);
    if (idx[dd] == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] -= sf[v * NFN + (2 * dd) * NF + f] * op.factor_1;
    }
    if (idx[dd] == N - 1) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] += sf[v * NFN + (2 * dd + 1) * NF + f] * op.factor_2;
    }
  }
  const double inv_jac = active ? d.inv_jac[e] : 1.0;
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] *= -inv_jac;
  if (d.src != TRIXIB200_SRC_NONE && active) {
    double x[3] = {0, 0, 0}, s[NV];
    if (d.node_coords) {
#pragma unroll
      for (int k = 0; k < ND; ++k) x[k] = d.node_coords[k + (size_t)ND * (n + (size_t)NN * e)];
    } else {
      double jac = 1.0 / inv_jac;
#pragma unroll
      for (int k = 0; k < ND; ++k) x[k] = __dadd_rn(d.centers[k + (size_t)ND * e], __dmul_rn(jac, op.nodes[idx[k]]));
    }
    Eq::source(d.src, un, x, t, d.prm, s);
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += s[v];
  }
  // ---- transpose back to AoS through shared memory, coalesced store
#pragma unroll
  for (int v = 0; v < NV; ++v) sx[NV * n + v] = acc[v];
  __syncthreads();
  if (active) {
    double* due = du + (size_t)NV * NN * e;
#pragma unroll
    for (int k = 0; k < NV; ++k) due[n + k * NN] = sx[n + k * NN];
  }
}

// ---------------------------------------------------------------------------------------------- host side
inline bool fused_available(const trixib200_config& c) {
  return (c.ndim == 2 || c.ndim == 3) && c.polydeg == 3;
}

template <class Eq, int VI, int VFLUX, int SFLUX, int FFLUX, bool NONCONS>
static int fused_launch_t(const Dev& d, double* du, const double* u, double t, const int* elems, int64_t count,
                          cudaStream_t stream) {
  using C = FusedCfg<Eq, VI>;
  auto kern = k_fused<Eq, VI, VFLUX, SFLUX, FFLUX, NONCONS>;
  static DeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
      { configured.undo(); return TRIXIB200_ECUDA; }
  }
  if (count <= 0) return 0;
  unsigned blocks = (unsigned)((count + C::EPB - 1) / C::EPB);
  kern<<<blocks, FUSED_THREADS, C::SMEM, stream>>>(d, du, u, t, elems, count);
  return cudaGetLastError() == cudaSuccess ? 0 : TRIXIB200_ECUDA;
}

template <class Eq, bool NONCONS>
static int fused_launch_eq(const trixib200_config& c, const Dev& d, double* du, const double* u, double t,
                           const int* elems, int64_t count, cudaStream_t s) {
  constexpr int R = TRIXIB200_FLUX_RANOCHA;
  switch (c.volume_integral) {
    case TRIXIB200_VI_WEAK_FORM:
      return fused_launch_t<Eq, TRIXIB200_VI_WEAK_FORM, -1, -1, -1, NONCONS>(d, du, u, t, elems, count, s);
    case TRIXIB200_VI_FLUX_DIFFERENCING:
      if (Eq::KIND == TRIXIB200_EQ_EULER && c.volume_flux == R && c.surface_flux == R)
        return fused_launch_t<Eq, TRIXIB200_VI_FLUX_DIFFERENCING, (Eq::KIND == TRIXIB200_EQ_EULER ? R : -1),
                              (Eq::KIND == TRIXIB200_EQ_EULER ? R : -1), -1, NONCONS>(d, du, u, t, elems, count, s);
      return fused_launch_t<Eq, TRIXIB200_VI_FLUX_DIFFERENCING, -1, -1, -1, NONCONS>(d, du, u, t, elems, count, s);
    default:
      if (Eq::KIND == TRIXIB200_EQ_EULER && c.volume_flux == R && c.surface_flux == R && c.volume_flux_fv == R)
        return fused_launch_t<Eq, TRIXIB200_VI_SHOCK_CAPTURING_HG, (Eq::KIND == TRIXIB200_EQ_EULER ? R : -1),
                              (Eq::KIND == TRIXIB200_EQ_EULER ? R : -1), (Eq::KIND == TRIXIB200_EQ_EULER ? R : -1),
                              NONCONS>(d, du, u, t, elems, count, s);
      return fused_launch_t<Eq, TRIXIB200_VI_SHOCK_CAPTURING_HG, -1, -1, -1, NONCONS>(d, du, u, t, elems, count, s);
  }
}

static int fused_launch(const trixib200_config& c, const Dev& d, double* du, const double* u, double t,
                        const int* elems, int64_t count, cudaStream_t s, int /*sm_count*/) {
  switch (c.equations) {
    case TRIXIB200_EQ_ADVECTION:
      if (c.ndim == 2) return fused_launch_eq<EqAdvection<2>, false>(c, d, du, u, t, elems, count, s);
      return fused_launch_eq<EqAdvection<3>, false>(c, d, du, u, t, elems, count, s);
    case TRIXIB200_EQ_EULER:
      if (c.ndim == 2) return fused_launch_eq<EqEuler<2>, false>(c, d, du, u, t, elems, count, s);
      return fused_launch_eq<EqEuler<3>, false>(c, d, du, u, t, elems, count, s);
    default:
      return fused_launch_eq<EqMhd3, true>(c, d, du, u, t, elems, count, s);
  }
}

}  // namespace tb
