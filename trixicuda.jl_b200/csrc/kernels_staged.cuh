// Staged kernels: one per rhs! stage, materialising interfaces.u / surface_flux_values / mortar buffers in
// Trixi's layouts. They cover every enumerated feature for any polydeg in 1D/2D/3D, back the per-stage
// entry points (trixib200_stage) and the faces the fused path delegates (boundaries, mortars).
// Stage semantics follow the reference's rhs_gpu! (src/solvers/dg_3d.jl:895-925) stage by stage; the
// formulas cite the reference kernels they replace.
#pragma once
#include "device.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------------
// Volume integral. Replaces flux_kernel!/weak_form_kernel!/flux_weak_form_kernel! (reference
// dg_3d_kernel.jl:8-120), volume_flux_kernel!/volume_integral_kernel!/volume_flux_integral_kernel!
// (:123-423, cons + noncons) and the six *_dgfv_* kernels (:426-1118). One thread per (node, element);
// du is overwritten (fused reset, reference dg_3d.jl:897-899).
// ------------------------------------------------------------------------------------------------
template <class Eq>
__global__ void k_volume(Dev d, double* __restrict__ du, const double* __restrict__ u) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int N = d.N, nn = d.nn;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.E * nn) return;
  int64_t e = gid / nn;
  int n = (int)(gid - e * nn);
  const double* ue = u + (size_t)NV * nn * e;
  const Ops& op = *d.ops;
  int idx[3] = {n % N, (n / N) % N, n / (N * N)};
  double un[NV], acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { un[v] = ue[NV * n + v]; acc[v] = 0; }

  if (d.volume_integral == TRIXIB200_VI_WEAK_FORM) {
    for (int dd = 0; dd < ND; ++dd) {
      int st = ipow_stride(N, dd), base = n - idx[dd] * st;
      for (int l = 0; l < N; ++l) {
        double ul[NV], f[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) ul[v] = ue[NV * (base + l * st) + v];
        Eq::flux(ul, dd + 1, d.prm, f);
        double w = op.Dhat[idx[dd] + N * l];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += w * f[v];
      }
    }
  } else {
    double a = 0.0;
    bool blend = false;
    if (d.volume_integral == TRIXIB200_VI_SHOCK_CAPTURING_HG) {
      a = d.alpha[e];
      // dg_only = isapprox(alpha, 0, atol = max(100 eps, eps^0.75))  (reference dg_3d.jl:189)
      blend = !(fabs(a) <= 1.8189894035458565e-12);
    }
    double scale = blend ? 1.0 - a : 1.0;
    for (int dd = 0; dd < ND; ++dd) {
      int st = ipow_stride(N, dd), base = n - idx[dd] * st, i = idx[dd];
      for (int l = 0; l < N; ++l) {
        if (l == i && !(Eq::HAS_NONCONS && d.noncons)) continue;
        double ul[NV], f[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) ul[v] = ue[NV * (base + l * st) + v];
        double w = scale * op.Dsplit[i + N * l];
        if (l != i) {
          // symmetric two-point flux, evaluated lower-node-first like Trixi's flux_differencing_kernel!
          if (l > i) Eq::two_point(d.vol_flux, un, ul, dd + 1, d.prm, f);
          else Eq::two_point(d.vol_flux, ul, un, dd + 1, d.prm, f);
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] += w * f[v];
        }
        if (Eq::HAS_NONCONS && d.noncons) {
          Eq::noncons(un, ul, dd + 1, d.prm, f);
          double w2 = scale * 0.5 * op.Dsplit[i + N * l];
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] += w2 * f[v];
        }
      }
      if (blend) {
        // FV sub-cell fluxes: fstar_L[i+1] - fstar_R[i] (reference dg_3d_kernel.jl:576-581)
        double fl[NV], fr[NV], g[NV], um[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { fl[v] = 0; fr[v] = 0; }
        if (i < N - 1) {
#pragma unroll
          for (int v = 0; v < NV; ++v) um[v] = ue[NV * (n + st) + v];
          Eq::two_point(d.fv_flux, un, um, dd + 1, d.prm, fl);
          if (Eq::HAS_NONCONS && d.noncons) {
            Eq::noncons(un, um, dd + 1, d.prm, g);
#pragma unroll
            for (int v = 0; v < NV; ++v) fl[v] += 0.5 * g[v];
          }
        }
        if (i > 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) um[v] = ue[NV * (n - st) + v];
          Eq::two_point(d.fv_flux, um, un, dd + 1, d.prm, fr);
          if (Eq::HAS_NONCONS && d.noncons) {
            Eq::noncons(un, um, dd + 1, d.prm, g);
#pragma unroll
            for (int v = 0; v < NV; ++v) fr[v] += 0.5 * g[v];
          }
        }
        double iw = op.inv_w[i];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += a * (iw * (fl[v] - fr[v]));
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) du[(size_t)NV * nn * e + NV * n + v] = acc[v];
}

// ------------------------------------------------------------------------------------------------
// Hennemann-Gassner indicator on the device (the reference runs Trixi's CPU loop after a D2H copy of u:
// src/solvers/indicators.jl:7-42, dg_3d.jl:187-188). One block per element, nn threads.
// ------------------------------------------------------------------------------------------------
template <class Eq>
__global__ void k_indicator(Dev d, const double* __restrict__ u) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  extern __shared__ double sm[];
  const int N = d.N, nn = d.nn;
  double* a = sm;
  double* b = sm + nn;
  int64_t e = blockIdx.x;
  int n = threadIdx.x;
  const Ops& op = *d.ops;
  if (n < nn) {
    double un[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) un[v] = u[(size_t)NV * nn * e + NV * n + v];
    a[n] = Eq::indicator_var(d.ind_var, un, d.prm);
  }
  __syncthreads();
  for (int dd = 0; dd < ND; ++dd) {
    if (n < nn) {
      int st = ipow_stride(N, dd);
      int id = (n / st) % N, base = n - id * st;
      double s = 0;
      for (int l = 0; l < N; ++l) s += op.invV[id + N * l] * a[base + l * st];
      b[n] = s;
    }
    __syncthreads();
    double* t = a; a = b; b = t;
  }
  if (n == 0) {
    double total = 0, clip1 = 0, clip2 = 0;
    for (int m = 0; m < nn; ++m) {
      int i0 = m % N, i1 = (m / N) % N, i2 = m / (N * N);
      int mx = max(i0, max(i1, i2));
      double m2 = a[m] * a[m];
      total += m2;
      if (mx < N - 1) clip1 += m2;
      if (mx < N - 2) clip2 += m2;
    }
    double f1 = (total != 0.0) ? (total - clip1) / total : 0.0;
    double f2 = (clip1 != 0.0) ? (clip1 - clip2) / clip1 : 0.0;
    double energy = fmax(f1, f2);
    // constants: reference src/solvers/indicators.jl:21-24
    double threshold = 0.5 * pow(10.0, -1.8 * pow((double)N, 0.25));
    double parameter_s = log((1 - 0.0001) / 0.0001);
    double al = 1 / (1 + exp(-parameter_s / threshold * (energy - threshold)));
    if (al < d.alpha_min) al = 0;
    if (al > 1 - d.alpha_min) al = 1;
    al = fmin(d.alpha_max, al);
    d.alpha[e] = al;
    d.alpha_tmp[e] = al;
  }
}

TB_D void atomic_max_nonneg(double* addr, double val) {
  // alpha >= 0: IEEE order == unsigned integer order
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(val));
}

// apply_smoothing!: alpha[e] = max(alpha_tmp[e], 0.5 * alpha_tmp[neighbours]) over interfaces and mortars
__global__ void k_alpha_smooth_interfaces(Dev d) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.I) return;
  int l = d.if_left[s], r = d.if_right[s];
  // a side on another rank: its indicator value arrived with the alpha halo exchange (slot = the face's halo slot);
  // only the local side is updated, the owner of the other side does the mirror-image update
  if (l >= 0) atomic_max_nonneg(&d.alpha[l], 0.5 * (r >= 0 ? d.alpha_tmp[r] : d.halo_alpha_recv[nb_halo_slot(r)]));
  if (r >= 0) atomic_max_nonneg(&d.alpha[r], 0.5 * (l >= 0 ? d.alpha_tmp[l] : d.halo_alpha_recv[nb_halo_slot(l)]));
}
// indicator value of the element behind every face this rank sends (same order as the trace exchange)
__global__ void k_pack_alpha(Dev d) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < d.nhalo_send) d.halo_alpha_send[s] = d.alpha_tmp[d.send_elem[s]];
}
__global__ void k_alpha_smooth_mortars(Dev d, int nsmall) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.M) return;
  int large = d.mo_ids[(nsmall + 1) * m + nsmall];
  const double a_large = large >= 0 ? d.alpha_tmp[large] : d.halo_alpha_recv[nb_halo_slot(large)];
  for (int q = 0; q < nsmall; ++q) {
    int sm_ = d.mo_ids[(nsmall + 1) * m + q];
    const double a_small = sm_ >= 0 ? d.alpha_tmp[sm_] : d.halo_alpha_recv[nb_halo_slot(sm_)];
    if (sm_ >= 0) atomic_max_nonneg(&d.alpha[sm_], 0.5 * a_large);
    if (large >= 0) atomic_max_nonneg(&d.alpha[large], 0.5 * a_small);
  }
}

// ------------------------------------------------------------------------------------------------
// prolong2interfaces (reference dg_3d_kernel.jl:1121-1152): thread per (v, face node, interface)
// ------------------------------------------------------------------------------------------------
template <int ND>
__global__ void k_prolong_interfaces(Dev d, const double* __restrict__ u) {
  const int N = d.N, nn = d.nn, nf = d.nf, nv = d.nv;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.I * nf * nv) return;
  int v = (int)(gid % nv);
  int f = (int)((gid / nv) % nf);
  int64_t s = gid / ((int64_t)nv * nf);
  int dim = d.if_dim[s];
  int l = d.if_left[s], r = d.if_right[s];
  double vl, vr;
  if (l >= 0) vl = u[(size_t)nv * nn * l + nv * face_node<ND>(N, dim, N - 1, f) + v];
  else vl = d.halo_recv[((size_t)nb_halo_slot(l) * nf + f) * nv + v];
  if (r >= 0) vr = u[(size_t)nv * nn * r + nv * face_node<ND>(N, dim, 0, f) + v];
  else vr = d.halo_recv[((size_t)nb_halo_slot(r) * nf + f) * nv + v];
  size_t o = 2 * (v + (size_t)nv * (f + (size_t)nf * s));
  d.interfaces_u[o] = vl;
  d.interfaces_u[o + 1] = vr;
}

// interface flux (reference dg_3d_kernel.jl:1155-1262): thread per (face node, interface); writes both
// surface_flux_values slots (left: direction 2o, right: 2o-1), noncons adds 0.5*nc(ll,rr) / 0.5*nc(rr,ll)
template <class Eq>
__global__ void k_interface_flux(Dev d) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int nf = d.nf;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.I * nf) return;
  int f = (int)(gid % nf);
  int64_t s = gid / nf;
  int dim = d.if_dim[s];
  int l = d.if_left[s], r = d.if_right[s];
  double ul[NV], ur[NV], fl[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    size_t o = 2 * (v + (size_t)NV * (f + (size_t)nf * s));
    ul[v] = d.interfaces_u[o];
    ur[v] = d.interfaces_u[o + 1];
  }
  Eq::two_point(d.surf_flux, ul, ur, dim + 1, d.prm, fl);
  double gl[NV], gr[NV];
  bool nc = Eq::HAS_NONCONS && d.noncons;
  if (nc) { Eq::noncons(ul, ur, dim + 1, d.prm, gl); Eq::noncons(ur, ul, dim + 1, d.prm, gr); }
  if (l >= 0) {
    double* sl = d.sfv + (size_t)NV * (f + (size_t)nf * ((2 * dim + 1) + (size_t)2 * ND * l));
#pragma unroll
    for (int v = 0; v < NV; ++v) sl[v] = nc ? fl[v] + 0.5 * gl[v] : fl[v];
  }
  if (r >= 0) {
    double* sr = d.sfv + (size_t)NV * (f + (size_t)nf * ((2 * dim) + (size_t)2 * ND * r));
#pragma unroll
    for (int v = 0; v < NV; ++v) sr[v] = nc ? fl[v] + 0.5 * gr[v] : fl[v];
  }
}

// ------------------------------------------------------------------------------------------------
// boundaries (reference dg_3d_kernel.jl:1265-1358)
// ------------------------------------------------------------------------------------------------
template <int ND>
__global__ void k_prolong_boundaries(Dev d, const double* __restrict__ u) {
  const int N = d.N, nn = d.nn, nf = d.nf, nv = d.nv;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.B * nf * nv) return;
  int v = (int)(gid % nv);
  int f = (int)((gid / nv) % nf);
  int64_t b = gid / ((int64_t)nv * nf);
  int side = d.bd_side[b];  // 1: element on the left (its +face), 2: element on the right
  int e = d.bd_elem[b];
  int n = face_node<ND>(N, d.bd_dim[b], side == 1 ? N - 1 : 0, f);
  size_t o = 2 * (v + (size_t)nv * (f + (size_t)nf * b));
  double val = u[(size_t)nv * nn * e + nv * n + v];
  d.boundaries_u[o + (side - 1)] = val;
  d.boundaries_u[o + (2 - side)] = 0.0;  // "Set to 0 instead of NaN" (reference dg_3d_kernel.jl:1287,1292)
}

template <class Eq>
__global__ void k_boundary_flux(Dev d, double t) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int nf = d.nf;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.B * nf) return;
  int f = (int)(gid % nf);
  int64_t b = gid / nf;
  int dir = d.bd_dir[b];  // 0-based direction
  if (d.bc[dir] == TRIXIB200_BC_PERIODIC) return;
  int side = d.bd_side[b], dim = d.bd_dim[b], e = d.bd_elem[b];
  double ui[NV], ub[NV], fl[NV], x[3] = {0, 0, 0};
#pragma unroll
  for (int v = 0; v < NV; ++v) ui[v] = d.boundaries_u[(side - 1) + 2 * (v + (size_t)NV * (f + (size_t)nf * b))];
#pragma unroll
  for (int q = 0; q < ND; ++q) x[q] = d.bd_coords[q + (size_t)ND * (f + (size_t)nf * b)];
  if (d.bc[dir] == TRIXIB200_BC_SLIP_WALL) {
    Eq::slip_wall_flux(ui, dim + 1, dir, d.prm, fl);   // create() admits it for compressible Euler only
  } else {
    // BoundaryConditionDirichlet(initial_condition): call contract reference dg_3d_kernel.jl:1327-1343
    Eq::initial_condition(d.ic, x, t, d.prm, ub);
    if (dir % 2 == 1) Eq::two_point(d.surf_flux, ui, ub, dim + 1, d.prm, fl);
    else Eq::two_point(d.surf_flux, ub, ui, dim + 1, d.prm, fl);
  }
  double* s = d.sfv + (size_t)NV * (f + (size_t)nf * (dir + (size_t)2 * ND * e));
#pragma unroll
  for (int v = 0; v < NV; ++v) s[v] = fl[v];
}

// ------------------------------------------------------------------------------------------------
// mortars (reference dg_3d_kernel.jl:1361-1770, 2D dg_2d_kernel.jl:1167-1400)
// ------------------------------------------------------------------------------------------------
// prolong2mortars: thread per (v, face node, mortar). Small faces are copied; the large face is
// interpolated with forward_{lower,upper} per face dimension (dim-1 operator applied first).
template <int ND>
__global__ void k_prolong_mortars(Dev d, const double* __restrict__ u) {
  const int N = d.N, nn = d.nn, nf = d.nf, nv = d.nv;
  constexpr int NS = (ND == 3) ? 4 : 2;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.M * nf * nv) return;
  int v = (int)(gid % nv);
  int f = (int)((gid / nv) % nf);
  int64_t m = gid / ((int64_t)nv * nf);
  int dim = d.mo_dim[m], ls = d.mo_side[m];
  const int* ids = d.mo_ids + (NS + 1) * m;
  int large = ids[NS];
  const Ops& op = *d.ops;
  int fixed_small = (ls == 1) ? 0 : N - 1, fixed_large = (ls == 1) ? N - 1 : 0;
  int a = f % N, b = (ND == 3) ? f / N : 0;
  // face value of a participant: a local element's node, or (element on another rank) its received trace [f][v]
  auto face_val = [&](int e, int fixed, int ff) -> double {
    return e >= 0 ? u[(size_t)nv * nn * e + nv * face_node<ND>(N, dim, fixed, ff) + v]
                  : d.halo_recv[(size_t)nv * (ff + (size_t)nf * nb_halo_slot(e)) + v];
  };
  for (int q = 0; q < NS; ++q) {
    size_t o = 2 * (v + (size_t)nv * (f + (size_t)nf * m));
    int e = ids[mortar_small_row<ND>(q)];
    d.mortar_u[q][o + (2 - ls)] = face_val(e, fixed_small, f);
    const double *M1, *M2;
    if (ND == 3) {
      M1 = (q == 0 || q == 2) ? op.fwd_l : op.fwd_u;
      M2 = (q == 0 || q == 1) ? op.fwd_u : op.fwd_l;
    } else {
      M1 = (q == 0) ? op.fwd_u : op.fwd_l;
      M2 = M1;
    }
    double s = 0;
    if (ND == 3) {
      for (int bb = 0; bb < N; ++bb) {
        double t1 = 0;
        for (int aa = 0; aa < N; ++aa)
          t1 += M1[a + N * aa] * face_val(large, fixed_large, aa + N * bb);
        s += M2[b + N * bb] * t1;
      }
    } else {
      for (int aa = 0; aa < N; ++aa)
        s += M1[a + N * aa] * face_val(large, fixed_large, aa);
    }
    d.mortar_u[q][o + (ls - 1)] = s;
  }
}

// mortar flux: thread per (face node, mortar, q). fstar_primary = fstar_secondary = surface flux;
// noncons: primary += 0.5 nc(large-side state, small-side state), secondary reversed
// (reference dg_3d_kernel.jl:1598-1642)
template <class Eq>
__global__ void k_mortar_flux(Dev d) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  constexpr int NS = (ND == 3) ? 4 : 2;
  const int nf = d.nf;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.M * nf * NS) return;
  int f = (int)(gid % nf);
  int q = (int)((gid / nf) % NS);
  int64_t m = gid / ((int64_t)nf * NS);
  int dim = d.mo_dim[m], ls = d.mo_side[m];
  double ul[NV], ur[NV], fl[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    size_t o = 2 * (v + (size_t)NV * (f + (size_t)nf * m));
    ul[v] = d.mortar_u[q][o];
    ur[v] = d.mortar_u[q][o + 1];
  }
  Eq::two_point(d.surf_flux, ul, ur, dim + 1, d.prm, fl);
  double* fp = d.fstar_p[q] + (size_t)NV * (f + (size_t)nf * m);
  double* fs = d.fstar_s[q] + (size_t)NV * (f + (size_t)nf * m);
  if (Eq::HAS_NONCONS && d.noncons) {
    double gp[NV], gs[NV];
    const double* u1 = (ls == 1) ? ul : ur;
    const double* u2 = (ls == 1) ? ur : ul;
    Eq::noncons(u1, u2, dim + 1, d.prm, gp);
    Eq::noncons(u2, u1, dim + 1, d.prm, gs);
#pragma unroll
    for (int v = 0; v < NV; ++v) { fp[v] = fl[v] + 0.5 * gp[v]; fs[v] = fl[v] + 0.5 * gs[v]; }
  } else {
#pragma unroll
    for (int v = 0; v < NV; ++v) { fp[v] = fl[v]; fs[v] = fl[v]; }
  }
}

// mortar_fluxes_to_elements: small elements take fstar_primary at direction 2o+ls-2; the large element takes
// the L2 projection of fstar_secondary at direction 2o-ls+1 (reference dg_3d_kernel.jl:1684-1766).
template <int ND>
__global__ void k_mortar_to_elements(Dev d) {
  const int N = d.N, nf = d.nf, nv = d.nv;
  constexpr int NS = (ND == 3) ? 4 : 2;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.M * nf * nv) return;
  int v = (int)(gid % nv);
  int f = (int)((gid / nv) % nf);
  int64_t m = gid / ((int64_t)nv * nf);
  int o1 = d.mo_dim[m] + 1, ls = d.mo_side[m];
  const int* ids = d.mo_ids + (NS + 1) * m;
  int dir_small = 2 * o1 + ls - 2 - 1, dir_large = 2 * o1 - ls + 1 - 1;  // 0-based
  const Ops& op = *d.ops;
  int a = f % N, b = (ND == 3) ? f / N : 0;
  double total = 0;
  for (int q = 0; q < NS; ++q) {
    int e = ids[mortar_small_row<ND>(q)];
    if (e >= 0)      // (a replicated mortar: elements of other ranks take their fluxes from their own copy)
      d.sfv[(size_t)nv * (f + (size_t)nf * (dir_small + (size_t)2 * ND * e)) + v] =
          d.fstar_p[q][(size_t)nv * (f + (size_t)nf * m) + v];
    const double *M1, *M2;
    if (ND == 3) {
      M1 = (q == 0 || q == 2) ? op.rev_l : op.rev_u;
      M2 = (q == 0 || q == 1) ? op.rev_u : op.rev_l;
    } else {
      M1 = (q == 0) ? op.rev_u : op.rev_l;
      M2 = M1;
    }
    const double* fs = d.fstar_s[q] + (size_t)nv * nf * m;
    double s = 0;
    if (ND == 3) {
      for (int bb = 0; bb < N; ++bb) {
        double t1 = 0;
        for (int aa = 0; aa < N; ++aa) t1 += M1[a + N * aa] * fs[nv * (aa + N * bb) + v];
        s += M2[b + N * bb] * t1;
      }
    } else {
      for (int aa = 0; aa < N; ++aa) s += M1[a + N * aa] * fs[nv * aa + v];
    }
    total += s;
  }
  int large = ids[NS];
  if (large >= 0) d.sfv[(size_t)nv * (f + (size_t)nf * (dir_large + (size_t)2 * ND * large)) + v] = total;
}

// ------------------------------------------------------------------------------------------------
// surface integral + Jacobian + sources (reference dg_3d_kernel.jl:1773-1844), one sweep over du.
// flags: 1 surface integral, 2 apply_jacobian, 4 sources
// ------------------------------------------------------------------------------------------------
template <class Eq>
__global__ void k_epilogue(Dev d, double* __restrict__ du, const double* __restrict__ u, double t, int flags) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int N = d.N, nn = d.nn, nf = d.nf;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.E * nn) return;
  int64_t e = gid / nn;
  int n = (int)(gid - e * nn);
  int idx[3] = {n % N, (n / N) % N, n / (N * N)};
  double acc[NV];
  size_t off = (size_t)NV * nn * e + NV * n;
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = du[off + v];
  if (flags & 1) {
    const Ops& op = *d.ops;
    for (int dd = 0; dd < ND; ++dd) {
      int f;
      if (ND == 1) f = 0;
      else if (ND == 2) f = idx[1 - dd];
      else f = (dd == 0) ? idx[1] + N * idx[2] : (dd == 1 ? idx[0] + N * idx[2] : idx[0] + N * idx[1]);
      if (idx[dd] == 0) {
        const double* s = d.sfv + (size_t)NV * (f + (size_t)nf * ((2 * dd) + (size_t)2 * ND * e));
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] -= s[v] * op.factor_1;
      }
      if (idx[dd] == N - 1) {
        const double* s = d.sfv + (size_t)NV * (f + (size_t)nf * ((2 * dd + 1) + (size_t)2 * ND * e));
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += s[v] * op.factor_2;
      }
    }
  }
  if (flags & 2) {
    double fac = -d.inv_jac[e];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] *= fac;
  }
  if ((flags & 4) && d.src != TRIXIB200_SRC_NONE) {
    double un[NV], s[NV], x[3] = {0, 0, 0};
#pragma unroll
    for (int v = 0; v < NV; ++v) un[v] = u[off + v];
    if (d.node_coords) {
#pragma unroll
      for (int q = 0; q < ND; ++q) x[q] = d.node_coords[q + (size_t)ND * (n + (size_t)nn * e)];
    } else {
      double jac = 1.0 / d.inv_jac[e];
#pragma unroll
      for (int q = 0; q < ND; ++q) x[q] = __dadd_rn(d.centers[q + (size_t)ND * e], __dmul_rn(jac, d.ops->nodes[idx[q]]));
    }
    Eq::source(d.src, un, x, t, d.prm, s);
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += s[v];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) du[off + v] = acc[v];
}

// ------------------------------------------------------------------------------------------------
// max_dt reduction (reference src/callbacks_step/stepsize_dg_3d.jl:20-45 copies u to the host). One warp
// per element: per-direction node maxima by shuffles, then inv_jac * sum; block max -> atomicMax.
// ------------------------------------------------------------------------------------------------
template <class Eq>
__global__ void k_max_dt(Dev d, const double* __restrict__ u, double* out) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int nn = d.nn;
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double best = 0.0;
  for (int64_t e = warp; e < d.E; e += nwarps) {
    double ml[3] = {0, 0, 0};
    for (int n = lane; n < nn; n += 32) {
      double un[NV], lam[3];
#pragma unroll
      for (int v = 0; v < NV; ++v) un[v] = u[(size_t)NV * nn * e + NV * n + v];
      Eq::max_abs_speeds(un, d.prm, lam);
#pragma unroll
      for (int q = 0; q < ND; ++q) ml[q] = fmax(ml[q], lam[q]);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < ND; ++q) {
      double m = ml[q];
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      s += m;
    }
    best = fmax(best, d.inv_jac[e] * s);
  }
  if (lane == 0) atomic_max_nonneg(out, best);
}

// nextfloat / prevfloat for finite doubles
TB_D double next_float(double x, bool up) {
  if (x == 0.0) return up ? 4.9406564584124654e-324 : -4.9406564584124654e-324;
  long long b = __double_as_longlong(x);
  b += ((x > 0) == up) ? 1 : -1;
  return __longlong_as_double(b);
}

// enumerated initial condition on the nodes (1D end-node nudging as Trixi's compute_coefficients!,
// mirrored at reference src/solvers/dg.jl:54-58)
template <class Eq>
__global__ void k_fill_ic(Dev d, double* __restrict__ u, double t) {
  constexpr int NV = Eq::NV, ND = Eq::NDIM;
  const int N = d.N, nn = d.nn;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.E * nn) return;
  int64_t e = gid / nn;
  int n = (int)(gid - e * nn);
  int idx[3] = {n % N, (n / N) % N, n / (N * N)};
  double x[3] = {0, 0, 0}, un[NV];
  if (d.node_coords) {
#pragma unroll
    for (int q = 0; q < ND; ++q) x[q] = d.node_coords[q + (size_t)ND * (n + (size_t)nn * e)];
  } else {
    double jac = 1.0 / d.inv_jac[e];
#pragma unroll
    for (int q = 0; q < ND; ++q) x[q] = __dadd_rn(d.centers[q + (size_t)ND * e], __dmul_rn(jac, d.ops->nodes[idx[q]]));
  }
  if (ND == 1) {
    if (n == 0) x[0] = next_float(x[0], true);
    else if (n == N - 1) x[0] = next_float(x[0], false);
  }
  Eq::initial_condition(d.ic, x, t, d.prm, un);
#pragma unroll
  for (int v = 0; v < NV; ++v) u[(size_t)NV * nn * e + NV * n + v] = un[v];
}

// 2N low-storage RK stage: tmp = a*tmp + dt*du; u += b*tmp
__global__ void k_rk2n_update(double* __restrict__ u, double* __restrict__ tmp, const double* __restrict__ du,
                              double a, double b, double dt, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    // a == 0 (first stage of a 2N scheme): tmp is not read, it may be uninitialised memory
    double t = a != 0.0 ? fma(a, tmp[i], dt * du[i]) : dt * du[i];
    tmp[i] = t;
    u[i] += b * t;
  }
}

// pack the face traces that peers need: halo_send[v, f, s] = u[v, face_node(dir), elem]
template <int ND>
__global__ void k_pack_halo(Dev d, const double* __restrict__ u) {
  const int N = d.N, nn = d.nn, nf = d.nf, nv = d.nv;
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= d.nhalo_send * nf * nv) return;
  int v = (int)(gid % nv);
  int f = (int)((gid / nv) % nf);
  int64_t s = gid / ((int64_t)nv * nf);
  int dir = d.send_dir[s];
  int n = face_node<ND>(N, dir / 2, (dir & 1) ? N - 1 : 0, f);
  d.halo_send[gid] = u[(size_t)nv * nn * d.send_elem[s] + nv * n + v];
}

}  // namespace tb
