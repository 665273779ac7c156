// TEST INFRASTRUCTURE ONLY -- CPU oracle ("parity unpinned", see basis.hpp header).
//
// The 10 rhs! stages in Trixi.jl's CPU order and loop/summation order, whose stage names and argument
// orders are visible in the reference's tests (/root/reference/test/tree_dgsem_3d/euler_ec.jl:55-120)
// and whose GPU restatement is /root/reference/src/solvers/dg_3d.jl:895-925 (2D dg_2d.jl:760, 1D dg_1d.jl:523).
// Also: HG indicator (/root/reference/src/solvers/indicators.jl:7-42), max_dt
// (/root/reference/src/callbacks_step/stepsize_dg_3d.jl:20-45), error norms
// (/root/reference/src/callbacks_step/analysis_dg_3d.jl:45-89).
// Layout everywhere is Trixi's: u[v, i, j, k, element] column-major, indices 1-based in the containers.
#pragma once
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include "basis.hpp"
#include "tree.hpp"
#include "equations.hpp"

namespace orc {

enum VolumeIntegralKind { VI_WEAK_FORM = 0, VI_FLUX_DIFFERENCING = 1, VI_SHOCK_CAPTURING_HG = 2 };
enum IndicatorVariable { IND_DENSITY = 0, IND_PRESSURE = 1, IND_DENSITY_PRESSURE = 2 };
enum BCKind { BC_PERIODIC = 0, BC_DIRICHLET_IC = 1, BC_SLIP_WALL = 2 };

struct SolverConfig {
  EqParams eq;
  int polydeg = 3;
  int volume_integral = VI_WEAK_FORM;
  int volume_flux = FLUX_CENTRAL;       // volume_flux (FD) / volume_flux_dg (SC)
  int volume_flux_fv = FLUX_LAX_FRIEDRICHS;
  int surface_flux = FLUX_LAX_FRIEDRICHS;
  int nonconservative = 0;              // 1: (flux, flux_nonconservative_powell) tuples
  double alpha_max = 0.5, alpha_min = 0.001;
  int alpha_smooth = 1;
  int indicator_variable = IND_DENSITY_PRESSURE;
  int bc[6] = {0, 0, 0, 0, 0, 0};
  int initial_condition = IC_CONVERGENCE_TEST;  // also used by BC_DIRICHLET_IC and error norms
  int source = SRC_NONE;
};

struct SolverBase {
  SolverConfig cfg;
  Basis basis;
  Containers c;
  int nd, N, nv, nn, nf;  // dims, nodes/dim, vars, nodes/element, nodes/face
  int stride[3];
  // cache (Trixi layouts)
  vec interfaces_u;          // [2, nv, nf, I]
  vec surface_flux_values;   // [nv, nf, 2*nd, E]
  vec boundaries_u;          // [2, nv, nf, B]
  vec mortar_u[4];           // 3D: upper_left, upper_right, lower_left, lower_right; 2D: upper, lower. each [2,nv,nf,M]
  vec fstar_primary[4], fstar_secondary[4];  // [nv, nf, M]
  vec alpha, alpha_tmp;      // [E]

  SolverBase(const SolverConfig& cfg_, Containers&& cc)
      : cfg(cfg_), basis(cfg_.polydeg), c(std::move(cc)) {
    nd = cfg.eq.ndim; N = basis.N; nv = cfg.eq.nvars;
    nn = 1; for (int d = 0; d < nd; ++d) nn *= N;
    nf = nn / N;
    stride[0] = 1; stride[1] = N; stride[2] = N * N;
    interfaces_u.assign((size_t)2 * nv * nf * c.ninterfaces, 0.0);
    surface_flux_values.assign((size_t)nv * nf * 2 * nd * c.nelements, 0.0);
    boundaries_u.assign((size_t)2 * nv * nf * c.nboundaries, 0.0);
    int nm = (nd == 3) ? 4 : (nd == 2 ? 2 : 0);
    for (int q = 0; q < nm; ++q) {
      mortar_u[q].assign((size_t)2 * nv * nf * c.nmortars, 0.0);
      fstar_primary[q].assign((size_t)nv * nf * c.nmortars, 0.0);
      fstar_secondary[q].assign((size_t)nv * nf * c.nmortars, 0.0);
    }
    alpha.assign(c.nelements, 0.0);
    alpha_tmp.assign(c.nelements, 0.0);
  }
  virtual ~SolverBase() {}

  size_t ndofs() const { return (size_t)nn * c.nelements; }
  size_t nunknowns() const { return (size_t)nv * nn * c.nelements; }

  // node index of face node f on face (dim d, position fixed)
  inline int face_node(int d, int fixed, int f) const {
    int a = f % N, b = f / N;
    if (nd == 1) return fixed;
    if (nd == 2) return d == 0 ? fixed + N * a : a + N * fixed;
    if (d == 0) return fixed + N * a + N * N * b;
    if (d == 1) return a + N * fixed + N * N * b;
    return a + N * b + N * N * fixed;
  }

  virtual void compute_coefficients(double t, double* u) const = 0;
  virtual void volume_integral(double* du, const double* u) = 0;
  virtual void calc_interface_flux() = 0;
  virtual void calc_boundary_flux(double t) = 0;
  virtual void calc_mortar_flux() = 0;
  virtual void calc_sources(double* du, const double* u, double t) const = 0;
  virtual void calc_indicator(const double* u) = 0;
  virtual double max_dt(const double* u) const = 0;
  virtual void calc_error_norms(const double* u, double t, double* l2, double* linf) const = 0;
  virtual void integrate_conserved(const double* u, double* out) const = 0;
  virtual double entropy_rate(const double* du, const double* u) const = 0;

  void reset_du(double* du) const { std::memset(du, 0, sizeof(double) * nunknowns()); }

  void prolong2interfaces(const double* u) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < c.ninterfaces; ++s) {
      int64_t left = c.if_neighbor_ids[2 * s] - 1, right = c.if_neighbor_ids[2 * s + 1] - 1;
      int d = (int)c.if_orientations[s] - 1;
      for (int f = 0; f < nf; ++f) {
        int nl = face_node(d, N - 1, f), nr = face_node(d, 0, f);
        for (int v = 0; v < nv; ++v) {
          interfaces_u[0 + 2 * (v + (size_t)nv * (f + (size_t)nf * s))] = u[v + (size_t)nv * (nl + (size_t)nn * left)];
          interfaces_u[1 + 2 * (v + (size_t)nv * (f + (size_t)nf * s))] = u[v + (size_t)nv * (nr + (size_t)nn * right)];
        }
      }
    }
  }

  void prolong2boundaries(const double* u) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t b = 0; b < c.nboundaries; ++b) {
      int64_t e = c.bd_neighbor_ids[b] - 1;
      int d = (int)c.bd_orientations[b] - 1;
      int side = (int)c.bd_neighbor_sides[b];
      for (int f = 0; f < nf; ++f) {
        int n = face_node(d, side == 1 ? N - 1 : 0, f);
        for (int v = 0; v < nv; ++v) {
          size_t o = 2 * (v + (size_t)nv * (f + (size_t)nf * b));
          // Trixi leaves the unused side uninitialised (NaN); the reference writes 0 there
          // (/root/reference/src/solvers/dg_3d_kernel.jl:1282-1291). Oracle keeps NaN; tests apply the NaN<->0 rule.
          boundaries_u[o + (side - 1)] = u[v + (size_t)nv * (n + (size_t)nn * e)];
          boundaries_u[o + (2 - side)] = nan;
        }
      }
    }
  }

  // out[v, a, b] = sum M1[a, aa] * M2[b, bb] * in[v, aa, bb]   (2D face);  1D face: out[v,a] = sum M[a,aa] in[v,aa]
  void face_apply(const Mat& M1, const Mat& M2, const double* in, double* out, bool add) const {
    std::vector<double> tmp((size_t)nv * nf);
    if (nd == 2) {
      for (int a = 0; a < N; ++a)
        for (int v = 0; v < nv; ++v) {
          double s = 0;
          for (int aa = 0; aa < N; ++aa) s += M1(a, aa) * in[v + nv * aa];
          if (add) out[v + nv * a] += s; else out[v + nv * a] = s;
        }
      return;
    }
    for (int b = 0; b < N; ++b)
      for (int a = 0; a < N; ++a)
        for (int v = 0; v < nv; ++v) {
          double s = 0;
          for (int aa = 0; aa < N; ++aa) s += M1(a, aa) * in[v + nv * (aa + N * b)];
          tmp[v + nv * (a + N * b)] = s;
        }
    for (int b = 0; b < N; ++b)
      for (int a = 0; a < N; ++a)
        for (int v = 0; v < nv; ++v) {
          double s = 0;
          for (int bb = 0; bb < N; ++bb) s += M2(b, bb) * tmp[v + nv * (a + N * bb)];
          if (add) out[v + nv * (a + N * b)] += s; else out[v + nv * (a + N * b)] = s;
        }
  }

  // mortar slot order: 3D storage q = 0 upper_left, 1 upper_right, 2 lower_left, 3 lower_right;
  // neighbor_ids rows (1-based): 1 lower_left, 2 lower_right, 3 upper_left, 4 upper_right, 5 large.
  // 2D storage q = 0 upper, 1 lower; neighbor_ids rows: 1 lower, 2 upper, 3 large.
  int mortar_small_row(int q) const {
    if (nd == 3) { static const int r[4] = {2, 3, 0, 1}; return r[q]; }
    static const int r2[2] = {1, 0};
    return r2[q];
  }

  void prolong2mortars(const double* u) {
    if (nd < 2) return;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    int nm = (nd == 3) ? 4 : 2, rows = nm + 1;
    for (int64_t m = 0; m < c.nmortars; ++m) {
      int d = (int)c.mo_orientations[m] - 1;
      int ls = (int)c.mo_large_sides[m];
      int64_t large = c.mo_neighbor_ids[rows * m + nm] - 1;
      // small to small: small elements on side (3 - ls)
      int small_side = 3 - ls;  // 1-based slot in dim 1 of u_*
      int fixed_small = (ls == 1) ? 0 : N - 1;
      for (int q = 0; q < nm; ++q) {
        int64_t e = c.mo_neighbor_ids[rows * m + mortar_small_row(q)] - 1;
        for (int f = 0; f < nf; ++f) {
          int n = face_node(d, fixed_small, f);
          for (int v = 0; v < nv; ++v)
            mortar_u[q][(small_side - 1) + 2 * (v + (size_t)nv * (f + (size_t)nf * m))] =
                u[v + (size_t)nv * (n + (size_t)nn * e)];
        }
      }
      // large to small
      int fixed_large = (ls == 1) ? N - 1 : 0;
      std::vector<double> ul((size_t)nv * nf), out((size_t)nv * nf);
      for (int f = 0; f < nf; ++f) {
        int n = face_node(d, fixed_large, f);
        for (int v = 0; v < nv; ++v) ul[v + nv * f] = u[v + (size_t)nv * (n + (size_t)nn * large)];
      }
      for (int q = 0; q < nm; ++q) {
        const Mat *M1, *M2;
        if (nd == 3) {
          // upper_left: (lower, upper); upper_right: (upper, upper); lower_left: (lower, lower); lower_right: (upper, lower)
          M1 = (q == 0 || q == 2) ? &basis.forward_lower : &basis.forward_upper;
          M2 = (q == 0 || q == 1) ? &basis.forward_upper : &basis.forward_lower;
        } else {
          M1 = (q == 0) ? &basis.forward_upper : &basis.forward_lower;
          M2 = M1;
        }
        face_apply(*M1, *M2, ul.data(), out.data(), false);
        for (int f = 0; f < nf; ++f)
          for (int v = 0; v < nv; ++v)
            mortar_u[q][(ls - 1) + 2 * (v + (size_t)nv * (f + (size_t)nf * m))] = out[v + nv * f];
      }
      (void)nan;
    }
  }

  void mortar_fluxes_to_elements() {
    int nm = (nd == 3) ? 4 : 2, rows = nm + 1;
    for (int64_t m = 0; m < c.nmortars; ++m) {
      int o = (int)c.mo_orientations[m];
      int ls = (int)c.mo_large_sides[m];
      int dir_small = 2 * o + ls - 2;   // 1-based direction
      int dir_large = 2 * o - ls + 1;
      for (int q = 0; q < nm; ++q) {
        int64_t e = c.mo_neighbor_ids[rows * m + mortar_small_row(q)] - 1;
        for (int f = 0; f < nf; ++f)
          for (int v = 0; v < nv; ++v)
            surface_flux_values[v + (size_t)nv * (f + (size_t)nf * ((dir_small - 1) + (size_t)2 * nd * e))] =
                fstar_primary[q][v + (size_t)nv * (f + (size_t)nf * m)];
      }
      int64_t large = c.mo_neighbor_ids[rows * m + nm] - 1;
      double* out = &surface_flux_values[(size_t)nv * nf * ((dir_large - 1) + (size_t)2 * nd * large)];
      for (int q = 0; q < nm; ++q) {
        const Mat *M1, *M2;
        if (nd == 3) {
          M1 = (q == 0 || q == 2) ? &basis.reverse_lower : &basis.reverse_upper;
          M2 = (q == 0 || q == 1) ? &basis.reverse_upper : &basis.reverse_lower;
        } else {
          M1 = (q == 0) ? &basis.reverse_upper : &basis.reverse_lower;
          M2 = M1;
        }
        face_apply(*M1, *M2, &fstar_secondary[q][(size_t)nv * nf * m], out, q != 0);
      }
    }
  }

  void surface_integral(double* du) const {
    double factor_1 = basis.boundary_interpolation(0, 0);
    double factor_2 = basis.boundary_interpolation(N - 1, 1);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < c.nelements; ++e)
      for (int f = 0; f < nf; ++f)
        for (int v = 0; v < nv; ++v)
          for (int d = 0; d < nd; ++d) {
            int n1 = face_node(d, 0, f), n2 = face_node(d, N - 1, f);
            const double* s = &surface_flux_values[(size_t)nv * nf * 2 * nd * e];
            du[v + (size_t)nv * (n1 + (size_t)nn * e)] -= s[v + nv * (f + nf * (2 * d))] * factor_1;
            du[v + (size_t)nv * (n2 + (size_t)nn * e)] += s[v + nv * (f + nf * (2 * d + 1))] * factor_2;
          }
  }

  void apply_jacobian(double* du) const {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < c.nelements; ++e) {
      double factor = -c.inverse_jacobian[e];
      for (int q = 0; q < nv * nn; ++q) du[q + (size_t)nv * nn * e] *= factor;
    }
  }

  void apply_smoothing() {
    alpha_tmp = alpha;
    for (int64_t s = 0; s < c.ninterfaces; ++s) {
      int64_t l = c.if_neighbor_ids[2 * s] - 1, r = c.if_neighbor_ids[2 * s + 1] - 1;
      alpha[l] = std::max(std::max(alpha_tmp[l], 0.5 * alpha_tmp[r]), alpha[l]);
      alpha[r] = std::max(std::max(alpha_tmp[r], 0.5 * alpha_tmp[l]), alpha[r]);
    }
    int nm = (nd == 3) ? 4 : (nd == 2 ? 2 : 0), rows = nm + 1;
    for (int64_t m = 0; m < c.nmortars; ++m) {
      int64_t large = c.mo_neighbor_ids[rows * m + nm] - 1;
      for (int q = 0; q < nm; ++q) {
        int64_t sm = c.mo_neighbor_ids[rows * m + q] - 1;
        alpha[sm] = std::max(std::max(alpha_tmp[sm], 0.5 * alpha_tmp[large]), alpha[sm]);
        alpha[large] = std::max(std::max(alpha_tmp[large], 0.5 * alpha_tmp[sm]), alpha[large]);
      }
    }
  }

  // Trixi.rhs!(du, u, t, mesh, equations, boundary_conditions, source_terms, dg, cache)
  void rhs(double* du, const double* u, double t) {
    reset_du(du);
    volume_integral(du, u);
    prolong2interfaces(u);
    calc_interface_flux();
    prolong2boundaries(u);
    calc_boundary_flux(t);
    prolong2mortars(u);
    calc_mortar_flux();
    surface_integral(du);
    apply_jacobian(du);
    calc_sources(du, u, t);
  }
};

// cons2prim_any / noncons shims (so generic code can ask any equation for (rho, ..., p))
template <class T> struct AdvectionX : Advection<T> {
  static void cons2prim_any(const T* u, const EqParams&, T* q) { for (int k = 0; k < 5; ++k) q[k] = u[0]; }
  static void noncons(const T*, const T*, int, const EqParams& p, T* g) { for (int v = 0; v < p.nvars; ++v) g[v] = 0; }
};
template <class T> struct EulerX : Euler<T> {
  static void cons2prim_any(const T* u, const EqParams& p, T* q) { Euler<T>::cons2prim(u, p, q); }
  static void noncons(const T*, const T*, int, const EqParams& p, T* g) { for (int v = 0; v < p.nvars; ++v) g[v] = 0; }
};
template <class T> struct MhdX : Mhd3D<T> {
  static void cons2prim_any(const T* u, const EqParams& p, T* q) { Mhd3D<T>::cons2prim(u, p, q); }
  static void noncons(const T* a, const T* b, int o, const EqParams& p, T* g) { Mhd3D<T>::noncons_powell(a, b, o, p, g); }
};

// -------------------------------------------------------------------------------------------------
template <template <class> class EqT> struct Solver : SolverBase {
  using Eq = EqT<double>;
  using SolverBase::SolverBase;

  void compute_coefficients(double t, double* u) const override {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < c.nelements; ++e)
      for (int n = 0; n < nn; ++n) {
        double x[3] = {0, 0, 0};
        for (int d = 0; d < nd; ++d) x[d] = c.node_coordinates[d + (size_t)nd * (n + (size_t)nn * e)];
        if (nd == 1) {  // Trixi nudges the end nodes inward by one ulp in 1D (dg_1d.jl compute_coefficients!;
                        // mirrored at /root/reference/src/solvers/dg.jl:54-58)
          if (n == 0) x[0] = std::nextafter(x[0], std::numeric_limits<double>::infinity());
          else if (n == N - 1) x[0] = std::nextafter(x[0], -std::numeric_limits<double>::infinity());
        }
        initial_condition(cfg.initial_condition, x, t, cfg.eq, &u[(size_t)nv * (n + (size_t)nn * e)]);
      }
  }

  void weak_form_kernel(double* du, const double* u, int64_t e) const {
    const double* ue = &u[(size_t)nv * nn * e];
    double* due = &du[(size_t)nv * nn * e];
    double f[9];
    for (int n = 0; n < nn; ++n) {
      int idx[3] = {n % N, (n / N) % N, n / (N * N)};
      for (int d = 0; d < nd; ++d) {
        Eq::flux(&ue[nv * n], d + 1, cfg.eq, f);
        int base = n - idx[d] * stride[d];
        for (int l = 0; l < N; ++l) {
          double w = basis.Dhat(l, idx[d]);
          double* t = &due[nv * (base + l * stride[d])];
          for (int v = 0; v < nv; ++v) t[v] += w * f[v];
        }
      }
    }
  }

  void flux_differencing_kernel(double* du, const double* u, int64_t e, double alpha_) const {
    const double* ue = &u[(size_t)nv * nn * e];
    double* due = &du[(size_t)nv * nn * e];
    double f[9];
    for (int n = 0; n < nn; ++n) {
      int idx[3] = {n % N, (n / N) % N, n / (N * N)};
      for (int d = 0; d < nd; ++d) {
        int base = n - idx[d] * stride[d];
        for (int l = idx[d] + 1; l < N; ++l) {
          int m = base + l * stride[d];
          Eq::two_point(cfg.volume_flux, &ue[nv * n], &ue[nv * m], d + 1, cfg.eq, f);
          double w1 = alpha_ * basis.Dsplit(idx[d], l), w2 = alpha_ * basis.Dsplit(l, idx[d]);
          for (int v = 0; v < nv; ++v) { due[nv * n + v] += w1 * f[v]; due[nv * m + v] += w2 * f[v]; }
        }
      }
    }
    if (cfg.nonconservative) {
      double g[9], acc[9];
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        for (int v = 0; v < nv; ++v) acc[v] = 0;
        for (int d = 0; d < nd; ++d) {
          int base = n - idx[d] * stride[d];
          for (int l = 0; l < N; ++l) {
            noncons(&ue[nv * n], &ue[nv * (base + l * stride[d])], d + 1, g);
            double w = basis.Dsplit(idx[d], l);
            for (int v = 0; v < nv; ++v) acc[v] += w * g[v];
          }
        }
        for (int v = 0; v < nv; ++v) due[nv * n + v] += alpha_ * 0.5 * acc[v];
      }
    }
  }

  void noncons(const double* a, const double* b, int o, double* g) const { Eq::noncons(a, b, o, cfg.eq, g); }

  void fv_kernel(double* du, const double* u, int64_t e, double alpha_) const {
    const double* ue = &u[(size_t)nv * nn * e];
    double* due = &du[(size_t)nv * nn * e];
    // fstar_d_{L,R}[v, node] with extent N+1 along d
    for (int d = 0; d < nd; ++d) {
      int ext[3] = {nd > 0 ? N : 1, nd > 1 ? N : 1, nd > 2 ? N : 1};
      ext[d] = N + 1;
      int st[3] = {1, ext[0], ext[0] * ext[1]};
      size_t tot = (size_t)ext[0] * ext[1] * ext[2];
      std::vector<double> fL(nv * tot, 0.0), fR(nv * tot, 0.0);
      double f[9], g[9];
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        if (idx[d] == 0) continue;
        int nl = n - stride[d];
        Eq::two_point(cfg.volume_flux_fv, &ue[nv * nl], &ue[nv * n], d + 1, cfg.eq, f);
        size_t q = idx[0] * st[0] + idx[1] * st[1] + idx[2] * st[2];
        for (int v = 0; v < nv; ++v) { fL[nv * q + v] = f[v]; fR[nv * q + v] = f[v]; }
        if (cfg.nonconservative) {
          noncons(&ue[nv * nl], &ue[nv * n], d + 1, g);
          for (int v = 0; v < nv; ++v) fL[nv * q + v] += 0.5 * g[v];
          noncons(&ue[nv * n], &ue[nv * nl], d + 1, g);
          for (int v = 0; v < nv; ++v) fR[nv * q + v] += 0.5 * g[v];
        }
      }
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        size_t q = idx[0] * st[0] + idx[1] * st[1] + idx[2] * st[2];
        double iw = basis.inverse_weights[idx[d]];
        for (int v = 0; v < nv; ++v)
          due[nv * n + v] += alpha_ * (iw * (fL[nv * (q + st[d]) + v] - fR[nv * q + v]));
      }
    }
  }

  void calc_indicator(const double* u) override {
    // /root/reference/src/solvers/indicators.jl:21-24 (constants) ; arithmetic = Trixi
    const double threshold = 0.5 * std::pow(10.0, -1.8 * std::pow((double)N, 0.25));
    const double parameter_s = std::log((1 - 0.0001) / 0.0001);
    const Mat& V = basis.inverse_vandermonde_legendre;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < c.nelements; ++e) {
      std::vector<double> ind(nn), modal(nn), tmp(nn);
      for (int n = 0; n < nn; ++n) {
        const double* un = &u[(size_t)nv * (n + (size_t)nn * e)];
        double q[9];
        Eq::cons2prim_any(un, cfg.eq, q);
        double rho = q[0], pr = q[cfg.eq.kind == EQ_MHD ? 4 : nd + 1];
        ind[n] = cfg.indicator_variable == IND_DENSITY ? rho
               : cfg.indicator_variable == IND_PRESSURE ? pr : rho * pr;
      }
      // multiply_scalar_dimensionwise!: apply V along each dim
      modal = ind;
      for (int d = 0; d < nd; ++d) {
        for (int n = 0; n < nn; ++n) {
          int id = (n / stride[d]) % N;
          int base = n - id * stride[d];
          double s = 0;
          for (int l = 0; l < N; ++l) s += V(id, l) * modal[base + l * stride[d]];
          tmp[n] = s;
        }
        modal = tmp;
      }
      double total = 0, clip1 = 0, clip2 = 0;
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        int mx = 0;
        for (int d = 0; d < nd; ++d) mx = std::max(mx, idx[d]);
        double m2 = modal[n] * modal[n];
        total += m2;
        if (mx < N - 1) clip1 += m2;
        if (mx < N - 2) clip2 += m2;
      }
      double f1 = (total != 0.0) ? (total - clip1) / total : 0.0;
      double f2 = (clip1 != 0.0) ? (clip1 - clip2) / clip1 : 0.0;
      double energy = std::max(f1, f2);
      double a = 1 / (1 + std::exp(-parameter_s / threshold * (energy - threshold)));
      if (a < cfg.alpha_min) a = 0;
      if (a > 1 - cfg.alpha_min) a = 1;
      alpha[e] = std::min(cfg.alpha_max, a);
    }
    if (cfg.alpha_smooth) apply_smoothing();
  }

  void volume_integral(double* du, const double* u) override {
    if (cfg.volume_integral == VI_WEAK_FORM) {
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < c.nelements; ++e) weak_form_kernel(du, u, e);
    } else if (cfg.volume_integral == VI_FLUX_DIFFERENCING) {
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < c.nelements; ++e) flux_differencing_kernel(du, u, e, 1.0);
    } else {
      calc_indicator(u);
      // /root/reference/src/solvers/dg_3d.jl:189
      const double eps = std::numeric_limits<double>::epsilon();
      const double atol = std::max(100 * eps, std::pow(eps, 0.75));
#pragma omp parallel for schedule(dynamic, 16)
      for (int64_t e = 0; e < c.nelements; ++e) {
        double a = alpha[e];
        bool dg_only = std::fabs(a) <= atol;
        if (dg_only) flux_differencing_kernel(du, u, e, 1.0);
        else {
          flux_differencing_kernel(du, u, e, 1 - a);
          fv_kernel(du, u, e, a);
        }
      }
    }
  }

  void calc_interface_flux() override {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < c.ninterfaces; ++s) {
      int64_t left = c.if_neighbor_ids[2 * s] - 1, right = c.if_neighbor_ids[2 * s + 1] - 1;
      int o = (int)c.if_orientations[s];
      int ldir = 2 * o - 1, rdir = 2 * o - 2;  // 0-based directions (2o, 2o-1 in Julia)
      double ul[9], ur[9], f[9], gl[9], gr[9];
      for (int fn = 0; fn < nf; ++fn) {
        for (int v = 0; v < nv; ++v) {
          ul[v] = interfaces_u[0 + 2 * (v + (size_t)nv * (fn + (size_t)nf * s))];
          ur[v] = interfaces_u[1 + 2 * (v + (size_t)nv * (fn + (size_t)nf * s))];
        }
        Eq::two_point(cfg.surface_flux, ul, ur, o, cfg.eq, f);
        double* sl = &surface_flux_values[(size_t)nv * (fn + (size_t)nf * (ldir + (size_t)2 * nd * left))];
        double* sr = &surface_flux_values[(size_t)nv * (fn + (size_t)nf * (rdir + (size_t)2 * nd * right))];
        if (cfg.nonconservative) {
          noncons(ul, ur, o, gl);
          noncons(ur, ul, o, gr);
          for (int v = 0; v < nv; ++v) { sl[v] = f[v] + 0.5 * gl[v]; sr[v] = f[v] + 0.5 * gr[v]; }
        } else {
          for (int v = 0; v < nv; ++v) { sl[v] = f[v]; sr[v] = f[v]; }
        }
      }
    }
  }

  void calc_boundary_flux(double t) override {
    int64_t first = 0;
    for (int dir = 1; dir <= 2 * nd; ++dir) {
      int64_t nb = c.n_boundaries_per_direction[dir - 1];
      for (int64_t b = first; b < first + nb; ++b) {
        if (cfg.bc[dir - 1] == BC_PERIODIC) continue;
        int64_t e = c.bd_neighbor_ids[b] - 1;
        int side = (int)c.bd_neighbor_sides[b];
        int o = (int)c.bd_orientations[b];
        double ui[9], ub[9], f[9];
        for (int fn = 0; fn < nf; ++fn) {
          for (int v = 0; v < nv; ++v) ui[v] = boundaries_u[(side - 1) + 2 * (v + (size_t)nv * (fn + (size_t)nf * b))];
          double x[3] = {0, 0, 0};
          for (int d = 0; d < nd; ++d) x[d] = c.bd_node_coordinates[d + (size_t)nd * (fn + (size_t)nf * b)];
          if (cfg.bc[dir - 1] == BC_SLIP_WALL) {
            Eq::slip_wall_flux(ui, o, dir, cfg.eq, f);   // boundary_condition_slip_wall (compressible Euler)
          } else {
            initial_condition(cfg.initial_condition, x, t, cfg.eq, ub);  // BoundaryConditionDirichlet(ic)
            if (dir % 2 == 0) Eq::two_point(cfg.surface_flux, ui, ub, o, cfg.eq, f);
            else Eq::two_point(cfg.surface_flux, ub, ui, o, cfg.eq, f);
          }
          double* s = &surface_flux_values[(size_t)nv * (fn + (size_t)nf * ((dir - 1) + (size_t)2 * nd * e))];
          for (int v = 0; v < nv; ++v) s[v] = f[v];
        }
      }
      first += nb;
    }
  }

  void calc_mortar_flux() override {
    if (nd < 2) return;
    int nm = (nd == 3) ? 4 : 2;
    for (int64_t m = 0; m < c.nmortars; ++m) {
      int o = (int)c.mo_orientations[m];
      int ls = (int)c.mo_large_sides[m];
      double ul[9], ur[9], f[9], gp[9], gs[9];
      for (int q = 0; q < nm; ++q)
        for (int fn = 0; fn < nf; ++fn) {
          for (int v = 0; v < nv; ++v) {
            ul[v] = mortar_u[q][0 + 2 * (v + (size_t)nv * (fn + (size_t)nf * m))];
            ur[v] = mortar_u[q][1 + 2 * (v + (size_t)nv * (fn + (size_t)nf * m))];
          }
          Eq::two_point(cfg.surface_flux, ul, ur, o, cfg.eq, f);
          double* fp = &fstar_primary[q][(size_t)nv * (fn + (size_t)nf * m)];
          double* fs = &fstar_secondary[q][(size_t)nv * (fn + (size_t)nf * m)];
          for (int v = 0; v < nv; ++v) { fp[v] = f[v]; fs[v] = f[v]; }
          if (cfg.nonconservative) {
            // primary: (large-side state, small-side state); secondary reversed
            // (/root/reference/src/solvers/dg_3d_kernel.jl:1598-1625)
            const double* u1 = (ls == 1) ? ul : ur;
            const double* u2 = (ls == 1) ? ur : ul;
            noncons(u1, u2, o, gp);
            noncons(u2, u1, o, gs);
            for (int v = 0; v < nv; ++v) { fp[v] += 0.5 * gp[v]; fs[v] += 0.5 * gs[v]; }
          }
        }
    }
    mortar_fluxes_to_elements();
  }

  void calc_sources(double* du, const double* u, double t) const override {
    if (cfg.source == SRC_NONE) return;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < c.nelements; ++e)
      for (int n = 0; n < nn; ++n) {
        double x[3] = {0, 0, 0}, s[9];
        for (int d = 0; d < nd; ++d) x[d] = c.node_coordinates[d + (size_t)nd * (n + (size_t)nn * e)];
        size_t o = (size_t)nv * (n + (size_t)nn * e);
        source_terms(cfg.source, &u[o], x, t, cfg.eq, s);
        for (int v = 0; v < nv; ++v) du[o + v] += s[v];
      }
  }

  double max_dt(const double* u) const override {
    double max_scaled_speed = std::numeric_limits<double>::denorm_min();  // nextfloat(0.0)
#pragma omp parallel for schedule(static) reduction(max : max_scaled_speed)
    for (int64_t e = 0; e < c.nelements; ++e) {
      double ml[3] = {0, 0, 0}, lam[3];
      for (int n = 0; n < nn; ++n) {
        Eq::max_abs_speeds(&u[(size_t)nv * (n + (size_t)nn * e)], cfg.eq, lam);
        for (int d = 0; d < nd; ++d) ml[d] = std::max(ml[d], lam[d]);
      }
      double s = 0;
      for (int d = 0; d < nd; ++d) s += ml[d];
      max_scaled_speed = std::max(max_scaled_speed, c.inverse_jacobian[e] * s);
    }
    return 2 / (N * max_scaled_speed);
  }

  // interpolate one element's nodal field (ncomp components, component-fastest) to the analysis nodes
  void to_analysis(const double* in, int ncomp, std::vector<double>& out) const {
    const int NA = basis.NA;
    int cur[3] = {nd > 0 ? N : 1, nd > 1 ? N : 1, nd > 2 ? N : 1};
    std::vector<double> a(in, in + (size_t)ncomp * nn), b;
    for (int d = 0; d < nd; ++d) {
      int nxt[3] = {cur[0], cur[1], cur[2]};
      nxt[d] = NA;
      b.assign((size_t)ncomp * nxt[0] * nxt[1] * nxt[2], 0.0);
      for (int k = 0; k < nxt[2]; ++k) for (int j = 0; j < nxt[1]; ++j) for (int i = 0; i < nxt[0]; ++i) {
        int id[3] = {i, j, k};
        for (int cc = 0; cc < ncomp; ++cc) {
          double s = 0;
          for (int l = 0; l < N; ++l) {
            int is[3] = {i, j, k};
            is[d] = l;
            s += basis.analysis_vandermonde(id[d], l) * a[cc + (size_t)ncomp * (is[0] + cur[0] * (is[1] + (size_t)cur[1] * is[2]))];
          }
          b[cc + (size_t)ncomp * (i + nxt[0] * (j + (size_t)nxt[1] * k))] = s;
        }
      }
      a.swap(b);
      cur[d] = NA;
    }
    out.swap(a);
  }

  void calc_error_norms(const double* u, double t, double* l2, double* linf) const override {
    const int NA = basis.NA;
    int na = 1; for (int d = 0; d < nd; ++d) na *= NA;
    for (int v = 0; v < nv; ++v) { l2[v] = 0; linf[v] = 0; }
    std::vector<double> ua, xa;
    for (int64_t e = 0; e < c.nelements; ++e) {
      to_analysis(&u[(size_t)nv * nn * e], nv, ua);
      to_analysis(&c.node_coordinates[(size_t)nd * nn * e], nd, xa);
      double vj = std::pow(1.0 / c.inverse_jacobian[e], nd);
      for (int q = 0; q < na; ++q) {
        int idx[3] = {q % NA, (q / NA) % NA, q / (NA * NA)};
        double w = 1;
        for (int d = 0; d < nd; ++d) w *= basis.analysis_weights[idx[d]];
        double x[3] = {0, 0, 0}, ue[9];
        for (int d = 0; d < nd; ++d) x[d] = xa[d + (size_t)nd * q];
        initial_condition(cfg.initial_condition, x, t, cfg.eq, ue);
        for (int v = 0; v < nv; ++v) {
          double diff = ue[v] - ua[v + (size_t)nv * q];
          l2[v] += diff * diff * (w * vj);
          linf[v] = std::max(linf[v], std::fabs(diff));
        }
      }
    }
    // total_volume(mesh) = length_level_0^ndims
    double lvl0 = 2.0 / c.inverse_jacobian[0] * double(1L << c.cell_levels[0]);
    double total_volume = std::pow(lvl0, nd);
    for (int v = 0; v < nv; ++v) l2[v] = std::sqrt(l2[v] / total_volume);
  }

  void integrate_conserved(const double* u, double* out) const override {
    for (int v = 0; v < nv; ++v) out[v] = 0;
    for (int64_t e = 0; e < c.nelements; ++e) {
      double vj = std::pow(1.0 / c.inverse_jacobian[e], nd);
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        double w = vj;
        for (int d = 0; d < nd; ++d) w *= basis.weights[idx[d]];
        for (int v = 0; v < nv; ++v) out[v] += w * u[(size_t)nv * (n + (size_t)nn * e) + v];
      }
    }
  }

  // d/dt of total mathematical entropy: sum_e J_e^nd sum_n w_n  w(u_n) . du_n   (Euler only; else 0)
  double entropy_rate(const double* du, const double* u) const override {
    if (cfg.eq.kind != EQ_EULER) return 0.0;
    double total = 0;
    for (int64_t e = 0; e < c.nelements; ++e) {
      double vj = std::pow(1.0 / c.inverse_jacobian[e], nd);
      for (int n = 0; n < nn; ++n) {
        int idx[3] = {n % N, (n / N) % N, n / (N * N)};
        double w = vj;
        for (int d = 0; d < nd; ++d) w *= basis.weights[idx[d]];
        const double* un = &u[(size_t)nv * (n + (size_t)nn * e)];
        const double* dun = &du[(size_t)nv * (n + (size_t)nn * e)];
        double q[9];
        Eq::cons2prim_any(un, cfg.eq, q);
        double rho = q[0], p = q[nd + 1], g = cfg.eq.gamma;
        double v2 = 0;
        for (int d = 0; d < nd; ++d) v2 += q[1 + d] * q[1 + d];
        double s = std::log(p) - g * std::log(rho);
        double rho_p = rho / p;
        double wv[5];
        wv[0] = (g - s) / (g - 1) - 0.5 * rho_p * v2;
        for (int d = 0; d < nd; ++d) wv[1 + d] = rho_p * q[1 + d];
        wv[nd + 1] = -rho_p;
        double dot = 0;
        for (int v = 0; v < nv; ++v) dot += wv[v] * dun[v];
        total += w * dot;
      }
    }
    return total;
  }
};

}  // namespace orc
