// TEST INFRASTRUCTURE ONLY -- C ABI of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs. Never linked or imported by the product library.
// "parity unpinned": see basis.hpp header and ORACLE_ASSUMPTIONS.md.
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dg.hpp"

using namespace orc;

extern "C" {

struct orc_config {
  int ndim, eq_kind, polydeg;
  int volume_integral, volume_flux, volume_flux_fv, surface_flux, nonconservative;
  int alpha_smooth, indicator_variable, initial_condition, source;
  int bc[6];
  int initial_refinement_level;
  int periodicity[3];
  int n_patches;
  double gamma, advection_velocity[3], c_h, alpha_max, alpha_min;
  double coordinates_min[3], coordinates_max[3];
  double patch_lo[4][3], patch_hi[4][3];
};

static thread_local std::string g_err;
const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(const orc_config* cc) {
  try {
    SolverConfig cfg;
    cfg.eq.kind = cc->eq_kind;
    cfg.eq.ndim = cc->ndim;
    cfg.eq.nvars = cc->eq_kind == EQ_ADVECTION ? 1 : (cc->eq_kind == EQ_EULER ? cc->ndim + 2 : 9);
    cfg.eq.gamma = cc->gamma;
    for (int d = 0; d < 3; ++d) cfg.eq.advection_velocity[d] = cc->advection_velocity[d];
    cfg.eq.c_h = cc->c_h;
    cfg.polydeg = cc->polydeg;
    cfg.volume_integral = cc->volume_integral;
    cfg.volume_flux = cc->volume_flux;
    cfg.volume_flux_fv = cc->volume_flux_fv;
    cfg.surface_flux = cc->surface_flux;
    cfg.nonconservative = cc->nonconservative;
    cfg.alpha_max = cc->alpha_max; cfg.alpha_min = cc->alpha_min;
    cfg.alpha_smooth = cc->alpha_smooth;
    cfg.indicator_variable = cc->indicator_variable;
    for (int d = 0; d < 6; ++d) cfg.bc[d] = cc->bc[d];
    cfg.initial_condition = cc->initial_condition;
    cfg.source = cc->source;
    if (cc->eq_kind == EQ_MHD && cc->ndim != 3) throw std::runtime_error("oracle: GLM-MHD only in 3D");

    Tree t;
    bool per[3] = {cc->periodicity[0] != 0, cc->periodicity[1] != 0, cc->periodicity[2] != 0};
    t.init(cc->ndim, cc->coordinates_min, cc->coordinates_max, per);
    t.refine_uniform(cc->initial_refinement_level);
    for (int p = 0; p < cc->n_patches; ++p) {
      RefinementBox b;
      for (int d = 0; d < 3; ++d) { b.lo[d] = cc->patch_lo[p][d]; b.hi[d] = cc->patch_hi[p][d]; }
      t.refine_box(b);
    }
    Basis basis(cc->polydeg);
    Containers c = build_containers(t, basis);
    SolverBase* s = nullptr;
    if (cc->eq_kind == EQ_ADVECTION) s = new Solver<AdvectionX>(cfg, std::move(c));
    else if (cc->eq_kind == EQ_EULER) s = new Solver<EulerX>(cfg, std::move(c));
    else s = new Solver<MhdX>(cfg, std::move(c));
    return s;
  } catch (std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void orc_destroy(void* h) { delete (SolverBase*)h; }

long long orc_size(void* h, const char* name) {
  SolverBase* s = (SolverBase*)h;
  std::string n(name);
  if (n == "nelements") return s->c.nelements;
  if (n == "ninterfaces") return s->c.ninterfaces;
  if (n == "nboundaries") return s->c.nboundaries;
  if (n == "nmortars") return s->c.nmortars;
  if (n == "nvars") return s->nv;
  if (n == "nnodes") return s->N;
  if (n == "ndofs") return (long long)s->ndofs();
  if (n == "nunknowns") return (long long)s->nunknowns();
  if (n == "nanalysis") return s->basis.NA;
  return -1;
}

static const vec* f64_array(SolverBase* s, const std::string& n) {
  Basis& b = s->basis;
  if (n == "nodes") return &b.nodes;
  if (n == "weights") return &b.weights;
  if (n == "inverse_weights") return &b.inverse_weights;
  if (n == "derivative_matrix") return &b.D.a;
  if (n == "derivative_dhat") return &b.Dhat.a;
  if (n == "derivative_split") return &b.Dsplit.a;
  if (n == "derivative_split_transpose") return &b.Dsplit_transpose.a;
  if (n == "boundary_interpolation") return &b.boundary_interpolation.a;
  if (n == "inverse_vandermonde_legendre") return &b.inverse_vandermonde_legendre.a;
  if (n == "forward_upper") return &b.forward_upper.a;
  if (n == "forward_lower") return &b.forward_lower.a;
  if (n == "reverse_upper") return &b.reverse_upper.a;
  if (n == "reverse_lower") return &b.reverse_lower.a;
  if (n == "analysis_nodes") return &b.analysis_nodes;
  if (n == "analysis_weights") return &b.analysis_weights;
  if (n == "analysis_vandermonde") return &b.analysis_vandermonde.a;
  if (n == "inverse_jacobian") return &s->c.inverse_jacobian;
  if (n == "node_coordinates") return &s->c.node_coordinates;
  if (n == "cell_centers") return &s->c.cell_centers;
  if (n == "boundaries.node_coordinates") return &s->c.bd_node_coordinates;
  if (n == "interfaces.u") return &s->interfaces_u;
  if (n == "boundaries.u") return &s->boundaries_u;
  if (n == "surface_flux_values") return &s->surface_flux_values;
  if (n == "alpha") return &s->alpha;
  const char* mu3[4] = {"mortars.u_upper_left", "mortars.u_upper_right", "mortars.u_lower_left", "mortars.u_lower_right"};
  const char* mu2[2] = {"mortars.u_upper", "mortars.u_lower"};
  for (int q = 0; q < 4; ++q) if (s->nd == 3 && n == mu3[q]) return &s->mortar_u[q];
  for (int q = 0; q < 2; ++q) if (s->nd == 2 && n == mu2[q]) return &s->mortar_u[q];
  return nullptr;
}

static const std::vector<int64_t>* i64_array(SolverBase* s, const std::string& n) {
  Containers& c = s->c;
  if (n == "cell_levels") return &c.cell_levels;
  if (n == "cell_icoords") return &c.cell_icoords;
  if (n == "interfaces.neighbor_ids") return &c.if_neighbor_ids;
  if (n == "interfaces.orientations") return &c.if_orientations;
  if (n == "boundaries.neighbor_ids") return &c.bd_neighbor_ids;
  if (n == "boundaries.orientations") return &c.bd_orientations;
  if (n == "boundaries.neighbor_sides") return &c.bd_neighbor_sides;
  if (n == "boundaries.n_boundaries_per_direction") return &c.n_boundaries_per_direction;
  if (n == "mortars.neighbor_ids") return &c.mo_neighbor_ids;
  if (n == "mortars.large_sides") return &c.mo_large_sides;
  if (n == "mortars.orientations") return &c.mo_orientations;
  return nullptr;
}

long long orc_len_f64(void* h, const char* name) {
  const vec* a = f64_array((SolverBase*)h, name);
  return a ? (long long)a->size() : -1;
}
long long orc_len_i64(void* h, const char* name) {
  auto* a = i64_array((SolverBase*)h, name);
  return a ? (long long)a->size() : -1;
}
long long orc_get_f64(void* h, const char* name, double* out, long long n) {
  const vec* a = f64_array((SolverBase*)h, name);
  if (!a || (long long)a->size() != n) return -1;
  std::memcpy(out, a->data(), sizeof(double) * n);
  return n;
}
long long orc_get_i64(void* h, const char* name, long long* out, long long n) {
  auto* a = i64_array((SolverBase*)h, name);
  if (!a || (long long)a->size() != n) return -1;
  std::memcpy(out, a->data(), sizeof(int64_t) * n);
  return n;
}

void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_compute_coefficients(void* h, double t, double* u) { ((SolverBase*)h)->compute_coefficients(t, u); }
void orc_rhs(void* h, double* du, const double* u, double t) { ((SolverBase*)h)->rhs(du, u, t); }

// one stage at a time, same names as the reference's per-stage tests
int orc_stage(void* h, const char* stage, double* du, const double* u, double t) {
  SolverBase* s = (SolverBase*)h;
  std::string n(stage);
  if (n == "reset_du") s->reset_du(du);
  else if (n == "calc_volume_integral") s->volume_integral(du, u);
  else if (n == "prolong2interfaces") s->prolong2interfaces(u);
  else if (n == "calc_interface_flux") s->calc_interface_flux();
  else if (n == "prolong2boundaries") s->prolong2boundaries(u);
  else if (n == "calc_boundary_flux") s->calc_boundary_flux(t);
  else if (n == "prolong2mortars") s->prolong2mortars(u);
  else if (n == "calc_mortar_flux") s->calc_mortar_flux();
  else if (n == "calc_surface_integral") s->surface_integral(du);
  else if (n == "apply_jacobian") s->apply_jacobian(du);
  else if (n == "calc_sources") s->calc_sources(du, u, t);
  else if (n == "calc_indicator") s->calc_indicator(u);
  else return -1;
  return 0;
}

double orc_max_dt(void* h, const double* u) { return ((SolverBase*)h)->max_dt(u); }
void orc_error_norms(void* h, const double* u, double t, double* l2, double* linf) {
  ((SolverBase*)h)->calc_error_norms(u, t, l2, linf);
}
void orc_integrate(void* h, const double* u, double* out) { ((SolverBase*)h)->integrate_conserved(u, out); }
double orc_entropy_rate(void* h, const double* du, const double* u) { return ((SolverBase*)h)->entropy_rate(du, u); }

// CarpenterKennedy2N54 with StepsizeCallback(cfl): dt = cfl * max_dt(u) before every step, last step clipped.
// Returns the number of steps taken.
long long orc_solve_ck2n54(void* h, double* u, double t0, double t1, double cfl, double fixed_dt, long long max_steps) {
  SolverBase* s = (SolverBase*)h;
  static const double A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                              -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
  static const double B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                              1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                              2277821191437.0 / 14882151754819.0};
  static const double C[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                              2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
  size_t n = s->nunknowns();
  std::vector<double> du(n), tmp(n, 0.0);
  double t = t0;
  long long steps = 0;
  while (t < t1 && steps < max_steps) {
    double dt = fixed_dt > 0 ? fixed_dt : cfl * s->max_dt(u);
    if (t + dt > t1 || std::fabs(t + dt - t1) < 100 * 2.2e-16 * std::max(1.0, std::fabs(t1))) dt = t1 - t;
    for (int st = 0; st < 5; ++st) {
      s->rhs(du.data(), u, t + C[st] * dt);
      double a = A[st], b = B[st];
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < (long long)n; ++i) {
        tmp[i] = a * tmp[i] + dt * du[i];
        u[i] += b * tmp[i];
      }
    }
    t += dt;
    ++steps;
  }
  return steps;
}

// two-point flux of any enumerated kind on raw states (unit tests of the flux restatements)
int orc_two_point_flux(int eq_kind, int ndim, double gamma, const double* adv, double c_h, int flux_kind,
                       const double* ul, const double* ur, int orientation, double* f) {
  EqParams p;
  p.kind = eq_kind; p.ndim = ndim; p.gamma = gamma; p.c_h = c_h;
  p.nvars = eq_kind == EQ_ADVECTION ? 1 : (eq_kind == EQ_EULER ? ndim + 2 : 9);
  for (int d = 0; d < 3; ++d) p.advection_velocity[d] = adv[d];
  bool ok;
  if (eq_kind == EQ_ADVECTION) ok = Advection<double>::two_point(flux_kind, ul, ur, orientation, p, f);
  else if (eq_kind == EQ_EULER) ok = Euler<double>::two_point(flux_kind, ul, ur, orientation, p, f);
  else ok = Mhd3D<double>::two_point(flux_kind, ul, ur, orientation, p, f);
  return ok ? 0 : -1;
}
int orc_noncons_powell(double gamma, double c_h, const double* ul, const double* ur, int orientation, double* f) {
  EqParams p; p.kind = EQ_MHD; p.ndim = 3; p.nvars = 9; p.gamma = gamma; p.c_h = c_h;
  Mhd3D<double>::noncons_powell(ul, ur, orientation, p, f);
  return 0;
}

// Timed rhs! loop for the CPU baseline: returns seconds for `reps` calls after `warm` warm-ups.
double orc_time_rhs(void* h, double* du, const double* u, double t, int warm, int reps) {
  SolverBase* s = (SolverBase*)h;
  for (int i = 0; i < warm; ++i) s->rhs(du, u, t);
  auto a = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) s->rhs(du, u, t);
  auto b = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(b - a).count();
}

}  // extern "C"
