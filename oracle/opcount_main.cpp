// TEST INFRASTRUCTURE ONLY -- counts the floating-point operations of ONE oracle rhs! per DOF for the five BASELINE.json
// configurations ("exact OpCount<double> figure from the oracle"). The oracle's headers are
// compiled a second time with `double` replaced by the instrumented scalar of opcount.hpp, so the counted algorithm is
// the oracle's own (Trixi.jl's CPU rhs!: symmetric flux differencing, every interface flux once). Single-threaded.
//   make -C oracle opcount && oracle/_build/opcount > profiles/r2_opcount.json
// Convention of SURVEY.md section 8(d): add = sub = mul = div = sqrt = log = exp = pow = 1 flop; the oracle is compiled
// with -ffp-contract=off, so there are no fused multiply-adds to count as 2; comparisons / abs / negation are listed
// but not counted as flops.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <array>
#include <map>
#include <cmath>
#include <limits>
#include <algorithm>
#include <stdexcept>
#include <memory>
#include <numeric>
#include <functional>
#include <cstdint>
#include <cassert>
#include "opcount.hpp"
#define double orc::OpCountD
#include "dg.hpp"
#undef double

using namespace orc;

struct Case {
  const char* key; int ndim, eq, level, vi, vflux, fvflux, sflux, noncons, ic; double gamma, c_h, lo, hi; bool patch;
};

static void run(const Case& k, bool last) {
  SolverConfig cfg;
  cfg.eq.kind = k.eq; cfg.eq.ndim = k.ndim;
  cfg.eq.nvars = k.eq == EQ_ADVECTION ? 1 : (k.eq == EQ_EULER ? k.ndim + 2 : 9);
  cfg.eq.gamma = k.gamma; cfg.eq.c_h = k.c_h;
  cfg.eq.advection_velocity[0] = 1.0; cfg.eq.advection_velocity[1] = 0.0; cfg.eq.advection_velocity[2] = 0.0;
  cfg.polydeg = 3;
  cfg.volume_integral = k.vi; cfg.volume_flux = k.vflux; cfg.volume_flux_fv = k.fvflux; cfg.surface_flux = k.sflux;
  cfg.nonconservative = k.noncons;
  cfg.alpha_max = 0.5; cfg.alpha_min = 0.001; cfg.alpha_smooth = 1; cfg.indicator_variable = 2;
  for (int d = 0; d < 6; ++d) cfg.bc[d] = 0;
  cfg.initial_condition = k.ic; cfg.source = SRC_NONE;
  Tree t;
  bool per[3] = {true, true, true};
  OpCountD cmin[3] = {k.lo, k.lo, k.lo}, cmax[3] = {k.hi, k.hi, k.hi};
  t.init(k.ndim, cmin, cmax, per);
  t.refine_uniform(k.level);
  if (k.patch) {
    RefinementBox b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = -0.5; b.hi[d] = 0.5; }
    t.refine_box(b);
  }
  Basis basis(3);
  Containers c = build_containers(t, basis);
  std::unique_ptr<SolverBase> s;
  if (k.eq == EQ_ADVECTION) s.reset(new Solver<AdvectionX>(cfg, std::move(c)));
  else if (k.eq == EQ_EULER) s.reset(new Solver<EulerX>(cfg, std::move(c)));
  else s.reset(new Solver<MhdX>(cfg, std::move(c)));
  std::vector<OpCountD> u(s->nunknowns()), du(s->nunknowns());
  s->compute_coefficients(OpCountD(0.0), u.data());
  opc() = OpCounters();
  s->rhs(du.data(), u.data(), OpCountD(0.0));
  const OpCounters n = opc();
  const double dof = (double)s->ndofs();
  const double flop = (double)(n.add + n.mul + n.div + n.sqrt_ + n.log_ + n.exp_ + n.pow_ + n.trig);
  std::printf("  \"%s\": {\"flop_per_dof\": %.3f, \"add_sub\": %.3f, \"mul\": %.3f, \"div\": %.3f, \"sqrt\": %.3f, \"log\": %.3f, "
              "\"exp\": %.3f, \"pow\": %.3f, \"compare\": %.3f, \"abs_neg\": %.3f, \"nelements\": %lld, \"level\": %d}%s\n",
              k.key, flop / dof, n.add / dof, n.mul / dof, n.div / dof, n.sqrt_ / dof, n.log_ / dof, n.exp_ / dof,
              n.pow_ / dof, n.cmp / dof, n.absneg / dof, (long long)s->c.nelements, k.level, last ? "" : ",");
}

int main() {
  const Case cases[] = {
      {"c1_advection_1d", 1, EQ_ADVECTION, 4, 0, FLUX_CENTRAL, FLUX_LAX_FRIEDRICHS, FLUX_LAX_FRIEDRICHS, 0, IC_CONVERGENCE_TEST, 1.4, 1.0, -1.0, 1.0, false},
      {"c2_euler_ec_2d", 2, EQ_EULER, 4, 1, FLUX_RANOCHA, FLUX_LAX_FRIEDRICHS, FLUX_RANOCHA, 0, IC_WEAK_BLAST_WAVE, 1.4, 1.0, -2.0, 2.0, false},
      {"c3_euler_sc_3d", 3, EQ_EULER, 3, 2, FLUX_RANOCHA, FLUX_RANOCHA, FLUX_RANOCHA, 0, IC_WEAK_BLAST_WAVE, 1.4, 1.0, -2.0, 2.0, false},
      {"c4_mhd_alfven_mortar_3d", 3, EQ_MHD, 2, 1, FLUX_HINDENLANG_GASSNER, FLUX_LAX_FRIEDRICHS, FLUX_HLLE, 1, IC_CONVERGENCE_TEST, 5.0 / 3, 1.3, -1.0, 1.0, true},
      {"c5_euler_ec_3d", 3, EQ_EULER, 3, 1, FLUX_RANOCHA, FLUX_LAX_FRIEDRICHS, FLUX_RANOCHA, 0, IC_WEAK_BLAST_WAVE, 1.4, 1.0, -2.0, 2.0, false},
      {"c5_euler_ec_3d_smooth", 3, EQ_EULER, 3, 1, FLUX_RANOCHA, FLUX_LAX_FRIEDRICHS, FLUX_RANOCHA, 0, IC_DENSITY_WAVE, 1.4, 1.0, -2.0, 2.0, false},
  };
  std::printf("{\n  \"_comment\": \"floating-point operations of ONE oracle rhs! per DOF (DOF = node of one field), counted by "
              "oracle/opcount_main.cpp with the instrumented scalar oracle/opcount.hpp; flop = add/sub + mul + div + sqrt + log + "
              "exp + pow (1 each, no FMA contraction in the oracle); per-DOF counts are level-independent on periodic uniform "
              "meshes; C3 depends on how many elements the indicator blends (weak blast wave at the counted level)\",\n");
  const int n = sizeof(cases) / sizeof(cases[0]);
  for (int i = 0; i < n; ++i) run(cases[i], i == n - 1);
  std::printf("}\n");
  return 0;
}
