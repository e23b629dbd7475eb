// C++ drop-in test: robotoc_b200::DirectMultipleShooting + DeviceRiccatiRecursion run one full hot-path iteration
// (evalKKT tail -> backward -> forward -> computeStepSizes -> integrateSolution) on the GPU and compare the step sizes and
// the updated solution with the CPU oracle (orc_*_batch).  Inputs: seeded synthetic linearisations obeying the structure
// the reference's linearize* halves produce (M SPD, diagonal Qaa, SE(3) blocks [[A,B],[0,D]], slack / dual > 0).
#include <cmath>
#include <cstdio>
#include <vector>

#include "robotoc_b200/riccati_recursion.hpp"

extern "C" {
int orc_condense_batch(const rbt_stage_dims*, const rbt_constraint_table*, const rbt_stage_ctrl*, int, int, const double*, double*,
                       double*, double*, int);
int orc_riccati_batch(const rbt_dims*, const rbt_stage_ctrl*, int, double, int, double*, double*, const double*, double*, int);
void orc_expand_batch(const rbt_stage_dims*, const rbt_constraint_table*, const rbt_stage_ctrl*, int, int, const double*,
                      const double*, const double*, double*, double*, double*, int);
void orc_update_batch(const rbt_stage_dims*, const rbt_constraint_table*, const rbt_stage_ctrl*, int, int, double*, double*, double*,
                      double*, double*, const double*, int);
void orc_perf_index_batch(const rbt_stage_dims*, const rbt_constraint_table*, const rbt_stage_ctrl*, int, int, const double*,
                          const double*, double*);
}
using namespace robotoc_b200;

static unsigned long long g_state = 77ULL;
static double urand() {
  g_state = g_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return double((g_state >> 11) & ((1ULL << 53) - 1)) / double(1ULL << 52) - 1.0;
}
static double pos() { return 0.01 + 0.495 * (urand() + 1.0); }

int main() {
  const int nv = 18, nu = 12, nx = 36, nf = 12, nvf = 30, batch = 2, n_grid = 6;
  rbt_dims dims = {nv, nu, 12, 6};
  rbt_constraint_table tab = {};
  tab.n_contacts = 4; tab.barrier = 1e-3; tab.fraction_to_boundary = 0.995;
  int r = 0;
  const int vars[3] = {RBT_VAR_Q, RBT_VAR_V, RBT_VAR_U}, offs[3] = {6, 6, 0};
  for (int v = 0; v < 3; ++v)
    for (int sgn = -1; sgn <= 1; sgn += 2)
      for (int j = 0; j < 12; ++j) { tab.box[r].var = vars[v]; tab.box[r].idx = offs[v] + j; tab.box[r].sign = sgn; ++r; }
  tab.n_box = r;
  rbt_stage_dims sd = {nv, nu, 6, 12, 12, 4, tab.n_box};
  rbt_stage_layout S; rbt_make_stage_layout(&sd, &S);
  rbt_layout L; rbt_make_layout(&dims, &L);
  std::vector<rbt_stage_ctrl> ctrl(n_grid);
  for (int i = 0; i < n_grid; ++i) {
    ctrl[i] = rbt_stage_ctrl();
    ctrl[i].type = (i == n_grid - 1) ? RBT_TERMINAL : RBT_INTERMEDIATE;
    ctrl[i].nf = nf; ctrl[i].ngrids_in_phase = 1; ctrl[i].contact_mask = 0xF; ctrl[i].dt = (i == n_grid - 1) ? 0.0 : 0.05;
  }
  const size_t per = size_t(batch) * n_grid;
  std::vector<double> lin(per * S.l_stride, 0.0), con(per * S.c_stride, 0.0), sol(per * S.s_stride, 0.0), dx0(size_t(batch) * nx);
  auto putspd = [&](double* dst, int n, int ld, double diag, double sc) {
    std::vector<double> t(size_t(n) * n);
    for (auto& x : t) x = urand();
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double acc = (i == j) ? diag : 0.0;
        for (int k = 0; k < n; ++k) acc += sc * t[i + size_t(k) * n] * t[j + size_t(k) * n];
        dst[i + size_t(j) * ld] = acc;
      }
    for (int j = 0; j < n; ++j)  // exactly symmetric, as the reference's containers are (the wire format sends the upper triangle)
      for (int i = j + 1; i < n; ++i) dst[i + size_t(j) * ld] = dst[j + size_t(i) * ld];
  };
  for (size_t o = 0; o < per; ++o) {
    double* rec = lin.data() + o * S.l_stride;
    const bool terminal = (o % n_grid) == size_t(n_grid - 1);
    putspd(rec + S.l_Qxx, nx, nx, 1.0, 1.0 / nx);
    for (int i = 0; i < nx; ++i) rec[S.l_lx + i] = urand();
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < 6; ++j)
        for (int i = 0; i < 6; ++i) {
          double v = 0.0;
          if ((i < 3) == (j < 3)) v = ((i == j) ? 1.0 : 0.0) + 0.1 * urand();
          else if (i < 3) v = 0.1 * urand();
          rec[S.l_se3 + 36 * k + i + 6 * j] = (k == 0) ? -v : v;
        }
    double* s = sol.data() + o * S.s_stride;
    for (int i = 0; i < S.nq; ++i) s[S.s_q + i] = urand();
    double nrm = 0; for (int i = 3; i < 7; ++i) nrm += s[S.s_q + i] * s[S.s_q + i];
    for (int i = 3; i < 7; ++i) s[S.s_q + i] /= std::sqrt(nrm);
    for (int i = 0; i < nv; ++i) { s[S.s_v + i] = urand(); s[S.s_a + i] = urand(); s[S.s_lmd + i] = urand(); s[S.s_gmm + i] = urand(); s[S.s_beta + i] = urand(); }
    for (int i = 0; i < nu; ++i) s[S.s_u + i] = urand();
    for (int i = 0; i < nf; ++i) { s[S.s_f + i] = urand(); s[S.s_mu + i] = urand(); }
    for (int i = 0; i < 6; ++i) s[S.s_nup + i] = urand();
    if (terminal) continue;
    putspd(rec + S.l_M, nv, nv, 1.0, 0.1);
    for (int j = 0; j < nv; ++j) for (int i = 0; i < nf; ++i) rec[S.l_J + i + 12 * j] = urand();
    for (int j = 0; j < nx; ++j) for (int i = 0; i < nvf; ++i) rec[S.l_D + i + 30 * j] = 0.5 * urand();
    for (int i = 0; i < nvf; ++i) rec[S.l_IDC + i] = 0.5 * urand();
    for (int i = 0; i < nv; ++i) { rec[S.l_Qaa + i] = pos(); rec[S.l_la + i] = urand(); }
    for (int i = 0; i < nf; ++i) { rec[S.l_Qff + i + 12 * i] = 1e-3; rec[S.l_lf + i] = urand(); }
    for (int i = 0; i < nx; ++i) rec[S.l_Fx + i] = 0.1 * urand();
    putspd(rec + S.l_Quu, nu, nu, 0.1, 1.0 / nu);
    for (int i = 0; i < nu; ++i) rec[S.l_lu + i] = urand();
    for (int i = 0; i < 6; ++i) rec[S.l_lup + i] = urand();
    rec[S.l_sc + 1] = 1.0;
    for (int c = 0; c < 4; ++c) {
      for (int e = 0; e < 5 * nv; ++e) rec[S.l_dgdq + c * 5 * nv + e] = 0.3 * urand();
      for (int e = 0; e < 15; ++e) rec[S.l_dgdf + c * 15 + e] = urand();
    }
    double* c_ = con.data() + o * S.c_stride;
    for (int i = 0; i < S.nc; ++i) { c_[S.c_slack + i] = pos(); c_[S.c_dual + i] = pos(); c_[S.c_res + i] = 0.1 * urand(); }
  }
  for (auto& x : dx0) x = 0.1 * urand();

  // ---- oracle
  std::vector<double> ocon = con, osol = sol, okkt(per * L.k_stride, 0.0), oex(per * S.e_stride, 0.0), oric(per * L.r_stride, 0.0),
                      odir(per * L.d_stride, 0.0), oxd(per * S.x_stride, 0.0), osteps(2 * batch, 0.0);
  if (orc_condense_batch(&sd, &tab, ctrl.data(), n_grid, batch, lin.data(), ocon.data(), okkt.data(), oex.data(), 0) != 0) return 2;
  if (orc_riccati_batch(&dims, ctrl.data(), n_grid, 0.1, batch, okkt.data(), oric.data(), dx0.data(), odir.data(), 0) != 0) return 2;
  orc_expand_batch(&sd, &tab, ctrl.data(), n_grid, batch, lin.data(), oex.data(), odir.data(), ocon.data(), oxd.data(), osteps.data(), 0);
  orc_update_batch(&sd, &tab, ctrl.data(), n_grid, batch, oex.data(), odir.data(), oxd.data(), ocon.data(), osol.data(), osteps.data(), 0);

  // ---- the adaptor (GPU)
  DeviceRiccatiRecursion riccati_recursion(dims, ctrl, batch, 0.1);
  DirectMultipleShooting dms(riccati_recursion, sd, tab);
  const std::vector<double> sol_in = sol, con_in = con;
  dms.evalKKT(lin, con);
  // OCPSolver::KKTError() (PerformanceIndex of evalKKT) against the oracle
  dms.evalPerformanceIndex();
  std::vector<double> operf(8 * size_t(batch), 0.0);
  orc_perf_index_batch(&sd, &tab, ctrl.data(), n_grid, batch, lin.data(), con.data(), operf.data());
  double kworst = 0.0;
  for (int b = 0; b < batch; ++b)
    for (int q = 1; q <= 5; ++q)
      kworst = std::fmax(kworst, std::fabs(dms.performanceIndex(b)[q] - operf[8 * b + q]) / std::fabs(operf[8 * b + q]));
  if (!(dms.KKTError(0) > 0.0)) return 4;
  riccati_recursion.backwardRiccatiRecursion();
  riccati_recursion.forwardRiccatiRecursion(dx0);
  dms.computeStepSizes();
  dms.integrateSolution(sol);
  double worst = 0.0, den = 1e-300;
  for (size_t i = 0; i < sol.size(); ++i) { worst = std::fmax(worst, std::fabs(sol[i] - osol[i])); den = std::fmax(den, std::fabs(osol[i])); }
  worst /= den;
  double sworst = 0.0;
  for (int b = 0; b < batch; ++b) {
    sworst = std::fmax(sworst, std::fabs(dms.maxPrimalStepSize(b) - osteps[2 * b]) / osteps[2 * b]);
    sworst = std::fmax(sworst, std::fabs(dms.maxDualStepSize(b) - osteps[2 * b + 1]) / osteps[2 * b + 1]);
  }
  std::printf("robotoc_b200::DirectMultipleShooting vs oracle: solution rel err %.3e, step sizes rel err %.3e (primal %.4f dual %.4f)\n",
              worst, sworst, dms.maxPrimalStepSize(0), dms.maxDualStepSize(0));
  // the one-call host path with resident solver state and wire records: same bits as the step-by-step path
  std::vector<double> sol_r, sd_r, res(per * S.ncp, 0.0);
  for (size_t o = 0; o < per; ++o)
    for (int i = 0; i < S.ncp; ++i) res[o * S.ncp + i] = con_in[o * S.c_stride + S.c_res + i];
  dms.setWireCostStructure(false);
  const std::vector<double> wire = dms.packWire(lin, ctrl);
  dms.setState(sol_in, con_in);
  dms.iterationHostResident(wire, std::vector<double>(), res, dx0, sol_r, sd_r);
  const int used = S.s_xi + S.nsm;
  bool same = wire.size() < lin.size();
  for (size_t o = 0; o < per && same; ++o)
    for (int i = 0; i < used; ++i) same = same && (sol_r[o * S.s_stride + i] == sol[o * S.s_stride + i]);
  for (int b = 0; b < batch; ++b) same = same && dms.maxPrimalStepSize(b) > 0.0 && dms.maxPrimalStepSize(b) <= 1.0;
  std::printf("KKT error rel err %.3e; resident wire path %s the step-by-step path (%zu vs %zu doubles up)\n", kworst,
              same ? "reproduces" : "DIFFERS FROM", wire.size() + res.size(), lin.size() + con.size() + sol.size());
  if (!same || !(kworst < 1e-10)) return 5;
  bool threw = false;
  try { std::vector<double> bad(3); dms.evalKKT(bad, con); } catch (const std::invalid_argument&) { threw = true; }
  if (!threw) return 3;
  return (worst < 1e-8 && sworst < 1e-9) ? 0 : 1;
}
