// C++ drop-in test: robotoc_b200::RiccatiRecursion (reference signatures, C ABI underneath) vs the CPU oracle.
// Mirrors the style of the reference's test/riccati/unconstr_riccati_recursion_test.cpp:62-105 (recursion class vs
// explicit loop, checks the mutated KKT, P, s, K, d) with a fixed seed and a horizon that has a Lift, an Impact and a
// switching-constraint stage.  Build: g++ -std=c++14 -Iinclude tests/cpp/test_riccati_recursion.cpp -Lrobotoc_b200
// -lrobotoc_b200 -Loracle -loracle  (done by tests/test_gpu_cpp_adaptor.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "robotoc_b200/riccati_recursion.hpp"

extern "C" {
int orc_riccati_backward(const rbt_dims*, const rbt_stage_ctrl*, int, double, double*, double*);
void orc_riccati_forward(const rbt_dims*, const rbt_stage_ctrl*, int, const double*, const double*, double*);
}
using namespace robotoc_b200;

static unsigned long long g_state = 20260924ULL;
static double urand() {  // U(-1,1), like Eigen::Random
  g_state = g_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return double((g_state >> 11) & ((1ULL << 53) - 1)) / double(1ULL << 52) - 1.0;
}
static void fill(Matrix& m, double sc = 1.0) { for (auto& v : m.a) v = sc * urand(); }
static void fill(Vector& v, double sc = 1.0) { for (auto& x : v) x = sc * urand(); }
static Matrix spd(int n) {
  Matrix s(n, n), h(n, n);
  fill(s);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double acc = 0;
      for (int k = 0; k < n; ++k) acc += s(i, k) * s(j, k);
      h(i, j) = acc;
    }
  return h;
}
static double relerr(const double* a, const double* b, int n) {
  double num = 0, den = 1e-300;
  for (int i = 0; i < n; ++i) { num = std::fmax(num, std::fabs(a[i] - b[i])); den = std::fmax(den, std::fabs(b[i])); }
  return num / den;
}

int main() {
  const int nv = 18, nu = 12, nx = 36, ns = 6;
  rbt_dims dims = {nv, nu, 12, 6};
  // N = 8 stages + events: [I, I(lift next), L, I, I(sw), I, Impact, I, I, I, I, Terminal]
  TimeDiscretization td(12);
  const GridType types[12] = {GridType::Intermediate, GridType::Intermediate, GridType::Lift, GridType::Intermediate,
                              GridType::Intermediate, GridType::Intermediate, GridType::Impact, GridType::Intermediate,
                              GridType::Intermediate, GridType::Intermediate, GridType::Intermediate, GridType::Terminal};
  for (int i = 0; i < 12; ++i) { td[i].type = types[i]; td[i].dt = (types[i] == GridType::Impact || i == 11) ? 0.0 : 0.05; td[i].dimf = 12; td[i].contact_mask = 0b1111; }
  td[4].switching_constraint = true; td[4].dims = ns;  // grid i+2 is the Impact
  const int n_grid = 12;
  KKTMatrix kkt_matrix(n_grid, SplitKKTMatrix(nv, nu, 12));
  KKTResidual kkt_residual(n_grid, SplitKKTResidual(nv, nu));
  for (int i = 0; i < n_grid; ++i) {
    SplitKKTMatrix& km = kkt_matrix[i];
    const double dt = td[i].dt;
    Matrix H = spd(nx + nu);
    for (int j = 0; j < nx; ++j) for (int r = 0; r < nx; ++r) km.Qxx(r, j) = H(r, j);
    for (int j = 0; j < nu; ++j) for (int r = 0; r < nx; ++r) km.Qxu(r, j) = H(r, nx + j);
    for (int j = 0; j < nu; ++j) for (int r = 0; r < nu; ++r) km.Quu(r, j) = H(nx + r, nx + j);
    fill(km.Fxx, dt > 0 ? dt : 0.05);
    for (int r = 0; r < nx; ++r) km.Fxx(r, r) += 1.0;
    for (int r = 0; r < nv; ++r) km.Fxx(r, nv + r) += dt;
    fill(km.Fvu, dt);
    fill(kkt_residual[i].Fx); fill(kkt_residual[i].lx); fill(kkt_residual[i].lu);
    if (td[i].switching_constraint) {
      km.setSwitchingConstraintDimension(ns);
      fill(km.Phix()); fill(km.Phiu());
      kkt_residual[i].P().assign(ns, 0.0); fill(kkt_residual[i].P());
    }
  }
  // ---- oracle on packed copies
  rbt_layout L; rbt_make_layout(&dims, &L);
  std::vector<rbt_stage_ctrl> ctrl(n_grid);
  for (int i = 0; i < n_grid; ++i) {
    ctrl[i] = rbt_stage_ctrl();
    ctrl[i].type = int(td[i].type); ctrl[i].ns = td[i].switching_constraint ? td[i].dims : 0; ctrl[i].nf = 12;
    ctrl[i].ngrids_in_phase = 1; ctrl[i].dt = td[i].dt;
  }
  std::vector<double> kk(size_t(n_grid) * L.k_stride, 0.0), rr(size_t(n_grid) * L.r_stride, 0.0), dd(size_t(n_grid) * L.d_stride, 0.0);
  for (int i = 0; i < n_grid; ++i) {
    double* rec = kk.data() + size_t(i) * L.k_stride;
    std::memcpy(rec + L.k_Qxx, kkt_matrix[i].Qxx.data(), 8 * nx * nx);
    std::memcpy(rec + L.k_lx, kkt_residual[i].lx.data(), 8 * nx);
    if (td[i].type == GridType::Terminal) continue;
    std::memcpy(rec + L.k_Fxx, kkt_matrix[i].Fxx.data(), 8 * nx * nx);
    std::memcpy(rec + L.k_Fx, kkt_residual[i].Fx.data(), 8 * nx);
    if (td[i].type == GridType::Impact) continue;
    std::memcpy(rec + L.k_Fvu, kkt_matrix[i].Fvu.data(), 8 * nv * nu);
    std::memcpy(rec + L.k_Qxu, kkt_matrix[i].Qxu.data(), 8 * nx * nu);
    std::memcpy(rec + L.k_Quu, kkt_matrix[i].Quu.data(), 8 * nu * nu);
    std::memcpy(rec + L.k_lu, kkt_residual[i].lu.data(), 8 * nu);
    if (ctrl[i].ns > 0) {
      std::memcpy(rec + L.k_Phix, kkt_matrix[i].Phix().data(), 8 * ns * nx);
      std::memcpy(rec + L.k_Phiu, kkt_matrix[i].Phiu().data(), 8 * ns * nu);
      std::memcpy(rec + L.k_p, kkt_residual[i].P().data(), 8 * ns);
    }
  }
  Direction d(n_grid, SplitDirection(nv, nu));
  fill(d[0].dx);
  std::memcpy(dd.data() + L.d_dx, d[0].dx.data(), 8 * nx);
  if (orc_riccati_backward(&dims, ctrl.data(), n_grid, 0.1, kk.data(), rr.data()) != 0) { std::printf("oracle chol failed\n"); return 2; }
  orc_riccati_forward(&dims, ctrl.data(), n_grid, kk.data(), rr.data(), dd.data());

  // ---- the adaptor (GPU)
  RiccatiFactorization factorization;
  RiccatiRecursion riccati_recursion(dims, n_grid, 0.1);
  riccati_recursion.backwardRiccatiRecursion(td, kkt_matrix, kkt_residual, factorization);
  riccati_recursion.forwardRiccatiRecursion(td, kkt_matrix, kkt_residual, factorization, d);
  const auto& lqr = riccati_recursion.getLQRPolicy();

  double worst = 0;
  for (int i = 0; i < n_grid; ++i) {
    const double* r = rr.data() + size_t(i) * L.r_stride;
    const double* dr = dd.data() + size_t(i) * L.d_stride;
    const double* k = kk.data() + size_t(i) * L.k_stride;
    worst = std::fmax(worst, relerr(factorization[i].P.data(), r + L.r_P, nx * nx));
    worst = std::fmax(worst, relerr(factorization[i].s.data(), r + L.r_s, nx));
    worst = std::fmax(worst, relerr(d[i].dx.data(), dr + L.d_dx, nx));
    worst = std::fmax(worst, relerr(d[i].dlmdgmm.data(), dr + L.d_dlmdgmm, nx));
    if (td[i].type == GridType::Intermediate || td[i].type == GridType::Lift) {
      worst = std::fmax(worst, relerr(lqr[i].Kt.data(), r + L.r_K, nx * nu));
      worst = std::fmax(worst, relerr(lqr[i].k.data(), r + L.r_k, nu));
      worst = std::fmax(worst, relerr(d[i].du.data(), dr + L.d_du, nu));
      // in-place mutation semantics: Qxx,Qxu,Quu,lu now hold F,H,G,lu' (unconstr_riccati_recursion_test.cpp:82-86 style)
      worst = std::fmax(worst, relerr(kkt_matrix[i].Qxx.data(), k + L.k_Qxx, nx * nx));
      worst = std::fmax(worst, relerr(kkt_matrix[i].Qxu.data(), k + L.k_Qxu, nx * nu));
      worst = std::fmax(worst, relerr(kkt_matrix[i].Quu.data(), k + L.k_Quu, nu * nu));
      worst = std::fmax(worst, relerr(kkt_residual[i].lu.data(), k + L.k_lu, nu));
      if (ctrl[i].ns > 0) {
        worst = std::fmax(worst, relerr(factorization[i].M().data(), r + L.r_M, ns * nx));
        worst = std::fmax(worst, relerr(d[i].dxi().data(), dr + L.d_dxi, ns));
      }
    }
  }
  std::printf("robotoc_b200::RiccatiRecursion vs oracle: worst relative error %.3e over %d grid points\n", worst, n_grid);
  // argument errors follow the reference's conventions
  bool threw = false;
  try { riccati_recursion.setRegularization(-1.0); } catch (const std::out_of_range&) { threw = true; }
  if (!threw) { std::printf("setRegularization(-1) did not throw\n"); return 3; }
  return worst < 1e-8 ? 0 : 1;
}
