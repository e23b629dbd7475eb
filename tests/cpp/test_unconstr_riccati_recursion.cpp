// C++ drop-in test for the unconstrained path: robotoc_b200::UnconstrRiccatiRecursion (reference class / method names over the
// C ABI) on an iiwa14-sized horizon (nv = 7, N = 20), compared with the CPU oracle (orc_unconstr_batch): P, s, K, k, dx, da,
// dlmdgmm and the mutated KKT blocks.  Recipe of the inputs: test/riccati/unconstr_riccati_recursion_test.cpp:33-46.
#include <cmath>
#include <cstdio>
#include <vector>

#include "robotoc_b200/unconstr_riccati_recursion.hpp"

extern "C" int orc_unconstr_batch(int, int, double, int, double*, double*, const double*, double*, int);
using namespace robotoc_b200;

static unsigned long long g_state = 99ULL;
static double urand() {
  g_state = g_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return double((g_state >> 11) & ((1ULL << 53) - 1)) / double(1ULL << 52) - 1.0;
}
static double worst = 0.0;
static bool close_to(const double* a, const double* b, int n, const char* what, int i) {
  double scale = 0.0, err = 0.0;
  for (int k = 0; k < n; ++k) { scale = std::fmax(scale, std::fabs(b[k])); err = std::fmax(err, std::fabs(a[k] - b[k])); }
  const double rel = err / (scale > 0 ? scale : 1.0);
  worst = std::fmax(worst, rel);
  if (rel > 1e-8) { std::printf("FAILED stage %d %s: rel err %.3e\n", i, what, rel); return false; }
  return true;
}

int main() {
  const int nv = 7, nx = 14, N = 20;
  const double T = 1.0;
  rbt_ulayout L; rbt_make_ulayout(nv, &L);
  std::vector<UnconstrSplitKKTMatrix> km(N + 1, UnconstrSplitKKTMatrix(nv));
  std::vector<UnconstrSplitKKTResidual> kr(N + 1, UnconstrSplitKKTResidual(nv));
  std::vector<double> kkt(size_t(N + 1) * L.k_stride, 0.0), dx0(nx);
  for (int i = 0; i <= N; ++i) {
    const int n3 = 3 * nv;
    std::vector<double> G(size_t(n3) * n3);
    for (auto& x : G) x = urand();
    auto H = [&](int a, int b) { double acc = 0; for (int k = 0; k < n3; ++k) acc += G[a + size_t(k) * n3] * G[b + size_t(k) * n3]; return acc; };
    for (int b = 0; b < nx; ++b) for (int a = 0; a < nx; ++a) km[i].Qxx(a, b) = H(a, b);
    for (int a = 0; a < nx; ++a) kr[i].lx[a] = urand();
    if (i < N) {
      for (int b = 0; b < nv; ++b) for (int a = 0; a < nx; ++a) km[i].Qxu(a, b) = H(a, nx + b);
      for (int b = 0; b < nv; ++b) for (int a = 0; a < nv; ++a) km[i].Qaa(a, b) = H(nx + a, nx + b);
      for (int a = 0; a < nx; ++a) kr[i].Fx[a] = urand();
      for (int a = 0; a < nv; ++a) kr[i].la[a] = urand();
    }
    double* rec = kkt.data() + size_t(i) * L.k_stride;
    for (int e = 0; e < nx * nx; ++e) rec[L.k_Qxx + e] = km[i].Qxx.data()[e];
    for (int e = 0; e < nx; ++e) rec[L.k_lx + e] = kr[i].lx[e];
    if (i < N) {
      for (int e = 0; e < nx * nv; ++e) rec[L.k_Qxu + e] = km[i].Qxu.data()[e];
      for (int e = 0; e < nv * nv; ++e) rec[L.k_Qaa + e] = km[i].Qaa.data()[e];
      for (int e = 0; e < nx; ++e) rec[L.k_Fx + e] = kr[i].Fx[e];
      for (int e = 0; e < nv; ++e) rec[L.k_la + e] = kr[i].la[e];
    }
  }
  for (auto& x : dx0) x = urand();
  // oracle
  std::vector<double> ric_o(size_t(N + 1) * L.r_stride, 0.0), dir_o(size_t(N + 1) * L.d_stride, 0.0);
  if (orc_unconstr_batch(nv, N, T / N, 1, kkt.data(), ric_o.data(), dx0.data(), dir_o.data(), 1) != 0) { std::printf("FAILED oracle\n"); return 1; }
  // adaptor
  try {
    UnconstrRiccatiRecursion rr(nv, N, T);
    std::vector<UnconstrSplitRiccatiFactorization> fact;
    std::vector<UnconstrSplitDirection> d(N + 1, UnconstrSplitDirection(nv));
    rr.backwardRiccatiRecursion(km, kr, fact);
    d[0].dx = dx0;
    rr.forwardRiccatiRecursion(kr, fact, d);
    bool ok = true;
    for (int i = 0; i <= N; ++i) {
      const double* r = ric_o.data() + size_t(i) * L.r_stride;
      const double* dd = dir_o.data() + size_t(i) * L.d_stride;
      const double* kk = kkt.data() + size_t(i) * L.k_stride;  // mutated by the oracle like the reference mutates it
      ok &= close_to(fact[i].P.data(), r + L.r_P, nx * nx, "P", i);
      ok &= close_to(fact[i].s.data(), r + L.r_s, nx, "s", i);
      ok &= close_to(d[i].dx.data(), dd + L.d_dx, nx, "dx", i);
      ok &= close_to(d[i].dlmdgmm.data(), dd + L.d_dlmdgmm, nx, "dlmdgmm", i);
      if (i == N) continue;
      ok &= close_to(rr.getLQRPolicy()[i].Kt.data(), r + L.r_K, nx * nv, "K", i);
      ok &= close_to(rr.getLQRPolicy()[i].k.data(), r + L.r_k, nv, "k", i);
      ok &= close_to(d[i].da.data(), dd + L.d_da, nv, "da", i);
      ok &= close_to(km[i].Qxx.data(), kk + L.k_Qxx, nx * nx, "mutated Qxx", i);
      ok &= close_to(km[i].Qaa.data(), kk + L.k_Qaa, nv * nv, "mutated Qaa", i);
      ok &= close_to(kr[i].la.data(), kk + L.k_la, nv, "mutated la", i);
    }
    // argument errors follow the reference's exception types
    bool threw = false;
    try { UnconstrRiccatiRecursion bad(nv, 0, T); } catch (const std::out_of_range&) { threw = true; }
    ok &= threw;
    if (!ok) return 1;
  } catch (const std::exception& e) {
    std::printf("FAILED exception: %s\n", e.what());
    return 1;
  }
  std::printf("ok: UnconstrRiccatiRecursion through the C ABI, worst rel err %.2e\n", worst);
  return 0;
}
