// Two ranks through the C ABI (one process, two GPUs, ncclCommInitAll): each rank solves its own batch of OCPs
// (backward + forward Riccati on seeded KKT records), then rbt_allgather_step gathers the packed Newton step of both ranks on
// both devices.  Checked: every rank's slice of the gathered buffer equals that rank's own direction records, bit for bit.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

extern "C" {
#include "robotoc_b200.h"
}

static unsigned long long g_state = 4242ULL;
static double urand() {
  g_state = g_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return double((g_state >> 11) & ((1ULL << 53) - 1)) / double(1ULL << 52) - 1.0;
}

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { std::printf("FAILED %s -> %d (line %d)\n", #x, rc_, __LINE__); return 1; } } while (0)

int main() {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < 2) { std::printf("SKIP: needs 2 GPUs, found %d\n", ndev); return 77; }
  const int nv = 18, nu = 12, nx = 36, batch = 3, n_grid = 5, R = 2;
  rbt_dims dims = {nv, nu, 12, 6};
  rbt_layout L; rbt_make_layout(&dims, &L);
  std::vector<rbt_stage_ctrl> ctrl(n_grid);
  for (int i = 0; i < n_grid; ++i) {
    ctrl[i] = rbt_stage_ctrl();
    ctrl[i].type = (i == n_grid - 1) ? RBT_TERMINAL : RBT_INTERMEDIATE;
    ctrl[i].nf = 12; ctrl[i].contact_mask = 0xF; ctrl[i].ngrids_in_phase = n_grid - 1; ctrl[i].dt = (i == n_grid - 1) ? 0.0 : 0.05;
  }
  int devs[R] = {0, 1};
  ncclComm_t comms[R];
  CK(ncclCommInitAll(comms, R, devs));
  rbt_handle* h[R];
  cudaStream_t st[R];
  double* all_dev[R];
  std::vector<std::vector<double>> dir_host(R);
  const size_t per = size_t(batch) * n_grid;
  for (int r = 0; r < R; ++r) {
    CK(rbt_create(&dims, n_grid, batch, devs[r], &h[r]));
    CK(rbt_set_schedule(h[r], ctrl.data(), n_grid, 0.1));
    cudaSetDevice(devs[r]);
    cudaStreamCreate(&st[r]);
    // seeded KKT records: Q = T T^T + I blocks, structured Fxx = [[I, dt I], [small, I + small]]
    std::vector<double> kkt(per * L.k_stride, 0.0), dx0(size_t(batch) * nx);
    for (size_t o = 0; o < per; ++o) {
      double* rec = kkt.data() + o * L.k_stride;
      const int i = int(o % n_grid);
      const int nq = nx + nu;
      std::vector<double> T(size_t(nq) * nq);
      for (auto& x : T) x = urand();
      auto Hq = [&](int a, int b) { double acc = (a == b) ? 1.0 : 0.0; for (int k = 0; k < nq; ++k) acc += 0.05 * T[a + size_t(k) * nq] * T[b + size_t(k) * nq]; return acc; };
      for (int b = 0; b < nx; ++b) for (int a = 0; a < nx; ++a) rec[L.k_Qxx + a + b * nx] = Hq(a, b);
      for (int a = 0; a < nx; ++a) rec[L.k_lx + a] = urand();
      if (i == n_grid - 1) continue;
      for (int b = 0; b < nu; ++b) for (int a = 0; a < nx; ++a) rec[L.k_Qxu + a + b * nx] = Hq(a, nx + b);
      for (int b = 0; b < nu; ++b) for (int a = 0; a < nu; ++a) rec[L.k_Quu + a + b * nu] = Hq(nx + a, nx + b);
      for (int a = 0; a < nu; ++a) rec[L.k_lu + a] = urand();
      for (int a = 0; a < nx; ++a) rec[L.k_Fx + a] = 0.1 * urand();
      for (int a = 0; a < nv; ++a) { rec[L.k_Fxx + a + a * nx] = 1.0; rec[L.k_Fxx + a + (nv + a) * nx] = 0.05; rec[L.k_Fxx + (nv + a) + (nv + a) * nx] = 1.0; }
      for (int b = 0; b < nx; ++b) for (int a = 0; a < nv; ++a) rec[L.k_Fxx + (nv + a) + b * nx] += 0.05 * urand();
      for (int b = 0; b < nu; ++b) for (int a = 0; a < nv; ++a) rec[L.k_Fvu + a + b * nv] = 0.05 * urand();
    }
    for (auto& x : dx0) x = urand();
    CK(rbt_upload(h[r], RBT_BUF_KKT, kkt.data(), st[r]));
    CK(rbt_upload(h[r], RBT_BUF_DX0, dx0.data(), st[r]));
    CK(rbt_riccati_backward(h[r], 0, st[r]));
    CK(rbt_riccati_forward(h[r], st[r]));
    CK(rbt_check_info(h[r], nullptr, st[r]));
    dir_host[r].resize(per * L.d_stride);
    CK(rbt_download(h[r], RBT_BUF_DIR, dir_host[r].data(), st[r]));
    CK(rbt_sync(h[r], st[r]));
    cudaMalloc(&all_dev[r], size_t(R) * per * rbt_step_doubles(h[r]) * sizeof(double));
  }
  const int step = rbt_step_doubles(h[0]);
  if (step != L.d_dts + 2) { std::printf("FAILED step size %d\n", step); return 1; }
  CK(ncclGroupStart());
  for (int r = 0; r < R; ++r) CK(rbt_allgather_step(h[r], comms[r], all_dev[r], st[r]));
  CK(ncclGroupEnd());
  int bad = 0;
  for (int r = 0; r < R; ++r) {
    CK(rbt_sync(h[r], st[r]));
    std::vector<double> all(size_t(R) * per * step);
    cudaSetDevice(devs[r]);
    cudaMemcpy(all.data(), all_dev[r], all.size() * sizeof(double), cudaMemcpyDeviceToHost);
    for (int src = 0; src < R; ++src)
      for (size_t o = 0; o < per; ++o)
        for (int k = 0; k < step; ++k)
          if (all[(size_t(src) * per + o) * step + k] != dir_host[src][o * L.d_stride + k]) ++bad;
  }
  for (int r = 0; r < R; ++r) { rbt_destroy(h[r]); ncclCommDestroy(comms[r]); }
  if (bad) { std::printf("FAILED: %d gathered entries differ from the owning rank's direction records\n", bad); return 1; }
  std::printf("ok: %d ranks x %d OCPs x %d grid points x %d doubles gathered on every rank\n", R, batch, n_grid, step);
  return 0;
}
