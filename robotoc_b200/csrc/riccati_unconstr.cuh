// riccati_unconstr.cuh -- batched Riccati recursion for fixed-base robots without contacts (iiwa14: nv=7, nx=14).
//
// Behaviour of  UnconstrRiccatiRecursion::{backward,forward}RiccatiRecursion   src/riccati/unconstr_riccati_recursion.cpp:26-48
//               UnconstrRiccatiFactorizer                                      src/riccati/unconstr_riccati_factorizer.cpp:26-58
//               UnconstrBackwardRiccatiRecursionFactorizer                     src/riccati/unconstr_backward_riccati_recursion_factorizer.cpp:27-70
// A = [[I, dt I],[0, I]], B = [0; dt I] are implicit, so there is no GEMM: the stage is element-wise block adds, a
// 7x7 Cholesky, 15 triangular solves and a rank-7 update.  The matrices are far below one DMMA tile band, so the
// kernel is one warp per OCP with the whole KKT record (3 KB) arriving by a single cp.async.bulk into a 2-deep ring.
#pragma once
#include "rbt_device.cuh"
#include "riccati_backward.cuh"  // warp_cholesky
#include "../../include/rbt_layout.h"

namespace rbt {

struct UParams {
  rbt_ulayout L;
  int N;
  int batch;
  double dt;
  const double* kkt;
  double* ric;
  double* fact;
  const double* dx0;
  double* dir;
  int* info;
};

template <int NV>
struct UCfg {
  static constexpr int NX = 2 * NV;
  static constexpr int NVP = (NV <= 8) ? 8 : ((NV <= 16) ? 16 : 32);
  static constexpr int KREC = ((NX * NX + 1) & ~1) + ((NX * NV + 1) & ~1) + ((NV * NV + 1) & ~1) + 2 * ((NX + 1) & ~1) +
                              ((NV + 1) & ~1);
  static constexpr int KSTRIDE = (KREC + 15) & ~15;
  static constexpr int o_in = 0;                 // 2 slots of KSTRIDE
  static constexpr int o_P = 2 * KSTRIDE;        // P+ (nx x nx)
  static constexpr int o_F = o_P + NX * NX;      // F
  static constexpr int o_H = o_F + NX * NX;      // H (nx x nv)
  static constexpr int o_G = o_H + NX * NV;      // G / L
  static constexpr int o_Y = o_G + NV * NV;      // Y (nv x nx)
  static constexpr int o_vec = o_Y + NV * NX;    // s+ (nx), la' (nv), k (nv), dinv (nv), PFx (nx)
  static constexpr int o_bar = (o_vec + 2 * NX + 3 * NV + 1) & ~1;
  static constexpr int SMEM_DOUBLES = o_bar + 2;
};

template <int NV>
__global__ void __launch_bounds__(32) unconstr_backward_kernel(const UParams p) {
  using C = UCfg<NV>;
  constexpr int NX = C::NX;
  static_assert(NX < 32, "one warp per OCP: nx must be below 32");
  __shared__ __align__(16) double smem[C::SMEM_DOUBLES];
  const rbt_ulayout& L = p.L;
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  if (b >= p.batch) return;
  const int N = p.N;
  const double dt = p.dt, dt2 = dt * dt;
  const double* kkt_b = p.kkt + size_t(b) * (N + 1) * L.k_stride;
  double* ric_b = p.ric + size_t(b) * (N + 1) * L.r_stride;
  double* fact_b = p.fact ? p.fact + size_t(b) * (N + 1) * L.f_stride : nullptr;
  double* sP = smem + C::o_P;
  double* sF = smem + C::o_F;
  double* sH = smem + C::o_H;
  double* sG = smem + C::o_G;
  double* sY = smem + C::o_Y;
  double* s_n = smem + C::o_vec;
  double* la2 = s_n + NX;
  double* kv = la2 + NV;
  double* dinv = kv + NV;
  double* PFx = dinv + NV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::o_bar);
  int bad = 0;

  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int st) {
    const int slot = st & 1;
    fence_proxy_async();
    mbar_expect_tx(&bars[slot], uint32_t(L.k_stride) * 8u);
    tma_load_1d(smem + C::o_in + slot * C::KSTRIDE, kkt_b + size_t(st) * L.k_stride, uint32_t(L.k_stride) * 8u,
                &bars[slot]);
  };
  if (lane == 0) {
    if (N >= 1) issue(N - 1);
    if (N >= 2) issue(N - 2);
  }
  {  // terminal: P_N = Qxx_N, s_N = -lx_N              unconstr_riccati_recursion.cpp:29-30
    const double* recN = kkt_b + size_t(N) * L.k_stride;
    double* ricN = ric_b + size_t(N) * L.r_stride;
    for (int e = lane; e < NX * NX; e += 32) {
      const double v = recN[L.k_Qxx + e];
      sP[e] = v;
      ricN[L.r_P + e] = v;
    }
    for (int e = lane; e < NX; e += 32) {
      const double v = -recN[L.k_lx + e];
      s_n[e] = v;
      ricN[L.r_s + e] = v;
    }
  }
  __syncwarp();
  int cnt = 0;
  for (int i = N - 1; i >= 0; --i, ++cnt) {
    const int slot = i & 1;
    const double* in = smem + C::o_in + slot * C::KSTRIDE;
    const double* Qxx = in + L.k_Qxx;
    const double* Qxu = in + L.k_Qxu;
    const double* Qaa = in + L.k_Qaa;
    const double* Fx = in + L.k_Fx;
    const double* lx = in + L.k_lx;
    const double* la = in + L.k_la;
    double* ric = ric_b + size_t(i) * L.r_stride;
    double* fct = fact_b ? fact_b + size_t(i) * L.f_stride : nullptr;
    mbar_wait(&bars[slot], uint32_t(cnt >> 1) & 1u);

    // factorizeKKTMatrix: unconstr_backward_riccati_recursion_factorizer.cpp:31-50
    for (int e = lane; e < NX * NX; e += 32) {
      const int r = e % NX, c = e / NX;
      double v = Qxx[e] + sP[e];
      if (r >= NV) v += dt * sP[(r - NV) + c * NX];
      if (c >= NV) v += dt * sP[r + (c - NV) * NX];
      if (r >= NV && c >= NV) v += dt2 * sP[(r - NV) + (c - NV) * NX];
      sF[e] = v;
    }
    for (int e = lane; e < NX * NV; e += 32) {
      const int r = e % NX, c = e / NX;
      double v = Qxu[e] + dt * sP[r + (NV + c) * NX];
      if (r >= NV) v += dt2 * sP[(r - NV) + (NV + c) * NX];
      sH[e] = v;
      if (fct) fct[L.f_H + e] = v;
    }
    for (int e = lane; e < NV * NV; e += 32) {
      const int r = e % NV, c = e / NV;
      const double v = Qaa[e] + dt2 * sP[(NV + r) + (NV + c) * NX];
      sG[e] = v;
      if (fct) fct[L.f_G + e] = v;
    }
    for (int r = lane; r < NX; r += 32) {  // PFx = P+ Fx
      double a = 0.0;
      for (int k = 0; k < NX; ++k) a = fma(sP[r + k * NX], Fx[k], a);
      PFx[r] = a;
    }
    __syncwarp();
    if (lane < NV) {  // la' = la + dt (P+ Fx)_v - dt s+_v        :45-47
      const double v = la[lane] + dt * PFx[NV + lane] - dt * s_n[NV + lane];
      la2[lane] = v;
      if (fct) fct[L.f_la + lane] = v;
    }
    __syncwarp();
    if (!warp_cholesky<C::NVP>(sG, NV, dinv)) bad |= 1;
    __syncwarp();
    // Y = L^-1 H^T ; K = -L^-T Y ; k = -G^-1 la'                  unconstr_riccati_factorizer.cpp:35-36
    if (lane <= NX) {
      const int c = lane;
      double y[NV];
#pragma unroll
      for (int u = 0; u < NV; ++u) y[u] = (c < NX) ? sH[c + u * NX] : la2[u];
#pragma unroll
      for (int a = 0; a < NV; ++a) {
        double v = y[a];
#pragma unroll
        for (int k = 0; k < a; ++k) v = fma(-sG[a + k * NV], y[k], v);
        y[a] = v * dinv[a];
      }
      if (c < NX) {
#pragma unroll
        for (int u = 0; u < NV; ++u) sY[u + c * NV] = y[u];
      }
#pragma unroll
      for (int a = NV - 1; a >= 0; --a) {
        double v = y[a];
#pragma unroll
        for (int k = a + 1; k < NV; ++k) v = fma(-sG[k + a * NV], y[k], v);
        y[a] = v * dinv[a];
      }
      if (c < NX) {
#pragma unroll
        for (int u = 0; u < NV; ++u) ric[L.r_K + c + u * NX] = -y[u];
      } else {
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          kv[u] = -y[u];
          ric[L.r_k + u] = -y[u];
        }
      }
    }
    __syncwarp();
    // F -= Y^T Y  (== K^T G K), fact F, then P = sym(F)             unconstr_backward_..factorizer.cpp:59-61
    for (int e = lane; e < NX * NX; e += 32) {
      const int r = e % NX, c = e / NX;
      double a = sF[e];
#pragma unroll
      for (int u = 0; u < NV; ++u) a = fma(-sY[u + r * NV], sY[u + c * NV], a);
      sF[e] = a;
      if (fct) fct[L.f_F + e] = a;
    }
    // s = s+ ; s_v += dt s+_q ; s -= P+ Fx ; s_v -= dt (P+ Fx)_q ; s -= lx ; s -= H k     :63-69
    double snew0 = 0.0, snew1 = 0.0;
    for (int r = lane, q = 0; r < NX; r += 32, ++q) {
      double v = s_n[r] - PFx[r] - lx[r];
      if (r >= NV) v += dt * s_n[r - NV] - dt * PFx[r - NV];
#pragma unroll
      for (int u = 0; u < NV; ++u) v = fma(-sH[r + u * NX], kv[u], v);
      if (q == 0) snew0 = v; else snew1 = v;
    }
    __syncwarp();
    for (int r = lane, q = 0; r < NX; r += 32, ++q) {
      const double v = (q == 0) ? snew0 : snew1;
      s_n[r] = v;
      ric[L.r_s + r] = v;
    }
    for (int e = lane; e < NX * NX; e += 32) {
      const int r = e % NX, c = e / NX;
      const double v = 0.5 * (sF[e] + sF[c + r * NX]);
      sP[e] = v;
      ric[L.r_P + e] = v;
    }
    __syncwarp();
    if (lane == 0 && i - 2 >= 0) issue(i - 2);
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (lane == 0 && bad) atomicOr(&p.info[b], bad);
}

// forward: da = K dx + k ; dx+ = Fx + dx ; dq+ += dt dv ; dv+ += dt da ; dlmdgmm = P dx - s
//   unconstr_riccati_factorizer.cpp:43-58.  Streams [P|s|K|k] per stage; Fx is read directly.
template <int NV>
__global__ void __launch_bounds__(32) unconstr_forward_kernel(const UParams p) {
  constexpr int NX = 2 * NV;
  constexpr int RPART = ((NX * NX + 1) & ~1) + ((NX + 1) & ~1) + ((NV * NX + 1) & ~1) + ((NV + 1) & ~1);
  constexpr int RREC = RPART + ((NX + 1) & ~1);  // [P|s|K|k] from the Riccati record + Fx from the KKT record
  static_assert(NX < 32, "one warp per OCP: nx must be below 32");
  __shared__ __align__(16) double smem[2 * RREC + 2 * NX + NV + 1 + 4];
  const rbt_ulayout& L = p.L;
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  if (b >= p.batch) return;
  const int N = p.N;
  const double dt = p.dt;
  const double* kkt_b = p.kkt + size_t(b) * (N + 1) * L.k_stride;
  const double* ric_b = p.ric + size_t(b) * (N + 1) * L.r_stride;
  double* dir_b = p.dir + size_t(b) * (N + 1) * L.d_stride;
  double* sdx = smem + 2 * RREC;
  double* sda = sdx + 2 * NX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((2 * RREC + 2 * NX + NV + 1) & ~1));
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  for (int r = lane; r < NX; r += 32) sdx[r] = p.dx0[size_t(b) * NX + r];
  __syncwarp();
  auto issue = [&](int st) {
    const int slot = st & 1;
    fence_proxy_async();
    mbar_expect_tx(&bars[slot], uint32_t(RREC) * 8u);
    tma_load_1d(smem + slot * RREC, ric_b + size_t(st) * L.r_stride + L.r_P, uint32_t(RPART) * 8u, &bars[slot]);
    tma_load_1d(smem + slot * RREC + RPART, kkt_b + size_t(st) * L.k_stride + L.k_Fx, uint32_t(RREC - RPART) * 8u,
                &bars[slot]);
  };
  if (lane == 0) {
    issue(0);
    if (N >= 1) issue(1);
  }
  for (int i = 0; i <= N; ++i) {
    const int slot = i & 1;
    const double* sPm = smem + slot * RREC;
    const double* ss = sPm + ((NX * NX + 1) & ~1);
    const double* sKt = ss + ((NX + 1) & ~1);
    const double* sk = sKt + ((NV * NX + 1) & ~1);
    double* dx = sdx + (i & 1) * NX;
    double* dxn = sdx + ((i + 1) & 1) * NX;
    double* drec = dir_b + size_t(i) * L.d_stride;
    const double* Fx = sPm + RPART;
    mbar_wait(&bars[slot], uint32_t(i >> 1) & 1u);
    if (i < N) {
      if (lane < NV) {
        double a = sk[lane];
        for (int k = 0; k < NX; ++k) a = fma(sKt[k + lane * NX], dx[k], a);
        sda[lane] = a;
        drec[L.d_da + lane] = a;
      }
      __syncwarp();
      for (int r = lane; r < NX; r += 32) {
        double v = Fx[r] + dx[r];
        v += (r < NV) ? dt * dx[NV + r] : dt * sda[r - NV];
        dxn[r] = v;
      }
    }
    for (int r = lane; r < NX; r += 32) {
      double a = -ss[r];
      for (int k = 0; k < NX; ++k) a = fma(sPm[r + k * NX], dx[k], a);
      drec[L.d_dlmdgmm + r] = a;
      drec[L.d_dx + r] = dx[r];
    }
    __syncwarp();
    if (lane == 0 && i + 2 <= N) issue(i + 2);
    __syncwarp();
  }
}

}  // namespace rbt
