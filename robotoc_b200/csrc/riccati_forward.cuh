// riccati_forward.cuh -- batched forward Riccati recursion (state / control / costate / multiplier directions).
//
// Behaviour of   RiccatiRecursion::forwardRiccatiRecursion     src/riccati/riccati_recursion.cpp:83-131
//                forwardRiccatiRecursion / computeSwitchingTimeDirection / computeCostateDirection /
//                computeLagrangeMultiplierDirection            src/riccati/riccati_factorizer.cpp:200-281
//
// One CTA (3 warps) per OCP walks i = 0..N.  The sweep is matrix-vector work at ~0.25 FLOP/B, i.e. pure HBM
// streaming: per stage [Fxx|Fvu|Fx] (KKT record) and [P|s|K|k] (Riccati record) arrive by two cp.async.bulk
// copies into a 2-deep ring (stage i+1 in flight while stage i computes), completion on mbarriers.
#pragma once
#include "rbt_device.cuh"
#include "../../include/rbt_layout.h"

namespace rbt {

struct FwdParams {
  rbt_layout L;
  const rbt_stage_ctrl* ctrl;
  int n_grid;
  int batch;
  const double* kkt;
  const double* ric;
  const double* dx0;  // [batch][nx]
  double* dir;        // [batch][n_grid][d_stride]
};

template <int NV, int NU, int NS>
struct FwdCfg {
  static constexpr int NX = 2 * NV;
  static constexpr int NTHREADS = 96;
  static constexpr int KPART = NX * NX + ((NV * NU + 1) & ~1) + ((NX + 1) & ~1);          // Fxx|Fvu|Fx
  static constexpr int RPART = NX * NX + ((NX + 1) & ~1) + ((NU * NX + 1) & ~1) + ((NU + 1) & ~1);  // P|s|K|k
  // extras mirror (per slot): [M|m] [Psi|Phi|T|W] [mt|mtn] [dtsdx|stosc] [fx]
  static constexpr int e_M = 0;
  static constexpr int e_m = e_M + ((NS * NX + 1) & ~1);
  static constexpr int e_Psi = e_m + ((NS + 1) & ~1);
  static constexpr int e_Phi = e_Psi + ((NX + 1) & ~1);
  static constexpr int e_T = e_Phi + ((NX + 1) & ~1);
  static constexpr int e_W = e_T + ((NU + 1) & ~1);
  static constexpr int e_mt = e_W + ((NU + 1) & ~1);
  static constexpr int e_mtn = e_mt + ((NS + 1) & ~1);
  static constexpr int e_pol = e_mtn + ((NS + 1) & ~1);  // dtsdx (nx) | dtsdts, dts0
  static constexpr int e_fx = e_pol + ((NX + 1) & ~1) + 2;
  static constexpr int ESZ = e_fx + ((NX + 1) & ~1);
  static constexpr int SLOT = KPART + RPART + ESZ;
  static constexpr int o_dx = 2 * SLOT;  // dx ping-pong (2 x NX), du (NU), scalars
  static constexpr int o_du = o_dx + 2 * NX;
  static constexpr int o_sc = o_du + ((NU + 1) & ~1);
  static constexpr int o_bar = o_sc + 4;
  static constexpr int SMEM_DOUBLES = o_bar + 4;
  static constexpr size_t SMEM_BYTES = size_t(SMEM_DOUBLES) * 8;
};

template <int NV, int NU, int NS>
__global__ void __launch_bounds__(FwdCfg<NV, NU, NS>::NTHREADS, 3) riccati_forward_kernel(const FwdParams p) {
  using C = FwdCfg<NV, NU, NS>;
  constexpr int NX = C::NX, NTHR = C::NTHREADS;
  constexpr int CO = ((NX + 15) / 16) * 16;  // first thread of the costate group
  static_assert(CO + NX <= NTHR, "costate thread group must fit");
  extern __shared__ __align__(16) double smem[];
  const rbt_layout& L = p.L;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  if (b >= p.batch) return;
  const int N = p.n_grid - 1;
  const double* kkt_b = p.kkt + size_t(b) * p.n_grid * L.k_stride;
  const double* ric_b = p.ric + size_t(b) * p.n_grid * L.r_stride;
  double* dir_b = p.dir + size_t(b) * p.n_grid * L.d_stride;
  double* sdx = smem + C::o_dx;
  double* sdu = smem + C::o_du;
  double* ssc = smem + C::o_sc;  // {dts, dts_next}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::o_bar);  // [0,1] main per slot, [2,3] extras per slot

  auto needs_pol = [&](int st) {
    const rbt_stage_ctrl c = p.ctrl[st];
    return (st == 0 && c.sto) || ((c.type == RBT_IMPACT || c.type == RBT_LIFT) && c.sto_next);
  };
  auto has_extras = [&](int st) {
    const rbt_stage_ctrl c = p.ctrl[st];
    return c.ns > 0 || c.sto || needs_pol(st);
  };
  // one elected thread: all copies of grid point st into ring slot st&1
  auto issue = [&](int st) {
    const int slot = st & 1;
    double* base = smem + slot * C::SLOT;
    const rbt_stage_ctrl c = p.ctrl[st];
    const double* krec = kkt_b + size_t(st) * L.k_stride;
    const double* rrec = ric_b + size_t(st) * L.r_stride;
    fence_proxy_async();
    if (st < N) {
      mbar_expect_tx(&bars[slot], uint32_t(C::KPART + C::RPART) * 8u);
      tma_load_1d(base, krec + L.k_Fxx, uint32_t(C::KPART) * 8u, &bars[slot]);
      tma_load_1d(base + C::KPART, rrec + L.r_P, uint32_t(C::RPART) * 8u, &bars[slot]);
    } else {  // terminal: only P|s
      const uint32_t by = uint32_t(NX * NX + ((NX + 1) & ~1)) * 8u;
      mbar_expect_tx(&bars[slot], by);
      tma_load_1d(base + C::KPART, rrec + L.r_P, by, &bars[slot]);
    }
    if (st < N && has_extras(st)) {
      double* ex = base + C::KPART + C::RPART;
      uint32_t bytes = 0;
      if (c.ns > 0) bytes += uint32_t(C::e_Psi - C::e_M) * 8u;
      if (c.sto) bytes += uint32_t(C::e_mt - C::e_Psi) * 8u + uint32_t(C::ESZ - C::e_fx) * 8u;
      if (c.sto && c.ns > 0) bytes += uint32_t(C::e_pol - C::e_mt) * 8u;
      if (needs_pol(st)) bytes += uint32_t(C::e_fx - C::e_pol) * 8u;
      mbar_expect_tx(&bars[2 + slot], bytes);
      if (c.ns > 0) tma_load_1d(ex + C::e_M, rrec + L.r_M, uint32_t(C::e_Psi - C::e_M) * 8u, &bars[2 + slot]);
      if (c.sto) {
        tma_load_1d(ex + C::e_Psi, rrec + L.r_Psi, uint32_t(C::e_mt - C::e_Psi) * 8u, &bars[2 + slot]);
        tma_load_1d(ex + C::e_fx, krec + L.k_fx, uint32_t(C::ESZ - C::e_fx) * 8u, &bars[2 + slot]);
      }
      if (c.sto && c.ns > 0) tma_load_1d(ex + C::e_mt, rrec + L.r_mt, uint32_t(C::e_pol - C::e_mt) * 8u, &bars[2 + slot]);
      if (needs_pol(st)) tma_load_1d(ex + C::e_pol, rrec + L.r_dtsdx, uint32_t(C::e_fx - C::e_pol) * 8u, &bars[2 + slot]);
    }
  };

  if (tid == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1);
    fence_mbar_init();
  }
  if (tid < NX) sdx[tid] = p.dx0[size_t(b) * NX + tid];
  if (tid == 0) {
    ssc[0] = 0.0;
    ssc[1] = 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    issue(0);
    if (N >= 1) issue(1);
  }
  uint32_t pe0 = 0, pe1 = 0;  // extras-barrier parity per slot

  for (int i = 0; i <= N; ++i) {
    const int slot = i & 1;
    const rbt_stage_ctrl c = p.ctrl[i];
    const double* base = smem + slot * C::SLOT;
    const double* sA = base;
    const double* sB = base + NX * NX;
    const double* sFx = sB + ((NV * NU + 1) & ~1);
    const double* sPm = base + C::KPART;
    const double* ss = sPm + NX * NX;
    const double* sKt = ss + ((NX + 1) & ~1);
    const double* sk = sKt + ((NU * NX + 1) & ~1);
    const double* ex = base + C::KPART + C::RPART;
    double* dx = sdx + (i & 1) * NX;
    double* dxn = sdx + ((i + 1) & 1) * NX;
    double* drec = dir_b + size_t(i) * L.d_stride;
    const bool terminal = (i == N);
    const bool impact = (!terminal && c.type == RBT_IMPACT);
    const bool lift = (!terminal && c.type == RBT_LIFT);
    const bool sto = !terminal && c.sto, sto_next = !terminal && c.sto_next;
    const bool extras = !terminal && has_extras(i);

    mbar_wait(&bars[slot], uint32_t(i >> 1) & 1u);
    if (extras) {
      if (slot) {
        mbar_wait(&bars[3], pe1);
        pe1 ^= 1;
      } else {
        mbar_wait(&bars[2], pe0);
        pe0 ^= 1;
      }
    }

    // ---- switching-time bookkeeping (riccati_recursion.cpp:88-118); entry state = final state of stage i-1
    double dts = ssc[0], dtsn = ssc[1];
    if (i == 0 && sto) {  // :90-93, has_prev_sto_phase = false
      dtsn = dot_serial(ex + C::e_pol, dx, NX) + ex[C::e_pol + ((NX + 1) & ~1) + 1];
    }
    if (lift) {
      dts = dtsn;
      dtsn = 0.0;
      if (sto_next) {
        dtsn = dot_serial(ex + C::e_pol, dx, NX) + ex[C::e_pol + ((NX + 1) & ~1) + 1];
        if (sto) dtsn += ex[C::e_pol + ((NX + 1) & ~1)] * dts;
      }
    }
    if (impact) {
      dts = dtsn;
      dtsn = 0.0;
    }
    const double delta = dtsn - dts;

    if (!terminal && !impact) {
      // du = K dx + k (+ T (dts_next - dts) - W dts_next)           riccati_factorizer.cpp:205-212
      matvec_T(sKt, NX, NX, NU, dx, tid, NTHR, [&](int u, double a) {
        double v = a + sk[u];
        if (sto) {
          v += ex[C::e_T + u] * delta;
          if (sto_next) v -= ex[C::e_W + u] * dtsn;
        }
        sdu[u] = v;
        drec[L.d_du + u] = v;
      });
      __syncthreads();
    }
    if (tid < NX) {
      const int r = tid;
      drec[L.d_dx + r] = dx[r];
      if (!terminal) {
        // dx+ = Fx + A dx (+ [0; Bv du]) (+ fx (dts_next - dts))      :213-218 / :227-229
        double a = sFx[r];
        for (int k = 0; k < NX; ++k) a = fma(sA[r + k * NX], dx[k], a);
        if (!impact) {
          if (r >= NV)
            for (int u = 0; u < NU; ++u) a = fma(sB[(r - NV) + u * NV], sdu[u], a);
          if (sto) a = fma(ex[C::e_fx + r], delta, a);
        }
        dxn[r] = a;
      }
    } else if (tid >= CO && tid < CO + NX && !(impact && sto_next)) {
      // dlmdgmm = P dx - s (+ Psi (dts_next-dts) - Phi dts_next)       :246-266
      const int r = tid - CO;
      double a = -ss[r];
      for (int k = 0; k < NX; ++k) a = fma(sPm[r + k * NX], dx[k], a);
      if (sto) {
        if (impact) {
          a -= ex[C::e_Phi + r] * dtsn;
        } else {
          a += ex[C::e_Psi + r] * delta;
          if (sto_next) a -= ex[C::e_Phi + r] * dtsn;
        }
      }
      drec[L.d_dlmdgmm + r] = a;
    }
    __syncthreads();

    if (impact && sto_next) {
      // impact with an STO phase after it: dts_next comes from the NEW state dx+      riccati_recursion.cpp:100-106
      const double e_dts = dts;  // d[i+1].dts = d[i-1].dts_next
      double e_dtsn = dot_serial(ex + C::e_pol, dxn, NX) + ex[C::e_pol + ((NX + 1) & ~1) + 1];
      if (sto) e_dtsn += ex[C::e_pol + ((NX + 1) & ~1)] * e_dts;
      dts = e_dts;
      dtsn = e_dtsn;
      if (tid >= CO && tid < CO + NX) {
        const int r = tid - CO;
        double a = -ss[r];
        for (int k = 0; k < NX; ++k) a = fma(sPm[r + k * NX], dx[k], a);
        if (sto) a -= ex[C::e_Phi + r] * dtsn;
        drec[L.d_dlmdgmm + r] = a;
      }
    }
    if (!terminal && !impact && c.ns > 0) {
      // dxi = M dx + m (+ mt (dts_next-dts) - mt_next dts_next)        riccati_factorizer.cpp:269-281
      const int ns = c.ns;
      for (int q = tid; q < ns; q += NTHR) {
        double a = ex[C::e_m + q];
        for (int k = 0; k < NX; ++k) a = fma(ex[C::e_M + q + k * ns], dx[k], a);
        if (sto) {
          a += ex[C::e_mt + q] * delta;
          if (sto_next) a -= ex[C::e_mtn + q] * dtsn;
        }
        drec[L.d_dxi + q] = a;
      }
    }
    if (tid == 0) {
      drec[L.d_dts + 0] = dts;
      drec[L.d_dts + 1] = dtsn;
    }
    __syncthreads();  // everyone is done with slot `slot` and with ssc
    if (tid == 0) {
      ssc[0] = dts;
      ssc[1] = dtsn;
      if (i + 2 <= N) issue(i + 2);
    }
    __syncthreads();
  }
}

}  // namespace rbt
