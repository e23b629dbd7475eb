// riccati_backward.cuh -- batched backward Riccati recursion, one CTA per OCP instance, sm_100a.
//
// Restates (as a different program) what the reference does in
//   RiccatiRecursion::backwardRiccatiRecursion                 src/riccati/riccati_recursion.cpp:32-80
//   RiccatiFactorizer::backwardRiccatiRecursion (all overloads) src/riccati/riccati_factorizer.cpp:44-197
//   BackwardRiccatiRecursionFactorizer::*                      src/riccati/backward_riccati_recursion_factorizer.cpp:31-174
//
// Design (not a translation):
//  * CTA = one OCP; the serial sweep i = N-1..0 runs inside the kernel with P_next resident in shared memory.
//  * [Fxx|Fvu|Fx|lx|lu] of a stage arrives by ONE cp.async.bulk (TMA 1-D) + mbarrier, issued as soon as the previous
//    stage stops reading the staging buffer, so the copy overlaps the latency-bound Cholesky / solve / symmetrise tail.
//    Qxx/Qxu/Quu never touch shared memory: they are loaded straight into DMMA accumulator fragments.
//  * All dense contractions (A^T P, (A^T P) A, B^T P, .. , Y^T Y) run on the fp64 tensor pipe (mma.sync m8n8k4):
//    warp w owns the 8-row band w of every product; extents that are not multiples of 8 use one pulled-back
//    (overlapping) tile instead of padding, so the bulk copy lands unpadded.
//  * algebra: z = s+ - P+ Fx makes  lu += B^T P+ Fx - B^T s+  ==  lu -= Bv^T z_v  and  s = A^T z - lx - H k;
//    with G = L L^T, Y = L^-1 H^T:  K = -L^-T Y and  P = sym(F - Y^T Y)  (the reference forms G K and K^T (G K)).
//    Rounding differs from the reference at the 1e-15 level; parity tolerance is 1e-6 relative (BASELINE.json).
//  * switching-constraint (Schur) stages and STO terms follow the reference formulas with plain shared-memory loops:
//    they are 2 of 47 stages (trot) / vector-only work.
#pragma once
#include "rbt_device.cuh"
#include "../../include/rbt_layout.h"

namespace rbt {

struct BwdParams {
  rbt_layout L;
  const rbt_stage_ctrl* ctrl;  // device, [n_grid]
  int n_grid;
  int batch;
  double max_dts0;
  const double* kkt;  // [batch][n_grid][k_stride]
  double* ric;        // [batch][n_grid][r_stride]
  double* fact;       // [batch][n_grid][f_stride] or nullptr
  int* info;          // [batch]
  int dbg;            // bring-up only: early-exit level (0 = run normally)
};

template <int NV, int NU, int NS>
struct BwdCfg {
  static constexpr int NX = 2 * NV;
  static constexpr int LDF = NX + 1;  // padded row stride of the F scratch (conflict-free transpose reads)
  static constexpr int TX = num_tiles(NX);
  static constexpr int TU = num_tiles(NU);
  static constexpr int NWARPS = TX;
  static constexpr int NTHREADS = 32 * NWARPS;
  // shared-memory carve-up (doubles)
  static constexpr int STAGE = NX * NX + ((NV * NU + 1) & ~1) + 2 * ((NX + 1) & ~1) + ((NU + 1) & ~1);
  static constexpr int EXTRA = ((NS * NX + 1) & ~1) + ((NS * NU + 1) & ~1) + ((NS + 1) & ~1) + 2 * ((NX + 1) & ~1) +
                               ((NU + 1) & ~1) + ((NS + 1) & ~1) + 4;
  static constexpr int o_P = 0;
  static constexpr int o_AtP = o_P + NX * NX;
  static constexpr int o_In = o_AtP + ((NX * LDF + 1) & ~1);
  static constexpr int o_Ex = o_In + STAGE;
  static constexpr int o_BtP = o_Ex + EXTRA;  // also Y / GK
  static constexpr int o_H = o_BtP + NU * NX;
  static constexpr int o_Kt = o_H + NX * NU;
  static constexpr int o_G = o_Kt + NX * NU;
  static constexpr int o_vec = o_G + NU * NU;
  // vectors
  static constexpr int v_sn = 0, v_z = v_sn + NX, v_t1 = v_z + NX, v_Psin = v_t1 + NX, v_Phin = v_Psin + NX,
                       v_psix = v_Phin + NX, v_phix = v_psix + NX, v_Psi = v_phix + NX, v_Phi = v_Psi + NX,
                       v_Pf = v_Phi + NX, v_lu2 = v_Pf + NX, v_k = v_lu2 + NU, v_psiu = v_k + NU, v_phiu = v_psiu + NU,
                       v_T = v_phiu + NU, v_W = v_T + NU, v_dinv = v_W + NU, v_m = v_dinv + NU, v_mt = v_m + NS,
                       v_mtn = v_mt + NS, v_dinvS = v_mtn + NS, v_Fxs = v_dinvS + NS, v_scn = v_Fxs + NX, v_sc = v_scn + 8,
                       v_end = v_sc + 8;
  static constexpr int o_bar = (o_vec + v_end + 1) & ~1;
  static constexpr int SMEM_DOUBLES = o_bar + 2;
  static constexpr size_t SMEM_BYTES = size_t(SMEM_DOUBLES) * 8;
  // Schur scratch inside the (dead) staging buffer
  static constexpr int x_Ginv = 0, x_DG = x_Ginv + NU * NU, x_S = x_DG + NS * NU, x_SDG = x_S + NS * NS,
                       x_M = x_SDG + NS * NU, x_DtM = x_M + NS * NX, x_Gc = x_DtM + NU * NX, x_end = x_Gc + NU * NU;
  static_assert(x_end <= STAGE, "Schur scratch must fit in the staging buffer");
  static_assert(NX >= 8 && NU >= 8, "DMMA tiling needs extents >= 8");
  static_assert(NX % 2 == 0, "nx even");
};

// In-place lower Cholesky of an n x n (n <= NMAX <= 32) column-major matrix in shared memory by ONE warp
// (lane = row).  dinv[j] = 1 / L_jj.  Returns false on a non-positive pivot (factor then holds garbage).
template <int NMAX>
__device__ __forceinline__ bool warp_cholesky(double* A, int n, double* dinv) {
  const int lane = threadIdx.x & 31;
  double a[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) a[k] = (lane < n && k <= lane && k < n) ? A[lane + k * n] : 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    if (j < n) {
      const double d = __shfl_sync(0xffffffffu, a[j], j);
      if (!(d > 0.0)) ok = false;
      const double sd = sqrt(d);
      const double inv = 1.0 / sd;
      const double lij = (lane == j) ? sd : a[j] * inv;
      a[j] = lij;
      if (lane == j) dinv[j] = inv;
#pragma unroll
      for (int k = j + 1; k < NMAX; ++k) {
        const double lkj = __shfl_sync(0xffffffffu, lij, k < 32 ? k : 0);
        if (k < n && lane >= k) a[k] = fma(-lij, lkj, a[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NMAX; ++k)
    if (lane < n && k <= lane && k < n) A[lane + k * n] = a[k];
  return ok;
}

// b <- (L L^T)^-1 b for a strided vector in shared memory; L column-major n x n, dinv = 1/diag(L).  One thread.
__device__ __forceinline__ void chol_solve_smem(const double* Lm, const double* dinv, int n, double* b, int stride) {
  for (int i = 0; i < n; ++i) {
    double v = b[i * stride];
    for (int k = 0; k < i; ++k) v = fma(-Lm[i + k * n], b[k * stride], v);
    b[i * stride] = v * dinv[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i * stride];
    for (int k = i + 1; k < n; ++k) v = fma(-Lm[k + i * n], b[k * stride], v);
    b[i * stride] = v * dinv[i];
  }
}

template <int NV, int NU, int NS>
__global__ void __launch_bounds__(BwdCfg<NV, NU, NS>::NTHREADS, 4) riccati_backward_kernel(const BwdParams p) {
  using C = BwdCfg<NV, NU, NS>;
  constexpr int NX = C::NX, LDF = C::LDF, TX = C::TX, TU = C::TU, NTHR = C::NTHREADS;
  extern __shared__ __align__(16) double smem[];
  double* sP = smem + C::o_P;
  double* sAtP = smem + C::o_AtP;  // AtP row-major (ld NX); later F scratch row-major (ld LDF)
  double* sIn = smem + C::o_In;
  double* sA = sIn;            // Fxx col-major
  double* sB = sIn + NX * NX;  // Fvu col-major (ld NV)
  double* sEx = smem + C::o_Ex;
  double* sBtP = smem + C::o_BtP;  // BtP row-major (ld NX); later Y / GK col-major (ld NU)
  double* sY = sBtP;
  double* sH = smem + C::o_H;    // H col-major (ld NX)
  double* sKt = smem + C::o_Kt;  // K^T col-major (ld NX)
  double* sG = smem + C::o_G;    // G, then its Cholesky factor
  double* vec = smem + C::o_vec;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::o_bar);

  const rbt_layout& L = p.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x;
  if (b >= p.batch) return;
  const int N = p.n_grid - 1;
  const double* kkt_b = p.kkt + size_t(b) * p.n_grid * L.k_stride;
  double* ric_b = p.ric + size_t(b) * p.n_grid * L.r_stride;
  double* fact_b = p.fact ? p.fact + size_t(b) * p.n_grid * L.f_stride : nullptr;

  const double* sFx = sIn + (L.k_Fx - L.k_Fxx);
  const double* slx = sIn + (L.k_lx - L.k_Fxx);
  const double* slu = sIn + (L.k_lu - L.k_Fxx);
  // extras (valid only on ns>0 / sto stages)
  const double* sPhix = sEx + (L.k_Phix - L.k_Phix);
  const double* sPhiu = sEx + (L.k_Phiu - L.k_Phix);
  const double* sp = sEx + (L.k_p - L.k_Phix);
  const double* sfx = sEx + (L.k_fx - L.k_Phix);
  const double* shx = sEx + (L.k_hx - L.k_Phix);
  const double* shu = sEx + (L.k_hu - L.k_Phix);
  const double* sPhit = sEx + (L.k_Phit - L.k_Phix);
  const double* sksc = sEx + (L.k_sc - L.k_Phix);

  double* s_n = vec + C::v_sn;
  double* z = vec + C::v_z;
  double* t1 = vec + C::v_t1;
  double* Psin = vec + C::v_Psin;
  double* Phin = vec + C::v_Phin;
  double* psix = vec + C::v_psix;
  double* phix = vec + C::v_phix;
  double* Psi = vec + C::v_Psi;
  double* Phi = vec + C::v_Phi;
  double* Pf = vec + C::v_Pf;
  double* lu2 = vec + C::v_lu2;
  double* kv = vec + C::v_k;
  double* psiu = vec + C::v_psiu;
  double* phiu = vec + C::v_phiu;
  double* Tv = vec + C::v_T;
  double* Wv = vec + C::v_W;
  double* dinv = vec + C::v_dinv;
  double* mvec = vec + C::v_m;
  double* mt = vec + C::v_mt;
  double* mtn = vec + C::v_mtn;
  double* dinvS = vec + C::v_dinvS;
  double* Fxs = vec + C::v_Fxs;  // copy of Fx (the staging buffer is reused as Schur scratch)
  double* scn = vec + C::v_scn;  // next-stage {xi, chi, rho, eta, iota}
  double* sc = vec + C::v_sc;

  int bad = 0;  // Cholesky failure flags (thread-local, OR-reduced at the end)

  auto issue_stage_load = [&](int st) {
    // one elected thread: stage st's [Fxx|Fvu|Fx|lx|lu] (+ extras when flagged)
    const rbt_stage_ctrl cs = p.ctrl[st];
    const double* rec = kkt_b + size_t(st) * L.k_stride;
    fence_proxy_async();
    mbar_expect_tx(&bars[0], uint32_t(L.k_stage_size) * 8u);
    tma_load_1d(sIn, rec + L.k_Fxx, uint32_t(L.k_stage_size) * 8u, &bars[0]);
    if (cs.ns > 0 || cs.sto) {
      mbar_expect_tx(&bars[1], uint32_t(L.k_extra_size) * 8u);
      tma_load_1d(sEx, rec + L.k_Phix, uint32_t(L.k_extra_size) * 8u, &bars[1]);
    }
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0 && N > 0) issue_stage_load(N - 1);

  // ---- terminal stage: P_N = Qxx_N, s_N = -lx_N          riccati_recursion.cpp:37-38
  {
    const double* recN = kkt_b + size_t(N) * L.k_stride;
    double* ricN = ric_b + size_t(N) * L.r_stride;
    for (int e = tid; e < NX * NX; e += NTHR) {
      const double v = recN[L.k_Qxx + e];
      sP[e] = v;
      ricN[L.r_P + e] = v;
    }
    for (int e = tid; e < NX; e += NTHR) {
      const double v = -recN[L.k_lx + e];
      s_n[e] = v;
      ricN[L.r_s + e] = v;
      Psin[e] = 0.0;
      Phin[e] = 0.0;
    }
    if (tid < 8) scn[tid] = 0.0;
  }
  __syncthreads();

  uint32_t par0 = 0, par1 = 0;
  if (p.dbg == 1) return;

  for (int i = N - 1; i >= 0; --i) {
    const rbt_stage_ctrl cs = p.ctrl[i];
    const bool impact = (cs.type == RBT_IMPACT);
    const int ns = impact ? 0 : cs.ns;
    const bool sto = cs.sto != 0, sto_next = cs.sto_next != 0;
    const bool extras = (cs.ns > 0 || cs.sto);
    const double* rec = kkt_b + size_t(i) * L.k_stride;
    double* ric = ric_b + size_t(i) * L.r_stride;
    double* fct = fact_b ? fact_b + size_t(i) * L.f_stride : nullptr;

    // ---- phase transition on the in-shared "next" factorization     riccati_recursion.cpp:42-62, riccati_factorizer.cpp:145-175
    {
      bool do_pt = false;
      double* pol = nullptr;
      if (impact) {
        do_pt = (i > 0 && p.ctrl[i - 1].sto) || sto;
        pol = ric;
      } else if (p.ctrl[i + 1].type == RBT_LIFT) {
        do_pt = sto || sto_next;
        pol = ric_b + size_t(i + 1) * L.r_stride;
      }
      if (do_pt) {
        const double xi = scn[0], chi = scn[1], rho = scn[2], eta = scn[3], iota = scn[4];
        __syncthreads();
        if (sto_next) {
          double sgm = xi - 2.0 * chi + rho;
          if ((sgm * p.max_dts0) < fabs(eta - iota) || sgm < 1.4901161193847656e-08 /* sqrt(DBL_EPSILON) */)
            sgm = fabs(sgm) + fabs(eta - iota) / p.max_dts0;
          const double is = 1.0 / sgm;
          if (tid < NX) {
            const double dpp = Psin[tid] - Phin[tid];
            pol[L.r_dtsdx + tid] = -is * dpp;
            s_n[tid] += is * dpp * (eta - iota);
            Phin[tid] = Psin[tid] - is * dpp * (xi - chi);  // m.Phi = Psi - (1/sgm)(Psi-Phi)(xi-chi)
            Psin[tid] = 0.0;
          }
          if (tid == 0) {
            pol[L.r_stosc + 0] = is * (xi - chi);
            pol[L.r_stosc + 1] = -is * (eta - iota);
            scn[0] = 0.0;
            scn[1] = 0.0;
            scn[2] = xi - is * (xi - chi) * (xi - chi);
            scn[3] = 0.0;
            scn[4] = eta - is * (xi - chi) * (eta - iota);
          }
        } else {
          if (tid < NX) {
            Phin[tid] = Psin[tid];
            Psin[tid] = 0.0;
          }
          if (tid == 0) {
            scn[0] = 0.0;
            scn[1] = 0.0;
            scn[2] = xi;
            scn[3] = 0.0;
            scn[4] = eta;
          }
        }
        __syncthreads();
      }
    }

    const int i0 = tile_off(warp, NX);  // this warp's row band

    // Qxx -> accumulator fragments of F (issued before the wait, consumed after GEMM1)
    double cF[TX][2];
#pragma unroll
    for (int n = 0; n < TX; ++n) {
      const int j0 = tile_off(n, NX);
      cF[n][0] = __ldg(rec + L.k_Qxx + (i0 + g) + (j0 + 2 * t) * NX);
      cF[n][1] = __ldg(rec + L.k_Qxx + (i0 + g) + (j0 + 2 * t + 1) * NX);
    }

    // ---- wait for this stage's blocks
    mbar_wait(&bars[0], par0);
    par0 ^= 1;
    if (extras) {
      mbar_wait(&bars[1], par1);
      par1 ^= 1;
    }
    if (p.dbg == 2) return;

    // ================= phase 1: AtP = A^T P+ (all warps), BtP = Bv^T P+[nv:,:] (warps < TU), z = s+ - P+ Fx
    {
      double acc[TX][2];
#pragma unroll
      for (int n = 0; n < TX; ++n) acc[n][0] = acc[n][1] = 0.0;
      warp_mma_band<NX, TX, NX>(
          acc, i0, [&](int ii, int k) { return sA[k + ii * NX]; }, [&](int k, int j) { return sP[k + j * NX]; });
#pragma unroll
      for (int n = 0; n < TX; ++n) {
        const int j0 = tile_off(n, NX);
        *reinterpret_cast<double2*>(&sAtP[(i0 + g) * NX + j0 + 2 * t]) = make_double2(acc[n][0], acc[n][1]);
      }
    }
    if (!impact && warp < TU) {
      const int u0 = tile_off(warp, NU);
      double acc[TX][2];
#pragma unroll
      for (int n = 0; n < TX; ++n) acc[n][0] = acc[n][1] = 0.0;
      warp_mma_band<NV, TX, NX>(
          acc, u0, [&](int u, int k) { return sB[k + u * NV]; }, [&](int k, int j) { return sP[NV + k + j * NX]; });
#pragma unroll
      for (int n = 0; n < TX; ++n) {
        const int j0 = tile_off(n, NX);
        *reinterpret_cast<double2*>(&sBtP[(u0 + g) * NX + j0 + 2 * t]) = make_double2(acc[n][0], acc[n][1]);
      }
    }
    // z = s+ - P+ Fx   (P+ symmetric; thread per row, conflict-free)
    matvec_N(sP, NX, NX, NX, sFx, tid, NTHR, [&](int r, double a) { z[r] = s_n[r] - a; });
    if (sto && !impact) matvec_N(sP, NX, NX, NX, sfx, tid, NTHR, [&](int r, double a) { Pf[r] = a; });
    if (sto && tid < NX) Fxs[tid] = sFx[tid];
    __syncthreads();
    if (p.dbg == 3) return;

    // ================= phase 2: F = Qxx + AtP A ; H = Qxu + AtP[:,nv:] Bv ; G = Quu + BtP[:,nv:] Bv ; vectors
    warp_mma_band<NX, TX, NX>(
        cF, i0, [&](int ii, int k) { return sAtP[ii * NX + k]; }, [&](int k, int j) { return sA[k + j * NX]; });
    if (!impact) {
      {  // H: this warp's row band, TU column tiles
        double cH[TU][2];
#pragma unroll
        for (int n = 0; n < TU; ++n) {
          const int u0 = tile_off(n, NU);
          cH[n][0] = __ldg(rec + L.k_Qxu + (i0 + g) + (u0 + 2 * t) * NX);
          cH[n][1] = __ldg(rec + L.k_Qxu + (i0 + g) + (u0 + 2 * t + 1) * NX);
        }
        warp_mma_band<NV, TU, NU>(
            cH, i0, [&](int ii, int k) { return sAtP[ii * NX + NV + k]; }, [&](int k, int u) { return sB[k + u * NV]; });
#pragma unroll
        for (int n = 0; n < TU; ++n) {
          const int u0 = tile_off(n, NU);
          sH[(i0 + g) + (u0 + 2 * t) * NX] = cH[n][0];
          sH[(i0 + g) + (u0 + 2 * t + 1) * NX] = cH[n][1];
          if (fct) {
            fct[L.f_H + (i0 + g) + (u0 + 2 * t) * NX] = cH[n][0];
            fct[L.f_H + (i0 + g) + (u0 + 2 * t + 1) * NX] = cH[n][1];
          }
        }
      }
      if (warp < TU) {  // G: row band `warp` of BtP[:,nv:] Bv
        const int u0 = tile_off(warp, NU);
        double cG[TU][2];
#pragma unroll
        for (int n = 0; n < TU; ++n) {
          const int v0 = tile_off(n, NU);
          cG[n][0] = __ldg(rec + L.k_Quu + (u0 + g) + (v0 + 2 * t) * NU);
          cG[n][1] = __ldg(rec + L.k_Quu + (u0 + g) + (v0 + 2 * t + 1) * NU);
        }
        warp_mma_band<NV, TU, NU>(
            cG, u0, [&](int u, int k) { return sBtP[u * NX + NV + k]; }, [&](int k, int v) { return sB[k + v * NV]; });
#pragma unroll
        for (int n = 0; n < TU; ++n) {
          const int v0 = tile_off(n, NU);
          sG[(u0 + g) + (v0 + 2 * t) * NU] = cG[n][0];
          sG[(u0 + g) + (v0 + 2 * t + 1) * NU] = cG[n][1];
          if (fct) {
            fct[L.f_G + (u0 + g) + (v0 + 2 * t) * NU] = cG[n][0];
            fct[L.f_G + (u0 + g) + (v0 + 2 * t + 1) * NU] = cG[n][1];
          }
        }
      }
      // lu' = lu - Bv^T z_v      (== lu + BtP Fx - Bv^T s+_v)
      matvec_T(sB, NV, NV, NU, z + NV, tid, NTHR, [&](int u, double a) {
        const double v = slu[u] - a;
        lu2[u] = v;
        if (fct) fct[L.f_lu + u] = v;
      });
    }
    // t1 = A^T z - lx
    matvec_T(sA, NX, NX, NX, z, tid, NTHR, [&](int c, double a) { t1[c] = a - slx[c]; });
    if (sto) {
      if (!impact) {
        // factorizeHamiltonian: backward_riccati_recursion_factorizer.cpp:48-66
        matvec_T(sAtP, NX, NX, NX, sfx, tid, NTHR, [&](int c, double a) { psix[c] = a + shx[c]; });   // AtP fx + hx
        matvec_T(sBtP, NX, NX, NU, sfx, tid, NTHR, [&](int u, double a) { psiu[u] = a + shu[u]; });   // BtP fx + hu
        __syncthreads();
        matvec_T(sA, NX, NX, NX, Psin, tid, NTHR, [&](int c, double a) { psix[c] += a; });
        matvec_T(sB, NV, NV, NU, Psin + NV, tid, NTHR, [&](int u, double a) { psiu[u] += a; });
        if (sto_next) {
          matvec_T(sA, NX, NX, NX, Phin, tid, NTHR, [&](int c, double a) { phix[c] = a; });
          matvec_T(sB, NV, NV, NU, Phin + NV, tid, NTHR, [&](int u, double a) { phiu[u] = a; });
        } else {
          if (tid < NX) phix[tid] = 0.0;
          if (tid < NU) phiu[tid] = 0.0;
        }
      } else {
        // impact + sto: Phi = A^T Phi+      backward_riccati_recursion_factorizer.cpp:160-174
        matvec_T(sA, NX, NX, NX, Phin, tid, NTHR, [&](int c, double a) { Phi[c] = a; });
      }
    }
    __syncthreads();

    // staging buffer is dead on plain stages: prefetch the next stage now (overlaps the serial tail)
    const bool early_prefetch = !extras;
    if (early_prefetch && tid == 0 && i > 0) issue_stage_load(i - 1);
    if (p.dbg == 4) return;

    if (!impact) {
      if (ns == 0) {
        // ================= phase 3: L L^T = G (warp 0)
        if (warp == 0) {
          if (!warp_cholesky<NU>(sG, NU, dinv)) bad |= 1;
        }
        __syncthreads();
        if (p.dbg == 5) return;
        // ================= phase 4: Y = L^-1 H^T, K = -L^-T Y, k = -G^-1 lu', [T, W]
        if (tid < NX + 3) {
          const int c = tid;
          double y[NU];
          bool active = true;
          if (c < NX) {
#pragma unroll
            for (int u = 0; u < NU; ++u) y[u] = sH[c + u * NX];
          } else if (c == NX) {
#pragma unroll
            for (int u = 0; u < NU; ++u) y[u] = lu2[u];
          } else if (c == NX + 1) {
            active = sto;
#pragma unroll
            for (int u = 0; u < NU; ++u) y[u] = sto ? psiu[u] : 0.0;
          } else {
            active = sto && sto_next;
#pragma unroll
            for (int u = 0; u < NU; ++u) y[u] = active ? phiu[u] : 0.0;
          }
          if (active) {
#pragma unroll
            for (int a = 0; a < NU; ++a) {
              double v = y[a];
#pragma unroll
              for (int k = 0; k < a; ++k) v = fma(-sG[a + k * NU], y[k], v);
              y[a] = v * dinv[a];
            }
            if (c < NX) {
#pragma unroll
              for (int u = 0; u < NU; ++u) sY[u + c * NU] = y[u];
            }
#pragma unroll
            for (int a = NU - 1; a >= 0; --a) {
              double v = y[a];
#pragma unroll
              for (int k = a + 1; k < NU; ++k) v = fma(-sG[k + a * NU], y[k], v);
              y[a] = v * dinv[a];
            }
          }
          if (c < NX) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              sKt[c + u * NX] = -y[u];
              ric[L.r_K + c + u * NX] = -y[u];
            }
          } else if (c == NX) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              kv[u] = -y[u];
              ric[L.r_k + u] = -y[u];
            }
          } else if (c == NX + 1) {
            if (sto) {
#pragma unroll
              for (int u = 0; u < NU; ++u) {
                Tv[u] = -y[u];
                ric[L.r_T + u] = -y[u];
              }
            }
          } else {
            if (sto) {
#pragma unroll
              for (int u = 0; u < NU; ++u) {
                Wv[u] = active ? -y[u] : 0.0;
                ric[L.r_W + u] = active ? -y[u] : 0.0;
              }
            }
          }
        }
        __syncthreads();
        if (p.dbg == 6) return;
        // ================= phase 5a: F -= Y^T Y  (tensor pipe), spill F to the scratch for symmetrisation
        warp_mma_band<NU, TX, NX>(
            cF, i0, [&](int ii, int k) { return -sY[k + ii * NU]; }, [&](int k, int j) { return sY[k + j * NU]; });
      } else {
        // ================= Schur-complement path (switching constraint)      riccati_factorizer.cpp:58-89
        double* xs = sIn;  // scratch in the dead staging buffer
        double* Ginv = xs + C::x_Ginv;
        double* DG = xs + C::x_DG;    // ns x nu (ld ns)
        double* Sm = xs + C::x_S;     // ns x ns
        double* SDG = xs + C::x_SDG;  // ns x nu (ld ns)
        double* Mm = xs + C::x_M;     // ns x nx (ld ns)
        double* DtM = xs + C::x_DtM;  // nu x nx (ld nu)
        double* Gc = xs + C::x_Gc;    // copy of G
        for (int e = tid; e < NU * NU; e += NTHR) {
          Gc[e] = sG[e];
          Ginv[e] = ((e % NU) == (e / NU)) ? 1.0 : 0.0;
        }
        for (int e = tid; e < ns * NU; e += NTHR) {  // DG^T <- D^T  (DG[r + u*ns] = D[r + u*ns])
          DG[e] = sPhiu[e];
        }
        __syncthreads();
        if (warp == 0) {
          if (!warp_cholesky<NU>(sG, NU, dinv)) bad |= 1;
        }
        __syncthreads();
        if (tid < NU) chol_solve_smem(sG, dinv, NU, Ginv + tid * NU, 1);            // Ginv = G^-1             :60
        else if (tid < NU + ns) chol_solve_smem(sG, dinv, NU, DG + (tid - NU), ns);  // DGinv^T = G^-1 D^T      :61
        __syncthreads();
        for (int e = tid; e < ns * ns; e += NTHR) {  // S = DGinv D^T                                          :62
          const int r = e % ns, q = e / ns;
          double a = 0.0;
          for (int u = 0; u < NU; ++u) a = fma(DG[r + u * ns], sPhiu[q + u * ns], a);
          Sm[e] = a;
        }
        for (int e = tid; e < ns * NU; e += NTHR) SDG[e] = DG[e];
        for (int e = tid; e < ns * NX; e += NTHR) Mm[e] = sPhix[e];
        if (tid < ns) mvec[tid] = sp[tid];
        if (tid < ns) mt[tid] = sto ? sPhit[tid] : 0.0;
        __syncthreads();
        if (warp == 0) {
          if (!warp_cholesky<NS>(Sm, ns, dinvS)) bad |= 2;
        }
        __syncthreads();
        // SDG = S^-1 DGinv (:65);  M = S^-1 C (:71);  m = S^-1 p (:73);  mt = S^-1 Phit (:116)
        for (int c = tid; c < NU + NX + 2; c += NTHR) {
          if (c < NU) chol_solve_smem(Sm, dinvS, ns, SDG + c * ns, 1);
          else if (c < NU + NX) chol_solve_smem(Sm, dinvS, ns, Mm + (c - NU) * ns, 1);
          else if (c == NU + NX) chol_solve_smem(Sm, dinvS, ns, mvec, 1);
          else if (sto) chol_solve_smem(Sm, dinvS, ns, mt, 1);
        }
        __syncthreads();
        for (int e = tid; e < NU * NU; e += NTHR) {  // Ginv -= SDG^T DGinv                                    :66
          const int a = e % NU, c = e / NU;
          double acc = 0.0;
          for (int r = 0; r < ns; ++r) acc = fma(SDG[r + a * ns], DG[r + c * ns], acc);
          Ginv[e] -= acc;
        }
        __syncthreads();
        // K = -Ginv H^T - SDG^T C (:67-68);  M -= SDG H^T (:72)
        for (int e = tid; e < NU * NX; e += NTHR) {
          const int j = e % NX, u = e / NX;
          double a = 0.0;
          for (int v = 0; v < NU; ++v) a = fma(Ginv[u + v * NU], sH[j + v * NX], a);
          for (int r = 0; r < ns; ++r) a = fma(SDG[r + u * ns], sPhix[r + j * ns], a);
          sKt[j + u * NX] = -a;
          ric[L.r_K + j + u * NX] = -a;
        }
        for (int e = tid; e < ns * NX; e += NTHR) {
          const int r = e % ns, j = e / ns;
          double a = 0.0;
          for (int u = 0; u < NU; ++u) a = fma(SDG[r + u * ns], sH[j + u * NX], a);
          Mm[e] -= a;
        }
        if (tid < NU) {  // k = -Ginv lu' - SDG^T p (:69-70); T, W (:111-114)
          const int u = tid;
          double a = 0.0, tt = 0.0, ww = 0.0;
          for (int v = 0; v < NU; ++v) {
            a = fma(Ginv[u + v * NU], lu2[v], a);
            if (sto) tt = fma(Ginv[u + v * NU], psiu[v], tt);
            if (sto && sto_next) ww = fma(Ginv[u + v * NU], phiu[v], ww);
          }
          for (int r = 0; r < ns; ++r) {
            a = fma(SDG[r + u * ns], sp[r], a);
            if (sto) tt = fma(SDG[r + u * ns], sPhit[r], tt);
          }
          kv[u] = -a;
          ric[L.r_k + u] = -a;
          if (sto) {
            Tv[u] = -tt;
            Wv[u] = -ww;
            ric[L.r_T + u] = -tt;
            ric[L.r_W + u] = -ww;
          }
        } else if (tid >= 32 && tid < 32 + ns) {  // m -= SDG lu' (:74); mt -= SDG psi_u (:117); mt_next = -SDG phi_u (:119)
          const int r = tid - 32;
          double a = 0.0, bq = 0.0, cq = 0.0;
          for (int u = 0; u < NU; ++u) {
            a = fma(SDG[r + u * ns], lu2[u], a);
            if (sto) bq = fma(SDG[r + u * ns], psiu[u], bq);
            if (sto && sto_next) cq = fma(SDG[r + u * ns], phiu[u], cq);
          }
          mvec[r] -= a;
          if (sto) {
            mt[r] -= bq;
            mtn[r] = -cq;
          }
        }
        __syncthreads();
        // outputs M, m (+ mt, mt_next);  GK = G K (:82) -> sY;  DtM = D^T M (:84)
        for (int e = tid; e < ns * NX; e += NTHR) ric[L.r_M + e] = Mm[e];
        if (tid < ns) {
          ric[L.r_m + tid] = mvec[tid];
          if (sto) {
            ric[L.r_mt + tid] = mt[tid];
            ric[L.r_mtn + tid] = mtn[tid];
          }
        }
        for (int e = tid; e < NU * NX; e += NTHR) {
          const int u = e % NU, j = e / NU;
          double a = 0.0, d2 = 0.0;
          for (int v = 0; v < NU; ++v) a = fma(Gc[u + v * NU], sKt[j + v * NX], a);
          for (int r = 0; r < ns; ++r) d2 = fma(sPhiu[r + u * ns], Mm[r + j * ns], d2);
          sY[e] = a;
          DtM[e] = d2;
        }
        __syncthreads();
        // F -= K^T (G K)                                                                     backward_..factorizer.cpp:83
        warp_mma_band<NU, TX, NX>(
            cF, i0, [&](int ii, int k) { return -sKt[ii + k * NX]; }, [&](int k, int j) { return sY[k + j * NU]; });
        if (fct) {  // the reference leaves Qxx = F - K^T G K here; spill it before the constraint correction
#pragma unroll
          for (int n = 0; n < TX; ++n) {
            const int j0 = tile_off(n, NX);
            fct[L.f_F + (i0 + g) + (j0 + 2 * t) * NX] = cF[n][0];
            fct[L.f_F + (i0 + g) + (j0 + 2 * t + 1) * NX] = cF[n][1];
          }
        }
        // P = sym(F) - KtDtM - KtDtM^T = sym(F - 2 K^T DtM)                                   riccati_factorizer.cpp:85-87
        warp_mma_band<NU, TX, NX>(
            cF, i0, [&](int ii, int k) { return -2.0 * sKt[ii + k * NX]; }, [&](int k, int j) { return DtM[k + j * NU]; });
      }
    }

    if (p.dbg == 7) return;
    // ================= phase 5b: spill F (row-major, ld LDF) ; s ; STO scalars
#pragma unroll
    for (int n = 0; n < TX; ++n) {
      const int j0 = tile_off(n, NX);
      sAtP[(i0 + g) * LDF + j0 + 2 * t] = cF[n][0];
      sAtP[(i0 + g) * LDF + j0 + 2 * t + 1] = cF[n][1];
    }
    if (tid < NX) {
      const int r = tid;
      double v = t1[r];
      if (!impact) {
#pragma unroll
        for (int u = 0; u < NU; ++u) v = fma(-sH[r + u * NX], kv[u], v);
        if (ns > 0) {
          for (int q = 0; q < ns; ++q) v = fma(-sPhix[q + r * ns], mvec[q], v);  // s -= C^T m    riccati_factorizer.cpp:88
        }
      }
      s_n[r] = v;  // old s+ is dead (folded into z)
      ric[L.r_s + r] = v;
    }
    if (sto) {
      if (!impact) {
        // factorizeSTOFactorization: backward_riccati_recursion_factorizer.cpp:94-143 (+ riccati_factorizer.cpp:136-141)
        if (tid >= 64 && tid < 64 + NX) {
          const int c = tid - 64;
          double a = psix[c], bq = sto_next ? phix[c] : 0.0;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            a = fma(sKt[c + u * NX], psiu[u], a);
            if (sto_next) bq = fma(sKt[c + u * NX], phiu[u], bq);
          }
          if (ns > 0) {
            const double* Mm = sIn + C::x_M;
            for (int q = 0; q < ns; ++q) a = fma(Mm[q + c * ns], sPhit[q], a);
          }
          Psi[c] = a;
          Phi[c] = bq;
        }
        if (tid == 128) {
          double xi = dot_serial(sfx, Pf, NX) + sksc[0] + 2.0 * dot_serial(Psin, sfx, NX) + dot_serial(Tv, psiu, NU) + scn[0];
          double chi = 0.0, rho = 0.0, iota = 0.0;
          if (sto_next) {
            chi = sksc[1] + dot_serial(Phin, sfx, NX) + dot_serial(Tv, phiu, NU) + scn[1];
            rho = dot_serial(Wv, phiu, NU) + scn[2];
          }
          // Pf2 = P+ Fx - s+ = -z
          double eta = -dot_serial(sfx, z, NX) + sksc[2] + dot_serial(Psin, Fxs, NX) + dot_serial(psiu, kv, NU) + scn[3];
          if (sto_next) iota = dot_serial(Phin, Fxs, NX) + dot_serial(phiu, kv, NU) + scn[4];
          if (ns > 0) {
            xi += dot_serial(mt, sPhit, ns);
            if (sto_next) chi += dot_serial(mtn, sPhit, ns);
            eta += dot_serial(mvec, sPhit, ns);
          }
          sc[0] = xi; sc[1] = chi; sc[2] = rho; sc[3] = eta; sc[4] = iota;
        }
      } else {
        if (tid == 128) {
          sc[0] = 0.0; sc[1] = 0.0; sc[2] = scn[2]; sc[3] = 0.0;
          sc[4] = scn[4] + dot_serial(Phin, Fxs, NX);
        }
        if (tid < NX) Psi[tid] = 0.0;
      }
    }
    __syncthreads();

    if (p.dbg == 8) return;
    // ================= phase 6: P = (F + F^T)/2 -> shared (next stage) and HBM ; roll the STO state
    for (int e = tid; e < NX * NX; e += NTHR) {
      const int r = e % NX, c = e / NX;
      const double f_rc = sAtP[r * LDF + c];
      const double v = 0.5 * (f_rc + sAtP[c * LDF + r]);
      sP[e] = v;
      ric[L.r_P + e] = v;
      if (fct && ns == 0) fct[L.f_F + e] = f_rc;
    }
    if (sto) {
      if (tid < NX) {
        Psin[tid] = Psi[tid];
        Phin[tid] = Phi[tid];
        ric[L.r_Psi + tid] = Psi[tid];
        ric[L.r_Phi + tid] = Phi[tid];
        if (!impact) {
          ric[L.r_psix + tid] = psix[tid];
          ric[L.r_phix + tid] = phix[tid];
        }
      }
      if (!impact && tid < NU) {
        ric[L.r_psiu + tid] = psiu[tid];
        ric[L.r_phiu + tid] = phiu[tid];
      }
      if (tid < 5) {
        scn[tid] = sc[tid];
        ric[L.r_sc + tid] = sc[tid];
      }
    } else {
      // !sto: Psi = 0, xi = chi = eta = 0 (riccati_factorizer.cpp:99-105); Phi, rho, iota keep their constructor zeros
      if (tid < NX) {
        Psin[tid] = 0.0;
        Phin[tid] = 0.0;
      }
      if (tid < 8) scn[tid] = 0.0;
    }
    __syncthreads();
    if (!early_prefetch && tid == 0 && i > 0) issue_stage_load(i - 1);
    if (p.dbg == 9) return;
  }

  // ---- final phase transition at stage 0                       riccati_recursion.cpp:75-79
  {
    const rbt_stage_ctrl c0 = p.ctrl[0];
    if (c0.sto && c0.sto_next && N > 0) {
      const double xi = scn[0], chi = scn[1], rho = scn[2], eta = scn[3], iota = scn[4];
      double sgm = xi - 2.0 * chi + rho;
      if ((sgm * p.max_dts0) < fabs(eta - iota) || sgm < 1.4901161193847656e-08) sgm = fabs(sgm) + fabs(eta - iota) / p.max_dts0;
      const double is = 1.0 / sgm;
      if (tid < NX) ric_b[L.r_dtsdx + tid] = -is * (Psin[tid] - Phin[tid]);
      if (tid == 0) {
        ric_b[L.r_stosc + 0] = is * (xi - chi);
        ric_b[L.r_stosc + 1] = -is * (eta - iota);
      }
    }
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (lane == 0 && bad) atomicOr(&p.info[b], bad);
}

}  // namespace rbt
