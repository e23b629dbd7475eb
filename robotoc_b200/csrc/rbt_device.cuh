// rbt_device.cuh -- sm_100a device helpers: 1-D TMA bulk copies + mbarrier, fp64 tensor-pipe tiles
// (mma.sync.m8n8k4.f64, DMMA), small warp-parallel dense helpers.  fp64 has no tcgen05 kind, so the
// tensor pipe is reached through DMMA; measured on B200: 37.1 TFLOP/s DMMA vs 33.5 DFMA vs ~18 for
// shared-memory-fed DFMA (gpurun_out/fp64_microbench.txt, tools/fp64_microbench.cu).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace rbt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier / TMA (cp.async.bulk, 1-D) -------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spins on try_wait (HW-suspended wait).  A bulk copy that never completes would hang the GPU box, so the spin is
// bounded: after ~2^24 failed probes (seconds) the kernel traps and the host sees a CUDA error instead of a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  uint32_t tries = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++tries > (1u << 24)) {
      if ((threadIdx.x & 31) == 0)
        printf("[rbt] mbarrier wait timed out: block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x,
               smem_u32(bar), parity);
      __trap();
    }
  } while (!ok);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned.  SASS: UBLKCP.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- DMMA m8n8k4: D(8x8) += A(8x4) * B(4x8), fp64.  lane = 4*g + t:
//   a = A[g][t],  b = B[t][g],  c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // volatile: the instruction is warp-convergent; it must never be cloned into the arms of a per-thread branch
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Bulk L2 prefetch (no shared-memory destination, no completion tracking): bytes multiple of 16, 16-byte aligned source.
__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// Tile origin for an extent M covered by 8-wide tiles: full tiles, then one tile pulled back to end at M
// (overlapping rows/cols are simply computed twice with identical operands -> identical bits).
__host__ __device__ constexpr int tile_off(int t, int M) { return (8 * t + 8 <= M) ? 8 * t : (M >= 8 ? M - 8 : 0); }
__host__ __device__ constexpr int num_tiles(int M) { return (M + 7) / 8; }

// One warp accumulates NT 8x8 tiles of one 8-row band:  acc[n] += sum_k A(i0+g, k) * B(k, joff(n)+g).
//   fa(i, k) / fb(k, j) return the operand element (shared-memory loads).  K need not be a multiple of 4.
template <int K, int NT, int N, class FA, class FB>
__device__ __forceinline__ void warp_mma_band(double (&acc)[NT][2], int i0, FA fa, FB fb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 4) {
    // K tail: load from a clamped (valid) index and zero the A operand arithmetically -- no per-thread branch
    // may surround the warp-convergent MMA.
    const int k = (K % 4 == 0) ? (k0 + t) : min(k0 + t, K - 1);
    const double msk = ((K % 4 == 0) || (k0 + t < K)) ? 1.0 : 0.0;
    const double a = fa(i0 + g, k) * msk;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const double b = fb(k, tile_off(n, N) + g);
      dmma884(acc[n][0], acc[n][1], a, b);
    }
  }
}

// ---- small dense helpers over shared memory (all threads of the CTA cooperate; tid/nthr given) ----
// y[c] = sum_k A[k + c*lda] * x[k]  (A^T x), 4 lanes per column, conflict-free for lda % 4 == 0 or 2.
// Call with all threads; uses shuffles inside aligned groups of 4 lanes.  Result written by lane q==0.
template <class Epi>
__device__ __forceinline__ void matvec_T(const double* A, int lda, int K, int C, const double* x, int tid, int nthr,
                                         Epi epi) {
  const int q = tid & 3;
  const int ngroups = nthr >> 2;
  const int Cpad = (C + ngroups - 1) / ngroups * ngroups;
  for (int c = tid >> 2; c < Cpad; c += ngroups) {
    double acc = 0.0;
    if (c < C)
      for (int k = q; k < K; k += 4) acc = fma(A[k + c * lda], x[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && c < C) epi(c, acc);
  }
}

// y[r] = sum_k A[r + k*lda] * x[k]  (A x), one thread per row (conflict-free: consecutive lanes, consecutive rows).
template <class Epi>
__device__ __forceinline__ void matvec_N(const double* A, int lda, int R, int K, const double* x, int tid, int nthr,
                                         Epi epi) {
  for (int r = tid; r < R; r += nthr) {
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(A[r + k * lda], x[k], acc);
    epi(r, acc);
  }
}

// y[r] = sum_k A[r + k*lda] * x[k]  (A x) with 4 lanes per row (k interleaved) + 2 shuffles: 4x shorter dependent chain
// than matvec_N.  Call with whole warps (nthr multiple of 32).
template <class Epi>
__device__ __forceinline__ void matvec_N4(const double* A, int lda, int R, int K, const double* x, int tid, int nthr,
                                          Epi epi) {
  const int q = tid & 3;
  const int ngroups = nthr >> 2;
  const int Rpad = (R + ngroups - 1) / ngroups * ngroups;
  for (int r = tid >> 2; r < Rpad; r += ngroups) {
    double acc = 0.0;
    if (r < R)
      for (int k = q; k < K; k += 4) acc = fma(A[r + k * lda], x[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && r < R) epi(r, acc);
  }
}

__device__ __forceinline__ double dot_serial(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s = fma(a[i], b[i], s);
  return s;
}

}  // namespace rbt
