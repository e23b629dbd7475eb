/*
 * orc_linalg.h -- TEST INFRASTRUCTURE (shared by the oracle sources): small dense GEMM for the CPU restatement.
 *
 * C(m x n) = beta*C + alpha * op(A) * op(B);  ta/tb: 0 = as is, 1 = transposed; column-major, leading dimensions lda/ldb/ldc.
 * The oracle is also the CPU baseline of bench.py, so this has to be a fair stand-in for the reference's Eigen kernels on
 * 6..36-sized blocks: a register-blocked 12 x 4 micro-kernel on 256-bit vectors (12 accumulator registers, A panel loaded
 * once per k, B broadcast), op(A) = A^T handled by transposing the (at most 36 x 36) block once.  Every output element is
 * still accumulated over k in ascending order, like the plain triple loop it replaces.
 */
#ifndef ORC_LINALG_H_
#define ORC_LINALG_H_
#include <stddef.h>
#include <string.h>

#define ORC_IDX(i, j, ld) ((i) + (size_t)(j) * (ld))

typedef double orc_v4d __attribute__((vector_size(32), aligned(8)));

#define ORC_UKERNEL(NAME, MV)                                                                                            \
  static inline void NAME(int k, int nr, const double* restrict A, int lda, const double* restrict B, size_t bsl,        \
                          size_t bsj, double alpha, double beta, double* restrict C, int ldc) {                           \
    orc_v4d acc[4][MV];                                                                                                   \
    for (int j = 0; j < 4; ++j)                                                                                           \
      for (int v = 0; v < MV; ++v) acc[j][v] = (orc_v4d){0.0, 0.0, 0.0, 0.0};                                             \
    if (nr == 4) {                                                                                                        \
      for (int l = 0; l < k; ++l) {                                                                                       \
        const double* a = A + (size_t)l * lda;                                                                            \
        orc_v4d av[MV];                                                                                                   \
        for (int v = 0; v < MV; ++v) av[v] = *(const orc_v4d*)(a + 4 * v);                                                \
        for (int j = 0; j < 4; ++j) {                                                                                     \
          const double b = B[l * bsl + j * bsj];                                                                          \
          const orc_v4d bv = (orc_v4d){b, b, b, b};                                                                       \
          for (int v = 0; v < MV; ++v) acc[j][v] += av[v] * bv;                                                           \
        }                                                                                                                 \
      }                                                                                                                   \
    } else {                                                                                                              \
      for (int l = 0; l < k; ++l) {                                                                                       \
        const double* a = A + (size_t)l * lda;                                                                            \
        orc_v4d av[MV];                                                                                                   \
        for (int v = 0; v < MV; ++v) av[v] = *(const orc_v4d*)(a + 4 * v);                                                \
        for (int j = 0; j < nr; ++j) {                                                                                    \
          const double b = B[l * bsl + j * bsj];                                                                          \
          const orc_v4d bv = (orc_v4d){b, b, b, b};                                                                       \
          for (int v = 0; v < MV; ++v) acc[j][v] += av[v] * bv;                                                           \
        }                                                                                                                 \
      }                                                                                                                   \
    }                                                                                                                     \
    for (int j = 0; j < nr; ++j) {                                                                                        \
      double* c = C + (size_t)j * ldc;                                                                                    \
      for (int v = 0; v < MV; ++v)                                                                                        \
        for (int q = 0; q < 4; ++q) {                                                                                     \
          const double x = alpha * acc[j][v][q];                                                                          \
          c[4 * v + q] = (beta == 0.0) ? x : beta * c[4 * v + q] + x;                                                     \
        }                                                                                                                 \
    }                                                                                                                     \
  }
ORC_UKERNEL(orc_uk12, 3)
ORC_UKERNEL(orc_uk8, 2)
ORC_UKERNEL(orc_uk4, 1)

static void gemm(int ta, int tb, int m, int n, int k, double alpha, const double* restrict A, int lda,
                 const double* restrict B, int ldb, double beta, double* restrict C, int ldc) {
  double At[36 * 48];
  if (ta) {
    if ((size_t)m * k > sizeof(At) / sizeof(double)) {  /* not met by the oracle's block sizes: plain loops */
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
          double acc = 0.0;
          for (int l = 0; l < k; ++l) acc += A[ORC_IDX(l, i, lda)] * (tb ? B[ORC_IDX(j, l, ldb)] : B[ORC_IDX(l, j, ldb)]);
          C[ORC_IDX(i, j, ldc)] = (beta == 0.0 ? 0.0 : beta * C[ORC_IDX(i, j, ldc)]) + alpha * acc;
        }
      return;
    }
    for (int i = 0; i < m; ++i)
      for (int l = 0; l < k; ++l) At[i + (size_t)l * m] = A[ORC_IDX(l, i, lda)];
    A = At;
    lda = m;
  }
  const size_t bsl = tb ? (size_t)ldb : 1, bsj = tb ? 1 : (size_t)ldb;
  for (int j0 = 0; j0 < n; j0 += 4) {
    const int nr = (n - j0 < 4) ? n - j0 : 4;
    const double* Bj = B + j0 * bsj;
    double* Cj = C + (size_t)j0 * ldc;
    int i0 = 0;
    for (; i0 + 12 <= m; i0 += 12) orc_uk12(k, nr, A + i0, lda, Bj, bsl, bsj, alpha, beta, Cj + i0, ldc);
    if (i0 + 8 <= m) { orc_uk8(k, nr, A + i0, lda, Bj, bsl, bsj, alpha, beta, Cj + i0, ldc); i0 += 8; }
    if (i0 + 4 <= m) { orc_uk4(k, nr, A + i0, lda, Bj, bsl, bsj, alpha, beta, Cj + i0, ldc); i0 += 4; }
    for (; i0 < m; ++i0)
      for (int j = 0; j < nr; ++j) {
        double acc = 0.0;
        for (int l = 0; l < k; ++l) acc += A[i0 + (size_t)l * lda] * Bj[l * bsl + j * bsj];
        Cj[i0 + (size_t)j * ldc] = (beta == 0.0 ? 0.0 : beta * Cj[i0 + (size_t)j * ldc]) + alpha * acc;
      }
  }
}
#endif /* ORC_LINALG_H_ */
