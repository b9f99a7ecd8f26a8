// CTA-cooperative LDL' of the augmented Riccati stage matrix [F | Gm] (development copy of the kernel in csrc/obca_kkt.h).
// 256 threads = 4 row groups x 64 columns; every thread keeps its entries (rows rg, rg+4, ...; one column) in registers;
// per pivot the pivot row and column are published through double-buffered shared memory: ONE barrier per pivot.
#include <math.h>
#ifndef PIVOT_TOL
#define PIVOT_TOL 1e-14
#endif
// 1 / d for a positive, normal d: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps (error 2^-80 < half an ulp of the
// product d * x; the result is within 1 ulp, which the refinement-free LDL' does not notice)
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}
struct LdlBuf {
  double col[2][32], row[2][64], inv[2], invd[32];
  int bad;
};
template <int KMAX>
__device__ __forceinline__ void cta_ldl(double* __restrict__ F, int nu, double* __restrict__ Gm, int nc, double* __restrict__ Ks,
                                        double* __restrict__ invd_out, const double* __restrict__ Rm, int* ok, LdlBuf* B) {
  const int t = threadIdx.x, rg = t >> 6, c = t & 63, ncols = nu + nc;
  double m[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int r = rg + 4 * k;
    m[k] = (r < nu && c < ncols) ? (c < nu ? F[r * nu + c] : Gm[r * nc + (c - nu)]) : 0.0;
  }
  // every column has one diagonal owner: thread (rg = c & 3, c); its pivot tolerance is known before the loop
  const double mytol = (c < nu && rg == (c & 3)) ? PIVOT_TOL * fmax(1.0, fabs(Rm[c * nu + c])) : 0.0;
  if (t == 0) B->bad = 0;
  __syncthreads();
  // publish pivot 0
  if (c == 0) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) B->col[0][rg + 4 * k] = m[k];
  }
  if (rg == 0) B->row[0][c] = m[0];
  if (t == 0) {
    double d = m[0];
    if (!(d > mytol)) B->bad = 1, d = 1.0;
    const double inv = fast_rcp(d);
    B->inv[0] = inv, B->invd[0] = inv;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4 * KMAX; ++j) {
    if (j < nu) {  // CTA-uniform
      const int p = j & 1;
      const double uc = c > j ? B->row[p][c] * B->inv[p] : 0.0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (4 * k + 3 > j) {  // static: some row of this slot is below the pivot
          const int r = rg + 4 * k;
          const double l = r > j ? B->col[p][r] : 0.0;
          m[k] = fma(-l, uc, m[k]);
        }
      }
      if (j + 1 < nu) {
        constexpr int dummy = 0;
        (void)dummy;
        const int jn = j + 1, kn = jn >> 2, gn = jn & 3;
        if (c == jn) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (4 * k + 3 > jn) B->col[p ^ 1][rg + 4 * k] = m[k];
        }
        if (rg == gn) {
          B->row[p ^ 1][c] = m[kn < KMAX ? kn : 0];
          if (c == jn) {
            double d = m[kn < KMAX ? kn : 0];
            if (!(d > mytol)) B->bad = 1, d = 1.0;
            const double inv = fast_rcp(d);
            B->inv[p ^ 1] = inv, B->invd[jn] = inv;
          }
        }
      }
      __syncthreads();
    }
  }
  // write back: unit L below the diagonal, Khat, Ks = -D^-1 Khat
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int r = rg + 4 * k;
    if (r < nu && c < ncols) {
      if (c < nu) {
        if (r > c) F[r * nu + c] = m[k] * B->invd[c];
      } else {
        Gm[r * nc + (c - nu)] = m[k];
        Ks[r * nc + (c - nu)] = -m[k] * B->invd[r];
      }
    }
  }
  if (t < nu) invd_out[t] = B->invd[t];
  if (t == 0 && B->bad) *ok = 0;
}
