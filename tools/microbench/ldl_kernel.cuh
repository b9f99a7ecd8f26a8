#include <math.h>
#define PIVOT_TOL 1e-14
template <int NR, bool TWO>
__device__ __forceinline__ void warp_ldl_regs(double* __restrict__ F, int nu, double* __restrict__ Gm, int nc, double* __restrict__ Ks,
                                              double* __restrict__ invd, const double* __restrict__ Rm, int* ok) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int ncols = nu + nc, c0 = lane, c1 = lane + 32;
  double a[NR], b[TWO ? NR : 1];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double v = 0.0;
    if (r < nu && c0 < ncols) v = c0 < nu ? F[r * nu + c0] : Gm[r * nc + (c0 - nu)];
    a[r] = v;
    if (TWO) b[r] = (r < nu && c1 < ncols) ? Gm[r * nc + (c1 - nu)] : 0.0;
  }
  double myinv = 0.0;
  bool good = true;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (j < nu) {  // warp-uniform
      double d = __shfl_sync(full, a[j], j);
      if (!(d > PIVOT_TOL * fmax(1.0, fabs(Rm[j * nu + j])))) good = false, d = 1.0;
      const double inv = 1.0 / d;
      if (lane == j) myinv = inv;
      if (lane == 0) invd[j] = inv;
      const double sa = c0 > j ? a[j] * inv : 0.0;
      const double sb = TWO ? b[j] * inv : 0.0;
#pragma unroll
      for (int r = j + 1; r < NR; ++r) {  // rows >= nu are zero padding: no branch inside the pivot step
        const double lr = __shfl_sync(full, a[r], j);
        a[r] = fma(-lr, sa, a[r]);
        if (TWO) b[r] = fma(-lr, sb, b[r]);
      }
    }
  }
  if (!good && lane == 0) *ok = 0;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r < nu) {
      const double ir = invd[r];
      if (c0 < nu) {
        if (r > c0) F[r * nu + c0] = a[r] * myinv;
      } else if (c0 < ncols) {
        Gm[r * nc + (c0 - nu)] = a[r];
        Ks[r * nc + (c0 - nu)] = -a[r] * ir;
      }
      if (TWO && c1 < ncols) {
        Gm[r * nc + (c1 - nu)] = b[r];
        Ks[r * nc + (c1 - nu)] = -b[r] * ir;
      }
    }
  }
}
template <int NR, bool TWO>
__global__ void k(double* F, int nu, double* Gm, int nc, double* Ks, double* invd, const double* Rm, int* ok, long long* t) {
  long long t0 = clock64();
  warp_ldl_regs<NR, TWO>(F, nu, Gm, nc, Ks, invd, Rm, ok);
  long long t1 = clock64();
  if (threadIdx.x == 0) *t = t1 - t0;
}
template __global__ void k<32, true>(double*, int, double*, int, double*, double*, const double*, int*, long long*);
template __global__ void k<24, true>(double*, int, double*, int, double*, double*, const double*, int*, long long*);
template __global__ void k<20, true>(double*, int, double*, int, double*, double*, const double*, int*, long long*);
template __global__ void k<8, false>(double*, int, double*, int, double*, double*, const double*, int*, long long*);
