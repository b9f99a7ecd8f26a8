// micro-benchmark of the Riccati stage factorisation kernels (one warp in registers vs the CTA-cooperative version)
#include <cstdio>
#include <vector>
#include <cmath>
#include "ldl_cta.cuh"
template <int KMAX>
__global__ void kc(double* gF, int nu, double* gG, int nc, double* gK, double* invd, const double* gR, int* ok, long long* t) {
  __shared__ double F[32 * 32], G[32 * 40], K[32 * 40], R[32 * 32];
  __shared__ LdlBuf B;
  for (int i = threadIdx.x; i < nu * nu; i += blockDim.x) F[i] = gF[i], R[i] = gR[i];
  for (int i = threadIdx.x; i < nu * nc; i += blockDim.x) G[i] = gG[i];
  __syncthreads();
  long long t0 = clock64();
  cta_ldl<KMAX>(F, nu, G, nc, K, invd, R, ok, &B);
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) *t = t1 - t0;
  for (int i = threadIdx.x; i < nu * nu; i += blockDim.x) gF[i] = F[i];
  for (int i = threadIdx.x; i < nu * nc; i += blockDim.x) gG[i] = G[i], gK[i] = K[i];
}
int main() {
  for (int nu : {5, 20, 32}) {
    int nc = nu == 5 ? 9 : 30;
    std::vector<double> F(nu * nu), G(nu * nc), R(nu * nu);
    for (int i = 0; i < nu; ++i) for (int j = 0; j < nu; ++j) F[i * nu + j] = (i == j ? 10.0 + i : 0.0) + 1.0 / (1 + abs(i - j)), R[i * nu + j] = F[i * nu + j];
    for (int i = 0; i < nu * nc; ++i) G[i] = 0.01 * (i % 17) - 0.05;
    // host reference
    std::vector<double> Fh = F, Gh = G, Kh(nu * nc), inv(nu);
    for (int j = 0; j < nu; ++j) {
      inv[j] = 1.0 / Fh[j * nu + j];
      for (int r = j + 1; r < nu; ++r) {
        double lr = Fh[r * nu + j];
        for (int c = j + 1; c < nu; ++c) Fh[r * nu + c] -= lr * (Fh[j * nu + c] * inv[j]);
        for (int c = 0; c < nc; ++c) Gh[r * nc + c] -= lr * (Gh[j * nc + c] * inv[j]);
      }
      for (int r = j + 1; r < nu; ++r) Fh[r * nu + j] *= inv[j];
    }
    double *dF, *dG, *dK, *dI, *dR; int* ok; long long* t;
    cudaMalloc(&dF, F.size() * 8); cudaMalloc(&dG, G.size() * 8); cudaMalloc(&dK, G.size() * 8); cudaMalloc(&dI, 64 * 8); cudaMalloc(&dR, R.size() * 8); cudaMalloc(&ok, 4); cudaMalloc(&t, 8);
    cudaMemcpy(dR, R.data(), R.size() * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemcpy(dF, F.data(), F.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice);
      if (nu <= 8) kc<2><<<1, 256>>>(dF, nu, dG, nc, dK, dI, dR, ok, t);
      else if (nu <= 24) kc<6><<<1, 256>>>(dF, nu, dG, nc, dK, dI, dR, ok, t);
      else kc<8><<<1, 256>>>(dF, nu, dG, nc, dK, dI, dR, ok, t);
      long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
      std::vector<double> Fo(nu * nu), Go(nu * nc);
      cudaMemcpy(Fo.data(), dF, Fo.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Go.data(), dG, Go.size() * 8, cudaMemcpyDeviceToHost);
      double eL = 0, eG = 0;
      for (int r = 0; r < nu; ++r) for (int c = 0; c < r; ++c) eL = fmax(eL, fabs(Fo[r * nu + c] - Fh[r * nu + c]));
      for (int i = 0; i < nu * nc; ++i) eG = fmax(eG, fabs(Go[i] - Gh[i]));
      printf("nu=%d cta_ldl cycles %lld  errL %.2e errK %.2e\n", nu, h, eL, eG);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
