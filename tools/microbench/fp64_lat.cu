// Developer micro-benchmark: latencies that bound the OBCA kernels on B200 (dependent DFMA, shared-memory load-to-use,
// warp shuffle, CTA barrier) and the DFMA throughput peak.  nvcc -arch=sm_100a -O3 fp64_lat.cu -o fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double* out, int n, long long* cyc) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  out[3] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_smem_chain(int* idx0, int n, long long* cyc) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 17 + 1) % 1024);
  __syncthreads();
  double v = (double)idx0[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) v = sm[(int)v];  // load -> cvt -> address
  long long t1 = clock64();
  idx0[1] = (int)v;
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_smem_fma_chain(double* out, int n, long long* cyc) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double acc = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) acc = fma(sm[(i * 33 + threadIdx.x) & 1023], acc, sm[(i * 7) & 1023]);  // dot-product-like: loads independent of the chain
  long long t1 = clock64();
  out[3] = acc;
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_shfl_chain(double* out, int n, long long* cyc) {
  double v = out[threadIdx.x & 3];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) v += __shfl_xor_sync(0xffffffffu, v, 1);
  long long t1 = clock64();
  out[4] = v;
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
__global__ void k_barrier(int n, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
}
__global__ void k_store_load(double* out, int n, long long* cyc) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = 1.0;
  __syncthreads();
  volatile double* p = sm;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) p[threadIdx.x & 63] = p[threadIdx.x & 63] * 1.0000001 + 1e-9;  // load -> fma -> store -> load ...
  long long t1 = clock64();
  out[5] = p[0];
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
}
__global__ void k_div_sqrt(double* out, int n, long long* cyc) {
  double a = out[0] + 3.0, b = out[1] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = b / a + 1.0;
  long long t1 = clock64();
  for (int i = 0; i < n; ++i) b = sqrt(b + 2.0);
  long long t2 = clock64();
  out[6] = a + b;
  if (threadIdx.x == 0) cyc[6] = t1 - t0, cyc[7] = t2 - t1;
}
__global__ void k_peak(double* out, int n) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < n; ++i) {
    a0 = fma(a0, b, c), a1 = fma(a1, b, c), a2 = fma(a2, b, c), a3 = fma(a3, b, c);
    a4 = fma(a4, b, c), a5 = fma(a5, b, c), a6 = fma(a6, b, c), a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  double* out;
  long long* cyc;
  int* idx;
  cudaMalloc(&out, 1 << 24);
  cudaMemset(out, 0, 1 << 24);
  cudaMalloc(&cyc, 64);
  cudaMalloc(&idx, 64);
  cudaMemset(idx, 0, 64);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    k_dfma_chain<<<1, 32>>>(out, n, cyc);
    k_smem_chain<<<1, 32>>>(idx, n, cyc);
    k_smem_fma_chain<<<1, 32>>>(out, n, cyc);
    k_shfl_chain<<<1, 32>>>(out, n, cyc);
    k_barrier<<<1, 256>>>(n, cyc);
    k_store_load<<<1, 32>>>(out, n, cyc);
    k_div_sqrt<<<1, 32>>>(out, n, cyc);
  }
  long long h[8];
  cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  printf("dependent DFMA           %.1f cycles\n", (double)h[0] / n);
  printf("smem load->addr chain    %.1f cycles (load + F2I)\n", (double)h[1] / n);
  printf("fma(smem, acc, smem)     %.1f cycles per step (loads off the chain)\n", (double)h[2] / n);
  printf("shfl + dadd chain        %.1f cycles\n", (double)h[3] / n);
  printf("__syncthreads (256 thr)  %.1f cycles\n", (double)h[4] / n);
  printf("smem load-fma-store loop %.1f cycles (volatile)\n", (double)h[5] / n);
  printf("dependent DDIV (+DADD)   %.1f cycles\n", (double)h[6] / n);
  printf("dependent DSQRT (+DADD)  %.1f cycles\n", (double)h[7] / n);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int blocks = p.multiProcessorCount * 8, thr = 256, it = 1 << 14;
  k_peak<<<blocks, thr>>>(out, it);
  cudaEventRecord(e0);
  k_peak<<<blocks, thr>>>(out, it);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("DFMA peak                %.2f TFLOP/s (%d SMs, %d MHz)\n", 2.0 * 8 * it * (double)blocks * thr / (ms * 1e-3) / 1e12, p.multiProcessorCount, p.clockRate / 1000);
  return 0;
}
