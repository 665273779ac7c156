// Measures the FP64 FMA pipe peak (register-resident DFMA loop) and a plain copy bandwidth on the current GPU.
// MEASURED_PEAKS.json has no FP64 entry; the FP64-pipe roofline in DESIGN.md uses this number.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void copyk(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) dst[i] = src[i];
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  dfma<<<blocks, threads>>>(out, 1000);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  double flops = 2.0 * 8 * iters * (double)blocks * threads;
  printf("{\"fp64_tflops\": %.2f, \"sms\": %d, \"dfma_ms\": %.3f", flops / best / 1e9, p.multiProcessorCount, best);
  size_t n = (size_t)1 << 28;  // 4 GiB per buffer of double2
  double2 *s, *d; cudaMalloc(&s, n * 16); cudaMalloc(&d, n * 16); cudaMemset(s, 1, n * 16);
  best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); copyk<<<p.multiProcessorCount * 16, 512>>>(d, s, n); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf(", \"copy_gbs\": %.1f}\n", 2.0 * n * 16 / best / 1e6);
  return 0;
}
