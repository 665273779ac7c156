// Probe: can the B200 scheduler issue integer/FP32 instructions in the shadow of half-rate FP64 instructions?
// Measures cycles per DFMA for (a) DFMA only, (b) DFMA + 1 independent IMAD each, (c) DFMA + 2 IMAD each,
// at 1, 2 and 4 warps per SM sub-partition. Output: JSON lines.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI>
__global__ void probe(double* out, int* iout, int iters, long long* cyc) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 5, i5 = i0 + 7, i6 = i0 + 11, i7 = i0 + 13;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      a0 = fma(a0, b, c); if (NI > 0) i0 = i0 * 3 + i1; if (NI > 1) i4 = i4 * 5 + i0;
      a1 = fma(a1, b, c); if (NI > 0) i1 = i1 * 3 + i2; if (NI > 1) i5 = i5 * 5 + i1;
      a2 = fma(a2, b, c); if (NI > 0) i2 = i2 * 3 + i3; if (NI > 1) i6 = i6 * 5 + i2;
      a3 = fma(a3, b, c); if (NI > 0) i3 = i3 * 3 + i0; if (NI > 1) i7 = i7 * 5 + i3;
      a4 = fma(a4, b, c); if (NI > 0) i0 = i0 * 7 + i2; if (NI > 1) i4 = i4 * 9 + i5;
      a5 = fma(a5, b, c); if (NI > 0) i1 = i1 * 7 + i3; if (NI > 1) i5 = i5 * 9 + i6;
      a6 = fma(a6, b, c); if (NI > 0) i2 = i2 * 7 + i0; if (NI > 1) i6 = i6 * 9 + i7;
      a7 = fma(a7, b, c); if (NI > 0) i3 = i3 * 7 + i1; if (NI > 1) i7 = i7 * 9 + i4;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI> void run(int warps_per_sm, int sms) {
  double* out; int* iout; long long* cyc;
  int threads = 32 * warps_per_sm, iters = 4000;
  cudaMalloc(&out, 8 * sms * threads); cudaMalloc(&iout, 4 * sms * threads); cudaMalloc(&cyc, 8);
  probe<NI><<<sms, threads>>>(out, iout, 100, cyc);
  probe<NI><<<sms, threads>>>(out, iout, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_smsp_dfma = (double)iters * 32 * (warps_per_sm / 4.0);   // DFMA warp-instructions per sub-partition
  printf("{\"imad_per_dfma\": %d, \"warps_per_smsp\": %d, \"cycles_per_dfma_per_smsp\": %.3f}\n", NI, warps_per_sm / 4,
         h / per_smsp_dfma);
  cudaFree(out); cudaFree(iout); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int w : {4, 8, 16}) { run<0>(w, p.multiProcessorCount); run<1>(w, p.multiProcessorCount); run<2>(w, p.multiProcessorCount); }
  return 0;
}
