// Probe: FP64 issue rate on B200 as a function of the number of distinct 64-bit REGISTER source operands.
// tools/fp64_peak.cu reaches ~2.1 cycles per warp-DFMA per sub-partition with ONE register source (the other two are
// constants). The rhs! kernels sit at 50 % of that peak whatever the occupancy; this checks whether DFMA/DADD/DMUL with
// 2 or 3 register sources issue at the same rate. Output: JSON lines (cycles per FP64 warp-instruction per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a = fma(a, B, C) constants      1: a = fma(a, b_i, C)      2: a = fma(a, b_i, c_i)
// MODE 3: a = fma(b_i, c_i, a)            4: a = a + b_i             5: a = a * b_i
// MODE 6: a = fma(b_i, K, a) (one register + constant-bank weight + accumulator: the flux-differencing accumulate)
// MODE 7: a = fma(a, a, c_i)              8: a_i = fma(a_j, b_i, c_i) (source rotates: no same-register reuse)
template <int MODE>
__global__ void probe(double* out, const double* in, int iters, long long* cyc) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = in[i] + threadIdx.x * 1e-3; b[i] = in[8 + i]; c[i] = in[16 + i]; }
  const double B = 1.0000001, C = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fma(a[i], B, C);
        if (MODE == 1) a[i] = fma(a[i], b[i], C);
        if (MODE == 2) a[i] = fma(a[i], b[i], c[i]);
        if (MODE == 3) a[i] = fma(b[i], c[i], a[i]);
        if (MODE == 4) a[i] = a[i] + b[i];
        if (MODE == 5) a[i] = a[i] * b[i];
        if (MODE == 6) a[i] = fma(b[i], B, a[i]);
        if (MODE == 7) a[i] = fma(a[i], a[i], c[i]);
      }
      if (MODE == 8) {
        double n[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) n[i] = fma(a[(i + 3) & 7], b[i], c[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = n[i];
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(int warps_per_sm, int sms, const double* in) {
  double* out; long long* cyc;
  int threads = 32 * warps_per_sm, iters = 4000;
  cudaMalloc(&out, 8 * sms * threads); cudaMalloc(&cyc, 8);
  probe<MODE><<<sms, threads>>>(out, in, 100, cyc);
  probe<MODE><<<sms, threads>>>(out, in, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_smsp = (double)iters * 32 * (warps_per_sm / 4.0);
  printf("{\"mode\": %d, \"warps_per_smsp\": %d, \"cycles_per_fp64_inst_per_smsp\": %.3f}\n", MODE, warps_per_sm / 4,
         h / per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double h[24]; for (int i = 0; i < 24; ++i) h[i] = i < 8 ? 0.5 + 0.01 * i : (i < 16 ? 1.0 - 1e-9 * i : 1e-9 * i);
  double* in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {4, 8, 12, 16}) {
    run<0>(w, p.multiProcessorCount, in); run<1>(w, p.multiProcessorCount, in); run<2>(w, p.multiProcessorCount, in);
    run<3>(w, p.multiProcessorCount, in); run<4>(w, p.multiProcessorCount, in); run<5>(w, p.multiProcessorCount, in);
    run<6>(w, p.multiProcessorCount, in); run<7>(w, p.multiProcessorCount, in); run<8>(w, p.multiProcessorCount, in);
  }
  return 0;
}
