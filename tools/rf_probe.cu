// Probe: does the register-file read bandwidth (not the FP64 pipe) bound FP64 code mixed with integer instructions?
// Each FP64 op below has TWO 64-bit register sources (a_i, b_i) like most FP64 ops of the rhs! kernels; NI integer ops
// with RI register sources each are interleaved per FP64 op. Output: cycles per FP64 warp-instruction per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int RI>
__global__ void probe(double* out, int* iout, const double* in, int iters, long long* cyc) {
  double a[8], b[8];
  unsigned x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = in[i] + threadIdx.x * 1e-3; b[i] = in[8 + i];
    x[i] = threadIdx.x * 3 + i; y[i] = threadIdx.x * 5 + 7 * i + 1; z[i] = threadIdx.x ^ (11 * i);
  }
  const double C = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = fma(a[i], b[i], C);
#pragma unroll
        for (int n = 0; n < NI; ++n) {
          const int j = (i + 3 * n + 1) & 7;
          if (RI == 1) x[j] = (x[j] << 1) ^ 0x9e3779b9u;                 // one register source (+ immediates)
          if (RI == 2) x[j] = x[j] ^ y[j];                               // LOP3 with two register sources
          if (RI == 3) x[j] = (x[j] & y[j]) ^ z[j];                      // LOP3 with three register sources
        }
      }
  }
  long long t1 = clock64();
  double s = 0; unsigned u = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; u += x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = u;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI, int RI> void run(int warps_per_sm, int sms, const double* in) {
  double* out; int* iout; long long* cyc;
  int threads = 32 * warps_per_sm, iters = 4000;
  cudaMalloc(&out, 8 * sms * threads); cudaMalloc(&iout, 4 * sms * threads); cudaMalloc(&cyc, 8);
  probe<NI, RI><<<sms, threads>>>(out, iout, in, 100, cyc);
  probe<NI, RI><<<sms, threads>>>(out, iout, in, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"int_ops_per_fp64\": %d, \"int_reg_sources\": %d, \"warps_per_smsp\": %d, \"cycles_per_fp64_per_smsp\": %.3f}\n", NI, RI,
         warps_per_sm / 4, h / ((double)iters * 32 * (warps_per_sm / 4.0)));
  cudaFree(out); cudaFree(iout); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double h[24]; for (int i = 0; i < 24; ++i) h[i] = i < 8 ? 0.5 + 0.01 * i : 1.0 - 1e-9 * i;
  double* in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {4, 12}) {
    run<0, 1>(w, p.multiProcessorCount, in);
    run<1, 1>(w, p.multiProcessorCount, in); run<1, 2>(w, p.multiProcessorCount, in); run<1, 3>(w, p.multiProcessorCount, in);
    run<2, 1>(w, p.multiProcessorCount, in); run<2, 2>(w, p.multiProcessorCount, in); run<2, 3>(w, p.multiProcessorCount, in);
  }
  return 0;
}
