// Probe: dependent-issue latency of FP64 instructions on B200 (one warp per SM sub-partition, NCH independent chains).
// cycles per instruction * NCH = latency when the chains are too few to fill the pipe.
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH, int OP>
__global__ void probe(double* out, const double* in, int iters, long long* cyc) {
  double a[NCH];
  const double b = in[8], c = in[16];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = in[i] + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (OP == 0) a[i] = fma(a[i], b, c);
        if (OP == 1) a[i] = a[i] + b;
        if (OP == 2) a[i] = a[i] * b;
        if (OP == 3) { double r_; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r_) : "d"(a[i])); a[i] = r_; }
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NCH, int OP> void run(const double* in, int warps) {
  double* out; long long* cyc;
  int iters = 2000;
  cudaMalloc(&out, 8 * 32 * warps); cudaMalloc(&cyc, 8);
  probe<NCH, OP><<<1, 32 * warps>>>(out, in, 10, cyc);
  probe<NCH, OP><<<1, 32 * warps>>>(out, in, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"dfma", "dadd", "dmul", "mufu_rcp64h"};
  printf("{\"op\": \"%s\", \"chains\": %d, \"warps_per_smsp\": %d, \"cycles_per_chain_step\": %.2f}\n", names[OP], NCH,
         warps / 4, (double)h / (iters * 16.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  double h[24]; for (int i = 0; i < 24; ++i) h[i] = i < 8 ? 0.5 + 0.01 * i : (i < 16 ? 1.0 - 1e-9 * i : 1e-9 * i);
  double* in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<1, 0>(in, 4); run<2, 0>(in, 4); run<4, 0>(in, 4); run<8, 0>(in, 4);
  run<1, 1>(in, 4); run<1, 2>(in, 4); run<1, 3>(in, 4); run<4, 3>(in, 4);
  run<1, 0>(in, 8); run<1, 0>(in, 12); run<4, 0>(in, 12);
  return 0;
}
