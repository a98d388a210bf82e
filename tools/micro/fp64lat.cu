// DFMA dependent latency and throughput vs (warps per SM, ILP): build/micro/fp64lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
void run(int warps) {
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<ILP><<<148, 32 * warps>>>(d, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<148, 32 * warps>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk * 1e3;
  double per_warp_instr = (double)iters * 8 * ILP;
  printf("warps/SM %2d ILP %d: %.2f cycles per dependent step, %.1f DFMA/clk/SM\n", warps, ILP, cycles / (iters * 8.0),
         per_warp_instr * warps * 32 / cycles);
  cudaFree(d);
}
int main() {
  for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
