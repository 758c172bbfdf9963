// micro-benchmark: throughput of MUFU ops on sm_100a (lanes/clk/SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 3) { float t; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a[i] * -2.885390f)); asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + t)); y = fmaf(y, 2.0f, -1.0f); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(a[i]));
      a[i] = y * 0.999f + 0.001f;
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name) {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 2000;
  k<OP><<<148 * 8, 256>>>(d, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * iters * 8;
  printf("%-10s %.3f ms  %.1f Gop/s  -> %.2f lanes/clk/SM @1.9GHz\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(d);
}
int main() { run<0>("tanh"); run<1>("ex2"); run<2>("rcp"); run<3>("ex2+rcp"); run<4>("fma(+fma)"); return 0; }
