// micro-benchmark: MUFU / FFMA2 throughput per SM at a given number of resident warps (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(float* out, int iters, long long* clk) {
  float a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      float y;
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(a[i]));
      if (OP == 4) { // tanh + 3 fma (mixed, as in the gate epilogue)
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
        y = fmaf(y, 0.5f, 0.5f); y = fmaf(y, a[i], 0.25f); y = fmaf(y, 0.3f, a[i]);
      }
      a[i] = y;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP>
__global__ void k2(float* out, int iters, long long* clk) {   // packed f32x2 fma, ILP 8
  uint64_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = (uint64_t)__float_as_uint(threadIdx.x * 1e-3f + i) * 0x100000001ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(a[i]));
  }
  long long t1 = clock64();
  uint64_t s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP, int ILP> void run(const char* name, int threads) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4); long long* c; cudaMalloc(&c, 8);
  const int iters = 4000;
  k<OP, ILP><<<148, threads>>>(d, 10, c);
  k<OP, ILP><<<148, threads>>>(d, iters, c); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double ops = (double)threads * iters * ILP * (OP == 4 ? 1 : 1);
  printf("%-12s threads/SM %4d ILP %d: %.2f lanes/clk/SM (%lld clk)\n", name, threads, ILP, ops / h, h);
  cudaFree(d); cudaFree(c);
}
void run2(int threads) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4); long long* c; cudaMalloc(&c, 8);
  const int iters = 4000;
  k2<0><<<148, threads>>>(d, 10, c);
  k2<0><<<148, threads>>>(d, iters, c); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-12s threads/SM %4d ILP 8: %.2f ffma2-lanes/clk/SM (x2 flop-lanes)\n", "ffma2", threads, (double)threads * iters * 8 / h);
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int th : {256, 512, 1024}) {
    run<0, 8>("tanh", th); run<1, 8>("ex2", th); run<2, 8>("rcp", th); run<3, 8>("ffma", th); run<4, 8>("tanh+3fma", th);
    run2(th);
  }
  run<0, 2>("tanh", 256); run<0, 4>("tanh", 256);
  return 0;
}
