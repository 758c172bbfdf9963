// micro-benchmark: packed half-precision MUFU / FMA throughput per SM (sm_100a): does tanh.approx.f16x2 / .bf16x2 produce two
// results per MUFU issue slot, and what do HFMA2 (f16x2 / bf16x2) sustain next to FFMA2 (f32x2)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o build/mufu_bench3 scratch/mufu_bench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(uint32_t* out, int iters, long long* clk) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x38003800u + threadIdx.x + i * 17;   // two small halves per word
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 2) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(a[i]));
      if (OP == 3) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(a[i]));
      if (OP == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 5) { float y = __uint_as_float(a[i]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(y)); a[i] = __float_as_uint(y); }
      if (OP == 6) {   // f32 pair -> f16x2 (one cvt) -> tanh.f16x2 -> two f32 (as a mixed-precision epilogue would use it)
        float lo = __uint_as_float(a[i]), hi = lo + 0.25f; uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h));
        uint16_t l16 = h & 0xffff, h16 = h >> 16; float fl, fh;
        asm volatile("cvt.f32.f16 %0, %1;" : "=f"(fl) : "h"(l16));
        asm volatile("cvt.f32.f16 %0, %1;" : "=f"(fh) : "h"(h16));
        a[i] = __float_as_uint(fl + fh);
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP> void run(const char* name, int threads, int elems) {
  uint32_t* d; cudaMalloc(&d, 148 * 1024 * 4); long long* c; cudaMalloc(&c, 8);
  const int iters = 4000;
  k<OP><<<148, threads>>>(d, 10, c);
  k<OP><<<148, threads>>>(d, iters, c); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads/SM %4d: %.2f instr-lanes/clk/SM = %.2f elements/clk/SM (%lld clk)\n", name, threads,
         (double)threads * iters * 8 / h, (double)threads * iters * 8 * elems / h, h);
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int th : {512, 1024}) {
    run<5>("tanh.approx.f32", th, 1); run<0>("tanh.approx.f16x2", th, 2); run<1>("tanh.approx.bf16x2", th, 2);
    run<4>("ex2.approx.f16x2", th, 2); run<2>("fma.rn.f16x2", th, 2); run<3>("fma.rn.bf16x2", th, 2);
    run<6>("cvt+tanh.f16x2+2cvt", th, 2);
  }
  return 0;
}
