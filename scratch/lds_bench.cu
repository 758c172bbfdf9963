// Shared-memory wavefront cost of the load patterns considered for the rollout epilogue (one CTA of 512 threads per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += 512) s[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  int off;   // in floats
  if (MODE == 0) off = 0;                       // uniform LDS.128
  else if (MODE == 1) off = (lane & 3) * 2;     // LDS.64, 4 distinct adjacent addresses
  else if (MODE == 2) off = lane * 2;           // LDS.64, 32 distinct consecutive
  else if (MODE == 3) off = (lane & 3) * 4;     // LDS.128, 4 distinct adjacent
  else if (MODE == 4) off = 0;                  // uniform LDS.64
  else if (MODE == 5) off = lane * 4;           // LDS.128, 32 distinct consecutive
  else off = (lane & 3);                        // LDS.32, 4 distinct
  const float* p = s + off;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const unsigned q = (unsigned)__cvta_generic_to_shared(p + ((it * 16 + j) & 63) * 64);
      float a, b, c, d;
      if (MODE == 0 || MODE == 3 || MODE == 5) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(q)); acc += a + d; }
      else if (MODE == 6) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(q)); acc += a; }
      else { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(q)); acc += a + b; }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
  out[blockIdx.x * 512 + threadIdx.x] = acc;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
  const char* names[] = {"uniform LDS.128", "LDS.64 4 distinct", "LDS.64 32 distinct", "LDS.128 4 distinct", "uniform LDS.64", "LDS.128 32 distinct", "LDS.32 4 distinct"};
  for (int m = 0; m < 7; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (m) {
        case 0: k<0><<<148, 512>>>(out, clk); break; case 1: k<1><<<148, 512>>>(out, clk); break;
        case 2: k<2><<<148, 512>>>(out, clk); break; case 3: k<3><<<148, 512>>>(out, clk); break;
        case 4: k<4><<<148, 512>>>(out, clk); break; case 5: k<5><<<148, 512>>>(out, clk); break;
        default: k<6><<<148, 512>>>(out, clk); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-22s %8.2f clk per warp-instruction (16 warps, 4096 loads each)\n", names[m], (double)c / (16.0 * 4096.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
