// micro-benchmark 3: does a second MMA-issuing thread (another warp) raise the tcgen05.mma issue rate for small N?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;
template <int N, int MODE>
__global__ void __launch_bounds__(384, 1) k(int passes, int nissuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = sbase + 160 * 1024, tslot = bar + 64;
  for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { for (int s = 0; s < 8; ++s) mbar_init(bar + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) tmem_alloc(tslot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 160 * 1024 + 64);
  const int me = warp - 9;
  if (me >= 0 && me < nissuers && (tid & 31) == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t d = tmem + me * 96;
    const long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
#pragma unroll
      for (int kc = 0; kc < 5; ++kc) {
        const uint64_t da = make_desc_sw128(sbase + kc * 16384);
        const uint64_t db = make_desc_sw128(sbase + 80 * 1024 + kc * 12288);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (MODE == 0) umma_bf16(d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
          else umma_bf16_ts(d, tmem + 320 + kc * 32 + ks * 8, db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
        }
        umma_commit(bar + 8 * (4 + me));
      }
    }
    umma_commit(bar + 8 * me);
    mbar_wait(bar + 8 * me, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) out[me] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}
template <int N, int MODE> void run(long long* d) {
  auto kern = k<N, MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024);
  const int passes = 20;
  for (int ni : {1, 2, 3}) {
    kern<<<148, 384, 162 * 1024>>>(passes, ni, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < ni; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s N=%3d issuers=%d: %.1f clk per MMA overall (nominal %.0f)\n", MODE ? "TS" : "SS", N, ni, (double)mx / (passes * 20 * ni), 128.0 * N / 256);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 1024);
  run<96, 1>(d); run<96, 0>(d); run<48, 1>(d); run<128, 1>(d);
  return 0;
}
