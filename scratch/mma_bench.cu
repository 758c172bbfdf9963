// micro-benchmark: tcgen05.mma issue/execution rate on sm_100a, one issuing thread per SM.
//   SS (A and B in smem, SWIZZLE_128B K-major) vs TS (A in TMEM), N in {32..256}, with/without a
//   tcgen05.commit per chunk of 4 MMAs, with/without 8 worker warps using shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int MODE, int COMMIT>
__global__ void __launch_bounds__(320, 1) k(int passes, int hammer, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = sbase + 160 * 1024, bar2 = bar + 8, tslot = bar + 64;
  for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); for (int s = 0; s < 5; ++s) mbar_init(bar2 + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) tmem_alloc(tslot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 160 * 1024 + 64);
  __shared__ volatile int done;
  if (tid == 0) done = 0;
  __syncthreads();
  if (warp == 9) {
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N);
      uint32_t ph = 0;
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
        for (int p = 0; p < passes; ++p) {
#pragma unroll
          for (int kc = 0; kc < 5; ++kc) {
            const uint64_t da = make_desc_sw128(sbase + kc * 16384);
            const uint64_t db = make_desc_sw128(sbase + 80 * 1024 + kc * 12288);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (MODE == 0) umma_bf16(tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
              else umma_ts(tmem, tmem + 256 + kc * 32 + ks * 8, db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
            }
            if (COMMIT) umma_commit(bar2 + 8 * kc);
          }
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, ph);
        ph ^= 1;
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
      }
      done = 1;
    }
  } else if (warp < 8 && hammer >= 3) {
    // 3: tcgen05.ld loop; 4: MUFU.TANH loop; 5: FFMA2 loop; 6: ld + tanh + ffma2 mix (like the gate epilogue)
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    float acc[8] = {0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f};
    while (!done) {
      if (hammer == 3 || hammer == 6) {
        float v[8];
        tmem_ld8(t_row + (warp >> 2) * 64, v);
        tmem_ld8(t_row + (warp >> 2) * 64 + 8, acc);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
      if (hammer == 4 || hammer == 6) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = tanh_fast(acc[i]);
      }
      if (hammer == 5 || hammer == 6) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float2 t = ffma2(make_float2(acc[i], acc[i + 1]), make_float2(0.99f, 0.98f), make_float2(0.01f, 0.02f));
            acc[i] = t.x; acc[i + 1] = t.y;
          }
      }
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6] + acc[7] == 0.12345f) out[101] = 1;
  } else if (warp < 8 && hammer) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint4* p = reinterpret_cast<uint4*>(smem + 144 * 1024);
    int it = 0;
    while (!done) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 v = p[(tid + 256 * ((it + j) & 3))];
        acc.x ^= v.x; acc.y += v.y;
        if (hammer > 1) p[(tid + 256 * ((it + j + 1) & 3))] = acc;
      }
      ++it;
    }
    if (acc.x == 0x12345) out[100] = acc.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

template <int N, int MODE, int COMMIT> void run(long long* d) {
  auto kern = k<N, MODE, COMMIT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024);
  const int passes = 20;
  for (int hammer : {0, 2, 3, 4, 5, 6}) {
    kern<<<148, 320, 162 * 1024>>>(passes, hammer, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("%s N=%3d commit/chunk=%d smem-traffic=%d: issue %.1f clk/MMA, complete %.1f clk/MMA (nominal %.0f)\n", MODE ? "TS" : "SS", N, COMMIT, hammer,
           (double)h[4] / (passes * 20), (double)h[5] / (passes * 20), 128.0 * N / 256);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 1024);
  run<96, 0, 1>(d); run<96, 1, 1>(d); run<128, 1, 1>(d); run<192, 1, 1>(d);
  return 0;
}
