// follow-up of pair_bench.cu (results: profiles/r01e_pair_bench2.md; written at the end of round 1, it
// ran first time on a B200).  pair_bench.cu showed that a CTA pair
// (cta_group::2, M256 N96 K16) issues at the nominal 48 clk per MMA but that forwarding "my half of the stage has
// landed" from the peer to the leader through a remote mbarrier.arrive costs 4x the MMA time.  Here both CTAs load
// their 48 of the 96 weight rows of a chunk with a tensor-map copy carrying .cta_group::2, whose complete_tx is
// delivered to the LEADER's mbarrier (shared::cluster address of rank 0): the leader's issuer waits on ONE local
// barrier per stage that expects the bytes of both halves, and nothing is forwarded in software.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pair_bench2 pair_bench2.cu && ./pair_bench2 [stream 0|1|2]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../multimodaltraj_2_b200/csrc/tc_common.cuh"
using namespace mmt;

constexpr int STAGE = 12288;   // a chunk = 96 weight rows x 64 bf16 (128 B, K-major); each CTA of the pair holds 48 rows = 6 KB
constexpr int ROWS_PER_CTA = 48, CHUNKS = 20, MAXSTAGE = 8;
constexpr int SM_BAR = 49152;                                 // W_FULL[8] W_EMPTY[8] DONE, tmem slot
constexpr int SM_TOTAL = SM_BAR + 256;

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// 48 rows x 128 B of the weight stream into this CTA's stage; the transaction bytes are credited to `leader_bar`
__device__ __forceinline__ void tma_2sm_load(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(x), "r"(y), "r"(leader_bar)
      : "memory");
}

// NSTAGE stages of STRIDE bytes per CTA: <4, 12288> is the ring of rollout_tc.cu with half of every stage unused,
// <8, 6144> the ring the pair affords in the same 48 KB
template <int NSTAGE, int STRIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __grid_constant__ CUtensorMap tm, int passes, int stream, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t W_FULL = sbase + SM_BAR, W_EMPTY = W_FULL + 8 * MAXSTAGE, DONE = W_EMPTY + 8 * MAXSTAGE, tslot = DONE + 16;
  const uint32_t rank = cluster_rank();
  for (int i = tid; i < SM_BAR / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 2 * MAXSTAGE + 1; ++s) mbar_init(W_FULL + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * (2 * MAXSTAGE) + 16);   // tslot
  const uint32_t total = (uint32_t)passes * 5;   // chunks (4 MMAs each)

  if (warp == 1 && lane == 0 && stream) {
    // one producer lane per CTA: chunk `it` -> stage it % 4 of BOTH CTAs (each its own 48 rows), refilled as soon as
    // the pair MMAs that read the stage have completed (the commit is multicast to both CTAs' W_EMPTY)
    uint32_t leader_full;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_full) : "r"(W_FULL), "r"(0));
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
      if (stream == 2) mbar_wait_spin(W_EMPTY + 8 * s, ph ^ 1u); else mbar_wait(W_EMPTY + 8 * s, ph ^ 1u);
      if (stream == 3) {
        // control experiment (timing only, the leader does not know when the peer's half has landed): every CTA
        // credits its OWN barrier, no transaction bytes cross the cluster
        mbar_arrive_expect_tx(W_FULL + 8 * s, STAGE / 2);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                sbase + s * STRIDE),
            "l"(&tm), "r"(0), "r"((int)((it % CHUNKS) * 96 + rank * ROWS_PER_CTA)), "r"(W_FULL + 8 * s)
            : "memory");
        continue;
      }
      if (rank == 0) mbar_arrive_expect_tx(W_FULL + 8 * s, STAGE);   // both halves are credited here
      tma_2sm_load(sbase + s * STRIDE, &tm, 0, (int)((it % CHUNKS) * 96 + rank * ROWS_PER_CTA), leader_full + 8 * s);
    }
  } else if (warp == 0 && rank == 0) {
    // issuer: the whole warp runs convergently, one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(256, 96);
    const long long t0 = clock64();
    uint32_t it = 0;
    for (int p = 0; p < passes; ++p) {
      const uint32_t d = tmem + 160 + (p & 1) * 96;
#pragma unroll
      for (int kc = 0; kc < 5; ++kc, ++it) {
        const uint32_t s = it % NSTAGE;
        if (stream) {
          if (stream == 2) mbar_wait_spin(W_FULL + 8 * s, (it / NSTAGE) & 1u); else mbar_wait(W_FULL + 8 * s, (it / NSTAGE) & 1u);
          __syncwarp();
        }
        const uint64_t db = make_desc_sw128(sbase + s * STRIDE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (kc | ks) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
              "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d),
              "r"(tmem + kc * 32 + ks * 8), "l"(db + (uint64_t)(ks * 2)), "r"(idesc), "r"(acc), "r"(0)
              : "memory");
        }
        asm volatile(
            "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|e, 0xffffffff;\n\t"
            "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
                W_EMPTY + 8 * s)
            : "memory");
      }
    }
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(DONE)
        : "memory");
    mbar_wait(DONE, 0);
    __syncwarp();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer's stages and barriers must outlive the leader's last MMA / commit
  if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NSTAGE, int STRIDE> void run(const CUtensorMap& tm, int stream, long long* d) {
  auto kern = k<NSTAGE, STRIDE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  const int passes = 200;
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<148, 128, SM_TOTAL>>>(tm, passes, stream, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s (stream %d)\n", cudaGetErrorString(e), stream); exit(1); }
  }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("CTA pair cta_group::2 M256 N96 K16, %d stages of %d B per SM, tensor-map 2SM weight stream %s: %.1f clk per MMA, %.0f clk per pass of 20 (nominal 48 / 960)\n",
         NSTAGE, STRIDE, stream == 3 ? "on, every CTA credits its own barrier (control)" : stream == 2 ? "on, spinning waits" : stream ? "on " : "off", (double)h / (passes * 20), (double)h / passes);
  fflush(stdout);
}

int main(int argc, char** argv) {   // pair_bench2 <stream 0|1|2>: one configuration per process
  const int stream = argc > 1 ? atoi(argv[1]) : 1;
  uint8_t* w; cudaMalloc(&w, (size_t)CHUNKS * STAGE); cudaMemset(w, 0, (size_t)CHUNKS * STAGE);
  long long* d; cudaMalloc(&d, 64);
  // the weight stream as a 2-D bf16 tensor [CHUNKS * 96 rows][64]; box = 48 rows x 64 (6 KB); the image in global
  // memory is already in the SWIZZLE_128B order the MMA descriptor expects (as in rollout_tc.cu), so no TMA swizzle
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tm;
  const cuuint64_t dims[2] = {64, (cuuint64_t)CHUNKS * 96}, strides[1] = {128};
  const cuuint32_t box[2] = {64, ROWS_PER_CTA}, estr[2] = {1, 1};
  CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  run<4, 12288>(tm, stream, d);
  if (stream == 1) { run<8, 6144>(tm, stream, d); run<6, 6144>(tm, stream, d); }
  return 0;
}
