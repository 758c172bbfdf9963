// TMA-fed tcgen05 GEMM in tf32 for the contractions of the training step (include/mmt.h: mmt_gemm_tf32):
//     C[M,N] (+)= op(A) op(B),  fp32 in memory, operands rounded to tf32 by the tensor core, fp32 accumulation in TMEM.
// The reference has no backward pass at all; the training step defined for it (App. C.5) needs
//     dW  += [e|h|mh]^T dz      M = 320, N = 384, K = rows   (both operands "transposed": MN-major, split over K)
//     dA   = dz W^T             M = rows, N = 320, K = 384   (both operands K-major)
// and their smaller relatives (head, embedding).  Round 1 ran them as library GEMMs.
//
// One CTA = one 128 x 128 tile of C over a K range:
//   warp 0   TMA producer: tensor-map copies (cp.async.bulk.tensor.2d, SWIZZLE_128B) of the A and B k-blocks
//            (32 fp32 = one 128-byte swizzle row per k-block) into a 4-stage ring, completion on mbarriers
//   warp 1   tensor-memory allocation + MMA issue: 4 x tcgen05.mma.kind::tf32 (M128 N128 K8) per stage
//   warps 2-5 epilogue: tcgen05.ld of the accumulator (lane quarter = warp % 4), store or red.add (split-K) to C
// Operand layouts in shared memory are exactly what the tensor map writes:
//   K-major  (A given as [M,K], B given as [N,K]): box {32 k, 128 rows}: row r at r * 128 B, 16-byte chunks XOR (r & 7)
//   MN-major (A given as [K,M], B given as [K,N]): four boxes {32 mn, 32 k}, swizzle 128B with a 32-byte base (the only
//            MN-major layout tf32 has): chunk j of 32 mn-elements at j * 4 KB, k-row at 128 B, four-row k-atoms of 512 B
#include <cuda.h>

#include <mutex>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int GB_M = 128, GB_N = 128, GB_K = 32;         // CTA tile; one k-block = 32 fp32 = 128 B
constexpr int GB_STAGES = 3;                             // 96 KB of stages: TWO CTAs per SM, so that one CTA's prologue / epilogue
                                                         // (latency-bound: these GEMMs have 10-12 k-blocks per tile) hides behind the other's main loop
constexpr int GB_TILE_BYTES = GB_M * GB_K * 4;           // 16 KB per operand per stage
constexpr int GB_SM_A = 0;
constexpr int GB_SM_B = GB_SM_A + GB_STAGES * GB_TILE_BYTES;
constexpr int GB_SM_BAR = GB_SM_B + GB_STAGES * GB_TILE_BYTES;
constexpr int GB_SM_TOTAL = GB_SM_BAR + 128;
constexpr int GB_THREADS = 192;

// instruction descriptor, kind::tf32: D = f32 (bits 4-5 = 1), A = B = tf32 (format 2 at bits 7-9 / 10-12),
// bit 15 / 16: A / B operand MN-major, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major tf32 operand: the only layout the tensor core accepts is SWIZZLE_128B with a 32-byte swizzle base
// (descriptor layout type 1; tensor map CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 128 bytes of MN per k-row, k-atoms of
// FOUR rows (32-byte chunk index XOR (k & 3)), so one K = 8 MMA spans two atoms: SBO = 512 B between them, LBO = 4 KB
// between the 32-element MN chunks.  (With the plain 16-byte-base SWIZZLE_128B the MMA reads zeros.)
__device__ __forceinline__ uint64_t gemm_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's row: 128 bytes = one full line of C per thread and store
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmArgs {
  float* C;
  int M, N, K, ldc;
  int kb_per_split;     // k-blocks (of 32) per grid.z slice
  int a_mn, b_mn;       // operand given "transposed" in memory (MN-major)
  int mode;             // 0: C = acc, 1: C += acc (plain read-modify-write), 2: red.add (split-K or shared C)
  float alpha;
  uint32_t* trap;
};

__global__ void __launch_bounds__(GB_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, a.trap, 6);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t FULL = sbase + GB_SM_BAR, EMPTY = FULL + 8 * GB_STAGES, ACC = EMPTY + 8 * GB_STAGES;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + GB_SM_BAR + 96);
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * GB_N;
  const int kb_total = (a.K + GB_K - 1) / GB_K;
  const int kb0 = blockIdx.z * a.kb_per_split;
  const int kb1 = min(kb_total, kb0 + a.kb_per_split);
  const int nkb = kb1 - kb0;

  if (tid == 0) {
    for (int s = 0; s < GB_STAGES; ++s) {
      mbar_init(FULL + 8 * s, 1);
      mbar_init(EMPTY + 8 * s, 1);
    }
    mbar_init(ACC, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(sbase + GB_SM_BAR + 96, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0 && nkb > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int i = 0; i < nkb; ++i) {
        const uint32_t s = i % GB_STAGES, ph = (i / GB_STAGES) & 1u;
        mbar_wait(EMPTY + 8 * s, ph ^ 1u, a.trap, 0x601);
        mbar_arrive_expect_tx(FULL + 8 * s, 2 * GB_TILE_BYTES);
        const int k0 = (kb0 + i) * GB_K;
        const uint32_t da = sbase + GB_SM_A + s * GB_TILE_BYTES, db = sbase + GB_SM_B + s * GB_TILE_BYTES;
        if (a.a_mn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(da + j * 4096, &tmA, m0 + 32 * j, k0, FULL + 8 * s);
        } else {
          tma_load_2d(da, &tmA, k0, m0, FULL + 8 * s);
        }
        if (a.b_mn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(db + j * 4096, &tmB, n0 + 32 * j, k0, FULL + 8 * s);
        } else {
          tma_load_2d(db, &tmB, k0, n0, FULL + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc_tf32(GB_M, GB_N, a.a_mn != 0, a.b_mn != 0);
      for (int i = 0; i < nkb; ++i) {
        const uint32_t s = i % GB_STAGES, ph = (i / GB_STAGES) & 1u;
        mbar_wait(FULL + 8 * s, ph, a.trap, 0x602);
        tc_fence_after();
        const uint32_t sa = sbase + GB_SM_A + s * GB_TILE_BYTES, sb = sbase + GB_SM_B + s * GB_TILE_BYTES;
        const uint64_t da = a.a_mn ? gemm_desc_mn(sa) : make_desc_sw128(sa);
        const uint64_t db = a.b_mn ? gemm_desc_mn(sb) : make_desc_sw128(sb);
#pragma unroll
        for (int ks = 0; ks < GB_K / 8; ++ks) {
          // one MMA consumes 8 k: 32 bytes inside the swizzle row (K-major) or one 8-row group of 1 KB (MN-major)
          const uint64_t oa = a.a_mn ? (uint64_t)(ks * 64) : (uint64_t)(ks * 2);
          const uint64_t ob = a.b_mn ? (uint64_t)(ks * 64) : (uint64_t)(ks * 2);
          umma_tf32(tmem_base, da + oa, db + ob, idesc, (i | ks) ? 1u : 0u);
        }
        umma_commit(EMPTY + 8 * s);
      }
      umma_commit(ACC);
    }
  } else {
    // ------------------------------- epilogue -------------------------------
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
    if (nkb > 0) {
      mbar_wait(ACC, 0, a.trap, 0x603);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < GB_N; c0 += 32) {
      if (n0 + c0 >= a.N) break;                  // columns beyond N (ragged last tile): nothing to store
      float v[32];
      if (nkb > 0) {
        tmem_ld32(t_row + c0, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (row < a.M) {
        float* cp = a.C + (size_t)row * a.ldc + n0 + c0;
        const bool full = n0 + c0 + 32 <= a.N && ((reinterpret_cast<uintptr_t>(cp) & 15u) == 0);
        if (full && a.mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(cp + j) = make_float4(a.alpha * v[j], a.alpha * v[j + 1], a.alpha * v[j + 2], a.alpha * v[j + 3]);
        } else if (full && a.mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = *reinterpret_cast<const float4*>(cp + j);
            o.x = fmaf(a.alpha, v[j], o.x); o.y = fmaf(a.alpha, v[j + 1], o.y);
            o.z = fmaf(a.alpha, v[j + 2], o.z); o.w = fmaf(a.alpha, v[j + 3], o.w);
            *reinterpret_cast<float4*>(cp + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + c0 + j < a.N) {
              const float x = a.alpha * v[j];
              if (a.mode == 0) cp[j] = x;
              else if (a.mode == 1) cp[j] += x;
              else atomicAdd(cp + j, x);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// 2-D fp32 tensor map over a row-major matrix [rows, cols] with leading dimension ld: inner dimension = cols
int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_inner, int box_outer, bool mn_major) {
  EncodeFn fn = encode_fn();
  if (!fn) {
    set_error("mmt_gemm_tf32: cuTensorMapEncodeTiled is not available from this driver");
    return MMT_ECUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mmt_gemm_tf32: cuTensorMapEncodeTiled failed (%d) for a [%d x %d] matrix, ld %d", (int)r, rows, cols, ld);
    return MMT_ECUDA;
  }
  return MMT_OK;
}
}  // namespace

}  // namespace mmt

extern "C" int mmt_gemm_tf32(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                             int M, int N, int K, float alpha, int accumulate, void* stream_) {
  using namespace mmt;
  MMT_REQUIRE(A && B && C, "A/B/C must not be NULL");
  MMT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "M, N, K must be >= 0");
  MMT_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && lda > 0 && ldb > 0 && ldc >= N,
              "leading dimensions: lda, ldb multiples of 4 floats (tensor-map strides are multiples of 16 bytes), ldc >= N");
  MMT_ALIGNED(A);
  MMT_ALIGNED(B);
  if (M == 0 || N == 0) return MMT_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (K == 0) {
    if (!accumulate) {
      const cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, stream);
      if (e != cudaSuccess) { set_error("mmt_gemm_tf32: memset: %s", cudaGetErrorString(e)); return MMT_ECUDA; }
    }
    return MMT_OK;
  }
  // A: op(A) is [M,K].  transA = 0: memory [M,K] (K contiguous: K-major, box {32 k, 128 m});
  //                     transA = 1: memory [K,M] (M contiguous: MN-major, boxes {32 m, 32 k})
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = transA ? make_map(&tmA, A, K, M, lda, 32, 32, true) : make_map(&tmA, A, M, K, lda, GB_K, GB_M, false))) return rc;
  // B: op(B) is [K,N].  transB = 0: memory [K,N] (N contiguous: MN-major); transB = 1: memory [N,K] (K-major)
  if ((rc = transB ? make_map(&tmB, B, N, K, ldb, GB_K, GB_N, false) : make_map(&tmB, B, K, N, ldb, 32, 32, true))) return rc;
  const int tm = (M + GB_M - 1) / GB_M, tn = (N + GB_N - 1) / GB_N, kb = (K + GB_K - 1) / GB_K;
  // split K so that a small output (the weight gradients: 3 x 3 tiles over hundreds of thousands of rows) still fills the GPU
  int splits = 1;
  const int sms = num_sms();
  if (tm * tn < sms) {
    splits = (2 * sms) / (tm * tn);
    if (splits > kb / 8) splits = kb / 8;       // at least 8 k-blocks (256 k) per CTA
    if (splits < 1) splits = 1;
  }
  GemmArgs a;
  a.C = C; a.M = M; a.N = N; a.K = K; a.ldc = ldc;
  a.kb_per_split = (kb + splits - 1) / splits;
  splits = (kb + a.kb_per_split - 1) / a.kb_per_split;
  a.a_mn = transA ? 1 : 0; a.b_mn = transB ? 0 : 1;
  a.alpha = alpha;
  a.mode = splits > 1 ? 2 : (accumulate ? 1 : 0);
  a.trap = trap_record();
  if (splits > 1 && !accumulate) {
    const cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, stream);
    if (e != cudaSuccess) { set_error("mmt_gemm_tf32: memset: %s", cudaGetErrorString(e)); return MMT_ECUDA; }
  }
  static DeviceMask smem_opted[1];
  if ((rc = opt_in_smem(reinterpret_cast<const void*>(&gemm_tf32_kernel), GB_SM_TOTAL + 1024, &smem_opted[0]))) return rc;
  gemm_tf32_kernel<<<dim3(tm, tn, splits), GB_THREADS, GB_SM_TOTAL + 1024, stream>>>(tmA, tmB, a);
  count_launch();
  return check_launch("gemm_tf32_kernel");
}
