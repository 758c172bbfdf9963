// Graph step of the per-step bf16 path on the tensor cores: pairwise kernel + adjacency + masked softmax +
// aggregation of neighbour states as tcgen05 GEMMs per 128-row tile.
//
//   D[128 rows i, 256] = A[i, j] * B[j, 256],  A = un-normalised attention e_ij = adj_ij * exp(kern_ij)
//                                              B = [h | c] of 128 agents (bf16)
//   mh_i = D[i, 0:128] / sum_j e_ij,  mc_i = D[i, 128:256] / sum_j e_ij
//
// 128 % N == 0: a tile holds 128/N whole scenes and A is block diagonal (the zero blocks are written once).
// N % 128 == 0: a scene spans N/128 tiles; the K loop walks the scene's tiles, staging their state and the
// attention columns one 128-agent block at a time and accumulating in TMEM.
// The N x N kernel matrix and the adjacency never exist in HBM.  A is K-major SWIZZLE_128B; B is MN-major
// SWIZZLE_128B ([agent][unit] rows of 128 B: agents along K), which is exactly a row copy of the tile-blocked state
// layout of cell_tc.cu (16-byte pieces of 8 units) -- no transposition while staging.  U == 128.
#include <cuda_bf16.h>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int GM_THREADS = 256;
constexpr int GM_BLK = 128 * 128;                // one [128 rows x 128 B] block
constexpr int GM_A_BYTES = 2 * GM_BLK;           // attention: 2 k-blocks (agents 0-63 | 64-127 of the K block)
constexpr int GM_B_BYTES = 4 * GM_BLK;           // [h units 0-63 | h 64-127 | c 0-63 | c 64-127] x [128 agents x 128 B]
constexpr int GM_MAXN = 1024;                    // largest scene of the multi-tile path
constexpr int GM_SM_A = 0;
constexpr int GM_SM_B = GM_SM_A + GM_A_BYTES;
constexpr int GM_SM_POS = GM_SM_B + GM_B_BYTES;  // float2[max(128, N)]
constexpr int GM_SM_SUM = GM_SM_POS + GM_MAXN * 8;   // float[2][128] partial row sums
constexpr int GM_SM_VAL = GM_SM_SUM + 1024;      // u8[max(128, N)]
constexpr int GM_SM_BAR = GM_SM_VAL + GM_MAXN;   // mbarrier + tmem ptr
constexpr int GM_SM_TOTAL = GM_SM_BAR + 32;
template <bool F16> constexpr uint32_t kIdescAggMN256 = make_idesc_op<F16>(128, 256) | (1u << 16);   // B operand MN-major

__device__ __forceinline__ uint64_t gm_desc_mn(uint32_t smem_addr) {   // LBO = one block, SBO = 8 agents x 128 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(GM_BLK >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// F16: fp16 instead of bf16 operands / state words (MMT_PREC_F16), see cell_tc.cu
template <bool F16>
__global__ void __launch_bounds__(GM_THREADS, 2) graph_aggregate_mma_kernel(
    const float* __restrict__ pos, const uint8_t* __restrict__ valid, const __nv_bfloat16* __restrict__ hb,
    const float* __restrict__ c, const float* __restrict__ score, int R, int N, float r2, float neg_inv_log2e,
    __nv_bfloat16* __restrict__ mhb, __nv_bfloat16* __restrict__ mcb, int num_tiles, uint32_t* trap) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];   // link-time constant base: uniform addresses / descriptors
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, trap, 3);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float2* spos = reinterpret_cast<float2*>(smem + GM_SM_POS);
  float* ssum = reinterpret_cast<float*>(smem + GM_SM_SUM);
  uint8_t* sval = smem + GM_SM_VAL;
  const uint32_t bar = sbase + GM_SM_BAR;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + GM_SM_BAR + 16);

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(sbase + GM_SM_BAR + 16, 256);
  // zero the attention operand once: with 128 % N == 0 the entries outside a row's own scene stay zero
  for (int i = tid; i < GM_A_BYTES / 16; i += GM_THREADS) reinterpret_cast<uint4*>(smem + GM_SM_A)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int i = tid & 127;               // attention row built by this thread
  const int jhalf = tid >> 7;            // which half of the columns
  const bool multi = N > 128;            // a scene spans nk tiles
  const int nk = multi ? N >> 7 : 1;
  const int sb = multi ? 0 : (i / N) * N;   // first row of this row's scene inside the tile
  const int jn = multi ? 64 : (N >> 1);     // columns per thread and K block
  const float LOG2E = 1.4426950408889634f;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tile0 = multi ? (tile / nk) * nk : tile;   // first tile of this tile's scene(s)
    const int li = multi ? (tile - tile0) * 128 + i : i; // this row's index among the staged positions
    // ---- positions / validity of every agent the rows of this tile can see
    for (int a = tid; a < (multi ? N : 128); a += GM_THREADS) {
      const int gr = tile0 * 128 + a;
      spos[a] = gr < R ? __ldg(reinterpret_cast<const float2*>(pos) + gr) : make_float2(0.f, 0.f);
      sval[a] = gr < R ? valid[gr] : 0;
    }
    float sum = 0.f;
    for (int kh = 0; kh < nk; ++kh, ++it) {
      const int kt = tile0 + kh;   // tile whose 128 agents are the K block
      // ---- B operand: piece (g, r) = units 8g..8g+7 of agent r: a 16-byte row copy into the MN-major image
      const uint4* hsrc = reinterpret_cast<const uint4*>(hb + (size_t)kt * 128 * 128);
      const uint4* csrc = reinterpret_cast<const uint4*>(c + (size_t)kt * 128 * 128);
      uint4 hv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hv[k] = __ldg(hsrc + tid + 256 * k);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 cv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cv[k] = __ldg(csrc + tid + 256 * (k + 8 * half));
        if (half == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int p = tid + 256 * k, g = p >> 7, r = p & 127;
            *reinterpret_cast<uint4*>(smem + GM_SM_B + (g >> 3) * GM_BLK + r * 128 + (((g & 7) ^ (r & 7)) << 4)) = hv[k];
          }
        }
        // c is fp32: quarter qd = 4 floats of piece (g, r); lane pairs hold the two quarters of one piece
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int qd = tid + 256 * (k + 8 * half), piece = qd >> 1, hf = qd & 1;
          const int g = piece >> 7, r = piece & 127;
          const uint32_t w0 = pack_op2<F16>(__uint_as_float(cv[k].x), __uint_as_float(cv[k].y));
          const uint32_t w1 = pack_op2<F16>(__uint_as_float(cv[k].z), __uint_as_float(cv[k].w));
          *reinterpret_cast<uint2*>(smem + GM_SM_B + (2 + (g >> 3)) * GM_BLK + r * 128 + (((g & 7) ^ (r & 7)) << 4) + hf * 8) =
              make_uint2(w0, w1);
        }
      }
      __syncthreads();  // spos / sval visible (first K block); the previous MMAs have been waited for below
      // ---- A operand: un-normalised attention of row i against this thread's columns of the K block
      {
        const float2 pi = spos[li];
        const bool vi = sval[li] != 0;
        const int cb = multi ? kh * 128 : sb;          // staged index of the K block's / scene's first agent
        const int jbeg = jhalf * jn;
        // relational variant (g2k_lstm_mcr): logits = kern + edge score; score[scene][i][j] row of this agent, columns of
        // this K block (read only where the adjacency holds: the edge kernel writes the edges only)
        const float* srow = nullptr;
        if (score != nullptr) {
          const size_t gi = (size_t)tile * 128 + i;                  // global agent row; gi / N = scene, gi % N = local index
          srow = score + gi * N + (multi ? kh * 128 : 0);
        }
        for (int j8 = jbeg; j8 < jbeg + jn; j8 += 8) {
          uint32_t pk[4];
          float sc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (srow != nullptr && (size_t)tile * 128 + i < (size_t)R) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(srow + j8));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(srow + j8 + 4));
            sc8[0] = s0.x; sc8[1] = s0.y; sc8[2] = s0.z; sc8[3] = s0.w;
            sc8[4] = s1.x; sc8[5] = s1.y; sc8[6] = s1.z; sc8[7] = s1.w;
          }
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            float e2[2];
#pragma unroll
            for (int z = 0; z < 2; ++z) {
              const int j = cb + j8 + q + z;
              const float2 pj = spos[j];
              const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y);
              const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
              const bool a = vi && sval[j] != 0 && j != li && d2 < r2;
              const float kern = ex2_fast(d2 * neg_inv_log2e);     // exp(-d2 / 2 sigma^2)
              e2[z] = a ? ex2_fast((kern + sc8[q + z]) * LOG2E) : 0.f;   // exp(logit); softmax numerator
            }
            pk[q >> 1] = pack_op2<F16>(e2[0], e2[1]);
            sum += op_lo<F16>(pk[q >> 1]) + op_hi<F16>(pk[q >> 1]);      // normalise by what the MMA really sums
          }
          const int jt = (multi ? 0 : sb) + j8;                    // column inside the K block
          *reinterpret_cast<uint4*>(smem + GM_SM_A + (jt >> 6) * GM_BLK + sw128_off(i, jt & 63)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = make_desc_sw128(sbase + GM_SM_A + (ks >> 2) * GM_BLK) + (uint64_t)((ks & 3) * 2);
          umma_bf16(tmem_base, da, gm_desc_mn(sbase + GM_SM_B + ks * 2048), kIdescAggMN256<F16>, (kh | ks) ? 1u : 0u);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, it & 1u, trap, 0x301);   // operands free for the next K block / accumulator complete
      tc_fence_after();
    }
    ssum[jhalf * 128 + i] = sum;
    __syncthreads();
    // ---- epilogue: thread = row (TMEM lane), warps 0-3 -> mh (cols 0..127), warps 4-7 -> mc (cols 128..255)
    {
      const int q = warp & 3, half = warp >> 2;
      const int r = q * 32 + lane;
      const float s = ssum[r] + ssum[128 + r];
      const float inv = s > 0.f ? 1.0f / s : 0.f;
      __nv_bfloat16* outp = half ? mcb : mhb;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
#pragma unroll
      for (int ch = 0; ch < 16; ch += 2) {
        float v0[8], v1[8];
        tmem_ld8(t_row + ch * 8, v0);
        tmem_ld8(t_row + ch * 8 + 8, v1);
        tmem_wait_ld();
        const size_t o0 = ((size_t)(tile * 16 + ch) * 128 + r) * 8;
        *reinterpret_cast<uint4*>(outp + o0) =
            make_uint4(pack_op2<F16>(v0[0] * inv, v0[1] * inv), pack_op2<F16>(v0[2] * inv, v0[3] * inv),
                       pack_op2<F16>(v0[4] * inv, v0[5] * inv), pack_op2<F16>(v0[6] * inv, v0[7] * inv));
        *reinterpret_cast<uint4*>(outp + o0 + 128 * 8) =
            make_uint4(pack_op2<F16>(v1[0] * inv, v1[1] * inv), pack_op2<F16>(v1[2] * inv, v1[3] * inv),
                       pack_op2<F16>(v1[4] * inv, v1[5] * inv), pack_op2<F16>(v1[6] * inv, v1[7] * inv));
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM drained, row sums consumed and smem operands free before the next tile
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// score: NULL (g2k_lstm_mc) or [S,N,N] edge scores added to the logits on the edges (g2k_lstm_mcr)
int launch_graph_aggregate_mma(const float* pos, const uint8_t* valid, const void* hb, const float* c, const float* score,
                               int S, int N, float r2, float inv_2sigma2, void* mhb, void* mcb, int f16, cudaStream_t stream) {
  const int R = S * N, tiles = (R + 127) / 128;
  static DeviceMask smem_opted[2];   // per kernel: devices already opted in
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&graph_aggregate_mma_kernel<false>), GM_SM_TOTAL + 1024, &smem_opted[0])) return rc;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&graph_aggregate_mma_kernel<true>), GM_SM_TOTAL + 1024, &smem_opted[1])) return rc;
  const int grid = tiles < 2 * num_sms() ? tiles : 2 * num_sms();
  if (f16)
    graph_aggregate_mma_kernel<true><<<grid, GM_THREADS, GM_SM_TOTAL + 1024, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, score, R, N, r2, -inv_2sigma2 * 1.4426950408889634f,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb), tiles, trap_record());
  else
    graph_aggregate_mma_kernel<false><<<grid, GM_THREADS, GM_SM_TOTAL + 1024, stream>>>(
        pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, score, R, N, r2, -inv_2sigma2 * 1.4426950408889634f,
        reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb), tiles, trap_record());
  count_launch();
  return check_launch("graph_aggregate_mma_kernel");
}

}  // namespace mmt
