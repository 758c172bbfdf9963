// Graph step of the bf16 rollout on the tensor cores: pairwise kernel + adjacency + masked softmax +
// aggregation of neighbour states as ONE tcgen05 GEMM per 128-row tile.
//
//   D[128 rows i, 256] = A[i, j] * B[j, 256],  A = un-normalised attention e_ij = adj_ij * exp(kern_ij)
//                                              B = [h | c] of the tile's 128 agents (bf16)
//   mh_i = D[i, 0:128] / sum_j e_ij,  mc_i = D[i, 128:256] / sum_j e_ij
//
// A tile holds 128/N whole scenes, so A is block diagonal: each thread builds the entries of one row
// against its own scene only (the zero blocks are written once).  The N x N kernel matrix and the
// adjacency therefore never exist in HBM, and the sparse gather of the CUDA-core version
// (~300 warp instructions per row) becomes 8 MMAs (M128 x N256 x K16) per tile.
// Both operands are K-major SWIZZLE_128B in shared memory; B needs agents along K, so the tile's
// state is transposed while it is staged (lane pairs exchange halves with one shuffle so the
// transposed elements are written as packed 32-bit words).  State layout in HBM is the tile-blocked
// one of cell_tc.cu: loads and the 16-byte epilogue stores are fully coalesced.
// Requires 128 % N == 0, U == 128.
#include <cuda_bf16.h>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int GM_THREADS = 256;
constexpr int GM_A_BYTES = 2 * 128 * 128;        // 2 k-blocks x [128 rows x 128 B]
constexpr int GM_B_BLOCK = 256 * 128;            // one k-block of B: [256 rows (h units | c units) x 128 B]
constexpr int GM_B_BYTES = 2 * GM_B_BLOCK;
constexpr int GM_SM_A = 0;
constexpr int GM_SM_B = GM_SM_A + GM_A_BYTES;
constexpr int GM_SM_POS = GM_SM_B + GM_B_BYTES;  // float2[128]
constexpr int GM_SM_SUM = GM_SM_POS + 1024;      // float[2][128] partial row sums
constexpr int GM_SM_VAL = GM_SM_SUM + 1024;      // u8[128]
constexpr int GM_SM_BAR = GM_SM_VAL + 128;       // mbarrier + tmem ptr
constexpr int GM_SM_TOTAL = GM_SM_BAR + 32;
constexpr uint32_t kIdescAgg = make_idesc_bf16(128, 256);

__global__ void __launch_bounds__(GM_THREADS, 2) graph_aggregate_mma_kernel(
    const float* __restrict__ pos, const uint8_t* __restrict__ valid, const __nv_bfloat16* __restrict__ hb,
    const float* __restrict__ c, int R, int N, float r2, float neg_inv_log2e, __nv_bfloat16* __restrict__ mhb,
    __nv_bfloat16* __restrict__ mcb, int num_tiles) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
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
  // zero the attention operand once: entries outside a row's own scene stay zero for every tile
  for (int i = tid; i < GM_A_BYTES / 16; i += GM_THREADS) reinterpret_cast<uint4*>(smem + GM_SM_A)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int i = tid & 127;               // attention row built by this thread
  const int jhalf = tid >> 7;            // which half of the scene's columns
  const int sb = (i / N) * N;            // first row of this row's scene inside the tile
  const int jn = N >> 1;                 // columns per thread (N >= 16 -> multiple of 8; N = 4, 8 handled below)
  const float LOG2E = 1.4426950408889634f;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int row0 = tile * 128;
    // ---- issue the state loads first (longest latency), 24 independent 16-byte loads per thread
    const uint4* hsrc = reinterpret_cast<const uint4*>(hb + (size_t)tile * 128 * 128);
    const uint4* csrc = reinterpret_cast<const uint4*>(c + (size_t)tile * 128 * 128);
    uint4 hv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = __ldg(hsrc + tid + 256 * k);
    uint4 cv0[8], cv1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cv0[k] = __ldg(csrc + tid + 256 * k);
    if (tid < 128) {
      const int gr = row0 + tid;
      spos[tid] = gr < R ? __ldg(reinterpret_cast<const float2*>(pos) + gr) : make_float2(0.f, 0.f);
      sval[tid] = gr < R ? valid[gr] : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cv1[k] = __ldg(csrc + tid + 256 * (k + 8));
    // ---- B operand, h part: piece (g, r) = units 8g..8g+7 of agent r -> B[kb = r/64][row 8g+u][k = r%64]
    //      lane pairs (r even/odd) swap halves so each lane writes 4 packed words (agents r, r+1 of one unit)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int p = tid + 256 * k, g = p >> 7, r = p & 127;
      const bool odd = r & 1;
      const uint32_t sx = odd ? hv[k].x : hv[k].z, sy = odd ? hv[k].y : hv[k].w;   // what the partner needs
      const uint32_t px = __shfl_xor_sync(0xffffffffu, sx, 1), py = __shfl_xor_sync(0xffffffffu, sy, 1);
      // even lane: units 0..3 = (mine x,y ; partner's x,y);  odd lane: units 4..7 = (partner's z,w ; mine z,w)
      const uint32_t e0 = odd ? px : hv[k].x, e1 = odd ? py : hv[k].y;   // even agent's two words
      const uint32_t o0 = odd ? hv[k].z : px, o1 = odd ? hv[k].w : py;   // odd agent's two words
      const int ubase = g * 8 + (odd ? 4 : 0);
      const int kk = (r & 63) & ~1, kb = r >> 6;
      uint8_t* dst = smem + GM_SM_B + kb * GM_B_BLOCK;
      *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 0, kk)) = __byte_perm(e0, o0, 0x5410);
      *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 1, kk)) = __byte_perm(e0, o0, 0x7632);
      *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 2, kk)) = __byte_perm(e1, o1, 0x5410);
      *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 3, kk)) = __byte_perm(e1, o1, 0x7632);
    }
    // ---- B operand, c part (fp32 in HBM -> bf16): qd = 16-byte quarter (4 floats) of piece (g, r)
    auto put_c = [&](const uint4 (&cv)[8], int half) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int qd = tid + 256 * (k + 8 * half), piece = qd >> 1, hf = qd & 1;
        const int g = piece >> 7, r = piece & 127;
        // this lane: 4 units (8g+4hf..+3) of agent r; partner lane^2 holds the same units of agent r^1
        const uint32_t w01 = pack_bf16x2(__uint_as_float(cv[k].x), __uint_as_float(cv[k].y));
        const uint32_t w23 = pack_bf16x2(__uint_as_float(cv[k].z), __uint_as_float(cv[k].w));
        const bool odd = r & 1;
        const uint32_t snd = odd ? w01 : w23;
        const uint32_t rcv = __shfl_xor_sync(0xffffffffu, snd, 2);
        const uint32_t ev = odd ? rcv : w01, ov = odd ? w23 : rcv;   // even / odd agent's word for my 2 units
        const int ubase = 128 + g * 8 + hf * 4 + (odd ? 2 : 0);
        const int kk = (r & 63) & ~1, kb = r >> 6;
        uint8_t* dst = smem + GM_SM_B + kb * GM_B_BLOCK;
        *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 0, kk)) = __byte_perm(ev, ov, 0x5410);
        *reinterpret_cast<uint32_t*>(dst + sw128_off(ubase + 1, kk)) = __byte_perm(ev, ov, 0x7632);
      }
    };
    put_c(cv0, 0);
    put_c(cv1, 1);
    __syncthreads();  // spos / sval visible
    // ---- A operand: un-normalised attention of row i against columns [j0, j0 + jn) of its scene
    {
      const float2 pi = spos[i];
      const bool vi = sval[i] != 0;
      float sum = 0.f;
      const int jbeg = jhalf * jn;
      for (int j8 = jbeg; j8 < jbeg + jn; j8 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          float e2[2];
#pragma unroll
          for (int z = 0; z < 2; ++z) {
            const int j = j8 + q + z;
            const float2 pj = spos[sb + j];
            const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const bool a = vi && sval[sb + j] != 0 && (sb + j) != i && d2 < r2;
            const float kern = ex2_fast(d2 * neg_inv_log2e);     // exp(-d2 / 2 sigma^2)
            e2[z] = a ? ex2_fast(kern * LOG2E) : 0.f;            // exp(kern); softmax numerator
          }
          pk[q >> 1] = pack_bf16x2(e2[0], e2[1]);
          sum += bf16_lo(pk[q >> 1]) + bf16_hi(pk[q >> 1]);      // normalise by what the MMA really sums
        }
        const int jt = sb + j8;                                  // tile-local column
        *reinterpret_cast<uint4*>(smem + GM_SM_A + (jt >> 6) * (128 * 128) + sw128_off(i, jt & 63)) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      ssum[jhalf * 128 + i] = sum;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = make_desc_sw128(sbase + GM_SM_A + kb * (128 * 128));
        const uint64_t db = make_desc_sw128(sbase + GM_SM_B + kb * GM_B_BLOCK);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem_base, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kIdescAgg, (kb | ks) ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, it & 1u);
    tc_fence_after();
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
            make_uint4(pack_bf16x2(v0[0] * inv, v0[1] * inv), pack_bf16x2(v0[2] * inv, v0[3] * inv),
                       pack_bf16x2(v0[4] * inv, v0[5] * inv), pack_bf16x2(v0[6] * inv, v0[7] * inv));
        *reinterpret_cast<uint4*>(outp + o0 + 128 * 8) =
            make_uint4(pack_bf16x2(v1[0] * inv, v1[1] * inv), pack_bf16x2(v1[2] * inv, v1[3] * inv),
                       pack_bf16x2(v1[4] * inv, v1[5] * inv), pack_bf16x2(v1[6] * inv, v1[7] * inv));
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM drained and smem operands free before the next tile overwrites them
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int launch_graph_aggregate_mma(const float* pos, const uint8_t* valid, const void* hb, const float* c, int S, int N,
                               float r2, float inv_2sigma2, void* mhb, void* mcb, cudaStream_t stream) {
  const int R = S * N, tiles = (R + 127) / 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(graph_aggregate_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SM_TOTAL + 1024);
    attr_set = true;
  }
  const int grid = tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs;
  graph_aggregate_mma_kernel<<<grid, GM_THREADS, GM_SM_TOTAL + 1024, stream>>>(
      pos, valid, reinterpret_cast<const __nv_bfloat16*>(hb), c, R, N, r2, -inv_2sigma2 * 1.4426950408889634f,
      reinterpret_cast<__nv_bfloat16*>(mhb), reinterpret_cast<__nv_bfloat16*>(mcb), tiles);
  count_launch();
  return check_launch("graph_aggregate_mma_kernel");
}

}  // namespace mmt
