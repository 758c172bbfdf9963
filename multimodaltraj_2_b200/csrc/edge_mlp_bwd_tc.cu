// Backward pass of the relational edge MLP on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in TMEM);
// training of g2k_lstm_mcr with Trainer(gemm="tc"), U = He = 128.  include/mmt.h: mmt_edge_mlp_backward_bf16.
// The fp32 CUDA-core version (edge_mlp_bwd.cu) states the algebra; reference: relational_inf_models/nri_learned.py:5-28,
// models/g2k_lstm_mcr.py:99-124 (the scores enter the attention logits), train.py:240-254 (the step that is differentiated).
//
// Per tile of 128 EDGES of the adjacency mask (compacted on the device as in edge_mlp_tc.cu; nothing goes to the host):
//   e1 tile (gathered node projections -> elu -> bf16 SWIZZLE_128B, rows = edges, k contiguous)
//   MMA 1  pre2[edge, c]  = e1[edge, k] W2[k, c]          A = e1 tile K-major,  B = packed W2 image K-major      -> TMEM D1
//   pass 1 (from TMEM)     z = w_out . elu(pre2 + b2) + b_out, s = sigmoid(z), du = d logit * s (1 - s)
//   pass 2 (from TMEM)     d pre2 = du w_out elu'(pre2) -> bf16 tile (rows = edges, c contiguous); column sums for g b2, g w_out
//   MMA 2  gW2[k, c]     += e1[edge, k] d pre2[edge, c]   A = e1 tile MN-major, B = d pre2 tile MN-major (K = edges) -> TMEM D2,
//                                                         accumulated over ALL tiles of the CTA, flushed once with red.global
//   MMA 3  d e1[edge, k]  = d pre2[edge, c] W2[k, c]      A = d pre2 tile K-major, B = the same W2 image read MN-major -> TMEM D1
//   pass 3 (from TMEM)     d pre1 = d e1 * elu'(e1) -> red.global.add.v4 into d a_i and d b_j
// So each of the two smem tiles and the one weight image serves two MMAs in two majornesses; no transposed copy exists.
#include <cuda_bf16.h>

#include "mmt_common.cuh"
#include "tc_common.cuh"

namespace mmt {

constexpr int EB_BLK = 128 * 128;                   // one [128 rows x 128 B] operand block (64 bf16 per row)
constexpr int EB_W1_BYTES = 2 * 256 * 128;          // offset of the W2 image inside the packed edge weights (edge_mlp_tc.cu)
constexpr int EB_W2_BYTES = 2 * EB_BLK;
constexpr int EB_LIST = 4096;                       // candidate pairs scanned per chunk (>= edges found)
constexpr int EB_SM_W2 = 0;                         // 32 KB  [k-block 2][c row 128][128 B]
constexpr int EB_SM_E1 = EB_SM_W2 + EB_W2_BYTES;    // 32 KB  [k-block 2][edge row 128][128 B]
constexpr int EB_SM_D2 = EB_SM_E1 + 2 * EB_BLK;     // 32 KB  [c-block 2][edge row 128][128 B]
constexpr int EB_SM_LIST = EB_SM_D2 + 2 * EB_BLK;   // u64[EB_LIST + 128]: scene << 32 | i << 16 | j; + a carried partial tile
constexpr int EB_SM_PAR = EB_SM_LIST + (EB_LIST + 128) * 8;   // b1[128] b2[128] w_out[128]
constexpr int EB_SM_PART = EB_SM_PAR + 3 * 128 * 4; // float[2][128]: per column half partial sums of z
constexpr int EB_SM_BAR = EB_SM_PART + 1024;
constexpr int EB_SM_TOTAL = EB_SM_BAR + 48 + 32;
constexpr uint32_t kIdescFwd = make_idesc_bf16(128, 128);                             // A K-major, B K-major
constexpr uint32_t kIdescGW2 = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);   // A MN-major, B MN-major
constexpr uint32_t kIdescDE1 = make_idesc_bf16(128, 128) | (1u << 16);                // A K-major, B MN-major

// MN-major SWIZZLE_128B descriptor over a [blocks of 64 MN elements][K rows][128 B] image: LBO = one block, SBO = 8 K rows
// (graph_mma.cu uses the same form for its state operand)
__device__ __forceinline__ uint64_t eb_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(EB_BLK >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float elu_fast_b(float x) { return x > 0.f ? x : ex2_fast(x * 1.4426950408889634f) - 1.0f; }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256, 1) edge_mlp_bwd_tc_kernel(
    const __nv_bfloat16* __restrict__ nab,   // [R, 256] = [a | b] node projections (node_proj_tc_kernel)
    const uint8_t* __restrict__ adj, const float* __restrict__ dlogit, const uint8_t* __restrict__ Wp,
    const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ w_out, const float* __restrict__ b_out,
    int S, int N, float* __restrict__ dab /* [R, 256] = [d a | d b], zeroed */, float* __restrict__ gW2,
    float* __restrict__ gb2, float* __restrict__ gw_out, float* __restrict__ gb_out, uint32_t* trap) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* const smem = smem_dyn;
  require_smem_alignment(smem, trap, 7);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = sbase + EB_SM_BAR, bar_mma = bar_w + 8;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + EB_SM_BAR + 16);
  int* s_wtot = reinterpret_cast<int*>(smem + EB_SM_BAR + 48);   // 8 warp totals of the edge scan
  unsigned long long* s_list = reinterpret_cast<unsigned long long*>(smem + EB_SM_LIST);
  float* s_b1 = reinterpret_cast<float*>(smem + EB_SM_PAR);
  float* s_b2 = s_b1 + 128;
  float* s_wo = s_b2 + 128;
  float* s_part = reinterpret_cast<float*>(smem + EB_SM_PART);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(bar_w, EB_W2_BYTES);
    bulk_g2s(sbase + EB_SM_W2, Wp + EB_W1_BYTES, EB_W2_BYTES, bar_w);
  }
  if (warp == 0) tmem_alloc(sbase + EB_SM_BAR + 16, 256);
  if (tid < 128) {
    s_b1[tid] = b1[tid];
    s_b2[tid] = b2[tid];
    s_wo[tid] = w_out[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const float bo = __ldg(b_out);
  mbar_wait(bar_w, 0, trap, 0x701);
  const int rows_per_chunk = EB_LIST / N > 0 ? EB_LIST / N : 1;
  uint32_t ph = 0;          // completed phases of bar_mma
  uint32_t tiles = 0;       // tiles this CTA has run (D2 accumulates from the second on)

  // epilogue role: TMEM lane quarter q = warp % 4 -> edge row r = 32 q + lane; column half = warp / 4 -> 64 columns
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + half * 64;
  float acc_b2[64], acc_wo[64], acc_bo = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) acc_b2[i] = acc_wo[i] = 0.f;

  auto edge_tile = [&](int t0, int nt) {
    // ---- e1 tile: warp w builds edges 16 w .. 16 w + 15; lane -> 4 consecutive k (one coalesced 256-byte row per gather)
    {
      const int k = lane * 4;
      const float4 c4 = *reinterpret_cast<const float4*>(s_b1 + k);
      uint8_t* blk = smem + EB_SM_E1 + (k >> 6) * EB_BLK + ((k & 7) << 1);
      const int chunk = (k & 63) >> 3;
#pragma unroll
      for (int e0 = 0; e0 < 16; e0 += 8) {
        uint2 av[8], bv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = warp * 16 + e0 + u;
          av[u] = bv[u] = make_uint2(0u, 0u);
          if (e < nt) {
            const unsigned long long en = s_list[t0 + e];
            const uint32_t ij = (uint32_t)en;
            const __nv_bfloat16* nab_s = nab + (size_t)(en >> 32) * N * 256;
            av[u] = __ldg(reinterpret_cast<const uint2*>(nab_s + (size_t)(ij >> 16) * 256) + lane);
            bv[u] = __ldg(reinterpret_cast<const uint2*>(nab_s + (size_t)(ij & 0xffffu) * 256 + 128) + lane);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = warp * 16 + e0 + u;   // rows beyond nt get elu(b1): finite, multiplied by d pre2 = 0 in MMA 2
          *reinterpret_cast<uint2*>(blk + e * 128 + ((chunk ^ (e & 7)) << 4)) =
              make_uint2(pack_bf16x2(elu_fast_b(bf16_lo(av[u].x) + bf16_lo(bv[u].x) + c4.x),
                                     elu_fast_b(bf16_hi(av[u].x) + bf16_hi(bv[u].x) + c4.y)),
                         pack_bf16x2(elu_fast_b(bf16_lo(av[u].y) + bf16_lo(bv[u].y) + c4.z),
                                     elu_fast_b(bf16_hi(av[u].y) + bf16_hi(bv[u].y) + c4.w)));
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    // ---- MMA 1: pre2 = e1 W2 -> D1 (TMEM columns 0..127)
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = make_desc_sw128(sbase + EB_SM_E1 + (ks >> 2) * EB_BLK) + (uint64_t)((ks & 3) * 2);
        const uint64_t db = make_desc_sw128(sbase + EB_SM_W2 + (ks >> 2) * EB_BLK) + (uint64_t)((ks & 3) * 2);
        umma_bf16(tmem_base, da, db, kIdescFwd, ks ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph & 1u, trap, 0x702);
    ++ph;
    tc_fence_after();
    // ---- pass 1: this thread's 64 columns of z = w_out . elu(pre2 + b2)
    {
      float part = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float v[8];
        tmem_ld8(t_row + ch * 8, v);
        tmem_wait_ld();
        const int c = half * 64 + ch * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) part = fmaf(elu_fast_b(v[i] + s_b2[c + i]), s_wo[c + i], part);
      }
      s_part[half * 128 + r] = part;
    }
    __syncthreads();
    float du = 0.f;
    unsigned long long en = 0ull;
    if (r < nt) {
      en = s_list[t0 + r];
      const uint32_t ij = (uint32_t)en;
      const float z = s_part[r] + s_part[128 + r] + bo;
      const float sc = 1.0f / (1.0f + __expf(-z));
      du = __ldg(dlogit + ((size_t)(en >> 32) * N + (ij >> 16)) * N + (ij & 0xffffu)) * sc * (1.0f - sc);
    }
    if (half == 0) acc_bo += du;
    // ---- pass 2: d pre2 = du w_out elu'(pre2) -> bf16 operand tile; column sums for g b2 and g w_out
    {
      uint8_t* drow = smem + EB_SM_D2 + half * EB_BLK + r * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float v[8], d2[8];
        tmem_ld8(t_row + ch * 8, v);
        tmem_wait_ld();
        const int c = half * 64 + ch * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pre = v[i] + s_b2[c + i];
          const float e2 = elu_fast_b(pre);
          d2[i] = du * s_wo[c + i] * (pre > 0.f ? 1.0f : e2 + 1.0f);
          acc_wo[ch * 8 + i] = fmaf(du, e2, acc_wo[ch * 8 + i]);
          acc_b2[ch * 8 + i] += d2[i];
        }
        *reinterpret_cast<uint4*>(drow + ((ch ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(d2[0], d2[1]), pack_bf16x2(d2[2], d2[3]), pack_bf16x2(d2[4], d2[5]), pack_bf16x2(d2[6], d2[7]));
      }
    }
    tc_fence_before();      // D1 has been read: MMA 3 may overwrite it
    fence_proxy_async();    // the d pre2 tile is visible to the tensor pipe
    __syncthreads();
    // ---- MMA 2: g W2 += e1^T d pre2 (D2, columns 128..255);  MMA 3: d e1 = d pre2 W2^T (D1)
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)      // K = 16 edges per instruction = 2 KB of each tile
        umma_bf16(tmem_base + 128, eb_desc_mn(sbase + EB_SM_E1 + ks * 2048), eb_desc_mn(sbase + EB_SM_D2 + ks * 2048), kIdescGW2,
                  (tiles | ks) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {    // K = 16 columns c: A advances 32 B inside a block, B 16 rows of the W2 image
        const uint64_t da = make_desc_sw128(sbase + EB_SM_D2 + (ks >> 2) * EB_BLK) + (uint64_t)((ks & 3) * 2);
        umma_bf16(tmem_base, da, eb_desc_mn(sbase + EB_SM_W2 + ks * 2048), kIdescDE1, ks ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph & 1u, trap, 0x703);
    ++ph;
    ++tiles;
    tc_fence_after();
    // ---- pass 3: d pre1 = d e1 * elu'(e1) -> d a_i, d b_j   (the TMEM loads are .sync.aligned: every lane runs them)
    {
      const uint32_t ij = (uint32_t)en;
      float* da = dab + ((size_t)(en >> 32) * N + (ij >> 16)) * 256 + half * 64;
      float* db = dab + ((size_t)(en >> 32) * N + (ij & 0xffffu)) * 256 + 128 + half * 64;
      const uint8_t* erow = smem + EB_SM_E1 + half * EB_BLK + r * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float v[8];
        tmem_ld8(t_row + ch * 8, v);
        tmem_wait_ld();
        if (r < nt) {
          const uint4 e8 = *reinterpret_cast<const uint4*>(erow + ((ch ^ (r & 7)) << 4));
          const uint32_t ew[4] = {e8.x, e8.y, e8.z, e8.w};
          float d1[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float lo = bf16_lo(ew[i]), hi = bf16_hi(ew[i]);
            d1[2 * i] = v[2 * i] * (lo > 0.f ? 1.0f : lo + 1.0f);
            d1[2 * i + 1] = v[2 * i + 1] * (hi > 0.f ? 1.0f : hi + 1.0f);
          }
          red_add_v4(da + ch * 8, d1[0], d1[1], d1[2], d1[3]);
          red_add_v4(da + ch * 8 + 4, d1[4], d1[5], d1[6], d1[7]);
          red_add_v4(db + ch * 8, d1[0], d1[1], d1[2], d1[3]);
          red_add_v4(db + ch * 8 + 4, d1[4], d1[5], d1[6], d1[7]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // D1 drained, both tiles and the list entries of this tile free
  };

  int carry = 0;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const uint8_t* ad = adj + (size_t)s * N * N;
    for (int r0 = 0; r0 < N; r0 += rows_per_chunk) {
      // ---- compact the edges of rows [r0, r1) (edge_mlp_tc.cu: 16 entries per thread + block-wide exclusive scan)
      const int r1 = min(N, r0 + rows_per_chunk);
      const int tot = (r1 - r0) * N;
      const uint8_t* src = ad + (size_t)r0 * N;
      const int e_base = tid * 16;
      uint32_t wv[4] = {0u, 0u, 0u, 0u};
      if (e_base < tot) {
        if (e_base + 16 <= tot && (reinterpret_cast<uintptr_t>(src + e_base) & 15u) == 0u) {
          const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(src + e_base));
          wv[0] = v4.x; wv[1] = v4.y; wv[2] = v4.z; wv[3] = v4.w;
        } else {
          for (int bb = 0; bb < 16; ++bb)
            if (e_base + bb < tot && src[e_base + bb] != 0) wv[bb >> 2] |= 1u << ((bb & 3) * 8);
        }
      }
      uint32_t em = 0;
#pragma unroll
      for (int bb = 0; bb < 16; ++bb) em |= ((wv[bb >> 2] >> ((bb & 3) * 8)) & 0xffu) ? (1u << bb) : 0u;
      const int cnt = __popc(em);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      __syncthreads();
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      int base = carry + incl - cnt, ne = 0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {
        const int tw = s_wtot[w8];
        if (w8 < warp) base += tw;
        ne += tw;
      }
      if (em) {
        int row = r0 + e_base / N, col = e_base % N;
        for (int bb = 0; bb < 16; ++bb) {
          if (em & (1u << bb)) s_list[base++] = ((unsigned long long)s << 32) | ((uint32_t)row << 16) | (uint32_t)col;
          if (++col == N) {
            col = 0;
            ++row;
          }
        }
      }
      __syncthreads();
      const int n_all = carry + ne;
      const int n_full = n_all & ~127;
      for (int t0 = 0; t0 < n_full; t0 += 128) edge_tile(t0, 128);
      __syncthreads();
      carry = n_all - n_full;
      if (n_full > 0 && tid < carry) s_list[tid] = s_list[n_full + tid];   // n_full >= 128 > carry: no overlap
      __syncthreads();
    }
  }
  if (carry > 0) edge_tile(0, carry);

  // ---- flush: g W2 from TMEM D2 (lane = k, columns = c); the register column sums; sum of du
  if (tiles > 0) {
    const uint32_t t2 = t_row + 128;
    float* grow = gW2 + (size_t)r * 128 + half * 64;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float v[8];
      tmem_ld8(t2 + ch * 8, v);
      tmem_wait_ld();
      red_add_v4(grow + ch * 8, v[0], v[1], v[2], v[3]);
      red_add_v4(grow + ch * 8 + 4, v[4], v[5], v[6], v[7]);
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float sb = warp_sum(acc_b2[i]), sw = warp_sum(acc_wo[i]);
      if (lane == 0) {
        atomicAdd(gb2 + half * 64 + i, sb);
        atomicAdd(gw_out + half * 64 + i, sw);
      }
    }
    const float so = warp_sum(acc_bo);
    if (lane == 0 && half == 0) atomicAdd(gb_out, so);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int launch_node_proj_tc(const float* h, int ld_h, const void* packed, int R, __nv_bfloat16* out, cudaStream_t stream);

int launch_edge_mlp_bwd_tc(const float* h, int ld_h, const uint8_t* adj, const float* dlogit, const void* packed,
                           const mmt_edge_weights* w, int S, int N, float* dab, float* gW2, float* gb2, float* gw_out, float* gb_out,
                           float* nab, cudaStream_t stream) {
  const int R = S * N;
  static DeviceMask smem_opted[1];
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(&edge_mlp_bwd_tc_kernel), EB_SM_TOTAL + 1024, &smem_opted[0])) return rc;
  __nv_bfloat16* nabh = reinterpret_cast<__nv_bfloat16*>(nab);
  if (int rc = launch_node_proj_tc(h, ld_h, packed, R, nabh, stream)) return rc;
  if (cudaMemsetAsync(dab, 0, sizeof(float) * 256 * (size_t)R, stream) != cudaSuccess) {
    set_error("mmt_edge_mlp_backward_bf16: cudaMemsetAsync failed");
    return MMT_ECUDA;
  }
  const int grid = S < num_sms() ? S : num_sms();
  edge_mlp_bwd_tc_kernel<<<grid, 256, EB_SM_TOTAL + 1024, stream>>>(nabh, adj, dlogit, reinterpret_cast<const uint8_t*>(packed),
                                                                    w->b1, w->b2, w->w_out, w->b_out, S, N, dab, gW2, gb2, gw_out,
                                                                    gb_out, trap_record());
  count_launch();
  return check_launch("edge_mlp_bwd_tc_kernel");
}

}  // namespace mmt

extern "C" int mmt_edge_mlp_backward_bf16(const float* h, const uint8_t* adj, const float* dlogit, const void* packed,
                                          const float* b1, const float* b2, const float* w_out, const float* b_out, int S, int N,
                                          int U, int He, float* dab, float* gW2, float* gb2, float* gw_out, float* gb_out,
                                          float* work, size_t work_bytes, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need 0 < N <= 1024, N % 4 == 0");
  MMT_REQUIRE(U == 128 && He == 128, "tensor-core edge MLP backward is built for U = 128, He = 128");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(h && adj && dlogit && packed && b1 && b2 && w_out && b_out && dab && gW2 && gb2 && gw_out && gb_out && work,
              "all pointers required");
  MMT_ALIGNED(h);
  MMT_ALIGNED(packed);
  MMT_ALIGNED(dab);
  MMT_ALIGNED(gW2);
  MMT_ALIGNED(work);
  if (work_bytes < sizeof(float) * 2 * (size_t)S * N * He) {
    set_error("mmt_edge_mlp_backward_bf16: workspace too small");
    return MMT_EWORKSPACE;
  }
  mmt_edge_weights w{nullptr, b1, nullptr, b2, w_out, b_out, He};
  return launch_edge_mlp_bwd_tc(h, U, adj, dlogit, packed, &w, S, N, dab, gW2, gb2, gw_out, gb_out, work, (cudaStream_t)stream);
}
