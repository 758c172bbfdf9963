// Pairwise pedestrian-distance kernel + adjacency (include/mmt.h).
//
// HBM-bound: per scene-frame it reads 9N bytes and writes 5N^2 (+4N) bytes.  Every thread produces
// quads of four consecutive j for one i (float4 position loads through L1) so that each warp-wide
// store instruction is one fully coalesced 512-byte (kern, float4) or 128-byte (adj, uchar4)
// segment.  Row degrees are reduced with segmented warp shuffles.  The distance uses __fmul_rn/__fadd_rn so no FMA contraction can
// change the rounding: adj/deg are bit-identical to the fp32 oracle.
#include "mmt_common.cuh"

namespace mmt {

// Flat mapping: quad q = (scene, i, j0/4) in memory order, so thread q's 16-byte kern store and 4-byte adj store
// land at kern + 4q / adj + 4q: every warp store instruction is one contiguous 512 B / 128 B segment, with no
// shared-memory staging and no block-level synchronisation (positions are 8N bytes per scene and stay in L1).
// Each thread owns UNROLL quads per iteration so several independent loads and stores are in flight.
// exp(x) for x <= 0 as one MUFU.EX2 (kern = 2^(d2 * -log2(e)/2sigma^2)); relative error ~2^-22
__device__ __forceinline__ float exp2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Fast path for power-of-two N in [16, 256]: a warp walks UN consecutive row groups of ONE scene, so the
// column positions / validity of a lane (its j-quads are the same in every iteration) are loaded once and only
// the row position changes: 2 small loads + 2 coalesced stores + ~30 ALU instructions per quad.  QPL = quads per
// lane per row (1 for N <= 128 where a row is lpr <= 32 consecutive lanes; 2 for N = 256).  Row degrees are
// gathered into consecutive lanes so the chunk's degrees leave as one contiguous store.
template <int UN, int QPL>
__global__ void __launch_bounds__(256) pairwise_rows_kernel(const float* __restrict__ pos,
                                                            const uint8_t* __restrict__ valid, long long nchunks, int N,
                                                            int lpr, int lpr_shift, int n_shift, float r2,
                                                            float neg_inv_log2e, float* __restrict__ kern,
                                                            uint8_t* __restrict__ adj, int32_t* __restrict__ deg) {
  const int lane = threadIdx.x & 31;
  const int seg = QPL == 1 ? lpr : 32;                       // lanes per row
  const int seg_shift = QPL == 1 ? lpr_shift : 5;
  const int jq = lane & (seg - 1), row_off = lane >> seg_shift, rows_per_iter = 32 >> seg_shift;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ch = warp0; ch < nchunks; ch += nwarps) {
    const long long row_base = ch * (UN * rows_per_iter);
    const long long s = row_base >> n_shift;
    float4 pa[QPL], pb[QPL];
    uchar4 vj[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
      const int j0 = (jq + 32 * k) << 2;
      pa[k] = __ldg(reinterpret_cast<const float4*>(pos + (s * N + j0) * 2));
      pb[k] = __ldg(reinterpret_cast<const float4*>(pos + (s * N + j0) * 2) + 1);
      vj[k] = __ldg(reinterpret_cast<const uchar4*>(valid + s * N + j0));
    }
    float2 pi[UN];
    uint8_t vi[UN];
#pragma unroll
    for (int it = 0; it < UN; ++it) {
      const long long row = row_base + it * rows_per_iter + row_off;
      pi[it] = __ldg(reinterpret_cast<const float2*>(pos) + row);
      vi[it] = valid[row];
    }
    int cnt[UN];
#pragma unroll
    for (int it = 0; it < UN; ++it) {
      const long long row = row_base + it * rows_per_iter + row_off;
      const int i = (int)(row - (s << n_shift));
      const bool v = vi[it] != 0;
      cnt[it] = 0;
#pragma unroll
      for (int k = 0; k < QPL; ++k) {
        const int j0 = (jq + 32 * k) << 2;
        float d2[4];
        {
          float dx, dy;
          dx = __fsub_rn(pi[it].x, pa[k].x); dy = __fsub_rn(pi[it].y, pa[k].y);
          d2[0] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi[it].x, pa[k].z); dy = __fsub_rn(pi[it].y, pa[k].w);
          d2[1] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi[it].x, pb[k].x); dy = __fsub_rn(pi[it].y, pb[k].y);
          d2[2] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi[it].x, pb[k].z); dy = __fsub_rn(pi[it].y, pb[k].w);
          d2[3] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        }
        const bool a0 = v && vj[k].x && (j0 + 0 != i) && (d2[0] < r2);
        const bool a1 = v && vj[k].y && (j0 + 1 != i) && (d2[1] < r2);
        const bool a2 = v && vj[k].z && (j0 + 2 != i) && (d2[2] < r2);
        const bool a3 = v && vj[k].w && (j0 + 3 != i) && (d2[3] < r2);
        const long long q = row * lpr + jq + 32 * k;
        if (kern != nullptr) {
          float4 kv;
          kv.x = a0 ? exp2_approx(d2[0] * neg_inv_log2e) : 0.0f;
          kv.y = a1 ? exp2_approx(d2[1] * neg_inv_log2e) : 0.0f;
          kv.z = a2 ? exp2_approx(d2[2] * neg_inv_log2e) : 0.0f;
          kv.w = a3 ? exp2_approx(d2[3] * neg_inv_log2e) : 0.0f;
          st_cs_f4(kern + (q << 2), kv);
        }
        if (adj != nullptr) __stcs(reinterpret_cast<uchar4*>(adj + (q << 2)), make_uchar4(a0, a1, a2, a3));
        cnt[it] += (int)a0 + (int)a1 + (int)a2 + (int)a3;
      }
    }
    if (deg != nullptr) {
      // butterfly inside each row segment, then lane L < rows-per-chunk picks row L's total: one coalesced store
      int mine = 0;
#pragma unroll
      for (int it = 0; it < UN; ++it) {
        int c = cnt[it];
        for (int o = seg >> 1; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        const int val = __shfl_sync(0xffffffffu, c, (lane & (rows_per_iter - 1)) << seg_shift);
        if ((lane >> (5 - seg_shift)) == it) mine = val;
      }
      if (lane < UN * rows_per_iter) deg[row_base + lane] = mine;
    }
  }
}

template <int UNROLL>
__global__ void __launch_bounds__(256) pairwise_adj_kernel(const float* __restrict__ pos,
                                                           const uint8_t* __restrict__ valid, long long nquads, int N,
                                                           int lpr, int lpr_shift, int n_shift, float r2,
                                                           float neg_inv_log2e, float* __restrict__ kern,
                                                           uint8_t* __restrict__ adj, int32_t* __restrict__ deg) {
  const bool seg = lpr_shift >= 0 && lpr <= 32;   // a row = lpr consecutive lanes of one warp instruction
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 - threadIdx.x % 32 < nquads;
       q0 += stride * UNROLL) {
#pragma unroll
    for (int un = 0; un < UNROLL; ++un) {
      const long long q = q0 + un * stride;
      const bool active = q < nquads;
      int cnt = 0;
      long long row = 0;
      if (active) {
        int jq;
        if (lpr_shift >= 0) {
          row = q >> lpr_shift;
          jq = (int)(q & (lpr - 1));
        } else {
          row = q / lpr;
          jq = (int)(q - row * lpr);
        }
        const long long s = n_shift >= 0 ? (row >> n_shift) : (row / N);
        const int i = (int)(row - s * N), j0 = jq << 2;
        const float2 pi = __ldg(reinterpret_cast<const float2*>(pos) + row);
        const float4 pa = __ldg(reinterpret_cast<const float4*>(pos + (s * N + j0) * 2));
        const float4 pb = __ldg(reinterpret_cast<const float4*>(pos + (s * N + j0) * 2) + 1);
        const bool vi = valid[row] != 0;
        const uchar4 vj = __ldg(reinterpret_cast<const uchar4*>(valid + s * N + j0));
        float d2[4];
        {
          float dx, dy;
          dx = __fsub_rn(pi.x, pa.x); dy = __fsub_rn(pi.y, pa.y);
          d2[0] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi.x, pa.z); dy = __fsub_rn(pi.y, pa.w);
          d2[1] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi.x, pb.x); dy = __fsub_rn(pi.y, pb.y);
          d2[2] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          dx = __fsub_rn(pi.x, pb.z); dy = __fsub_rn(pi.y, pb.w);
          d2[3] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        }
        const bool a0 = vi && vj.x && (j0 + 0 != i) && (d2[0] < r2);
        const bool a1 = vi && vj.y && (j0 + 1 != i) && (d2[1] < r2);
        const bool a2 = vi && vj.z && (j0 + 2 != i) && (d2[2] < r2);
        const bool a3 = vi && vj.w && (j0 + 3 != i) && (d2[3] < r2);
        cnt = (int)a0 + (int)a1 + (int)a2 + (int)a3;
        if (kern != nullptr) {
          float4 k;
          k.x = a0 ? exp2_approx(d2[0] * neg_inv_log2e) : 0.0f;
          k.y = a1 ? exp2_approx(d2[1] * neg_inv_log2e) : 0.0f;
          k.z = a2 ? exp2_approx(d2[2] * neg_inv_log2e) : 0.0f;
          k.w = a3 ? exp2_approx(d2[3] * neg_inv_log2e) : 0.0f;
          st_cs_f4(kern + (q << 2), k);
        }
        if (adj != nullptr) __stcs(reinterpret_cast<uchar4*>(adj + (q << 2)), make_uchar4(a0, a1, a2, a3));
      }
      if (deg != nullptr) {
        if (seg) {
          for (int o = lpr >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          if (active && (lane & (lpr - 1)) == 0) deg[row] = cnt;
        } else if (active && cnt) {
          atomicAdd(&deg[row], cnt);   // deg is zeroed by the launcher on this path
        }
      }
    }
  }
}

// one warp per (scene,row): ordered compaction of the adjacency row into neighbour indices
__global__ void __launch_bounds__(256) neighbor_index_kernel(const uint8_t* __restrict__ adj, int rows, int N,
                                                             int max_nbr, int32_t* __restrict__ nbr,
                                                             int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const uint8_t* row = adj + (size_t)r * N;
    int32_t* out = nbr + (size_t)r * max_nbr;
    int n = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const bool a = j < N && row[j] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, a);
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      if (a && pos < max_nbr) out[pos] = j;
      n += __popc(m);
    }
    const int w = n < max_nbr ? n : max_nbr;
    for (int k = w + lane; k < max_nbr; k += 32) out[k] = -1;
    if (lane == 0) cnt[r] = w;
  }
}

}  // namespace mmt

extern "C" int mmt_pairwise_adj_f32(const float* pos, const uint8_t* valid, int S, int N, float r2,
                                    float inv_2sigma2, float* kern, uint8_t* adj, int32_t* deg, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N % 4 == 0 && N <= 1024, "need S >= 0, 0 < N <= 1024, N % 4 == 0");
  if (S == 0) return MMT_OK;  // empty batch: nothing to read or write
  MMT_REQUIRE(pos && valid, "pos/valid must not be NULL");
  MMT_ALIGNED(pos);
  MMT_ALIGNED(kern);
  if (adj && (reinterpret_cast<uintptr_t>(adj) & 3u)) {
    set_error("mmt_pairwise_adj_f32: adj not 4-byte aligned");
    return MMT_EALIGN;
  }
  if (S == 0) return MMT_OK;
  const int lpr = N >> 2;
  auto log2_exact = [](int v) { int sh = 0; while ((1 << sh) < v) ++sh; return (1 << sh) == v ? sh : -1; };
  const int lpr_shift = log2_exact(lpr), n_shift = log2_exact(N);
  const long long nquads = (long long)S * N * lpr;
  if (deg != nullptr && !(lpr_shift >= 0 && lpr <= 32)) cudaMemsetAsync(deg, 0, sizeof(int32_t) * (size_t)S * N, (cudaStream_t)stream);
  const float neg_inv_log2e = -inv_2sigma2 * 1.4426950408889634f;
  constexpr int UN = 4;
  if (n_shift >= 0 && N >= 16 && N <= 256) {
    // rows per chunk = UN * rows_per_iter divides N: N = 16 -> UN 2 (16 rows), 32/64/128 -> UN 4, 256 -> UN 4 (QPL 2)
    const long long rows = (long long)S * N;
    const int rpi = N <= 128 ? (32 >> lpr_shift) : 1;
    const int un = N == 16 ? 2 : UN;
    const long long nchunks = rows / (un * rpi);
    long long blocks = (nchunks + 7) / 8;
    int grid = blocks < (long long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 256)
      pairwise_rows_kernel<UN, 2><<<grid, 256, 0, st>>>(pos, valid, nchunks, N, lpr, lpr_shift, n_shift, r2,
                                                        neg_inv_log2e, kern, adj, deg);
    else if (N == 16)
      pairwise_rows_kernel<2, 1><<<grid, 256, 0, st>>>(pos, valid, nchunks, N, lpr, lpr_shift, n_shift, r2,
                                                       neg_inv_log2e, kern, adj, deg);
    else
      pairwise_rows_kernel<UN, 1><<<grid, 256, 0, st>>>(pos, valid, nchunks, N, lpr, lpr_shift, n_shift, r2,
                                                        neg_inv_log2e, kern, adj, deg);
    count_launch();
    return check_launch("pairwise_rows_kernel");
  }
  long long blocks = (nquads + 256LL * UN - 1) / (256LL * UN);
  // a multiple of the SM count with 8 resident CTAs per SM; larger inputs grid-stride
  int grid = blocks < (long long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  pairwise_adj_kernel<UN><<<grid, 256, 0, (cudaStream_t)stream>>>(pos, valid, nquads, N, lpr, lpr_shift, n_shift, r2,
                                                                  neg_inv_log2e, kern, adj, deg);
  count_launch();
  return check_launch("pairwise_adj_kernel");
}

extern "C" int mmt_neighbor_index_i32(const uint8_t* adj, int S, int N, int max_nbr, int32_t* nbr, int32_t* cnt,
                                      void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && max_nbr > 0, "need S >= 0, N > 0, max_nbr > 0");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(adj && nbr && cnt, "adj/nbr/cnt must not be NULL");
  if (S == 0) return MMT_OK;
  const long rows = (long)S * N;
  long blocks = (rows + 7) / 8;
  int grid = blocks < num_sms() * 8 ? (int)blocks : num_sms() * 8;
  neighbor_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(adj, (int)rows, N, max_nbr, nbr, cnt);
  count_launch();
  return check_launch("neighbor_index_kernel");
}
