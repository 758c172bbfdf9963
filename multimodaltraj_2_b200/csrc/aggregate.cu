// Graph aggregation of neighbour states (generalises train.py:240-247).
//
// a = softmax over {j : adj_ij} of (logits_ij [+ logits2_ij]); out_i = sum_j a_ij feat_j.
// One warp per row: the masked softmax is warp-shuffle reduced, the (sparse) neighbour set is
// compacted with ballots into shared memory, then the warp streams the neighbours' feature rows
// as float4 (coalesced 512 B per neighbour chunk) and accumulates in registers.  HBM-bound:
// reads 5N^2 + 4NC, writes 4NC (+4N^2 when attn is requested) bytes per scene.
#include "mmt_common.cuh"

namespace mmt {

constexpr int kAggWarps = 8;

__global__ void __launch_bounds__(kAggWarps * 32) aggregate_kernel(
    const float* __restrict__ logits, const float* __restrict__ logits2, const uint8_t* __restrict__ adj,
    const float* __restrict__ feat, int rows, int N, int C, int ld_feat, float* __restrict__ attn,
    float* __restrict__ out, int ld_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* nb_idx = reinterpret_cast<int*>(smem_raw) + warp * N;                    // [warps][N]
  float* nb_w = reinterpret_cast<float*>(smem_raw) + kAggWarps * N + warp * N;  // [warps][N]

  for (int r = blockIdx.x * kAggWarps + warp; r < rows; r += gridDim.x * kAggWarps) {
    const int s = r / N;
    const float* lrow = logits + (size_t)r * N;
    const float* l2row = logits2 ? logits2 + (size_t)r * N : nullptr;
    const uint8_t* arow = adj + (size_t)r * N;
    // pass 1: compact neighbours, track max
    int n = 0;
    float mx = -INFINITY;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const bool a = j < N && arow[j] != 0;
      float v = 0.0f;
      if (a) {
        v = lrow[j];
        if (l2row) v += l2row[j];
      }
      const unsigned m = __ballot_sync(0xffffffffu, a);
      if (a) {
        const int p = n + __popc(m & ((1u << lane) - 1u));
        nb_idx[p] = j;
        nb_w[p] = v;
        mx = fmaxf(mx, v);
      }
      n += __popc(m);
    }
    mx = warp_max(mx);
    __syncwarp();
    // pass 2: exp / sum
    float sum = 0.0f;
    for (int k = lane; k < n; k += 32) {
      const float e = expf(nb_w[k] - mx);
      nb_w[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = n > 0 ? 1.0f / sum : 0.0f;
    __syncwarp();
    for (int k = lane; k < n; k += 32) nb_w[k] *= inv;
    __syncwarp();
    if (attn != nullptr) {
      float* arow_out = attn + (size_t)r * N;
      for (int j = lane; j < N; j += 32) arow_out[j] = 0.0f;
      __syncwarp();
      for (int k = lane; k < n; k += 32) arow_out[nb_idx[k]] = nb_w[k];
    }
    // pass 3: weighted sum of neighbour feature rows
    const float* fbase = feat + (size_t)s * N * ld_feat;
    float* orow = out + (size_t)r * ld_out;
    for (int c0 = lane * 4; c0 < C; c0 += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < n; ++k) {
        const float w = nb_w[k];
        const float4 f = __ldg(reinterpret_cast<const float4*>(fbase + (size_t)nb_idx[k] * ld_feat + c0));
        acc.x = fmaf(w, f.x, acc.x);
        acc.y = fmaf(w, f.y, acc.y);
        acc.z = fmaf(w, f.z, acc.z);
        acc.w = fmaf(w, f.w, acc.w);
      }
      *reinterpret_cast<float4*>(orow + c0) = acc;
    }
    __syncwarp();
  }
}

// Transposed aggregation of the backward pass: out_j = sum_i a_ij d_i (the adjoint of out_i = sum_j a_ij feat_j),
// i.e. att^T x d per scene.  One warp per output row j: the (sparse) column j of the attention matrix is compacted
// with ballots, then the warp streams the rows d_i of its in-neighbours as float4 -- a gather in a fixed order
// (ascending i), so the result is deterministic (no atomics).
__global__ void __launch_bounds__(kAggWarps * 32) aggregate_transpose_kernel(
    const float* __restrict__ att, const float* __restrict__ d, int rows, int N, int C, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* nb_idx = reinterpret_cast<int*>(smem_raw) + warp * N;
  float* nb_w = reinterpret_cast<float*>(smem_raw) + kAggWarps * N + warp * N;
  for (int r = blockIdx.x * kAggWarps + warp; r < rows; r += gridDim.x * kAggWarps) {
    const int s = r / N, j = r - s * N;
    const float* acol = att + (size_t)s * N * N + j;
    int n = 0;
    for (int i0 = 0; i0 < N; i0 += 32) {
      const int i = i0 + lane;
      const float w = i < N ? __ldg(acol + (size_t)i * N) : 0.0f;
      const bool a = w != 0.0f;
      const unsigned m = __ballot_sync(0xffffffffu, a);
      if (a) {
        const int p = n + __popc(m & ((1u << lane) - 1u));
        nb_idx[p] = i;
        nb_w[p] = w;
      }
      n += __popc(m);
    }
    __syncwarp();
    const float* dbase = d + (size_t)s * N * C;
    float* orow = out + (size_t)r * C;
    for (int c0 = lane * 4; c0 < C; c0 += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < n; ++k) {
        const float w = nb_w[k];
        const float4 f = __ldg(reinterpret_cast<const float4*>(dbase + (size_t)nb_idx[k] * C + c0));
        acc.x = fmaf(w, f.x, acc.x);
        acc.y = fmaf(w, f.y, acc.y);
        acc.z = fmaf(w, f.z, acc.z);
        acc.w = fmaf(w, f.w, acc.w);
      }
      *reinterpret_cast<float4*>(orow + c0) = acc;
    }
    __syncwarp();
  }
}

int launch_aggregate(const float* logits, const float* logits2, const uint8_t* adj, const float* feat, int S, int N,
                     int C, int ld_feat, float* attn, float* out, int ld_out, cudaStream_t stream) {
  const long rows = (long)S * N;
  long blocks = (rows + kAggWarps - 1) / kAggWarps;
  int grid = blocks < (long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  const size_t smem = (size_t)kAggWarps * N * 8;
  if (smem > 48 * 1024) {   // N > 768: beyond the default dynamic shared-memory limit (N <= 1024 -> 64 KB)
    static DeviceMask smem_opted[1];
    if (int rc = opt_in_smem(reinterpret_cast<const void*>(&aggregate_kernel), kAggWarps * 1024 * 8, &smem_opted[0])) return rc;
  }
  aggregate_kernel<<<grid, kAggWarps * 32, smem, stream>>>(logits, logits2, adj, feat, (int)rows, N, C, ld_feat, attn,
                                                           out, ld_out);
  count_launch();
  return check_launch("aggregate_kernel");
}

}  // namespace mmt

extern "C" int mmt_aggregate_f32(const float* logits, const uint8_t* adj, const float* feat, int S, int N, int C,
                                 float* attn, float* out, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N <= 1024 && C > 0 && C % 4 == 0, "need 0 < N <= 1024, C % 4 == 0");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(logits && adj && feat && out, "logits/adj/feat/out must not be NULL");
  MMT_ALIGNED(feat);
  MMT_ALIGNED(out);
  if (S == 0) return MMT_OK;
  return launch_aggregate(logits, nullptr, adj, feat, S, N, C, C, attn, out, C, (cudaStream_t)stream);
}

extern "C" int mmt_aggregate_transpose_f32(const float* attn, const float* d, int S, int N, int C, float* out,
                                           void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && N > 0 && N <= 1024 && C > 0 && C % 4 == 0, "need 0 < N <= 1024, C % 4 == 0");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(attn && d && out, "attn/d/out must not be NULL");
  MMT_ALIGNED(d);
  MMT_ALIGNED(out);
  const long rows = (long)S * N;
  const long blocks = (rows + kAggWarps - 1) / kAggWarps;
  const int grid = blocks < (long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  const size_t smem = (size_t)kAggWarps * N * 8;
  if (smem > 48 * 1024) {
    static DeviceMask smem_opted[1];
    if (int rc = opt_in_smem(reinterpret_cast<const void*>(&aggregate_transpose_kernel), kAggWarps * 1024 * 8, &smem_opted[0])) return rc;
  }
  aggregate_transpose_kernel<<<grid, kAggWarps * 32, smem, (cudaStream_t)stream>>>(attn, d, (int)rows, N, C, out);
  count_launch();
  return check_launch("aggregate_transpose_kernel");
}
