// Reference-compatible ADE/FDE reductions and the small relational helpers (include/mmt.h).
//   mmt_mean_error_f32        sample.get_mean_error (sample.py:21-82), value for value
//   mmt_train_val_scores_f32  the validation scores of train.py:639-674 (spectral norm per agent)
//   mmt_mcr_forward_f32       g2k_lstm_mcr.forward alone (models/g2k_lstm_mcr.py:99-124) on given `outputs`
//   mmt_sigmoid_f32 / mmt_rowsoftmax_f32   nri_learned.infer_rlns / eval_rln_ngh (nri_learned.py:16-28)
// All are tiny latency-bound reductions: one CTA (or one warp per row) each.
#include "mmt_common.cuh"

namespace mmt {

// predicted / true: [n, L, 2] agent-major.  error_t = sum_j (true - pred)[t, j] for t in [obs, L);
// ade = mean_t ||error_t|| / counter, counter = (L-obs)*maxNumPeds; fde = mean_j ||(true-pred)[L-1, j]|| / maxNumPeds
__global__ void __launch_bounds__(256) mean_error_kernel(const float* __restrict__ pred, const float* __restrict__ tru,
                                                         int n, int L, int obs, int maxp, float* __restrict__ out) {
  __shared__ float s_part[256];
  const int tid = threadIdx.x;
  float ade_acc = 0.f;
  for (int t = obs; t < L; ++t) {
    float ex = 0.f, ey = 0.f;
    for (int j = tid; j < maxp; j += blockDim.x) {
      ex += tru[((size_t)j * L + t) * 2] - pred[((size_t)j * L + t) * 2];
      ey += tru[((size_t)j * L + t) * 2 + 1] - pred[((size_t)j * L + t) * 2 + 1];
    }
    s_part[tid] = ex;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) s_part[tid] += s_part[tid + o];
      __syncthreads();
    }
    const float sx = s_part[0];
    __syncthreads();
    s_part[tid] = ey;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) s_part[tid] += s_part[tid + o];
      __syncthreads();
    }
    const float sy = s_part[0];
    __syncthreads();
    ade_acc += sqrtf(sx * sx + sy * sy);
  }
  float f = 0.f;
  for (int j = tid; j < maxp; j += blockDim.x) {
    const float dx = tru[((size_t)j * L + L - 1) * 2] - pred[((size_t)j * L + L - 1) * 2];
    const float dy = tru[((size_t)j * L + L - 1) * 2 + 1] - pred[((size_t)j * L + L - 1) * 2 + 1];
    f += sqrtf(dx * dx + dy * dy);
  }
  s_part[tid] = f;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_part[tid] += s_part[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    const float counter = (float)(L - obs) * (float)maxp;
    out[0] = ade_acc / (float)(L - obs) / counter;
    out[1] = s_part[0] / (float)maxp / (float)maxp;
    out[2] = counter;
  }
}

// per agent i: M = pred[i,:Li] - tgt[i,:Li] ([Li,2]); euc_i = sigma_max(M)/12 (closed form from the 2x2 Gram
// matrix), err_i = M[Li-1].  out: euc[n], err[n,2].  len[i] = Li (<= P).  short tracks also divide by n_targets.
__global__ void __launch_bounds__(128) train_val_scores_kernel(const float* __restrict__ pred,
                                                               const float* __restrict__ tgt,
                                                               const int32_t* __restrict__ len, int n, int P,
                                                               int n_targets, float* __restrict__ euc,
                                                               float* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int Li = len[i] < P ? len[i] : P;
  float a = 0.f, b = 0.f, c = 0.f, lx = 0.f, ly = 0.f;
  for (int t = 0; t < Li; ++t) {
    const float dx = pred[((size_t)i * P + t) * 2] - tgt[((size_t)i * P + t) * 2];
    const float dy = pred[((size_t)i * P + t) * 2 + 1] - tgt[((size_t)i * P + t) * 2 + 1];
    a += dx * dx;
    b += dx * dy;
    c += dy * dy;
    lx = dx;
    ly = dy;
  }
  const float tr = a + c, det = a * c - b * b;
  const float disc = sqrtf(fmaxf(tr * tr * 0.25f - det, 0.f));
  float e = sqrtf(fmaxf(tr * 0.5f + disc, 0.f)) / 12.0f;
  if (len[i] < P) e /= (float)n_targets;
  euc[i] = e;
  err[i * 2] = lx;
  err[i * 2 + 1] = ly;
}

// A.4 alone: outputs[D+2,D] rel[2,D] ngh[D,T] per scene -> attn[D,D], cost[T,T], band[2P,n]
__global__ void __launch_bounds__(128) mcr_forward_kernel(const float* __restrict__ outputs,
                                                          const float* __restrict__ rel, const float* __restrict__ ngh,
                                                          const float* __restrict__ W_v, const float* __restrict__ b_v,
                                                          const float* __restrict__ W_r, const float* __restrict__ W_c,
                                                          const float* __restrict__ W_o, int S, int n, int D, int T, int P,
                                                          float lam, int variant, float* __restrict__ attn,
                                                          float* __restrict__ cost, float* __restrict__ band) {
  __shared__ float sEo[16 * 16], sM[16 * 16], sNg[16 * 16], sCost[16 * 16], sWC[32 * 16];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const float* o = outputs + (size_t)s * (D + 2) * D;
    const float* r = rel + (size_t)s * 2 * D;
    for (int i = tid; i < D * T; i += nt) sNg[i] = lam * ngh[(size_t)s * D * T + i];
    for (int i = tid; i < T * D; i += nt) {
      const int t = i / D, d = i - t * D;
      float acc = 0.f;
      for (int k = 0; k < D + 2; ++k) acc = fmaf(W_v[t * (D + 2) + k], o[k * D + d], acc);
      acc += b_v[d];
      sEo[i] = acc;
      sM[i] = acc * (W_r[t * 2] * r[d] + W_r[t * 2 + 1] * r[D + d]);
    }
    __syncthreads();
    for (int i = tid; i < D * D + T * T; i += nt) {
      if (i < D * D) {
        const int a = i / D, d = i - a * D;
        float acc = 0.f;
        for (int k = 0; k < T; ++k) acc = fmaf(sNg[a * T + k], sM[k * D + d], acc);
        if (attn) attn[(size_t)s * D * D + i] = acc;
      } else {
        const int q = i - D * D, a = q / T, t = q - a * T;
        float acc = 0.f;
        if (variant == 0)
          for (int k = 0; k < D; ++k) acc = fmaf(sEo[a * D + k], sNg[k * T + t], acc);
        sCost[q] = acc;
        if (cost) cost[(size_t)s * T * T + q] = acc;
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * P * T; i += nt) {
      const int a = i / T, t = i - a * T;
      float acc = 0.f;
      for (int k = 0; k < T; ++k) acc = fmaf(W_c[a * T + k], sCost[k * T + t], acc);
      sWC[i] = acc;
    }
    __syncthreads();
    if (band)
      for (int i = tid; i < 2 * P * n; i += nt) {
        const int a = i / n, c = i - a * n;
        float acc = 0.f;
        for (int k = 0; k < T; ++k) acc = fmaf(sWC[a * T + k], W_o[k * n + c], acc);
        band[(size_t)s * 2 * P * n + i] = acc;
      }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = sigmoid_acc(x[i]);
}

__global__ void __launch_bounds__(256) rowsoftmax_kernel(const float* __restrict__ x, float* __restrict__ y, int rows,
                                                         int cols) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* xr = x + (size_t)r * cols;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += expf(xr[c] - mx);
    sum = warp_sum(sum);
    for (int c = lane; c < cols; c += 32) y[(size_t)r * cols + c] = expf(xr[c] - mx) / sum;
  }
}


// ADE / FDE in world coordinates (metres): the data files hold pixel coordinates divided by (480, 640)
// (data/eth/univ/getPixelCoordinates.m:27-30, written as pinv(H) * world there); a point goes back through
// p = (pos0 * s0, pos1 * s1, 1), w = H p, (X, Y) = w[0:2] / w[2].  HBM-bound: 16 P + 8 bytes per agent.  A group of
// 16 lanes (32 when P > 16) walks one agent's steps (coalesced float2 loads), errors are combined by shuffles in a
// fixed order; sums[3] = (sum ADE, sum FDE, number of valid agents): per-warp partial sums in registers, one shared-
// memory pass per CTA, three atomics per CTA (one atomic per agent serialised the whole kernel on a single address:
// 1.2 ms for 262 144 agents).
__global__ void __launch_bounds__(256) ade_fde_world_kernel(const float* pred, const float* gt, const uint8_t* valid, int n,
                                                            int P, const float* Hm, float s0, float s1, float* ade,
                                                            float* fde, float* sums) {
  __shared__ float s_part[8][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpa = P <= 16 ? 16 : 32;            // lanes per agent
  const int gpw = 32 / lpa;                     // agents per warp and iteration
  const int gl = lane & (lpa - 1), grp = lane / lpa;
  const int gpb = (blockDim.x >> 5) * gpw;      // agents per CTA and iteration
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = __ldg(Hm + i);
  float sa = 0.f, sf = 0.f, sn = 0.f;           // this group's running sums (held by its lane 0)
  const int iters = (n + gridDim.x * gpb - 1) / (gridDim.x * gpb);
  for (int it = 0; it < iters; ++it) {          // every lane takes part in every shuffle
    const int ag = (it * gridDim.x + blockIdx.x) * gpb + warp * gpw + grp;
    const bool in = ag < n;
    const bool v = in && (valid == nullptr || valid[ag] != 0);
    float acc = 0.f, last = 0.f;
    if (in)
      for (int t = gl; t < P; t += lpa) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(pred) + (size_t)ag * P + t);
        const float2 b = __ldg(reinterpret_cast<const float2*>(gt) + (size_t)ag * P + t);
        auto world = [&](float2 q) {
          const float u = q.x * s0, w = q.y * s1;
          const float X = fmaf(h[0], u, fmaf(h[1], w, h[2])), Y = fmaf(h[3], u, fmaf(h[4], w, h[5])),
                      Z = fmaf(h[6], u, fmaf(h[7], w, h[8]));
          return make_float2(__fdiv_rn(X, Z), __fdiv_rn(Y, Z));
        };
        const float2 wa = world(a), wb = world(b);
        const float dx = wa.x - wb.x, dy = wa.y - wb.y;
        const float d = sqrtf(fmaf(dx, dx, dy * dy));
        acc += d;
        if (t == P - 1) last = d;               // exactly one lane of the group holds the final-step error
      }
    for (int o = lpa >> 1; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      last += __shfl_xor_sync(0xffffffffu, last, o);
    }
    if (in && gl == 0) {
      const float a_ = v ? acc / (float)P : 0.f, f_ = v ? last : 0.f;
      ade[ag] = a_;
      fde[ag] = f_;
      sa += a_;
      sf += f_;
      sn += v ? 1.f : 0.f;
    }
  }
  if (sums) {
    sa = warp_sum(sa); sf = warp_sum(sf); sn = warp_sum(sn);     // lanes other than the group leaders hold 0
    if (lane == 0) { s_part[warp][0] = sa; s_part[warp][1] = sf; s_part[warp][2] = sn; }
    __syncthreads();
    if (threadIdx.x < 3) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w][threadIdx.x];
      atomicAdd(sums + threadIdx.x, t);
    }
  }
}

}  // namespace mmt

extern "C" int mmt_mean_error_f32(const float* predicted, const float* truth, int n, int L, int observed_length,
                                  int maxNumPeds, float* out3, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(predicted && truth && out3, "pointers required");
  MMT_REQUIRE(n > 0 && L > observed_length && observed_length >= 0 && maxNumPeds > 0 && maxNumPeds <= n,
              "need L > observed_length, 0 < maxNumPeds <= n");
  mean_error_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(predicted, truth, n, L, observed_length, maxNumPeds, out3);
  count_launch();
  return check_launch("mean_error_kernel");
}

extern "C" int mmt_train_val_scores_f32(const float* pred, const float* tgt, const int32_t* len, int n, int P,
                                        int n_targets, float* euc, float* err, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(n >= 0 && P > 0 && n_targets > 0, "need P > 0, n_targets > 0");
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(pred && tgt && len && euc && err, "pointers required");
  train_val_scores_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pred, tgt, len, n, P, n_targets, euc, err);
  count_launch();
  return check_launch("train_val_scores_kernel");
}

extern "C" int mmt_mcr_forward_f32(const float* outputs, const float* rel, const float* ngh, const mmt_mcr_weights* w,
                                   int S, int n, int D, int T, int P, float lam, int variant, float* attn, float* cost,
                                   float* band, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(S >= 0 && n > 0 && D > 0 && D <= 16 && T > 0 && T <= 16 && P > 0 && 2 * P <= 32,
              "need D <= 16, T <= 16, 2P <= 32");
  if (S == 0) return MMT_OK;
  MMT_REQUIRE(outputs && rel && ngh && w && w->W_v && w->b_v && w->W_r && w->W_c && w->W_o, "pointers required");
  int grid = S < num_sms() * 8 ? S : num_sms() * 8;
  mcr_forward_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(outputs, rel, ngh, w->W_v, w->b_v, w->W_r, w->W_c, w->W_o,
                                                             S, n, D, T, P, lam, variant, attn, cost, band);
  count_launch();
  return check_launch("mcr_forward_kernel");
}

extern "C" int mmt_sigmoid_f32(const float* x, float* y, size_t n, void* stream) {
  using namespace mmt;
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(x && y, "pointers required");
  size_t blocks = (n + 255) / 256;
  int grid = blocks < (size_t)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  sigmoid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n);
  count_launch();
  return check_launch("sigmoid_kernel");
}

extern "C" int mmt_rowsoftmax_f32(const float* x, float* y, int rows, int cols, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(rows >= 0 && cols > 0, "need cols > 0");
  if (rows == 0) return MMT_OK;
  MMT_REQUIRE(x && y, "pointers required");
  int blocks = (rows + 7) / 8;
  int grid = blocks < num_sms() * 8 ? blocks : num_sms() * 8;
  rowsoftmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols);
  count_launch();
  return check_launch("rowsoftmax_kernel");
}

extern "C" int mmt_ade_fde_world_f32(const float* pred, const float* gt, const uint8_t* valid, int n, int P,
                                     const float* Hm, float scale0, float scale1, float* ade, float* fde, float* sums,
                                     void* stream) {
  using namespace mmt;
  MMT_REQUIRE(n >= 0 && P > 0, "need P > 0");
  if (n == 0) return MMT_OK;
  MMT_REQUIRE(pred && gt && Hm && ade && fde, "pointers required");
  MMT_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 7u) == 0 && (reinterpret_cast<uintptr_t>(gt) & 7u) == 0,
              "pred / gt must be 8-byte aligned");
  if (sums) cudaMemsetAsync(sums, 0, 3 * sizeof(float), (cudaStream_t)stream);
  const int blocks = (n + 15) / 16;
  const int grid = blocks < num_sms() * 8 ? blocks : num_sms() * 8;
  ade_fde_world_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, valid, n, P, Hm, scale0, scale1, ade, fde, sums);
  count_launch();
  return check_launch("ade_fde_world_kernel");
}
