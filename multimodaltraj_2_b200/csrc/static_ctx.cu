// Static-context branch: the scene image correlated with one huge filter, the way
// train.py:93-110 does it once per dataset, and the neighbourhood input ngh = conv x stat_mask of train.py:154-158.
//   imgp   = pad(img[H,W,C], rows (1,1), cols (0,1))                                   (train.py:97-99)
//   conv   = lambda * VALID-correlate(imgp, filt[FH,FW,C]),  FH = H+2-D+1, FW = W+1-D+1  -> [D,D]   (train.py:103-109)
//   ngh    = conv @ stat_mask,  stat_mask[D,T] rows = range(0, 1, 1/T)                   -> [D,T]   (train.py:154-158)
// The reference draws the filter with an unseeded tf.random_normal, so the filter is an INPUT here (the host mirror
// draws it from a seeded generator).  Every output is a dot product over FH*FW*C ~ 1.2 M taps: 2*D*D*FH*FW*C FLOP
// (0.64 GFLOP at 576x720x3, D = 16) over 2*FH*FW*C*4 B read once -> fp32-FMA bound, not HBM bound.
//
// Kernel 1: one CTA per filter row, thread = one of the D*D outputs; the D image rows under the filter row and the
// filter row itself are staged through shared memory in tiles of SC_BCH filter columns (thread (i, j) reads word
// 3 (j + b) + c: stride-3 over j, conflict free).  Fixed summation order -> deterministic.  Kernel 2 adds the FH
// partial sums per output in double, scales by lambda and forms ngh.
#include "mmt_common.cuh"

namespace mmt {

constexpr int SC_BCH = 64;

struct ScArgs {
  const float *img, *filt;
  int H, W, C, FH, FW, D;
  float* part;   // [FH, D*D]
};

__global__ void __launch_bounds__(256) static_ctx_partial_kernel(ScArgs a) {
  extern __shared__ __align__(16) float sc_sm[];
  const int D = a.D, C = a.C, DD = D * D;
  const int colsmax = SC_BCH + D - 1;
  float* s_img = sc_sm;                       // [D][colsmax][C]
  float* s_f = sc_sm + D * colsmax * C;       // [SC_BCH][C]
  const int tid = threadIdx.x;
  const int oi = tid / D, oj = tid - oi * D;
  for (int fa = blockIdx.x; fa < a.FH; fa += gridDim.x) {
    float acc = 0.f;
    for (int b0 = 0; b0 < a.FW; b0 += SC_BCH) {
      const int nb = min(SC_BCH, a.FW - b0);
      const int cols = nb + D - 1;
      for (int idx = tid; idx < D * cols * C; idx += blockDim.x) {
        const int yy = idx / (cols * C), rem = idx - yy * (cols * C);
        const int xx = rem / C, c = rem - xx * C;
        const int iy = fa + yy - 1, ix = b0 + xx;     // padded -> image coordinates (one zero row on top, none on the left)
        s_img[(yy * colsmax + xx) * C + c] =
            (iy >= 0 && iy < a.H && ix < a.W) ? __ldg(a.img + ((size_t)iy * a.W + ix) * C + c) : 0.f;
      }
      for (int idx = tid; idx < nb * C; idx += blockDim.x) s_f[idx] = __ldg(a.filt + ((size_t)fa * a.FW + b0) * C + idx);
      __syncthreads();
      if (tid < DD) {
        const float* row = s_img + (oi * colsmax + oj) * C;
        for (int k = 0; k < nb * C; ++k) acc = fmaf(row[k], s_f[k], acc);   // (b, c) contiguous in both operands
      }
      __syncthreads();
    }
    if (tid < DD) a.part[(size_t)fa * DD + tid] = acc;
  }
}

__global__ void __launch_bounds__(256) static_ctx_reduce_kernel(const float* part, int FH, int D, int T, float lam,
                                                                float* conv, float* ngh) {
  __shared__ double s_conv[256];
  const int tid = threadIdx.x, DD = D * D;
  if (tid < DD) {
    double s = 0.0;
    for (int fa = 0; fa < FH; ++fa) s += (double)part[(size_t)fa * DD + tid];
    s *= (double)lam;
    s_conv[tid] = s;
    conv[tid] = (float)s;
  }
  __syncthreads();
  if (tid < D * T) {
    const int i = tid / T, t = tid - i * T;
    double rs = 0.0;
    for (int j = 0; j < D; ++j) rs += s_conv[i * D + j];
    ngh[tid] = (float)(rs * ((double)t * (1.0 / (double)T)));   // stat_mask[j, t] = t / T for every j
  }
}

}  // namespace mmt

extern "C" size_t mmt_static_context_workspace_bytes(int H, int D) {
  const int FH = H + 2 - D + 1;
  return FH > 0 && D > 0 ? (size_t)FH * D * D * sizeof(float) : 0;
}

extern "C" int mmt_static_context_f32(const float* img, int H, int W, int C, const float* filt, int D, int T, float lam,
                                      float* conv, float* ngh, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(img && filt && conv && ngh && workspace, "pointers required");
  MMT_REQUIRE(D > 0 && D <= 16 && T > 0 && D * T <= 256 && C > 0 && C <= 4, "need 0 < D <= 16, D*T <= 256, 0 < C <= 4");
  const int FH = H + 2 - D + 1, FW = W + 1 - D + 1;
  MMT_REQUIRE(H > 0 && W > 0 && FH > 0 && FW > 0, "image smaller than the neighbourhood grid");
  if (workspace_bytes < mmt_static_context_workspace_bytes(H, D)) {
    set_error("mmt_static_context_f32: workspace too small");
    return MMT_EWORKSPACE;
  }
  ScArgs a;
  a.img = img; a.filt = filt; a.H = H; a.W = W; a.C = C; a.FH = FH; a.FW = FW; a.D = D;
  a.part = static_cast<float*>(workspace);
  const size_t smem = sizeof(float) * ((size_t)D * (SC_BCH + D - 1) * C + (size_t)SC_BCH * C);
  const int grid = FH < num_sms() * 4 ? FH : num_sms() * 4;
  static_ctx_partial_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  int rc = check_launch("static_ctx_partial_kernel");
  if (rc) return rc;
  static_ctx_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a.part, FH, D, T, lam, conv, ngh);
  count_launch();
  return check_launch("static_ctx_reduce_kernel");
}
