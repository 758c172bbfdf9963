// Track A, batched: the reference's as-written g2k_lstm_mcr / g2k_lstm_mc scene-frame step and
// the GridLSTMCell gate equations exactly as helper.py instantiates them.
//
// These are tiny per-scene matrix chains (<= 16x16, 24x8, 16x128): one CTA per scene keeps every
// intermediate in shared memory; the only HBM traffic is the scene's inputs and outputs
// (~24 KB per scene-frame) so the kernel is HBM/latency-bound, not FLOP-bound.
#include "mmt_common.cuh"

namespace mmt {

constexpr int MAXD = 16, MAXT = 16, MAXN = 64, MAXH = 128, MAX2P = 32;

struct McrArgs {
  const float *X, *V, *C, *Hs, *vemb_prev;
  const float *W_i, *W_ii, *W_v, *b_v, *W_r, *W_c, *W_o;
  int S, n, D, T, P, H, variant;
  float lam;
  float *attn, *cost, *band, *Hs_out, *adj, *vemb_out;
};

__global__ void __launch_bounds__(128) mcr_step_kernel(McrArgs a) {
  __shared__ float sX[MAXT * MAXN];           // X[T,n]
  __shared__ float sXW[MAXT * MAXD];          // X W_i  [T,D]
  __shared__ float sOut[(MAXD + 2) * MAXD];   // outputs [D+2,D]
  __shared__ float sVrel[2 * MAXD];
  __shared__ float sNgh[MAXD * MAXT];         // ngh' [D,T]
  __shared__ float sEo[MAXT * MAXD];          // Eo [T,D]
  __shared__ float sM[MAXT * MAXD];           // Eo * (W_r vrel)
  __shared__ float sAttn[MAXD * MAXD];
  __shared__ float sCost[MAXT * MAXT];
  __shared__ float sWC[MAX2P * MAXT];         // W_c cost [2P,T]
  __shared__ float sA[MAXD * MAXD];           // a
  __shared__ float sH[MAXD * MAXH];           // softmax(Hs) then a@.
  __shared__ float sH2[MAXD * MAXH];
  __shared__ float sAdj[MAXD];

  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = a.n, D = a.D, T = a.T, P = a.P, H = a.H;
  for (int s = blockIdx.x; s < a.S; s += gridDim.x) {
    const float* X = a.X + (size_t)s * T * n;
    const float* V = a.V + (size_t)s * 2 * n;
    const float* C = a.C + (size_t)s * D * D;
    const float* Hs = a.Hs + (size_t)s * D * H;
    for (int i = tid; i < T * n; i += nt) sX[i] = X[i];
    for (int i = tid; i < D * H; i += nt) sH[i] = Hs[i];
    __syncthreads();
    // XW = X W_i [T,D];  vemb = V W_i [2,D] -> outputs rows D, D+1  (train.py:179,183)
    for (int i = tid; i < (T + 2) * D; i += nt) {
      const int r = i / D, d = i - r * D;
      float acc = 0.f;
      if (r < T) {
        for (int k = 0; k < n; ++k) acc = fmaf(sX[r * n + k], __ldg(a.W_i + k * D + d), acc);
        sXW[r * D + d] = acc;
      } else {
        for (int k = 0; k < n; ++k) acc = fmaf(__ldg(V + (r - T) * n + k), __ldg(a.W_i + k * D + d), acc);
        sOut[(D + r - T) * D + d] = acc;
      }
    }
    // ngh' = lam * ((lam C) stat_mask): ngh'[i,t] = lam * sum_j (lam C_ij) (t/T)   (train.py:110,154-158; mcr:102)
    for (int i = tid; i < D * T; i += nt) {
      const int r = i / T, t = i - r * T;
      const float m = (float)t / (float)T;
      float acc = 0.f;
      for (int j = 0; j < D; ++j) acc = fmaf(a.lam * __ldg(C + r * D + j), m, acc);
      sNgh[i] = a.lam * acc;
    }
    __syncthreads();
    // I = W_ii XW [D,D] -> outputs rows 0..D-1 (train.py:180); vrel = vemb_prev * vemb (train.py:194-195)
    for (int i = tid; i < D * D; i += nt) {
      const int r = i / D, d = i - r * D;
      float acc = 0.f;
      for (int k = 0; k < T; ++k) acc = fmaf(__ldg(a.W_ii + r * T + k), sXW[k * D + d], acc);
      sOut[i] = acc;
    }
    for (int i = tid; i < 2 * D; i += nt) {
      const float ve = sOut[D * D + i];
      const float pv = a.vemb_prev ? a.vemb_prev[(size_t)s * 2 * D + i] : ve;
      sVrel[i] = pv * ve;
      if (a.vemb_out) a.vemb_out[(size_t)s * 2 * D + i] = ve;
    }
    __syncthreads();
    // Eo = W_v outputs + b_v [T,D] (mcr:105);  M = Eo * (W_r vrel)
    for (int i = tid; i < T * D; i += nt) {
      const int t = i / D, d = i - t * D;
      float acc = 0.f;
      for (int k = 0; k < D + 2; ++k) acc = fmaf(__ldg(a.W_v + t * (D + 2) + k), sOut[k * D + d], acc);
      acc += __ldg(a.b_v + d);
      sEo[i] = acc;
      const float rr = __ldg(a.W_r + t * 2) * sVrel[d] + __ldg(a.W_r + t * 2 + 1) * sVrel[D + d];
      sM[i] = acc * rr;
    }
    __syncthreads();
    // attn = ngh' M [D,D] (mcr:105-106);  cost = Eo ngh' [T,T] (mcr:112-113; mc: 0)
    for (int i = tid; i < D * D + T * T; i += nt) {
      if (i < D * D) {
        const int r = i / D, d = i - r * D;
        float acc = 0.f;
        for (int k = 0; k < T; ++k) acc = fmaf(sNgh[r * T + k], sM[k * D + d], acc);
        sAttn[i] = acc;
        if (a.attn) a.attn[(size_t)s * D * D + i] = acc;
      } else {
        const int q = i - D * D;
        const int r = q / T, t = q - r * T;
        float acc = 0.f;
        if (a.variant == 0)
          for (int k = 0; k < D; ++k) acc = fmaf(sEo[r * D + k], sNgh[k * T + t], acc);
        sCost[q] = acc;
        if (a.cost) a.cost[(size_t)s * T * T + q] = acc;
      }
    }
    __syncthreads();
    // WC = W_c cost [2P,T] (mcr:122);  a = softmax_rows(exp(A)/cumsum0(exp(A))) (train.py:240)
    for (int i = tid; i < 2 * P * T; i += nt) {
      const int r = i / T, t = i - r * T;
      float acc = 0.f;
      for (int k = 0; k < T; ++k) acc = fmaf(__ldg(a.W_c + r * T + k), sCost[k * T + t], acc);
      sWC[i] = acc;
    }
    if (tid < D) {  // column cumsum of exp(A): thread per column
      float cs = 0.f;
      for (int r = 0; r < D; ++r) {
        const float e = expf(sAttn[r * D + tid]);
        cs += e;
        sA[r * D + tid] = e / cs;
      }
    }
    __syncthreads();
    if (tid < D) {  // row softmax of the ratio
      float mx = -INFINITY;
      for (int d = 0; d < D; ++d) mx = fmaxf(mx, sA[tid * D + d]);
      float sum = 0.f;
      for (int d = 0; d < D; ++d) {
        const float e = expf(sA[tid * D + d] - mx);
        sA[tid * D + d] = e;
        sum += e;
      }
      for (int d = 0; d < D; ++d) sA[tid * D + d] /= sum;
    }
    // band = WC W_o [2P,n] (mcr:122-124), written row-major == reshape (2,P,n)
    if (a.band)
      for (int i = tid; i < 2 * P * n; i += nt) {
        const int r = i / n, c = i - r * n;
        float acc = 0.f;
        for (int k = 0; k < T; ++k) acc = fmaf(sWC[r * T + k], __ldg(a.W_o + k * n + c), acc);
        a.band[(size_t)s * 2 * P * n + i] = acc;
      }
    // softmax_rows(Hs) (train.py:243): one warp per row
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      for (int r = warp; r < D; r += nw) {
        float mx = -INFINITY;
        for (int c = lane; c < H; c += 32) mx = fmaxf(mx, sH[r * H + c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < H; c += 32) {
          const float e = expf(sH[r * H + c] - mx);
          sH[r * H + c] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        for (int c = lane; c < H; c += 32) sH[r * H + c] /= sum;
      }
    }
    __syncthreads();
    // Hs = a softmax(Hs) (train.py:247)
    for (int i = tid; i < D * H; i += nt) {
      const int r = i / H, c = i - r * H;
      float acc = 0.f;
      for (int k = 0; k < D; ++k) acc = fmaf(sA[r * D + k], sH[k * H + c], acc);
      sH2[i] = acc;
    }
    __syncthreads();
    // adj = rowsum(softmax(Hs)) (train.py:248-249);  Hs = adj * Hs (train.py:252)
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      for (int r = warp; r < D; r += nw) {
        float mx = -INFINITY;
        for (int c = lane; c < H; c += 32) mx = fmaxf(mx, sH2[r * H + c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < H; c += 32) sum += expf(sH2[r * H + c] - mx);
        sum = warp_sum(sum);
        float tot = 0.f;
        for (int c = lane; c < H; c += 32) tot += expf(sH2[r * H + c] - mx) / sum;
        tot = warp_sum(tot);
        if (lane == 0) {
          sAdj[r] = tot;
          if (a.adj) a.adj[(size_t)s * D + r] = tot;
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < D * H; i += nt) a.Hs_out[(size_t)s * D * H + i] = sAdj[i / H] * sH2[i];
    __syncthreads();
  }
}

// GridLSTMCell as instantiated by helper.py:31-39 / :131-139.  One thread per batch row;
// the frequency blocks are sequential (block f consumes block f-1's m_freq / c_freq).
constexpr int GL_MAXU = 16;
__global__ void __launch_bounds__(128) gridlstm_step_kernel(const float* __restrict__ inputs, int in_stride,
                                                            const float* __restrict__ state, int st_stride,
                                                            const float* __restrict__ W_f, const float* __restrict__ B_f,
                                                            const float* __restrict__ w_If, const float* __restrict__ w_It,
                                                            const float* __restrict__ w_Of, const float* __restrict__ w_Ot,
                                                            int B, int U, int F, int peep, float* __restrict__ m_out,
                                                            float* __restrict__ state_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  float m_f[GL_MAXU], c_f[GL_MAXU];
  for (int u = 0; u < U; ++u) m_f[u] = c_f[u] = 0.f;
  const int G3 = 3 * U;
  for (int f = 0; f < F; ++f) {
    const float* x = inputs + (size_t)r * in_stride + 4 * f;
    const float* ct = state + (size_t)r * st_stride + 2 * f * U;
    const float* mt = ct + U;
    float nm_f[GL_MAXU], nc_f[GL_MAXU];
    for (int u = 0; u < U; ++u) {
      float z[3];
      for (int g = 0; g < 3; ++g) {
        const int col = g * U + u;
        float acc = 0.f;
        for (int k = 0; k < 4; ++k) acc = fmaf(x[k], __ldg(W_f + k * G3 + col), acc);
        for (int k = 0; k < U; ++k) acc = fmaf(mt[k], __ldg(W_f + (4 + k) * G3 + col), acc);
        for (int k = 0; k < U; ++k) acc = fmaf(m_f[k], __ldg(W_f + (4 + U + k) * G3 + col), acc);
        z[g] = acc + __ldg(B_f + col);
      }
      const float cto = ct[u];
      const float g_ = peep ? sigmoid_acc(z[0] + w_If[u] * c_f[u] + w_It[u] * cto) : sigmoid_acc(z[0]);
      const float tj = tanhf(z[1]);
      const float cfreq = (1.f - g_) * c_f[u] + g_ * tj;
      const float ctime = (1.f - g_) * cto + g_ * tj;
      const float q = peep ? sigmoid_acc(z[2] + w_Of[u] * cfreq + w_Ot[u] * ctime) : sigmoid_acc(z[2]);
      const float mfreq = q * tanhf(cfreq), mtime = q * tanhf(ctime);
      float* so = state_out + (size_t)r * 2 * U * F + 2 * f * U;
      so[u] = ctime;
      so[U + u] = mtime;
      float* mo = m_out + (size_t)r * 2 * U * F + 2 * f * U;
      mo[u] = mtime;
      mo[U + u] = mfreq;
      nm_f[u] = mfreq;
      nc_f[u] = cfreq;
    }
    for (int u = 0; u < U; ++u) {
      m_f[u] = nm_f[u];
      c_f[u] = nc_f[u];
    }
  }
}

}  // namespace mmt

extern "C" int mmt_mcr_step_f32(const float* X, const float* V, const float* C, const float* Hs,
                                const float* vemb_prev, const mmt_mcr_weights* w, int S, int n, int D, int T, int P,
                                int H, float lam, int variant, float* attn, float* cost, float* band, float* Hs_out,
                                float* adj, float* vemb_out, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(X && V && C && Hs && w && Hs_out, "X/V/C/Hs/w/Hs_out must not be NULL");
  MMT_REQUIRE(w->W_i && w->W_ii && w->W_v && w->b_v && w->W_r && w->W_c && w->W_o, "all weights required");
  MMT_REQUIRE(S >= 0 && n > 0 && n <= MAXN && D > 0 && D <= MAXD && T > 0 && T <= MAXT && H > 0 && H <= MAXH &&
                  P > 0 && 2 * P <= MAX2P,
              "need n <= 64, D <= 16, T <= 16, H <= 128, 2P <= 32");
  MMT_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (mcr) or 1 (mc)");
  MMT_REQUIRE(Hs_out != Hs, "Hs_out must not alias Hs");
  if (S == 0) return MMT_OK;
  McrArgs a;
  a.X = X; a.V = V; a.C = C; a.Hs = Hs; a.vemb_prev = vemb_prev;
  a.W_i = w->W_i; a.W_ii = w->W_ii; a.W_v = w->W_v; a.b_v = w->b_v; a.W_r = w->W_r; a.W_c = w->W_c; a.W_o = w->W_o;
  a.S = S; a.n = n; a.D = D; a.T = T; a.P = P; a.H = H; a.variant = variant; a.lam = lam;
  a.attn = attn; a.cost = cost; a.band = band; a.Hs_out = Hs_out; a.adj = adj; a.vemb_out = vemb_out;
  int grid = S < num_sms() * 8 ? S : num_sms() * 8;
  mcr_step_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return check_launch("mcr_step_kernel");
}

extern "C" int mmt_gridlstm_step_f32(const float* inputs, int in_stride, const float* state, int st_stride,
                                     const float* W_f, const float* B_f, const float* w_If, const float* w_It,
                                     const float* w_Of, const float* w_Ot, int B, int U, int F, int peepholes,
                                     float* m_out, float* state_out, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(inputs && state && W_f && B_f && m_out && state_out, "inputs/state/W_f/B_f/outputs must not be NULL");
  MMT_REQUIRE(!peepholes || (w_If && w_It && w_Of && w_Ot), "peephole diagonals required when peepholes != 0");
  MMT_REQUIRE(B >= 0 && U > 0 && U <= GL_MAXU && F > 0, "need 0 < U <= 16, F > 0");
  MMT_REQUIRE(in_stride >= 4 * F && st_stride >= 2 * U * F, "strides too small for F blocks");
  MMT_REQUIRE(state_out != state, "state_out must not alias state");
  if (B == 0) return MMT_OK;
  gridlstm_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(inputs, in_stride, state, st_stride, W_f, B_f,
                                                                          w_If, w_It, w_Of, w_Ot, B, U, F, peepholes,
                                                                          m_out, state_out);
  count_launch();
  return check_launch("gridlstm_step_kernel");
}
