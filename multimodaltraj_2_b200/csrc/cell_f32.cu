// gsk_lstm_cell, fp32 parity mode: tiled CUDA-core SGEMM [R, E+2U] x [E+2U, 3U] with the gate
// update fused as the epilogue (include/mmt.h mmt_gsk_cell).
//
// The A operand [e | h | mh] is never materialised: e = relu(x W_e + b_e) is computed while
// the tile is staged.  Each thread owns (4 rows) x (2 units) x (3 gates) so i, j, o of a unit
// meet in one thread and the cell update needs no exchange.  This is the 1e-4-parity path;
// the throughput path is the tcgen05 kernel in cell_tc.cu.
#include "mmt_common.cuh"

namespace mmt {

constexpr int CBM = 64;   // rows per CTA
constexpr int CBU = 32;   // units per CTA (=> 96 gate columns)
constexpr int CBK = 16;   // k-chunk

struct CellArgs {
  const float *x, *h, *c, *mh, *mc;
  const uint8_t* valid;
  const float *W_e, *b_e, *W, *b, *w_If, *w_It, *w_Of, *w_Ot;
  float *h_out, *c_out, *mf_out;
  int R, E, U, ld, ld_mf;  // ld = row stride (floats) of h/c/mh/mc/h_out/c_out; ld_mf of mf_out
};

__global__ void __launch_bounds__(256) gsk_cell_f32_kernel(CellArgs a) {
  __shared__ __align__(16) float As[2][CBK][CBM + 4];
  __shared__ __align__(16) float Bs[2][CBK][3 * CBU];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int row0 = blockIdx.x * CBM;
  const int u0 = blockIdx.y * CBU;
  const int E = a.E, U = a.U, Kt = E + 2 * U, G3 = 3 * U;
  const int nchunks = Kt / CBK;

  // A staging: thread -> (row lr = tid/4, k offset lk = (tid%4)*4), one float4 along k
  const int lr = tid >> 2, lk = (tid & 3) << 2;
  const int grow = row0 + lr;
  const bool rok = grow < a.R;
  float4 xrow = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rok) xrow = __ldg(reinterpret_cast<const float4*>(a.x) + grow);
  // B staging: thread -> (k = tid/16, cols (tid%16)*2 of each gate segment)
  const int bk = tid >> 4, bc = (tid & 15) << 1;

  float acc[4][3][2];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int g = 0; g < 3; ++g) acc[r][g][0] = acc[r][g][1] = 0.f;

  float4 areg;
  float2 breg[3];
  auto load_chunk = [&](int ch) {
    const int k0 = ch * CBK;
    const int k = k0 + lk;
    areg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rok) {
      if (k < E) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kk = k + q;
          float s = __ldg(a.b_e + kk);
          s = fmaf(xrow.x, __ldg(a.W_e + 0 * E + kk), s);
          s = fmaf(xrow.y, __ldg(a.W_e + 1 * E + kk), s);
          s = fmaf(xrow.z, __ldg(a.W_e + 2 * E + kk), s);
          s = fmaf(xrow.w, __ldg(a.W_e + 3 * E + kk), s);
          v[q] = fmaxf(s, 0.f);
        }
        areg = make_float4(v[0], v[1], v[2], v[3]);
      } else if (k < E + U) {
        areg = *reinterpret_cast<const float4*>(a.h + (size_t)grow * a.ld + (k - E));
      } else {
        areg = *reinterpret_cast<const float4*>(a.mh + (size_t)grow * a.ld + (k - E - U));
      }
    }
    const float* wrow = a.W + (size_t)(k0 + bk) * G3 + u0 + bc;
#pragma unroll
    for (int g = 0; g < 3; ++g) breg[g] = __ldg(reinterpret_cast<const float2*>(wrow + g * U));
  };
  auto store_chunk = [&](int buf) {
    As[buf][lk + 0][lr] = areg.x;
    As[buf][lk + 1][lr] = areg.y;
    As[buf][lk + 2][lr] = areg.z;
    As[buf][lk + 3][lr] = areg.w;
#pragma unroll
    for (int g = 0; g < 3; ++g) *reinterpret_cast<float2*>(&Bs[buf][bk][g * CBU + bc]) = breg[g];
  };

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) load_chunk(ch + 1);
#pragma unroll
    for (int k = 0; k < CBK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      float2 bv[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) bv[g] = *reinterpret_cast<const float2*>(&Bs[buf][k][g * CBU + tx * 2]);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          acc[r][g][0] = fmaf(ar[r], bv[g].x, acc[r][g][0]);
          acc[r][g][1] = fmaf(ar[r], bv[g].y, acc[r][g][1]);
        }
    }
    if (ch + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- fused gate epilogue (App. B equations)
  const int u = u0 + tx * 2;
  const float2 bi = *reinterpret_cast<const float2*>(a.b + u);
  const float2 bj = *reinterpret_cast<const float2*>(a.b + U + u);
  const float2 bo = *reinterpret_cast<const float2*>(a.b + 2 * U + u);
  const float2 pIf = *reinterpret_cast<const float2*>(a.w_If + u);
  const float2 pIt = *reinterpret_cast<const float2*>(a.w_It + u);
  const float2 pOf = *reinterpret_cast<const float2*>(a.w_Of + u);
  const float2 pOt = *reinterpret_cast<const float2*>(a.w_Ot + u);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + ty * 4 + r;
    if (row >= a.R) continue;
    const size_t off = (size_t)row * a.ld + u;
    float2 ho = make_float2(0.f, 0.f), co = ho, fo = ho;
    if (a.valid[row]) {
      const float2 cv = *reinterpret_cast<const float2*>(a.c + off);
      const float2 mcv = *reinterpret_cast<const float2*>(a.mc + off);
      const float cc[2] = {cv.x, cv.y}, mm[2] = {mcv.x, mcv.y};
      const float bI[2] = {bi.x, bi.y}, bJ[2] = {bj.x, bj.y}, bO[2] = {bo.x, bo.y};
      const float wIf[2] = {pIf.x, pIf.y}, wIt[2] = {pIt.x, pIt.y}, wOf[2] = {pOf.x, pOf.y}, wOt[2] = {pOt.x, pOt.y};
      float hh[2], c2[2], ff[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float zi = acc[r][0][q] + bI[q], zj = acc[r][1][q] + bJ[q], zo = acc[r][2][q] + bO[q];
        const float g = sigmoid_acc(zi + wIf[q] * mm[q] + wIt[q] * cc[q]);
        const float tj = tanhf(zj);
        const float cf = (1.0f - g) * mm[q] + g * tj;
        const float ct = (1.0f - g) * cc[q] + g * tj;
        const float qq = sigmoid_acc(zo + wOf[q] * cf + wOt[q] * ct);
        ff[q] = qq * tanhf(cf);
        hh[q] = qq * tanhf(ct);
        c2[q] = ct;
      }
      ho = make_float2(hh[0], hh[1]);
      co = make_float2(c2[0], c2[1]);
      fo = make_float2(ff[0], ff[1]);
    }
    *reinterpret_cast<float2*>(a.h_out + off) = ho;
    *reinterpret_cast<float2*>(a.c_out + off) = co;
    *reinterpret_cast<float2*>(a.mf_out + (size_t)row * a.ld_mf + u) = fo;
  }
}

// head: y = [m_t | m_f] W_h + b_h -> (mu_x, mu_y, exp, exp, tanh); next_pos = cur + mu.  One warp per row.
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ mt, int ld,
                                                   const float* __restrict__ mf, int ld_mf,
                                                   const uint8_t* __restrict__ valid, const float* __restrict__ W_h,
                                                   const float* __restrict__ b_h, int R, int U,
                                                   const float* __restrict__ cur_pos, float* __restrict__ params_out,
                                                   int params_stride, float* __restrict__ next_pos) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    float y[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const bool v = valid[r] != 0;
    if (v) {
      for (int u = lane; u < U; u += 32) {
        const float a = mt[(size_t)r * ld + u], b = mf[(size_t)r * ld_mf + u];
        const float* wa = W_h + (size_t)u * 5;
        const float* wb = W_h + (size_t)(U + u) * 5;
#pragma unroll
        for (int q = 0; q < 5; ++q) y[q] = fmaf(a, __ldg(wa + q), fmaf(b, __ldg(wb + q), y[q]));
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) y[q] = warp_sum(y[q]);
    if (lane == 0) {
      float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (v) {
        o[0] = y[0] + b_h[0];
        o[1] = y[1] + b_h[1];
        o[2] = expf(y[2] + b_h[2]);
        o[3] = expf(y[3] + b_h[3]);
        o[4] = tanhf(y[4] + b_h[4]);
      }
      float* po = params_out + (size_t)r * params_stride;
#pragma unroll
      for (int q = 0; q < 5; ++q) po[q] = o[q];
      if (next_pos != nullptr) {
        const float2 cp = *reinterpret_cast<const float2*>(cur_pos + (size_t)r * 2);
        *reinterpret_cast<float2*>(next_pos + (size_t)r * 2) = make_float2(cp.x + o[0], cp.y + o[1]);
      }
    }
  }
}

int launch_cell_f32(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                    const uint8_t* valid, const mmt_cell_weights* w, int R, int ld, float* h_out, float* c_out,
                    float* mf_out, int ld_mf, cudaStream_t stream) {
  CellArgs a;
  a.x = x; a.h = h; a.c = c; a.mh = mh; a.mc = mc; a.valid = valid;
  a.W_e = w->W_e; a.b_e = w->b_e; a.W = w->W; a.b = w->b;
  a.w_If = w->w_If; a.w_It = w->w_It; a.w_Of = w->w_Of; a.w_Ot = w->w_Ot;
  a.h_out = h_out; a.c_out = c_out; a.mf_out = mf_out;
  a.R = R; a.E = w->E; a.U = w->U; a.ld = ld; a.ld_mf = ld_mf;
  dim3 grid((R + CBM - 1) / CBM, w->U / CBU);
  gsk_cell_f32_kernel<<<grid, 256, 0, stream>>>(a);
  count_launch();
  return check_launch("gsk_cell_f32_kernel");
}

int launch_head(const float* mt, int ld, const float* mf, int ld_mf, const uint8_t* valid, const mmt_cell_weights* w, int R,
                const float* cur_pos, float* params_out, int params_stride, float* next_pos, cudaStream_t stream) {
  long blocks = ((long)R + 7) / 8;
  int grid = blocks < (long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  head_kernel<<<grid, 256, 0, stream>>>(mt, ld, mf, ld_mf, valid, w->W_h, w->b_h, R, w->U, cur_pos, params_out,
                                        params_stride, next_pos);
  count_launch();
  return check_launch("head_kernel");
}

}  // namespace mmt
