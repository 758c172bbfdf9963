// Training step of the Track-B path (SURVEY App. C.5 "Training loss (fills F2)"; section 8e: data-parallel training
// with one gradient all-reduce): the two fused backward kernels.  The loss is the teacher-forced mean bivariate-
// Gaussian NLL (oracle/train_b.py); back-propagation through time is orchestrated by multimodaltraj_2_b200/train.py
// (Trainer): these kernels do the element-wise work of a step, the weight-gradient / input-gradient contractions
// are plain library GEMMs.
//
//   mmt_head_nll_f32            y = [m_t | m_f] W_h + b_h ; nll(y, target) ; dy = d nll / d y          (warp per row)
//   mmt_gsk_gates_f32           forward gate update from pre-activations z = [e|h|mh] W + b computed by a library GEMM:
//                               h', c', m_f (the training forward keeps z for the backward instead of recomputing it)
//   mmt_gsk_cell_backward_f32   gates re-evaluated from the saved pre-activations z, then d z, d c, d mc and the
//                               peephole gradients (block-reduced, one atomicAdd per unit and block)
#include "mmt_common.cuh"

namespace mmt {

__global__ void __launch_bounds__(256) head_nll_kernel(const float* __restrict__ m_t, const float* __restrict__ m_f,
                                                       const uint8_t* __restrict__ valid, const float* __restrict__ W_h,
                                                       const float* __restrict__ b_h, const float* __restrict__ target,
                                                       int R, int U, float scale, float* __restrict__ loss_sum,
                                                       float* __restrict__ dy) {
  __shared__ float s_loss[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wloss = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
    float y[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < 2 * U; k += 32) {
      const float v = k < U ? m_t[(size_t)r * U + k] : m_f[(size_t)r * U + (k - U)];
#pragma unroll
      for (int z = 0; z < 5; ++z) y[z] = fmaf(v, __ldg(W_h + (size_t)k * 5 + z), y[z]);
    }
#pragma unroll
    for (int z = 0; z < 5; ++z) y[z] = warp_sum(y[z]) + __ldg(b_h + z);
    if (lane == 0) {
      float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid[r]) {
        const float sx = expf(y[2]), sy = expf(y[3]), rho = tanhf(y[4]);
        const float zx = (target[(size_t)r * 2] - y[0]) / sx, zy = (target[(size_t)r * 2 + 1] - y[1]) / sy;
        const float om = 1.f - rho * rho, q = zx * zx - 2.f * rho * zx * zy + zy * zy;
        wloss += (1.8378770664093453f + y[2] + y[3] + 0.5f * logf(om) + q / (2.f * om)) * scale;
        g[0] = -(zx - rho * zy) / (om * sx) * scale;
        g[1] = -(zy - rho * zx) / (om * sy) * scale;
        g[2] = (1.f - (zx * zx - rho * zx * zy) / om) * scale;
        g[3] = (1.f - (zy * zy - rho * zx * zy) / om) * scale;
        g[4] = (-rho + (-zx * zy * om + rho * q) / om) * scale;
      }
#pragma unroll
      for (int z = 0; z < 5; ++z) dy[(size_t)r * 5 + z] = g[z];
    }
  }
  if (lane == 0) s_loss[warp] = wloss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    if (t != 0.f) atomicAdd(loss_sum, t);
  }
}

// forward gate update of helper.py:31-39 from the pre-activations: thread = 4 consecutive units of a row
__global__ void __launch_bounds__(256) gsk_gates_kernel(const float* __restrict__ z, const float* __restrict__ c,
                                                        const float* __restrict__ mc, const uint8_t* __restrict__ valid,
                                                        const float* __restrict__ w_If, const float* __restrict__ w_It,
                                                        const float* __restrict__ w_Of, const float* __restrict__ w_Ot,
                                                        int R, int U, float* __restrict__ h_out, float* __restrict__ c_out,
                                                        float* __restrict__ mf_out) {
  const int U4 = U >> 2;
  const long total = (long)R * U4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / U4), u = (int)(i - (long)r * U4) * 4;
    const size_t o = (size_t)r * U + u, oz = (size_t)r * 3 * U + u;
    float4 ho = make_float4(0.f, 0.f, 0.f, 0.f), co = ho, fo = ho;
    if (valid[r]) {
      const float4 zi = *reinterpret_cast<const float4*>(z + oz), zj = *reinterpret_cast<const float4*>(z + oz + U),
                   zo = *reinterpret_cast<const float4*>(z + oz + 2 * U);
      const float4 cp = *reinterpret_cast<const float4*>(c + o), m = *reinterpret_cast<const float4*>(mc + o);
      const float4 pIf = *reinterpret_cast<const float4*>(w_If + u), pIt = *reinterpret_cast<const float4*>(w_It + u),
                   pOf = *reinterpret_cast<const float4*>(w_Of + u), pOt = *reinterpret_cast<const float4*>(w_Ot + u);
      auto one = [](float zi_, float zj_, float zo_, float cp_, float m_, float pIf_, float pIt_, float pOf_, float pOt_,
                    float& h_, float& c_, float& f_) {
        const float g = sigmoid_acc(zi_ + pIf_ * m_ + pIt_ * cp_);
        const float tj = tanhf(zj_);
        const float cf = (1.f - g) * m_ + g * tj, ct = (1.f - g) * cp_ + g * tj;
        const float q = sigmoid_acc(zo_ + pOf_ * cf + pOt_ * ct);
        h_ = q * tanhf(ct);
        c_ = ct;
        f_ = q * tanhf(cf);
      };
      one(zi.x, zj.x, zo.x, cp.x, m.x, pIf.x, pIt.x, pOf.x, pOt.x, ho.x, co.x, fo.x);
      one(zi.y, zj.y, zo.y, cp.y, m.y, pIf.y, pIt.y, pOf.y, pOt.y, ho.y, co.y, fo.y);
      one(zi.z, zj.z, zo.z, cp.z, m.z, pIf.z, pIt.z, pOf.z, pOt.z, ho.z, co.z, fo.z);
      one(zi.w, zj.w, zo.w, cp.w, m.w, pIf.w, pIt.w, pOf.w, pOt.w, ho.w, co.w, fo.w);
    }
    *reinterpret_cast<float4*>(h_out + o) = ho;
    *reinterpret_cast<float4*>(c_out + o) = co;
    *reinterpret_cast<float4*>(mf_out + o) = fo;
  }
}


// thread = unit u of a row; a block walks rows blockIdx.x, + gridDim.x, ... and keeps the four peephole partial
// sums of its unit in registers
__global__ void __launch_bounds__(128) gsk_cell_backward_kernel(
    const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ mc, const uint8_t* __restrict__ valid,
    const float* __restrict__ w_If, const float* __restrict__ w_It, const float* __restrict__ w_Of,
    const float* __restrict__ w_Ot, const float* __restrict__ d_mt, const float* __restrict__ d_mf,
    const float* __restrict__ d_ct, int R, int U, float* __restrict__ dz, float* __restrict__ dc,
    float* __restrict__ dmc, float* __restrict__ dpeep, float* __restrict__ db) {
  const int u = threadIdx.x;
  if (u >= U) return;
  const float pIf = w_If[u], pIt = w_It[u], pOf = w_Of[u], pOt = w_Ot[u];
  float aIf = 0.f, aIt = 0.f, aOf = 0.f, aOt = 0.f;
  float bI = 0.f, bJ = 0.f, bO = 0.f;   // column sums of dz = the bias gradient (was a separate reduction pass over dz)
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const size_t o = (size_t)r * U + u, oz = (size_t)r * 3 * U + u;
    float di = 0.f, dj = 0.f, dO = 0.f, dcp = 0.f, dm = 0.f;
    if (valid[r]) {
      const float cp = c[o], m = mc[o];
      const float g = sigmoid_acc(z[oz] + pIf * m + pIt * cp);
      const float tj = tanhf(z[oz + U]);
      const float cf = (1.f - g) * m + g * tj, ct = (1.f - g) * cp + g * tj;
      const float q = sigmoid_acc(z[oz + 2 * U] + pOf * cf + pOt * ct);
      const float tcf = tanhf(cf), tct = tanhf(ct);
      const float gmt = d_mt[o], gmf = d_mf ? d_mf[o] : 0.f, gct = d_ct ? d_ct[o] : 0.f;
      const float dq = gmt * tct + gmf * tcf;
      float dct = gmt * q * (1.f - tct * tct) + gct;
      float dcf = gmf * q * (1.f - tcf * tcf);
      const float dpo = dq * q * (1.f - q);
      dcf += dpo * pOf;
      dct += dpo * pOt;
      aOf += dpo * cf;
      aOt += dpo * ct;
      const float dg = dcf * (tj - m) + dct * (tj - cp);
      dm = dcf * (1.f - g);
      dcp = dct * (1.f - g);
      dj = (dcf + dct) * g * (1.f - tj * tj);
      const float dpi = dg * g * (1.f - g);
      dm += dpi * pIf;
      dcp += dpi * pIt;
      aIf += dpi * m;
      aIt += dpi * cp;
      di = dpi;
      dO = dpo;
    }
    dz[oz] = di;
    dz[oz + U] = dj;
    dz[oz + 2 * U] = dO;
    bI += di;
    bJ += dj;
    bO += dO;
    dc[o] = dcp;
    dmc[o] = dm;
  }
  atomicAdd(dpeep + u, aIf);
  atomicAdd(dpeep + U + u, aIt);
  atomicAdd(dpeep + 2 * U + u, aOf);
  atomicAdd(dpeep + 3 * U + u, aOt);
  if (db) {
    atomicAdd(db + u, bI);
    atomicAdd(db + U + u, bJ);
    atomicAdd(db + 2 * U + u, bO);
  }
}

}  // namespace mmt

extern "C" int mmt_head_nll_f32(const float* m_t, const float* m_f, const uint8_t* valid, const float* W_h,
                                const float* b_h, const float* target, int R, int U, float scale, float* loss_sum,
                                float* dy, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && U > 0, "need R >= 0, U > 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(m_t && m_f && valid && W_h && b_h && target && loss_sum && dy, "all pointers required");
  const int grid = (R + 7) / 8 < num_sms() * 8 ? (R + 7) / 8 : num_sms() * 8;
  head_nll_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m_t, m_f, valid, W_h, b_h, target, R, U, scale, loss_sum, dy);
  count_launch();
  return check_launch("head_nll_kernel");
}

extern "C" int mmt_gsk_cell_backward_f32(const float* z, const float* c, const float* mc, const uint8_t* valid,
                                         const float* w_If, const float* w_It, const float* w_Of, const float* w_Ot,
                                         const float* d_mt, const float* d_mf, const float* d_ct, int R, int U,
                                         float* dz, float* dc, float* dmc, float* dpeep, float* db, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && U > 0 && U <= 128, "need R >= 0, 0 < U <= 128");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(z && c && mc && valid && w_If && w_It && w_Of && w_Ot && d_mt && dz && dc && dmc && dpeep,
              "z/c/mc/valid/peepholes/d_mt/outputs required");
  const int grid = R < num_sms() * 16 ? R : num_sms() * 16;
  gsk_cell_backward_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z, c, mc, valid, w_If, w_It, w_Of, w_Ot, d_mt, d_mf,
                                                                   d_ct, R, U, dz, dc, dmc, dpeep, db);
  count_launch();
  return check_launch("gsk_cell_backward_kernel");
}

extern "C" int mmt_gsk_gates_f32(const float* z, const float* c, const float* mc, const uint8_t* valid, const float* w_If,
                                 const float* w_It, const float* w_Of, const float* w_Ot, int R, int U, float* h_out,
                                 float* c_out, float* mf_out, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && U > 0 && U % 4 == 0, "need R >= 0, U % 4 == 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(z && c && mc && valid && w_If && w_It && w_Of && w_Ot && h_out && c_out && mf_out, "all pointers required");
  MMT_ALIGNED(z); MMT_ALIGNED(c); MMT_ALIGNED(mc); MMT_ALIGNED(h_out); MMT_ALIGNED(c_out); MMT_ALIGNED(mf_out);
  MMT_ALIGNED(w_If); MMT_ALIGNED(w_It); MMT_ALIGNED(w_Of); MMT_ALIGNED(w_Ot);
  const long blocks = ((long)R * (U / 4) + 255) / 256;
  const int grid = blocks < (long)num_sms() * 16 ? (int)blocks : num_sms() * 16;
  gsk_gates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, c, mc, valid, w_If, w_It, w_Of, w_Ot, R, U, h_out, c_out, mf_out);
  count_launch();
  return check_launch("gsk_gates_kernel");
}
