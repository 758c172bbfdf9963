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
// c, mc: row stride ld_c (U for separate arrays, 2U for the c / mc halves of packed [h | c], [mh | mc] rows); bias: NULL (z holds
// it already) or b[3U] added here (packed training path: z comes straight out of the GEMM); hc_out: NULL or packed [h' | c']
// rows [R, 2U] written besides h_out / c_out (either of which may then be NULL)
__global__ void __launch_bounds__(256) gsk_gates_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                                                        const float* __restrict__ c, const float* __restrict__ mc, int ld_c,
                                                        const uint8_t* __restrict__ valid,
                                                        const float* __restrict__ w_If, const float* __restrict__ w_It,
                                                        const float* __restrict__ w_Of, const float* __restrict__ w_Ot,
                                                        int R, int U, float* __restrict__ h_out, float* __restrict__ c_out,
                                                        float* __restrict__ mf_out, float* __restrict__ hc_out) {
  const int U4 = U >> 2;
  const long total = (long)R * U4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / U4), u = (int)(i - (long)r * U4) * 4;
    const size_t o = (size_t)r * U + u, oz = (size_t)r * 3 * U + u, oc = (size_t)r * ld_c + u;
    float4 ho = make_float4(0.f, 0.f, 0.f, 0.f), co = ho, fo = ho;
    if (valid[r]) {
      float4 zi = *reinterpret_cast<const float4*>(z + oz), zj = *reinterpret_cast<const float4*>(z + oz + U),
             zo = *reinterpret_cast<const float4*>(z + oz + 2 * U);
      if (bias) {
        const float4 bi = *reinterpret_cast<const float4*>(bias + u), bj = *reinterpret_cast<const float4*>(bias + U + u),
                     bo = *reinterpret_cast<const float4*>(bias + 2 * U + u);
        zi = make_float4(zi.x + bi.x, zi.y + bi.y, zi.z + bi.z, zi.w + bi.w);
        zj = make_float4(zj.x + bj.x, zj.y + bj.y, zj.z + bj.z, zj.w + bj.w);
        zo = make_float4(zo.x + bo.x, zo.y + bo.y, zo.z + bo.z, zo.w + bo.w);
      }
      const float4 cp = *reinterpret_cast<const float4*>(c + oc), m = *reinterpret_cast<const float4*>(mc + oc);
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
    if (h_out) *reinterpret_cast<float4*>(h_out + o) = ho;
    if (c_out) *reinterpret_cast<float4*>(c_out + o) = co;
    *reinterpret_cast<float4*>(mf_out + o) = fo;
    if (hc_out) {
      *reinterpret_cast<float4*>(hc_out + (size_t)r * 2 * U + u) = ho;
      *reinterpret_cast<float4*>(hc_out + (size_t)r * 2 * U + U + u) = co;
    }
  }
}


// thread = unit u of a row; a block walks rows blockIdx.x, + gridDim.x, ... and keeps the four peephole partial
// sums of its unit in registers
// Packed training path: bias != NULL (z without bias), c / mc with row stride ld_c, d_head[R, 2U] (NULL or the head's gradient
// w.r.t. [m_t | m_f]: its first half is added to d_mt, its second half is d_mf), dmc written with row stride ld_dmc (2U: the
// second half of the packed [d mh | d mc] rows the adjoint aggregation reads).
__global__ void __launch_bounds__(128) gsk_cell_backward_kernel(
    const float* __restrict__ z, const float* __restrict__ bias, const float* __restrict__ c, const float* __restrict__ mc,
    int ld_c, const uint8_t* __restrict__ valid,
    const float* __restrict__ w_If, const float* __restrict__ w_It, const float* __restrict__ w_Of,
    const float* __restrict__ w_Ot, const float* __restrict__ d_mt, const float* __restrict__ d_mf,
    const float* __restrict__ d_head, const float* __restrict__ d_ct, int R, int U, float* __restrict__ dz,
    float* __restrict__ dc, float* __restrict__ dmc, int ld_dmc, float* __restrict__ dpeep, float* __restrict__ db) {
  const int u = threadIdx.x;
  if (u >= U) return;
  const float pIf = w_If[u], pIt = w_It[u], pOf = w_Of[u], pOt = w_Ot[u];
  const float bzi = bias ? bias[u] : 0.f, bzj = bias ? bias[U + u] : 0.f, bzo = bias ? bias[2 * U + u] : 0.f;
  float aIf = 0.f, aIt = 0.f, aOf = 0.f, aOt = 0.f;
  float bI = 0.f, bJ = 0.f, bO = 0.f;   // column sums of dz = the bias gradient (was a separate reduction pass over dz)
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const size_t o = (size_t)r * U + u, oz = (size_t)r * 3 * U + u;
    float di = 0.f, dj = 0.f, dO = 0.f, dcp = 0.f, dm = 0.f;
    if (valid[r]) {
      const float cp = c[(size_t)r * ld_c + u], m = mc[(size_t)r * ld_c + u];
      const float g = sigmoid_acc(z[oz] + bzi + pIf * m + pIt * cp);
      const float tj = tanhf(z[oz + U] + bzj);
      const float cf = (1.f - g) * m + g * tj, ct = (1.f - g) * cp + g * tj;
      const float q = sigmoid_acc(z[oz + 2 * U] + bzo + pOf * cf + pOt * ct);
      const float tcf = tanhf(cf), tct = tanhf(ct);
      float gmt = d_mt[o], gmf = d_mf ? d_mf[o] : 0.f;
      const float gct = d_ct ? d_ct[o] : 0.f;
      if (d_head) {
        gmt += d_head[(size_t)r * 2 * U + u];
        gmf += d_head[(size_t)r * 2 * U + U + u];
      }
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
    dmc[(size_t)r * ld_dmc + u] = dm;
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

// ---- element-wise glue of the packed training path (Trainer(gemm="tc")): what used to be library slicing / concatenation /
// broadcast kernels between this library's kernels (8 of the step's 21 ms, profiles/r02_train_launches.csv)

// teacher-forced inputs of frame t: cur = pos[:, t]; x = [cur - pos[:, t-1] | vis[:, min(t, T-1)]]; target = pos[:, t+1] - cur
__global__ void __launch_bounds__(256) train_frame_inputs_kernel(const float* __restrict__ pos, const float* __restrict__ vis,
                                                                 int R, int F, int T, int t, float* __restrict__ cur,
                                                                 float* __restrict__ x, float* __restrict__ target) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float2* pr = reinterpret_cast<const float2*>(pos) + (size_t)r * F;
  const float2 c = __ldg(pr + t);
  reinterpret_cast<float2*>(cur)[r] = c;
  float2 d = make_float2(0.f, 0.f);
  if (t > 0) {
    const float2 p = __ldg(pr + t - 1);
    d = make_float2(c.x - p.x, c.y - p.y);
  }
  const float2 v = __ldg(reinterpret_cast<const float2*>(vis) + (size_t)r * T + (t < T ? t : T - 1));
  reinterpret_cast<float4*>(x)[r] = make_float4(d.x, d.y, v.x, v.y);
  if (target && t + 1 < F) {
    const float2 n = __ldg(pr + t + 1);
    reinterpret_cast<float2*>(target)[r] = make_float2(n.x - c.x, n.y - c.y);
  }
}

// A[r] = [relu(x_r W_e + b_e) | hc[r, :U] | mhc[r, :U]]: the gate GEMM's input.  Thread = 4 consecutive columns of a row.
__global__ void __launch_bounds__(256) train_gate_input_kernel(const float* __restrict__ x, const float* __restrict__ hc,
                                                               const float* __restrict__ mhc, const float* __restrict__ W_e,
                                                               const float* __restrict__ b_e, int R, int E, int U,
                                                               float* __restrict__ A) {
  const int K4 = (E + 2 * U) >> 2;
  const long total = (long)R * K4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K4), k = (int)(i - (long)r * K4) * 4;
    float4 o;
    if (k < E) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + r);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W_e + k)), w1 = __ldg(reinterpret_cast<const float4*>(W_e + E + k)),
                   w2 = __ldg(reinterpret_cast<const float4*>(W_e + 2 * E + k)),
                   w3 = __ldg(reinterpret_cast<const float4*>(W_e + 3 * E + k)), bb = __ldg(reinterpret_cast<const float4*>(b_e + k));
      // x W_e + b_e in the accumulation order of the library GEMM it replaces does not matter at the 2e-2 tolerance of the mode
      o.x = fmaxf(fmaf(xv.w, w3.x, fmaf(xv.z, w2.x, fmaf(xv.y, w1.x, fmaf(xv.x, w0.x, bb.x)))), 0.f);
      o.y = fmaxf(fmaf(xv.w, w3.y, fmaf(xv.z, w2.y, fmaf(xv.y, w1.y, fmaf(xv.x, w0.y, bb.y)))), 0.f);
      o.z = fmaxf(fmaf(xv.w, w3.z, fmaf(xv.z, w2.z, fmaf(xv.y, w1.z, fmaf(xv.x, w0.z, bb.z)))), 0.f);
      o.w = fmaxf(fmaf(xv.w, w3.w, fmaf(xv.z, w2.w, fmaf(xv.y, w1.w, fmaf(xv.x, w0.w, bb.w)))), 0.f);
    } else if (k < E + U) {
      o = *reinterpret_cast<const float4*>(hc + (size_t)r * 2 * U + (k - E));
    } else {
      o = *reinterpret_cast<const float4*>(mhc + (size_t)r * 2 * U + (k - E - U));
    }
    *reinterpret_cast<float4*>(A + (size_t)r * (E + 2 * U) + k) = o;
  }
}

// after dA = dz W^T: dpre[r] = dA[r, :E] * (A[r, :E] > 0) (+ its column sums into gbe[E]); dmhc[r, :U] = dA[r, E+U:]
// (the d mh half of the packed rows whose d mc half the cell backward wrote).  Thread = 4 consecutive columns of [E + U];
// blockDim is a multiple of (E + U) / 4, so a thread keeps ITS column quad over the grid-stride loop and the bias
// gradient costs four atomicAdd per thread.
__global__ void train_backward_split_kernel(const float* __restrict__ dA, const float* __restrict__ A, int R, int E, int U,
                                            float* __restrict__ dpre, float* __restrict__ dmhc, float* __restrict__ gbe) {
  const int C4 = (E + U) >> 2, K = E + 2 * U;
  const int q = (int)(threadIdx.x % C4) * 4;
  const int rows_per_block = blockDim.x / C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = blockIdx.x * rows_per_block + threadIdx.x / C4; r < R; r += gridDim.x * rows_per_block) {
    if (q < E) {
      const float4 d = *reinterpret_cast<const float4*>(dA + (size_t)r * K + q), e = *reinterpret_cast<const float4*>(A + (size_t)r * K + q);
      const float4 o = make_float4(e.x > 0.f ? d.x : 0.f, e.y > 0.f ? d.y : 0.f, e.z > 0.f ? d.z : 0.f, e.w > 0.f ? d.w : 0.f);
      *reinterpret_cast<float4*>(dpre + (size_t)r * E + q) = o;
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    } else {
      *reinterpret_cast<float4*>(dmhc + (size_t)r * 2 * U + (q - E)) = *reinterpret_cast<const float4*>(dA + (size_t)r * K + U + q);
    }
  }
  // bias gradient: the block's threads of one column quad meet in shared memory, then ONE global atomic per column and block
  // (per-thread global atomics -- 380 k of them on 64 addresses -- made this kernel 136 us instead of ~25)
  if (gbe) {
    __shared__ float s_acc[256];
    for (int i = threadIdx.x; i < E; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    if (q < E) {
      atomicAdd(&s_acc[q], acc.x); atomicAdd(&s_acc[q + 1], acc.y); atomicAdd(&s_acc[q + 2], acc.z); atomicAdd(&s_acc[q + 3], acc.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) atomicAdd(gbe + i, s_acc[i]);
  }
}

// after back = att^T [d mh | d mc]: Gh = dA[:, E:E+U] + back[:, :U] (gradient w.r.t. the previous h), Gc = dc + back[:, U:]
__global__ void __launch_bounds__(256) train_backward_merge_kernel(const float* __restrict__ dA, const float* __restrict__ back,
                                                                   const float* __restrict__ dc, int R, int E, int U,
                                                                   float* __restrict__ Gh, float* __restrict__ Gc) {
  const int U4 = U >> 2, K = E + 2 * U;
  const long total = (long)R * U4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / U4), u = (int)(i - (long)r * U4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(dA + (size_t)r * K + E + u);
    const float4 bh = *reinterpret_cast<const float4*>(back + (size_t)r * 2 * U + u);
    const float4 bc = *reinterpret_cast<const float4*>(back + (size_t)r * 2 * U + U + u);
    const float4 d = *reinterpret_cast<const float4*>(dc + (size_t)r * U + u);
    *reinterpret_cast<float4*>(Gh + (size_t)r * U + u) = make_float4(a.x + bh.x, a.y + bh.y, a.z + bh.z, a.w + bh.w);
    *reinterpret_cast<float4*>(Gc + (size_t)r * U + u) = make_float4(d.x + bc.x, d.y + bc.y, d.z + bc.z, d.w + bc.w);
  }
}

static int ew_grid(long items) {
  const long blocks = (items + 255) / 256;
  return blocks < (long)num_sms() * 16 ? (int)(blocks > 0 ? blocks : 1) : num_sms() * 16;
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
  gsk_cell_backward_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z, nullptr, c, mc, U, valid, w_If, w_It, w_Of, w_Ot, d_mt, d_mf,
                                                                   nullptr, d_ct, R, U, dz, dc, dmc, U, dpeep, db);
  count_launch();
  return check_launch("gsk_cell_backward_kernel");
}

extern "C" int mmt_gsk_cell_backward_packed_f32(const float* z, const float* b, const float* hc, const float* mhc,
                                                const uint8_t* valid, const float* w_If, const float* w_It, const float* w_Of,
                                                const float* w_Ot, const float* d_mt, const float* d_head, const float* d_ct,
                                                int R, int U, float* dz, float* dc, float* dmhc, float* dpeep, float* db,
                                                void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && U > 0 && U <= 128, "need R >= 0, 0 < U <= 128");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(z && b && hc && mhc && valid && w_If && w_It && w_Of && w_Ot && d_mt && dz && dc && dmhc && dpeep,
              "z/b/hc/mhc/valid/peepholes/d_mt/outputs required");
  // blocks per SM: 3 / 6 / 16 / 32 / 64 measured 324 / 176 / 148 / 132 / 137 us at 65 536 rows (a thread walks rows serially)
  const int grid = R < num_sms() * 32 ? R : num_sms() * 32;
  gsk_cell_backward_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z, b, hc + U, mhc + U, 2 * U, valid, w_If, w_It, w_Of, w_Ot,
                                                                   d_mt, nullptr, d_head, d_ct, R, U, dz, dc, dmhc + U, 2 * U,
                                                                   dpeep, db);
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
  gsk_gates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, nullptr, c, mc, U, valid, w_If, w_It, w_Of, w_Ot, R, U, h_out, c_out,
                                                           mf_out, nullptr);
  count_launch();
  return check_launch("gsk_gates_kernel");
}

extern "C" int mmt_gsk_gates_packed_f32(const float* z, const float* b, const float* hc, const float* mhc, const uint8_t* valid,
                                        const float* w_If, const float* w_It, const float* w_Of, const float* w_Ot, int R, int U,
                                        float* hc_out, float* h_out, float* mf_out, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && U > 0 && U % 4 == 0, "need R >= 0, U % 4 == 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(z && b && hc && mhc && valid && w_If && w_It && w_Of && w_Ot && hc_out && h_out && mf_out, "all pointers required");
  MMT_ALIGNED(z); MMT_ALIGNED(b); MMT_ALIGNED(hc); MMT_ALIGNED(mhc); MMT_ALIGNED(hc_out); MMT_ALIGNED(h_out); MMT_ALIGNED(mf_out);
  MMT_ALIGNED(w_If); MMT_ALIGNED(w_It); MMT_ALIGNED(w_Of); MMT_ALIGNED(w_Ot);
  gsk_gates_kernel<<<ew_grid((long)R * (U / 4)), 256, 0, (cudaStream_t)stream>>>(z, b, hc + U, mhc + U, 2 * U, valid, w_If, w_It, w_Of,
                                                                                 w_Ot, R, U, h_out, nullptr, mf_out, hc_out);
  count_launch();
  return check_launch("gsk_gates_kernel");
}

extern "C" int mmt_train_frame_inputs_f32(const float* pos, const float* vis, int R, int F, int T, int t, float* cur, float* x,
                                          float* target, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && T >= 1 && F >= T && t >= 0 && t < F, "need 0 <= t < F, 1 <= T <= F");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(pos && vis && cur && x, "pos/vis/cur/x required");
  MMT_ALIGNED(x);
  train_frame_inputs_kernel<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pos, vis, R, F, T, t, cur, x, target);
  count_launch();
  return check_launch("train_frame_inputs_kernel");
}

extern "C" int mmt_train_gate_input_f32(const float* x, const float* hc, const float* mhc, const float* W_e, const float* b_e,
                                        int R, int E, int U, float* A, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && E > 0 && E % 4 == 0 && U > 0 && U % 4 == 0, "need E % 4 == 0, U % 4 == 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(x && hc && mhc && W_e && b_e && A, "all pointers required");
  MMT_ALIGNED(x); MMT_ALIGNED(hc); MMT_ALIGNED(mhc); MMT_ALIGNED(W_e); MMT_ALIGNED(b_e); MMT_ALIGNED(A);
  train_gate_input_kernel<<<ew_grid((long)R * ((E + 2 * U) / 4)), 256, 0, (cudaStream_t)stream>>>(x, hc, mhc, W_e, b_e, R, E, U, A);
  count_launch();
  return check_launch("train_gate_input_kernel");
}

extern "C" int mmt_train_backward_split_f32(const float* dA, const float* A, int R, int E, int U, float* dpre, float* dmhc,
                                            float* gbe, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && E > 0 && E % 4 == 0 && U > 0 && U % 4 == 0, "need E % 4 == 0, U % 4 == 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(dA && A && dpre && dmhc, "dA/A/dpre/dmhc required");
  MMT_ALIGNED(dA); MMT_ALIGNED(A); MMT_ALIGNED(dpre); MMT_ALIGNED(dmhc);
  const int C4 = (E + U) / 4;
  MMT_REQUIRE(C4 <= 256 && E <= 256, "need E + U <= 1024, E <= 256");
  const int threads = (256 / C4) * C4, rows_per_block = threads / C4;
  const long blocks = ((long)R + rows_per_block - 1) / rows_per_block;
  const int grid = blocks < (long)num_sms() * 8 ? (int)blocks : num_sms() * 8;
  train_backward_split_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(dA, A, R, E, U, dpre, dmhc, gbe);
  count_launch();
  return check_launch("train_backward_split_kernel");
}

extern "C" int mmt_train_backward_merge_f32(const float* dA, const float* back, const float* dc, int R, int E, int U, float* Gh,
                                            float* Gc, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(R >= 0 && E > 0 && E % 4 == 0 && U > 0 && U % 4 == 0, "need E % 4 == 0, U % 4 == 0");
  if (R == 0) return MMT_OK;
  MMT_REQUIRE(dA && back && dc && Gh && Gc, "all pointers required");
  MMT_ALIGNED(dA); MMT_ALIGNED(back); MMT_ALIGNED(dc); MMT_ALIGNED(Gh); MMT_ALIGNED(Gc);
  train_backward_merge_kernel<<<ew_grid((long)R * (U / 4)), 256, 0, (cudaStream_t)stream>>>(dA, back, dc, R, E, U, Gh, Gc);
  count_launch();
  return check_launch("train_backward_merge_kernel");
}
