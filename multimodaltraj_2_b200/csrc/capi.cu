// C-ABI plumbing of libmmt: version, thread-local error string, launch counter, and the
// mmt_gsk_cell entry point that dispatches between the fp32 and the tcgen05/bf16 cell kernels.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "mmt_common.cuh"

namespace mmt {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MMT_ECUDA;
  }
  return MMT_OK;
}

int launch_cell_f32(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                    const uint8_t* valid, const mmt_cell_weights* w, int R, int ld, float* h_out, float* c_out,
                    float* mf_out, int ld_mf, cudaStream_t stream);
int launch_head(const float* mt, int ld, const float* mf, int ld_mf, const uint8_t* valid, const mmt_cell_weights* w,
                int R, const float* cur_pos, float* params_out, int params_stride, float* next_pos,
                cudaStream_t stream);
int launch_cell_tc(const float* x, const float* h, const float* c, const float* mh, const float* mc, int ld,
                   const uint8_t* valid, const mmt_cell_weights* w, int R, float* h_out, float* c_out, float* mf_out,
                   int ld_mf, const float* cur_pos, float* params_out, int params_stride, float* next_pos,
                   cudaStream_t stream);

}  // namespace mmt

extern "C" int mmt_version(void) { return MMT_VERSION; }
extern "C" const char* mmt_last_error(void) { return mmt::g_err; }
extern "C" uint64_t mmt_launch_count(void) { return mmt::g_launches.load(std::memory_order_relaxed); }

extern "C" int mmt_gsk_cell(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                            const uint8_t* valid, const mmt_cell_weights* w, int R, int prec, float* h_out,
                            float* c_out, float* mf_out, const float* cur_pos, float* params_out, int params_stride,
                            float* next_pos, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(x && h && c && mh && mc && valid && w && h_out && c_out && mf_out, "state/weight/output pointers required");
  MMT_REQUIRE(w->W_e && w->b_e && w->W && w->b && w->w_If && w->w_It && w->w_Of && w->w_Ot, "cell weights required");
  MMT_REQUIRE(R >= 0, "R must be >= 0");
  MMT_REQUIRE(w->E == 64 && w->U == 128, "cell is built for E = 64, U = 128");
  MMT_REQUIRE(prec == MMT_PREC_F32 || prec == MMT_PREC_BF16, "unknown precision mode");
  MMT_REQUIRE(!params_out || (w->W_h && w->b_h && cur_pos && params_stride >= 5), "head needs W_h, b_h, cur_pos");
  MMT_REQUIRE(h_out != h && c_out != c, "outputs must not alias the input state (rows are re-read by other CTAs)");
  MMT_ALIGNED(x); MMT_ALIGNED(h); MMT_ALIGNED(c); MMT_ALIGNED(mh); MMT_ALIGNED(mc);
  MMT_ALIGNED(h_out); MMT_ALIGNED(c_out); MMT_ALIGNED(mf_out);
  if (R == 0) return MMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int U = w->U;
  if (prec == MMT_PREC_F32) {
    int rc = launch_cell_f32(x, h, c, mh, mc, valid, w, R, U, h_out, c_out, mf_out, U, st);
    if (rc) return rc;
    if (params_out) rc = launch_head(h_out, U, mf_out, U, valid, w, R, cur_pos, params_out, params_stride, next_pos, st);
    return rc;
  }
  MMT_REQUIRE(w->W_packed_bf16, "bf16 mode needs W_packed_bf16 (mmt_pack_gate_weights_bf16)");
  return launch_cell_tc(x, h, c, mh, mc, U, valid, w, R, h_out, c_out, mf_out, U, cur_pos, params_out, params_stride,
                        next_pos, st);
}
