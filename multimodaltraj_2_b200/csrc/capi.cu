// C-ABI plumbing of libmmt: version, thread-local error string, launch counter, and the
// mmt_gsk_cell entry point that dispatches between the fp32 and the tcgen05/bf16 cell kernels.
#include <atomic>
#include <cstdarg>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

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

static std::atomic<uint32_t*> g_trap{nullptr};
static std::mutex g_mu;

uint32_t* trap_record() {
  uint32_t* p = g_trap.load(std::memory_order_acquire);
  if (p) return p;
  std::lock_guard<std::mutex> lk(g_mu);
  p = g_trap.load(std::memory_order_relaxed);
  if (p) return p;
  void* h = nullptr;
  if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();   // not sticky: the kernels then trap without a record
    return nullptr;
  }
  memset(h, 0, 64);
  g_trap.store(static_cast<uint32_t*>(h), std::memory_order_release);
  return static_cast<uint32_t*>(h);
}

// "<kernel>/<wait>" of a trap site code (tc_common.cuh: trap_report).  Kernel ids: 1 rollout_tc, 2 gsk_cell_tc,
// 3 graph_aggregate_mma, 4 node_proj_tc, 5 edge_mlp_tc, 6 gemm_tf32; wait 0xFF = shared-memory base not 1024-byte aligned.
static void describe_trap(char* out, size_t n) {
  const uint32_t* r = g_trap.load(std::memory_order_acquire);
  if (!r || r[0] == 0) { out[0] = 0; return; }
  static const char* kern[] = {"?", "rollout_tc_kernel", "gsk_cell_tc_kernel", "graph_aggregate_mma_kernel",
                               "node_proj_tc_kernel", "edge_mlp_tc_kernel", "gemm_tf32_kernel", "edge_mlp_bwd_tc_kernel"};
  const uint32_t k = r[0] >> 8, w = r[0] & 0xFFu;
  snprintf(out, n, " [device trap: %s site 0x%02x%s, CTA %u thread %u, barrier smem 0x%x parity %u]",
           k < 8 ? kern[k] : "?", w, w == 0xFF ? " (smem base misaligned)" : " (bounded mbarrier wait expired)", r[1], r[2],
           r[3], r[4]);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char trap[256];
    describe_trap(trap, sizeof(trap));
    set_error("%s: %s%s", what, cudaGetErrorString(e), trap);
    return MMT_ECUDA;
  }
  return MMT_OK;
}

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

int opt_in_smem(const void* func, int bytes, DeviceMask* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return check_launch("cudaGetDevice");
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask->load(std::memory_order_acquire) & bit) return MMT_OK;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) on device %d: %s", bytes, dev, cudaGetErrorString(e));
    return MMT_ECUDA;
  }
  mask->fetch_or(bit, std::memory_order_release);
  return MMT_OK;
}

int env_int_once(const char* name, std::atomic<int>* c) {
  int v = c->load(std::memory_order_relaxed);
  if (v != INT_MIN) return v;
  const char* s = getenv(name);
  v = s ? atoi(s) : 0;
  c->store(v, std::memory_order_relaxed);
  return v;
}

int launch_cell_f32(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                    const uint8_t* valid, const mmt_cell_weights* w, int R, int ld, float* h_out, float* c_out,
                    float* mf_out, int ld_mf, cudaStream_t stream);
int launch_head(const float* mt, int ld, const float* mf, int ld_mf, const uint8_t* valid, const mmt_cell_weights* w,
                int R, const float* cur_pos, float* params_out, int params_stride, float* next_pos,
                cudaStream_t stream);
int launch_cell_tc(const float* x, const float* h, const float* c, const float* mh, const float* mc, int ld,
                   const uint8_t* valid, const mmt_cell_weights* w, int R, float* h_out, float* c_out, float* mf_out,
                   int ld_mf, const float* cur_pos, float* params_out, int params_stride, float* next_pos, int x3,
                   cudaStream_t stream);

}  // namespace mmt

extern "C" int mmt_version(void) { return MMT_VERSION; }
extern "C" const char* mmt_last_error(void) { return mmt::g_err; }
extern "C" uint64_t mmt_launch_count(void) { return mmt::g_launches.load(std::memory_order_relaxed); }
extern "C" int mmt_last_trap(uint32_t out[8]) {
  const uint32_t* r = mmt::g_trap.load(std::memory_order_acquire);
  for (int i = 0; i < 8; ++i) out[i] = r ? r[i] : 0u;
  return (r && r[0]) ? 1 : 0;
}
extern "C" int mmt_num_sms(void) { return mmt::num_sms(); }

extern "C" int mmt_gsk_cell(const float* x, const float* h, const float* c, const float* mh, const float* mc,
                            const uint8_t* valid, const mmt_cell_weights* w, int R, int prec, float* h_out,
                            float* c_out, float* mf_out, const float* cur_pos, float* params_out, int params_stride,
                            float* next_pos, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(x && h && c && mh && mc && valid && w && h_out && c_out && mf_out, "state/weight/output pointers required");
  MMT_REQUIRE(w->W_e && w->b_e && w->W && w->b && w->w_If && w->w_It && w->w_Of && w->w_Ot, "cell weights required");
  MMT_REQUIRE(R >= 0, "R must be >= 0");
  MMT_REQUIRE(w->E == 64 && w->U == 128, "cell is built for E = 64, U = 128");
  MMT_REQUIRE(prec == MMT_PREC_F32 || prec == MMT_PREC_BF16 || prec == MMT_PREC_BF16X3 || prec == MMT_PREC_F16,
              "unknown precision mode");
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
  const int x3 = prec == MMT_PREC_BF16X3 ? 1 : prec == MMT_PREC_F16 ? 2 : 0;   // operand mode of launch_cell_tc
  MMT_REQUIRE(x3 == 1 ? w->W_packed_bf16x3 != nullptr : x3 == 2 ? w->W_packed_f16 != nullptr : w->W_packed_bf16 != nullptr,
              "tensor-core modes need the packed operand image (mmt_pack_gate_weights_bf16 / _bf16x3 / _f16)");
  return launch_cell_tc(x, h, c, mh, mc, U, valid, w, R, h_out, c_out, mf_out, U, cur_pos, params_out, params_stride,
                        next_pos, x3, st);
}
