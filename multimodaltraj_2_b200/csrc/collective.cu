// The one collective of the path: the gradient all-reduce of data-parallel training (and the few
// floats of ADE/FDE partial sums), as a thin C-ABI wrapper over the NCCL that is ALREADY loaded in the process
// (include/mmt.h: mmt_allreduce_f32).  Inference shards scenes with no data-path collective at all.
//
// libmmt does not link NCCL: the communicator comes from the caller (torch.distributed's ProcessGroupNCCL, which
// bundles its own libnccl.so.2), and a communicator is only valid inside the library instance that created it.  The
// entry points are therefore resolved at first use from the loaded library (dlopen RTLD_NOLOAD by soname).
#include <dlfcn.h>

#include <atomic>
#include <mutex>

#include "mmt_common.cuh"

namespace mmt {
namespace {

// stable NCCL 2.x ABI (nccl.h: ncclDataType_t, ncclRedOp_t)
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclMax = 2;
using AllReduceFn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using ErrStrFn = const char* (*)(int);

struct Nccl {
  AllReduceFn all_reduce = nullptr;
  ErrStrFn err_str = nullptr;
};

const Nccl* nccl() {
  static Nccl api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_NOLOAD);          // the instance the caller's communicator lives in
      if (h) break;
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);   // nobody loaded NCCL yet: the system library
    if (!h) return;
    api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
    api.err_str = reinterpret_cast<ErrStrFn>(dlsym(h, "ncclGetErrorString"));
  });
  return api.all_reduce ? &api : nullptr;
}

int all_reduce(void* comm, float* buf, size_t n, int op, cudaStream_t stream, const char* what) {
  const Nccl* api = nccl();
  if (!api) {
    set_error("%s: NCCL is not loaded in this process (libnccl.so.2 not found)", what);
    return MMT_ENCCL;
  }
  const int rc = api->all_reduce(buf, buf, n, kNcclFloat32, op, comm, stream);
  if (rc != 0) {
    set_error("%s: ncclAllReduce failed (%d): %s", what, rc, api->err_str ? api->err_str(rc) : "?");
    return MMT_ENCCL;
  }
  return MMT_OK;
}

}  // namespace
}  // namespace mmt

extern "C" int mmt_allreduce_f32(void* comm, float* buf, size_t n, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(comm, "comm (ncclComm_t) must not be NULL");
  MMT_REQUIRE(buf || n == 0, "buf must not be NULL");
  if (n == 0) return MMT_OK;
  return all_reduce(comm, buf, n, kNcclSum, (cudaStream_t)stream, "mmt_allreduce_f32");
}

extern "C" int mmt_allreduce_max_f32(void* comm, float* buf, size_t n, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(comm, "comm (ncclComm_t) must not be NULL");
  MMT_REQUIRE(buf || n == 0, "buf must not be NULL");
  if (n == 0) return MMT_OK;
  return all_reduce(comm, buf, n, kNcclMax, (cudaStream_t)stream, "mmt_allreduce_max_f32");
}

// generic workspace query: one entry point over the per-operation ones
extern "C" int mmt_workspace_bytes(int op, const mmt_shape* s, size_t* out) {
  using namespace mmt;
  MMT_REQUIRE(s && out, "shape / out must not be NULL");
  switch (op) {
    case MMT_OP_FORECAST: {
      mmt_forecast_cfg cfg = {};
      cfg.S = s->S; cfg.N = s->N; cfg.T = s->T; cfg.P = s->P; cfg.K = s->K;
      cfg.relational = s->relational; cfg.prec = s->prec;
      *out = mmt_forecast_workspace_bytes(&cfg, s->U, s->He);
      return MMT_OK;
    }
    case MMT_OP_EDGE_MLP:
      *out = (size_t)2 * s->S * s->N * s->He * sizeof(float);
      return MMT_OK;
    case MMT_OP_STATIC_CONTEXT:
      *out = mmt_static_context_workspace_bytes(s->img_h, s->D);
      return MMT_OK;
    case MMT_OP_GATE_WEIGHTS_BF16:
      *out = mmt_gate_weights_packed_bytes(s->E, s->U);
      return MMT_OK;
    case MMT_OP_GATE_WEIGHTS_BF16X3:
      *out = mmt_gate_weights_packed_x3_bytes(s->E, s->U);
      return MMT_OK;
    case MMT_OP_EDGE_WEIGHTS_BF16:
      *out = mmt_edge_weights_packed_bytes(s->U, s->He);
      return MMT_OK;
    default:
      set_error("mmt_workspace_bytes: unknown op %d", op);
      return MMT_EARG;
  }
}
