// Device-side padded scene batching (include/mmt.h mmt_scene_batch_f32).
//
// Replaces the dict walking of load_traj.py:153-224 (DataLoader.next_step) and
// networkx_graph.py:30-73,114-129 (ConstructGraph / setNodes) for window extraction: the
// trajectory table lives on the device sorted by (frame, ped) with a CSR index over frames; one
// warp builds one scene window [N, F, 2] with a stable slot order (ascending ped id of the
// window's first frame) and a validity mask.  Integer outputs (slots, mask) are bit-exact
// against the oracle's Python restatement.
#include "mmt_common.cuh"

namespace mmt {

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) scene_batch_kernel(const int32_t* __restrict__ frame_ids,
                                                          const int32_t* __restrict__ frame_row_start, int n_frames,
                                                          const int32_t* __restrict__ ped_id,
                                                          const float* __restrict__ xy, const float* __restrict__ vis,
                                                          const int32_t* __restrict__ win_start, int S, int N, int F,
                                                          int fstride, float* __restrict__ pos,
                                                          float* __restrict__ visout, uint8_t* __restrict__ valid,
                                                          int32_t* __restrict__ ped_of_slot) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < S; s += gridDim.x * wpb) {
    // clear the scene
    float* ps = pos + (size_t)s * N * F * 2;
    for (int i = lane; i < N * F * 2; i += 32) ps[i] = 0.f;
    if (visout) {
      float* vs = visout + (size_t)s * N * F * 2;
      for (int i = lane; i < N * F * 2; i += 32) vs[i] = 0.f;
    }
    for (int i = lane; i < N; i += 32) {
      valid[(size_t)s * N + i] = 0;
      ped_of_slot[(size_t)s * N + i] = -1;
    }
    __syncwarp();
    // lane k locates frame k of the window (F <= 32)
    int rs = 0, re = 0;
    bool present = true;
    if (lane < F) {
      const int fid = win_start[s] + lane * fstride;
      const int idx = lower_bound_i32(frame_ids, 0, n_frames, fid);
      present = idx < n_frames && frame_ids[idx] == fid;
      if (present) {
        rs = frame_row_start[idx];
        re = frame_row_start[idx + 1];
      }
    }
    if (!__all_sync(0xffffffffu, present)) continue;
    const int rs0 = __shfl_sync(0xffffffffu, rs, 0), re0 = __shfl_sync(0xffffffffu, re, 0);
    int nslots = 0;
    for (int c0 = rs0; c0 < re0 && nslots < N; c0 += 32) {
      const int row0 = c0 + lane;
      const bool cand = row0 < re0;
      const int ped = cand ? ped_id[row0] : 0;
      bool ok = cand;
      int rows[32];
      rows[0] = row0;
      for (int k = 1; k < F; ++k) {
        const int ks = __shfl_sync(0xffffffffu, rs, k), ke = __shfl_sync(0xffffffffu, re, k);
        int r = -1;
        if (ok) {
          const int p = lower_bound_i32(ped_id, ks, ke, ped);
          if (p < ke && ped_id[p] == ped) r = p; else ok = false;
        }
        rows[k] = r;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int slot = nslots + __popc(m & ((1u << lane) - 1u));
      if (ok && slot < N) {
        valid[(size_t)s * N + slot] = 1;
        ped_of_slot[(size_t)s * N + slot] = ped;
        for (int k = 0; k < F; ++k) {
          const float2 p = *reinterpret_cast<const float2*>(xy + (size_t)rows[k] * 2);
          *reinterpret_cast<float2*>(ps + ((size_t)slot * F + k) * 2) = p;
          if (visout && vis)
            *reinterpret_cast<float2*>(visout + (((size_t)s * N + slot) * F + k) * 2) =
                *reinterpret_cast<const float2*>(vis + (size_t)rows[k] * 2);
        }
      }
      nslots += __popc(m);
    }
  }
}

}  // namespace mmt

extern "C" int mmt_scene_batch_f32(const int32_t* frame_ids_sorted, const int32_t* frame_row_start, int n_frames,
                                   const int32_t* ped_id, const float* xy, const float* vis, const int32_t* win_start,
                                   int S, int N, int F, int fstride, float* pos, float* visout, uint8_t* valid,
                                   int32_t* ped_of_slot, void* stream) {
  using namespace mmt;
  MMT_REQUIRE(frame_ids_sorted && frame_row_start && ped_id && xy && win_start && pos && valid && ped_of_slot,
              "table / window / output pointers must not be NULL");
  MMT_REQUIRE(S >= 0 && N > 0 && F > 0 && F <= 32 && fstride > 0 && n_frames >= 0, "need 0 < F <= 32, fstride > 0");
  MMT_REQUIRE(!visout || vis, "visout requires vis");
  if (S == 0) return MMT_OK;
  int blocks = (S + 3) / 4;
  int grid = blocks < num_sms() * 8 ? blocks : num_sms() * 8;
  scene_batch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(frame_ids_sorted, frame_row_start, n_frames, ped_id, xy,
                                                             vis, win_start, S, N, F, fstride, pos, visout, valid,
                                                             ped_of_slot);
  count_launch();
  return check_launch("scene_batch_kernel");
}
