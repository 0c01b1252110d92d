// Slot selection of the DEVIAS head (model/modeling_slot.py:396-404, model/modeling_slot_fusion.py:376-386):
//   probs = softmax(head(slots)) per (clip, slot) row over all C + Cs logits
//   action slot = argmax_s max_{c < C} probs[s, c] ;  scene slot = argmax_s max_{C <= c < C + Cs} probs[s, c]
// One CTA per clip, one warp per slot: row maximum and sum of exponentials by warp shuffles (two passes over the row), the
// per-slot scores exp(max_group - max_row) / sum meet in shared memory and thread 0 takes the first maximum (torch.argmax
// tie order).  Replaces softmax + 2 x (slice, max, argmax) = ~10 launches on B*S rows by one.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kSelMaxSlots = 8;

__global__ void __launch_bounds__(32 * kSelMaxSlots) slot_select_kernel(const float* __restrict__ logits, long long ld, int slots,
                                                                        int n_action, int n_scene, long long* __restrict__ a_idx,
                                                                        long long* __restrict__ s_idx) {
  pdl_trigger();
  pdl_wait();
  __shared__ float score[2][kSelMaxSlots];
  const int b = blockIdx.x, s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = n_action + n_scene;
  if (s < slots) {
    const float* row = logits + ((long long)b * slots + s) * ld;
    float ma = -INFINITY, ms = -INFINITY;
    for (int c = lane; c < n; c += 32) {
      const float v = __ldg(row + c);
      if (c < n_action) ma = fmaxf(ma, v); else ms = fmaxf(ms, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
      ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
    }
    const float mx = fmaxf(ma, ms);
    float sum = 0.f;
    for (int c = lane; c < n; c += 32) sum += expf(__ldg(row + c) - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      score[0][s] = expf(ma - mx) / sum;
      score[1][s] = expf(ms - mx) / sum;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int g = threadIdx.x;
    int best = 0;
    float bv = score[g][0];
    for (int i = 1; i < slots; ++i)
      if (score[g][i] > bv) { bv = score[g][i]; best = i; }
    (g == 0 ? a_idx : s_idx)[b] = best;
  }
}

}  // namespace dv

extern "C" int devias_slot_select(const float* logits, int64_t ld, int batch, int slots, int n_action, int n_scene,
                                  long long* action_idx, long long* scene_idx, void* stream) {
  using namespace dv;
  DV_REQUIRE(logits && action_idx && scene_idx, "null pointer");
  DV_REQUIRE(batch > 0 && slots > 0 && slots <= kSelMaxSlots, "slots must be 1..8");
  DV_REQUIRE(n_action > 0 && n_scene > 0 && ld >= n_action + n_scene, "bad class counts / row stride");
  DV_CHECK_CUDA(launch_k(slot_select_kernel, dim3((unsigned)(batch)), dim3((unsigned)(32 * slots)), (size_t)(0), (cudaStream_t)stream, logits, ld, slots, n_action, n_scene, action_idx, scene_idx));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
