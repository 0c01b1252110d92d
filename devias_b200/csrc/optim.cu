// Optimizer update over the flat parameter arena: ONE pass reads {param, grad, exp_avg, exp_avg_sq} and writes {param, exp_avg,
// exp_avg_sq, bf16 shadow of the parameter, grad = 0}.  Replaces, per training step, torch's multi-tensor AdamW launches, the
// stand-alone weight cast that refreshed the tensor-core operand copies, and the gradient memset
// (reference: torch.optim.AdamW built by utils/optim_factory.py:94-178, stepped at engine/engine_for_slot.py:147-166 with the
// per-group `lr * lr_scale` / `weight_decay` written every iteration at :91-97 and optional clip_grad_norm_ via
// utils/utils.py NativeScaler).  Hyper-parameters live in DEVICE memory so that a captured CUDA graph follows the schedule.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

// hyper layout (floats): [0] 1 - beta1^t   [1] 1 - beta2^t   [2] beta1   [3] beta2   [4] eps   [5] max_norm (<= 0: no clipping)
//                        [6] gradient pre-scale (1 / update_freq etc.)   [7] unused;  then per group g: [8 + 2g] lr, [9 + 2g] weight decay
constexpr int HYPER_GLOBAL = 8;
constexpr int MAX_SEGS = 1024;

// sum of squares of the gradient arena -> *out (fp32, pre-zeroed); the clip coefficient is derived from it inside the update
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

// seg_start[s] (in 8-element granules, ascending, seg_start[n_seg] = total granules), seg_group[s] = hyper-parameter group
__global__ void __launch_bounds__(256) adamw_arena_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, __nv_bfloat16* __restrict__ p16,
                                                          const int* __restrict__ seg_start, const int* __restrict__ seg_group,
                                                          int n_seg, const float* __restrict__ hyper,
                                                          const float* __restrict__ sumsq, long long n8, int zero_grad,
                                                          const __nv_bfloat16* __restrict__ g16) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_start[MAX_SEGS + 1];
  __shared__ unsigned char s_group[MAX_SEGS];
  for (int i = threadIdx.x; i <= n_seg; i += blockDim.x) s_start[i] = seg_start[i];
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) s_group[i] = (unsigned char)seg_group[i];
  __syncthreads();
  const float bc1 = hyper[0], bc2 = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4], max_norm = hyper[5];
  float gscale = hyper[6];
  if (sumsq != nullptr && max_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float coef = max_norm / (sqrtf(*sumsq) * fabsf(gscale) + 1e-6f);
    gscale *= fminf(coef, 1.f);
  }
  const float rs_bc2 = rsqrtf(bc2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  int seg = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    // segment of granule i: binary search (first granule of this thread), then walk forward
    if (!(s_start[seg] <= i && i < s_start[seg + 1])) {
      int lo = 0, hi = n_seg - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_start[mid] <= i) lo = mid; else hi = mid - 1;
      }
      seg = lo;
    }
    const int grp = s_group[seg];
    const float lr = hyper[HYPER_GLOBAL + 2 * grp], wd = hyper[HYPER_GLOBAL + 2 * grp + 1];
    const float decay = 1.f - lr * wd, step = lr / bc1;
    float pp[8], gg[8], mm[8], vv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 a = reinterpret_cast<const float4*>(p)[2 * i + h];
      float4 b;
      if (g16 != nullptr) {       // gradients exchanged in bf16 (data parallel): the reduced values live in the bf16 buffer
        const uint2 q = __ldcs(reinterpret_cast<const uint2*>(g16) + 2 * i + h);
        const float2 lo = unpack_bf16(q.x), hi = unpack_bf16(q.y);
        b = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        b = __ldcs(reinterpret_cast<const float4*>(g) + 2 * i + h);
      }
      const float4 c = __ldcs(reinterpret_cast<const float4*>(m) + 2 * i + h), d = __ldcs(reinterpret_cast<const float4*>(v) + 2 * i + h);
      pp[4 * h] = a.x; pp[4 * h + 1] = a.y; pp[4 * h + 2] = a.z; pp[4 * h + 3] = a.w;
      gg[4 * h] = b.x; gg[4 * h + 1] = b.y; gg[4 * h + 2] = b.z; gg[4 * h + 3] = b.w;
      mm[4 * h] = c.x; mm[4 * h + 1] = c.y; mm[4 * h + 2] = c.z; mm[4 * h + 3] = c.w;
      vv[4 * h] = d.x; vv[4 * h + 1] = d.y; vv[4 * h + 2] = d.z; vv[4 * h + 3] = d.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gr = gg[e] * gscale;
      pp[e] *= decay;                                         // decoupled weight decay (AdamW)
      mm[e] = fmaf(b1, mm[e], (1.f - b1) * gr);
      vv[e] = fmaf(b2, vv[e], (1.f - b2) * gr * gr);
      const float denom = sqrtf(vv[e]) * rs_bc2 + eps;
      pp[e] -= step * (mm[e] / denom);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      reinterpret_cast<float4*>(p)[2 * i + h] = make_float4(pp[4 * h], pp[4 * h + 1], pp[4 * h + 2], pp[4 * h + 3]);
      __stcs(reinterpret_cast<float4*>(m) + 2 * i + h, make_float4(mm[4 * h], mm[4 * h + 1], mm[4 * h + 2], mm[4 * h + 3]));
      __stcs(reinterpret_cast<float4*>(v) + 2 * i + h, make_float4(vv[4 * h], vv[4 * h + 1], vv[4 * h + 2], vv[4 * h + 3]));
      if (zero_grad) reinterpret_cast<float4*>(g)[2 * i + h] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p16 != nullptr)
      reinterpret_cast<uint4*>(p16)[i] = make_uint4(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]), pack_bf16(pp[4], pp[5]),
                                                    pack_bf16(pp[6], pp[7]));
  }
}

}  // namespace dv

extern "C" int devias_sumsq_f32(const float* x, int64_t n, float* out, void* stream) {
  using namespace dv;
  DV_REQUIRE(x && out, "null pointer");
  DV_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "length must be a multiple of 4, 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DV_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  if (n <= 0) return DEVIAS_OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  DV_CHECK_CUDA(launch_k(sumsq_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), s, x, n / 4, out));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_adamw_arena(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                                  const int32_t* seg_start, const int32_t* seg_group, int n_seg, const float* hyper,
                                  const float* grad_sumsq, int64_t n, int zero_grad, const void* grad_bf16, void* stream) {
  using namespace dv;
  DV_REQUIRE(param && grad && exp_avg && exp_avg_sq && seg_start && seg_group && hyper, "null pointer");
  DV_REQUIRE(n % 8 == 0, "arena length must be a multiple of 8 elements");
  DV_REQUIRE(n_seg >= 1 && n_seg <= MAX_SEGS, "1..1024 parameter segments");
  DV_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
               reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(param_bf16)) & 15) == 0, "16-byte alignment");
  if (n <= 0) return DEVIAS_OK;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  DV_CHECK_CUDA(launch_k(adamw_arena_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), static_cast<cudaStream_t>(stream), 
      param, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(param_bf16), seg_start, seg_group, n_seg, hyper, grad_sumsq,
      n8, zero_grad, static_cast<const __nv_bfloat16*>(grad_bf16)));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
