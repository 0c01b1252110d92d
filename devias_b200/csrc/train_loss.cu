// The DEVIAS training objective ('matching' branch of utils/loss/train_loss.py:85-187) as ONE forward and ONE backward launch.
// Per clip (one CTA each; everything in the objective is per clip except the batch means and one batch-wide scalar):
//   slots_head [S, W = C + 365] -> softmax -> cost[s] = (-p[s, target], -p[s, scene_target]) -> assignment of distinct slots (i, j)
//   minimising cost[i, 0] + cost[j, 1] (:112-122: scipy linear_sum_assignment on an S x 2 matrix; first minimum in (i, j) order)
//   action  : cross_entropy(head[i], target)                                                        (:150)
//   scene   : kl_div(log_softmax(head[j]), log_softmax([var x C | teacher]), 'batchmean', log_target) * 2000  (:160-165; on a
//             1-D row 'batchmean' divides by W) or cross_entropy(head[j], scene_target)             (:157)
//   distill : mse(mean_heads attn[i], fg_frames) * 3                                                (:145)
//   maskpred: binary_cross_entropy_with_logits(mask_predictions[i], fg) * 1                         (:146-149)
//   cosine  : mean over ordered pairs i != j of <s_i / |s_i|, s_j / |s_j|>                          (:173-178)
// each summed over the batch and divided by the batch size.  `var` = min over the WHOLE batch of the teacher logits - 1 (:103) is
// the one cross-clip quantity: the caller passes it as a device scalar.
// The reference evaluates this with B host synchronisations (scipy per clip) and O(B * S) tiny kernels; the torch restatement in
// devias_b200/loss.py still cost ~100 launches forward + as many backward.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kTlThreads = 256;
constexpr int kTlMaxS = 8;

__device__ __forceinline__ float tl_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kTlThreads / 32; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float tl_block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int i = 0; i < kTlThreads / 32; ++i) t = fmaxf(t, red[i]);
  return t;
}
// log-sum-exp of a row of n floats (+ `extra` copies of the value xv), all threads get the result
__device__ __forceinline__ float tl_row_lse(const float* row, int n, int extra, float xv, float* red) {
  float mx = extra > 0 ? xv : -INFINITY;
  for (int i = threadIdx.x; i < n; i += kTlThreads) mx = fmaxf(mx, row[i]);
  mx = tl_block_max(mx, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kTlThreads) s += expf(row[i] - mx);
  s = tl_block_sum(s, red) + (float)extra * expf(xv - mx);
  return mx + logf(s);
}

struct TlParams {
  const float* head;        // [B*S, W]
  const float* attn;        // [B*H, S, N]
  const float* maskp;       // [B*S, P]
  const float* slots;       // [B*S, D]
  const long long* target;  // [B]
  const float* teacher;     // [B, T]  (T = W - C scene classes)
  const float* var;         // device scalar: min(teacher) - 1
  const float* fg;          // [B, P]
  const float* fgf;         // [B, N]
  int B, S, W, C, H, N, P, D;
  int scene_ce;             // 0: KL (the DEVIAS recipes), 1: CE
  float w_scene, w_maskpred, w_distill;
};

// shared per-clip analysis used by forward and backward: lse of every slot row, teacher arg-max / lse, the assignment
struct TlClip {
  float lse[kTlMaxS];
  float lse_t;
  int scene_target, ai, si;
};

__device__ __forceinline__ void tl_analyse(const TlParams& p, int b, TlClip& c, float* red, int* ired) {
  const int T = p.W - p.C;
  const float* trow = p.teacher + (long long)b * T;
  // arg-max of the teacher row (first maximum, as torch.argmax)
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < T; i += kTlThreads) {
    const float v = trow[i];
    if (v > best) { best = v; bi = i; }
  }
  const float mx = tl_block_max(best, red);
  int cand = (best == mx) ? bi : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) ired[threadIdx.x >> 5] = cand;
  __syncthreads();
  int am = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kTlThreads / 32; ++i) am = min(am, ired[i]);
  c.scene_target = p.C + am;
  const float var = *p.var;
  c.lse_t = tl_row_lse(trow, T, p.C, var, red);
  const int tgt = (int)p.target[b];
  float ca[kTlMaxS], cs[kTlMaxS];
  for (int s = 0; s < p.S; ++s) {
    const float* row = p.head + ((long long)b * p.S + s) * p.W;
    c.lse[s] = tl_row_lse(row, p.W, 0, 0.f, red);
    ca[s] = -expf(row[tgt] - c.lse[s]);
    cs[s] = -expf(row[c.scene_target] - c.lse[s]);
  }
  float bestc = INFINITY;
  c.ai = 0; c.si = 1;
  for (int i = 0; i < p.S; ++i)
    for (int j = 0; j < p.S; ++j) {
      if (i == j) continue;
      const float v = ca[i] + cs[j];
      if (v < bestc) { bestc = v; c.ai = i; c.si = j; }
    }
}

// out[0..4] += (action, scene, cosine, maskpred, distill) / B, out[5] += their sum ;  idx[b] = (action slot, scene slot)
__global__ void __launch_bounds__(kTlThreads) train_loss_fwd_kernel(const TlParams p, float* __restrict__ out, long long* __restrict__ idx) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kTlThreads / 32];
  __shared__ int ired[kTlThreads / 32];
  __shared__ float inv_norm[kTlMaxS];
  const int b = blockIdx.x, tid = threadIdx.x;
  TlClip c;
  tl_analyse(p, b, c, red, ired);
  const int tgt = (int)p.target[b];
  const float* arow = p.head + ((long long)b * p.S + c.ai) * p.W;
  const float* srow = p.head + ((long long)b * p.S + c.si) * p.W;
  const float action = c.lse[c.ai] - arow[tgt];
  float scene;
  if (p.scene_ce) {
    scene = c.lse[c.si] - srow[c.scene_target];
  } else {
    const float var = *p.var;
    const float* trow = p.teacher + (long long)b * (p.W - p.C);
    float acc = 0.f;
    for (int w = tid; w < p.W; w += kTlThreads) {
      const float lt = (w < p.C ? var : trow[w - p.C]) - c.lse_t;
      const float ls = srow[w] - c.lse[c.si];
      acc += expf(lt) * (lt - ls);
    }
    scene = tl_block_sum(acc, red) / (float)p.W * p.w_scene;
  }
  // distillation of the action slot's attention (mean over heads) against the per-frame foreground mask
  float md = 0.f;
  {
    const float invh = 1.0f / (float)p.H;
    for (int n = tid; n < p.N; n += kTlThreads) {
      float a = 0.f;
      for (int h = 0; h < p.H; ++h) a += p.attn[(((long long)b * p.H + h) * p.S + c.ai) * p.N + n];
      const float d = a * invh - p.fgf[(long long)b * p.N + n];
      md = fmaf(d, d, md);
    }
    md = tl_block_sum(md, red) / (float)p.N * p.w_distill;
  }
  float mp = 0.f;
  {
    const float* x = p.maskp + ((long long)b * p.S + c.ai) * p.P;
    for (int k = tid; k < p.P; k += kTlThreads) {
      const float v = x[k], y = p.fg[(long long)b * p.P + k];
      mp += fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));      // binary_cross_entropy_with_logits
    }
    mp = tl_block_sum(mp, red) / (float)p.P * p.w_maskpred;
  }
  // cosine similarity between the slots of the clip
  float cosl = 0.f;
  {
    for (int s = 0; s < p.S; ++s) {
      const float* v = p.slots + ((long long)b * p.S + s) * p.D;
      float q = 0.f;
      for (int d = tid; d < p.D; d += kTlThreads) q = fmaf(v[d], v[d], q);
      q = tl_block_sum(q, red);
      if (tid == 0) inv_norm[s] = 1.0f / fmaxf(sqrtf(q), 1e-12f);
    }
    __syncthreads();
    float acc = 0.f;
    for (int i = 0; i < p.S; ++i)
      for (int j = i + 1; j < p.S; ++j) {
        const float* vi = p.slots + ((long long)b * p.S + i) * p.D;
        const float* vj = p.slots + ((long long)b * p.S + j) * p.D;
        float q = 0.f;
        for (int d = tid; d < p.D; d += kTlThreads) q = fmaf(vi[d], vj[d], q);
        acc += 2.0f * q * inv_norm[i] * inv_norm[j];
      }
    cosl = tl_block_sum(acc, red) / (float)(p.S * (p.S - 1));
  }
  if (tid == 0) {
    const float ib = 1.0f / (float)p.B;
    atomicAdd(out + 0, action * ib);
    atomicAdd(out + 1, scene * ib);
    atomicAdd(out + 2, cosl * ib);
    atomicAdd(out + 3, mp * ib);
    atomicAdd(out + 4, md * ib);
    atomicAdd(out + 5, (action + scene + cosl + mp + md) * ib);
    idx[2 * b] = c.ai;
    idx[2 * b + 1] = c.si;
  }
}

// gradients of the TOTAL (sum of the five parts) times *gtot, written in full (zeros where a tensor does not take part)
__global__ void __launch_bounds__(kTlThreads) train_loss_bwd_kernel(const TlParams p, const float* __restrict__ gtot, float* __restrict__ dhead,
                                                                    float* __restrict__ dattn, float* __restrict__ dmaskp,
                                                                    float* __restrict__ dslots) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kTlThreads / 32];
  __shared__ int ired[kTlThreads / 32];
  __shared__ float inv_norm[kTlMaxS];
  __shared__ float dots[kTlMaxS][kTlMaxS];
  const int b = blockIdx.x, tid = threadIdx.x;
  TlClip c;
  tl_analyse(p, b, c, red, ired);
  const float g = *gtot / (float)p.B;
  const int tgt = (int)p.target[b];
  const float var = *p.var;
  const float* trow = p.teacher + (long long)b * (p.W - p.C);
  for (int s = 0; s < p.S; ++s) {
    const float* row = p.head + ((long long)b * p.S + s) * p.W;
    float* drow = dhead + ((long long)b * p.S + s) * p.W;
    for (int w = tid; w < p.W; w += kTlThreads) {
      float d = 0.f;
      if (s == c.ai) d += g * (expf(row[w] - c.lse[s]) - (w == tgt ? 1.f : 0.f));
      if (s == c.si) {
        const float ps = expf(row[w] - c.lse[s]);
        if (p.scene_ce) d += g * (ps - (w == c.scene_target ? 1.f : 0.f));
        else d += g * p.w_scene / (float)p.W * (ps - expf((w < p.C ? var : trow[w - p.C]) - c.lse_t));
      }
      drow[w] = d;
    }
  }
  {
    const float invh = 1.0f / (float)p.H;
    const float k = g * p.w_distill * 2.0f / (float)p.N * invh;
    for (int n = tid; n < p.N; n += kTlThreads) {
      float a = 0.f;
      for (int h = 0; h < p.H; ++h) a += p.attn[(((long long)b * p.H + h) * p.S + c.ai) * p.N + n];
      const float d = k * (a * invh - p.fgf[(long long)b * p.N + n]);
      for (int h = 0; h < p.H; ++h)
        for (int s = 0; s < p.S; ++s) dattn[(((long long)b * p.H + h) * p.S + s) * p.N + n] = (s == c.ai) ? d : 0.f;
    }
  }
  for (int s = 0; s < p.S; ++s) {
    const float* x = p.maskp + ((long long)b * p.S + s) * p.P;
    float* dx = dmaskp + ((long long)b * p.S + s) * p.P;
    for (int k = tid; k < p.P; k += kTlThreads)
      dx[k] = (s == c.ai) ? g * p.w_maskpred / (float)p.P * (1.0f / (1.0f + expf(-x[k])) - p.fg[(long long)b * p.P + k]) : 0.f;
  }
  // cosine: L = sum_{i != j} <n_i, n_j> / (S (S - 1)),  n_i = s_i / |s_i|:  dL/ds_i = 2 / (S (S - 1)) sum_{j != i} (n_j - <n_i, n_j> n_i) / |s_i|
  for (int i = 0; i < p.S; ++i)
    for (int j = i; j < p.S; ++j) {
      const float* vi = p.slots + ((long long)b * p.S + i) * p.D;
      const float* vj = p.slots + ((long long)b * p.S + j) * p.D;
      float q = 0.f;
      for (int d = tid; d < p.D; d += kTlThreads) q = fmaf(vi[d], vj[d], q);
      q = tl_block_sum(q, red);
      if (tid == 0) { dots[i][j] = q; dots[j][i] = q; }
    }
  __syncthreads();
  if (tid < p.S) inv_norm[tid] = 1.0f / fmaxf(sqrtf(dots[tid][tid]), 1e-12f);
  __syncthreads();
  const float kc = g * 2.0f / (float)(p.S * (p.S - 1));
  for (int i = 0; i < p.S; ++i) {
    const float* vi = p.slots + ((long long)b * p.S + i) * p.D;
    float* di = dslots + ((long long)b * p.S + i) * p.D;
    float self = 0.f;                                   // sum_{j != i} <n_i, n_j>
    for (int j = 0; j < p.S; ++j)
      if (j != i) self += dots[i][j] * inv_norm[i] * inv_norm[j];
    for (int d = tid; d < p.D; d += kTlThreads) {
      float acc = 0.f;
      for (int j = 0; j < p.S; ++j)
        if (j != i) acc += p.slots[((long long)b * p.S + j) * p.D + d] * inv_norm[j];
      di[d] = kc * inv_norm[i] * (acc - self * vi[d] * inv_norm[i]);
    }
  }
}

static int tl_check(const TlParams& p) {
  DV_REQUIRE(p.head && p.attn && p.maskp && p.slots && p.target && p.teacher && p.var && p.fg && p.fgf, "null pointer");
  DV_REQUIRE(p.B > 0 && p.S >= 2 && p.S <= kTlMaxS && p.C > 0 && p.W > p.C && p.H > 0 && p.N > 0 && p.P > 0 && p.D > 0,
             "bad sizes (2..8 slots)");
  return DEVIAS_OK;
}

}  // namespace dv

extern "C" int devias_train_loss_fwd(const float* head, const float* attn, const float* maskp, const float* slots,
                                     const int64_t* target, const float* teacher, const float* var_scalar, const float* fg,
                                     const float* fgf, int batch, int slots_per_clip, int width, int n_action, int heads,
                                     int n_tokens, int n_patches, int dim, int scene_ce, float w_scene, float w_maskpred,
                                     float w_distill, float* out6, int64_t* slot_idx, void* stream) {
  using namespace dv;
  TlParams p{head, attn, maskp, slots, reinterpret_cast<const long long*>(target), teacher, var_scalar, fg, fgf, batch, slots_per_clip,
             width, n_action, heads, n_tokens, n_patches, dim, scene_ce, w_scene, w_maskpred, w_distill};
  int rc = tl_check(p);
  if (rc) return rc;
  DV_REQUIRE(out6 && slot_idx, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DV_CHECK_CUDA(cudaMemsetAsync(out6, 0, 6 * sizeof(float), s));
  DV_CHECK_CUDA(launch_k(train_loss_fwd_kernel, dim3(batch), dim3(kTlThreads), (size_t)0, s, p, out6, reinterpret_cast<long long*>(slot_idx)));
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_train_loss_bwd(const float* head, const float* attn, const float* maskp, const float* slots,
                                     const int64_t* target, const float* teacher, const float* var_scalar, const float* fg,
                                     const float* fgf, int batch, int slots_per_clip, int width, int n_action, int heads,
                                     int n_tokens, int n_patches, int dim, int scene_ce, float w_scene, float w_maskpred,
                                     float w_distill, const float* grad_total, float* dhead, float* dattn, float* dmaskp,
                                     float* dslots, void* stream) {
  using namespace dv;
  TlParams p{head, attn, maskp, slots, reinterpret_cast<const long long*>(target), teacher, var_scalar, fg, fgf, batch, slots_per_clip,
             width, n_action, heads, n_tokens, n_patches, dim, scene_ce, w_scene, w_maskpred, w_distill};
  int rc = tl_check(p);
  if (rc) return rc;
  DV_REQUIRE(grad_total && dhead && dattn && dmaskp && dslots, "null pointer");
  DV_CHECK_CUDA(launch_k(train_loss_bwd_kernel, dim3(batch), dim3(kTlThreads), (size_t)0, static_cast<cudaStream_t>(stream), p, grad_total,
                         dhead, dattn, dmaskp, dslots));
  count_launch();
  return DEVIAS_OK;
}
