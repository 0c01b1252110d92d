// The O(S) algebra around the streaming slot-attention kernel (devias_b200/slot_attention.py), fused: per layer and direction
// about thirty tiny torch launches (scale, gamma product, row sums, dot with beta; the context expression and its quotient rule)
// become one kernel each.  Rows are the B*4*S (clip, head, slot) vectors of 768 channels.
//
//   fold : g = s qt * gamma,  G = sum_c g,  c0 = s sum_c qt beta                      (qt = Wk_h^T q, agg_block/attention.py:121-131 folded)
//   ctx  : cbar = (gamma * (U - m) + beta A) / (A + eps)                               (the token-axis renormalised context, :132-136)
//
// One CTA = kRows rows x 768 channels (thread <-> channels t, t+256, t+512): row reductions by warp shuffle + shared memory,
// column reductions (d gamma, d beta) in registers over the CTA's rows, then one atomicAdd per channel and CTA.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kGD = 768;
constexpr int kGT = 256;     // threads
constexpr int kGRows = 4;    // rows per CTA

// sum of `v` over the 256 threads of the CTA, result in every thread (slot: one of two shared arrays, alternated by the caller)
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kGT / 32; ++i) s += red[i];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(kGT) slot_fold_fwd_kernel(const float* __restrict__ qt, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float scale, float* __restrict__ g,
                                                            float* __restrict__ G, float* __restrict__ c0, int rows) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kGT / 32];
  const int t = threadIdx.x;
  float gm[3], bt[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { gm[k] = __ldg(gamma + t + kGT * k); bt[k] = __ldg(beta + t + kGT * k); }
  for (int r = blockIdx.x * kGRows; r < min(rows, (blockIdx.x + 1) * kGRows); ++r) {
    float sg = 0.f, sc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float q = scale * __ldg(qt + (long long)r * kGD + t + kGT * k);
      const float gv = q * gm[k];
      g[(long long)r * kGD + t + kGT * k] = gv;
      sg += gv;
      sc = fmaf(q, bt[k], sc);
    }
    sg = block_sum(sg, red);
    sc = block_sum(sc, red);
    if (t == 0) { G[r] = sg; c0[r] = sc; }
  }
}

__global__ void __launch_bounds__(kGT) slot_fold_bwd_kernel(const float* __restrict__ qt, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float scale, const float* __restrict__ dg,
                                                            const float* __restrict__ dG, const float* __restrict__ dc0,
                                                            float* __restrict__ dqt, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int rows) {
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x;
  float gm[3], bt[3], ag[3] = {0.f, 0.f, 0.f}, ab[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) { gm[k] = __ldg(gamma + t + kGT * k); bt[k] = __ldg(beta + t + kGT * k); }
  for (int r = blockIdx.x * kGRows; r < min(rows, (blockIdx.x + 1) * kGRows); ++r) {
    const float dGr = __ldg(dG + r), dcr = __ldg(dc0 + r);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long i = (long long)r * kGD + t + kGT * k;
      const float u = __ldg(dg + i) + dGr;                 // d g[r, c] including the row-sum path
      const float q = scale * __ldg(qt + i);
      dqt[i] = scale * fmaf(gm[k], u, bt[k] * dcr);
      ag[k] = fmaf(q, u, ag[k]);
      ab[k] = fmaf(q, dcr, ab[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    atomicAdd(dgamma + t + kGT * k, ag[k]);
    atomicAdd(dbeta + t + kGT * k, ab[k]);
  }
}

__global__ void __launch_bounds__(kGT) slot_ctx_fwd_kernel(const float* __restrict__ U, const float* __restrict__ m,
                                                           const float* __restrict__ A, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float* __restrict__ cbar, int rows) {
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x;
  float gm[3], bt[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { gm[k] = __ldg(gamma + t + kGT * k); bt[k] = __ldg(beta + t + kGT * k); }
  for (int r = blockIdx.x * kGRows; r < min(rows, (blockIdx.x + 1) * kGRows); ++r) {
    const float mr = __ldg(m + r), Ar = __ldg(A + r);
    const float inv = 1.0f / (Ar + eps);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long i = (long long)r * kGD + t + kGT * k;
      cbar[i] = fmaf(gm[k], __ldg(U + i) - mr, bt[k] * Ar) * inv;
    }
  }
}

__global__ void __launch_bounds__(kGT) slot_ctx_bwd_kernel(const float* __restrict__ dcbar, const float* __restrict__ U,
                                                           const float* __restrict__ m, const float* __restrict__ A,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                           float* __restrict__ dU, float* __restrict__ dm, float* __restrict__ dA,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int rows) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kGT / 32];
  const int t = threadIdx.x;
  float gm[3], bt[3], ag[3] = {0.f, 0.f, 0.f}, ab[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) { gm[k] = __ldg(gamma + t + kGT * k); bt[k] = __ldg(beta + t + kGT * k); }
  for (int r = blockIdx.x * kGRows; r < min(rows, (blockIdx.x + 1) * kGRows); ++r) {
    const float mr = __ldg(m + r), Ar = __ldg(A + r);
    const float inv = 1.0f / (Ar + eps);
    float s_gd = 0.f, s_bd = 0.f, s_nc = 0.f;               // sum gamma dnum, sum beta dnum, sum num dcbar
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long i = (long long)r * kGD + t + kGT * k;
      const float dc = __ldg(dcbar + i), um = __ldg(U + i) - mr;
      const float dn = dc * inv;                            // d numerator
      const float num = fmaf(gm[k], um, bt[k] * Ar);
      dU[i] = gm[k] * dn;
      s_gd = fmaf(gm[k], dn, s_gd);
      s_bd = fmaf(bt[k], dn, s_bd);
      s_nc = fmaf(num, dc, s_nc);
      ag[k] = fmaf(um, dn, ag[k]);
      ab[k] = fmaf(Ar, dn, ab[k]);
    }
    s_gd = block_sum(s_gd, red);
    s_bd = block_sum(s_bd, red);
    s_nc = block_sum(s_nc, red);
    if (t == 0) {
      dm[r] = -s_gd;
      dA[r] = s_bd - s_nc * inv * inv;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    atomicAdd(dgamma + t + kGT * k, ag[k]);
    atomicAdd(dbeta + t + kGT * k, ab[k]);
  }
}

static inline int glue_grid(int rows) { return (rows + kGRows - 1) / kGRows; }

}  // namespace dv

using namespace dv;

extern "C" int devias_slot_fold_fwd(const float* qt, const float* gamma, const float* beta, float scale, float* g, float* G, float* c0,
                                    int rows, int dim, void* stream) {
  DV_REQUIRE(qt && gamma && beta && g && G && c0, "null pointer");
  DV_REQUIRE(dim == kGD && rows > 0, "dim must be 768");
  DV_CHECK_CUDA(launch_k(slot_fold_fwd_kernel, dim3((unsigned)(glue_grid(rows))), dim3((unsigned)(kGT)), (size_t)(0), (cudaStream_t)stream, qt, gamma, beta, scale, g, G, c0, rows));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_slot_fold_bwd(const float* qt, const float* gamma, const float* beta, float scale, const float* dg,
                                    const float* dG, const float* dc0, float* dqt, float* dgamma, float* dbeta, int rows, int dim,
                                    void* stream) {
  DV_REQUIRE(qt && gamma && beta && dg && dG && dc0 && dqt && dgamma && dbeta, "null pointer");
  DV_REQUIRE(dim == kGD && rows > 0, "dim must be 768");
  DV_CHECK_CUDA(launch_k(slot_fold_bwd_kernel, dim3((unsigned)(glue_grid(rows))), dim3((unsigned)(kGT)), (size_t)(0), (cudaStream_t)stream, qt, gamma, beta, scale, dg, dG, dc0, dqt, dgamma, dbeta, rows));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_slot_ctx_fwd(const float* U, const float* m, const float* A, const float* gamma, const float* beta, float eps,
                                   float* cbar, int rows, int dim, void* stream) {
  DV_REQUIRE(U && m && A && gamma && beta && cbar, "null pointer");
  DV_REQUIRE(dim == kGD && rows > 0, "dim must be 768");
  DV_CHECK_CUDA(launch_k(slot_ctx_fwd_kernel, dim3((unsigned)(glue_grid(rows))), dim3((unsigned)(kGT)), (size_t)(0), (cudaStream_t)stream, U, m, A, gamma, beta, eps, cbar, rows));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_slot_ctx_bwd(const float* dcbar, const float* U, const float* m, const float* A, const float* gamma,
                                   const float* beta, float eps, float* dU, float* dm, float* dA, float* dgamma, float* dbeta, int rows,
                                   int dim, void* stream) {
  DV_REQUIRE(dcbar && U && m && A && gamma && beta && dU && dm && dA && dgamma && dbeta, "null pointer");
  DV_REQUIRE(dim == kGD && rows > 0, "dim must be 768");
  DV_CHECK_CUDA(launch_k(slot_ctx_bwd_kernel, dim3((unsigned)(glue_grid(rows))), dim3((unsigned)(kGT)), (size_t)(0), (cudaStream_t)stream, dcbar, U, m, A, gamma, beta, eps, dU, dm, dA, dgamma, dbeta, rows));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
