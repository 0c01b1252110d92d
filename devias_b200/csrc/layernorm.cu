// LayerNorm over the channel dim (D = 768 on this path), forward and backward, memory-bound streaming kernels.
//   forward : x fp32 [M,D] -> y (bf16 or fp32) [M,D], mean/rstd fp32 [M]         (model/modeling_slot.py:126,132,373)
//   backward: dx = d_resid + LN'(dy)  (fp32) [+ bf16 copy], d_gamma/d_beta += column sums, colsum(dx) -> bias grad of the
//             residual-producing linear layer (Block.forward's x = x + f(LN(x)), model/modeling_slot.py:150-151)
// One warp per row, the whole row in registers (two-pass statistics: exact mean, then centred variance).
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

template <int D>
struct RowRegs {
  static constexpr int V = D / 128;  // float4 per lane
};

template <int D, bool OUT_BF16>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, void* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int M,
                                                            float eps) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = RowRegs<D>::V;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = __ldcs(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float r = rsqrtf(q * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = r;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    const float o0 = fmaf((v[i].x - mu) * r, g.x, b.x), o1 = fmaf((v[i].y - mu) * r, g.y, b.y);
    const float o2 = fmaf((v[i].z - mu) * r, g.z, b.z), o3 = fmaf((v[i].w - mu) * r, g.w, b.w);
    if constexpr (OUT_BF16) {
      uint2* yr = reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + (long long)row * D);
      yr[lane + 32 * i] = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
    } else {
      float4* yr = reinterpret_cast<float4*>(static_cast<float*>(y) + (long long)row * D);
      yr[lane + 32 * i] = make_float4(o0, o1, o2, o3);
    }
  }
}

// Backward.  dy: gradient w.r.t. the LN output (bf16 or fp32).  d_resid (optional, fp32) is added to the result
// (the skip connection); dx may alias d_resid.  The three column reductions (dgamma, dbeta, colsum(dx)) are kept in
// per-warp shared-memory rows instead of registers (which held occupancy to one block per SM and the kernel to ~40 % of
// HBM bandwidth); they are combined across the block's warps at the end and flushed with one atomicAdd per column.
// All three streams of a row (x, dy, d_resid: 7.5 KiB per warp) are requested up front, so a row costs ONE memory round trip;
// with d_resid fetched after the row reductions (and three blocks of 85 registers per SM) the kernel sat at 0.69 of the HBM
// peak, with two blocks of 128 registers and the early fetch it runs at 0.96 (B = 32: 137 -> 98 us).
constexpr int kLnBwdBlocksPerSM = 2;
template <int D, bool DY_BF16>
__global__ void __launch_bounds__(256, kLnBwdBlocksPerSM) layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* d_resid,
                                                               float* dx, __nv_bfloat16* __restrict__ dx_bf16,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               float* __restrict__ dx_colsum,
                                                               const float* __restrict__ bscale, int rows_per_scale, int M) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = RowRegs<D>::V;
  extern __shared__ float4 ln_acc[];                 // [3][8 warps][D/4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float4* accg = ln_acc + (0 * 8 + warp) * (D / 4);
  float4* accb = ln_acc + (1 * 8 + warp) * (D / 4);
  float4* accc = ln_acc + (2 * 8 + warp) * (D / 4);
#pragma unroll
  for (int i = 0; i < V; ++i) accg[lane + 32 * i] = accb[lane + 32 * i] = accc[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
    const float mu = __ldg(mean + row), r = __ldg(rstd + row);
    // per-sample factor of the branch that consumes this gradient next (drop-path): applied to the bf16 copy and its column sums
    const float sc = bscale != nullptr ? __ldg(bscale + row / rows_per_scale) : 1.0f;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    float4 xh[V], d[V];
    float4 rr[V];
#pragma unroll
    for (int i = 0; i < V; ++i)
      rr[i] = d_resid != nullptr ? __ldcs(reinterpret_cast<const float4*>(d_resid + (long long)row * D) + lane + 32 * i)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 xv = __ldcs(xr + lane + 32 * i);
      xh[i] = make_float4((xv.x - mu) * r, (xv.y - mu) * r, (xv.z - mu) * r, (xv.w - mu) * r);
      if constexpr (DY_BF16) {
        const uint2 p = __ldcs(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy) + (long long)row * D) + lane + 32 * i);
        const float2 a = unpack_bf16(p.x), b = unpack_bf16(p.y);
        d[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d[i] = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(dy) + (long long)row * D) + lane + 32 * i);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 ab = accb[lane + 32 * i], ag = accg[lane + 32 * i];
      ab.x += d[i].x; ab.y += d[i].y; ab.z += d[i].z; ab.w += d[i].w;
      ag.x = fmaf(d[i].x, xh[i].x, ag.x); ag.y = fmaf(d[i].y, xh[i].y, ag.y);
      ag.z = fmaf(d[i].z, xh[i].z, ag.z); ag.w = fmaf(d[i].w, xh[i].w, ag.w);
      accb[lane + 32 * i] = ab;
      accg[lane + 32 * i] = ag;
      const float4 g = __ldg(g4 + lane + 32 * i);
      d[i].x *= g.x; d[i].y *= g.y; d[i].z *= g.z; d[i].w *= g.w;  // dy * gamma
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / D);
    s2 *= (1.0f / D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = r * (d[i].x - s1 - xh[i].x * s2); o.y = r * (d[i].y - s1 - xh[i].y * s2);
      o.z = r * (d[i].z - s1 - xh[i].z * s2); o.w = r * (d[i].w - s1 - xh[i].w * s2);
      o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w;
      if (dx != nullptr) reinterpret_cast<float4*>(dx + (long long)row * D)[lane + 32 * i] = o;
      o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;
      if (dx_bf16 != nullptr)
        reinterpret_cast<uint2*>(dx_bf16 + (long long)row * D)[lane + 32 * i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      if (dx_colsum != nullptr) {
        float4 ac = accc[lane + 32 * i];
        ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
        accc[lane + 32 * i] = ac;
      }
    }
  }
  __syncthreads();
  const float* accf = reinterpret_cast<const float*>(ln_acc);
  auto flush = [&](int which, float* __restrict__ gout) {
    if (gout == nullptr) return;  // uniform across the block
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += accf[(which * 8 + w) * D + c];
      atomicAdd(gout + c, s);
    }
  };
  flush(0, dgamma);
  flush(1, dbeta);
  flush(2, dx_colsum);
}

// column sums of a bf16 [M,N] matrix into fp32 [N] (bias gradients of qkv / fc1): out[n] += sum_m a[m,n]
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long lda, int M, int N,
                                                          float* __restrict__ out, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  // thread handles 8 consecutive columns (one 16-byte load per row); block covers 32*8 = 256 columns x rows_per_block rows
  const int colgrp = blockIdx.x * 32 + (threadIdx.x & 31);
  const int col0 = colgrp * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col0 < N) {
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) {
      const uint4 p = __ldcs(reinterpret_cast<const uint4*>(a + (long long)r * lda + col0));
      const float2 a0 = unpack_bf16(p.x), a1 = unpack_bf16(p.y), a2 = unpack_bf16(p.z), a3 = unpack_bf16(p.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
  }
  __shared__ float red[8][32][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x >> 5][threadIdx.x & 31][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 256) {
    const int cg = threadIdx.x >> 3, i = threadIdx.x & 7;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][cg][i];
    const int col = (blockIdx.x * 32 + cg) * 8 + i;
    if (col < N) atomicAdd(out + col, s);
  }
}

}  // namespace dv

extern "C" int devias_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_is_bf16,
                                    float* mean, float* rstd, int rows, int dim, float eps, void* stream) {
  using namespace dv;
  DV_REQUIRE(x && gamma && beta && y, "null pointer");
  DV_REQUIRE(dim == 768, "only dim = 768 is instantiated (ViT-B / DEVIAS slot dim)");
  if (rows <= 0) return DEVIAS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (rows + 7) / 8;
  if (y_is_bf16) DV_CHECK_CUDA(launch_k(layernorm_fwd_kernel<768, true>, dim3(grid), dim3((unsigned)(256)), (size_t)(0), s, x, gamma, beta, y, mean, rstd, rows, eps));
  else DV_CHECK_CUDA(launch_k(layernorm_fwd_kernel<768, false>, dim3(grid), dim3((unsigned)(256)), (size_t)(0), s, x, gamma, beta, y, mean, rstd, rows, eps));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                    const float* gamma, const float* d_resid, float* dx, void* dx_bf16, float* dgamma,
                                    float* dbeta, float* dx_colsum, const float* bf16_row_scale, int rows_per_scale, int rows, int dim,
                                    void* stream) {
  using namespace dv;
  DV_REQUIRE(dy && x && mean && rstd && gamma, "null pointer");
  DV_REQUIRE(bf16_row_scale == nullptr || rows_per_scale > 0, "rows_per_scale");
  if (bf16_row_scale == nullptr) rows_per_scale = 1;
  DV_REQUIRE(dim == 768, "only dim = 768 is instantiated");
  if (rows <= 0) return DEVIAS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = sm_count() * kLnBwdBlocksPerSM;
  const int need = (rows + 7) / 8;
  if (grid > need) grid = need;
  constexpr int kLnSmem = 3 * 8 * 768 * 4;   // 72 KiB: three per-warp accumulator rows
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<768, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmem));
    DV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<768, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmem));
    attr_done = true;
  }
  if (dy_is_bf16)
    DV_CHECK_CUDA(launch_k(layernorm_bwd_kernel<768, true>, dim3(grid), dim3((unsigned)(256)), (size_t)(kLnSmem), s, dy, x, mean, rstd, gamma, d_resid, dx,
                                                         static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dx_colsum, bf16_row_scale,
                                                         rows_per_scale, rows));
  else
    DV_CHECK_CUDA(launch_k(layernorm_bwd_kernel<768, false>, dim3(grid), dim3((unsigned)(256)), (size_t)(kLnSmem), s, dy, x, mean, rstd, gamma, d_resid, dx,
                                                          static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dx_colsum, bf16_row_scale,
                                                         rows_per_scale, rows));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_colsum_bf16(const void* a, int64_t lda, int rows, int cols, float* out, void* stream) {
  using namespace dv;
  DV_REQUIRE(a && out, "null pointer");
  DV_REQUIRE(cols % 8 == 0 && lda % 8 == 0, "cols and lda must be multiples of 8");
  if (rows <= 0) return DEVIAS_OK;
  const int rows_per_block = 128;   // 768 x 12544 -> 294 blocks (>= 1 per SM), 16 rows in flight per thread
  dim3 grid((cols + 255) / 256, (rows + rows_per_block - 1) / rows_per_block);
  DV_CHECK_CUDA(launch_k(colsum_bf16_kernel, dim3(grid), dim3((unsigned)(256)), (size_t)(0), static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(a), lda, rows, cols,
                                                                         out, rows_per_block));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
