// Small memory-bound helpers around the GEMMs: fp32 -> bf16 weight cast and the tube-patch gather that turns the
// Conv3d(k=s=(2,16,16)) of model/modeling_slot.py:167-176 into a plain GEMM operand.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dv {

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                            long long n8, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8)) out[n8 * 8 + threadIdx.x] = __float2bfloat16(in[n8 * 8 + threadIdx.x]);
}

// out_bf16[r, :] = bf16(in[r, :] * row_scale[r / rows_per_scale])  (drop-path scaled gradient operand)
__global__ void __launch_bounds__(256) scale_rows_cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                              long long n8, int cols8, const float* __restrict__ row_scale,
                                                              int rows_per_scale) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float s = __ldg(row_scale + (i / cols8) / rows_per_scale);
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16(a.x * s, a.y * s), pack_bf16(a.z * s, a.w * s),
                                                  pack_bf16(b.x * s, b.y * s), pack_bf16(b.z * s, b.w * s));
  }
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(p)), b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
  const float2 f0 = unpack_bf16(a.x), f1 = unpack_bf16(a.y), f2 = unpack_bf16(a.z), f3 = unpack_bf16(a.w);
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// clip [B, C, T, H, W] -> patch matrix bf16 [B * (T/2)*(H/16)*(W/16), C*2*16*16];
// row = token (t*196 + h*14 + w for 224^2), col = c*512 + dt*256 + dy*16 + dx  
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ clip, __nv_bfloat16* __restrict__ out, int B, int C,
                                                       int F, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int segs = W / 8;
  const long long total = (long long)B * C * F * H * segs;
  const int hp = H / 16, wp = W / 16, tp = F / 2;
  const int Kdim = C * 512;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int seg = (int)(i % segs);
    long long r = i / segs;
    const int y = (int)(r % H); r /= H;
    const int f = (int)(r % F); r /= F;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    float v[8];
    load8<T>(clip + i * 8, v);
    const int x0 = seg * 8;
    const long long row = (((long long)b * tp + (f >> 1)) * hp + (y >> 4)) * wp + (x0 >> 4);
    const int col = c * 512 + (f & 1) * 256 + (y & 15) * 16 + (x0 & 15);
    *reinterpret_cast<uint4*>(out + row * Kdim + col) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}


// bf16 -> fp32, 8 elements per thread (the fp32 slot-attention backward consumes an upcast copy of bf16 tokens)
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                                            long long n8, long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i * 8);
    const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
    *reinterpret_cast<float4*>(out + i * 8) = make_float4(a.x, a.y, b.x, b.y);
    *reinterpret_cast<float4*>(out + i * 8 + 4) = make_float4(c.x, c.y, d.x, d.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n8 * 8; i < n; ++i) out[i] = __bfloat162float(in[i]);
}

}  // namespace dv

extern "C" int devias_cast_bf16_f32(const void* in, float* out, int64_t n, void* stream) {
  using namespace dv;
  DV_REQUIRE(in && out, "null pointer");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "16-byte alignment");
  if (n <= 0) return DEVIAS_OK;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (blocks < 1) blocks = 1;
  DV_CHECK_CUDA(launch_k(cast_bf16_f32_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(in), out, n8, (long long)n));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  using namespace dv;
  DV_REQUIRE(in && out, "null pointer");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "16-byte alignment");
  if (n <= 0) return DEVIAS_OK;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (blocks < 1) blocks = 1;
  DV_CHECK_CUDA(launch_k(cast_f32_bf16_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), static_cast<cudaStream_t>(stream), in, static_cast<__nv_bfloat16*>(out), n8, n));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_scale_rows_cast(const float* in, void* out, int rows, int cols, const float* row_scale,
                                      int rows_per_scale, void* stream) {
  using namespace dv;
  DV_REQUIRE(in && out && row_scale && rows_per_scale > 0, "null pointer");
  DV_REQUIRE(cols % 8 == 0, "cols must be a multiple of 8");
  if (rows <= 0) return DEVIAS_OK;
  const long long n8 = (long long)rows * cols / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  DV_CHECK_CUDA(launch_k(scale_rows_cast_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), static_cast<cudaStream_t>(stream), in, static_cast<__nv_bfloat16*>(out), n8,
                                                                                   cols / 8, row_scale, rows_per_scale));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_patchify(const void* clip, int clip_dtype, void* out, int batch, int chans, int frames, int height,
                               int width, void* stream) {
  using namespace dv;
  DV_REQUIRE(clip && out, "null pointer");
  DV_REQUIRE(frames % 2 == 0 && height % 16 == 0 && width % 16 == 0, "clip must tile into 2x16x16 tubes");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(clip) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "16-byte alignment");
  if (batch <= 0) return DEVIAS_OK;
  const long long total = (long long)batch * chans * frames * height * (width / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > sm_count() * 32) blocks = sm_count() * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (clip_dtype) {
    case DEVIAS_DTYPE_F32: DV_CHECK_CUDA(launch_k(patchify_kernel<float>, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), s, static_cast<const float*>(clip), o, batch, chans, frames, height, width)); break;
    case DEVIAS_DTYPE_BF16: DV_CHECK_CUDA(launch_k(patchify_kernel<__nv_bfloat16>, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), s, static_cast<const __nv_bfloat16*>(clip), o, batch, chans, frames, height, width)); break;
    case DEVIAS_DTYPE_F16: DV_CHECK_CUDA(launch_k(patchify_kernel<__half>, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), s, static_cast<const __half*>(clip), o, batch, chans, frames, height, width)); break;
    default: set_last_error("clip_dtype", "unknown dtype id", __FILE__, __LINE__); return DEVIAS_ERR_ARG;
  }
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
