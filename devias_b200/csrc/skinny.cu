// Slot-side linear algebra (fp32): products whose row count is M = clips x slots (a few dozen) against the aggregation
// block's and the heads' weights -- agg_block/attention.py:120-141 (to_q, the folded to_k / to_v, to_out), :81-82
// (FeedForward), model/modeling_slot.py:390-410 (head, mask predictor) and their gradients.  With so few rows every
// product is bound by reading (or, for the weight gradients, writing) the weight matrix once, so the three kernels
// below are organised around one coalesced pass over the weights; the M rows live in shared memory / registers.
//
//   nt    : y[m, n]  = sum_k x[m, k] w[n, k] (+ bias[n])          weights [N, K]  (nn.Linear forward)
//   nn    : y[m, n] += sum_k x[m, k] w[k, n]                      weights [K, N]  (input gradient; folded key projection)
//   outer : c[i, j]  = sum_m a[m, i] b[m, j],  colsum[i] = sum_m a[m, i]          (weight / bias gradients)
//
// Rows of x / y / a / b are addressed through a two-level map  off(m) = (m / inner) * outer + (m % inner) * ld  plus a
// per-problem (head) offset, which expresses the '(b s) (h d)' <-> 'b h s c' views of the folded attention without copies.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

struct RowMap {
  long long outer, ld, batch;
  int inner;
  __device__ __forceinline__ long long off(int m, int z) const {
    return (long long)(m / inner) * outer + (long long)(m % inner) * ld + (long long)z * batch;
  }
};

constexpr int kSkMT = 16;          // rows per pass
constexpr int kSkKC = 1024;        // K chunk held in shared memory (64 KiB)
constexpr int kSkWR = kSkKC / 128; // weight float4s per lane, column and chunk

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;    // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------- nt
// One warp <-> two weight rows (output columns); the lanes stride K with 16-byte loads.  Per K chunk the warp's weight
// slice goes to registers and the x tile to shared memory (cp.async), both in flight together: one L2 round trip per chunk.
__global__ void __launch_bounds__(256) skinny_nt_kernel(const float* __restrict__ x, RowMap xm, const float* __restrict__ w,
                                                        long long w_batch, const float* __restrict__ bias, float* __restrict__ y,
                                                        RowMap ym, int M, int N, int K) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float4 xs4[];                                  // [16][kc / 4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x;
  const int z = blockIdx.z, m0 = blockIdx.y * kSkMT;
  const int n0 = blockIdx.x * (nthreads >> 4) + warp * 2;
  const float* wz = w + (long long)z * w_batch;
  const float4* w0 = reinterpret_cast<const float4*>(wz + (long long)min(n0, N - 1) * K);
  const float4* w1 = reinterpret_cast<const float4*>(wz + (long long)min(n0 + 1, N - 1) * K);
  const uint32_t xs_u = smem_u32(xs4);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kSkKC) {
    const int kc4 = min(kSkKC, K - k0) >> 2;
    __syncthreads();                                               // the previous chunk's x tile is no longer read
    float4 wa[kSkWR], wb[kSkWR];
#pragma unroll
    for (int q = 0; q < kSkWR; ++q) {
      const int c = lane + 32 * q;
      const bool ok = c < kc4;
      wa[q] = ok ? __ldg(w0 + (k0 >> 2) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      wb[q] = ok ? __ldg(w1 + (k0 >> 2) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int m = warp; m < kSkMT; m += (nthreads >> 5)) {
      const bool ok = m0 + m < M;
      const float4* src = reinterpret_cast<const float4*>(x + xm.off(ok ? m0 + m : m0, z) + k0);
      for (int c = lane; c < kc4; c += 32) cp_async16(xs_u + (m * kc4 + c) * 16, src + c, ok);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSkWR; ++q) {
      const int c = lane + 32 * q;
      if (c < kc4) {
#pragma unroll
        for (int m = 0; m < kSkMT; ++m) {
          const float4 v = xs4[m * kc4 + c];
          acc[m] = fmaf(v.x, wa[q].x, fmaf(v.y, wa[q].y, fmaf(v.z, wa[q].z, fmaf(v.w, wa[q].w, acc[m]))));
          acc[16 + m] = fmaf(v.x, wb[q].x, fmaf(v.y, wb[q].y, fmaf(v.z, wb[q].z, fmaf(v.w, wb[q].w, acc[16 + m]))));
        }
      }
    }
  }
  // butterfly: 31 shuffles leave in lane l the warp total of acc[l]  (l = column * 16 + row)
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float lo = acc[i], hi = acc[i + n / 2];
      const float other = __shfl_xor_sync(0xffffffffu, up ? lo : hi, off);
      acc[i] = (up ? hi : lo) + other;
    }
  }
  const int m = m0 + (lane & 15), n = n0 + (lane >> 4);
  if (m < M && n < N) y[ym.off(m, z) + n] = acc[0] + (bias != nullptr ? __ldg(bias + n) : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------- nn
constexpr int kSkNnThreads = 128;  // one output column per thread
constexpr int kSkNnKS = 64;        // k rows per CTA

__global__ void __launch_bounds__(kSkNnThreads) skinny_nn_kernel(const float* __restrict__ x, RowMap xm, const float* __restrict__ w,
                                                                 long long w_batch, float* __restrict__ y, RowMap ym, int M,
                                                                 int N, int K, int mtiles) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 xs4[kSkMT * kSkNnKS / 4];                      // [16][64 k], zero-padded (K may be any size: C + 365 logits)
  const int tid = threadIdx.x;
  const int z = blockIdx.z / mtiles, m0 = (blockIdx.z % mtiles) * kSkMT;
  const int k0 = blockIdx.y * kSkNnKS, ks = min(kSkNnKS, K - k0);
  const int n = blockIdx.x * kSkNnThreads + tid;
  const float* wp = w + (long long)z * w_batch + (long long)k0 * N + min(n, N - 1);
  float bw[kSkNnKS / 2];
#pragma unroll
  for (int q = 0; q < kSkNnKS / 2; ++q) bw[q] = (q < ks) ? __ldg(wp + (long long)q * N) : 0.f;    // first half: in flight with x
  {
    float xv[kSkMT * kSkNnKS / kSkNnThreads];
#pragma unroll
    for (int j = 0; j < kSkMT * kSkNnKS / kSkNnThreads; ++j) {
      const int i = tid + j * kSkNnThreads, m = i / kSkNnKS, k = i % kSkNnKS;
      xv[j] = (m0 + m < M && k < ks) ? __ldg(x + xm.off(m0 + m, z) + k0 + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kSkMT * kSkNnKS / kSkNnThreads; ++j) reinterpret_cast<float*>(xs4)[tid + j * kSkNnThreads] = xv[j];
  }
  __syncthreads();
  float acc[kSkMT];
#pragma unroll
  for (int m = 0; m < kSkMT; ++m) acc[m] = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float bn[kSkNnKS / 2];
    if (half == 0) {
#pragma unroll
      for (int q = 0; q < kSkNnKS / 2; ++q) bn[q] = (kSkNnKS / 2 + q < ks) ? __ldg(wp + (long long)(kSkNnKS / 2 + q) * N) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < kSkNnKS / 8; ++c) {
#pragma unroll
      for (int m = 0; m < kSkMT; ++m) {
        const float4 v = xs4[m * (kSkNnKS / 4) + half * (kSkNnKS / 8) + c];
        acc[m] = fmaf(v.x, bw[4 * c], fmaf(v.y, bw[4 * c + 1], fmaf(v.z, bw[4 * c + 2], fmaf(v.w, bw[4 * c + 3], acc[m]))));
      }
    }
    if (half == 0) {
#pragma unroll
      for (int q = 0; q < kSkNnKS / 2; ++q) bw[q] = bn[q];
    }
  }
  if (n >= N) return;
#pragma unroll
  for (int m = 0; m < kSkMT; ++m)
    if (m0 + m < M) atomicAdd(y + ym.off(m0 + m, z) + n, acc[m]);
}

// ------------------------------------------------------------------------------------------------------------- outer
constexpr int kSkOutRows = 8;      // rows i of c per CTA

__global__ void __launch_bounds__(256) skinny_outer_kernel(const float* __restrict__ a, RowMap am, const float* __restrict__ b, RowMap bm,
                                                           float* __restrict__ c, long long c_batch, float* __restrict__ colsum,
                                                           long long colsum_batch, int M, int I, int J, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 as4[kSkOutRows * kSkMT / 4];                   // [8 rows i][16 m]
  const int tid = threadIdx.x, z = blockIdx.z;
  const int i0 = blockIdx.y * kSkOutRows;
  const int j4 = blockIdx.x * blockDim.x + tid;                    // float4 column of b / c
  const bool active = 4 * j4 < J;
  float4 acc[kSkOutRows];
#pragma unroll
  for (int i = 0; i < kSkOutRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float cs = 0.f;
  for (int m0 = 0; m0 < M; m0 += kSkMT) {
    float4 bv[kSkMT];                                              // issued first: in flight together with the a tile
#pragma unroll
    for (int mm = 0; mm < kSkMT; ++mm)
      bv[mm] = (active && m0 + mm < M) ? __ldg(reinterpret_cast<const float4*>(b + bm.off(m0 + mm, z)) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int t = tid; t < kSkOutRows * kSkMT; t += blockDim.x) {
      const int i = t & (kSkOutRows - 1), m = t / kSkOutRows;      // consecutive threads <-> consecutive i: coalesced
      const float v = (m0 + m < M && i0 + i < I) ? __ldg(a + am.off(m0 + m, z) + i0 + i) : 0.f;
      reinterpret_cast<float*>(as4)[i * kSkMT + m] = v;
    }
    __syncthreads();
    if (colsum != nullptr && blockIdx.x == 0 && tid < kSkOutRows) {
#pragma unroll
      for (int q = 0; q < kSkMT / 4; ++q) {
        const float4 v = as4[tid * (kSkMT / 4) + q];
        cs += (v.x + v.y) + (v.z + v.w);
      }
    }
    if (active) {
#pragma unroll
      for (int q = 0; q < kSkMT / 4; ++q) {
#pragma unroll
        for (int i = 0; i < kSkOutRows; ++i) {
          const float4 av = as4[i * (kSkMT / 4) + q];
          const float4 b0 = bv[4 * q], b1 = bv[4 * q + 1], b2 = bv[4 * q + 2], b3 = bv[4 * q + 3];
          acc[i].x = fmaf(av.x, b0.x, fmaf(av.y, b1.x, fmaf(av.z, b2.x, fmaf(av.w, b3.x, acc[i].x))));
          acc[i].y = fmaf(av.x, b0.y, fmaf(av.y, b1.y, fmaf(av.z, b2.y, fmaf(av.w, b3.y, acc[i].y))));
          acc[i].z = fmaf(av.x, b0.z, fmaf(av.y, b1.z, fmaf(av.z, b2.z, fmaf(av.w, b3.z, acc[i].z))));
          acc[i].w = fmaf(av.x, b0.w, fmaf(av.y, b1.w, fmaf(av.z, b2.w, fmaf(av.w, b3.w, acc[i].w))));
        }
      }
    }
  }
  if (active) {
    float* cz = c + (long long)z * c_batch;
#pragma unroll
    for (int i = 0; i < kSkOutRows; ++i)
      if (i0 + i < I) {
        float4* dst = reinterpret_cast<float4*>(cz + (long long)(i0 + i) * J) + j4;
        if (accumulate) {       // gradient arena mode: every (i, j) is owned by exactly one thread, plain read-modify-write
          const float4 o = *dst;
          acc[i].x += o.x; acc[i].y += o.y; acc[i].z += o.z; acc[i].w += o.w;
        }
        *dst = acc[i];
      }
  }
  if (colsum != nullptr && blockIdx.x == 0 && tid < kSkOutRows && i0 + tid < I) {
    float* d = colsum + (long long)z * colsum_batch + i0 + tid;
    *d = accumulate ? *d + cs : cs;
  }
}

// ------------------------------------------------------------------------------------------------------- many rows (M > 32)
// K400 training has 64 slot rows per GPU, an evaluation batch of 256 has 512: there the products stop being weight-read bound
// and the row-tile-at-a-time kernels above re-read the weights per 16 rows.  Register-tiled fp32 SIMT GEMM instead:
//   MODE 0 (nt)   : C[i, j] = sum_l x[i, l] w[j, l]        i = row m (mapped), j = n          both operands contiguous along l
//   MODE 1 (nn)   : C[i, j] += sum_l x[i, l] w[l, j]                                        B contiguous along j
//   MODE 2 (outer): C[i, j] (+)= sum_l a[l, i] b[l, j]     l = row m (mapped)               both operands contiguous along i / j
// k-chunks staged in shared memory as [l][i] / [l][j], the next chunk's 16-byte loads in flight while the current one is
// multiplied, split over the contraction where the output has too few tiles (partial sums leave through
// red.global.add.v4.f32: a quarter of the L2 atomic operations of scalar atomics).  Two shapes of the same template, chosen
// per mode by measurement at 64 rows (a product is 0.2-0.3 GFLOP, i.e. ~4 us of the whole GPU's FFMA rate: latency and the
// shared-memory return path decide, not the FMA pipe):
//   R = 1: CTA tile 64 x 64, 256 threads x (4 x 4) outputs, chunks of 32.  Many small CTAs: best where the contraction is short
//          (weight gradients: 64 rows) or the output narrow (input gradients).
//   R = 2: CTA tile 64 x 128, 128 threads x (8 x 8) outputs, chunks of 16.  The shared-memory -> register return path (128 B per
//          clock and SM; an LDS.128 costs four of its cycles even when the lanes read the same 16 bytes) carries 2 bytes per FMA
//          with 4 x 4 outputs per thread -- twice the FFMA time, ncu: 3.3 of 4 warps per scheduler waiting on shared loads --
//          and 1 byte with 8 x 8.  Forward products (nt, both operands stored transposed): 22 -> 16 us at 768 -> 3072.
constexpr int kTgI = 64, kTgPad = 4, kTgSplitK = 32;   // rows per CTA tile; split granularity of the contraction

struct TileArgs {
  const float* A; const float* B; float* C;
  RowMap am, bm, cm;                 // row maps of the row-mapped operands / output (unused ones are ignored)
  long long a_batch, b_batch, c_batch;
  int I, J, L;                       // output rows, output columns, contraction length
  int ldb, ldc;                      // MODE 0: w row stride (= K); MODE 1: w row stride (= N); MODE 2: c row stride (= J)
  int l_per_split;                   // a multiple of kTgSplitK
  int accumulate;                    // MODE 2: c += ; (MODE 0 / 1 always add atomically onto a pre-filled output)
  float* colsum; long long colsum_batch;   // MODE 2: colsum[i] (+)= sum_l a[l, i]
};

__device__ __forceinline__ float4 ld4_guard(const float* p, int valid, bool vec) {
  if (vec && valid >= 4) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid > 0) v.x = __ldg(p);
  if (valid > 1) v.y = __ldg(p + 1);
  if (valid > 2) v.z = __ldg(p + 2);
  if (valid > 3) v.w = __ldg(p + 3);
  return v;
}
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int R>
struct TileShape {
  static constexpr int NT = 256 / R;         // threads: (16 / R) x 16
  static constexpr int TJ = 64 * R;          // columns per CTA tile
  static constexpr int KC = 32 / R;          // contraction chunk
  static constexpr int PA = 2, PB = 2 * R;   // 16-byte pieces per thread and chunk of the A / B operand
};

template <int MODE, int R>
__global__ void __launch_bounds__(TileShape<R>::NT, R == 1 ? 4 : 3) tile_gemm_kernel(const TileArgs p, int splits) {
  using T = TileShape<R>;
  constexpr int NT = T::NT, TJ = T::TJ, KC = T::KC, PA = T::PA, PB = T::PB, MR = 4 * R;   // MR x MR outputs per thread
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float As[KC][kTgI + kTgPad];
  __shared__ __align__(16) float Bs[KC][TJ + kTgPad];
  const int tid = threadIdx.x;
  const int z = blockIdx.z / splits, split = blockIdx.z % splits;
  const int i0 = blockIdx.y * kTgI, j0 = blockIdx.x * TJ;
  const int l_begin = split * p.l_per_split, l_end = min(p.L, l_begin + p.l_per_split);
  if (l_begin >= l_end) return;                       // (the launchers size the split count so that this does not happen)
  // thread (ti, tj): rows {(64 / R) ra + 4 ti + a}, columns {64 rb + 4 tj + b}   (ra, rb < R;  a, b < 4)
  const int ti = tid >> 4, tj = tid & 15;
  float acc[MR][MR];
#pragma unroll
  for (int a = 0; a < MR; ++a)
#pragma unroll
    for (int b = 0; b < MR; ++b) acc[a][b] = 0.f;

  // loader roles: the chunk of an operand is (rows x KC) floats = 16-byte pieces idx = tid + NT q
  //  "along l" (operand rows contiguous in l): row = idx / (KC / 4), l4 = (idx % (KC / 4)) * 4  (NT q keeps l4), stored transposed
  //  "along i/j" (operand contiguous across its rows): l = idx / (W / 4), c4 = (idx % (W / 4)) * 4  (W = 64 / TJ), stored as is
  constexpr int LQ = KC / 4;
  const int ll4 = (tid % LQ) * 4;
  const float* a_row[PA];
  const float* b_row[PB];
#pragma unroll
  for (int q = 0; q < PA; ++q) {
    const int r = i0 + (tid + NT * q) / LQ;
    a_row[q] = ((MODE == 0 || MODE == 1) && r < p.I) ? p.A + p.am.off(r, z) : nullptr;
  }
#pragma unroll
  for (int q = 0; q < PB; ++q) {
    const int r = j0 + (tid + NT * q) / LQ;
    b_row[q] = (MODE == 0 && r < p.J) ? p.B + (long long)z * p.b_batch + (long long)r * p.ldb : nullptr;
  }
  float csum = 0.f;                                   // MODE 2: column sums of a (thread tid < 64 <-> i = i0 + tid)
  auto fetch = [&](int l0, float4 (&av)[PA], float4 (&bv)[PB]) {   // this thread's pieces of the A and B chunk starting at l0
#pragma unroll
    for (int q = 0; q < PA; ++q) {
      av[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0 || MODE == 1) {
        const float* src = a_row[q] + l0 + ll4;
        if (a_row[q] != nullptr) av[q] = ld4_guard(src, l_end - (l0 + ll4), al16(src));
      } else {
        const int idx = tid + NT * q, l = l0 + idx / (kTgI / 4), c4 = (idx % (kTgI / 4)) * 4;
        if (l < l_end) {
          const float* src = p.A + p.am.off(l, z) + i0 + c4;
          av[q] = ld4_guard(src, p.I - (i0 + c4), al16(src));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      bv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        const float* src = b_row[q] + l0 + ll4;
        if (b_row[q] != nullptr) bv[q] = ld4_guard(src, l_end - (l0 + ll4), al16(src));
      } else {
        const int idx = tid + NT * q, l = l0 + idx / (TJ / 4), c4 = (idx % (TJ / 4)) * 4;
        if (l < l_end) {
          const float* src = (MODE == 1 ? p.B + (long long)z * p.b_batch + (long long)l * p.ldb : p.B + p.bm.off(l, z)) + j0 + c4;
          bv[q] = ld4_guard(src, p.J - (j0 + c4), al16(src));
        }
      }
    }
  };
  float4 av[PA], bv[PB];
  fetch(l_begin, av, bv);
  for (int l0 = l_begin; l0 < l_end; l0 += KC) {
    __syncthreads();                                  // the previous chunk has been consumed
#pragma unroll
    for (int q = 0; q < PA; ++q) {
      const int idx = tid + NT * q;
      if (MODE == 0 || MODE == 1) {
        const int r = idx / LQ;
        As[ll4][r] = av[q].x; As[ll4 + 1][r] = av[q].y; As[ll4 + 2][r] = av[q].z; As[ll4 + 3][r] = av[q].w;
      } else {
        *reinterpret_cast<float4*>(&As[idx / (kTgI / 4)][(idx % (kTgI / 4)) * 4]) = av[q];
      }
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int idx = tid + NT * q;
      if (MODE == 0) {
        const int r = idx / LQ;
        Bs[ll4][r] = bv[q].x; Bs[ll4 + 1][r] = bv[q].y; Bs[ll4 + 2][r] = bv[q].z; Bs[ll4 + 3][r] = bv[q].w;
      } else {
        *reinterpret_cast<float4*>(&Bs[idx / (TJ / 4)][(idx % (TJ / 4)) * 4]) = bv[q];
      }
    }
    __syncthreads();
    if (l0 + KC < l_end) fetch(l0 + KC, av, bv);      // next chunk in flight while this one is multiplied
#pragma unroll
    for (int l = 0; l < KC; ++l) {
      float aa[MR], bb[MR];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[l][(kTgI / R) * r + ti * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[l][64 * r + tj * 4]);
        aa[4 * r] = a4.x; aa[4 * r + 1] = a4.y; aa[4 * r + 2] = a4.z; aa[4 * r + 3] = a4.w;
        bb[4 * r] = b4.x; bb[4 * r + 1] = b4.y; bb[4 * r + 2] = b4.z; bb[4 * r + 3] = b4.w;
      }
#pragma unroll
      for (int a = 0; a < MR; ++a)
#pragma unroll
        for (int b = 0; b < MR; ++b) acc[a][b] = fmaf(aa[a], bb[b], acc[a][b]);
    }
    if (MODE == 2 && p.colsum != nullptr && blockIdx.x == 0 && tid < kTgI) {
#pragma unroll
      for (int l = 0; l < KC; ++l) csum += As[l][tid];
    }
  }
  // ---- output
#pragma unroll
  for (int a = 0; a < MR; ++a) {
    const int i = i0 + (a >> 2) * (kTgI / R) + ti * 4 + (a & 3);
    if (i >= p.I) continue;
    float* crow = (MODE == 2) ? p.C + (long long)z * p.c_batch + (long long)i * p.ldc : p.C + p.cm.off(i, z);
#pragma unroll
    for (int h = 0; h < R; ++h) {
      const int jt = j0 + h * 64 + tj * 4;
      const float* v = &acc[a][h * 4];
      const bool vec = jt + 3 < p.J && al16(crow + jt);
      if (MODE == 2 && !p.accumulate) {
        if (vec) {
          *reinterpret_cast<float4*>(crow + jt) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (jt + b < p.J) crow[jt + b] = v[b];
        }
      } else if (vec) {                               // (MODE 2: each output belongs to one CTA, so the reduction is a plain += )
        red_add_v4_f32(crow + jt, v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (jt + b < p.J) atomicAdd(crow + jt + b, v[b]);
      }
    }
  }
  if (MODE == 2 && p.colsum != nullptr && blockIdx.x == 0 && tid < kTgI && i0 + tid < p.I) {
    float* d = p.colsum + (long long)z * p.colsum_batch + i0 + tid;
    *d = p.accumulate ? *d + csum : csum;
  }
}

// y[m, :] = bias (or 0) through the row map: the pre-fill the split nt product adds onto
__global__ void __launch_bounds__(256) rows_fill_kernel(float* __restrict__ y, RowMap ym, const float* __restrict__ bias, int M, int N) {
  pdl_trigger();
  pdl_wait();
  const int z = blockIdx.z;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < M * N; idx += gridDim.x * blockDim.x) {
    const int m = idx / N, n = idx - m * N;
    y[ym.off(m, z) + n] = bias != nullptr ? __ldg(bias + n) : 0.f;
  }
}

static int skinny_max_rows() {       // up to here the weight-streaming kernels; above, the register-tiled ones
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEVIAS_SKINNY_MAX_ROWS"); v = e ? atoi(e) : 32; }
  return v;
}

// split of the contraction over CTAs: ~`per_sm` CTAs per SM, whole kTgSplitK pieces per split, no empty split
static void tile_split(int tiles, int L, int per_sm, int* splits, int* l_per_split) {
  const int chunks = (L + kTgSplitK - 1) / kTgSplitK;
  int want = (per_sm * sm_count() + tiles - 1) / tiles;
  if (want > chunks) want = chunks;
  if (want < 1) want = 1;
  const int per = (chunks + want - 1) / want;
  *splits = (chunks + per - 1) / per;
  *l_per_split = per * kTgSplitK;
}

template <int MODE, int R>
static cudaError_t launch_tile(TileArgs& t, int batch, bool split, cudaStream_t stream) {
  using T = TileShape<R>;
  const int gx = (t.J + T::TJ - 1) / T::TJ, gy = (t.I + kTgI - 1) / kTgI;
  int splits = 1;
  t.l_per_split = (t.L + kTgSplitK - 1) / kTgSplitK * kTgSplitK;
  if (split) tile_split(gx * gy * batch, t.L, R == 1 ? 2 : 3, &splits, &t.l_per_split);
  return launch_k(tile_gemm_kernel<MODE, R>, dim3(gx, gy, batch * splits), dim3(T::NT), (size_t)0, stream, t, splits);
}

static RowMap make_map(const int64_t* m) {
  RowMap r;
  r.outer = m[0]; r.inner = (int)m[1]; r.ld = m[2]; r.batch = m[3];
  return r;
}
static bool map_ok(const int64_t* m) { return m != nullptr && m[1] > 0 && m[0] % 4 == 0 && m[2] % 4 == 0 && m[3] % 4 == 0; }

}  // namespace dv

using namespace dv;

extern "C" int devias_skinny_nt(const float* x, const int64_t* x_map, const float* w, int64_t w_batch, const float* bias, float* y,
                                const int64_t* y_map, int M, int N, int K, int batch, void* stream) {
  DV_REQUIRE(x && w && y, "null pointer");
  DV_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && K % 4 == 0, "skinny_nt: K must be a multiple of 4");
  DV_REQUIRE(map_ok(x_map) && y_map && y_map[1] > 0 && w_batch % 4 == 0, "skinny_nt: row maps must keep 16-byte alignment");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) % 16 == 0, "skinny_nt: x / w must be 16-byte aligned");
  if (M > skinny_max_rows()) {
    TileArgs t{};
    t.A = x; t.B = w; t.C = y; t.am = make_map(x_map); t.cm = make_map(y_map); t.b_batch = w_batch;
    t.I = M; t.J = N; t.L = K; t.ldb = K;
    int fill_blocks = (M * N + 255) / 256;
    if (fill_blocks > 4 * sm_count()) fill_blocks = 4 * sm_count();
    DV_CHECK_CUDA(launch_k(rows_fill_kernel, dim3(fill_blocks, 1, batch), dim3(256), (size_t)0, (cudaStream_t)stream, y, t.cm, bias, M, N));
    DV_CHECK_CUDA((launch_tile<0, 2>(t, batch, true, (cudaStream_t)stream)));
    count_launch(2);
    return DEVIAS_OK;
  }
  static bool attr = false;
  if (!attr) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(skinny_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkMT * kSkKC * 4));
    attr = true;
  }
  // 16 output columns per CTA (8 warps) when that still fills the GPU, else 8 (4 warps)
  const int threads = ((long long)((N + 15) / 16) * ((M + kSkMT - 1) / kSkMT) * batch >= sm_count()) ? 256 : 128;
  const int cols = threads >> 4;
  const dim3 grid((N + cols - 1) / cols, (M + kSkMT - 1) / kSkMT, batch);
  const size_t smem = (size_t)kSkMT * (K < kSkKC ? K : kSkKC) * 4;
  DV_CHECK_CUDA(launch_k(skinny_nt_kernel, dim3(grid), dim3((unsigned)(threads)), (size_t)(smem), (cudaStream_t)stream, x, make_map(x_map), w, w_batch, bias, y, make_map(y_map), M, N, K));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_skinny_nn(const float* x, const int64_t* x_map, const float* w, int64_t w_batch, float* y, const int64_t* y_map,
                                int M, int N, int K, int batch, void* stream) {
  DV_REQUIRE(x && w && y, "null pointer");
  DV_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "skinny_nn: empty problem");
  DV_REQUIRE(x_map && x_map[1] > 0 && y_map && y_map[1] > 0, "skinny_nn: row maps need inner > 0");
  if (M > skinny_max_rows()) {
    TileArgs t{};
    t.A = x; t.B = w; t.C = y; t.am = make_map(x_map); t.cm = make_map(y_map); t.b_batch = w_batch;
    t.I = M; t.J = N; t.L = K; t.ldb = N;
    DV_CHECK_CUDA((launch_tile<1, 1>(t, batch, true, (cudaStream_t)stream)));
    count_launch();
    return DEVIAS_OK;
  }
  const int mtiles = (M + kSkMT - 1) / kSkMT;
  const dim3 grid((N + kSkNnThreads - 1) / kSkNnThreads, (K + kSkNnKS - 1) / kSkNnKS, batch * mtiles);
  DV_CHECK_CUDA(launch_k(skinny_nn_kernel, dim3(grid), dim3((unsigned)(kSkNnThreads)), (size_t)(0), (cudaStream_t)stream, x, make_map(x_map), w, w_batch, y, make_map(y_map), M, N, K, mtiles));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_skinny_outer(const float* a, const int64_t* a_map, const float* b, const int64_t* b_map, float* c,
                                   int64_t c_batch, float* colsum, int64_t colsum_batch, int M, int I, int J, int batch,
                                   int accumulate, void* stream) {
  DV_REQUIRE(a && b && c, "null pointer");
  DV_REQUIRE(M > 0 && I > 0 && J > 0 && batch > 0 && J % 4 == 0, "skinny_outer: J must be a multiple of 4");
  DV_REQUIRE(a_map && a_map[1] > 0 && map_ok(b_map) && c_batch % 4 == 0, "skinny_outer: row maps must keep 16-byte alignment");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) % 16 == 0, "skinny_outer: b / c must be 16-byte aligned");
  if (M > skinny_max_rows()) {
    TileArgs t{};
    t.A = a; t.B = b; t.C = c; t.am = make_map(a_map); t.bm = make_map(b_map); t.c_batch = c_batch;
    t.I = I; t.J = J; t.L = M; t.ldc = J; t.accumulate = accumulate;
    t.colsum = colsum; t.colsum_batch = colsum_batch;
    DV_CHECK_CUDA((launch_tile<2, 1>(t, batch, false, (cudaStream_t)stream)));   // each output belongs to one CTA: no split
    count_launch();
    return DEVIAS_OK;
  }
  const int j4 = J / 4;
  const int threads = j4 >= 256 ? 256 : (j4 + 31) / 32 * 32;
  const dim3 grid((j4 + threads - 1) / threads, (I + kSkOutRows - 1) / kSkOutRows, batch);
  DV_CHECK_CUDA(launch_k(skinny_outer_kernel, dim3(grid), dim3((unsigned)(threads)), (size_t)(0), (cudaStream_t)stream, a, make_map(a_map), b, make_map(b_map), c, c_batch, colsum,
                                                                 colsum_batch, M, I, J, accumulate));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
