// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1,
// 128 x BN x 16) -> fp32 accumulators double-buffered in tensor memory -> fused epilogue straight from TMEM.
// CTAs run as clusters of two along M: the pair shares its B (weight) tile -- each CTA fetches half of it and TMA-multicasts
// it into both shared memories -- which cuts L2->SM operand traffic per FLOP by a third (ncu: the 1-CTA version was
// L2-bandwidth bound at ~45% tensor-pipe utilisation).
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM lane quarter = warp_idx % 4), overlapped with the next tile's MMAs
//
// Replaces the cuBLAS calls behind F.linear in model/modeling_slot.py:101,113,61,65 (and their autograd
// dgrad / wgrad) -- see include/devias_b200.h for the epilogue catalogue.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

struct GemmParams {
  int M, N, K;
  int splits;
  void* out; long long ldo;
  void* out2; long long ldo2;
  const float* bias;
  const void* aux; long long ldaux; int aux_row_mod;
  const float* row_scale; int rows_per_scale;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KiB alignment
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, const uint32_t (&acc)[32]) {
  // one thread: row `row`, 32 consecutive columns starting at col0
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if constexpr (EPI != DEVIAS_EPI_DGELU_BF16 && EPI != DEVIAS_EPI_ATOMIC_F32) {
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    }
  }
  if constexpr (EPI == DEVIAS_EPI_STORE_BF16) {
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                        pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  } else if constexpr (EPI == DEVIAS_EPI_STORE_F32) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (EPI == DEVIAS_EPI_GELU_BF16) {
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col0);
    uint4* o2 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out2) + (long long)row * p.ldo2 + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                        pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = gelu_fast(v[8 * i + j]);
      o2[i] = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
    }
  } else if constexpr (EPI == DEVIAS_EPI_DGELU_BF16) {
    const uint4* a = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col0);
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 h = __ldg(a + i);
      const float2 h0 = unpack_bf16(h.x), h1 = unpack_bf16(h.y), h2 = unpack_bf16(h.z), h3 = unpack_bf16(h.w);
      o[i] = make_uint4(pack_bf16(v[8 * i] * gelu_fast_grad(h0.x), v[8 * i + 1] * gelu_fast_grad(h0.y)),
                        pack_bf16(v[8 * i + 2] * gelu_fast_grad(h1.x), v[8 * i + 3] * gelu_fast_grad(h1.y)),
                        pack_bf16(v[8 * i + 4] * gelu_fast_grad(h2.x), v[8 * i + 5] * gelu_fast_grad(h2.y)),
                        pack_bf16(v[8 * i + 6] * gelu_fast_grad(h3.x), v[8 * i + 7] * gelu_fast_grad(h3.y)));
    }
  } else if constexpr (EPI == DEVIAS_EPI_RESID_F32) {
    const int arow = p.aux_row_mod > 0 ? row % p.aux_row_mod : row;
    const float s = p.row_scale != nullptr ? __ldg(p.row_scale + row / p.rows_per_scale) : 1.0f;
    const float4* a = reinterpret_cast<const float4*>(static_cast<const float*>(p.aux) + (long long)arow * p.ldaux + col0);
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 r = __ldg(a + i);
      o[i] = make_float4(fmaf(s, v[4 * i], r.x), fmaf(s, v[4 * i + 1], r.y), fmaf(s, v[4 * i + 2], r.z),
                         fmaf(s, v[4 * i + 3], r.w));
    }
  } else if constexpr (EPI == DEVIAS_EPI_ATOMIC_F32) {
    float* o = static_cast<float*>(p.out) + (long long)row * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < 8; ++i) red_add_v4_f32(o + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int rank = (int)cluster_ctarank();          // 0/1: which half of the 256-row pair tile / which half of B we fetch
  const int m_blks = (p.M + kBM - 1) / kBM;
  const int m_pairs = (m_blks + 1) / 2;
  const int n_blks = (p.N + BN - 1) / BN;
  const int k_blks = (p.K + kBK - 1) / kBK;
  const int kb_per_split = (k_blks + p.splits - 1) / p.splits;
  const int tiles = m_pairs * n_blks * p.splits;    // pair tiles; both CTAs of a cluster walk the same sequence
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 2);   // released by the MMA warps of BOTH CTAs (the peer multicasts into our stage)
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers are initialised before anything is multicast into it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < tiles; t += n_clusters) {
        const int split = t / (m_pairs * n_blks);
        const int rem = t - split * (m_pairs * n_blks);
        const int m_blk = 2 * (rem / n_blks) + rank, n_blk = rem % n_blks;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(k_blks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBK, m_blk * kBM);
          } else {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i)
              tma_load_2d(sa + i * (kBK * 128), &tmA, &full_bar[stage], m_blk * kBM + i * 64, kb * kBK);
          }
          // our half of the shared B tile, multicast into both CTAs of the pair
          if constexpr (!B_MN) {
            tma_load_2d_mcast(sb + rank * (BN / 2) * 128, &tmB, &full_bar[stage], kb * kBK, n_blk * BN + rank * (BN / 2), 3);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) {
              const int bx = rank * (BN / 128) + i;
              tma_load_2d_mcast(sb + bx * (kBK * 128), &tmB, &full_bar[stage], n_blk * BN + bx * 64, kb * kBK, 3);
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO); one 128 B swizzle atom along K (LBO unused).
      // MN-major: 64-element MN atoms kBK*128 B apart (LBO); 8-k groups 1024 B apart (SBO).
      constexpr uint32_t lbo_a = A_MN ? kBK * 128 : 0, lbo_b = B_MN ? kBK * 128 : 0;
      constexpr uint32_t kstep_a = A_MN ? (16 * 128) >> 4 : 32 >> 4;  // descriptor start-address units (16 B) per UMMA_K
      constexpr uint32_t kstep_b = B_MN ? (16 * 128) >> 4 : 32 >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < tiles; t += n_clusters, ++it) {
        const int split = t / (m_pairs * n_blks);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(k_blks, kb0 + kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa, lbo_a, 1024);
          const uint64_t db = umma_desc_sw128(sa + Cfg::A_BYTES, lbo_b, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss(tmem_d, da + (uint64_t)(k * kstep_a), db + (uint64_t)(k * kstep_b), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_mcast(&empty_bar[stage], 3);  // slot reusable (in both CTAs) once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = cluster_id; t < tiles; t += n_clusters, ++it) {
      const int split = t / (m_pairs * n_blks);
      const int rem = t - split * (m_pairs * n_blks);
      const int m_blk = 2 * (rem / n_blks) + rank, n_blk = rem % n_blks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m_blk * kBM + q * 32 + (int)lane_id();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + c * 32, acc);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row < p.M && col0 < p.N) epilogue_chunk<EPI>(p, row, col0, acc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody leaves while the peer may still multicast into / signal this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const void* a, long long lda, const void* b, long long ldb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, a, p.K, p.M, lda * 2, kBK, kBM);
  else rc = make_tmap_2d_bf16(&tmA, a, p.M, p.K, lda * 2, 64, kBK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d_bf16(&tmB, b, p.K, p.N, ldb * 2, kBK, BN / 2);   // each CTA of a pair fetches half of B
  else rc = make_tmap_2d_bf16(&tmB, b, p.N, p.K, ldb * 2, 64, kBK);
  if (rc) return rc;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int m_blks = (p.M + kBM - 1) / kBM, n_blks = (p.N + BN - 1) / BN;
  const int pair_tiles = ((m_blks + 1) / 2) * n_blks * p.splits;
  const int max_clusters = sm_count() / 2;
  const int grid = 2 * (pair_tiles < max_clusters ? pair_tiles : max_clusters);
  const int prof = prof_begin(DEVIAS_PROF_GEMM, 2.0 * p.M * (double)p.N * p.K, stream);
  kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  prof_end(prof, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

template <int BN, int EPI>
static int dispatch_layout(bool a_mn, bool b_mn, const void* a, long long lda, const void* b, long long ldb,
                           const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EPI>(a, lda, b, ldb, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EPI>(a, lda, b, ldb, p, s);
  if (a_mn && b_mn) return launch_gemm<BN, true, true, EPI>(a, lda, b, ldb, p, s);
  set_last_error("layout", "A mn-major with B k-major is not instantiated (no caller on the DEVIAS path)", __FILE__, __LINE__);
  return DEVIAS_ERR_UNSUPPORTED;
}

template <int EPI>
static int dispatch_bn(int bn, bool a_mn, bool b_mn, const void* a, long long lda, const void* b, long long ldb,
                       const GemmParams& p, cudaStream_t s) {
  if (bn == 256) return dispatch_layout<256, EPI>(a_mn, b_mn, a, lda, b, ldb, p, s);
  return dispatch_layout<128, EPI>(a_mn, b_mn, a, lda, b, ldb, p, s);
}

}  // namespace dv

extern "C" int devias_gemm_bf16(const void* a, int64_t lda, int a_mn_major, const void* b, int64_t ldb, int b_mn_major, int m,
                                int n, int k, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2,
                                const float* bias, const void* aux, int64_t ldaux, int aux_row_mod, const float* row_scale,
                                int rows_per_scale, int split_k, void* stream) {
  using namespace dv;
  DV_REQUIRE(a && b && out, "null operand");
  DV_REQUIRE(m > 0 && n > 0 && k > 0, "empty problem");
  DV_REQUIRE(n % 32 == 0, "n must be a multiple of 32");
  DV_REQUIRE(k % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "k and leading dims must be multiples of 8 (16-byte TMA strides)");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "operands must be 16-byte aligned");
  DV_REQUIRE(ldo % 8 == 0, "ldo must be a multiple of 8");
  if (epilogue == DEVIAS_EPI_GELU_BF16) DV_REQUIRE(out2 && ldo2 % 8 == 0, "GELU epilogue needs out2");
  if (epilogue == DEVIAS_EPI_DGELU_BF16) DV_REQUIRE(aux && ldaux % 8 == 0, "DGELU epilogue needs aux");
  if (epilogue == DEVIAS_EPI_RESID_F32) DV_REQUIRE(aux && ldaux % 4 == 0, "RESID epilogue needs aux");
  if (row_scale) DV_REQUIRE(rows_per_scale > 0, "rows_per_scale");
  const int k_blks = (k + kBK - 1) / kBK;
  int splits = split_k < 1 ? 1 : split_k;
  if (epilogue != DEVIAS_EPI_ATOMIC_F32) splits = 1;
  if (splits > k_blks) splits = k_blks;
  {  // every split must own at least one k-block
    const int per = (k_blks + splits - 1) / splits;
    splits = (k_blks + per - 1) / per;
  }
  GemmParams p{m, n, k, splits, out, (long long)ldo, out2, (long long)ldo2, bias, aux, (long long)ldaux, aux_row_mod,
               row_scale, rows_per_scale};
  // BN = 256 maximises operand reuse; fall back to 128 when that leaves too few tiles for 148 SMs
  const int m_blks = (m + kBM - 1) / kBM;
  int bn = 256;
  if (n % 256 != 0 || ((m_blks + 1) / 2) * (n / 256) * splits < sm_count() / 2) bn = 128;
  if (n % 128 != 0 && bn == 128) bn = 128;  // tail columns are masked per 32-col chunk
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
  switch (epilogue) {
    case DEVIAS_EPI_STORE_BF16: return dispatch_bn<DEVIAS_EPI_STORE_BF16>(bn, amn, bmn, a, lda, b, ldb, p, s);
    case DEVIAS_EPI_STORE_F32: return dispatch_bn<DEVIAS_EPI_STORE_F32>(bn, amn, bmn, a, lda, b, ldb, p, s);
    case DEVIAS_EPI_GELU_BF16: return dispatch_bn<DEVIAS_EPI_GELU_BF16>(bn, amn, bmn, a, lda, b, ldb, p, s);
    case DEVIAS_EPI_DGELU_BF16: return dispatch_bn<DEVIAS_EPI_DGELU_BF16>(bn, amn, bmn, a, lda, b, ldb, p, s);
    case DEVIAS_EPI_RESID_F32: return dispatch_bn<DEVIAS_EPI_RESID_F32>(bn, amn, bmn, a, lda, b, ldb, p, s);
    case DEVIAS_EPI_ATOMIC_F32: return dispatch_bn<DEVIAS_EPI_ATOMIC_F32>(bn, amn, bmn, a, lda, b, ldb, p, s);
    default: break;
  }
  set_last_error("epilogue", "unknown epilogue id", __FILE__, __LINE__);
  return DEVIAS_ERR_ARG;
}
