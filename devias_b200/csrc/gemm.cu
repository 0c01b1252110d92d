// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma.cta_group::2
// (256 x BN x 16 over the two SMs of a cluster) -> fp32 accumulators double-buffered in tensor memory -> fused epilogue ->
// TMA store / TMA reduce-add.  CTAs run as pairs along M: each CTA holds its own 128 rows of A and HALF of the B (weight)
// tile; the leader CTA issues one MMA for both and the tensor cores read the other half of B from the peer's shared memory.
// Per k-step and SM that is 32 KiB written by TMA and 32 KiB read by the MMA -- the single-CTA version (128 x 256 tiles, B
// multicast) moved 48 + 48 KiB, above what the 128 B/clk shared memory sustains next to a full-rate tensor pipe, and topped
// out near 1000 TFLOP/s.
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  : epilogue (TMEM lane quarter = warp_idx % 4; the two warps of a quarter split the tile's columns),
//                 overlapped with the next tile's MMAs.  Each warp owns a 32-row x 128-byte staging box: it reads its
//                 accumulator rows from TMEM (thread = row), applies the epilogue, writes the box 128B-swizzled and one lane
//                 issues a bulk tensor store (or reduce-add for split-K weight gradients); residual / pre-activation
//                 operands arrive the same way (bulk tensor load into a second box, one box ahead).  Global memory is thus
//                 only touched by TMA in full lines -- the first version's row-per-thread LDG/STG epilogue was L1-tag bound
//                 (32 sectors per instruction; ncu l1tex ~50 %, K=768 GEMMs at 370-850 TFLOP/s).
//
// Replaces the cuBLAS calls behind F.linear in model/modeling_slot.py:101,113,61,65 (and their autograd
// dgrad / wgrad) -- see include/devias_b200.h for the epilogue catalogue.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

struct GemmParams {
  int M, N, K;
  int splits;
  const float* bias;
  int aux_row_mod;
  const float* row_scale; int rows_per_scale;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int kBoxBytes = 32 * 128; // one epilogue staging box: 32 rows x 128 B

template <int EPI>
struct EpiTraits {
  static constexpr bool OUT_F32 = EPI == DEVIAS_EPI_STORE_F32 || EPI == DEVIAS_EPI_RESID_F32 || EPI == DEVIAS_EPI_ATOMIC_F32;
  static constexpr bool HAS_AUX = EPI == DEVIAS_EPI_RESID_F32 || EPI == DEVIAS_EPI_DGELU_BF16;
  static constexpr bool TWO_OUT = EPI == DEVIAS_EPI_GELU_BF16;
  static constexpr int BOX_COLS = OUT_F32 ? 32 : 64;   // 128 bytes of output per row
  static constexpr int BOXES_PER_WARP_SMEM = (HAS_AUX || TWO_OUT) ? 2 : 1;
};

template <int BN, int EPI>
struct GemmCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = (BN / 2) * kBK * 2;      // this CTA's half of the pair's B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 8 * EpiTraits<EPI>::BOXES_PER_WARP_SMEM * kBoxBytes;   // 32 or 64 KiB
  static constexpr int FIT = (227 * 1024 - 2048 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 6 ? 6 : FIT;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;  // +1024: manual 1 KiB alignment
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// one staging row (128 B) <-> registers, 128B-swizzled: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
__device__ __forceinline__ void stage_write_row(uint32_t box, int r, const uint32_t (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) sts128(box + r * 128 + ((c ^ (r & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void stage_read_row(uint32_t box, int r, uint32_t (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 t = lds128(box + r * 128 + ((c ^ (r & 7)) << 4));
    v[4 * c] = __float_as_uint(t.x); v[4 * c + 1] = __float_as_uint(t.y);
    v[4 * c + 2] = __float_as_uint(t.z); v[4 * c + 3] = __float_as_uint(t.w);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                 const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  pdl_trigger();
  using Cfg = GemmCfg<BN, EPI>;
  using ET = EpiTraits<EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;                 // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 8);

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
    prefetch_tmap(&tmOut);
    if constexpr (ET::TWO_OUT) prefetch_tmap(&tmOut2);
    if constexpr (ET::HAS_AUX) prefetch_tmap(&tmAux);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);    // used in the leader only: its producer's arrive + the bytes of BOTH CTAs' loads
        mbar_init(&empty_bar[s], 1);   // released in both CTAs by the leader's MMA commit
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);   // the leader's commit, multicast
        mbar_init(&tempty_bar[s], 16); // used in the leader only: the epilogue warps of both CTAs
      }
      for (int s = 0; s < 8; ++s) mbar_init(&aux_bar[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();              // everything above overlapped the previous kernel's tail; global memory is touched only below

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
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);   // the leader's barrier counts the pair's bytes
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, lbar, kb * kBK, m_blk * kBM);
          } else {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i) tma_load_2d_2sm(sa + i * (kBK * 128), &tmA, lbar, m_blk * kBM + i * 64, kb * kBK);
          }
          // our half (BN/2 rows of N) of the pair's B tile
          if constexpr (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, lbar, kb * kBK, n_blk * BN + rank * (BN / 2));
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i)
              tma_load_2d_2sm(sb + i * (kBK * 128), &tmB, lbar, n_blk * BN + (rank * (BN / 128) + i) * 64, kb * kBK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of the pair only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, BN, A_MN, B_MN);
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
            umma_ss_2sm(tmem_d, da + (uint64_t)(k * kstep_a), db + (uint64_t)(k * kstep_b), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 3);    // slot reusable (in both CTAs) once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[as], 3);         // accumulator complete (both CTAs' epilogues)
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;              // 0..7
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;            // which half of the tile's columns this warp drains
    const int lane = (int)lane_id();
    constexpr int kBoxes = (BN / 2) / ET::BOX_COLS;          // output boxes per warp per tile
    uint8_t* box_out_p = smem + Cfg::OFF_EPI + ew * ET::BOXES_PER_WARP_SMEM * kBoxBytes;
    uint8_t* box_aux_p = box_out_p + kBoxBytes;              // second box: aux operand (RESID/DGELU) or second output (GELU)
    const uint32_t box_out = smem_u32(box_out_p);
    const uint32_t box_aux = box_out + kBoxBytes;
    uint64_t* my_aux_bar = &aux_bar[ew];
    uint32_t aux_uses = 0;
    int it = 0;
    for (int t = cluster_id; t < tiles; t += n_clusters, ++it) {
      const int split = t / (m_pairs * n_blks);
      const int rem = t - split * (m_pairs * n_blks);
      const int m_blk = 2 * (rem / n_blks) + rank, n_blk = rem % n_blks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * kBM + q * 32;                 // first of this warp's 32 rows
      const int row = row0 + lane;
      const int colbase = n_blk * BN + chalf * (BN / 2);
      const bool active = row0 < p.M && colbase < p.N;       // warp-uniform
      float rs = 1.0f;
      if constexpr (EPI == DEVIAS_EPI_RESID_F32) {
        if (active && p.row_scale != nullptr && row < p.M) rs = __ldg(p.row_scale + row / p.rows_per_scale);
      }
      auto issue_aux = [&](int b) {   // lane 0: bulk tensor load of aux box b (rows of `aux` may repeat modulo aux_row_mod)
        const int arow0 = p.aux_row_mod > 0 ? row0 % p.aux_row_mod : row0;
        mbar_arrive_expect_tx(my_aux_bar, kBoxBytes);
        tma_load_2d(box_aux_p, &tmAux, my_aux_bar, colbase + b * ET::BOX_COLS, arow0);
      };
      if constexpr (ET::HAS_AUX) {
        if (active && lane == 0) issue_aux(0);
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + chalf * (BN / 2);
      if (active) {
#pragma unroll 1
        for (int b = 0; b < kBoxes; ++b) {
          const int col0 = colbase + b * ET::BOX_COLS;
          if (col0 >= p.N) break;                            // warp-uniform (ragged N)
          // ---- auxiliary operand of this box -> registers; then prefetch the next one
          uint32_t aux[ET::HAS_AUX ? 32 : 1];
          if constexpr (ET::HAS_AUX) {
            mbar_wait(my_aux_bar, aux_uses & 1);
            ++aux_uses;
            stage_read_row(box_aux, lane, aux);
            __syncwarp();
            if (lane == 0 && b + 1 < kBoxes && col0 + ET::BOX_COLS < p.N) issue_aux(b + 1);
          }
          // ---- accumulator -> registers, epilogue math, packed into `outv` (32 x 32-bit = one 128-byte row)
          uint32_t outv[32];
          uint32_t outv2[ET::TWO_OUT ? 32 : 1];
          if constexpr (ET::OUT_F32) {
            uint32_t acc[32];
            tmem_ld_32x32b_x32(taddr + b * 32, acc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 v = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                                     __uint_as_float(acc[4 * i + 3]));
              if constexpr (EPI != DEVIAS_EPI_ATOMIC_F32) {
                if (p.bias != nullptr) {
                  const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
                  v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                }
              }
              if constexpr (EPI == DEVIAS_EPI_RESID_F32) {
                v.x = fmaf(rs, v.x, __uint_as_float(aux[4 * i])); v.y = fmaf(rs, v.y, __uint_as_float(aux[4 * i + 1]));
                v.z = fmaf(rs, v.z, __uint_as_float(aux[4 * i + 2])); v.w = fmaf(rs, v.w, __uint_as_float(aux[4 * i + 3]));
              }
              outv[4 * i] = __float_as_uint(v.x); outv[4 * i + 1] = __float_as_uint(v.y);
              outv[4 * i + 2] = __float_as_uint(v.z); outv[4 * i + 3] = __float_as_uint(v.w);
            }
          } else {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {              // 64 output columns = two 32-column TMEM reads
              uint32_t acc[32];
              tmem_ld_32x32b_x32(taddr + b * 64 + hlf * 32, acc);
              tmem_ld_wait();
              const int cc = col0 + hlf * 32;
              const bool has_bias = EPI != DEVIAS_EPI_DGELU_BF16 && p.bias != nullptr && cc < p.N;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 v = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                                       __uint_as_float(acc[4 * i + 3]));
                if (has_bias) {
                  const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + cc) + i);
                  v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                }
                if constexpr (EPI == DEVIAS_EPI_DGELU_BF16) {
                  const float2 h0 = unpack_bf16(aux[hlf * 16 + 2 * i]), h1 = unpack_bf16(aux[hlf * 16 + 2 * i + 1]);
                  const uint64_t g0 = f2_mul(f2_pack(v.x, v.y), gelu_grad_pair(f2_pack(h0.x, h0.y)));
                  const uint64_t g1 = f2_mul(f2_pack(v.z, v.w), gelu_grad_pair(f2_pack(h1.x, h1.y)));
                  v.x = f2_lo(g0); v.y = f2_hi(g0); v.z = f2_lo(g1); v.w = f2_hi(g1);
                }
                outv[hlf * 16 + 2 * i] = pack_bf16(v.x, v.y);
                outv[hlf * 16 + 2 * i + 1] = pack_bf16(v.z, v.w);
                if constexpr (ET::TWO_OUT) {
                  const uint64_t a0 = gelu_pair(f2_pack(v.x, v.y)), a1 = gelu_pair(f2_pack(v.z, v.w));
                  outv2[hlf * 16 + 2 * i] = pack_bf16(f2_lo(a0), f2_hi(a0));
                  outv2[hlf * 16 + 2 * i + 1] = pack_bf16(f2_lo(a1), f2_hi(a1));
                }
              }
            }
          }
          // ---- staging box free again? (the previous bulk store of this warp has finished READING it)
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          stage_write_row(box_out, lane, outv);
          if constexpr (ET::TWO_OUT) stage_write_row(box_aux, lane, outv2);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if constexpr (EPI == DEVIAS_EPI_ATOMIC_F32) tma_reduce_add_2d(&tmOut, box_out_p, col0, row0);
            else tma_store_2d(&tmOut, box_out_p, col0, row0);
            if constexpr (ET::TWO_OUT) tma_store_2d(&tmOut2, box_aux_p, col0, row0);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));   // the leader's MMA warp waits for both CTAs
    }
    if (lane == 0) bulk_wait0();   // all of this warp's stores / reductions are complete before the CTA may exit
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody leaves while the peer may still read our B half / signal this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
struct GemmHostArgs {
  const void* a; long long lda; const void* b; long long ldb;
  void* out; long long ldo; void* out2; long long ldo2;
  const void* aux; long long ldaux;
};

static int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t str[1] = {row_stride_elems * elem_bytes};
  const uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_nd(out, dt, 2, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const GemmHostArgs& h, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  using ET = EpiTraits<EPI>;
  static_assert(Cfg::STAGES >= 3, "pipeline too shallow");
  CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
  int rc;
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, h.a, p.K, p.M, h.lda * 2, kBK, kBM);
  else rc = make_tmap_2d_bf16(&tmA, h.a, p.M, p.K, h.lda * 2, 64, kBK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d_bf16(&tmB, h.b, p.K, p.N, h.ldb * 2, kBK, BN / 2);   // each CTA of a pair fetches half of B
  else rc = make_tmap_2d_bf16(&tmB, h.b, p.N, p.K, h.ldb * 2, 64, kBK);
  if (rc) return rc;
  const CUtensorMapDataType odt = ET::OUT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int obytes = ET::OUT_F32 ? 4 : 2;
  rc = make_tmap_2d(&tmOut, odt, obytes, h.out, p.N, p.M, h.ldo, ET::BOX_COLS, 32);
  if (rc) return rc;
  tmOut2 = tmOut;
  tmAux = tmOut;
  if (ET::TWO_OUT) {
    rc = make_tmap_2d(&tmOut2, odt, obytes, h.out2, p.N, p.M, h.ldo2, ET::BOX_COLS, 32);
    if (rc) return rc;
  }
  if (ET::HAS_AUX) {
    const uint64_t arows = p.aux_row_mod > 0 ? (uint64_t)p.aux_row_mod : (uint64_t)p.M;
    rc = make_tmap_2d(&tmAux, odt, obytes, h.aux, p.N, arows, h.ldaux, ET::BOX_COLS, 32);
    if (rc) return rc;
  }
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
  DV_CHECK_CUDA(launch_k(kern, dim3((unsigned)(grid)), dim3((unsigned)(kGemmThreads)), (size_t)(Cfg::SMEM_BYTES), stream, tmA, tmB, tmOut, tmOut2, tmAux, p));
  prof_end(prof, stream);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

template <int BN, int EPI>
static int dispatch_layout(bool a_mn, bool b_mn, const GemmHostArgs& h, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EPI>(h, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EPI>(h, p, s);
  if (a_mn && b_mn) return launch_gemm<BN, true, true, EPI>(h, p, s);
  set_last_error("layout", "A mn-major with B k-major is not instantiated (no caller on the DEVIAS path)", __FILE__, __LINE__);
  return DEVIAS_ERR_UNSUPPORTED;
}

template <int EPI>
static int dispatch_bn(int bn, bool a_mn, bool b_mn, const GemmHostArgs& h, const GemmParams& p, cudaStream_t s) {
  if (bn == 256) return dispatch_layout<256, EPI>(a_mn, b_mn, h, p, s);
  return dispatch_layout<128, EPI>(a_mn, b_mn, h, p, s);
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
  if (epilogue == DEVIAS_EPI_GELU_BF16)
    DV_REQUIRE(out2 && ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(out2) & 15) == 0, "GELU epilogue needs out2");
  if (epilogue == DEVIAS_EPI_DGELU_BF16)
    DV_REQUIRE(aux && ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, "DGELU epilogue needs aux");
  if (epilogue == DEVIAS_EPI_RESID_F32) {
    DV_REQUIRE(aux && ldaux % 4 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, "RESID epilogue needs aux");
    DV_REQUIRE(aux_row_mod == 0 || aux_row_mod % 32 == 0, "aux_row_mod must be a multiple of 32 (epilogue boxes are 32 rows)");
  }
  if (row_scale) DV_REQUIRE(rows_per_scale > 0, "rows_per_scale");
  const int k_blks = (k + kBK - 1) / kBK;
  int splits = split_k < 1 ? 1 : split_k;
  if (epilogue != DEVIAS_EPI_ATOMIC_F32) splits = 1;
  if (splits > k_blks) splits = k_blks;
  {  // every split must own at least one k-block
    const int per = (k_blks + splits - 1) / splits;
    splits = (k_blks + per - 1) / per;
  }
  GemmParams p{m, n, k, splits, bias, aux_row_mod, row_scale, rows_per_scale};
  GemmHostArgs h{a, (long long)lda, b, (long long)ldb, out, (long long)ldo, out2, (long long)ldo2, aux, (long long)ldaux};
  // BN = 256 maximises operand reuse; fall back to 128 when that leaves too few tiles for the 74 clusters
  const int m_blks = (m + kBM - 1) / kBM;
  int bn = 256;
  if (n % 256 != 0 || ((m_blks + 1) / 2) * (n / 256) * splits < sm_count() / 2) bn = 128;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
  switch (epilogue) {
    case DEVIAS_EPI_STORE_BF16: return dispatch_bn<DEVIAS_EPI_STORE_BF16>(bn, amn, bmn, h, p, s);
    case DEVIAS_EPI_STORE_F32: return dispatch_bn<DEVIAS_EPI_STORE_F32>(bn, amn, bmn, h, p, s);
    case DEVIAS_EPI_GELU_BF16: return dispatch_bn<DEVIAS_EPI_GELU_BF16>(bn, amn, bmn, h, p, s);
    case DEVIAS_EPI_DGELU_BF16: return dispatch_bn<DEVIAS_EPI_DGELU_BF16>(bn, amn, bmn, h, p, s);
    case DEVIAS_EPI_RESID_F32: return dispatch_bn<DEVIAS_EPI_RESID_F32>(bn, amn, bmn, h, p, s);
    case DEVIAS_EPI_ATOMIC_F32: return dispatch_bn<DEVIAS_EPI_ATOMIC_F32>(bn, amn, bmn, h, p, s);
    default: break;
  }
  set_last_error("epilogue", "unknown epilogue id", __FILE__, __LINE__);
  return DEVIAS_ERR_ARG;
}
