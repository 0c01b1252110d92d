// Tube patch embedding as an IMPLICIT GEMM (model/modeling_slot.py:167-177 Conv3d(k = s = (2, 16, 16)) + flatten/transpose, plus the
// bias and the fixed sin-cos position table of :181-191 / :354-355 in the epilogue).
//
//   x0[b, t*196 + h*14 + w, d] = sum_{c, dt, dy, dx} clip[b, c, 2t + dt, 16h + dy, 16w + dx] * W[d, c, dt, dy, dx] + bias[d] + pos[token, d]
//
// Kernel size == stride, so the im2col matrix is a pure re-indexing of the clip: the A operand is fetched by TMA straight from
// the NCTHW fp32 clip through a 5-D tensor map  (dx:16 | dy:16 | w:14 | h:14 | z = (b*3 + c)*16 + frame)  with box
// [16, 1, 14, 7, 1]: one box = the 98 tokens of half a frame pair x 16 consecutive k (one dy row of 16 dx), landing in shared
// memory as 98 rows of 64 bytes -- exactly a K-major, 64B-swizzled UMMA operand tile (the inner box extent must equal the swizzle
// span: with the 128-byte swizzle the TMA unit pads every 64-byte dx row to 128 bytes).  No patch matrix is materialised and the
// clip is never converted: the MMAs run as kind::tf32 on the fp32 data (10-bit mantissa, finer than the bf16 operands of the rest
// of the encoder), fp32 accumulation in tensor memory.  k = c*512 + dt*256 + dy*16 + dx is also the memory order of the Conv3d
// weight [768, 3, 2, 16, 16], so B is a plain 2-D box of the fp32 master weight.
//
// The WEIGHT GRADIENT keeps the explicit route (devias_patchify in the backward + the bf16 MN-major GEMM): its contraction runs
// over tokens, which forces MN-major operands, and tcgen05 kind::tf32 accepts K-major operands only (an MN-major tf32 prototype
// of dW = dX0^T im2col(clip) over the same 5-D boxes returned exact zeros on sm_100a; the 16-bit kinds transpose, the 32-bit one
// does not) -- an in-kernel fp32 -> bf16 conversion stage would be needed to drop the patch matrix there too.
//
//   CTA  = (98-token tile, 256-column block of the 768 outputs); 96 k-chunks of 16 through a 3-stage TMA ring, two CTAs per SM
//   warp 0 : TMA producer      warp 1 : tcgen05.mma issuer (M = 128 rows of which 98 are tokens, N = 256, K = 8 per instruction,
//                              two per chunk)
//   warps 2..5 : epilogue -- TMEM -> registers, + bias + position rows (bulk tensor load of the [32 x 32] box), 128B-swizzled
//                staging box -> bulk tensor store; tensor maps over [tile][98][768] views clip the 30 padding rows of every tile
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kPeTok = 98;            // tokens per tile: 7 patch rows x 14 patch columns of one frame pair
constexpr int kPeBN = 256;
constexpr int kPeKC = 16;             // fp32 elements per k-chunk = one 64-byte swizzle row = the 16 dx of one patch row
constexpr int kPeStages = 3;            // 104 KiB per CTA: two CTAs per SM (their TMA streams and epilogues overlap)
constexpr int kPeThreads = 192;
struct PeSmem {
  static constexpr int A_BYTES = 128 * 64;                   // 98 rows are written by the TMA box, 128 are addressed by the MMA
  static constexpr int A_TX = kPeTok * 64;
  static constexpr int B_BYTES = kPeBN * 64;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int OFF_EPI = kPeStages * STAGE;          // 4 warps x (output box + position box), 4 KiB each
  static constexpr int OFF_BAR = OFF_EPI + 4 * 2 * 4096;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};
static_assert(PeSmem::BYTES <= 227 * 1024, "patch embed shared memory");

struct PeParams {
  int tiles_per_clip;     // 16 = 8 frame pairs x 2 halves
  const float* bias;      // [768]
};

// kind::tf32 instruction descriptor: [4,6) D fmt (1 = f32) | [7,10) A fmt (2 = tf32) | [10,13) B fmt | 15/16 majors | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// shared-memory descriptor, K-major, 64-byte swizzle: 8-row x 64-byte atoms, sbo = 512 bytes between 8-row groups, layout type 4
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((512 >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kPeThreads, 2)
patch_embed_fwd_kernel(const __grid_constant__ CUtensorMap tmClip, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPos, const PeParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PeSmem::OFF_BAR);
  uint64_t* empty_bar = full_bar + kPeStages;
  uint64_t* tfull_bar = empty_bar + kPeStages;
  uint64_t* pos_bar = tfull_bar + 1;                 // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pos_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int n_blk = blockIdx.x;                      // 0..2: the three column blocks of a tile are neighbours (A stays in L2)
  const int tile = blockIdx.y;                       // (b * 8 + t) * 2 + half
  const int b = tile / p.tiles_per_clip, tt = tile % p.tiles_per_clip;
  const int t = tt >> 1, half = tt & 1;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmClip);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmOut);
    prefetch_tmap(&tmPos);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < kPeStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tfull_bar, 1);
      for (int s = 0; s < 4; ++s) mbar_init(&pos_bar[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kPeBN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  constexpr int kChunks = 3 * 2 * 16;                // (c, dt) x 16 dy rows

  if (warp == 0) {
    if (elect_one()) {
      for (int kc = 0; kc < kChunks; ++kc) {
        const int st = kc % kPeStages;
        mbar_wait(&empty_bar[st], ((kc / kPeStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[st], PeSmem::A_TX + PeSmem::B_BYTES);
        uint8_t* sa = smem + st * PeSmem::STAGE;
        const int cdt = kc >> 4, c = cdt >> 1, dt = cdt & 1;
        tma_load_5d(sa, &tmClip, &full_bar[st], 0, kc & 15, 0, 7 * half, (b * 3 + c) * 16 + 2 * t + dt);
        tma_load_2d(sa + PeSmem::A_BYTES, &tmW, &full_bar[st], kc * kPeKC, n_blk * kPeBN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, kPeBN);
      for (int kc = 0; kc < kChunks; ++kc) {
        const int st = kc % kPeStages;
        mbar_wait(&full_bar[st], (kc / kPeStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * PeSmem::STAGE);
        const uint64_t da = umma_desc_sw64(sa), db = umma_desc_sw64(sa + PeSmem::A_BYTES);
#pragma unroll
        for (int k = 0; k < kPeKC / 8; ++k) umma_ss_tf32(tmem, da + 2 * k, db + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[st]);
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    const int ew = warp - 2, q = warp & 3, lane = (int)lane_id();
    uint8_t* box_out_p = smem + PeSmem::OFF_EPI + ew * 2 * 4096;
    uint8_t* box_pos_p = box_out_p + 4096;
    const uint32_t box_out = smem_u32(box_out_p), box_pos = smem_u32(box_pos_p);
    const int row0 = q * 32;                          // rows >= 98 are clipped by the tensor maps (loads zero-fill, stores drop)
    const bool active = row0 < kPeTok;
    auto issue_pos = [&](int bx) {
      mbar_arrive_expect_tx(&pos_bar[ew], 4096);
      // position rows of this tile: pos viewed as [16 tiles per clip][98][768]
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(box_pos),
          "l"(reinterpret_cast<uint64_t>(&tmPos)), "r"(smem_u32(&pos_bar[ew])), "r"(n_blk * kPeBN + bx * 32), "r"(row0), "r"(tt)
          : "memory");
    };
    if (active && lane == 0) issue_pos(0);
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (active) {
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int bx = 0; bx < kPeBN / 32; ++bx) {
        const int col0 = n_blk * kPeBN + bx * 32;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + bx * 32, acc);
        mbar_wait(&pos_bar[ew], bx & 1);
        float pos[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(box_pos + lane * 128 + ((c ^ (lane & 7)) << 4));
          pos[4 * c] = v.x; pos[4 * c + 1] = v.y; pos[4 * c + 2] = v.z; pos[4 * c + 3] = v.w;
        }
        __syncwarp();
        if (lane == 0 && bx + 1 < kPeBN / 32) issue_pos(bx + 1);
        tmem_ld_wait();
        if (lane == 0) bulk_wait_read0();             // the previous store has finished reading the output box
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + c);
          sts128f(box_out + lane * 128 + ((c ^ (lane & 7)) << 4),
                  make_float4(__uint_as_float(acc[4 * c]) + bb.x + pos[4 * c], __uint_as_float(acc[4 * c + 1]) + bb.y + pos[4 * c + 1],
                              __uint_as_float(acc[4 * c + 2]) + bb.z + pos[4 * c + 2], __uint_as_float(acc[4 * c + 3]) + bb.w + pos[4 * c + 3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmOut, box_out_p, col0, row0, tile);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kPeBN>(tmem);
  }
}

}  // namespace dv

extern "C" int devias_patch_embed_fwd(const float* clip, const float* weight, const float* bias, const float* pos, float* out,
                                      int batch, int chans, int frames, int height, int width, int dim, void* stream) {
  using namespace dv;
  DV_REQUIRE(clip && weight && bias && pos && out, "null pointer");
  DV_REQUIRE(chans == 3 && frames == 16 && height == 224 && width == 224 && dim == 768,
             "the implicit-GEMM patch embedding is instantiated for 3 x 16 x 224 x 224 clips and 768 outputs (DEVIAS / VideoMAE ViT-B/16)");
  DV_REQUIRE(((reinterpret_cast<uintptr_t>(clip) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(pos) |
               reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "16-byte alignment");
  if (batch <= 0) return DEVIAS_OK;
  CUtensorMap tmClip, tmW, tmOut, tmPos;
  int rc;
  {  // (dx | dy | w | h | z): a token is (h, w), its 16 k-values of one chunk are the dx of one dy row
    const uint64_t dims[5] = {16, 16, 14, 14, (uint64_t)batch * 48};
    const uint64_t str[4] = {224 * 4, 16 * 4, 224 * 16 * 4, 224 * 224 * 4};
    const uint32_t box[5] = {16, 1, 14, 7, 1};
    rc = make_tmap_nd(&tmClip, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, clip, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {1536, 768};
    const uint64_t str[1] = {1536 * 4};
    const uint32_t box[2] = {kPeKC, kPeBN};
    rc = make_tmap_nd(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, weight, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {768, kPeTok, (uint64_t)batch * 16};
    const uint64_t str[2] = {768 * 4, (uint64_t)kPeTok * 768 * 4};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap_nd(&tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const uint64_t pdims[3] = {768, kPeTok, 16};
    rc = make_tmap_nd(&tmPos, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pos, pdims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PeSmem::BYTES));
    attr_done = true;
  }
  PeParams p{16, bias};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DV_CHECK_CUDA(launch_k(patch_embed_fwd_kernel, dim3(768 / kPeBN, batch * 16), dim3(kPeThreads), (size_t)PeSmem::BYTES, s, tmClip,
                         tmW, tmOut, tmPos, p));
  count_launch();
  return DEVIAS_OK;
}
