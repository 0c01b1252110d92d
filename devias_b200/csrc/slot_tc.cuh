// Shared by the tcgen05 slot-attention kernels for bf16 tokens (slot_attn_tc.cu forward, slot_attn_tc_bwd.cu backward):
// tile geometry, small shared-memory / tensor-memory helpers, warp-convergent MMA issue, the token tensor map.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kTD = 768;                            // channels
constexpr int kTT = 32;                             // tokens per tile
constexpr int kTBoxes = kTD / 64;                   // 12 boxes of 64 bf16 channels (128 bytes)
constexpr int kTBoxBytes = kTT * 128;               // 4 KiB
constexpr int kTTileBytes = kTBoxes * kTBoxBytes;   // 48 KiB

__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// a bf16 pair held in one 32-bit word -> (low element, high element) as a packed fp32 pair
__device__ __forceinline__ uint64_t bf16x2_to_f2(uint32_t w) {
  return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// Issued by every lane of a converged warp with warp-uniform operands; only the elected lane (leader != 0) executes the MMA.
// Keeping the issuing code convergent lets the compiler hold descriptors in uniform registers: inside an `if (lane == 0)`
// region every tcgen05.mma is wrapped in an R2UR / ELECT / BRA.U.ANY waterfall loop (~60 cycles per MMA, measured).
__device__ __forceinline__ void umma_ss_lead(uint32_t leader, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_lead(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}
template <int HSP>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[HSP]) {
  if constexpr (HSP == 16) tmem_ld_32x32b_x16(taddr, r);
  else tmem_ld_32x32b_x32(taddr, r);
}

// 4-D view of the bf16 tokens [B, N, 768] as (64 channels | N tokens | 12 channel boxes | B); one box [64, 32, 12, 1] per tile
static inline int make_token_tmap_bf16(CUtensorMap* tm, const void* tokens, int B, int N) {
  const uint64_t dims[4] = {64, (uint64_t)N, (uint64_t)kTBoxes, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)kTD * 2, 128, (uint64_t)N * kTD * 2};
  const uint32_t box[4] = {64, kTT, kTBoxes, 1};
  return make_tmap_nd(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, tokens, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}


}  // namespace dv
