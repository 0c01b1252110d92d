// Fused softmax attention for the encoder (model/modeling_slot.py:102-112): out = softmax(scale * Q K^T) V per (clip, head),
// head_dim 64, sequence 1568 (any N), no mask, no dropout -- the 12 x N x N probability matrix never touches HBM.
//
// Forward: flash_fwd2_kernel (four small CTAs per SM, P kept in tensor memory); backward: flash_bwd_kernel (transposed
// formulation, P^T / dS^T consumed from tensor memory).  Both are described at their definitions below.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kHD = 64;        // head dim
constexpr int kQT = 128;       // queries per CTA
constexpr int kKT = 64;        // keys per tile

struct FaParams {
  int B, N, H;
  int Npad;                         // row length of lse2 / delta: N rounded up to a multiple of 128
  float scale_log2;                 // head_dim^-0.5 * log2(e)
  __nv_bfloat16* out; long long ldo;  // [B*N, H*64]
  float* lse2;                      // [B, H, N]  log2-domain log-sum-exp of the scaled scores
};

// =====================================================================================================================
// Forward: FOUR small CTAs per SM.
// ncu on the first-generation kernel (two 6-warp CTAs per SM, removed): MUFU (ex2) 48 % and tensor pipe 21 % busy, one softmax
// warp per scheduler and CTA stalled on its own serial chain (TMEM load -> max -> 64 ex2 -> pack -> TMEM store -> barrier), i.e. latency bound with two CTAs per SM.
// Here a CTA is 5 warps with <= 96 registers and 128 TMEM columns (S/P 64 + O 64), so four fit on an SM and their chains
// interleave on the shared MUFU / tensor pipes:
//   warp 0     : control -- one lane issues the TMA loads (Q once, K_j / V_j through a 2-stage ring) AND the MMAs
//                (S_j = Q K_j^T;  after the softmax of tile j:  O += P_j V_j, then at once S_{j+1});  tcgen05 ops of one thread
//                retire in order, so S_{j+1} cannot overwrite P_j early and `s_full` of tile j implies P_{j-1} V_{j-1} landed
//   warps 1..4 : softmax, thread = query row, in 16-column chunks to stay small: pass 1 row maximum, lazy rescale of O
//                (only when the maximum moved by more than 8 in the exp2 domain), pass 2 re-reads the chunk, ex2, packs P_j
//                as bf16 pairs over the S columns already consumed (the MMA reads P from tensor memory)
constexpr int kFa2Threads = 160;
struct Fa2Smem {
  static constexpr int OFF_Q = 0;                                  // 16 KiB
  static constexpr int OFF_K = OFF_Q + kQT * kHD * 2;              // 2 x 8 KiB
  static constexpr int OFF_V = OFF_K + 2 * kKT * kHD * 2;          // 2 x 8 KiB
  static constexpr int OFF_BAR = OFF_V + 2 * kKT * kHD * 2;
  static constexpr int BYTES = OFF_BAR + 128 + 1024;
};

__global__ void __launch_bounds__(kFa2Threads, 4)
flash_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const FaParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fa2Smem::OFF_BAR);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // 2
  uint64_t* kv_empty = bars + 3;           // 2
  uint64_t* s_full = bars + 5;             // 1
  uint64_t* p_full = bars + 6;             // 1 (4 warp arrivals)
  uint64_t* o_done = bars + 7;             // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int q_tiles = (p.N + kQT - 1) / kQT;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * kQT;
  const int T = (p.N + kKT - 1) / kKT;     // key tiles
  const int D = p.H * kHD;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      mbar_init(s_full, 1);
      mbar_init(p_full, 4);
      mbar_init(o_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<128>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_s = tmem, tm_o = tmem + 64;
  pdl_wait();              // barrier init / TMEM allocation overlapped the previous kernel's tail

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kQT, kKT, false, false);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(kQT, kHD, false, true);    // O = P V   : A from TMEM, B (=V) MN-major
      constexpr uint32_t kTile = kKT * kHD * 2;
      auto load_kv = [&](int j) {
        const int st = j & 1;
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
        tma_load_3d(smem + Fa2Smem::OFF_K + st * kTile, &tmKV, &kv_full[st], D + h * kHD, j * kKT, b);
        tma_load_3d(smem + Fa2Smem::OFF_V + st * kTile, &tmKV, &kv_full[st], 2 * D + h * kHD, j * kKT, b);
      };
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(smem + Fa2Smem::OFF_Q), 0, 1024);
        const uint64_t db = umma_desc_sw128(smem_u32(smem + Fa2Smem::OFF_K + st * kTile), 0, 1024);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(tm_s, da + 2 * k, db + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_arrive_expect_tx(q_full, kQT * kHD * 2);
      tma_load_3d(smem + Fa2Smem::OFF_Q, &tmQ, q_full, h * kHD, q0, b);
      load_kv(0);
      if (T > 1) load_kv(1);
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);            // P_j is in TMEM and any rescaling of O has been fenced
        tc_fence_after();
        const uint64_t db = umma_desc_sw128(smem_u32(smem + Fa2Smem::OFF_V + st * kTile), kKT * 128, 1024);
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k) umma_ts(tm_o, tm_s + 8 * k, db + 128 * k, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        if (j + 1 < T) {
          issue_s(j + 1);                    // queued behind P_j V_j on the tensor pipe
          if (j + 2 < T) {
            mbar_wait(&kv_empty[st], (j >> 1) & 1);   // P_j V_j retired: its K/V stage is free for tile j + 2
            load_kv(j + 2);
          }
        } else {
          umma_commit(o_done);
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (warps 1..4 -> 1, 2, 3, 0)
    const int lane = (int)lane_id();
    const int r = q * 32 + lane;                     // query row inside the tile == TMEM lane
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    const uint64_t cc = f2_pack(p.scale_log2, p.scale_log2);
    for (int j = 0; j < T; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid = p.N - j * kKT;               // keys valid in this tile (>= 64 except the last)
      uint64_t sum2 = 0ull;
      auto tile = [&](auto ragged_tag) {             // the ragged last tile masks its tail; all others take the bare path
        constexpr bool RAGGED = decltype(ragged_tag)::value;
        // ---- pass 1: row maximum.  Chunk loads are software-pipelined: tcgen05.wait::ld covers every load issued so far, so
        // the next chunk is requested right after the wait and its latency hides behind the arithmetic on the current one.
        uint32_t sa[16], sb[16];
        float mx = -INFINITY;
        tmem_ld_32x32b_x16(tm_s + lane_sel, sa);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&cur)[16] = (c & 1) ? sb : sa;
          uint32_t (&nxt)[16] = (c & 1) ? sa : sb;
          tmem_ld_wait();
          tmem_ld_32x32b_x16(tm_s + lane_sel + 16 * ((c + 1) & 3), nxt);      // c = 3: chunk 0 again, for pass 2
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = (!RAGGED || 16 * c + i < valid) ? __uint_as_float(cur[i]) : -INFINITY;
            m4[i & 3] = fmaxf(m4[i & 3], v);
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
        const float m_run = fmaxf(m, mx * p.scale_log2);   // scale > 0: max(c s) = c max(s)
        if (__any_sync(0xffffffffu, m_run > m + 8.0f)) {
          const float alpha = fast_exp2(m - m_run);       // 0 on the first tile (m = -inf)
          if (j > 0) {                                    // (s_full of tile j implies P_{j-1} V_{j-1} has been accumulated)
            tmem_ld_wait();                               // chunk 0 of pass 2 (in sa) has landed; its registers stay live
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t t[16];
              tmem_ld_32x32b_x16(tm_o + lane_sel + 16 * c, t);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
              tmem_st_32x32b_x16(tm_o + lane_sel + 16 * c, t);
            }
          }
          l *= alpha;
          m = m_run;
        }
        // ---- pass 2: P_j = exp2(c s - m) as bf16 pairs, written over the S columns this thread has already consumed
        const uint64_t negm = f2_pack(-m, -m);
        sum2 = 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&cur)[16] = (c & 1) ? sb : sa;
          uint32_t (&nxt)[16] = (c & 1) ? sa : sb;
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32b_x16(tm_s + lane_sel + 16 * (c + 1), nxt);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = (!RAGGED || 16 * c + 2 * i < valid) ? __uint_as_float(cur[2 * i]) : -INFINITY;
            const float x1 = (!RAGGED || 16 * c + 2 * i + 1 < valid) ? __uint_as_float(cur[2 * i + 1]) : -INFINITY;
            const uint64_t y = f2_fma(f2_pack(x0, x1), cc, negm);
            // three of every eight pairs take the polynomial path on the FMA pipe, the rest MUFU.EX2 (ncu: XU 67 % busy)
            const uint64_t ev = (i == 1 || i == 4 || i == 6) ? exp2_poly_pair(y) : f2_pack(fast_exp2(f2_lo(y)), fast_exp2(f2_hi(y)));
            sum2 = f2_add(sum2, ev);
            pk[i] = pack_bf16(f2_lo(ev), f2_hi(ev));
          }
          tmem_st_32x32b_x8(tm_s + lane_sel + 8 * c, pk);
        }
      };
      if (valid < kKT) tile(std::true_type{}); else tile(std::false_type{});
      l += f2_lo(sum2) + f2_hi(sum2);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_done, 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int qi = q0 + r;
    __nv_bfloat16* dst = p.out + ((long long)b * p.N + qi) * p.ldo + h * kHD;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t t[16];
      tmem_ld_32x32b_x16(tm_o + lane_sel + 16 * c, t);
      tmem_ld_wait();
      if (qi < p.N) {
        uint4 v0, v1;
        v0.x = pack_bf16(__uint_as_float(t[0]) * inv, __uint_as_float(t[1]) * inv);
        v0.y = pack_bf16(__uint_as_float(t[2]) * inv, __uint_as_float(t[3]) * inv);
        v0.z = pack_bf16(__uint_as_float(t[4]) * inv, __uint_as_float(t[5]) * inv);
        v0.w = pack_bf16(__uint_as_float(t[6]) * inv, __uint_as_float(t[7]) * inv);
        v1.x = pack_bf16(__uint_as_float(t[8]) * inv, __uint_as_float(t[9]) * inv);
        v1.y = pack_bf16(__uint_as_float(t[10]) * inv, __uint_as_float(t[11]) * inv);
        v1.z = pack_bf16(__uint_as_float(t[12]) * inv, __uint_as_float(t[13]) * inv);
        v1.w = pack_bf16(__uint_as_float(t[14]) * inv, __uint_as_float(t[15]) * inv);
        reinterpret_cast<uint4*>(dst)[2 * c] = v0;
        reinterpret_cast<uint4*>(dst)[2 * c + 1] = v1;
      }
    }
    if (p.lse2 != nullptr) {
      if (qi < p.N) p.lse2[((long long)b * p.H + h) * p.Npad + qi] = m + log2f(l);
      else if (qi < p.Npad) p.lse2[((long long)b * p.H + h) * p.Npad + qi] = INFINITY;   // padded queries: P = 0 in the backward
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

// =====================================================================================================================
// Backward, TRANSPOSED formulation.  One CTA per (clip, head, 128-key tile), loop over 128-query tiles in two 64-query halves.
// The first backward kept thread = query row (S = Q K^T) and moved P and dS through shared memory three times (A operands of
// dV, dK and dQ): 336 KiB of shared-memory traffic per (128 q x 128 k) step against a 128 B/clk port, twice the tensor time
// (profiles/r01_ncu_flash.md).  Here the scores are computed transposed, thread = KEY row (TMEM lane = key):
//   S^T = K Q^T,  dP^T = V dO^T      A = K / V live in TENSOR MEMORY for the whole CTA (copied once), B = Q_i / dO_i tiles
//   P^T = exp2(c S^T - lse2[q]),  dS^T = P^T (dP^T - delta[q])  -> bf16 pairs written IN PLACE over the fp32 columns
//   dV += P^T dO_i,  dK += dS^T Q_i  A = P^T / dS^T straight from tensor memory (TS MMAs): no shared-memory round trip
//   dQ_i = dS K                      the only product that needs dS in shared memory: the thread's row of dS^T is exactly the
//                                    MN-major A operand ([key rows][queries contiguous]); dQ_i leaves by red.global.add.v4
// lse2[q] / delta[q] are per-COLUMN values in this orientation.  Instead of 32 broadcast loads per thread and step they ride
// in the MMAs: one extra k-step multiplies an all-ones A block with a [queries x 16] B block holding -lse2/c (resp. -delta) as
// a three-term bf16 split (exact to fp32), so tensor memory already holds S^T - lse2/c and dP^T - delta.
// Per step this moves ~208 KiB through shared memory (TMA in 40, MMA operand reads 136, dS^T out 32) instead of 336.
//   warp 0 : TMA (K, V once; Q_i, dO_i and the [lse | delta] operand blocks through a 2-stage ring)
//   warp 1 : tcgen05 issuer            warps 2..17 : compute (key-row quarter = warp % 4, 16-query column group = (warp-2)/4)
//   warps 18..21 : drain dQ_i (TMEM -> swizzled staging boxes -> TMA reduce-add), off the compute warps' critical path
constexpr int kBwdThreads = 704;     // TMA warp, MMA warp, 16 compute warps, 4 dQ-drain warps
struct FbSmem {
  static constexpr int TILE = 128 * kHD * 2;            // 16 KiB: a [128 x 64] bf16 operand tile
  static constexpr int STAGES = 3;                      // Q_i / dO_i / [lse | delta] ring: two tiles of TMA latency in flight
  static constexpr int AUG = 2 * 2048;                  // [lse | delta] k-step operands of one query tile: 2 x [128 q][8 bf16]
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE;            // after the prologue (V lives in tensor memory): dQ staging boxes 0..3
  static constexpr int OFF_Q = OFF_V + TILE;
  static constexpr int OFF_DO = OFF_Q + STAGES * TILE;
  static constexpr int OFF_DS = OFF_DO + STAGES * TILE; // 2 buffers x [2 query atoms][128 key rows][128 B]
  static constexpr int OFF_AUG = OFF_DS + 4 * TILE;
  static constexpr int OFF_ZERO = OFF_AUG + STAGES * AUG;   // 1 KiB of zeros: second k-half of every [lse | delta] operand
  static constexpr int OFF_ONES = OFF_ZERO + 1024;      // [2 k-halves][128 rows][8 bf16] of 1.0
  static constexpr int OFF_DQ2 = OFF_ONES + 4096;       // dQ staging boxes 4..7 (32 rows x 32 fp32 each, 128B-swizzled)
  static constexpr int OFF_BAR = OFF_DQ2 + 4 * 4096;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};
static_assert(FbSmem::BYTES <= 227 * 1024, "flash backward shared memory");

struct FbParams {
  int B, N, H, Npad, T;
  float scale, scale_log2;
  const __nv_bfloat16* aug;                 // [B, H, T][lse | delta][128][8] (flash_bwd_prep_kernel)
  float* dq;                                // fp32 [B, N, H*64], zero-filled, accumulated by red.global.add
  __nv_bfloat16* dqkv; long long ld;        // [B*N, 3*H*64]
};

// shared-memory descriptor without swizzle, K-major: 8-row x 16-byte core matrices; lbo = byte distance between the two core
// matrices of a 16-element k-step, sbo = byte distance between consecutive 8-row groups
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const FbParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FbSmem::OFF_BAR);
  uint64_t* kv_full = bars;            // 1   K, V tiles landed
  uint64_t* kvt_full = bars + 1;       // 1   (16 warps) K, V copied to tensor memory, ones block written
  uint64_t* qdo_full = bars + 2;       // 3
  uint64_t* qdo_empty = bars + 5;      // 3
  uint64_t* sdp_full = bars + 8;       // 2   S^T / dP^T of a 64-query half are in tensor memory
  uint64_t* pds_full = bars + 10;      // 2   (16 warps) P^T / dS^T written back (+ dS^T in shared memory)
  uint64_t* dq_full = bars + 12;       // 1
  uint64_t* dq_empty = bars + 13;      // 1   (4 drain warps)
  uint64_t* acc_full = bars + 14;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int T = p.T;                     // 128-query tiles (= 128-key tiles)
  const int kt = blockIdx.x % T;
  const int bh = blockIdx.x / T;
  const int h = bh % p.H, b = bh / p.H;
  const int k0 = kt * 128;
  const int U = 2 * T;                   // 64-query half steps
  const int D = p.H * kHD;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1);
      mbar_init(kvt_full, 16);
      for (int i = 0; i < FbSmem::STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&pds_full[i], 16); }
      mbar_init(dq_full, 1); mbar_init(dq_empty, 4);
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_dk = tmem, tm_dv = tmem + 64, tm_dq = tmem + 384, tm_kt = tmem + 448, tm_vt = tmem + 480;
  pdl_wait();              // barrier init / TMEM allocation overlapped the previous kernel's tail
  // S^T -> P^T and dP^T -> dS^T (bf16 pairs over the first 8 of every 16 fp32 columns), double-buffered per 64-query half
  auto tm_st_ = [&](int buf) { return tmem + 128 + 64 * buf; };
  auto tm_dpt_ = [&](int buf) { return tmem + 256 + 64 * buf; };

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * FbSmem::TILE);
      tma_load_3d(smem + FbSmem::OFF_K, &tmQKV, kv_full, D + h * kHD, k0, b);
      tma_load_3d(smem + FbSmem::OFF_V, &tmQKV, kv_full, 2 * D + h * kHD, k0, b);
      const __nv_bfloat16* aug = p.aug + (long long)bh * T * (FbSmem::AUG / 2);
      for (int i = 0; i < T; ++i) {
        const int st = i % FbSmem::STAGES;
        mbar_wait(&qdo_empty[st], ((i / FbSmem::STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * FbSmem::TILE + FbSmem::AUG);
        tma_load_3d(smem + FbSmem::OFF_Q + st * FbSmem::TILE, &tmQKV, &qdo_full[st], h * kHD, i * 128, b);
        tma_load_3d(smem + FbSmem::OFF_DO + st * FbSmem::TILE, &tmDO, &qdo_full[st], h * kHD, i * 128, b);
        bulk_load_1d(smem + FbSmem::OFF_AUG + st * FbSmem::AUG, aug + (long long)i * (FbSmem::AUG / 2), FbSmem::AUG, &qdo_full[st]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, false, false);   // [keys x 64 queries]: A K-major (TMEM / ones), B = Q / dO K-major
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, kHD, false, true);  // [keys x hd]: A = P^T / dS^T from TMEM, B = dO / Q MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHD, true, true);    // [queries x hd]: A = dS^T read MN-major, B = K MN-major
      const uint32_t sk = smem_u32(smem + FbSmem::OFF_K);
      const uint64_t ones = umma_desc_nosw(smem_u32(smem + FbSmem::OFF_ONES), 2048, 128);
      const uint32_t zero = smem_u32(smem + FbSmem::OFF_ZERO);
      auto issue_sdp = [&](int u) {
        const int i = u >> 1, hf = u & 1, st = i % FbSmem::STAGES, buf = u & 1;
        if (hf == 0) {
          mbar_wait(&qdo_full[st], (i / FbSmem::STAGES) & 1);
          tc_fence_after();
        }
        const uint64_t bq = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_Q + st * FbSmem::TILE) + hf * 8192, 0, 1024);
        const uint64_t bdo = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_DO + st * FbSmem::TILE) + hf * 8192, 0, 1024);
        const uint32_t aug = smem_u32(smem + FbSmem::OFF_AUG + st * FbSmem::AUG) + hf * 1024;   // k-half 0; k-half 1 = the zero block
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tm_st_(buf), tm_kt + 8 * k, bq + 2 * k, idesc_t, k > 0 ? 1u : 0u);
        umma_ss(tm_st_(buf), ones, umma_desc_nosw(aug, zero - aug, 128), idesc_t, 1u);               // - lse2[q] / c
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tm_dpt_(buf), tm_vt + 8 * k, bdo + 2 * k, idesc_t, k > 0 ? 1u : 0u);
        umma_ss(tm_dpt_(buf), ones, umma_desc_nosw(aug + 2048, zero - aug - 2048, 128), idesc_t, 1u);  // - delta[q]
        umma_commit(&sdp_full[buf]);
      };
      mbar_wait(kvt_full, 0);
      tc_fence_after();
      issue_sdp(0);
      issue_sdp(1);
      for (int u = 0; u < U; ++u) {
        const int i = u >> 1, hf = u & 1, st = i % FbSmem::STAGES, buf = u & 1;
        mbar_wait(&pds_full[buf], (u >> 1) & 1);
        tc_fence_after();
        // B operands MN-major over the 64 query rows of this half (16 rows per k-step)
        const uint64_t bdo = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_DO + st * FbSmem::TILE) + hf * 8192, 128 * 128, 1024);
        const uint64_t bq = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_Q + st * FbSmem::TILE) + hf * 8192, 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tm_dv, tm_st_(buf) + 16 * k, bdo + 128 * k, idesc_kv, (u > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tm_dk, tm_dpt_(buf) + 16 * k, bq + 128 * k, idesc_kv, (u > 0 || k > 0) ? 1u : 0u);
        if (hf == 1) umma_commit(&qdo_empty[st]);   // Q_i / dO_i have had their last reader: the stage is refilled while dQ_i runs
        if (u + 2 < U) issue_sdp(u + 2);            // overwrites P^T / dS^T of this half: queued behind the dV / dK MMAs above
        if (hf == 1) {
          // dQ_i[q, hd] = sum_keys dS[q, key] K[key, hd]: A = dS^T tile read MN-major (M = queries: 2 atoms; K = key rows), B = K MN-major
          if (i > 0) {
            mbar_wait(dq_empty, (i - 1) & 1);
            tc_fence_after();
          }
          const uint64_t ads = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_DS + (i & 1) * 2 * FbSmem::TILE), FbSmem::TILE, 1024);
          const uint64_t bk = umma_desc_sw128(sk, 128 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tm_dq, ads + 128 * k, bk + 128 * k, idesc_q, k > 0 ? 1u : 0u);
          umma_commit(dq_full);
        }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else if (warp >= 18) {
    // ---- dQ drain: rows of this warp's lane quarter, all 64 columns as two [32 rows x 32 fp32] boxes: TMEM -> 128B-swizzled box
    // -> TMA reduce-add into the fp32 buffer (full 128-byte lines at L2; per-thread red.global.add.v4 of 16-byte pieces
    // quadrupled the L2 request count and stalled the issuing warps on the LSU queue)
    const int dw = warp - 18, q4 = warp & 3, lane = (int)lane_id();
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    // boxes 0..3 reuse the V tile (dead after the prologue), 4..7 have their own 16 KiB
    uint8_t* box0_p = smem + (dw < 2 ? FbSmem::OFF_V + 2 * dw * 4096 : FbSmem::OFF_DQ2 + (2 * dw - 4) * 4096);
    uint8_t* box1_p = box0_p + 4096;
    const uint32_t b0 = smem_u32(box0_p) + lane * 128, b1 = b0 + 4096;
    mbar_wait(kvt_full, 0);                    // V has been copied to tensor memory: its tile may be overwritten
    for (int i = 0; i < T; ++i) {
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t t[32];
      tmem_ld_32x32b_x32(tm_dq + lane_sel, t);
      tmem_ld_wait();
      if (lane == 0) bulk_wait_read0();        // the previous reductions have finished reading the boxes
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(b0 + ((c ^ (lane & 7)) << 4), t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
      tmem_ld_32x32b_x32(tm_dq + lane_sel + 32, t);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);    // the accumulator may be overwritten by dQ_{i+1}
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(b1 + ((c ^ (lane & 7)) << 4), t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmDQ, box0_p, h * kHD, i * 128 + q4 * 32, b);
        tma_reduce_add_3d(&tmDQ, box1_p, h * kHD + 32, i * 128 + q4 * 32, b);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  } else {
    const int cw = warp - 2;                 // 0..15
    const int q4 = warp & 3;                 // TMEM lane quarter = key-row quarter
    const int cg = cw >> 2;                  // 16-column group
    const int lane = (int)lane_id();
    const int r = q4 * 32 + lane;            // key row of this thread
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t sw0 = (uint32_t)(((2 * cg) ^ (r & 7)) << 4), sw1 = (uint32_t)(((2 * cg + 1) ^ (r & 7)) << 4);
    const uint64_t cc = f2_pack(p.scale_log2, p.scale_log2);

    // ---- prologue: K and V rows into tensor memory (A operands of S^T / dP^T), the all-ones k-step block into shared memory
    {
      const int t = cw * 32 + lane;
      if (t < 256) sts128(smem_u32(smem + FbSmem::OFF_ONES) + t * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      else if (t < 320) sts128(smem_u32(smem + FbSmem::OFF_ZERO) + (t - 256) * 16, 0u, 0u, 0u, 0u);
      mbar_wait(kv_full, 0);
      const uint32_t krow = smem_u32(smem + FbSmem::OFF_K) + r * 128, vrow = smem_u32(smem + FbSmem::OFF_V) + r * 128;
      const float4 ka = lds128(krow + sw0), kb = lds128(krow + sw1), va = lds128(vrow + sw0), vb = lds128(vrow + sw1);
      const uint32_t kw[8] = {__float_as_uint(ka.x), __float_as_uint(ka.y), __float_as_uint(ka.z), __float_as_uint(ka.w),
                              __float_as_uint(kb.x), __float_as_uint(kb.y), __float_as_uint(kb.z), __float_as_uint(kb.w)};
      const uint32_t vw[8] = {__float_as_uint(va.x), __float_as_uint(va.y), __float_as_uint(va.z), __float_as_uint(va.w),
                              __float_as_uint(vb.x), __float_as_uint(vb.y), __float_as_uint(vb.z), __float_as_uint(vb.w)};
      tmem_st_32x32b_x8(tm_kt + lane_sel + 8 * cg, kw);
      tmem_st_32x32b_x8(tm_vt + lane_sel + 8 * cg, vw);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(kvt_full);
    }
    for (int u = 0; u < U; ++u) {
      const int i = u >> 1, hf = u & 1, buf = u & 1;
      mbar_wait(&sdp_full[buf], (u >> 1) & 1);
      tc_fence_after();
      uint32_t sv_[16], dp_[16], pp[8], dd[8];
      tmem_ld_32x32b_x16(tm_st_(buf) + lane_sel + 16 * cg, sv_);
      tmem_ld_32x32b_x16(tm_dpt_(buf) + lane_sel + 16 * cg, dp_);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t x = f2_mul(f2_pack(__uint_as_float(sv_[2 * k]), __uint_as_float(sv_[2 * k + 1])), cc);
        // MUFU.EX2 (16 per clock and SM) alone would cost 1024 cycles per 128 x 128 step: three of every eight pairs go through
        // the polynomial on the FMA pipe instead
        const uint64_t pv = (k == 1 || k == 4 || k == 6) ? exp2_poly_pair(x) : f2_pack(fast_exp2(f2_lo(x)), fast_exp2(f2_hi(x)));
        const float p0 = f2_lo(pv), p1 = f2_hi(pv);
        const uint64_t ds = f2_mul(pv, f2_pack(__uint_as_float(dp_[2 * k]), __uint_as_float(dp_[2 * k + 1])));
        pp[k] = pack_bf16(p0, p1);
        dd[k] = pack_bf16(f2_lo(ds), f2_hi(ds));
      }
      tmem_st_32x32b_x8(tm_st_(buf) + lane_sel + 16 * cg, pp);
      tmem_st_32x32b_x8(tm_dpt_(buf) + lane_sel + 16 * cg, dd);
      // dS^T row of this key, queries [64 hf + 16 cg, +16): the MN-major A operand of dQ_i = dS K
      const uint32_t dsrow = smem_u32(smem + FbSmem::OFF_DS + (i & 1) * 2 * FbSmem::TILE + hf * FbSmem::TILE) + r * 128;
      sts128(dsrow + sw0, dd[0], dd[1], dd[2], dd[3]);
      sts128(dsrow + sw1, dd[4], dd[5], dd[6], dd[7]);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[buf]);
    }
    // epilogue: dK (scaled) and dV rows of this key tile (thread = key row; 16 of the 64 columns of each per warp)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int ki = k0 + r;
    uint32_t tk[16], tv[16];
    tmem_ld_32x32b_x16(tm_dk + lane_sel + 16 * cg, tk);
    tmem_ld_32x32b_x16(tm_dv + lane_sel + 16 * cg, tv);
    tmem_ld_wait();
    if (ki < p.N) {
      __nv_bfloat16* row = p.dqkv + ((long long)b * p.N + ki) * p.ld + h * kHD + 16 * cg;
      uint4* dk_dst = reinterpret_cast<uint4*>(row + D);
      uint4* dv_dst = reinterpret_cast<uint4*>(row + 2 * D);
      const float sc = p.scale;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        dk_dst[c] = make_uint4(pack_bf16(sc * __uint_as_float(tk[8 * c]), sc * __uint_as_float(tk[8 * c + 1])),
                               pack_bf16(sc * __uint_as_float(tk[8 * c + 2]), sc * __uint_as_float(tk[8 * c + 3])),
                               pack_bf16(sc * __uint_as_float(tk[8 * c + 4]), sc * __uint_as_float(tk[8 * c + 5])),
                               pack_bf16(sc * __uint_as_float(tk[8 * c + 6]), sc * __uint_as_float(tk[8 * c + 7])));
        dv_dst[c] = make_uint4(pack_bf16(__uint_as_float(tv[8 * c]), __uint_as_float(tv[8 * c + 1])),
                               pack_bf16(__uint_as_float(tv[8 * c + 2]), __uint_as_float(tv[8 * c + 3])),
                               pack_bf16(__uint_as_float(tv[8 * c + 4]), __uint_as_float(tv[8 * c + 5])),
                               pack_bf16(__uint_as_float(tv[8 * c + 6]), __uint_as_float(tv[8 * c + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// Operand preparation for the backward: delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d] and the forward's lse2[b,h,q], written as
// the extra k-step operand blocks the MMAs read ([B, H, T][lse | delta][128 queries][8 bf16]: the first k-half in no-swizzle
// core-matrix order, the second k-half is a shared block of zeros): row q holds (hi, mid, lo, 0, ...) with hi + mid + lo = -lse2[q] / c (resp. -delta[q]) exact to fp32.  Padded queries
// (q >= N) get -30000 in the lse block, so their probabilities vanish.
__device__ __forceinline__ uint4 split3_bf16(float v) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const float r1 = v - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16(r1);
  const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
  const uint32_t w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
  const uint32_t w1 = (uint32_t)__bfloat16_as_ushort(lo);
  return make_uint4(w0, w1, 0u, 0u);
}
// One CTA = 32 consecutive queries of a clip x all heads: 8 lanes share a (query, head) row of 64 channels (one 16-byte load of
// O and dO each, 512 contiguous bytes per warp and instruction, four rows in flight per lane), the 32 x H sums go through
// shared memory so that lse2 is read and the operand blocks are written along q (512 contiguous bytes per head).  (Thread =
// (query, head) with 128-byte strides between lanes and scattered 16-byte writes ran at 3.1 TB/s of DRAM traffic.)
constexpr int kPrepQ = 32, kPrepMaxH = 16;
__global__ void __launch_bounds__(256) flash_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                             const float* __restrict__ lse2, __nv_bfloat16* __restrict__ aug, int B,
                                                             int N, int H, int Npad, float inv_c) {
  pdl_trigger();
  pdl_wait();
  __shared__ float delta_s[kPrepQ * kPrepMaxH];                  // [q][h]
  const int tiles_q = Npad / kPrepQ;
  const int b = blockIdx.x / tiles_q, q0 = (blockIdx.x % tiles_q) * kPrepQ;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int nchunk = kPrepQ * H;                                 // (q, h) rows of this CTA, h fastest = memory order
  const long long base = ((long long)b * N + q0) * (H * kHD) + sub * 8;
  for (int c0 = 0; c0 < nchunk; c0 += 128) {
    uint4 a[4], d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32 + grp;
      a[u] = make_uint4(0u, 0u, 0u, 0u);
      d[u] = a[u];
      if (c < nchunk && q0 + c / H < N) {
        a[u] = __ldg(reinterpret_cast<const uint4*>(o + base + (long long)c * kHD));
        d[u] = __ldg(reinterpret_cast<const uint4*>(dout + base + (long long)c * kHD));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 a0 = unpack_bf16(a[u].x), a1 = unpack_bf16(a[u].y), a2 = unpack_bf16(a[u].z), a3 = unpack_bf16(a[u].w);
      const float2 d0 = unpack_bf16(d[u].x), d1 = unpack_bf16(d[u].y), d2 = unpack_bf16(d[u].z), d3 = unpack_bf16(d[u].w);
      float s = (a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y) + (a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      const int c = c0 + u * 32 + grp;
      if (sub == 0 && c < nchunk) delta_s[c] = s;
    }
  }
  __syncthreads();
  const int T = Npad / 128;
  for (int t = threadIdx.x; t < nchunk; t += blockDim.x) {
    const int ql = t & (kPrepQ - 1), h = t / kPrepQ, q = q0 + ql;
    float s = 0.f, nl = -30000.f;
    if (q < N) {
      s = delta_s[ql * H + h];
      nl = -__ldg(lse2 + ((long long)b * H + h) * Npad + q) * inv_c;
    }
    uint4* tile = reinterpret_cast<uint4*>(aug + ((((long long)b * H + h) * T + (q >> 7)) * 2048));   // 256 uint4 per tile pair
    const int row = q & 127;
    tile[row] = split3_bf16(nl);          // lse block   (first k-half; the second one is all zeros and shared)
    tile[128 + row] = split3_bf16(-s);    // delta block
  }
}

// dqkv[:, 0:D] = bf16(scale * dq_acc)
__global__ void __launch_bounds__(256) flash_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv,
                                                               long long rows, int D, long long ld, float scale) {
  pdl_trigger();
  pdl_wait();
  const long long n8 = rows * (D / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (D / 8);
    const int c8 = (int)(i % (D / 8));
    const float4 a = __ldcs(reinterpret_cast<const float4*>(acc + row * D + c8 * 8));
    const float4 c = __ldcs(reinterpret_cast<const float4*>(acc + row * D + c8 * 8) + 1);
    *reinterpret_cast<uint4*>(dqkv + row * ld + c8 * 8) =
        make_uint4(pack_bf16(a.x * scale, a.y * scale), pack_bf16(a.z * scale, a.w * scale), pack_bf16(c.x * scale, c.y * scale),
                   pack_bf16(c.z * scale, c.w * scale));
  }
}

static int make_qkv_tmap(CUtensorMap* tm, const void* qkv, int B, int N, int D3, uint32_t box_rows) {
  const uint64_t dims[3] = {(uint64_t)D3, (uint64_t)N, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)D3 * 2, (uint64_t)N * D3 * 2};
  const uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_nd(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace dv

extern "C" int devias_flash_attn_fwd(const void* qkv, void* out, float* lse2, int batch, int seq, int heads, int head_dim,
                                     float scale, void* stream) {
  using namespace dv;
  DV_REQUIRE(qkv && out, "null pointer");
  DV_REQUIRE(head_dim == 64, "head_dim 64 only (ViT-B/16)");
  DV_REQUIRE(batch > 0 && seq > 0 && heads > 0, "empty problem");
  const int D = heads * kHD;
  CUtensorMap tmQ, tmKV;
  int rc = make_qkv_tmap(&tmQ, qkv, batch, seq, 3 * D, kQT);
  if (rc) return rc;
  rc = make_qkv_tmap(&tmKV, qkv, batch, seq, 3 * D, kKT);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa2Smem::BYTES));
    attr_done = true;
  }
  FaParams p{batch, seq, heads, (seq + 127) / 128 * 128, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out),
             (long long)D, lse2};
  const int q_tiles = (seq + kQT - 1) / kQT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int prof = prof_begin(DEVIAS_PROF_ATTN, 4.0 * batch * heads * (double)seq * seq * kHD, s);
  DV_CHECK_CUDA(launch_k(flash_fwd2_kernel, dim3((unsigned)(batch * heads * q_tiles)), dim3((unsigned)(kFa2Threads)), (size_t)(Fa2Smem::BYTES), s, tmQ, tmKV, p));
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_flash_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                                     void* aug_ws, float* dq_ws, int batch, int seq, int heads, int head_dim, float scale,
                                     void* stream) {
  using namespace dv;
  DV_REQUIRE(qkv && out && dout && lse2 && dqkv && aug_ws && dq_ws, "null pointer");
  DV_REQUIRE(head_dim == 64, "head_dim 64 only (ViT-B/16)");
  DV_REQUIRE(batch > 0 && seq > 0 && heads > 0, "empty problem");
  DV_REQUIRE((reinterpret_cast<uintptr_t>(aug_ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(dq_ws) & 15) == 0, "16-byte alignment");
  const int D = heads * kHD;
  const int Npad = (seq + 127) / 128 * 128;
  const int T = Npad / 128;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tmQKV, tmDO;
  int rc = make_qkv_tmap(&tmQKV, qkv, batch, seq, 3 * D, 128);
  if (rc) return rc;
  rc = make_qkv_tmap(&tmDO, dout, batch, seq, D, 128);
  if (rc) return rc;
  CUtensorMap tmDQ;
  {  // fp32 dQ accumulator [B, N, H*64]: boxes of 32 floats x 32 rows; rows past N are clipped by the TMA unit
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)seq, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)D * 4, (uint64_t)seq * D * 4};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap_nd(&tmDQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dq_ws, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FbSmem::BYTES));
    attr_done = true;
  }
  DV_CHECK_CUDA(cudaMemsetAsync(dq_ws, 0, (size_t)batch * seq * D * sizeof(float), s));
  const float scale_log2 = scale * 1.4426950408889634f;
  {
    DV_REQUIRE(heads <= kPrepMaxH, "flash_attn_bwd: at most 16 heads");
    DV_CHECK_CUDA(launch_k(flash_bwd_prep_kernel, dim3((unsigned)(batch * (Npad / kPrepQ))), dim3((unsigned)(256)), (size_t)(0), s, static_cast<const __nv_bfloat16*>(out),
                                                                     static_cast<const __nv_bfloat16*>(dout), lse2,
                                                                     static_cast<__nv_bfloat16*>(aug_ws), batch, seq, heads, Npad,
                                                                     1.0f / scale_log2));
  }
  FbParams p{batch, seq, heads, Npad, T, scale, scale_log2, static_cast<const __nv_bfloat16*>(aug_ws), dq_ws,
             static_cast<__nv_bfloat16*>(dqkv), (long long)3 * D};
  const int prof = prof_begin(DEVIAS_PROF_ATTN, 10.0 * batch * heads * (double)seq * seq * kHD, s);
  DV_CHECK_CUDA(launch_k(flash_bwd_kernel, dim3((unsigned)(batch * heads * T)), dim3((unsigned)(kBwdThreads)), (size_t)(FbSmem::BYTES), s, tmQKV, tmDO, tmDQ, p));
  prof_end(prof, s);
  {
    const long long n8 = (long long)batch * seq * (D / 8);
    long long blocks = (n8 + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    DV_CHECK_CUDA(launch_k(flash_dq_convert_kernel, dim3((unsigned)((int)blocks)), dim3((unsigned)(256)), (size_t)(0), s, dq_ws, static_cast<__nv_bfloat16*>(dqkv), (long long)batch * seq, D,
                                                        (long long)3 * D, scale));
  }
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return DEVIAS_OK;
}
