// Fused softmax attention for the encoder (model/modeling_slot.py:102-112): out = softmax(scale * Q K^T) V per (clip, head),
// head_dim 64, sequence 1568 (any N), no mask, no dropout -- the 12 x N x N probability matrix never touches HBM.
//
// Forward kernel, one CTA per (clip, head, 128-query tile), two CTAs resident per SM:
//   warp 0      : TMA producer   (Q once; K_j / V_j 64-key tiles through a 3-stage ring; 3-D tensor map over the packed
//                                 qkv activation [B, N, 3*H*64] so rows past N are zero-filled)
//   warp 1      : tcgen05 issuer (S_j = Q K_j^T -> TMEM;  O_j = P_j V_j -> TMEM;  S_{j+1} is issued before P_j is awaited)
//   warps 2..5  : softmax        (thread = query row: TMEM -> registers, online max/sum in the exp2 domain, P_j -> bf16 ->
//                                 128B-swizzled smem as the A operand of the second MMA, running O in registers)
// Backward kernels live below (dQ / dK / dV with recomputed probabilities).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kHD = 64;        // head dim
constexpr int kQT = 128;       // queries per CTA
constexpr int kKT = 64;        // keys per tile
constexpr int kKVStages = 3;
constexpr int kFaThreads = 192;

struct FaSmem {
  static constexpr int Q_BYTES = kQT * kHD * 2;        // 16 KiB
  static constexpr int K_BYTES = kKT * kHD * 2;        // 8 KiB
  static constexpr int V_BYTES = kKT * kHD * 2;        // 8 KiB
  static constexpr int P_BYTES = kQT * kKT * 2;        // 16 KiB
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + kKVStages * K_BYTES;
  static constexpr int OFF_P = OFF_V + kKVStages * V_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};

struct FaParams {
  int B, N, H;
  int Npad;                         // row length of lse2 / delta: N rounded up to a multiple of 128
  float scale_log2;                 // head_dim^-0.5 * log2(e)
  __nv_bfloat16* out; long long ldo;  // [B*N, H*64]
  float* lse2;                      // [B, H, N]  log2-domain log-sum-exp of the scaled scores
};

__global__ void __launch_bounds__(kFaThreads, 2)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FaSmem::OFF_BAR);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // 3
  uint64_t* kv_empty = bars + 4;           // 3
  uint64_t* s_full = bars + 7;             // 2
  uint64_t* s_empty = bars + 9;            // 2
  uint64_t* p_full = bars + 11;            // 2
  uint64_t* p_empty = bars + 13;           // 2
  uint64_t* pv_done = bars + 15;           // 4-deep ring: pv_done[j % 4] completes when P_j V_j has been accumulated into O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int q_tiles = (p.N + kQT - 1) / kQT;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * kQT;
  const int T = (p.N + kKT - 1) / kKT;     // key tiles
  const int D = p.H * kHD;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmKV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int i = 0; i < kKVStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
        mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
      }
      for (int i = 0; i < 4; ++i) mbar_init(&pv_done[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_s[2] = {tmem, tmem + 64};
  const uint32_t tm_o = tmem + 128;        // running (unnormalised) output, accumulated by the MMAs across key tiles

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, FaSmem::Q_BYTES);
      tma_load_3d(smem + FaSmem::OFF_Q, &tmQ, q_full, h * kHD, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j % kKVStages;
        mbar_wait(&kv_empty[st], ((j / kKVStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], FaSmem::K_BYTES + FaSmem::V_BYTES);
        tma_load_3d(smem + FaSmem::OFF_K + st * FaSmem::K_BYTES, &tmKV, &kv_full[st], D + h * kHD, j * kKT, b);
        tma_load_3d(smem + FaSmem::OFF_V + st * FaSmem::V_BYTES, &tmKV, &kv_full[st], 2 * D + h * kHD, j * kKT, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kQT, kKT, false, false);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(kQT, kHD, false, true);    // O = P V   : A K-major, B (=V) MN-major
      const uint32_t sq = smem_u32(smem + FaSmem::OFF_Q);
      auto issue_s = [&](int j) {
        const int st = j % kKVStages;
        mbar_wait(&kv_full[st], (j / kKVStages) & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(sq, 0, 1024);
        const uint64_t db = umma_desc_sw128(smem_u32(smem + FaSmem::OFF_K + st * FaSmem::K_BYTES), 0, 1024);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(tm_s[j & 1], da + 2 * k, db + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j % kKVStages;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);   // P_j is in TMEM and any rescaling of O (tcgen05.st) has been fenced
        tc_fence_after();
        // A = P_j straight from tensor memory (bf16 pairs packed over the first 32 columns of the S_j buffer): the
        // shared-memory port, which bounds this head_dim-64 kernel, only carries the V operand
        const uint64_t db = umma_desc_sw128(smem_u32(smem + FaSmem::OFF_V + st * FaSmem::V_BYTES), kKT * 128, 1024);
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k) umma_ts(tm_o, tm_s[j & 1] + 8 * k, db + 128 * k, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&pv_done[j & 3]);
        umma_commit(&kv_empty[st]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int lane = (int)lane_id();
    const int r = q * 32 + lane;                     // query row inside the tile == TMEM lane
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    // Lazy rescaling: exponentials are taken relative to a reference maximum `m` that is only moved (and the TMEM
    // accumulator rescaled) when the running maximum exceeds it by more than 8 (log2 units), i.e. P <= 256.
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row_base = smem_u32(smem + FaSmem::OFF_P + r * 128);
    const uint64_t cc = f2_pack(p.scale_log2, p.scale_log2);
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(tm_s[j & 1] + lane_sel, s0);
      tmem_ld_32x32b_x32(tm_s[j & 1] + lane_sel + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      const int valid = p.N - j * kKT;               // keys valid in this tile (>= 64 except the last)
      if (valid < kKT) {                             // warp-uniform: only the ragged last tile masks
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) s0[i] = 0xff800000u;       // -inf
          if (i + 32 >= valid) s1[i] = 0xff800000u;
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains (latency, not throughput, matters)
#pragma unroll
      for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_run = fmaxf(m, mx * p.scale_log2);   // scale > 0: max(c s) = c max(s)
      if (__any_sync(0xffffffffu, m_run > m + 8.0f)) {
        const float alpha = fast_exp2(m - m_run);       // 0 on the first tile (m = -inf)
        if (j > 0) {
          // P_{j-1} V_{j-1} must have landed in O.  pv_done is a 4-deep ring, so its parity cannot alias: the MMA warp is
          // never more than one tile ahead of this warp (it needs our p_full arrival for tile j).
          mbar_wait(&pv_done[(j - 1) & 3], ((j - 1) >> 2) & 1);
          tc_fence_after();
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(tm_o + lane_sel + 32 * hlf, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st_32x32b_x32(tm_o + lane_sel + 32 * hlf, t);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m = m_run;
      }
      const uint64_t negm = f2_pack(-m, -m);
      // P_j (bf16 pairs) -> tensor memory, over the S_j columns this thread has already consumed (lane = query row)
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = 8 * (c & 3) + 2 * i;
          const uint64_t x = c < 4 ? f2_pack(__uint_as_float(s0[e]), __uint_as_float(s0[e + 1]))
                                   : f2_pack(__uint_as_float(s1[e]), __uint_as_float(s1[e + 1]));
          const uint64_t y = f2_fma(x, cc, negm);
          const float e0 = fast_exp2(f2_lo(y)), e1 = fast_exp2(f2_hi(y));
          sum2[i] = f2_add(sum2[i], f2_pack(e0, e1));
          pk[4 * c + i] = pack_bf16(e0, e1);
        }
      }
      tmem_st_32x32b_x32(tm_s[j & 1] + lane_sel, pk);
      const uint64_t st = f2_add(f2_add(sum2[0], sum2[1]), f2_add(sum2[2], sum2[3]));
      l += f2_lo(st) + f2_hi(st);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    float o[kHD];
    {
      mbar_wait(&pv_done[(T - 1) & 3], ((T - 1) >> 2) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(tm_o + lane_sel + 32 * hlf, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[32 * hlf + i] = __uint_as_float(t[i]) * inv;
      }
    }
    const int qi = q0 + r;
    if (qi < p.N) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.N + qi) * p.ldo + h * kHD);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        dst[c] = make_uint4(pack_bf16(o[8 * c], o[8 * c + 1]), pack_bf16(o[8 * c + 2], o[8 * c + 3]),
                            pack_bf16(o[8 * c + 4], o[8 * c + 5]), pack_bf16(o[8 * c + 6], o[8 * c + 7]));
      if (p.lse2 != nullptr) p.lse2[((long long)b * p.H + h) * p.Npad + qi] = m + log2f(l);
    } else if (p.lse2 != nullptr && qi < p.Npad) {
      p.lse2[((long long)b * p.H + h) * p.Npad + qi] = INFINITY;   // padded queries: exp2(s - inf) = 0 in the backward
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}


// =====================================================================================================================
// Forward, second generation: FOUR small CTAs per SM instead of two big ones.
// ncu on the first kernel: MUFU (ex2) 48 % and tensor pipe 21 % busy, one softmax warp per scheduler and CTA stalled on its
// own serial chain (TMEM load -> max -> 64 ex2 -> pack -> TMEM store -> barrier), i.e. latency bound with two CTAs per SM.
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
// Backward.  One CTA per (clip, head, 128-key tile); loop over 128-query tiles.  dK/dV accumulate in TMEM for the whole
// loop; each iteration's dQ partial goes TMEM -> smem -> TMA reduce-add into an fp32 [B, N, H*64] buffer (L2 resident).
//   S = Q_i K^T, dP = dO_i V^T                 (TMEM, thread = query row, so lse2[q] / delta[q] are per-thread scalars)
//   P = exp2(c S - lse2[q]),  dS = P * (dP - delta[q])       -> bf16 -> smem, K-major over the keys (128B swizzle)
//   dV += P^T dO_i,  dK += dS^T Q_i   (the P / dS tiles are read as MN-major A operands),   dQ_i = dS K
// The softmax scale is applied to dK in the epilogue and to dQ in the fp32 -> bf16 conversion kernel.
//   warp 0 : TMA (K,V once; Q_i, dO_i through a 2-stage ring)   warp 1 : tcgen05 issuer
//   warps 2..17 : compute (query-row quarter = warp % 4, 32-key column group = (warp - 2) / 4): four warps per scheduler
//                 hide the TMEM / MUFU / shared-memory latencies of the recompute chain (8 warps left it latency-bound)
constexpr int kBwdThreads = 576;   // TMA warp + MMA warp + 16 compute warps (4 per TMEM lane quarter)
struct FbSmem {
  static constexpr int TILE = 128 * kHD * 2;            // 16 KiB: a [128 x 64] bf16 operand tile
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE;
  static constexpr int OFF_Q = OFF_V + TILE;            // 2 stages
  static constexpr int OFF_DO = OFF_Q + 2 * TILE;       // 2 stages
  static constexpr int OFF_P = OFF_DO + 2 * TILE;       // P  [128 queries x 128 keys] = 2 key atoms x 16 KiB
  static constexpr int OFF_DS = OFF_P + 2 * TILE;       // dS
  static constexpr int OFF_DQ = OFF_DS + 2 * TILE;      // dQ staging: 8 warps x (32 rows x 32 fp32)
  static constexpr int OFF_BAR = OFF_DQ + 8 * 4096;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};

struct FbParams {
  int B, N, H, Npad;
  float scale, scale_log2;
  const float* lse2; const float* delta;   // [B, H, Npad]
  __nv_bfloat16* dqkv; long long ld;        // [B*N, 3*H*64]
  long long* dbg;                           // DEVIAS_FLASH_DEBUG=1: per-phase SM clock stamps of CTA 0 (timeline debugging)
};
#define FB_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x == 0 && i < 8) p.dbg[(slot) * 8 + i] = clock64(); } while (0)

__global__ void __launch_bounds__(kBwdThreads, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const FbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FbSmem::OFF_BAR);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 1;       // 2
  uint64_t* qdo_empty = bars + 3;      // 2
  uint64_t* sdp_full = bars + 5;       // 1
  uint64_t* sdp_empty = bars + 6;      // 1 (8 warps)
  uint64_t* pds_full = bars + 7;       // 1 (8 warps)
  uint64_t* pds_empty = bars + 8;      // 1
  uint64_t* dq_full = bars + 9;        // 2
  uint64_t* dq_empty = bars + 11;      // 2 (8 warps)
  uint64_t* acc_full = bars + 13;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int k_tiles = (p.N + 127) / 128;
  const int kt = blockIdx.x % k_tiles;
  const int bh = blockIdx.x / k_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int k0 = kt * 128;
  const int T = (p.N + 127) / 128;       // query tiles
  const int D = p.H * kHD;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1);
        mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 8);
      }
      mbar_init(sdp_full, 1); mbar_init(sdp_empty, 16);
      mbar_init(pds_full, 16); mbar_init(pds_empty, 1);
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
  const uint32_t tm_dk = tmem, tm_dv = tmem + 64, tm_s = tmem + 128, tm_dp = tmem + 256;
  const uint32_t tm_dq[2] = {tmem + 384, tmem + 448};

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * FbSmem::TILE);
      tma_load_3d(smem + FbSmem::OFF_K, &tmQKV, kv_full, D + h * kHD, k0, b);
      tma_load_3d(smem + FbSmem::OFF_V, &tmQKV, kv_full, 2 * D + h * kHD, k0, b);
      for (int i = 0; i < T; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * FbSmem::TILE);
        tma_load_3d(smem + FbSmem::OFF_Q + st * FbSmem::TILE, &tmQKV, &qdo_full[st], h * kHD, i * 128, b);
        tma_load_3d(smem + FbSmem::OFF_DO + st * FbSmem::TILE, &tmDO, &qdo_full[st], h * kHD, i * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);   // [queries x keys] = Q K^T / dO V^T
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, kHD, true, true);    // [keys x hd]: A = P / dS read MN-major, B = dO / Q MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHD, false, true);    // [queries x hd]: A = dS K-major, B = K MN-major
      const uint32_t sk = smem_u32(smem + FbSmem::OFF_K), sv = smem_u32(smem + FbSmem::OFF_V);
      const uint32_t sp = smem_u32(smem + FbSmem::OFF_P), sds = smem_u32(smem + FbSmem::OFF_DS);
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        mbar_wait(&qdo_full[st], (i >> 1) & 1);
        mbar_wait(sdp_empty, (i & 1) ^ 1);
        tc_fence_after();
        const uint64_t dq_ = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_Q + st * FbSmem::TILE), 0, 1024);
        const uint64_t ddo = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_DO + st * FbSmem::TILE), 0, 1024);
        const uint64_t dk_ = umma_desc_sw128(sk, 0, 1024), dv_ = umma_desc_sw128(sv, 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm_s, dq_ + 2 * k, dk_ + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm_dp, ddo + 2 * k, dv_ + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int i = 0; i < T; ++i) {
        FB_STAMP(0);
        if (i + 1 < T) issue_sdp(i + 1);
        FB_STAMP(1);
        const int st = i & 1;
        mbar_wait(pds_full, i & 1);
        FB_STAMP(2);
        tc_fence_after();
        // P / dS tiles: [128 query rows][2 key atoms of 64]; read MN-major: M = keys (atoms 16 KiB apart), K = query rows
        const uint64_t ap_mn = umma_desc_sw128(sp, FbSmem::TILE, 1024), ads_mn = umma_desc_sw128(sds, FbSmem::TILE, 1024);
        // B operands, MN-major over the 128 query rows of this stage (one 64-wide atom, 16 rows per UMMA_K)
        const uint64_t bdo = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_DO + st * FbSmem::TILE), 128 * 128, 1024);
        const uint64_t bq = umma_desc_sw128(smem_u32(smem + FbSmem::OFF_Q + st * FbSmem::TILE), 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tm_dv, ap_mn + 128 * k, bdo + 128 * k, idesc_kv, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tm_dk, ads_mn + 128 * k, bq + 128 * k, idesc_kv, (i > 0 || k > 0) ? 1u : 0u);
        // dQ_i[q, hd] = sum_keys dS[q, key] K[key, hd]: A = dS K-major (keys: 2 atoms x 4 k-steps), B = K MN-major
        mbar_wait(&dq_empty[i & 1], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t ads = umma_desc_sw128(sds, 0, 1024);
        const uint64_t bk = umma_desc_sw128(sk, 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t aoff = (uint64_t)((k >> 2) * (FbSmem::TILE >> 4) + (k & 3) * 2);
          umma_ss(tm_dq[i & 1], ads + aoff, bk + 128 * k, idesc_q, k > 0 ? 1u : 0u);
        }
        umma_commit(&dq_full[i & 1]);
        umma_commit(pds_empty);
        umma_commit(&qdo_empty[st]);
        FB_STAMP(3);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    const int cw = warp - 2;                 // 0..15
    const int q4 = warp & 3;                 // TMEM lane quarter = query-row quarter
    const int cg = cw >> 2;                  // 32-key column group
    const int lane = (int)lane_id();
    const int r = q4 * 32 + lane;            // query row inside the tile (S/dP/dQ) or key row (epilogue)
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t prow = smem_u32(smem + FbSmem::OFF_P + (cg >> 1) * FbSmem::TILE + r * 128);
    const uint32_t dsrow = smem_u32(smem + FbSmem::OFF_DS + (cg >> 1) * FbSmem::TILE + r * 128);
    const int g = cg;                        // dQ column half handled by the warps with cg < 2
    uint8_t* dq_box_p = smem + FbSmem::OFF_DQ + (cw & 7) * 4096;
    const uint32_t dq_box = smem_u32(dq_box_p);
    const long long stat_row = ((long long)b * p.H + h) * p.Npad;
    const uint64_t cc = f2_pack(p.scale_log2, p.scale_log2);
    auto reduce_dq = [&](int i) {
      // dQ_i rows of this warp, columns [32 g, 32 g + 32): TMEM -> swizzled smem box -> TMA reduce-add (fp32, at L2)
      if (cg >= 2) return;
      tc_fence_after();
      uint32_t t[32];
      tmem_ld_32x32b_x32(tm_dq[i & 1] + lane_sel + 32 * g, t);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&dq_empty[i & 1]);
        bulk_wait_read0();                   // previous reduce has finished reading the box
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(dq_box + lane * 128 + ((c ^ (lane & 7)) << 4), t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmDQ, dq_box_p, h * kHD + 32 * g, i * 128 + q4 * 32, b);
        bulk_commit();
      }
    };
    for (int i = 0; i < T; ++i) {
      const int qi = i * 128 + r;
      const float lse = __ldg(p.lse2 + stat_row + qi);        // padded rows: +inf -> P = 0
      const float del = __ldg(p.delta + stat_row + qi);       // padded rows: 0
      const uint64_t nl = f2_pack(-lse, -lse), nd = f2_pack(-del, -del);
      if (warp == 2 && lane == 0) FB_STAMP(4);
      mbar_wait(sdp_full, i & 1);
      if (warp == 2 && lane == 0) FB_STAMP(5);
      tc_fence_after();
      uint32_t pp[16], dd[16];                     // packed bf16x2: P and dS for this thread's 32 keys
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv_[16], dp_[16];
        tmem_ld_32x32b_x16(tm_s + lane_sel + 32 * cg + 16 * c, sv_);
        tmem_ld_32x32b_x16(tm_dp + lane_sel + 32 * cg + 16 * c, dp_);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t x = f2_fma(f2_pack(__uint_as_float(sv_[2 * k]), __uint_as_float(sv_[2 * k + 1])), cc, nl);
          const float p0 = fast_exp2(f2_lo(x)), p1 = fast_exp2(f2_hi(x));
          const uint64_t pv = f2_pack(p0, p1);
          const uint64_t ds = f2_mul(pv, f2_add(f2_pack(__uint_as_float(dp_[2 * k]), __uint_as_float(dp_[2 * k + 1])), nd));
          pp[8 * c + k] = pack_bf16(p0, p1);
          dd[8 * c + k] = pack_bf16(f2_lo(ds), f2_hi(ds));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty);
      if (warp == 2 && lane == 0) FB_STAMP(6);
      // dQ_{i-1} is the last MMA of tile i-1: once it has retired, P / dS of tile i-1 have been consumed as well
      if (i > 0) mbar_wait(&dq_full[(i - 1) & 1], ((i - 1) >> 1) & 1);
      if (warp == 2 && lane == 0) FB_STAMP(7);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int off = (((cg & 1) * 4 + c) ^ (r & 7)) << 4;
        sts128(prow + off, pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        sts128(dsrow + off, dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (warp == 2 && lane == 0) FB_STAMP(8);
      if (i > 0) reduce_dq(i - 1);
      if (warp == 2 && lane == 0) FB_STAMP(9);
    }
    mbar_wait(&dq_full[(T - 1) & 1], ((T - 1) >> 1) & 1);
    reduce_dq(T - 1);
    // epilogue: dK (scaled) and dV rows of this key tile (thread = key row; 32 of the 64 columns per warp)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int ki = k0 + r;
    {
      const bool is_dk = cg < 2;
      uint32_t t0[32];
      tmem_ld_32x32b_x32((is_dk ? tm_dk : tm_dv) + lane_sel + 32 * (cg & 1), t0);
      tmem_ld_wait();
      if (ki < p.N) {
        const float sc = is_dk ? p.scale : 1.0f;
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((long long)b * p.N + ki) * p.ld + (is_dk ? D : 2 * D) + h * kHD + 32 * (cg & 1));
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(sc * __uint_as_float(t0[8 * c]), sc * __uint_as_float(t0[8 * c + 1])),
                              pack_bf16(sc * __uint_as_float(t0[8 * c + 2]), sc * __uint_as_float(t0[8 * c + 3])),
                              pack_bf16(sc * __uint_as_float(t0[8 * c + 4]), sc * __uint_as_float(t0[8 * c + 5])),
                              pack_bf16(sc * __uint_as_float(t0[8 * c + 6]), sc * __uint_as_float(t0[8 * c + 7])));
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (rows q in [N, Npad) are written as 0)
__global__ void __launch_bounds__(256) flash_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                          float* __restrict__ delta, int B, int N, int H, int Npad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, q, h) with h fastest
  const long long total = (long long)B * Npad * H;
  if (idx >= total) return;
  const int h = (int)(idx % H);
  const long long bq = idx / H;
  const int q = (int)(bq % Npad), b = (int)(bq / Npad);
  float s = 0.f;
  if (q < N) {
    const long long off = ((long long)b * N + q) * (H * kHD) + h * kHD;
    const uint4* po = reinterpret_cast<const uint4*>(o + off);
    const uint4* pd = reinterpret_cast<const uint4*>(dout + off);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 a = __ldg(po + c), d = __ldg(pd + c);
      const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
      const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
      s += a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y;
    }
  }
  delta[((long long)b * H + h) * Npad + q] = s;
}

// dqkv[:, 0:D] = bf16(scale * dq_acc)
__global__ void __launch_bounds__(256) flash_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv,
                                                               long long rows, int D, long long ld, float scale) {
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
  static int generation = 2;               // DEVIAS_FLASH_FWD=1 selects the first-generation kernel (two CTAs per SM)
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FaSmem::BYTES));
    DV_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa2Smem::BYTES));
    const char* e = getenv("DEVIAS_FLASH_FWD");
    if (e != nullptr && e[0] == '1') generation = 1;
    attr_done = true;
  }
  FaParams p{batch, seq, heads, (seq + 127) / 128 * 128, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out),
             (long long)D, lse2};
  const int q_tiles = (seq + kQT - 1) / kQT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int prof = prof_begin(DEVIAS_PROF_ATTN, 4.0 * batch * heads * (double)seq * seq * kHD, s);
  if (generation == 2) flash_fwd2_kernel<<<batch * heads * q_tiles, kFa2Threads, Fa2Smem::BYTES, s>>>(tmQ, tmKV, p);
  else flash_fwd_kernel<<<batch * heads * q_tiles, kFaThreads, FaSmem::BYTES, s>>>(tmQ, tmKV, p);
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

extern "C" int devias_flash_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                                     float* delta_ws, float* dq_ws, int batch, int seq, int heads, int head_dim, float scale,
                                     int delta_ready, void* stream) {
  using namespace dv;
  DV_REQUIRE(qkv && out && dout && lse2 && dqkv && delta_ws && dq_ws, "null pointer");
  DV_REQUIRE(head_dim == 64, "head_dim 64 only (ViT-B/16)");
  DV_REQUIRE(batch > 0 && seq > 0 && heads > 0, "empty problem");
  const int D = heads * kHD;
  const int Npad = (seq + 127) / 128 * 128;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tmQKV, tmDO, tmDQ;
  int rc = make_qkv_tmap(&tmQKV, qkv, batch, seq, 3 * D, 128);
  if (rc) return rc;
  rc = make_qkv_tmap(&tmDO, dout, batch, seq, D, 128);
  if (rc) return rc;
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
  if (!delta_ready) {
    const long long total = (long long)batch * Npad * heads;
    flash_delta_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(out),
                                                                  static_cast<const __nv_bfloat16*>(dout), delta_ws, batch, seq,
                                                                  heads, Npad);
  }
  static long long* dbg_dev = nullptr;
  const bool dbg_on = getenv("DEVIAS_FLASH_DEBUG") != nullptr;
  if (dbg_on && dbg_dev == nullptr) {
    DV_CHECK_CUDA(cudaMalloc(&dbg_dev, 10 * 8 * sizeof(long long)));
    DV_CHECK_CUDA(cudaMemset(dbg_dev, 0, 10 * 8 * sizeof(long long)));
  }
  FbParams p{batch, seq, heads, Npad, scale, scale * 1.4426950408889634f, lse2, delta_ws,
             static_cast<__nv_bfloat16*>(dqkv), (long long)3 * D, dbg_on ? dbg_dev : nullptr};
  const int k_tiles = (seq + 127) / 128;
  const int prof = prof_begin(DEVIAS_PROF_ATTN, 10.0 * batch * heads * (double)seq * seq * kHD, s);
  flash_bwd_kernel<<<batch * heads * k_tiles, kBwdThreads, FbSmem::BYTES, s>>>(tmQKV, tmDO, tmDQ, p);
  prof_end(prof, s);
  {
    const long long n8 = (long long)batch * seq * (D / 8);
    long long blocks = (n8 + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    flash_dq_convert_kernel<<<(int)blocks, 256, 0, s>>>(dq_ws, static_cast<__nv_bfloat16*>(dqkv), (long long)batch * seq, D,
                                                        (long long)3 * D, scale);
  }
  DV_CHECK_CUDA(cudaGetLastError());
  if (dbg_on) {
    long long hst[80];
    DV_CHECK_CUDA(cudaDeviceSynchronize());
    DV_CHECK_CUDA(cudaMemcpy(hst, dbg_dev, sizeof(hst), cudaMemcpyDeviceToHost));
    static int printed = 0;
    if (printed++ < 2) {
      const char* names[10] = {"mma: loop top", "mma: S/dP(i+1) issued", "mma: pds_full seen", "mma: dV dK dQ issued", "cmp: wait sdp_full",
                               "cmp: sdp_full seen", "cmp: reads+math done", "cmp: dq_full(i-1) seen", "cmp: P/dS written", "cmp: dQ reduced"};
      const long long t0 = hst[4 * 8 + 0];
      for (int sl = 0; sl < 10; ++sl) {
        printf("%-24s", names[sl]);
        for (int i = 0; i < 8; ++i) printf(" %7lld", hst[sl * 8 + i] - t0);
        printf("\n");
      }
    }
  }
  count_launch(3);
  return DEVIAS_OK;
}
