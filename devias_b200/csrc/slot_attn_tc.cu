// Streaming slot attention for BF16 context tokens on the tensor cores (agg_block/attention.py:120-141 under PreNorm :32-40,
// folded form of devias_b200/slot_attention.py; BASELINE config 5 "fp32 and bf16", S = 2 / 4 / 8).
//
// The fp32 kernels of slot_attn.cu spend 11 (S = 2) to 59 (S = 8) fp32 FMAs per token element on the CUDA cores because the 1e-5
// contract of the fp32 path rules out reduced-precision products.  With bf16 tokens the inputs already carry 8 mantissa bits, so
// both contractions go to tcgen05 with the token tile as it lands from TMA -- no conversion pass, each byte read from HBM once:
//
//   phase 1   D1[token, sh] = sum_c t[token, c] g[sh, c]          M = 64 (32 tokens used), N = HS, K = 768; A = token tile (K-major)
//   softmax   logits = r (D1 - mu G) + c0, softmax over the S slots of each head, w = a r          thread <-> token (TMEM lane)
//   phase 2   U[sh, c] += sum_token w[sh, token] t[token, c]      M = 128 channels x 6, N = HS, K = 32; A = the SAME tile bytes read
//                                                                 MN-major, B = w (bf16, K-major, written by the softmax threads)
//
// The 128-byte-swizzled TMA boxes [64 channels x 32 tokens] are at the same time the K-major operand of phase 1 (rows = tokens)
// and the MN-major operand of phase 2 (rows = contraction index).  LayerNorm statistics of the tokens (the only per-element
// work left on the CUDA cores: ~3 instructions per element) are taken from the tile by the four compute warps while the MMAs
// of phase 1 run.  U accumulates in tensor memory over all tiles of a clip and is flushed with red.global.add per clip.
// Warp roles: 0, 1, 4, 5 = slot-axis softmax (two heads each), 2 = TMA producer, 3 / 16 = MMA issuers (phase 1 / phase 2),
// 4-7 = g image / drain, 8-15 = LayerNorm statistics.
#include "common.cuh"
#include "ptx.cuh"
#include "slot_tc.cuh"

namespace dv {

constexpr int kTcThreads = 544;                     // 16 warps, roles in the header comment
constexpr int kTcStatSlots = 4;

template <int HS>
struct SlotTcCfg {
  static constexpr int HSP = HS < 16 ? 16 : HS;                      // MMA N (phase 2 has M = 128: N % 16 == 0); extra rows are zero
  static constexpr int STAGES = HS <= 16 ? 4 : 3;
  static constexpr int OFF_TILE = 0;
  static constexpr int G_BOX = HSP * 128;                            // one channel box of the bf16 g image: [HSP rows][128 B]
  static constexpr int OFF_G = STAGES * kTTileBytes;
  static constexpr int W_BYTES = HSP * 128;                          // w[HSP rows][32 tokens] bf16 in 128-byte rows (first 64 B used)
  static constexpr int OFF_W = OFF_G + kTBoxes * G_BOX;
  static constexpr int OFF_STAT = OFF_W + 2 * W_BYTES;               // (mu, r)[kTcStatSlots][32 tokens]
  static constexpr int OFF_VEC = OFF_STAT + kTcStatSlots * kTT * 8;  // G[32], c0[32] of the clip
  static constexpr int OFF_BAR = OFF_VEC + 2 * 32 * 4;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
  // All 512 columns: the allocation then necessarily starts at column 0, which makes every tensor-memory address of the kernel a
  // compile-time constant (a base read back from shared memory costs an R2UR waterfall loop -- ~60 cycles -- per MMA issued).
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t TM_D1 = 0;                               // 2 x HSP columns: logits of two tiles in flight
  static constexpr uint32_t TM_D2 = 2 * HSP;                         // 6 x HSP columns: U[channel block][sh]
};

struct SlotTcParams {
  int B, N, tiles_per_clip;
  const float* g;              // [B, HS, 768]
  const float* G;              // [B, HS]
  const float* c0;             // [B, HS]
  float* U;                    // [B, HS, 768]  (+=)
  float* m;                    // [B, HS]       (+=)
  float* A;                    // [B, HS]       (+=)
  float* attn;                 // [B, HS, N] or null
  float* mu;                   // [B, N] or null
  float* rstd;                 // [B, N] or null
  float eps;
};

template <int HS>
__global__ void __launch_bounds__(kTcThreads, 1)
slot_stream_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotTcParams p) {
  pdl_trigger();
  using Cfg = SlotTcCfg<HS>;
  constexpr int S = HS / 4, HSP = Cfg::HSP, ST = Cfg::STAGES, NS = kTcStatSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                        // ST   TMA -> everyone
  uint64_t* tile_free = full + ST;              // ST   phase-2 MMAs of the tile retired (1) + the stats warps are done with it (8)
  uint64_t* d1_full = tile_free + ST;           // 2    phase-1 MMAs retired
  uint64_t* d1_free = d1_full + 2;              // 2    (2 softmax warps) logits read out of tensor memory
  uint64_t* w_full = d1_free + 2;               // 2    (2 softmax warps) weights of the tile are in shared memory
  uint64_t* w_free = w_full + 2;                // 2    phase-2 MMAs that read them retired
  uint64_t* st_full = w_free + 2;               // NS   (8 stats warps) LayerNorm statistics of a tile
  uint64_t* st_free = st_full + NS;             // NS   (2 softmax warps) consumed
  uint64_t* d2_full = st_free + NS;             // 1    last phase-2 MMA of a clip segment retired
  uint64_t* d2_free = d2_full + 1;              // 1    (4 slot warps) U of the segment flushed
  uint64_t* g_ready = d2_free + 1;              // 1    bf16 image of the segment's g is in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_ready + 1);
  float* stat = reinterpret_cast<float*>(smem + Cfg::OFF_STAT);
  float* Gs = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
  float* c0s = Gs + 32;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // tells the compiler the role dispatch is warp-uniform
  const int tpc = p.tiles_per_clip;
  const long long total = (long long)p.B * tpc;
  const int start = (int)(total * blockIdx.x / gridDim.x), end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (start >= end) return;
  const int n = end - start;
  const uint32_t tile_u = smem_u32(smem + Cfg::OFF_TILE), g_u = smem_u32(smem + Cfg::OFF_G), w_u = smem_u32(smem + Cfg::OFF_W);

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmTok);
      for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&tile_free[s], 9); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&d1_full[s], 1); mbar_init(&d1_free[s], 4);
        mbar_init(&w_full[s], 4); mbar_init(&w_free[s], 1);
      }
      for (int s = 0; s < NS; ++s) { mbar_init(&st_full[s], 8); mbar_init(&st_free[s], 4); }
      mbar_init(d2_full, 1); mbar_init(d2_free, 4); mbar_init(g_ready, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int i = tid; i < 2 * Cfg::W_BYTES / 16; i += kTcThreads) sts128(w_u + 16 * i, 0u, 0u, 0u, 0u);   // padding rows stay zero
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();               // see TMEM_COLS
  pdl_wait();

  if (warp == 2) {
    // =============================================================== TMA producer
    if (lane == 0) {
      for (int it = 0; it < n; ++it) {
        const int gt = start + it, st = it % ST;
        mbar_wait(&tile_free[st], ((it / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], kTTileBytes);
        tma_load_4d(smem + Cfg::OFF_TILE + st * kTTileBytes, &tmTok, &full[st], 0, (gt % tpc) * kTT, 0, gt / tpc);
      }
    }
    __syncwarp();
  } else if (warp == 3 || warp == 16) {
    // =============================================================== MMA issuers (whole warp converged, one elected lane issues)
    {
      const uint32_t lead = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc1 = umma_idesc_bf16(64, HSP, false, false);    // D1[tokens x HSP]: A = tile K-major, B = g K-major
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, HSP, true, false);    // D2[channels x HSP]: A = tile MN-major, B = w K-major
      int seg1 = 0, seg2 = 0;                                                // clip segments begun by phase 1 / finished by phase 2
      // stage / buffer indices are passed as values that are compile-time constants after unrolling (see the loop below): the
      // operand descriptors are then shared-memory base + constant, which the compiler keeps in uniform registers (descriptors
      // built from a run-time stage index live in vector registers and cost two R2UR + their latency per MMA)
      auto phase2 = [&](int it, int st, int buf) {
        const int gt = start + it;
        const bool first = it == 0 || gt % tpc == 0, last = it == n - 1 || (gt + 1) % tpc == 0;
        mbar_wait(&w_full[buf], (it >> 1) & 1);
        if (first) mbar_wait(d2_free, (seg2 & 1) ^ 1);                       // U of the previous segment has left tensor memory
        tc_fence_after();
        const uint64_t db = umma_desc_sw128(w_u + buf * Cfg::W_BYTES, 0, 1024);
        const uint64_t da0 = umma_desc_sw128(tile_u + st * kTTileBytes, kTBoxBytes, 1024);
        const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll
        for (int mb = 0; mb < 6; ++mb) {
          const uint64_t da = da0 + (uint64_t)(2 * mb * (kTBoxBytes >> 4));
          umma_ss_lead(lead, Cfg::TM_D2 + mb * HSP, da, db, idesc2, acc0);
          umma_ss_lead(lead, Cfg::TM_D2 + mb * HSP, da + 128, db + 2, idesc2, 1u);
        }
        umma_commit_lead(lead, &tile_free[st]);
        umma_commit_lead(lead, &w_free[buf]);
        if (last) { umma_commit_lead(lead, d2_full); ++seg2; }
      };
      auto phase1 = [&](int it, int st, int buf) {
        mbar_wait(&full[st], (it / ST) & 1);
        mbar_wait(&d1_free[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t da0 = umma_desc_sw128(tile_u + st * kTTileBytes, 0, 1024);
        const uint64_t db0 = umma_desc_sw128(g_u, 0, 1024);
        const uint32_t d1 = Cfg::TM_D1 + buf * HSP;
#pragma unroll
        for (int box = 0; box < kTBoxes; ++box) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss_lead(lead, d1, da0 + (uint64_t)(box * (kTBoxBytes >> 4) + 2 * k),
                         db0 + (uint64_t)(box * (Cfg::G_BOX >> 4) + 2 * k), idesc1, (box > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit_lead(lead, &d1_full[buf]);
      };
      constexpr int UF = (ST % 2 == 0) ? ST : 2 * ST;                        // unroll: stage and buffer index periodic in it
      // Two issuing warps: warp 3 queues the logits MMAs of a tile as soon as it has landed, warp 16 the weighted-sum MMAs as soon
      // as the softmax weights exist.  (One thread issuing both had to wait for softmax(it - 1) before it could queue phase 1 of
      // tile it + 1: the token ring ran dry behind that serialisation.)
      if (warp == 3) {
        for (int base = 0; base < n; base += UF) {
#pragma unroll
          for (int j = 0; j < UF; ++j) {
            const int it = base + j;
            if (it < n) {
              if (it == 0 || (start + it) % tpc == 0) { mbar_wait(g_ready, seg1 & 1); ++seg1; }
              phase1(it, j % ST, j & 1);
            }
          }
        }
      } else {
        for (int base = 0; base < n; base += UF) {
#pragma unroll
          for (int j = 0; j < UF; ++j) {
            const int it = base + j;
            if (it < n) phase2(it, j % ST, j & 1);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8 && warp < 16) {
    // =============================================================== stats warps: LayerNorm moments of the tokens, one tile ahead
    const int tl = (warp - 8) * 4 + (lane >> 3), pq = lane & 7;              // 8 threads per token, one 16-byte unit per box each
    for (int it = 0; it < n; ++it) {
      const int gt = start + it, st = it % ST, slot = it % NS;
      const int b = gt / tpc, tok = (gt % tpc) * kTT + tl;
      mbar_wait(&full[st], (it / ST) & 1);
      const uint32_t row_u = tile_u + st * kTTileBytes + tl * 128;
      const float x0 = __uint_as_float(lds32u(row_u + ((tl & 7) << 4)) << 16);   // shift of the one-pass moments: channel 0
      const uint64_t nx0 = f2_pack(-x0, -x0);
      uint64_t s1 = 0ull, s2 = 0ull;
      const uint32_t col_u = row_u + ((pq ^ (tl & 7)) << 4);                  // a quarter warp reads one token row: conflict-free
#pragma unroll
      for (int box = 0; box < kTBoxes; ++box) {
        const uint4 v = lds128u(col_u + box * kTBoxBytes);
        const uint64_t d0 = f2_add(bf16x2_to_f2(v.x), nx0), d1 = f2_add(bf16x2_to_f2(v.y), nx0);
        const uint64_t d2 = f2_add(bf16x2_to_f2(v.z), nx0), d3 = f2_add(bf16x2_to_f2(v.w), nx0);
        s1 = f2_add(s1, f2_add(f2_add(d0, d1), f2_add(d2, d3)));
        s2 = f2_fma(d0, d0, f2_fma(d1, d1, f2_fma(d2, d2, f2_fma(d3, d3, s2))));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_free[st]);                             // this warp has read what it needs of the tile
      float a1 = f2_lo(s1) + f2_hi(s1), a2 = f2_lo(s2) + f2_hi(s2);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      mbar_wait(&st_free[slot], ((it / NS) & 1) ^ 1);
      if (pq == 0) {
        const float d1 = a1 * (1.0f / kTD);
        const float mu = x0 + d1;
        const float r = rsqrtf(fmaxf(a2 * (1.0f / kTD) - d1 * d1, 0.f) + p.eps);
        stat[(slot * kTT + tl) * 2] = mu;
        stat[(slot * kTT + tl) * 2 + 1] = r;
        if (p.mu != nullptr && tok < p.N) { p.mu[(long long)b * p.N + tok] = mu; p.rstd[(long long)b * p.N + tok] = r; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&st_full[slot]);
    }
  } else {
    // =============================================================== softmax warps (0, 1: heads 0-1; 4, 5: heads 2-3) and slot warps
    // (4-7: bf16 image of g, U out of tensor memory).  D1 (M = 64) keeps token 16 q + i in lane i < 16 of lane quarter q.
    const int q = warp & 3;                                                  // the TMEM lane quarter this warp may read
    const bool slot_warp = warp >= 4, sm_warp = q < 2;
    const int tc = tid - 128;                                                // thread index among the slot warps
    const int hg = warp >> 2;                                                // head group of a softmax warp
    const int tk = 16 * q + (lane & 15);
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    constexpr int HH = HS / 2;                                               // (head, slot) rows per softmax warp
    const int hl = lane >> 4;                                                // lanes 0-15: first head of the group, 16-31: second
    float accA[S], accM[S];
    int seg = 0;
    for (int it = 0; it < n;) {
      const int b = (start + it) / tpc;
      const int seg_n = min(n - it, tpc - (start + it) % tpc);               // tiles of this clip in our range
      if (slot_warp) {
        // ---- bf16 image of g[b]: [12 boxes][HSP rows][128 B], 16-byte chunks swizzled by (row & 7); rows >= HS are zero
        for (int i = tc; i < HSP * 96; i += 128) {
          const int row = i / 96, c8 = i - row * 96;
          uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
          if (row < HS) {
            const float4* src = reinterpret_cast<const float4*>(p.g + ((long long)b * HS + row) * kTD + 8 * c8);
            const float4 x = __ldg(src), y = __ldg(src + 1);
            w0 = pack_bf16(x.x, x.y); w1 = pack_bf16(x.z, x.w); w2 = pack_bf16(y.x, y.y); w3 = pack_bf16(y.z, y.w);
          }
          sts128(g_u + (c8 >> 3) * Cfg::G_BOX + row * 128 + (((c8 & 7) ^ (row & 7)) << 4), w0, w1, w2, w3);
        }
        if (tc < HS) { Gs[tc] = __ldg(p.G + b * HS + tc); c0s[tc] = __ldg(p.c0 + b * HS + tc); }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tc == 0) mbar_arrive(g_ready);
      } else {
        mbar_wait(g_ready, seg & 1);                                         // G / c0 of this clip are in shared memory
      }
      if (sm_warp) {
        const int row0 = hg * HH + hl * S;                                   // this thread's (head, slot) rows: row0 .. row0 + S - 1
        float Gr[S], cr[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { accA[s] = 0.f; accM[s] = 0.f; Gr[s] = Gs[row0 + s]; cr[s] = c0s[row0 + s] * 1.4426950408889634f; }
        for (int e = it + seg_n, i2 = it; i2 < e; ++i2) {
          const int buf = i2 & 1, slot = i2 % NS;
          const int tok = ((start + i2) % tpc) * kTT + tk;
          // ---- softmax over the slots of one head, thread <-> (token, head): TMEM lane i < 16 holds token 16 q + i; lanes 16-31
          //      take the group's second head of the same tokens through a shuffle
          mbar_wait(&d1_full[buf], (i2 >> 1) & 1);
          tc_fence_after();
          uint32_t d[HSP];
          tmem_ld_cols<HSP>(Cfg::TM_D1 + lane_sel + buf * HSP, d);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&d1_free[buf]);
          float a[S];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            // (d is indexed with compile-time constants on both head groups so that it stays in registers)
            const uint32_t lo = hg == 0 ? d[s] : d[HH + s], hi = hg == 0 ? d[S + s] : d[HH + S + s];
            const uint32_t other = __shfl_sync(0xffffffffu, hi, lane & 15);
            a[s] = __uint_as_float(hl == 0 ? lo : other);
          }
          mbar_wait(&st_full[slot], (i2 / NS) & 1);
          const float mu = stat[(slot * kTT + tk) * 2], r = stat[(slot * kTT + tk) * 2 + 1];
          const bool valid = tok < p.N;
          float mx = -INFINITY;
          const float r2 = r * 1.4426950408889634f;                          // logits in the log2 domain: one MUFU.EX2 per weight
#pragma unroll
          for (int s = 0; s < S; ++s) {
            a[s] = fmaf(r2, a[s] - mu * Gr[s], cr[s]);
            mx = fmaxf(mx, a[s]);
          }
          float sum = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) { a[s] = fast_exp2(a[s] - mx); sum += a[s]; }
          const float inv = valid ? __frcp_rn(sum) : 0.f;
          mbar_wait(&w_free[buf], ((i2 >> 1) & 1) ^ 1);
          const uint32_t wt = w_u + buf * Cfg::W_BYTES + (tk & 7) * 2;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int i = row0 + s;
            const float av = a[s] * inv, w = av * r;
            const __nv_bfloat16 wb = __float2bfloat16_rn(w);
            sts16(wt + i * 128 + (((tk >> 3) ^ (i & 7)) << 4), *reinterpret_cast<const uint16_t*>(&wb));
            accA[s] += av;
            accM[s] = fmaf(w, mu, accM[s]);
            if (valid && p.attn != nullptr) p.attn[((long long)b * HS + i) * p.N + tok] = av;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&w_full[buf]); mbar_arrive(&st_free[slot]); }
        }
        // ---- A / m of the clip segment: sums over the 16 tokens of each half warp
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float a = accA[s], mm = accM[s];
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            mm += __shfl_xor_sync(0xffffffffu, mm, o);
          }
          if ((lane & 15) == 0) { atomicAdd(p.A + b * HS + row0 + s, a); atomicAdd(p.m + b * HS + row0 + s, mm); }
        }
      }
      it += seg_n;
      if (slot_warp) {
        // ---- U of the segment out of tensor memory
        mbar_wait(d2_full, seg & 1);
        tc_fence_after();
        float* dst = p.U + (long long)b * HS * kTD + 32 * q + lane;
#pragma unroll 1
        for (int mb = 0; mb < 6; ++mb) {
          uint32_t u[HSP];
          tmem_ld_cols<HSP>(Cfg::TM_D2 + lane_sel + mb * HSP, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HS; ++i) red_add_f32(dst + i * kTD + 128 * mb, __uint_as_float(u[i]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_free);
      }
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(0u);
}

template <int HS>
static int launch_slot_tc(const CUtensorMap& tm, const SlotTcParams& p, cudaStream_t s) {
  using Cfg = SlotTcCfg<HS>;
  static_assert(Cfg::BYTES <= 227 * 1024, "tensor-core slot forward does not fit in shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_tc_fwd_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    attr_done = true;
  }
  const long long total = (long long)p.B * p.tiles_per_clip;      // persistent: an even share of all tiles per SM
  long long grid = sm_count();
  if (grid > total) grid = total;
  const double bytes = (double)p.B * p.N * kTD * 2 + (p.attn ? (double)p.B * HS * p.N * 4 : 0.0);
  const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
  DV_CHECK_CUDA(launch_k(slot_stream_tc_fwd_kernel<HS>, dim3((unsigned)grid), dim3((unsigned)kTcThreads), (size_t)Cfg::BYTES, s, tm, p));
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

}  // namespace dv

extern "C" int devias_slot_stream_fwd_bf16(const void* tokens, const float* g, const float* G, const float* c0, float* U, float* m,
                                           float* A, float* attn, float* mu, float* rstd, int batch, int n_tokens, int dim,
                                           int num_slots, float eps, void* stream) {
  using namespace dv;
  DV_REQUIRE(tokens && g && G && c0 && U && m && A, "null pointer");
  DV_REQUIRE(dim == kTD, "token dim must be 768");
  DV_REQUIRE(num_slots == 2 || num_slots == 4 || num_slots == 8, "num_slots must be 2, 4 or 8 (4 heads x S query vectors)");
  DV_REQUIRE((mu == nullptr) == (rstd == nullptr), "mu and rstd go together");
  DV_REQUIRE(batch > 0 && n_tokens > 0, "empty problem");
  DV_REQUIRE(reinterpret_cast<uintptr_t>(tokens) % 16 == 0, "tokens must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm;
  int rc = make_token_tmap_bf16(&tm, tokens, batch, n_tokens);
  if (rc) return rc;
  SlotTcParams p{batch, n_tokens, (n_tokens + kTT - 1) / kTT, g, G, c0, U, m, A, attn, mu, rstd, eps};
  switch (num_slots) {
    case 2: return launch_slot_tc<8>(tm, p, s);
    case 4: return launch_slot_tc<16>(tm, p, s);
    default: return launch_slot_tc<32>(tm, p, s);
  }
}
