// Streaming slot attention (agg_block/attention.py:120-141 under PreNorm :32-40) in the folded form of
// devias_b200/slot_attention.py: per layer the 1568 x 768 context tokens of a clip are read from HBM exactly once.
//
//   sim[sh, j] = r_j * (g[sh] . t_j - mu_j * G[sh]) + c0[sh]        (sh = head * S + slot;  mu_j, r_j = LayerNorm stats of t_j)
//   a = softmax over the S slots of each head;   A[sh] = sum_j a,   U[sh] = sum_j (a r_j) t_j,   m[sh] = sum_j a r_j mu_j
//
// One CTA streams a contiguous range of 16-token tiles of one clip through a TMA ring (128B-swizzled boxes of
// 32 floats x 16 tokens, so both access patterns below are bank-conflict free):
//   phase 1 (8 warps)  : lane <-> token, warp <-> 96-wide channel slice: HS dot products + shifted LayerNorm moments
//   phase 1b (warp 0)  : combine the slices, LayerNorm stats, logits, slot-axis softmax, weights
//   phase 2 (6 warps)  : thread <-> 4 channels: U[sh][d] += w[sh][j] * t[j][d]
// Partial U / m / A of the CTA are flushed with red.global.add (few CTAs per clip).  fp32 FFMA throughout: the 1e-5
// parity budget of BASELINE.json rules out tf32/bf16 tensor-core products for the fp32 path.
#include "common.cuh"
#include "ptx.cuh"

namespace dv {

constexpr int kSD = 768;          // channels
constexpr int kST = 16;           // tokens per tile
constexpr int kSBoxes = kSD / 32; // 24 TMA boxes (32 floats) per tile
constexpr int kSTileBytes = kST * kSD * 4;   // 48 KiB
constexpr int kSlotThreads = 256;

template <int HS>
struct SlotCfg {
  static constexpr int STAGES = (HS <= 16) ? 3 : 2;
  static constexpr int OFF_TILE = 0;
  static constexpr int OFF_G = STAGES * kSTileBytes;                 // g[HS][768] fp32
  static constexpr int OFF_PART = OFF_G + HS * kSD * 4;              // partial[8 warps][16 tokens][HS + 2]
  static constexpr int PART_BYTES = 8 * kST * (HS + 2) * 4;
  static constexpr int OFF_W = OFF_PART + PART_BYTES;                // w[16 tokens][HS]  (a * r)
  static constexpr int OFF_ACC = OFF_W + kST * HS * 4;               // running A / m partials [2][16 tokens][HS]
  static constexpr int OFF_BAR = OFF_ACC + 2 * kST * HS * 4;
  static constexpr int BYTES = OFF_BAR + 64 + 1024;
};

struct SlotParams {
  int B, N, S;                 // HS = 4 * S
  int tiles_per_cta, tiles_per_clip;
  const float* g;              // [B, HS, 768]
  const float* G;              // [B, HS]
  const float* c0;             // [B, HS]
  float* U;                    // [B, HS, 768]  (+=)
  float* m;                    // [B, HS]       (+=)
  float* A;                    // [B, HS]       (+=)
  float* attn;                 // [B, HS, N] or null
  float* mu;                   // [B, N] or null (LayerNorm stats written for the backward)
  float* rstd;                 // [B, N] or null
  float eps;
};

// shared-space address of the 16-byte chunk holding channels [4*c4, 4*c4+4) of token `tok` inside a swizzled tile
__device__ __forceinline__ void st_global_v2(float* p, uint64_t v) {
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(f2_lo(v)), "f"(f2_hi(v)) : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint32_t tile_chunk(uint32_t tile, int tok, int c4) {
  const int box = c4 >> 3, chunk = c4 & 7;
  return tile + box * (kST * 128) + tok * 128 + ((chunk ^ (tok & 7)) << 4);
}

template <int HS>
__global__ void __launch_bounds__(kSlotThreads, 1)
slot_stream_fwd_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotParams p) {
  pdl_trigger();
  pdl_wait();
  using Cfg = SlotCfg<HS>;
  constexpr int S = HS / 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* g_s = reinterpret_cast<float*>(smem + Cfg::OFF_G);
  float* part = reinterpret_cast<float*>(smem + Cfg::OFF_PART);
  float* w_s = reinterpret_cast<float*>(smem + Cfg::OFF_W);
  float* accA = reinterpret_cast<float*>(smem + Cfg::OFF_ACC);   // [16][HS]
  float* accM = accA + kST * HS;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.tiles_per_clip - tile0);
  if (ntiles <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  // g of this clip -> smem (plain coalesced loads; 24..96 KiB once per CTA)
  {
    const float4* src = reinterpret_cast<const float4*>(p.g + (long long)b * HS * kSD);
    float4* dst = reinterpret_cast<float4*>(g_s);
    for (int i = tid; i < HS * kSD / 4; i += kSlotThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < 2 * kST * HS; i += kSlotThreads) accA[i] = 0.f;
  }
  __syncthreads();

  auto issue = [&](int it) {   // one elected thread: 24 boxes of 32 floats x 16 tokens
    const int st = it % Cfg::STAGES;
    mbar_arrive_expect_tx(&full[st], kSTileBytes);
    uint8_t* dst = smem + Cfg::OFF_TILE + st * kSTileBytes;
    const int tok0 = (tile0 + it) * kST;
    tma_load_4d(dst, &tmTok, &full[st], 0, tok0, 0, b);   // one bulk tensor copy: [24 channel boxes][16 tokens][32 floats]
  };
  if (tid == 0) {
    for (int it = 0; it < Cfg::STAGES - 1 && it < ntiles; ++it) issue(it);
  }

  // phase-2 accumulators: thread t < 192 owns channels [4t, 4t+4) for every sh
  float4 acc[HS];
#pragma unroll
  for (int i = 0; i < HS; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t g_u = smem_u32(g_s), w_u = smem_u32(w_s);
  const float* Gs = p.G + b * HS;
  const float* c0s = p.c0 + b * HS;

  const int tok_l = lane & 15, half = lane >> 4;
  const int slice = warp * 2 + half;                  // 16 channel slices of 48
  for (int it = 0; it < ntiles; ++it) {
    const int st = it % Cfg::STAGES;
    // keep the ring full: the stage being refilled was consumed in iteration it-1 (all threads passed its last barrier)
    if (tid == 0 && it + Cfg::STAGES - 1 < ntiles) issue(it + Cfg::STAGES - 1);
    mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
    const uint32_t tile = smem_u32(smem + Cfg::OFF_TILE + st * kSTileBytes);
    const int tok_base = (tile0 + it) * kST;

    // ---------------- phase 1: partial dots over this lane's 48 channels of token tok_l
    {
      float dot[HS];
#pragma unroll
      for (int i = 0; i < HS; ++i) dot[i] = 0.f;
      const float x0 = lds32(tile_chunk(tile, tok_l, 0));   // shift for the one-pass moments (same for all slices of a token)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int c = 0; c < 12; ++c) {
        const int c4 = slice * 12 + c;
        const float4 t = lds128(tile_chunk(tile, tok_l, c4));
        const float a0 = t.x - x0, a1 = t.y - x0, a2 = t.z - x0, a3 = t.w - x0;
        s1 += (a0 + a1) + (a2 + a3);
        s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
#pragma unroll
        for (int i = 0; i < HS; ++i) {
          const float4 gv = lds128(g_u + (i * kSD + c4 * 4) * 4);
          dot[i] = fmaf(t.x, gv.x, fmaf(t.y, gv.y, fmaf(t.z, gv.z, fmaf(t.w, gv.w, dot[i]))));
        }
      }
      // combine the two 48-channel halves of this warp, lanes 0..15 publish
#pragma unroll
      for (int i = 0; i < HS; ++i) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], 16);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      if (half == 0) {
        float* pp = part + (warp * kST + tok_l) * (HS + 2);
#pragma unroll
        for (int i = 0; i < HS; ++i) pp[i] = dot[i];
        pp[HS] = s1;
        pp[HS + 1] = s2;
      }
    }
    __syncthreads();
    // ---------------- phase 1b: warp 0, lanes 0..15 <-> tokens
    if (warp == 0 && lane < kST) {
      float dot[HS];
#pragma unroll
      for (int i = 0; i < HS; ++i) dot[i] = 0.f;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int sl = 0; sl < 8; ++sl) {
        const float* pp = part + (sl * kST + lane) * (HS + 2);
#pragma unroll
        for (int i = 0; i < HS; ++i) dot[i] += pp[i];
        s1 += pp[HS];
        s2 += pp[HS + 1];
      }
      const float x0 = lds32(tile_chunk(tile, lane, 0));
      const float d1 = s1 * (1.0f / kSD);
      const float mu = x0 + d1;
      const float var = fmaxf(s2 * (1.0f / kSD) - d1 * d1, 0.f);
      const float r = rsqrtf(var + p.eps);
      const int tok = tok_base + lane;
      const bool valid = tok < p.N;
      if (valid && p.mu != nullptr) { p.mu[(long long)b * p.N + tok] = mu; p.rstd[(long long)b * p.N + tok] = r; }
      float a[HS];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int i = h * S + s;
          a[i] = fmaf(r, dot[i] - mu * __ldg(Gs + i), __ldg(c0s + i));
          mx = fmaxf(mx, a[i]);
        }
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) { a[h * S + s] = expf(a[h * S + s] - mx); sum += a[h * S + s]; }
        const float inv = valid ? 1.0f / sum : 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) a[h * S + s] *= inv;
      }
#pragma unroll
      for (int i = 0; i < HS; ++i) {
        const float w = a[i] * r;
        w_s[lane * HS + i] = w;
        accA[lane * HS + i] += a[i];
        accM[lane * HS + i] = fmaf(w, mu, accM[lane * HS + i]);
        if (p.attn != nullptr && valid) p.attn[((long long)b * HS + i) * p.N + tok] = a[i];
      }
    }
    __syncthreads();
    // ---------------- phase 2: U[sh][4t..4t+3] += w[j][sh] * token_j[4t..4t+3]
    if (tid < kSD / 4) {
#pragma unroll 4
      for (int j = 0; j < kST; ++j) {
        const float4 t = lds128(tile_chunk(tile, j, tid));
#pragma unroll
        for (int i4 = 0; i4 < HS / 4; ++i4) {
          const float4 w = lds128(w_u + (j * HS + 4 * i4) * 4);
          float4& a0 = acc[4 * i4]; float4& a1 = acc[4 * i4 + 1]; float4& a2 = acc[4 * i4 + 2]; float4& a3 = acc[4 * i4 + 3];
          a0.x = fmaf(w.x, t.x, a0.x); a0.y = fmaf(w.x, t.y, a0.y); a0.z = fmaf(w.x, t.z, a0.z); a0.w = fmaf(w.x, t.w, a0.w);
          a1.x = fmaf(w.y, t.x, a1.x); a1.y = fmaf(w.y, t.y, a1.y); a1.z = fmaf(w.y, t.z, a1.z); a1.w = fmaf(w.y, t.w, a1.w);
          a2.x = fmaf(w.z, t.x, a2.x); a2.y = fmaf(w.z, t.y, a2.y); a2.z = fmaf(w.z, t.z, a2.z); a2.w = fmaf(w.z, t.w, a2.w);
          a3.x = fmaf(w.w, t.x, a3.x); a3.y = fmaf(w.w, t.y, a3.y); a3.z = fmaf(w.w, t.z, a3.z); a3.w = fmaf(w.w, t.w, a3.w);
        }
      }
    }
    __syncthreads();   // tile + w_s + part fully consumed: the stage may be refilled
  }
  // ---------------- flush partials
  if (tid < kSD / 4) {
    float* dst = p.U + (long long)b * HS * kSD + tid * 4;
#pragma unroll
    for (int i = 0; i < HS; ++i) red_add_v4_f32(dst + i * kSD, acc[i].x, acc[i].y, acc[i].z, acc[i].w);
  }
  if (tid < HS) {   // the loop's final barrier made warp 0's running sums visible
    float a = 0.f, mm = 0.f;
#pragma unroll
    for (int j = 0; j < kST; ++j) { a += accA[j * HS + tid]; mm += accM[j * HS + tid]; }
    atomicAdd(p.A + b * HS + tid, a);
    atomicAdd(p.m + b * HS + tid, mm);
  }
}

// =====================================================================================================================
// Forward, version 2 (HS = 8, 16).
//  * warp-specialised and software-pipelined: group A (4 warps) takes the dot products / moments of tile t+1 and turns them
//    into softmax weights (w ring, 2 deep) while group B (8 warps) accumulates tile t; one thread of A is the TMA producer.
//  * S-1 instead of S vectors per head: the slot-axis softmax only depends on logit DIFFERENCES, so phase 1 takes the dots
//    with gd[h,s] = g[h,s] - g[h,0] (s >= 1); sum_s a[h,s,j] = 1, so U[h,0] = R - sum_{s>=1} U[h,s] with R = sum_j r_j t_j
//    and phase 2 accumulates the S-1 slots of each head plus the row R  (S = 2: 4 + 5 instead of 8 + 8 vector passes).
//  * shared-memory return path economy (what bounds the kernel on the SM side, see the backward below): in phase 1 a lane
//    owns 4 (S = 2) or 2 (S = 4) tokens so that one broadcast LDS.128 of a gd chunk feeds that many tokens; the partial sums
//    are combined across the lanes sharing a token with a halving butterfly; weights are stored once, not as FFMA2 pairs.
//  * persistent CTAs: the B x 98 tiles are split evenly over the SMs; a CTA's range may span clips (group A reloads gd and
//    both groups flush their per-clip sums at the boundary while the TMA ring keeps running).
template <int HS>
struct SlotCfg2 {
  static constexpr int S = HS / 4;
  static constexpr int HE = HS - 4;                                  // effective vectors: 4 heads x (S - 1) slots
  static constexpr int TPL = (HS == 8) ? 4 : 2;                      // tokens per lane in phase 1
  static constexpr int NG = kST / TPL;                               // token groups per warp ...
  static constexpr int NE = 32 / NG;                                 // ... each split over NE lanes along the channels
  static constexpr int NV1 = TPL * (HE + 2);                         // per-lane partial sums: dots[HE], s1, s2 per token
  static constexpr int NFIN = NV1 / NE;                              // left per lane after the butterfly
  static constexpr int WS = (HE + 1 + 3) / 4 * 4;                    // floats per token in the w ring: w[HE], R, pad
  static constexpr int STAGES = (HS <= 8) ? 4 : 3;
  static constexpr int NSUB = (HS <= 8) ? 2 : 1;                     // group-A sub-groups (4 warps each) taking alternate tiles
  static constexpr int NBW = (HS <= 8) ? 4 : 8;                      // group-B warps
  static constexpr int TOKB = kST / (NBW / 4);                       // tokens of a tile per group-B warp
  static constexpr bool OWN_PRODUCER = NSUB > 1;                     // a TMA producer warp of its own (else: thread 0 of group A)
  static constexpr int PRODUCER = 4 * NSUB + NBW;
  static constexpr int THREADS = (PRODUCER + (OWN_PRODUCER ? 1 : 0)) * 32;
  static constexpr int OFF_TILE = 0;
  static constexpr int OFF_G = STAGES * kSTileBytes;                 // gd[NSUB][HE][768] fp32 (one copy per sub-group)
  static constexpr int PART_STRIDE = HE + 4;                         // dots[HE], s1, s2, x0, pad
  static constexpr int OFF_PART = OFF_G + NSUB * HE * kSD * 4;       // partial[NSUB][4 warps][16 tokens][PART_STRIDE]
  static constexpr int OFF_W = OFF_PART + NSUB * 4 * kST * PART_STRIDE * 4; // w ring [2][16 tokens][WS]
  static constexpr int OFF_BAR = OFF_W + 2 * kST * WS * 4;
  static constexpr int BYTES = OFF_BAR + 128 + 1024;
  static_assert(NV1 % NE == 0, "butterfly needs an even split");
};

template <int HS>
__global__ void __launch_bounds__(SlotCfg2<HS>::THREADS, 1)
slot_stream_fwd2_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotParams p) {
  pdl_trigger();
  pdl_wait();
  using Cfg = SlotCfg2<HS>;
  constexpr int S = HS / 4;
  constexpr int HE = Cfg::HE, WS = Cfg::WS, TPL = Cfg::TPL, NG = Cfg::NG, NE = Cfg::NE;
  constexpr int TPW = 32 / S;                 // tokens per warp-unit in phase 1b
  constexpr int UNITS = 4 * (kST / TPW);      // (head, token group) units per tile
  constexpr int UPW = UNITS / 4;              // units per group-A warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                       // STAGES
  uint64_t* tile_empty = bars + Cfg::STAGES;   // STAGES (4 + NBW warp arrivals)
  uint64_t* w_full = tile_empty + Cfg::STAGES; // 2 (4 arrivals)
  uint64_t* w_empty = w_full + 2;              // 2 (NBW arrivals)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tpc = p.tiles_per_clip;
  const long long total = (long long)p.B * tpc;
  const int start = (int)(total * blockIdx.x / gridDim.x), end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (start >= end) return;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&tile_empty[s], 4 + Cfg::NBW); }
    for (int s = 0; s < 2; ++s) { mbar_init(&w_full[s], 4); mbar_init(&w_empty[s], Cfg::NBW); }
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t w_u = smem_u32(smem + Cfg::OFF_W);
  auto issue = [&](int it) {                   // ring index it <-> global tile start + it
    const int gt = start + it, st = it % Cfg::STAGES;
    mbar_wait(&tile_empty[st], ((it / Cfg::STAGES) & 1) ^ 1);
    mbar_arrive_expect_tx(&full[st], kSTileBytes);
    tma_load_4d(smem + Cfg::OFF_TILE + st * kSTileBytes, &tmTok, &full[st], 0, (gt % tpc) * kST, 0, gt / tpc);
  };

  constexpr int PRE = Cfg::STAGES - 2;         // tiles in flight ahead of group A when one of its threads is the producer
  if (Cfg::OWN_PRODUCER && warp == Cfg::PRODUCER) {
    // =============================================================== TMA producer: tiles in order, as far ahead as the ring allows
    if (lane == 0) {
      for (int it = 0; start + it < end; ++it) issue(it);
    }
  } else if (warp < 4 * Cfg::NSUB) {
    // =============================================================== group A: sub-group `sub` takes the tiles it = sub (mod NSUB)
    const int sub = warp >> 2, wa = warp & 3, ta = tid & 127;
    const uint32_t g_u = smem_u32(smem + Cfg::OFF_G) + sub * (HE * kSD * 4);
    const uint32_t part_u = smem_u32(smem + Cfg::OFF_PART) + sub * (4 * kST * Cfg::PART_STRIDE * 4);
    if (!Cfg::OWN_PRODUCER && tid == 0) {
      for (int it = 0; it < PRE && start + it < end; ++it) issue(it);
    }
    const int tg = lane % NG, e = lane / NG;   // phase 1: token group (tokens tg + NG i) and channel split
    int pbase = 0;                             // first of the NFIN sums this lane owns after the butterfly
    {
      int n = Cfg::NV1;
#pragma unroll
      for (int off = 16; off >= NG; off >>= 1) { n >>= 1; if (lane & off) pbase += n; }
    }
    float accA[UPW], accM[UPW], Gd_[UPW], cd_[UPW];
    int cur_b = -1;
    auto flush = [&]() {
#pragma unroll
      for (int k = 0; k < UPW; ++k) {
        float a = accA[k], mm = accM[k];
#pragma unroll
        for (int o = 1; o < TPW; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          mm += __shfl_xor_sync(0xffffffffu, mm, o);
        }
        if (lane % TPW == 0) {
          const int unit = wa + 4 * k;
          const int sh = (unit % 4) * S + lane / TPW;
          atomicAdd(p.A + cur_b * HS + sh, a);
          atomicAdd(p.m + cur_b * HS + sh, mm);
        }
      }
    };
    for (int gt = start + sub; gt < end; gt += Cfg::NSUB) {
      const int it = gt - start, st = it % Cfg::STAGES;
      const int b = gt / tpc;
      if (b != cur_b) {                        // clip boundary: per-clip sums out, gd of the new clip in
        named_bar_sync(1 + sub, 128);          // every warp of the sub-group is done with the old gd
        if (cur_b >= 0) flush();
        cur_b = b;
        const float4* src = reinterpret_cast<const float4*>(p.g + (long long)b * HS * kSD);
        for (int i = ta; i < HE * kSD / 4; i += 128) {
          const int ev = i / (kSD / 4), c = i % (kSD / 4);
          const int h = ev / (S - 1), s = ev % (S - 1) + 1;
          const float4 a = __ldg(src + (h * S + s) * (kSD / 4) + c), r0 = __ldg(src + (h * S) * (kSD / 4) + c);
          sts128f(g_u + i * 16, make_float4(a.x - r0.x, a.y - r0.y, a.z - r0.z, a.w - r0.w));
        }
#pragma unroll
        for (int k = 0; k < UPW; ++k) {
          accA[k] = 0.f; accM[k] = 0.f;
          const int h = (wa + 4 * k) % 4, sh = h * S + lane / TPW;     // G / c0 of this lane's (head, slot) relative to slot 0
          Gd_[k] = __ldg(p.G + b * HS + sh) - __ldg(p.G + b * HS + h * S);
          cd_[k] = __ldg(p.c0 + b * HS + sh) - __ldg(p.c0 + b * HS + h * S);
        }
        named_bar_sync(1 + sub, 128);
      }
      if (!Cfg::OWN_PRODUCER && tid == 0 && gt + PRE < end) issue(it + PRE);
      mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
      const uint32_t tile = smem_u32(smem + Cfg::OFF_TILE + st * kSTileBytes);
      const int tok_base = (gt % tpc) * kST;
      // ---- phase 1: dots with gd and shifted moments of TPL tokens over this lane's share of the warp's 48 chunks
      {
        uint64_t dot2[TPL][HE], s1[TPL], s2[TPL], nx0[TPL];
        float x0[TPL];
#pragma unroll
        for (int i = 0; i < TPL; ++i) {
          x0[i] = lds32(tile_chunk(tile, tg + NG * i, 0));
          nx0[i] = f2_pack(-x0[i], -x0[i]);
          s1[i] = 0ull; s2[i] = 0ull;
#pragma unroll
          for (int v = 0; v < HE; ++v) dot2[i][v] = 0ull;
        }
#pragma unroll 2
        for (int c = 0; c < 48 / NE; ++c) {
          const int c4 = wa * 48 + NE * c + e;
          uint64_t t01[TPL], t23[TPL];
#pragma unroll
          for (int i = 0; i < TPL; ++i) {
            const float4 t = lds128(tile_chunk(tile, tg + NG * i, c4));
            t01[i] = f2_pack(t.x, t.y); t23[i] = f2_pack(t.z, t.w);
            const uint64_t a01 = f2_add(t01[i], nx0[i]), a23 = f2_add(t23[i], nx0[i]);
            s1[i] = f2_add(s1[i], f2_add(a01, a23));
            s2[i] = f2_fma(a01, a01, f2_fma(a23, a23, s2[i]));
          }
#pragma unroll
          for (int v = 0; v < HE; ++v) {
            const float4 gv = lds128(g_u + (v * kSD + c4 * 4) * 4);
            const uint64_t g01 = f2_pack(gv.x, gv.y), g23 = f2_pack(gv.z, gv.w);
#pragma unroll
            for (int i = 0; i < TPL; ++i) dot2[i][v] = f2_fma(t01[i], g01, f2_fma(t23[i], g23, dot2[i][v]));
          }
        }
        float x[Cfg::NV1];
#pragma unroll
        for (int i = 0; i < TPL; ++i) {
#pragma unroll
          for (int v = 0; v < HE; ++v) x[i * (HE + 2) + v] = f2_lo(dot2[i][v]) + f2_hi(dot2[i][v]);
          x[i * (HE + 2) + HE] = f2_lo(s1[i]) + f2_hi(s1[i]);
          x[i * (HE + 2) + HE + 1] = f2_lo(s2[i]) + f2_hi(s2[i]);
        }
        // halving butterfly over the NE lanes that share a token group
#pragma unroll
        for (int off = 16, n = Cfg::NV1; off >= NG; off >>= 1, n >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float lo = x[i], hi = x[i + n / 2];
            const float other = __shfl_xor_sync(0xffffffffu, up ? lo : hi, off);
            x[i] = (up ? hi : lo) + other;
          }
        }
#pragma unroll
        for (int j = 0; j < Cfg::NFIN; ++j) {
          const int idx = pbase + j, ti = idx / (HE + 2), v = idx - ti * (HE + 2);
          sts32(part_u + ((wa * kST + tg + NG * ti) * Cfg::PART_STRIDE + v) * 4, x[j]);
        }
        if (wa == 0 && e == 0) {
#pragma unroll
          for (int i = 0; i < TPL; ++i) sts32(part_u + ((tg + NG * i) * Cfg::PART_STRIDE + HE + 2) * 4, x0[i]);
        }
      }
      // this warp is done with the token tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_empty[st]);
      named_bar_sync(1 + sub, 128);                             // partials of the sub-group's four warps are visible
      // ---- phase 1b: thread <-> (token, slot) of one head
      const int buf = it & 1;
      mbar_wait(&w_empty[buf], ((it >> 1) & 1) ^ 1);
#pragma unroll
      for (int k = 0; k < UPW; ++k) {
        const int unit = wa + 4 * k;
        const int h = unit % 4, tgrp = unit / 4;
        const int s_l = lane / TPW, tk = tgrp * TPW + (lane % TPW);
        const int sh = h * S + s_l;
        const int ev = h * (S - 1) + (s_l > 0 ? s_l - 1 : 0);
        float dot = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t pp = part_u + ((w * kST + tk) * Cfg::PART_STRIDE) * 4;
          dot += lds32(pp + 4 * ev);
          const float2 sv = lds64(pp + 4 * HE);
          s1 += sv.x; s2 += sv.y;
        }
        const float x0 = lds32(part_u + (tk * Cfg::PART_STRIDE) * 4 + 4 * HE + 8);
        const float d1 = s1 * (1.0f / kSD);
        const float mu = x0 + d1;
        const float r = rsqrtf(fmaxf(s2 * (1.0f / kSD) - d1 * d1, 0.f) + p.eps);
        const int tok = tok_base + tk;
        const bool valid = tok < p.N;
        const float logit = s_l > 0 ? fmaf(r, dot - mu * Gd_[k], cd_[k]) : 0.f;      // relative to slot 0 of the head
        float mx = logit;
#pragma unroll
        for (int o = TPW; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float ex = expf(logit - mx);
        float sum = ex;
#pragma unroll
        for (int o = TPW; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float a = valid ? ex / sum : 0.f;
        const float w = a * r;
        const uint32_t wt = w_u + ((buf * kST + tk) * WS) * 4;
        if (s_l > 0) sts32(wt + 4 * ev, w);
        else if (h == 0) sts32(wt + 4 * HE, valid ? r : 0.f);   // the R row: weight r_j (0 for padded tokens)
        accA[k] += a;
        accM[k] = fmaf(w, mu, accM[k]);
        if (valid) {
          if (p.attn != nullptr) p.attn[((long long)b * HS + sh) * p.N + tok] = a;
          if (p.mu != nullptr && s_l == 0 && h == 0) { p.mu[(long long)b * p.N + tok] = mu; p.rstd[(long long)b * p.N + tok] = r; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_full[buf]);
      named_bar_sync(1 + sub, 128);                             // partial[] may be overwritten by the sub-group's next tile
    }
    if (cur_b >= 0) flush();
  } else {
    // =============================================================== group B
    const int tb = tid - 128 * Cfg::NSUB;
    const int u = tb & 127, hb = tb >> 7;                       // owns channel pairs 2u, 256 + 2u, 512 + 2u of TOKB tokens of a tile
    uint64_t acc[HE + 1][3];                                    // S-1 slots of every head, then R
    const int cchunk = u >> 1;                                  // 16-byte chunk index of the first pair (0..63)
    const int cin = (u & 1) * 8;                                // byte offset inside the chunk
    int cur_b = -1;
    auto flush = [&]() {                                        // U[h, s >= 1] += acc ;  U[h, 0] += R - sum_{s >= 1} acc[h, s]
      float* dst = p.U + (long long)cur_b * HS * kSD + 2 * u;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          uint64_t rest = acc[HE][k];
#pragma unroll
          for (int s = 1; s < S; ++s) {
            const uint64_t v = acc[h * (S - 1) + s - 1][k];
            rest = f2_add(rest, f2_mul(v, f2_pack(-1.0f, -1.0f)));
            red_add_v2_f32(dst + (h * S + s) * kSD + 256 * k, f2_lo(v), f2_hi(v));
          }
          red_add_v2_f32(dst + (h * S) * kSD + 256 * k, f2_lo(rest), f2_hi(rest));
        }
      }
    };
    for (int gt = start; gt < end; ++gt) {
      const int it = gt - start, st = it % Cfg::STAGES, buf = it & 1;
      const int b = gt / tpc;
      if (b != cur_b) {
        if (cur_b >= 0) flush();
        cur_b = b;
#pragma unroll
        for (int i = 0; i <= HE; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0ull;
      }
      mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
      mbar_wait(&w_full[buf], (it >> 1) & 1);
      const uint32_t tile = smem_u32(smem + Cfg::OFF_TILE + st * kSTileBytes);
#pragma unroll 4
      for (int jj = 0; jj < Cfg::TOKB; ++jj) {
        const int j = hb * Cfg::TOKB + jj;
        uint64_t t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float2 v = lds64(tile_chunk(tile, j, cchunk + 64 * k) + cin);
          t[k] = f2_pack(v.x, v.y);
        }
        const uint32_t wj = w_u + ((buf * kST + j) * WS) * 4;
        float wv[WS];
#pragma unroll
        for (int i4 = 0; i4 < HE / 4; ++i4) {
          const float4 q4 = lds128(wj + 16 * i4);
          wv[4 * i4] = q4.x; wv[4 * i4 + 1] = q4.y; wv[4 * i4 + 2] = q4.z; wv[4 * i4 + 3] = q4.w;
        }
        wv[HE] = lds32(wj + 4 * HE);
#pragma unroll
        for (int i = 0; i <= HE; ++i) {
          const uint64_t w2 = f2_pack(wv[i], wv[i]);
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[i][k] = f2_fma(w2, t[k], acc[i][k]);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tile_empty[st]);
        mbar_arrive(&w_empty[buf]);
      }
    }
    if (cur_b >= 0) flush();
  }
}

// =====================================================================================================================
// Backward of the streaming step.  Saved from the forward: the slot-softmax a[sh, j] and the token statistics (mu, r).
//   f[sh]    = dU[sh] . t_j + dm[sh] mu_j                      e[sh] = g[sh] . t_j - mu_j G[sh]
//   da[sh]   = r_j f[sh] + dA[sh] + dattn[sh, j]
//   dsim[sh] = a[sh] (da[sh] - sum_{s' in head} a[s'] da[s'])
//   dg[sh]  += r_j dsim[sh] t_j ;  dG[sh] -= r_j dsim[sh] mu_j ;  dc0[sh] += dsim[sh]
//   dr = sum_sh dsim[sh] e[sh] + a[sh] f[sh] ;  dmu = r_j sum_sh (a[sh] dm[sh] - dsim[sh] G[sh])
//   dt_j = sum_sh (r_j dsim[sh]) g[sh] + (r_j a[sh]) dU[sh]  +  dmu/D  -  dr r_j^3 (t_j - mu_j)/D
// Same tile ring / phase structure as the forward; phase 2 (thread <-> 4 channels) keeps its g / dU columns in registers,
// produces dt_j with coalesced 16-byte accesses (optionally accumulating onto the gradient of earlier layers) and
// accumulates dg in registers.
template <int HS>
struct SlotBwdCfg {
  static constexpr int STAGES = (HS <= 8) ? 3 : 2;
  static constexpr int OFF_TILE = 0;
  static constexpr int OFF_G = STAGES * kSTileBytes;                 // g[HS][768]
  static constexpr int OFF_DU = OFF_G + HS * kSD * 4;                // dU[HS][768]
  static constexpr int OFF_PART = OFF_DU + HS * kSD * 4;             // partial[8 warps][16 tokens][2 HS]
  static constexpr int OFF_COEF = OFF_PART + 8 * kST * 2 * HS * 4;   // coef[16 tokens][2 HS + 4]: alpha[HS], beta[HS], kappa, lambda
  static constexpr int COEF_STRIDE = 2 * HS + 4;
  static constexpr int OFF_ACC = OFF_COEF + kST * COEF_STRIDE * 4;   // running dG / dc0 partials [2][16][HS]
  static constexpr int OFF_BAR = OFF_ACC + 2 * kST * HS * 4;
  static constexpr int BYTES = OFF_BAR + 64 + 1024;
};

struct SlotBwdParams {
  int B, N, S;
  int tiles_per_cta, tiles_per_clip;
  const float* mu; const float* rstd;     // [B, N]
  const float* g; const float* G;         // [B, HS, 768], [B, HS]
  const float* a;                         // [B, HS, N]
  const float* dU; const float* dm; const float* dA;   // [B, HS, 768], [B, HS], [B, HS]
  const float* dattn;                     // [B, HS, N] or null
  float* dt; int accumulate;              // [B, N, 768]
  float* dg; float* dG; float* dc0;       // (+=)
  int hs_total, hs_off;                   // rows per clip in g / G / a / dU / ... and the first row this launch handles
};

// HEADS heads x S = HS / HEADS slots per launch; the (head, slot) rows handled are [hs_off, hs_off + HS) of hs_total per clip.
// Every term of dt / dmu / dr is a sum over (head, slot), so S = 8 (32 rows: g and dU of all heads do not fit in shared memory)
// runs as two launches over two heads each, the second accumulating onto the first.
template <int HS, int HEADS>
__global__ void __launch_bounds__(kSlotThreads, 1)
slot_stream_bwd_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotBwdParams p) {
  pdl_trigger();
  pdl_wait();
  using Cfg = SlotBwdCfg<HS>;
  constexpr int S = HS / HEADS;
  const long long row0 = (long long)blockIdx.y * p.hs_total + p.hs_off;      // first (head, slot) row of this clip / launch
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* g_s = reinterpret_cast<float*>(smem + Cfg::OFF_G);
  float* du_s = reinterpret_cast<float*>(smem + Cfg::OFF_DU);
  float* part = reinterpret_cast<float*>(smem + Cfg::OFF_PART);
  float* coef = reinterpret_cast<float*>(smem + Cfg::OFF_COEF);
  float* accG = reinterpret_cast<float*>(smem + Cfg::OFF_ACC);
  float* accC = accG + kST * HS;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.tiles_per_clip - tile0);
  if (ntiles <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  {
    const float4* sg = reinterpret_cast<const float4*>(p.g + row0 * kSD);
    const float4* sd = reinterpret_cast<const float4*>(p.dU + row0 * kSD);
    for (int i = tid; i < HS * kSD / 4; i += kSlotThreads) {
      reinterpret_cast<float4*>(g_s)[i] = __ldg(sg + i);
      reinterpret_cast<float4*>(du_s)[i] = __ldg(sd + i);
    }
    for (int i = tid; i < 2 * kST * HS; i += kSlotThreads) accG[i] = 0.f;
  }
  __syncthreads();

  auto issue = [&](int it) {
    const int st = it % Cfg::STAGES;
    mbar_arrive_expect_tx(&full[st], kSTileBytes);
    uint8_t* dst = smem + Cfg::OFF_TILE + st * kSTileBytes;
    const int tok0 = (tile0 + it) * kST;
    tma_load_4d(dst, &tmTok, &full[st], 0, tok0, 0, b);   // one bulk tensor copy: [24 channel boxes][16 tokens][32 floats]
  };
  if (tid == 0) {
    for (int it = 0; it < Cfg::STAGES - 1 && it < ntiles; ++it) issue(it);
  }

  const uint32_t g_u = smem_u32(g_s), du_u = smem_u32(du_s), coef_u = smem_u32(coef);
  // phase-2 state: thread t < 192 owns channels [4t, 4t+4): its g / dU columns and the dg accumulators
  float4 gr[HS], dur[HS], dgacc[HS];
  if (tid < kSD / 4) {
#pragma unroll
    for (int i = 0; i < HS; ++i) {
      gr[i] = lds128(g_u + (i * kSD + tid * 4) * 4);
      dur[i] = lds128(du_u + (i * kSD + tid * 4) * 4);
      dgacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float* Gs = p.G + row0;
  const float* dms = p.dm + row0;
  const float* dAs = p.dA + row0;

  const int tok_l = lane & 15, half = lane >> 4;
  const int slice = warp * 2 + half;
  for (int it = 0; it < ntiles; ++it) {
    const int st = it % Cfg::STAGES;
    if (tid == 0 && it + Cfg::STAGES - 1 < ntiles) issue(it + Cfg::STAGES - 1);
    mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
    const uint32_t tile = smem_u32(smem + Cfg::OFF_TILE + st * kSTileBytes);
    const int tok_base = (tile0 + it) * kST;

    // ---------------- phase 1: g . t and dU . t over this lane's 48 channels
    {
      float dg_[HS], df_[HS];
#pragma unroll
      for (int i = 0; i < HS; ++i) { dg_[i] = 0.f; df_[i] = 0.f; }
#pragma unroll 2
      for (int c = 0; c < 12; ++c) {
        const int c4 = slice * 12 + c;
        const float4 t = lds128(tile_chunk(tile, tok_l, c4));
#pragma unroll
        for (int i = 0; i < HS; ++i) {
          const float4 gv = lds128(g_u + (i * kSD + c4 * 4) * 4);
          const float4 dv_ = lds128(du_u + (i * kSD + c4 * 4) * 4);
          dg_[i] = fmaf(t.x, gv.x, fmaf(t.y, gv.y, fmaf(t.z, gv.z, fmaf(t.w, gv.w, dg_[i]))));
          df_[i] = fmaf(t.x, dv_.x, fmaf(t.y, dv_.y, fmaf(t.z, dv_.z, fmaf(t.w, dv_.w, df_[i]))));
        }
      }
#pragma unroll
      for (int i = 0; i < HS; ++i) {
        dg_[i] += __shfl_xor_sync(0xffffffffu, dg_[i], 16);
        df_[i] += __shfl_xor_sync(0xffffffffu, df_[i], 16);
      }
      if (half == 0) {
        float* pp = part + (warp * kST + tok_l) * (2 * HS);
#pragma unroll
        for (int i = 0; i < HS; ++i) { pp[i] = dg_[i]; pp[HS + i] = df_[i]; }
      }
    }
    __syncthreads();
    // ---------------- phase 1b: warp 0, lanes 0..15 <-> tokens: per-token coefficients
    if (warp == 0 && lane < kST) {
      float e[HS], f[HS];
#pragma unroll
      for (int i = 0; i < HS; ++i) { e[i] = 0.f; f[i] = 0.f; }
#pragma unroll 2
      for (int sl = 0; sl < 8; ++sl) {
        const float* pp = part + (sl * kST + lane) * (2 * HS);
#pragma unroll
        for (int i = 0; i < HS; ++i) { e[i] += pp[i]; f[i] += pp[HS + i]; }
      }
      const int tok = tok_base + lane;
      const bool valid = tok < p.N;
      const float mu = valid ? __ldg(p.mu + (long long)b * p.N + tok) : 0.f;
      const float r = valid ? __ldg(p.rstd + (long long)b * p.N + tok) : 0.f;
      float dr = 0.f, dmu = 0.f;
      float* cf = coef + lane * Cfg::COEF_STRIDE;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        float av[S], da[S];
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int i = h * S + s;
          av[s] = valid ? __ldg(p.a + (row0 + i) * p.N + tok) : 0.f;
          e[i] -= mu * __ldg(Gs + i);
          f[i] += mu * __ldg(dms + i);
          da[s] = fmaf(r, f[i], __ldg(dAs + i));
          if (p.dattn != nullptr && valid) da[s] += __ldg(p.dattn + (row0 + i) * p.N + tok);
          dot = fmaf(av[s], da[s], dot);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int i = h * S + s;
          const float dsim = av[s] * (da[s] - dot);
          dr += dsim * e[i] + av[s] * f[i];
          dmu += av[s] * __ldg(dms + i) - dsim * __ldg(Gs + i);
          const float alpha = r * dsim;
          cf[i] = alpha;
          cf[HS + i] = r * av[s];
          accG[lane * HS + i] -= alpha * mu;
          accC[lane * HS + i] += dsim;
        }
      }
      dmu *= r;
      const float lambda = -dr * r * r * r * (1.0f / kSD);
      cf[2 * HS] = dmu * (1.0f / kSD) - lambda * mu;   // kappa'
      cf[2 * HS + 1] = lambda;
    }
    __syncthreads();
    // ---------------- phase 2: dt_j and dg accumulation, thread <-> 4 channels
    if (tid < kSD / 4) {
#pragma unroll 2
      for (int j = 0; j < kST; ++j) {
        const int tok = tok_base + j;
        if (tok >= p.N) break;
        const float4 t = lds128(tile_chunk(tile, j, tid));
        const uint32_t cj = coef_u + j * Cfg::COEF_STRIDE * 4;
        const float2 kl = lds64(cj + 2 * HS * 4);
        float4 o = make_float4(fmaf(kl.y, t.x, kl.x), fmaf(kl.y, t.y, kl.x), fmaf(kl.y, t.z, kl.x), fmaf(kl.y, t.w, kl.x));
        float* dst = p.dt + ((long long)b * p.N + tok) * kSD + tid * 4;
        float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.accumulate) old = *reinterpret_cast<const float4*>(dst);
#pragma unroll
        for (int i4 = 0; i4 < HS / 4; ++i4) {
          const float4 al = lds128(cj + 16 * i4);
          const float4 be = lds128(cj + HS * 4 + 16 * i4);
          const float alv[4] = {al.x, al.y, al.z, al.w};
          const float bev[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = 4 * i4 + u;
            o.x = fmaf(alv[u], gr[i].x, fmaf(bev[u], dur[i].x, o.x));
            o.y = fmaf(alv[u], gr[i].y, fmaf(bev[u], dur[i].y, o.y));
            o.z = fmaf(alv[u], gr[i].z, fmaf(bev[u], dur[i].z, o.z));
            o.w = fmaf(alv[u], gr[i].w, fmaf(bev[u], dur[i].w, o.w));
            dgacc[i].x = fmaf(alv[u], t.x, dgacc[i].x); dgacc[i].y = fmaf(alv[u], t.y, dgacc[i].y);
            dgacc[i].z = fmaf(alv[u], t.z, dgacc[i].z); dgacc[i].w = fmaf(alv[u], t.w, dgacc[i].w);
          }
        }
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        *reinterpret_cast<float4*>(dst) = o;
      }
    }
    __syncthreads();
  }
  if (tid < kSD / 4) {
    float* dst = p.dg + row0 * kSD + tid * 4;
#pragma unroll
    for (int i = 0; i < HS; ++i) red_add_v4_f32(dst + i * kSD, dgacc[i].x, dgacc[i].y, dgacc[i].z, dgacc[i].w);
  }
  if (tid < HS) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int j = 0; j < kST; ++j) { a += accG[j * HS + tid]; c += accC[j * HS + tid]; }
    atomicAdd(p.dG + row0 + tid, a);
    atomicAdd(p.dc0 + row0 + tid, c);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward, S = 2 (the DEVIAS recipes): second generation.
//  * softmax shift invariance again: with two slots per head  sum_s dsim[h,s] = 0  and  a[h,0] = 1 - a[h,1],  so per head only
//    gd = g[h,1] - g[h,0],  dUd = dU[h,1] - dU[h,0]  and the single vector sum_h dU[h,0] enter dt, and dg[h,0] = -dg[h,1]:
//        dt_j = kappa + lambda t_j + sum_h alpha_h gd_h + sum_h beta_h dUd_h + r_j sum_h dU[h,0]      (alpha = r dsim_1, beta = r a_1)
//    14 packed FFMA2 per channel pair and token instead of 25 scalar FMA per channel;
//  * persistent CTAs: the B x 98 token tiles are split evenly over the SMs (a CTA's range may span clips; the per-clip
//    vectors are reloaded at the boundary while the TMA ring keeps running), no wave quantisation;
//  * the saved a / mu / r of a tile are fetched before its dot products (latency hidden), the per-token coefficient
//    math is spread over all 256 threads, and dt leaves through st / red.global.add.v2 (no read of the running gradient).
struct SlotBwd2Cfg {
  static constexpr int THREADS = 256;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int STAGES = 3;
  static constexpr int NV = 12;                                       // gd[4], dU[8]
  static constexpr int OFF_TILE = 0;
  // two groups of four warps take alternate tiles (each runs all three phases of its tile, so the phases of one group
  // fill the latencies of the other); vec / partial / coef exist once per group
  static constexpr int OFF_VEC = STAGES * kSTileBytes;                // vec[2][12][768]
  static constexpr int OFF_PART = OFF_VEC + 2 * NV * kSD * 4;         // partial[2][4 warps][16 tokens][12]
  static constexpr int OFF_COEF = OFF_PART + 2 * 4 * kST * NV * 4;    // coef[2][16 tokens][12]: alpha[4], beta[4], r, lambda, kappa, pad
  static constexpr int OFF_BAR = OFF_COEF + 2 * kST * 12 * 4;
  static constexpr int BYTES = OFF_BAR + 64 + 1024;
};

__device__ __forceinline__ uint64_t f2_dup(float w) { return f2_pack(w, w); }

// What bounds these kernels on the SM side is the shared-memory -> register return path (128 B per clock per SM): an
// LDS.128 costs four of its cycles even when every lane reads the same 16 bytes.  Hence (a) in the dot-product phase a lane
// owns FOUR tokens, so one broadcast load of a slot-vector chunk feeds 8 FFMA2 instead of 2, and the 48 partial sums are
// combined over the 8 lanes sharing a token with a halving butterfly (42 shuffles, not 144); (b) the per-token
// coefficients are stored once (not as duplicated FFMA2 pairs) and each thread applies them to three channel pairs.
__global__ void __launch_bounds__(SlotBwd2Cfg::THREADS, 1)
slot_stream_bwd2_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotBwdParams p) {
  pdl_trigger();
  pdl_wait();
  using Cfg = SlotBwd2Cfg;
  constexpr int HS = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> 7, tg = tid & 127, warp = (tid >> 5) & 3;     // group, thread and warp within the group
  const uint32_t vec_u = smem_u32(smem + Cfg::OFF_VEC) + grp * (Cfg::NV * kSD * 4);
  const uint32_t part_u = smem_u32(smem + Cfg::OFF_PART) + grp * (4 * kST * Cfg::NV * 4);
  const uint32_t coef_u = smem_u32(smem + Cfg::OFF_COEF) + grp * (kST * 12 * 4);
  const int tpc = p.tiles_per_clip;
  const long long total = (long long)p.B * tpc;
  const int start = (int)(total * blockIdx.x / gridDim.x), end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (start >= end) return;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int it) {                   // ring index it <-> global tile start + it
    const int gt = start + it, st = it % Cfg::STAGES;
    mbar_arrive_expect_tx(&full[st], kSTileBytes);
    tma_load_4d(smem + Cfg::OFF_TILE + st * kSTileBytes, &tmTok, &full[st], 0, (gt % tpc) * kST, 0, gt / tpc);
  };
  if (tid == 0) {
    for (int it = 0; it < Cfg::STAGES && start + it < end; ++it) issue(it);
  }

  // phase 1: lane <-> (token quad {tq, tq+4, tq+8, tq+12}, eighth e); a warp covers 48 of the 192 channel chunks
  const int tq = lane & 3, e8 = lane >> 2;
  // after the butterfly the lane holds 6 of the quad's 48 sums: indices pbase .. pbase+5 (index = token_i * 12 + vector)
  const int pbase = ((lane >> 4) & 1) * 24 + ((lane >> 3) & 1) * 12 + ((lane >> 2) & 1) * 6;
  // phase 1b: thread <-> (token, head, slice); slice s sums the partials of warps 2s, 2s+1
  const int tk1 = tg >> 3, h1 = (tg >> 1) & 3, sl1 = tg & 1;
  // phase 2: thread <-> channel pairs 2u, 256 + 2u, 512 + 2u of all 16 tokens
  const int u = tg;
  const int cchunk = u >> 1, cin = (u & 1) * 8;

  for (int gt = start + grp; gt < end;) {
    const int b = gt / tpc;
    const int seg_end = min(end, (b + 1) * tpc);
    named_bar_sync(1 + grp, 128);                                    // previous clip: phase 2 and its register reads are done
    {  // vec rows 0..3: gd[h] = g[h,1] - g[h,0];  rows 4..11: dU[h,s]
      const float4* sg = reinterpret_cast<const float4*>(p.g + (long long)b * HS * kSD);
      const float4* sd = reinterpret_cast<const float4*>(p.dU + (long long)b * HS * kSD);
      for (int i = tg; i < 4 * (kSD / 4); i += 128) {
        const int h = i / (kSD / 4), c = i % (kSD / 4);
        const float4 a1 = __ldg(sg + (2 * h + 1) * (kSD / 4) + c), a0 = __ldg(sg + (2 * h) * (kSD / 4) + c);
        sts128f(vec_u + i * 16, make_float4(a1.x - a0.x, a1.y - a0.y, a1.z - a0.z, a1.w - a0.w));
      }
      for (int i = tg; i < HS * (kSD / 4); i += 128) sts128f(vec_u + (4 * (kSD / 4) + i) * 16, __ldg(sd + i));
    }
    named_bar_sync(1 + grp, 128);
    uint64_t gd[4][3], dud[4][3], du0[3], dgacc[4][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ch = 2 * u + 256 * k;
      uint64_t z = 0ull;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 gv = lds64(vec_u + (h * kSD + ch) * 4);
        const float2 d0 = lds64(vec_u + ((4 + 2 * h) * kSD + ch) * 4), d1 = lds64(vec_u + ((5 + 2 * h) * kSD + ch) * 4);
        gd[h][k] = f2_pack(gv.x, gv.y);
        dud[h][k] = f2_pack(d1.x - d0.x, d1.y - d0.y);
        z = f2_add(z, f2_pack(d0.x, d0.y));
        dgacc[h][k] = 0ull;
      }
      du0[k] = z;
    }
    const float G0 = __ldg(p.G + b * HS + 2 * h1), G1 = __ldg(p.G + b * HS + 2 * h1 + 1);
    const float dm0 = __ldg(p.dm + b * HS + 2 * h1), dm1 = __ldg(p.dm + b * HS + 2 * h1 + 1);
    const float dA0 = __ldg(p.dA + b * HS + 2 * h1), dA1 = __ldg(p.dA + b * HS + 2 * h1 + 1);
    float accG0 = 0.f, accG1 = 0.f, accC0 = 0.f, accC1 = 0.f;
    // saved forward values of this thread's (token, head), fetched ONE TILE AHEAD (their DRAM latency would otherwise sit
    // on the critical path of every tile)
    float n_a0 = 0.f, n_a1 = 0.f, n_mu = 0.f, n_r = 0.f, n_d0 = 0.f, n_d1 = 0.f;
    auto fetch = [&](int g_tile) {
      const int tok1 = (g_tile % tpc) * kST + tk1;
      const bool v = g_tile < seg_end && tok1 < p.N;
      const long long arow = ((long long)b * HS + 2 * h1) * p.N + tok1;
      n_a0 = v ? __ldg(p.a + arow) : 0.f; n_a1 = v ? __ldg(p.a + arow + p.N) : 0.f;
      n_mu = v ? __ldg(p.mu + (long long)b * p.N + tok1) : 0.f;
      n_r = v ? __ldg(p.rstd + (long long)b * p.N + tok1) : 0.f;
      n_d0 = (v && p.dattn != nullptr) ? __ldg(p.dattn + arow) : 0.f;
      n_d1 = (v && p.dattn != nullptr) ? __ldg(p.dattn + arow + p.N) : 0.f;
    };
    fetch(gt);

    for (; gt < seg_end; gt += 2) {
      const int it = gt - start, st = it % Cfg::STAGES;
      const int tok_base = (gt % tpc) * kST;
      const float a0 = n_a0, a1 = n_a1, mu = n_mu, r = n_r;
      float da0 = dA0 + n_d0, da1 = dA1 + n_d1;
      fetch(gt + 2);

      mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
      const uint32_t tile = smem_u32(smem + Cfg::OFF_TILE + st * kSTileBytes);
      // ---------------- phase 1: the 12 dot products of four tokens over this lane's 6 chunks
      {
        uint64_t d[4][Cfg::NV];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int v = 0; v < Cfg::NV; ++v) d[i][v] = 0ull;
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          const int c4 = warp * 48 + 8 * c + e8;
          uint64_t t01[4], t23[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = lds128(tile_chunk(tile, tq + 4 * i, c4));
            t01[i] = f2_pack(t.x, t.y); t23[i] = f2_pack(t.z, t.w);
          }
#pragma unroll
          for (int v = 0; v < Cfg::NV; ++v) {
            const float4 x = lds128(vec_u + (v * kSD + c4 * 4) * 4);
            const uint64_t v01 = f2_pack(x.x, x.y), v23 = f2_pack(x.z, x.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i][v] = f2_fma(t01[i], v01, f2_fma(t23[i], v23, d[i][v]));
          }
        }
        float x[48];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int v = 0; v < Cfg::NV; ++v) x[i * 12 + v] = f2_lo(d[i][v]) + f2_hi(d[i][v]);
        // halving butterfly over the 8 lanes of a token quad (lane bits 4, 3, 2)
#pragma unroll
        for (int off = 16, n = 48; off >= 4; off >>= 1, n >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float lo = x[i], hi = x[i + n / 2];
            const float other = __shfl_xor_sync(0xffffffffu, up ? lo : hi, off);
            x[i] = (up ? hi : lo) + other;
          }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int idx = pbase + j, ti = idx / 12, v = idx - ti * 12;
          sts32(part_u + ((warp * kST + tq + 4 * ti) * Cfg::NV + v) * 4, x[j]);
        }
      }
      named_bar_sync(1 + grp, 128);                                  // partials of the group's four warps are visible
      // ---------------- phase 1b: per-token coefficients
      {
        float ed = 0.f, f0 = 0.f, f1 = 0.f;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t pp = part_u + (((2 * sl1 + w) * kST + tk1) * Cfg::NV) * 4;
          ed += lds32(pp + 4 * h1);
          const float2 fv = lds64(pp + 4 * (4 + 2 * h1));
          f0 += fv.x; f1 += fv.y;
        }
        ed += __shfl_xor_sync(0xffffffffu, ed, 1); f0 += __shfl_xor_sync(0xffffffffu, f0, 1); f1 += __shfl_xor_sync(0xffffffffu, f1, 1);
        ed -= mu * (G1 - G0);                                        // e[h,1] - e[h,0]
        f0 = fmaf(mu, dm0, f0); f1 = fmaf(mu, dm1, f1);
        da0 = fmaf(r, f0, da0); da1 = fmaf(r, f1, da1);
        const float dot = fmaf(a0, da0, a1 * da1);
        const float ds0 = a0 * (da0 - dot), ds1 = a1 * (da1 - dot);
        float dr = fmaf(ds1, ed, fmaf(a0, f0, a1 * f1));
        float dmu = fmaf(a0, dm0, a1 * dm1) - fmaf(ds0, G0, ds1 * G1);
        dr += __shfl_xor_sync(0xffffffffu, dr, 2); dmu += __shfl_xor_sync(0xffffffffu, dmu, 2);
        dr += __shfl_xor_sync(0xffffffffu, dr, 4); dmu += __shfl_xor_sync(0xffffffffu, dmu, 4);
        dmu *= r;
        const float lambda = -dr * r * r * r * (1.0f / kSD);
        const float kappa = dmu * (1.0f / kSD) - lambda * mu;
        const float alpha = r * ds1, beta = r * a1;
        if (sl1 == 0) {
          const uint32_t cf = coef_u + tk1 * 48;
          sts32(cf + 4 * h1, alpha);
          sts32(cf + 4 * (4 + h1), beta);
          if (h1 == 0) { sts32(cf + 32, r); sts32(cf + 36, lambda); sts32(cf + 40, kappa); }
          accG0 = fmaf(-r * ds0, mu, accG0); accG1 = fmaf(-alpha, mu, accG1);
          accC0 += ds0; accC1 += ds1;
        }
      }
      named_bar_sync(1 + grp, 128);
      // ---------------- phase 2: dt and the dg accumulators
      {
#pragma unroll 2
        for (int j = 0; j < kST; ++j) {
          const int tok = tok_base + j;
          if (tok >= p.N) break;
          uint64_t t[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float2 v = lds64(tile_chunk(tile, j, cchunk + 64 * k) + cin);
            t[k] = f2_pack(v.x, v.y);
          }
          const uint32_t cf = coef_u + j * 48;
          const float4 c_a = lds128(cf), c_b = lds128(cf + 16), c_r = lds128(cf + 32);
          const uint64_t al[4] = {f2_dup(c_a.x), f2_dup(c_a.y), f2_dup(c_a.z), f2_dup(c_a.w)};
          const uint64_t be[4] = {f2_dup(c_b.x), f2_dup(c_b.y), f2_dup(c_b.z), f2_dup(c_b.w)};
          const uint64_t rr = f2_dup(c_r.x), lam = f2_dup(c_r.y), kap = f2_dup(c_r.z);
          float* dst = p.dt + ((long long)b * p.N + tok) * kSD + 2 * u;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            uint64_t o = f2_fma(lam, t[k], kap);
            uint64_t o2 = f2_mul(rr, du0[k]);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              o = f2_fma(al[h], gd[h][k], o);
              o2 = f2_fma(be[h], dud[h][k], o2);
              dgacc[h][k] = f2_fma(al[h], t[k], dgacc[h][k]);
            }
            o = f2_add(o, o2);
            if (p.accumulate) red_add_v2_f32(dst + 256 * k, f2_lo(o), f2_hi(o));
            else st_global_v2(dst + 256 * k, o);
          }
        }
      }
      named_bar_sync(1 + grp, 128);                                  // the group is done with the stage: refill it (tile it + 3,
      if (tg == 0 && gt + Cfg::STAGES < end) issue(it + Cfg::STAGES); // which the other group will consume)
    }
    // ---- flush this clip's share: dg[h,1] += acc, dg[h,0] -= acc; dG, dc0
    {
      float* dst = p.dg + (long long)b * HS * kSD + 2 * u;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float x = f2_lo(dgacc[h][k]), y = f2_hi(dgacc[h][k]);
          red_add_v2_f32(dst + (2 * h + 1) * kSD + 256 * k, x, y);
          red_add_v2_f32(dst + (2 * h) * kSD + 256 * k, -x, -y);
        }
      }
      // the four tokens of a warp (lanes l, l ^ 8, l ^ 16, l ^ 24), then one atomic per (warp, head, slot)
      accG0 += __shfl_xor_sync(0xffffffffu, accG0, 8); accG1 += __shfl_xor_sync(0xffffffffu, accG1, 8);
      accC0 += __shfl_xor_sync(0xffffffffu, accC0, 8); accC1 += __shfl_xor_sync(0xffffffffu, accC1, 8);
      accG0 += __shfl_xor_sync(0xffffffffu, accG0, 16); accG1 += __shfl_xor_sync(0xffffffffu, accG1, 16);
      accC0 += __shfl_xor_sync(0xffffffffu, accC0, 16); accC1 += __shfl_xor_sync(0xffffffffu, accC1, 16);
      if (lane < 8 && sl1 == 0) {
        atomicAdd(p.dG + b * HS + 2 * h1, accG0); atomicAdd(p.dG + b * HS + 2 * h1 + 1, accG1);
        atomicAdd(p.dc0 + b * HS + 2 * h1, accC0); atomicAdd(p.dc0 + b * HS + 2 * h1 + 1, accC1);
      }
    }
  }
}

static int launch_slot_bwd2(const CUtensorMap& tm, const SlotBwdParams& p, cudaStream_t s) {
  using Cfg = SlotBwd2Cfg;
  static_assert(Cfg::BYTES <= 227 * 1024, "slot backward does not fit in shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    attr_done = true;
  }
  const long long total = (long long)p.B * p.tiles_per_clip;
  long long grid = sm_count();
  if (grid > (total + 1) / 2) grid = (total + 1) / 2;             // at least two tiles per CTA
  const double bytes = (double)p.B * p.N * kSD * 4 * (p.accumulate ? 3.0 : 2.0);
  const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
  DV_CHECK_CUDA(launch_k(slot_stream_bwd2_kernel, dim3((unsigned)grid), dim3((unsigned)(Cfg::THREADS)), (size_t)(Cfg::BYTES), s, tm, p));
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

template <int HS, int HEADS>
static int launch_slot_bwd(const CUtensorMap& tm, const SlotBwdParams& p, int splits, cudaStream_t s) {
  using Cfg = SlotBwdCfg<HS>;
  static_assert(Cfg::BYTES <= 227 * 1024, "slot backward does not fit in shared memory");
  auto kern = slot_stream_bwd_kernel<HS, HEADS>;
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    attr_done = true;
  }
  const double bytes = (double)p.B * p.N * kSD * 4 * (p.accumulate ? 3.0 : 2.0);
  const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
  DV_CHECK_CUDA(launch_k(kern, dim3(splits, p.B), dim3((unsigned)(kSlotThreads)), (size_t)(Cfg::BYTES), s, tm, p));
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

// 4-D view of the tokens [B, N, 768] as (32 floats | N tokens | 24 channel boxes | B): a single box [32, 16, 24, 1] lands in
// shared memory as [channel box][token][32 floats] with the 128-byte swizzle keyed on the token index -- the layout both
// access patterns of the kernels want -- with ONE TMA instruction per tile (24 separate 2 KiB boxes per tile throttled the
// first version to ~5000 cycles per tile).
static int make_token_tmap(CUtensorMap* tm, const float* tokens, int B, int N) {
  const uint64_t dims[4] = {32, (uint64_t)N, (uint64_t)kSBoxes, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)kSD * 4, 128, (uint64_t)N * kSD * 4};
  const uint32_t box[4] = {32, kST, kSBoxes, 1};
  return make_tmap_nd(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, tokens, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int HS, bool V2>
static int launch_slot_fwd(const CUtensorMap& tm, const SlotParams& p, int splits, cudaStream_t s) {
  const double bytes = (double)p.B * p.N * kSD * 4 + (p.attn ? (double)p.B * HS * p.N * 4 : 0.0);
  static bool attr_done = false;
  if constexpr (V2) {
    using Cfg = SlotCfg2<HS>;
    static_assert(Cfg::BYTES <= 227 * 1024, "slot forward does not fit in shared memory");
    if (!attr_done) {
      DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_fwd2_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
      attr_done = true;
    }
    const long long total = (long long)p.B * p.tiles_per_clip;     // persistent: an even share of all tiles per SM
    long long grid = sm_count();
    if (grid > (total + 1) / 2) grid = (total + 1) / 2;
    const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
    DV_CHECK_CUDA(launch_k(slot_stream_fwd2_kernel<HS>, dim3((unsigned)grid), dim3((unsigned)(Cfg::THREADS)), (size_t)(Cfg::BYTES), s, tm, p));
    prof_end(prof, s);
  } else {
    using Cfg = SlotCfg<HS>;
    static_assert(Cfg::BYTES <= 227 * 1024, "slot forward does not fit in shared memory");
    if (!attr_done) {
      DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_fwd_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
      attr_done = true;
    }
    const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
    DV_CHECK_CUDA(launch_k(slot_stream_fwd_kernel<HS>, dim3(splits, p.B), dim3((unsigned)(kSlotThreads)), (size_t)(Cfg::BYTES), s, tm, p));
    prof_end(prof, s);
  }
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}


// Diagnostic: the token stream alone (same tensor map, ring and persistent tile split as the kernels above, no arithmetic):
// the ceiling the TMA path itself gives the streaming kernels.  One thread issues, every warp touches each tile once.
__global__ void __launch_bounds__(256, 1)
slot_stream_probe_kernel(const __grid_constant__ CUtensorMap tmTok, int B, int tpc, int stages, float* out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * kSTileBytes);
  uint64_t* empty = full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = (long long)B * tpc;
  const int start = (int)(total * blockIdx.x / gridDim.x), end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (start >= end) return;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int it) {
    const int gt = start + it, st = it % stages;
    mbar_wait(&empty[st], ((it / stages) & 1) ^ 1);
    mbar_arrive_expect_tx(&full[st], kSTileBytes);
    tma_load_4d(smem + st * kSTileBytes, &tmTok, &full[st], 0, (gt % tpc) * kST, 0, gt / tpc);
  };
  const int n = end - start;
  if (tid == 0) for (int it = 0; it < stages - 1 && it < n; ++it) issue(it);
  float acc = 0.f;
  for (int it = 0; it < n; ++it) {
    const int st = it % stages;
    if (tid == 0 && it + stages - 1 < n) issue(it + stages - 1);
    mbar_wait(&full[st], (it / stages) & 1);
    acc += lds32(smem_u32(smem + st * kSTileBytes) + tid * 16);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  if (acc == 123.456f) out[0] = acc;
  (void)warp;
}

}  // namespace dv

extern "C" int devias_slot_stream_fwd(const float* tokens, const float* g, const float* G, const float* c0, float* U, float* m,
                                      float* A, float* attn, float* mu, float* rstd, int batch, int n_tokens, int dim,
                                      int num_slots, float eps, void* stream) {
  using namespace dv;
  DV_REQUIRE(tokens && g && G && c0 && U && m && A, "null pointer");
  DV_REQUIRE(dim == kSD, "token dim must be 768");
  DV_REQUIRE(num_slots == 2 || num_slots == 4 || num_slots == 8, "num_slots must be 2, 4 or 8 (4 heads x S query vectors)");
  DV_REQUIRE((mu == nullptr) == (rstd == nullptr), "mu and rstd go together");
  DV_REQUIRE(batch > 0 && n_tokens > 0, "empty problem");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm;
  int rc = make_token_tmap(&tm, tokens, batch, n_tokens);
  if (rc) return rc;
  const int tiles = (n_tokens + kST - 1) / kST;
  // enough CTAs for ~2 waves over the SMs, but at least 4 tiles each so the per-CTA flush stays cheap
  int splits = (2 * sm_count() + batch - 1) / batch;
  if (splits > (tiles + 3) / 4) splits = (tiles + 3) / 4;
  if (splits < 1) splits = 1;
  const int per = (tiles + splits - 1) / splits;
  splits = (tiles + per - 1) / per;
  SlotParams p{batch, n_tokens, num_slots, per, tiles, g, G, c0, U, m, A, attn, mu, rstd, eps};
  switch (num_slots) {
    case 2: return launch_slot_fwd<8, true>(tm, p, splits, s);
    case 4: return launch_slot_fwd<16, true>(tm, p, splits, s);
    default: return launch_slot_fwd<32, false>(tm, p, splits, s);
  }
}

extern "C" int devias_slot_stream_bwd(const float* tokens, const float* mu, const float* rstd, const float* g, const float* G,
                                      const float* attn, const float* dU, const float* dm, const float* dA, const float* dattn,
                                      float* dtokens, int accumulate_dtokens, float* dg, float* dG, float* dc0, int batch,
                                      int n_tokens, int dim, int num_slots, void* stream) {
  using namespace dv;
  DV_REQUIRE(tokens && mu && rstd && g && G && attn && dU && dm && dA && dtokens && dg && dG && dc0, "null pointer");
  DV_REQUIRE(dim == kSD, "token dim must be 768");
  DV_REQUIRE(num_slots == 2 || num_slots == 4 || num_slots == 8, "num_slots must be 2, 4 or 8");
  DV_REQUIRE(batch > 0 && n_tokens > 0, "empty problem");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm;
  int rc = make_token_tmap(&tm, tokens, batch, n_tokens);
  if (rc) return rc;
  const int tiles = (n_tokens + kST - 1) / kST;
  int splits = (2 * sm_count() + batch - 1) / batch;
  if (splits > (tiles + 3) / 4) splits = (tiles + 3) / 4;
  if (splits < 1) splits = 1;
  const int per = (tiles + splits - 1) / splits;
  splits = (tiles + per - 1) / per;
  SlotBwdParams p{batch, n_tokens, num_slots, per, tiles, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens, accumulate_dtokens,
                  dg, dG, dc0, 4 * num_slots, 0};
  if (num_slots == 2) return launch_slot_bwd2(tm, p, s);
  if (num_slots == 4) return launch_slot_bwd<16, 4>(tm, p, splits, s);
  // S = 8: heads {0, 1} then heads {2, 3}; the second pass adds its share of dt onto the first
  rc = launch_slot_bwd<16, 2>(tm, p, splits, s);
  if (rc) return rc;
  p.hs_off = 16;
  p.accumulate = 1;
  return launch_slot_bwd<16, 2>(tm, p, splits, s);
}

extern "C" int devias_debug_token_stream(const float* tokens, int batch, int n_tokens, int stages, float* scratch, void* stream) {
  using namespace dv;
  DV_REQUIRE(tokens && scratch && batch > 0 && n_tokens > 0 && stages >= 2 && stages <= 4, "bad arguments");
  CUtensorMap tm;
  int rc = make_token_tmap(&tm, tokens, batch, n_tokens);
  if (rc) return rc;
  const int tiles = (n_tokens + kST - 1) / kST;
  const int bytes = stages * kSTileBytes + 256 + 1024;
  DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  long long grid = sm_count();
  const long long total = (long long)batch * tiles;
  if (grid > total) grid = total;
  DV_CHECK_CUDA(launch_k(slot_stream_probe_kernel, dim3((unsigned)grid), dim3((unsigned)(256)), (size_t)(bytes), (cudaStream_t)stream, tm, batch, tiles, stages, scratch));
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}
