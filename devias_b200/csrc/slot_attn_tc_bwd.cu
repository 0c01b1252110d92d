// Backward of the streaming slot attention for BF16 context tokens on the tensor cores (the gradient of slot_attn_tc.cu; same
// folded contract as slot_attn.cu's fp32 backward: agg_block/attention.py:32-40,120-141 differentiated).
//
//   f[sh]    = dU[sh] . t_j + dm[sh] mu_j                      e[sh] = g[sh] . t_j - mu_j G[sh]
//   da[sh]   = r_j f[sh] + dA[sh] + dattn[sh, j]               dsim[sh] = a[sh] (da[sh] - sum_{s' in head} a[s'] da[s'])
//   dg[sh]  += r_j dsim[sh] t_j ;  dG[sh] -= r_j dsim[sh] mu_j ;  dc0[sh] += dsim[sh]
//   dt_j     = sum_sh alpha[sh] g[sh] + beta[sh] dU[sh]  +  kappa  +  lambda t_j          (alpha = r dsim, beta = r a)
//
// Three contractions per 32-token tile, all on tcgen05 off the SAME TMA-landed token tile and ONE bf16 image of [g; dU]:
//   phase 1   D1[token, e | f] = t . [g; dU]^T       M = 64 (32 tokens used), N = 2 HS, K = 768; A = tile (K-major), B = image (K-major)
//   coef      per token: alpha, beta, kappa, lambda (fp32 math, thread <-> (token, two heads))
//   phase 2   dg[sh, c] += sum_tok alpha[sh, tok] t[tok, c]            A = tile read MN-major, B = alpha (K-major)
//   phase 3   dt^T[c, tok] = sum_r [g; dU][r, c] coef[tok, r]  +  sum_tok' t[tok', c] diag(lambda)[tok', tok]
//                            A = the image read MN-major (M = channels)   A = tile read MN-major, B = diag(lambda)
//             kappa is added when dt leaves tensor memory (fp32); dt goes out as st.global / red.global.add, 128 bytes per warp.
// Warp roles: 0, 1 / 4, 5 = coefficient warps (the pairs take alternate tiles), 2 = TMA producer, 3 = phase-1 issuer,
// 7 = phase-2/3 issuer, 4-6 = image of [g; dU], 8-11 = dt drain (and dg at the end of a clip).
#include "slot_tc.cuh"

namespace dv {

constexpr int kTbThreads = 384;   // 12 warps

template <int HS>
struct SlotTcBwdCfg {
  static constexpr int R = 2 * HS;                                   // rows of the image: g[HS], dU[HS]  (16 / 32 / 64)
  static constexpr int HSP = HS < 16 ? 16 : HS;                      // N of the dg product (phase 2 has M = 128: N % 16 == 0)
  static constexpr int STAGES = HS <= 16 ? 3 : 2;
  static constexpr int G_BOX = R * 128;                              // one channel box of the image: [R rows][128 B]
  static constexpr int OFF_TILE = 0;
  static constexpr int OFF_G = STAGES * kTTileBytes;
  static constexpr int W_BYTES = HSP * 128;                          // alpha[HSP rows][32 tokens] in 128-byte rows
  static constexpr int OFF_W = OFF_G + kTBoxes * G_BOX;
  static constexpr int C_BYTES = kTT * 128;                          // coef[32 tokens][alpha HS | beta HS] in 128-byte rows
  static constexpr int OFF_C = OFF_W + 2 * W_BYTES;
  static constexpr int OFF_L = OFF_C + 2 * C_BYTES;                  // diag(lambda)[32 tokens][32 tokens] in 128-byte rows
  static constexpr int OFF_KAP = OFF_L + 2 * C_BYTES;                // kappa[2][32 tokens] fp32
  static constexpr int OFF_VEC = OFF_KAP + 2 * kTT * 4;              // G[32], dm[32], dA[32] of the clip
  static constexpr int OFF_BAR = OFF_VEC + 3 * 32 * 4;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = 512;                         // everything: base address 0 (see slot_attn_tc.cu)
  static constexpr uint32_t TM_D1 = 0;                               // 2 x R columns
  static constexpr uint32_t TM_D2 = 2 * R;                           // 6 x HSP columns: dg[channel block][sh]
  static constexpr uint32_t TM_D3 = 2 * R + 6 * HSP;                 // 6 x 32 columns: dt^T[channel block][token]
  static_assert(TM_D3 + 6 * kTT <= 512, "tensor memory");
};

struct SlotTcBwdParams {
  int B, N, tiles_per_clip;
  const float* mu; const float* rstd;     // [B, N]
  const float* g; const float* G;         // [B, HS, 768], [B, HS]
  const float* a;                         // [B, HS, N]
  const float* dU; const float* dm; const float* dA;   // [B, HS, 768], [B, HS], [B, HS]
  const float* dattn;                     // [B, HS, N] or null
  float* dt; int accumulate;              // [B, N, 768]
  float* dg; float* dG; float* dc0;       // (+=)
};

__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int HS>
__global__ void __launch_bounds__(kTbThreads, 1)
slot_stream_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmTok, const SlotTcBwdParams p) {
  pdl_trigger();
  using Cfg = SlotTcBwdCfg<HS>;
  constexpr int S = HS / 4, R = Cfg::R, HSP = Cfg::HSP, ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                        // ST   TMA -> everyone
  uint64_t* tile_free = full + ST;              // ST   every MMA reading the tile retired
  uint64_t* d1_full = tile_free + ST;           // 2    phase-1 MMAs retired
  uint64_t* d1_free = d1_full + 2;              // 2    (2 coefficient warps) dots read out of tensor memory
  uint64_t* c_full = d1_free + 2;               // 2    (2 coefficient warps) alpha / coef / lambda / kappa of the tile written
  uint64_t* c_free = c_full + 2;                // 2    phase-2/3 MMAs that read them retired
  uint64_t* kap_free = c_free + 2;              // 2    (4 drain warps) kappa consumed
  uint64_t* d3_full = kap_free + 2;             // 1    dt^T of a tile is in tensor memory
  uint64_t* d3_free = d3_full + 1;              // 1    (4 drain warps) ... and has been read out
  uint64_t* d2_full = d3_free + 1;              // 1    last phase-2 MMA of a clip segment retired
  uint64_t* d2_free = d2_full + 1;              // 1    (4 drain warps) dg of the segment flushed
  uint64_t* g_ready = d2_free + 1;              // 1    image of the segment's [g; dU] is in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_ready + 1);
  float* kap_s = reinterpret_cast<float*>(smem + Cfg::OFF_KAP);
  float* Gs = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
  float* dms = Gs + 32;
  float* dAs = Gs + 64;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tpc = p.tiles_per_clip;
  const long long total = (long long)p.B * tpc;
  const int start = (int)(total * blockIdx.x / gridDim.x), end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (start >= end) return;
  const int n = end - start;
  const uint32_t tile_u = smem_u32(smem + Cfg::OFF_TILE), g_u = smem_u32(smem + Cfg::OFF_G), w_u = smem_u32(smem + Cfg::OFF_W);
  const uint32_t c_u = smem_u32(smem + Cfg::OFF_C), l_u = smem_u32(smem + Cfg::OFF_L);

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmTok);
      for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&tile_free[s], 1); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&d1_full[s], 1); mbar_init(&d1_free[s], 2);
        mbar_init(&c_full[s], 2); mbar_init(&c_free[s], 1); mbar_init(&kap_free[s], 4);
      }
      mbar_init(d3_full, 1); mbar_init(d3_free, 4);
      mbar_init(d2_full, 1); mbar_init(d2_free, 4); mbar_init(g_ready, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  // alpha padding rows and the off-diagonal of diag(lambda) stay zero for the whole kernel
  for (int i = tid; i < (2 * Cfg::W_BYTES + 4 * Cfg::C_BYTES) / 16; i += kTbThreads) sts128(w_u + 16 * i, 0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();
  pdl_wait();

  constexpr int UF = (ST % 2 == 0) ? ST : 2 * ST;                            // stage and buffer index periodic in the tile index
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
  } else if (warp == 3) {
    // =============================================================== phase-1 issuer (warp converged, elected lane issues)
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc1 = umma_idesc_bf16(64, R, false, false);        // D1[tokens x 2 HS]: A = tile K-major, B = image K-major
    int seg1 = 0;
    auto phase1 = [&](int it, int st, int buf) {
      mbar_wait(&full[st], (it / ST) & 1);
      mbar_wait(&d1_free[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint64_t da0 = umma_desc_sw128(tile_u + st * kTTileBytes, 0, 1024);
      const uint64_t db0 = umma_desc_sw128(g_u, 0, 1024);
      const uint32_t d1 = Cfg::TM_D1 + buf * R;
#pragma unroll
      for (int box = 0; box < kTBoxes; ++box) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_lead(lead, d1, da0 + (uint64_t)(box * (kTBoxBytes >> 4) + 2 * k), db0 + (uint64_t)(box * (Cfg::G_BOX >> 4) + 2 * k),
                       idesc1, (box > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit_lead(lead, &d1_full[buf]);
    };
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
    __syncwarp();
  } else if (warp == 7) {
    // =============================================================== phase-2 / phase-3 issuer
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, HSP, true, false);      // dg[channels x HSP]: A = tile MN-major, B = alpha K-major
    constexpr uint32_t idesc3 = umma_idesc_bf16(128, kTT, true, false);      // dt^T[channels x tokens]: A = image / tile MN-major
    int seg2 = 0;
    auto phase23 = [&](int it, int st, int buf) {
      const int gt = start + it;
      const bool first = it == 0 || gt % tpc == 0, last = it == n - 1 || (gt + 1) % tpc == 0;
      mbar_wait(&c_full[buf], (it >> 1) & 1);
      if (first) mbar_wait(d2_free, (seg2 & 1) ^ 1);                         // dg of the previous segment has left tensor memory
      mbar_wait(d3_free, (it & 1) ^ 1);                                      // dt^T of the previous tile has been read out
      tc_fence_after();
      const uint64_t at0 = umma_desc_sw128(tile_u + st * kTTileBytes, kTBoxBytes, 1024);   // tile, MN-major, two boxes per M = 128
      const uint64_t ag0 = umma_desc_sw128(g_u, Cfg::G_BOX, 1024);                           // image, MN-major
      const uint64_t bw = umma_desc_sw128(w_u + buf * Cfg::W_BYTES, 0, 1024);
      const uint64_t bc = umma_desc_sw128(c_u + buf * Cfg::C_BYTES, 0, 1024);
      const uint64_t bl = umma_desc_sw128(l_u + buf * Cfg::C_BYTES, 0, 1024);
      const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll
      for (int mb = 0; mb < 6; ++mb) {
        const uint64_t at = at0 + (uint64_t)(2 * mb * (kTBoxBytes >> 4));
        const uint64_t ag = ag0 + (uint64_t)(2 * mb * (Cfg::G_BOX >> 4));
        umma_ss_lead(lead, Cfg::TM_D2 + mb * HSP, at, bw, idesc2, acc0);
        umma_ss_lead(lead, Cfg::TM_D2 + mb * HSP, at + 128, bw + 2, idesc2, 1u);
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          umma_ss_lead(lead, Cfg::TM_D3 + mb * kTT, ag + (uint64_t)(128 * k), bc + (uint64_t)(2 * k), idesc3, k > 0 ? 1u : 0u);
        umma_ss_lead(lead, Cfg::TM_D3 + mb * kTT, at, bl, idesc3, 1u);
        umma_ss_lead(lead, Cfg::TM_D3 + mb * kTT, at + 128, bl + 2, idesc3, 1u);
      }
      umma_commit_lead(lead, &tile_free[st]);
      umma_commit_lead(lead, &c_free[buf]);
      umma_commit_lead(lead, d3_full);
      if (last) { umma_commit_lead(lead, d2_full); ++seg2; }
    };
    for (int base = 0; base < n; base += UF) {
#pragma unroll
      for (int j = 0; j < UF; ++j) {
        const int it = base + j;
        if (it < n) phase23(it, j % ST, j & 1);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // =============================================================== drain warps: dt of every tile, dg at the end of a clip segment
    const int q = warp & 3;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    int seg = 0;
    for (int it = 0; it < n;) {
      const int b = (start + it) / tpc;
      const int seg_n = min(n - it, tpc - (start + it) % tpc);
      for (int e = it + seg_n; it < e; ++it) {
        const int buf = it & 1;
        const int tok_base = ((start + it) % tpc) * kTT;
        mbar_wait(d3_full, it & 1);
        tc_fence_after();
        float kap[kTT];
#pragma unroll
        for (int j4 = 0; j4 < kTT / 4; ++j4) {
          const float4 k4 = *reinterpret_cast<const float4*>(kap_s + buf * kTT + 4 * j4);
          kap[4 * j4] = k4.x; kap[4 * j4 + 1] = k4.y; kap[4 * j4 + 2] = k4.z; kap[4 * j4 + 3] = k4.w;
        }
        float* dst = p.dt + ((long long)b * p.N + tok_base) * kTD + 32 * q + lane;
        const int nvalid = min(kTT, p.N - tok_base);
#pragma unroll 1
        for (int mb = 0; mb < 6; ++mb) {
          uint32_t u[kTT];
          tmem_ld_32x32b_x32(Cfg::TM_D3 + lane_sel + mb * kTT, u);
          tmem_ld_wait();
          if (mb == 5) {                                  // everything of this tile has been read: D3 and kappa may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(d3_free); mbar_arrive(&kap_free[buf]); }
          }
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < kTT; ++j)
              if (j < nvalid) red_add_f32(dst + (long long)j * kTD + 128 * mb, __uint_as_float(u[j]) + kap[j]);
          } else {
#pragma unroll
            for (int j = 0; j < kTT; ++j)
              if (j < nvalid) dst[(long long)j * kTD + 128 * mb] = __uint_as_float(u[j]) + kap[j];
          }
        }
      }
      // ---- dg of the segment
      mbar_wait(d2_full, seg & 1);
      tc_fence_after();
      {
        float* dst = p.dg + (long long)b * HS * kTD + 32 * q + lane;
#pragma unroll 1
        for (int mb = 0; mb < 6; ++mb) {
          uint32_t u[HSP];
          tmem_ld_cols<HSP>(Cfg::TM_D2 + lane_sel + mb * HSP, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HS; ++i) red_add_f32(dst + i * kTD + 128 * mb, __uint_as_float(u[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_free);
      ++seg;
    }
  } else {
    // =============================================================== coefficient warps (0, 1: even tiles; 4, 5: odd tiles) and the
    // slot warps (4-7) that build the image of [g; dU].  D1 (M = 64) keeps token 16 q + i in lane i < 16 of lane quarter q;
    // lanes 0-15 take heads 0, 1 of their token, lanes 16-31 heads 2, 3 (fetched through a shuffle).
    const int q = warp & 3;
    const bool slot_warp = warp >= 4, cf_warp = q < 2;                       // slot warps: 4, 5, 6 (warp 7 issues MMAs)
    const int pair = warp >> 2;                                              // = buffer index of this warp's tiles
    const int tc = tid - 128;
    const int tk = 16 * q + (lane & 15), hl = lane >> 4;
    const int row0 = hl * 2 * S;                                             // this lane's (head, slot) rows: row0 .. row0 + 2 S - 1
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    float accG[2 * S], accC[2 * S];
    int seg = 0;
    for (int it = 0; it < n;) {
      const int b = (start + it) / tpc;
      const int seg_n = min(n - it, tpc - (start + it) % tpc);
      if (slot_warp) {
        // ---- bf16 image of [g[b]; dU[b]]: [12 boxes][R rows][128 B], 16-byte chunks swizzled by (row & 7)
        if (seg > 0) mbar_wait(d2_full, (seg - 1) & 1);                      // every MMA reading the previous clip's image has retired
        for (int i = tc; i < R * 96; i += 96) {
          const int row = i / 96, c8 = i - row * 96;
          const float* base = row < HS ? p.g + ((long long)b * HS + row) * kTD : p.dU + ((long long)b * HS + row - HS) * kTD;
          const float4* src = reinterpret_cast<const float4*>(base + 8 * c8);
          const float4 x = __ldg(src), y = __ldg(src + 1);
          sts128(g_u + (c8 >> 3) * Cfg::G_BOX + row * 128 + (((c8 & 7) ^ (row & 7)) << 4), pack_bf16(x.x, x.y), pack_bf16(x.z, x.w),
                 pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        }
        if (tc < HS) {
          Gs[tc] = __ldg(p.G + b * HS + tc); dms[tc] = __ldg(p.dm + b * HS + tc); dAs[tc] = __ldg(p.dA + b * HS + tc);
        }
        fence_proxy_async();
        named_bar_sync(1, 96);
        if (tc == 0) mbar_arrive(g_ready);
      } else {
        mbar_wait(g_ready, seg & 1);
      }
      if (cf_warp) {
        float Gr[2 * S], dmr[2 * S], dAr[2 * S];
#pragma unroll
        for (int i = 0; i < 2 * S; ++i) {
          accG[i] = 0.f; accC[i] = 0.f;
          Gr[i] = Gs[row0 + i]; dmr[i] = dms[row0 + i]; dAr[i] = dAs[row0 + i];
        }
        for (int e_it = it + seg_n, i2 = it; i2 < e_it; ++i2) {
          if ((i2 & 1) != pair) continue;
          const int u_ = i2 >> 1;                                            // use count of this pair's buffers
          const int tok = ((start + i2) % tpc) * kTT + tk;
          const bool valid = tok < p.N;
          // saved forward values of this lane's (token, heads): issued before the wait on the dots
          float av[2 * S], da[2 * S];
          const long long arow = ((long long)b * HS + row0) * p.N + tok;
#pragma unroll
          for (int i = 0; i < 2 * S; ++i) {
            av[i] = valid ? __ldg(p.a + arow + (long long)i * p.N) : 0.f;
            da[i] = (valid && p.dattn != nullptr) ? __ldg(p.dattn + arow + (long long)i * p.N) : 0.f;
          }
          const float mu = valid ? __ldg(p.mu + (long long)b * p.N + tok) : 0.f;
          const float r = valid ? __ldg(p.rstd + (long long)b * p.N + tok) : 0.f;
          mbar_wait(&d1_full[pair], u_ & 1);
          tc_fence_after();
          uint32_t d[R];
          if constexpr (R == 64) {
            uint32_t (&dlo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&d[0]);
            uint32_t (&dhi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&d[32]);
            tmem_ld_32x32b_x32(Cfg::TM_D1 + lane_sel + pair * R, dlo);
            tmem_ld_32x32b_x32(Cfg::TM_D1 + lane_sel + pair * R + 32, dhi);
          } else {
            tmem_ld_cols<R>(Cfg::TM_D1 + lane_sel + pair * R, d);
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&d1_free[pair]);
          float ev[2 * S], fv[2 * S];
#pragma unroll
          for (int i = 0; i < 2 * S; ++i) {
            const uint32_t oe = __shfl_sync(0xffffffffu, d[2 * S + i], lane & 15);
            const uint32_t of = __shfl_sync(0xffffffffu, d[HS + 2 * S + i], lane & 15);
            ev[i] = __uint_as_float(hl == 0 ? d[i] : oe) - mu * Gr[i];
            fv[i] = fmaf(mu, dmr[i], __uint_as_float(hl == 0 ? d[HS + i] : of));
            da[i] += fmaf(r, fv[i], dAr[i]);
          }
          float dr = 0.f, dmu = 0.f;
          float alpha[2 * S], beta[2 * S];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float dot = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) dot = fmaf(av[h * S + s], da[h * S + s], dot);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int i = h * S + s;
              const float dsim = av[i] * (da[i] - dot);
              dr += dsim * ev[i] + av[i] * fv[i];
              dmu += av[i] * dmr[i] - dsim * Gr[i];
              alpha[i] = r * dsim;
              beta[i] = r * av[i];
              accG[i] = fmaf(-alpha[i], mu, accG[i]);
              accC[i] += dsim;
            }
          }
          dr += __shfl_xor_sync(0xffffffffu, dr, 16);
          dmu += __shfl_xor_sync(0xffffffffu, dmu, 16);
          dmu *= r;
          const float lambda = -dr * r * r * r * (1.0f / kTD);
          const float kappa = dmu * (1.0f / kTD) - lambda * mu;
          mbar_wait(&c_free[pair], (u_ & 1) ^ 1);
          mbar_wait(&kap_free[pair], (u_ & 1) ^ 1);
          const uint32_t wt = w_u + pair * Cfg::W_BYTES + (tk & 7) * 2;
          const uint32_t ct = c_u + pair * Cfg::C_BYTES + tk * 128;
#pragma unroll
          for (int i = 0; i < 2 * S; ++i) {
            const int sh = row0 + i;
            const __nv_bfloat16 ab = __float2bfloat16_rn(alpha[i]);
            sts16(wt + sh * 128 + (((tk >> 3) ^ (sh & 7)) << 4), *reinterpret_cast<const uint16_t*>(&ab));
          }
#pragma unroll
          for (int i = 0; i < 2 * S; i += 2) {                               // coefficient row of the token: alpha | beta, pairs of bf16
            const int ka = row0 + i, kb = HS + row0 + i;
            sts32u(ct + (((ka >> 3) ^ (tk & 7)) << 4) + (ka & 7) * 2, pack_bf16(alpha[i], alpha[i + 1]));
            sts32u(ct + (((kb >> 3) ^ (tk & 7)) << 4) + (kb & 7) * 2, pack_bf16(beta[i], beta[i + 1]));
          }
          if (hl == 0) {
            const __nv_bfloat16 lb = __float2bfloat16_rn(lambda);
            sts16(l_u + pair * Cfg::C_BYTES + tk * 128 + (((tk >> 3) ^ (tk & 7)) << 4) + (tk & 7) * 2,
                  *reinterpret_cast<const uint16_t*>(&lb));
            kap_s[pair * kTT + tk] = kappa;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_full[pair]);
        }
        // ---- dG / dc0 of the clip segment: sums over the 16 tokens of each half warp
#pragma unroll
        for (int i = 0; i < 2 * S; ++i) {
          float x = accG[i], y = accC[i];
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
          }
          if ((lane & 15) == 0) { atomicAdd(p.dG + b * HS + row0 + i, x); atomicAdd(p.dc0 + b * HS + row0 + i, y); }
        }
      }
      it += seg_n;
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(0u);
}

template <int HS>
static int launch_slot_tc_bwd(const CUtensorMap& tm, const SlotTcBwdParams& p, cudaStream_t s) {
  using Cfg = SlotTcBwdCfg<HS>;
  static_assert(Cfg::BYTES <= 227 * 1024, "tensor-core slot backward does not fit in shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    DV_CHECK_CUDA(cudaFuncSetAttribute(slot_stream_tc_bwd_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    attr_done = true;
  }
  const long long total = (long long)p.B * p.tiles_per_clip;
  long long grid = sm_count();
  if (grid > total) grid = total;
  const double bytes = (double)p.B * p.N * kTD * (2.0 + (p.accumulate ? 8.0 : 4.0));
  const int prof = prof_begin(DEVIAS_PROF_SLOT, bytes, s);
  DV_CHECK_CUDA(launch_k(slot_stream_tc_bwd_kernel<HS>, dim3((unsigned)grid), dim3((unsigned)kTbThreads), (size_t)Cfg::BYTES, s, tm, p));
  prof_end(prof, s);
  DV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DEVIAS_OK;
}

}  // namespace dv

extern "C" int devias_slot_stream_bwd_bf16(const void* tokens, const float* mu, const float* rstd, const float* g, const float* G,
                                           const float* attn, const float* dU, const float* dm, const float* dA,
                                           const float* dattn, float* dtokens, int accumulate_dtokens, float* dg, float* dG,
                                           float* dc0, int batch, int n_tokens, int dim, int num_slots, void* stream) {
  using namespace dv;
  DV_REQUIRE(tokens && mu && rstd && g && G && attn && dU && dm && dA && dtokens && dg && dG && dc0, "null pointer");
  DV_REQUIRE(dim == kTD, "token dim must be 768");
  DV_REQUIRE(num_slots == 2 || num_slots == 4 || num_slots == 8, "num_slots must be 2, 4 or 8");
  DV_REQUIRE(batch > 0 && n_tokens > 0, "empty problem");
  DV_REQUIRE(reinterpret_cast<uintptr_t>(tokens) % 16 == 0, "tokens must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm;
  int rc = make_token_tmap_bf16(&tm, tokens, batch, n_tokens);
  if (rc) return rc;
  SlotTcBwdParams p{batch, n_tokens, (n_tokens + kTT - 1) / kTT, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens,
                    accumulate_dtokens, dg, dG, dc0};
  switch (num_slots) {
    case 2: return launch_slot_tc_bwd<8>(tm, p, s);
    case 4: return launch_slot_tc_bwd<16>(tm, p, s);
    default: return launch_slot_tc_bwd<32>(tm, p, s);
  }
}
