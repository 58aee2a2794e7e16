// hb_env_cta.cuh -- CTA-cooperative pieces built on hb_env.cuh: Philox randomness for a new episode and the
// strided observation writer.  Device-only (included by the .cu files of libhanabi_b200.so).
#pragma once
#include "hb_engine.h"

// ---------------------------------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al. 2011).  Replaces the per-env std::mt19937 of the reference
// (hanabi_game.cc:43-53): stream = (seed, game), counter = (episode, draw block, purpose).
__device__ __forceinline__ uint4 hb_philox(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

enum { HB_RNG_EPISODE = 0, HB_RNG_ACT = 1, HB_RNG_TEST = 2, HB_RNG_SAMPLE = 3 };

struct HbRng {  // sequential 32-bit draws out of one Philox stream
  uint2 key;
  uint4 ctr;
  uint4 buf;
  int have;
  __device__ HbRng(uint64_t seed, uint32_t stream, uint32_t a, uint32_t purpose) {
    key = make_uint2((uint32_t)seed ^ (stream * 0x9E3779B1u), (uint32_t)(seed >> 32) + stream);
    ctr = make_uint4(0u, a, purpose, stream);
    have = 0;
  }
  __device__ uint32_t next() {
    if (have == 0) { buf = hb_philox(ctr, key); ++ctr.x; have = 4; }
    const uint32_t v = have == 4 ? buf.x : (have == 3 ? buf.y : (have == 2 ? buf.z : buf.w));
    --have;
    return v;
  }
  __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }  // [0,n)
  __device__ float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }           // [0,1)
};

// Randomness of a new episode (HanabiEnv::reset, hanabi_env.cc:12-44): deck order, eps index per player,
// colour permutations.  The reference draws cards one at a time with probability proportional to the
// remaining counts (hanabi_state.cc:285-289, 316-328), which is a uniformly random order of the 50 cards --
// a Fisher-Yates shuffle is distribution-identical.  One thread.
__device__ inline void hb_new_episode_random(HbGame& s, uint8_t* deck, const HbEnvCfg& cfg, uint64_t seed, int game) {
  HbRng rng(seed, (uint32_t)game, s.episode, HB_RNG_EPISODE);
  int n = 0;
  for (int c = 0; c < HB_NC; ++c)
    for (int r = 0; r < HB_NR; ++r)
      for (int k = hb_card_mult(r); k > 0; --k) deck[n++] = (uint8_t)(c * HB_NR + r);
  for (int i = HB_DECK - 1; i > 0; --i) {
    const int j = (int)rng.below((uint32_t)i + 1);
    const uint8_t t = deck[i]; deck[i] = deck[j]; deck[j] = t;
  }
  for (int p = 0; p < cfg.g.P; ++p) s.eps_idx[p] = (uint8_t)(cfg.n_eps > 1 ? rng.below((uint32_t)cfg.n_eps) : 0);
  const int fix = cfg.shuffle_color ? (int)rng.below((uint32_t)cfg.g.P) : -1;
  for (int p = 0; p < HB_MAX_P; ++p) {
    int pm[HB_NC] = {0, 1, 2, 3, 4};
    if (cfg.shuffle_color && p < cfg.g.P && p != fix) {
      for (int i = HB_NC - 1; i > 0; --i) {
        const int j = (int)rng.below((uint32_t)i + 1);
        const int t = pm[i]; pm[i] = pm[j]; pm[j] = t;
      }
    }
    uint16_t fw = 0, inv = 0;
    for (int c = 0; c < HB_NC; ++c) { fw |= (uint16_t)(pm[c] << (3 * c)); inv |= (uint16_t)(c << (3 * pm[c])); }
    s.perm[p] = fw; s.inv_perm[p] = inv;
  }
}

// Start a new episode on `s` (one thread): injected randomness if the host fixed it, Philox otherwise.
__device__ inline void hb_begin_episode(HbGame& s, uint8_t* deck, HbInject* inj, const HbEnvCfg& cfg, uint64_t seed, int game) {
  if (inj != nullptr && inj->flag) {
    for (int i = 0; i < HB_DECK; ++i) deck[i] = inj->deck[i];
    for (int p = 0; p < HB_MAX_P; ++p) { s.eps_idx[p] = inj->eps_idx[p]; s.perm[p] = inj->perm[p]; s.inv_perm[p] = inj->inv_perm[p]; }
    inj->flag = 0;
  } else {
    hb_new_episode_random(s, deck, cfg, seed, game);
  }
  hb_reset_game(s, cfg.g, deck);
}

// Per-game encoder tables (needs >= 32 threads; caller syncs before and after).
__device__ __forceinline__ void hb_cta_build_tables(const HbGame& s, HbEncTables& t, const HbGeom& g) {
  const int tid = threadIdx.x;
  if (tid < HB_NCARD) t.pub_count[tid] = (uint8_t)hb_pub_count(s, tid);
}
__device__ __forceinline__ void hb_cta_build_totals(const HbGame& s, HbEncTables& t, const HbGeom& g) {
  const int tid = threadIdx.x;
  if (tid < HB_MAX_P * HB_MAX_H) {
    const int p = tid / HB_MAX_H, i = tid % HB_MAX_H;
    t.belief_total[p][i] = (p < g.P && i < s.hand_len[p]) ? hb_belief_total(s, t, p, i) : 0.f;
  }
}

// All threads of the CTA write the obs dict of one game (hanabi_env.cc:115-205) with coalesced stores: observers
// [p0, p0 + np) into buffers whose first row is observer p0 (np = P for the actors and VDN batches, 1 for an IQL entry).
__device__ __forceinline__ void hb_cta_write_obs(const HbGame& s, const HbEncTables& t, const HbEnvCfg& cfg,
                                                 float* __restrict__ priv_s, float* __restrict__ legal,
                                                 float* __restrict__ own, float* __restrict__ eps,
                                                 const float* __restrict__ eps_list,
                                                 __nv_bfloat16* __restrict__ s_hi = nullptr, __nv_bfloat16* __restrict__ s_lo = nullptr,
                                                 int KS = 0, int p0 = 0, int np = -1) {
  const HbGeom& g = cfg.g;
  if (np < 0) np = g.P;
  const int nt = blockDim.x, tid = threadIdx.x;
  const int PF = np * g.F;
  for (int i = tid; i < PF; i += nt) {
    const int o = i / g.F, f = i - o * g.F;
    const float v = hb_feature(s, t, cfg, p0 + o, f);
    if (priv_s != nullptr) priv_s[i] = v;   // null: only the GEMM operand is wanted (fused rollout, see hb_refresh_obs)
    if (s_hi != nullptr) {  // GEMM operand: everything outside the belief block is 0/1, i.e. exact in bf16 (lo stays 0)
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      s_hi[o * KS + f] = hi;
      if (f >= g.off_belief && f < g.off_sad) s_lo[o * KS + f] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
  const int PA = np * g.A;
  for (int i = tid; i < PA; i += nt) legal[i] = hb_legal_elem(s, cfg, p0 + i / g.A, i % g.A);
  const int PO = np * 3 * g.H;
  if (own != nullptr)
    for (int i = tid; i < PO; i += nt) own[i] = hb_own_hand_elem(s, p0 + i / (3 * g.H), i % (3 * g.H));
  if (tid < np) eps[tid] = eps_list[s.eps_idx[p0 + tid]];
}

// Zero the recurrent state of one game's agents (rows g*P .. g*P+P-1, every layer) -- all threads of the CTA.
__device__ __forceinline__ void hb_cta_zero_hidden(const HbHidPtrs& hp, int g, int P) {
  if (hp.c == nullptr) return;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const int per_row_c = HB_HID * 4 / 16, per_row_h = HB_HID * 2 / 16;  // uint4 stores per row
  for (int l = 0; l < HB_LAYERS; ++l) {
    for (int p = 0; p < P; ++p) {
      const size_t row = (size_t)l * hp.rows_pad + (size_t)g * P + p;
      uint4* c4 = reinterpret_cast<uint4*>(hp.c + row * HB_HID);
      uint4* hh = reinterpret_cast<uint4*>(hp.h_hi + row * HB_HID);
      uint4* hl = reinterpret_cast<uint4*>(hp.h_lo + row * HB_HID);
      for (int i = threadIdx.x; i < per_row_c; i += blockDim.x) c4[i] = z;
      for (int i = threadIdx.x; i < per_row_h; i += blockDim.x) { hh[i] = z; hl[i] = z; }
    }
  }
}
