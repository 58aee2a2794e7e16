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

// The same stream read from a table filled in parallel (hb_cta_fill_draws): draw k = word k % 4 of Philox block k / 4.
#define HB_EPISODE_DRAW_BLOCKS 20   // 49 (shuffle) + P (eps) + 1 + 4 (P-1) (colour permutations) <= 71 draws -> 18 blocks
struct HbRngTable {
  const uint32_t* v;
  int i;
  __device__ HbRngTable(const uint32_t* table) : v(table), i(0) {}
  __device__ uint32_t next() { return v[i++]; }
  __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};
// All threads of the CTA: the first 4 * HB_EPISODE_DRAW_BLOCKS draws of (seed, game, episode, HB_RNG_EPISODE), one Philox block
// per thread, into shared memory -- the 13+ blocks a reset consumes no longer run back to back in one thread.
__device__ __forceinline__ void hb_cta_fill_draws(uint32_t* table, uint64_t seed, int game, uint32_t episode) {
  for (int b = threadIdx.x; b < HB_EPISODE_DRAW_BLOCKS; b += blockDim.x) {
    const uint32_t stream = (uint32_t)game;
    const uint2 key = make_uint2((uint32_t)seed ^ (stream * 0x9E3779B1u), (uint32_t)(seed >> 32) + stream);
    const uint4 r = hb_philox(make_uint4((uint32_t)b, episode, HB_RNG_EPISODE, stream), key);
    table[4 * b] = r.x; table[4 * b + 1] = r.y; table[4 * b + 2] = r.z; table[4 * b + 3] = r.w;
  }
}

// Randomness of a new episode (HanabiEnv::reset, hanabi_env.cc:12-44): deck order, eps index per player,
// colour permutations.  The reference draws cards one at a time with probability proportional to the
// remaining counts (hanabi_state.cc:285-289, 316-328), which is a uniformly random order of the 50 cards --
// a Fisher-Yates shuffle is distribution-identical.  One thread.
template <class RNG>
__device__ inline void hb_new_episode_draw(HbGame& s, uint8_t* deck, const HbEnvCfg& cfg, RNG& rng) {
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
__device__ inline void hb_new_episode_random(HbGame& s, uint8_t* deck, const HbEnvCfg& cfg, uint64_t seed, int game) {
  HbRng rng(seed, (uint32_t)game, s.episode, HB_RNG_EPISODE);
  hb_new_episode_draw(s, deck, cfg, rng);
}

// Start a new episode on `s` (one thread): injected randomness if the host fixed it, Philox otherwise -- drawn here, or read
// from `draws` (hb_cta_fill_draws for this game's CURRENT s.episode) when the caller prepared them in parallel.
__device__ inline void hb_begin_episode(HbGame& s, uint8_t* deck, HbInject* inj, const HbEnvCfg& cfg, uint64_t seed, int game,
                                        const uint32_t* draws = nullptr) {
  if (inj != nullptr && inj->flag) {
    for (int i = 0; i < HB_DECK; ++i) deck[i] = inj->deck[i];
    for (int p = 0; p < HB_MAX_P; ++p) { s.eps_idx[p] = inj->eps_idx[p]; s.perm[p] = inj->perm[p]; s.inv_perm[p] = inj->inv_perm[p]; }
    inj->flag = 0;
  } else if (draws != nullptr) {
    HbRngTable rng(draws);
    hb_new_episode_draw(s, deck, cfg, rng);
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
  const int PF = (priv_s != nullptr || s_hi != nullptr) ? np * g.F : 0;
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

// ---------------------------------------------------------------------------------------- fast operand encoder
// The fused tick only needs priv_s as the bf16 hi/lo operand of the fc GEMM.  Evaluating every feature with hb_feature
// costs ~118 warp instructions per 32 features (section cascade, index decomposition, an IEEE division per belief entry,
// P-fold recomputation of observer-independent values).  This writer produces the SAME values (asserted against
// hb_feature on the device in tests/test_rollout_parity.py::test_fast_operand_equals_feature_encoder) differently:
//   1. every 0/1 feature of an observer's row is a bit of a shared-memory mask that starts as zero and receives the few
//      ones by scatter (one-hot cards, thermometers as bit ranges, the <= 9 ones of a last-action block, hint one-hots);
//   2. the belief fractions plausible * count / total (canonical_encoders.cc:553-563) are computed ONCE per game and card
//      slot in real colour space, already split into bf16 hi/lo -- observers only differ by rotation and colour permutation;
//   3. one coalesced pass writes the rows: a mask bit -> 0x3F80 / 0, a belief position -> the table entry.
#define HB_MASK_WORDS 46          // ceil(max F / 32): 5 players, hand 4, sad -> F = 1439
#define HB_BEL_ENTRIES (HB_MAX_P * HB_MAX_H * HB_NCARD)

struct HbFastEnc {
  uint32_t mask[HB_MAX_P][HB_MASK_WORDS];
  uint32_t bel[HB_BEL_ENTRIES];   // (lo << 16) | hi for (player, slot, real card type)
};

__device__ __forceinline__ void hb_mask_set(uint32_t* row, int bit) { atomicOr(&row[bit >> 5], 1u << (bit & 31)); }
__device__ __forceinline__ void hb_mask_range(uint32_t* row, int bit, int n) {  // bits [bit, bit + n), n <= 64
  while (n > 0) {
    const int w = bit >> 5, b = bit & 31, take = min(n, 32 - b);
    atomicOr(&row[w], (take == 32 ? 0xFFFFFFFFu : ((1u << take) - 1u)) << b);
    bit += take; n -= take;
  }
}
// the ones of one last-action block (canonical_encoders.cc:293-422) for `observer`, block starting at bit `base`
__device__ __forceinline__ void hb_mask_last_action(uint32_t* row, int base, const HbLastMove& lm, const HbGeom& g, int observer,
                                                    uint16_t perm, bool shuffle) {
  if (!lm.valid) return;
  const int P = g.P, H = g.H;
  const int rel = (lm.player - observer + P) % P;
  const bool hint = lm.type == HB_MV_REVEAL_COLOR || lm.type == HB_MV_REVEAL_RANK;
  const bool card_move = lm.type == HB_MV_PLAY || lm.type == HB_MV_DISCARD;
  int o = base;
  hb_mask_set(row, o + rel); o += P;
  if (lm.type >= 1 && lm.type <= 4) hb_mask_set(row, o + lm.type - 1);
  o += 4;
  if (hint) hb_mask_set(row, o + (rel + lm.target_offset) % P);
  o += P;
  if (lm.type == HB_MV_REVEAL_COLOR) hb_mask_set(row, o + (shuffle ? hb_perm_get(perm, lm.color) : lm.color));
  o += HB_NC;
  if (lm.type == HB_MV_REVEAL_RANK && lm.rank < HB_NR) hb_mask_set(row, o + lm.rank);
  o += HB_NR;
  if (hint) for (int j = 0; j < H; ++j) if ((lm.reveal_mask >> j) & 1) hb_mask_set(row, o + j);
  o += H;
  if (card_move && lm.card_index < H) hb_mask_set(row, o + lm.card_index);
  o += H;
  if (card_move) {
    const int shown = shuffle ? hb_perm_get(perm, lm.card_color) : lm.card_color;
    const int k = shown * HB_NR + lm.card_rank;
    if (k >= 0 && k < HB_NCARD) hb_mask_set(row, o + k);
  }
  o += HB_NCARD;
  if (lm.type == HB_MV_PLAY) { if (lm.scored) hb_mask_set(row, o); if (lm.info_token) hb_mask_set(row, o + 1); }
}

// All threads of the CTA; `t` must hold pub_count and belief_total (hb_cta_build_tables / _totals + sync).  Writes the rows
// of all P observers: hi for every feature, lo for the belief block (elsewhere lo stays 0 from allocation, as before).
__device__ __forceinline__ void hb_cta_write_operand_fast(const HbGame& s, const HbEncTables& t, const HbEnvCfg& cfg, HbFastEnc& E,
                                                          __nv_bfloat16* __restrict__ s_hi, __nv_bfloat16* __restrict__ s_lo, int KS) {
  const HbGeom& g = cfg.g;
  const int P = g.P, H = g.H, nt = blockDim.x, tid = threadIdx.x;
  const bool shuffle = cfg.shuffle_color != 0;
  for (int i = tid; i < P * HB_MASK_WORDS; i += nt) E.mask[i / HB_MASK_WORDS][i % HB_MASK_WORDS] = 0u;
  // belief fractions per (player, slot, real card type): a warp takes a card slot, its lanes the 25 card types (no per-element
  // index arithmetic: the flat loop spent a fifth of the kernel's instructions on divisions by 25, 5 and H)
  const int wlane = tid & 31, wid = tid >> 5, nwarp = (nt + 31) >> 5;
  const int lane_c = wlane / HB_NR, lane_r = wlane - lane_c * HB_NR;
  for (int ps = wid; ps < P * H; ps += nwarp) {
    if (wlane >= HB_NCARD) continue;
    const int p = ps / H, slot = ps - p * H;
    float v = 0.f;
    if (slot < s.hand_len[p]) {
      const unsigned kn = s.know[p][slot];
      if (((kn >> lane_c) & 1u) && ((kn >> (5 + lane_r)) & 1u)) {
        const float total = t.belief_total[p][slot];
        v = total > 0.f ? HB_FDIV((float)t.pub_count[wlane], total) : 0.f;
      }
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    E.bel[ps * HB_NCARD + wlane] = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
  }
  __syncthreads();
  // ---- scatter the ones (atomicOr on shared memory).  One KIND of task per warp, so that no warp walks through the other
  // kinds' branches: warp 0 card slots (hands block + belief hint one-hots), warp 1 discard thermometers, warp 2 the
  // "hand short" bits and the board block, warp 3 the last-action and SAD blocks.
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  if (warp == 0 % nw) {
    for (int i = lane; i < P * P * H; i += 32) {
      const int o = i / (P * H), k = i - o * (P * H), rel = k / H, slot = k - rel * H, p = (o + rel) % P;
      if (slot >= s.hand_len[p]) continue;
      uint32_t* row = E.mask[o];
      const uint16_t perm = s.perm[o];
      if (rel != 0) {  // own cards are hidden (canonical_encoders.cc:88-95)
        const int card = s.hand_card[p][slot], c = card / HB_NR, r = card - c * HB_NR;
        hb_mask_set(row, k * HB_NCARD + (shuffle ? hb_perm_get(perm, c) : c) * HB_NR + r);
      }
      const unsigned kn = s.know[p][slot];
      const int hc = (kn >> 10) & 7, hr = (kn >> 13) & 7;
      const int b0 = g.off_belief + k * 35 + HB_NCARD;
      if (hc != 7) hb_mask_set(row, b0 + (shuffle ? hb_perm_get(perm, hc) : hc));
      if (hr != 7) hb_mask_set(row, b0 + HB_NC + hr);
    }
  }
  if (warp == 1 % nw) {
    for (int i = lane; i < P * HB_NCARD; i += 32) {  // discard thermometer of shown colour sc, rank r: widths 3,2,2,2,1 (:252-280)
      const int o = i / HB_NCARD, k = i - o * HB_NCARD, sc = k / HB_NR, r = k - sc * HB_NR;
      const int real_c = shuffle ? hb_perm_get(s.inv_perm[o], sc) : sc;
      const int n = min((int)s.discard_count[real_c * HB_NR + r], hb_card_mult(r));
      hb_mask_range(E.mask[o], g.off_discard + sc * 10 + (r == 0 ? 0 : 2 * r + 1), n);
    }
  }
  if (warp == 2 % nw) {
    for (int i = lane; i < P * P; i += 32) {
      const int o = i / P, k = i - o * P;
      if (s.hand_len[(o + k) % P] < H) hb_mask_set(E.mask[o], P * H * HB_NCARD + k);
    }
    for (int i = lane; i < P * HB_NC; i += 32) {
      const int o = i / HB_NC, sc = i - o * HB_NC;
      const int fw = s.fireworks[shuffle ? hb_perm_get(s.inv_perm[o], sc) : sc];
      if (fw > 0) hb_mask_set(E.mask[o], g.off_board + g.deck_bits + sc * HB_NR + fw - 1);
    }
    for (int i = lane; i < 3 * P; i += 32) {
      const int o = i / 3, k = i - o * 3;
      const int bit = k == 0 ? g.off_board : (k == 1 ? g.off_board + g.deck_bits + HB_NCARD : g.off_board + g.deck_bits + HB_NCARD + HB_MAX_INFO);
      const int n = k == 0 ? min(g.deck_bits, HB_DECK - (int)s.deck_pos) : (k == 1 ? min((int)s.info, HB_MAX_INFO) : min((int)s.life, HB_MAX_LIFE));
      hb_mask_range(E.mask[o], bit, n);
    }
  }
  if (warp == 3 % nw) {
    if (lane < 2 * P && (lane < P || g.sad)) {
      const int o = lane < P ? lane : lane - P;
      const bool sad_block = lane >= P;
      hb_mask_last_action(E.mask[o], sad_block ? g.off_sad : g.off_last, (sad_block && s.greedy_valid) ? s.greedy : s.last, g, o, s.perm[o], shuffle);
    }
  }
  __syncthreads();
  // ---- rows out.  Pass 1: every 32-bit mask word becomes 32 bf16 (0x3F80 / 0) = four 16-byte stores; KS is a multiple of
  // 64, so a row is exactly KS/32 words (bits at and beyond F are zero).  Pass 2 (after the barrier, which orders the two
  // writes of the same CTA to the same addresses) overwrites the belief fractions and writes their lo halves.
  const int words = KS >> 5;
  for (int i = tid; i < P * words; i += nt) {
    const int o = i / words, w = i - o * words;
    const uint32_t m = w < HB_MASK_WORDS ? E.mask[o][w] : 0u;
    // two 32-byte stores (st.global.v8.b32, sm_100): each lane writes whole sectors -- with 16-byte stores at this 64-byte
    // lane stride every request touched 32 half sectors and the kernel moved twice the sectors it needed
    __nv_bfloat16* dst = s_hi + (size_t)o * KS + w * 32;
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t b2 = (m >> (16 * q2 + 2 * j)) & 3u;
        v[j] = ((b2 & 1u) ? 0x3F80u : 0u) | ((b2 & 2u) ? 0x3F800000u : 0u);
      }
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "l"(dst + 16 * q2), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
  }
  __syncthreads();
  for (int j = wid; j < P * P * H; j += nwarp) {   // j = observer * (P*H) + rs, rs = rel * H + slot; lanes = the 25 shown card types
    const int o = j / (P * H), rs = j - o * (P * H), rel = rs / H, slot = rs - rel * H;
    if (wlane >= HB_NCARD) continue;
    int p = o + rel;
    if (p >= P) p -= P;
    const int real_c = shuffle ? hb_perm_get(s.inv_perm[o], lane_c) : lane_c;
    const uint32_t e = E.bel[(p * H + slot) * HB_NCARD + real_c * HB_NR + lane_r];
    const size_t at = (size_t)o * KS + g.off_belief + rs * 35 + wlane;
    s_hi[at] = __ushort_as_bfloat16((unsigned short)(e & 0xFFFFu));
    s_lo[at] = __ushort_as_bfloat16((unsigned short)(e >> 16));
  }
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
