// hb_env_kernels.cu -- stand-alone environment kernels (VectorEnv::reset / VectorEnv::step semantics,
// rela/env.h:48-87 over cpp/hanabi_env.cc:9-205).  One CTA per game: thread 0 advances the 256-byte board
// record held in shared memory, then all threads evaluate the P*F observation features as pure functions of
// that record and store them coalesced.  The fused rollout tick (hb_rollout.cu) reuses the same device code.
#include "hb_env_cta.cuh"

#define HB_ENV_THREADS 128

__global__ void __launch_bounds__(HB_ENV_THREADS)
hb_k_env(HbGame* __restrict__ games, uint8_t* __restrict__ decks, HbInject* __restrict__ inject, HbEnvCfg cfg,
         uint64_t seed, int do_reset, int do_step, const int64_t* __restrict__ a, const int64_t* __restrict__ greedy_a,
         HbObsPtrs obs, const float* __restrict__ eps_list, float* __restrict__ reward, uint8_t* __restrict__ terminal,
         int* __restrict__ flags, HbHidPtrs hid) {
  __shared__ HbGame s;
  __shared__ int did_reset;
  __shared__ HbEncTables tab;
  __shared__ __align__(16) uint8_t deck[HB_DECK_STRIDE];
  const int g = blockIdx.x, tid = threadIdx.x;
  const HbGeom& geo = cfg.g;
  if (tid < 16) reinterpret_cast<uint4*>(&s)[tid] = reinterpret_cast<const uint4*>(games + g)[tid];
  else if (tid < 20) reinterpret_cast<uint4*>(deck)[tid - 16] = reinterpret_cast<const uint4*>(decks + (size_t)g * HB_DECK_STRIDE)[tid - 16];
  __syncthreads();
  if (tid == 0) {
    did_reset = 0;
    if (do_step) {
      if (!s.terminated) {
        const int cur = s.cur_player < geo.P ? s.cur_player : 0;
        const int au = (int)a[g * geo.P + cur];
        const int gu = greedy_a != nullptr ? (int)greedy_a[g * geo.P + cur] : au;
        const bool t = hb_step_game(s, cfg, deck, au, gu);
        reward[g] = s.reward;
        terminal[g] = t ? 1 : 0;
        if (s.illegal) atomicAdd(&flags[1], 1);
      } else {  // the reference asserts !terminated() (hanabi_env.cc:51); keep the game frozen and report it
        reward[g] = 0.f;
        terminal[g] = 1;
      }
    }
    if (do_reset && s.terminated) { hb_begin_episode(s, deck, inject + g, cfg, seed, g); did_reset = 1; }
    if (s.terminated) atomicOr(&flags[0], 1);
    else atomicAdd(&flags[4], 1);   // games still being played (hb_eval_rollout's stop condition)
  }
  __syncthreads();
  if (did_reset) hb_cta_zero_hidden(hid, g, geo.P);
  hb_cta_build_tables(s, tab, geo);
  __syncthreads();
  hb_cta_build_totals(s, tab, geo);
  __syncthreads();
  hb_cta_write_obs(s, tab, cfg, obs.priv_s + (size_t)g * geo.P * geo.F, obs.legal_move + (size_t)g * geo.P * geo.A,
                   obs.own_hand + (size_t)g * geo.P * 3 * geo.H, obs.eps + (size_t)g * geo.P, eps_list,
                   obs.s_hi ? obs.s_hi + (size_t)g * geo.P * obs.KS : nullptr, obs.s_lo ? obs.s_lo + (size_t)g * geo.P * obs.KS : nullptr, obs.KS);
  if (tid < 16) reinterpret_cast<uint4*>(games + g)[tid] = reinterpret_cast<const uint4*>(&s)[tid];
  else if (tid < 20) reinterpret_cast<uint4*>(decks + (size_t)g * HB_DECK_STRIDE)[tid - 16] = reinterpret_cast<const uint4*>(deck)[tid - 16];
}

// Encode-only pass: the fp32 obs dict of every game from its current board record (same device code as hb_k_env / the
// fused tick, so the values are the ones those kernels would have written).
__global__ void __launch_bounds__(HB_ENV_THREADS) hb_k_encode(const HbGame* __restrict__ games, HbEnvCfg cfg, HbObsPtrs obs, const float* __restrict__ eps_list) {
  __shared__ HbGame s;
  __shared__ HbEncTables tab;
  const int g = blockIdx.x, tid = threadIdx.x;
  const HbGeom& geo = cfg.g;
  if (tid < 16) reinterpret_cast<uint4*>(&s)[tid] = reinterpret_cast<const uint4*>(games + g)[tid];
  __syncthreads();
  hb_cta_build_tables(s, tab, geo);
  __syncthreads();
  hb_cta_build_totals(s, tab, geo);
  __syncthreads();
  hb_cta_write_obs(s, tab, cfg, obs.priv_s + (size_t)g * geo.P * geo.F, obs.legal_move + (size_t)g * geo.P * geo.A,
                   obs.own_hand + (size_t)g * geo.P * 3 * geo.H, obs.eps + (size_t)g * geo.P, eps_list);
}

int hb_refresh_obs(hb_engine* e) {
  if (!e->obs_stale) return 0;
  hb_k_encode<<<e->G, HB_ENV_THREADS, 0, e->stream>>>(e->d_games, e->env, e->obs, e->d_eps_list);
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  e->obs_stale = 0;
  return 0;
}

// Uniform-random legal action per agent (the role of `legal_move.multinomial(1)` in r2d2.py:273), one thread
// per (game, player).  greedy_a gets an independent draw so the SAD block is exercised.
__global__ void hb_k_random_actions(const float* __restrict__ legal, int rows, int A, uint64_t seed, uint64_t counter,
                                    int64_t* __restrict__ a, int64_t* __restrict__ greedy_a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  HbRng rng(seed, (uint32_t)r, (uint32_t)counter, HB_RNG_TEST);
  const float* lm = legal + (size_t)r * A;
  int n = 0;
  for (int u = 0; u < A; ++u) n += lm[u] != 0.f;
  int k1 = (int)rng.below((uint32_t)n), k2 = (int)rng.below((uint32_t)n);
  int a1 = A - 1, a2 = A - 1;
  for (int u = 0, seen = 0; u < A; ++u) {
    if (lm[u] != 0.f) {
      if (seen == k1) a1 = u;
      if (seen == k2) a2 = u;
      ++seen;
    }
  }
  a[r] = a1;
  greedy_a[r] = a2;
}

// Audit kernel: one thread per game (see hb_env_check_invariants in include/hanabi_b200.h).
__global__ void hb_k_check_invariants(const HbGame* __restrict__ games, const uint8_t* __restrict__ decks, int G, HbGeom geo,
                                      int* __restrict__ flags) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const HbGame& s = games[g];
  if (s.episode == 0) return;  // never reset: nothing to audit
  int cnt[HB_NCARD];
  for (int i = 0; i < HB_NCARD; ++i) cnt[i] = s.discard_count[i];
  bool bad = false;
  for (int c = 0; c < HB_NC; ++c) {
    bad |= s.fireworks[c] > HB_NR;
    for (int r = 0; r < s.fireworks[c] && r < HB_NR; ++r) ++cnt[c * HB_NR + r];
  }
  for (int p = 0; p < geo.P; ++p) {
    bad |= s.hand_len[p] > geo.H;
    for (int i = 0; i < HB_MAX_H; ++i) {
      const int card = s.hand_card[p][i];
      if (i < s.hand_len[p]) { if (card < HB_NCARD) ++cnt[card]; else bad = true; }
      else bad |= card != HB_NO_CARD;
    }
  }
  bad |= s.deck_pos > HB_DECK;
  for (int i = s.deck_pos; i < HB_DECK; ++i) {
    const int card = decks[(size_t)g * HB_DECK_STRIDE + i];
    if (card < HB_NCARD) ++cnt[card]; else bad = true;
  }
  for (int i = 0; i < HB_NCARD; ++i) bad |= cnt[i] != hb_card_mult(i % HB_NR);
  bad |= s.info > HB_MAX_INFO || s.life > HB_MAX_LIFE || s.turns_to_play > geo.P || s.illegal != 0;
  if (!s.terminated) bad |= s.cur_player >= geo.P || s.next_player >= geo.P;
  if (bad) atomicAdd(&flags[2], 1);
}

int hb_launch_check_invariants(hb_engine* e) {
  HB_CUDA(cudaMemsetAsync(e->d_flags + 2, 0, sizeof(int), e->stream));
  hb_k_check_invariants<<<(e->G + 127) / 128, 128, 0, e->stream>>>(e->d_games, e->d_decks, e->G, e->env.g, e->d_flags);
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}

int hb_launch_env(hb_engine* e, int do_reset, int do_step, const int64_t* a_dev, const int64_t* greedy_a_dev) {
  HB_CUDA(cudaMemsetAsync(e->d_flags, 0, 2 * sizeof(int), e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_flags + 4, 0, sizeof(int), e->stream));
  hb_k_env<<<e->G, HB_ENV_THREADS, 0, e->stream>>>(e->d_games, e->d_decks, e->d_inject, e->env, e->cfg.seed, do_reset, do_step,
                                                   a_dev, greedy_a_dev, e->obs, e->d_eps_list, e->d_reward, e->d_terminal,
                                                   e->d_flags, hb_policy_hidden_ptrs(e));
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  e->obs_stale = 0;   // hb_k_env writes the whole obs dict
  return 0;
}

int hb_launch_random_actions(hb_engine* e, uint64_t counter) {
  const int rows = e->rows;
  hb_k_random_actions<<<(rows + 127) / 128, 128, 0, e->stream>>>(e->obs.legal_move, rows, e->A, e->cfg.seed, counter, e->d_a,
                                                                 e->d_greedy_a);
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}
