// hb_rollout.cu -- the fused actor tick: one launch of hb_k_tick does, for every game of the engine, what one
// iteration of HanabiThreadLoop::mainLoop (cpp/thread_loop.h:42-88) spreads over VectorEnv::step, R2D2Actor::postAct,
// MultiStepBuffer / R2D2Buffer, PrioritizedReplay::add and VectorEnv::reset:
//
//   1. apply the action the policy chose last tick (HanabiEnv::step, hanabi_env.cc:49-108): reward, terminal;
//   2. record (a, greedy_a, reward, Q_online, Q_target) of that step in the episode slot of the replay ring;
//   3. on terminal: n-step returns + bootstrap flags (transition_buffer.h:51-99), per-step priorities
//      (r2d2.py:344-358), eta-aggregate (r2d2_actor.h:10-21), weight = priority^alpha, commit the slot
//      (prioritized_replay.h:192-197); zero the agents' LSTM state (r2d2_actor.h:113-126); start the next episode
//      (HanabiEnv::reset, hanabi_env.cc:9-47) in a freshly claimed slot;
//   4. encode the observation (hanabi_env.cc:115-205) ONCE into the bf16 hi/lo operand of the policy's first GEMM (plus
//      legal_move / eps for the act kernel; the fp32 priv_s / own_hand of the obs dict only on demand, hb_refresh_obs); the replay slot receives the 256-byte board record the observation is a function of
//      (hb_replay.h), at step index ep_len.
//
// The policy forward (hb_policy.cu: 3 tcgen05 GEMM launches + head/act kernel) follows on the same stream; nothing
// returns to the host between ticks.
#include "hb_engine.h"
#include "hb_env_cta.cuh"
#include "hb_policy.h"
#include "hb_replay.h"

// Threads per game (= per CTA).  2-3 players: ONE WARP -- no cross-warp barrier behind thread 0's serial sections, 32 CTAs / SM =
// one wave at 4096 games (2 players, 4096 games: 96 threads 54 us per tick, 64: 57, 32: 46).  4-5 players have 4-5 observers' rows
// and 4-6x the belief entries per game and far fewer games per GPU (1024 at BASELINE's C4): three warps (5 players, 1024 games:
// 32 threads 47 us, 64: 38, 96: 37, 128: 35; 4 players 43 -> 36 us; 3 players gain nothing).  HB_TICK_THREADS overrides both.
#ifdef HB_TICK_THREADS
#define HB_TICK_NT(P_) HB_TICK_THREADS
#else
#define HB_TICK_NT(P_) ((P_) >= 4 ? 96 : 32)
#endif
#define HB_TICK_MIN_CTAS(NT_) ((NT_) >= 96 ? 10 : ((NT_) >= 64 ? 24 : 32))   // 96 threads: 68 registers, no spills (16 CTAs / SM left 40)

struct HbTickArgs {
  HbGame* games;
  uint8_t* decks;
  HbInject* inject;
  HbEnvCfg cfg;
  uint64_t seed;
  int do_step, do_reset, has_replay;
  const int64_t* a;
  const int64_t* greedy_a;
  HbObsPtrs obs;
  const float* eps_list;
  float* reward;
  uint8_t* terminal;
  int* flags;
  HbHidPtrs hid;
  const float* oq;
  const float* tq;
  HbRing ring;
  int do_head;         // run the head / act step of the previous forward first (its output is the action applied below)
  HbHeadArgs head;
};

__device__ __forceinline__ unsigned long long hb_ld_counter(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }

// A slot for the episode that starts now: a free one, or a committed one that is no longer sampleable -- free-running ring:
// not among the `cap_slots` most recent commits (commit order, not ring order: episodes differ in length); replay_block:
// already popped by sample() (prioritized_replay.h:326-332).  Everything else is skipped; -1 only if the ring has no such
// slot within reach (counted in HB_CNT_DROPPED by the caller).
__device__ __forceinline__ int hb_claim_slot(const HbRing& R) {
  for (int tries = 0; tries < 4 * R.phys_slots; ++tries) {
    const unsigned long long id = atomicAdd(&R.counters[HB_CNT_HEAD], 1ULL);
    const int slot = (int)(id % (unsigned long long)R.phys_slots);
    const int old = atomicCAS(&R.state[slot], HB_SLOT_FREE, HB_SLOT_INFLIGHT);
    if (old == HB_SLOT_FREE) return slot;
    if (old != HB_SLOT_COMMITTED) continue;
    const long long limit = R.block ? (long long)hb_ld_counter(&R.counters[HB_CNT_POPPED])
                                    : (long long)hb_ld_counter(&R.counters[HB_CNT_COMMIT]) - R.cap_slots;
    if (*(const volatile long long*)&R.commit_seq[slot] >= limit) continue;   // still held by the replay
    if (atomicCAS(&R.state[slot], HB_SLOT_COMMITTED, HB_SLOT_INFLIGHT) == HB_SLOT_COMMITTED) {
      for (int e = 0; e < R.NE; ++e) R.weight[(size_t)slot * R.NE + e] = 0.f;
      R.commit_seq[slot] = -1;
      return slot;
    }
  }
  return -1;
}

// Turn the finished episode of game g (length Len, raw data in the sc_* scratch rows) into replay form and commit it.
// All threads of the CTA; `red` is shared scratch of 2 * (blockDim.x / 32) floats.  *committed (shared) = 0 if the ring is
// full under replay_block: nothing is published, the caller retries next tick (the computation is idempotent).
__device__ __forceinline__ void hb_cta_finalize_episode(const HbRing& R, int g, int slot, int Len, float* red, int* committed) {
  const int T = R.T, P = R.P, n = R.n_step;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const float* r = R.sc_reward + (size_t)g * T;
  for (int e = 0; e < R.NE; ++e) {
    float pmax = 0.f, psum = 0.f;
    for (int t = tid; t < T; t += nt) {
      float rew = 0.f, boot = 0.f, prio = 0.f;
      if (t < Len) {
        boot = t < Len - n ? 1.f : 0.f;                       // no terminal inside the next n steps (transition_buffer.h:66-79)
        const int K = min(n - 1, Len - 1 - t);
        float acc = 0.f;
        // Horner from the far end (transition_buffer.h:83-90); separately rounded multiply and add, as the
        // reference compiled without FMA contraction (the oracle/_ref build) computes it
        for (int s = K; s >= 0; --s) acc = __fadd_rn(r[t + s], __fmul_rn(R.gamma, acc));
        rew = acc;
        if (R.uniform_priority) {
          prio = 1.f;
        } else {
          float oq = 0.f, tq = 0.f;
          if (R.NE == 1) {                                    // VDN: Q summed over the players (r2d2.py:351-354)
            for (int p = 0; p < P; ++p) oq += R.sc_oq[((size_t)g * T + t) * P + p];
            if (boot != 0.f) for (int p = 0; p < P; ++p) tq += R.sc_tq[((size_t)g * T + t + n) * P + p];
          } else {
            oq = R.sc_oq[((size_t)g * T + t) * P + e];
            if (boot != 0.f) tq = R.sc_tq[((size_t)g * T + t + n) * P + e];
          }
          prio = fabsf(rew + boot * R.gamma_n * tq - oq);     // r2d2.py:356-357
        }
      }
      if (e == 0) { R.reward[(size_t)slot * T + t] = rew; R.bootstrap[(size_t)slot * T + t] = boot; }
      pmax = fmaxf(pmax, prio);
      psum += prio;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, k)); psum += __shfl_xor_sync(0xffffffffu, psum, k); }
    if (lane == 0) { red[warp] = pmax; red[nw + warp] = psum; }
    __syncthreads();
    if (tid == 0) {
      float m = 0.f, s = 0.f;
      for (int w = 0; w < nw; ++w) { m = fmaxf(m, red[w]); s += red[nw + w]; }
      const float agg = R.eta * m + (1.f - R.eta) * (s / (float)Len);  // aggregatePriority (r2d2_actor.h:10-21)
      R.weight[(size_t)slot * R.NE + e] = powf(agg, R.alpha);           // PrioritizedReplay::add (prioritized_replay.h:192-197)
    }
    __syncthreads();
  }
  if (tid == 0) {
    R.seq_len[slot] = Len;
    bool ok = true;
    unsigned long long seq;
    if (R.block) {   // blockAppend: size_ + 1 <= int(1.25 * capacity), else wait (prioritized_replay.h:44-48)
      seq = hb_ld_counter(&R.counters[HB_CNT_COMMIT]);
      for (;;) {
        if ((long long)(seq - hb_ld_counter(&R.counters[HB_CNT_POPPED])) >= (long long)R.limit_slots) { ok = false; break; }
        const unsigned long long prev = atomicCAS(&R.counters[HB_CNT_COMMIT], seq, seq + 1ULL);
        if (prev == seq) break;
        seq = prev;
      }
    } else {
      seq = atomicAdd(&R.counters[HB_CNT_COMMIT], 1ULL);
    }
    if (ok) {
      __threadfence();
      R.commit_seq[slot] = (long long)seq;
      __threadfence();
      atomicExch(&R.state[slot], HB_SLOT_COMMITTED);
    }
    *committed = ok ? 1 : 0;
  }
}

// TP / TH / TSAD > 0 bake the game geometry into the kernel (feature offsets, F, A become literals: the per-feature index
// arithmetic is mul-shift instead of runtime division); TP == 0 is the generic fallback for unusual configurations.
// 96 threads, 16 resident CTAs / SM (40 registers): the kernel is latency bound -- all threads wait at barriers while
// thread 0 applies the move / starts an episode -- so what helps is more CTAs in flight per SM, not more threads per CTA
// (7 CTAs of 128 threads: 82 us per 4096-game tick; 12: 71 us; with the fast encoder 58 us, as 16 x 96 threads 54 us).
template <int TP, int TH, int TSAD, int NT>
__global__ void __launch_bounds__(NT, HB_TICK_MIN_CTAS(NT)) hb_k_tick(const __grid_constant__ HbTickArgs A) {
  __shared__ HbGame s;
  __shared__ HbEncTables tab;
  __shared__ HbFastEnc enc;
  __shared__ __align__(16) uint8_t deck[HB_DECK_STRIDE];
  __shared__ int sh_did_step, sh_t, sh_term, sh_reset, sh_slot, sh_drop, sh_retry, sh_committed;
  __shared__ float red[2 * ((NT + 31) / 32)];
  __shared__ uint32_t sh_draws[4 * HB_EPISODE_DRAW_BLOCKS];
  const int g = blockIdx.x, tid = threadIdx.x;
  HbEnvCfg cfg = A.cfg;
  if (TP > 0) cfg.g = hb_make_geom(TP, TH, TSAD);
  const HbGeom& geo = cfg.g;
  const HbRing& R = A.ring;
  const int P = geo.P;
  if (tid < 16) reinterpret_cast<uint4*>(&s)[tid] = reinterpret_cast<const uint4*>(A.games + g)[tid];
  else if (tid < 20) reinterpret_cast<uint4*>(deck)[tid - 16] = reinterpret_cast<const uint4*>(A.decks + (size_t)g * HB_DECK_STRIDE)[tid - 16];
  if (A.do_head) {
    // R2D2Agent.act's tail for this game's agents (hb_head.cuh), deferred from the previous tick's forward: one launch and
    // one round trip of (a, greedy_a, Q) through a separate kernel less per tick.  The values it writes are read below by
    // other threads of this CTA: L2 loads (__ldcg) after the barrier.
    for (int p = tid >> 5; p < P; p += (NT + 31) / 32) hb_head_row(A.head, g * P + p, tid & 31);
  }
  __syncthreads();
  if (tid == 0) {
    sh_did_step = 0; sh_term = 0; sh_reset = 0; sh_drop = 0; sh_t = 0; sh_retry = 0; sh_committed = 1;
    const int raw = A.has_replay ? R.game_slot[g] : -1;
    sh_slot = raw >= 0 ? (raw & ~HB_SLOT_PENDING) : -1;
    if (raw >= 0 && (raw & HB_SLOT_PENDING)) sh_retry = 1;   // finished episode still waiting for room in the ring
    if (g == 0 && A.has_replay) atomicAdd(&R.counters[HB_CNT_TICK], 1ULL);
    if (A.do_step && !s.terminated) {
      const int t = s.ep_len;
      const int cur = s.cur_player < P ? s.cur_player : 0;
      const bool term = hb_step_game(s, cfg, deck, (int)__ldcg(A.a + g * P + cur), (int)__ldcg(A.greedy_a + g * P + cur));
      s.ep_len = (int16_t)(t + 1);
      sh_did_step = 1; sh_t = t; sh_term = term ? 1 : 0;
      A.reward[g] = s.reward;
      A.terminal[g] = term ? 1 : 0;
      if (s.illegal) { atomicAdd(&A.flags[1], 1); atomicAdd(&A.flags[3], 1); sh_drop = 1; }   // flags[3]: sticky, reported by hb_sync / hb_rollout
    } else if (A.do_step) {
      A.reward[g] = 0.f;
      A.terminal[g] = 1;
    }
  }
  __syncthreads();
  int slot = sh_slot;
  if (sh_did_step && slot >= 0 && !sh_drop) {
    const int t = sh_t;
    if (t < R.T) {
      if (tid < P) {
        const size_t o = ((size_t)slot * R.T + t) * P + tid;
        R.a[o] = __ldcg(A.a + g * P + tid);
        R.greedy_a[o] = __ldcg(A.greedy_a + g * P + tid);
        R.sc_oq[((size_t)g * R.T + t) * P + tid] = __ldcg(A.oq + g * P + tid);
        R.sc_tq[((size_t)g * R.T + t) * P + tid] = A.tq != nullptr ? __ldcg(A.tq + g * P + tid) : 0.f;
      }
      if (tid == 0) R.sc_reward[(size_t)g * R.T + t] = s.reward;
    }
    if (sh_term) {
      __syncthreads();  // the scratch rows written above are read by other threads below
      hb_cta_finalize_episode(R, g, slot, min(t + 1, R.T), red, &sh_committed);
    }
  } else if (sh_retry && slot >= 0) {
    hb_cta_finalize_episode(R, g, slot, min((int)s.ep_len, R.T), red, &sh_committed);
  }
  __syncthreads();
  // a game that restarts now needs ~60 Philox draws (deck shuffle, eps, colour permutations): one block per thread instead of
  // thirteen in a row on thread 0 (every other thread of the CTA would wait at the next barrier meanwhile)
  if (A.do_reset && s.terminated && sh_committed) {
    hb_cta_fill_draws(sh_draws, A.seed, g, s.episode);
    __syncthreads();
  }
  if (tid == 0) {
    if (sh_drop && slot >= 0) { atomicExch(&R.state[slot], HB_SLOT_FREE); atomicAdd(&R.counters[HB_CNT_DROPPED], 1ULL); }
    const bool stalled = !sh_committed;
    if (stalled) {   // the actor waits in blockAppend: no new episode, same slot, try again next tick
      R.game_slot[g] = slot | HB_SLOT_PENDING;
      atomicAdd(&R.counters[HB_CNT_STALLED], 1ULL);
    }
    if (A.do_reset && s.terminated && !stalled) {
      hb_begin_episode(s, deck, A.inject + g, cfg, A.seed, g, sh_draws);
      s.illegal = 0;   // an illegal action was counted (flags[3]) and its episode dropped; the seat carries on with a fresh game
      sh_reset = 1;
      if (A.has_replay) {
        sh_slot = hb_claim_slot(R);
        R.game_slot[g] = sh_slot;
        if (sh_slot < 0) atomicAdd(&R.counters[HB_CNT_DROPPED], 1ULL);
      }
    }
    if (s.terminated) atomicOr(&A.flags[0], 1);
  }
  __syncthreads();
  slot = sh_slot;
  if (sh_reset) hb_cta_zero_hidden(A.hid, g, P);
  hb_cta_build_tables(s, tab, geo);
  __syncthreads();
  hb_cta_build_totals(s, tab, geo);
  __syncthreads();
  const HbObsPtrs& O = A.obs;
  const int t_obs = s.ep_len;
  const bool to_ring = slot >= 0 && !s.terminated && t_obs < R.T;
  // the fp32 obs dict (priv_s, own_hand) has no reader inside the rollout: the policy consumes the bf16 operand, the replay
  // the board record.  It is materialised on demand by hb_refresh_obs when the host asks for it (hb_env_observe*).
  hb_cta_write_obs(s, tab, cfg, nullptr, O.legal_move + (size_t)g * P * geo.A, nullptr, O.eps + (size_t)g * P, A.eps_list);
  if (O.s_hi != nullptr) hb_cta_write_operand_fast(s, tab, cfg, enc, O.s_hi + (size_t)g * P * O.KS, O.s_lo + (size_t)g * P * O.KS, O.KS);
  constexpr int RT0 = NT >= 64 ? 32 : 0;   // the threads that copy the record into the ring (a warp of its own if there is one)
  if (tid >= RT0 && tid < RT0 + 16 && to_ring)   // the replay keeps the record, not the observation (re-encoded by hb_k_replay_gather)
    reinterpret_cast<uint4*>(R.states + (size_t)slot * R.T + t_obs)[tid - RT0] = reinterpret_cast<const uint4*>(&s)[tid - RT0];
  if (tid < 16) reinterpret_cast<uint4*>(A.games + g)[tid] = reinterpret_cast<const uint4*>(&s)[tid];
  else if (tid < 20) reinterpret_cast<uint4*>(A.decks + (size_t)g * HB_DECK_STRIDE)[tid - 16] = reinterpret_cast<const uint4*>(deck)[tid - 16];
}

HbRing hb_replay_ring(hb_engine* e);  // hb_replay.cu

int hb_launch_tick(hb_engine* e, int do_step, int do_reset, int clear_flags) {
  HbTickArgs a;
  memset(&a, 0, sizeof(a));
  a.games = e->d_games; a.decks = e->d_decks; a.inject = e->d_inject; a.cfg = e->env; a.seed = e->cfg.seed;
  a.do_step = do_step; a.do_reset = do_reset; a.has_replay = e->replay != nullptr;
  a.a = e->d_a; a.greedy_a = e->d_greedy_a; a.obs = e->obs; a.eps_list = e->d_eps_list; a.reward = e->d_reward; a.terminal = e->d_terminal;
  a.flags = e->d_flags; a.hid = hb_policy_hidden_ptrs(e);
  a.oq = e->policy ? e->policy->oq : nullptr;
  a.tq = e->policy && e->policy->have_weights[1] && e->cfg.priority_mode != 1 ? e->policy->tq : nullptr;
  if (e->replay) a.ring = hb_replay_ring(e);
  if (e->policy && e->policy->head_pending) {   // the previous forward left its head / act step to this launch
    a.do_head = 1;
    a.head = e->policy->pending_head;
    e->policy->head_pending = 0;
  }
  // flags[0..1] describe the LAST tick of a call (hb_env_any_terminated, the per-launch illegal count): cleared before the first
  // and before the last tick only, not by a memset between every two launches of the loop
  if (clear_flags) HB_CUDA(cudaMemsetAsync(e->d_flags, 0, 2 * sizeof(int), e->stream));
  {
    HbProfScope ps(e, HB_PROF_TICK);
    const int key = e->P * 100 + e->H * 10 + (e->env.g.sad ? 1 : 0);
#define HB_TICK_CASE(P_, H_, S_) case P_ * 100 + H_ * 10 + S_: hb_k_tick<P_, H_, S_, HB_TICK_NT(P_)><<<e->G, HB_TICK_NT(P_), 0, e->stream>>>(a); break
    switch (key) {
      HB_TICK_CASE(2, 5, 1); HB_TICK_CASE(2, 5, 0); HB_TICK_CASE(3, 5, 1); HB_TICK_CASE(3, 5, 0);
      HB_TICK_CASE(4, 4, 1); HB_TICK_CASE(4, 4, 0); HB_TICK_CASE(5, 4, 1); HB_TICK_CASE(5, 4, 0);
      default: hb_k_tick<0, 0, 0, HB_TICK_NT(2)><<<e->G, HB_TICK_NT(2), 0, e->stream>>>(a); break;
    }
#undef HB_TICK_CASE
  }
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}

extern "C" {

// n iterations of HanabiThreadLoop::mainLoop (cpp/thread_loop.h:42-88) for every game at once, entirely on the device.
int hb_rollout(hb_engine* e, int n_ticks) {
  if (!e) { hb_set_error("hb_rollout: null engine"); return -1; }
  if (!e->policy || !e->policy->have_weights[0]) { hb_set_error("hb_rollout: no policy weights (hb_policy_set_weights)"); return -1; }
  if (e->replay && e->cfg.priority_mode != 1 && !e->policy->have_weights[1]) {
    // compute_priority needs Q_target (r2d2.py:345-348); without it every priority would silently use a zero bootstrap
    hb_set_error("hb_rollout: the replay computes priorities from the target network (priority_mode %d) but net 1 has no weights", e->cfg.priority_mode);
    return -1;
  }
  HB_CUDA(cudaSetDevice(e->device));
  { const int src = hb_status_poll(e, false); if (src) return src; }   // a guard that fired during an EARLIER call
  for (int i = 0; i < n_ticks; ++i) {
    int rc = hb_launch_tick(e, e->pending_actions, 1, (i == 0 || i + 1 == n_ticks) ? 1 : 0);
    if (rc) return rc;
    // all but the last forward of this call leave their head / act step to the next tick's prologue; in profiling mode every
    // kernel class keeps its own launch so that the per-class timings stay comparable
    rc = hb_policy_forward(e, 0, (i + 1 < n_ticks && !e->prof_on) ? 1 : 0);
    if (rc) return rc;
    e->pending_actions = 1;
    e->obs_stale = 1;
    e->num_act += e->G;
    if (e->prof_on) {  // profiling mode: one sync per tick, accumulate the five launch durations
      HB_CUDA(cudaStreamSynchronize(e->stream));
      for (int k = 0; k < HB_PROF_N; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->prof_ev[2 * k], e->prof_ev[2 * k + 1]) == cudaSuccess) { e->prof_ms[k] += ms; e->prof_n[k] += 1; }
      }
    }
  }
  return hb_status_post(e);
}

// Device time per kernel class of the fused tick, measured with CUDA events on the engine stream while `on`:
// index 0 tick (env+replay+encode), 1 fc GEMM, 2 LSTM layer-0 GEMM, 3 LSTM layer-1 GEMM, 4 head/act.  Turning it on
// clears the accumulators and makes hb_rollout synchronise after every tick (do not time throughput in this mode).
int hb_profile(hb_engine* e, int on, double* ms_sum, int64_t* launches) {
  if (!e) { hb_set_error("hb_profile: null engine"); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  if (ms_sum) for (int k = 0; k < HB_PROF_N; ++k) ms_sum[k] = e->prof_ms[k];
  if (launches) for (int k = 0; k < HB_PROF_N; ++k) launches[k] = e->prof_n[k];
  if (on && !e->prof_on) {
    for (int k = 0; k < 2 * HB_PROF_N; ++k) if (!e->prof_ev[k]) HB_CUDA(cudaEventCreate(&e->prof_ev[k]));
    for (int k = 0; k < HB_PROF_N; ++k) { e->prof_ms[k] = 0; e->prof_n[k] = 0; }
  }
  e->prof_on = on ? 1 : 0;
  return 0;
}

}  // extern "C"
