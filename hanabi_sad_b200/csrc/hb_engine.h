// hb_engine.h -- internal (C++) definition of hb_engine shared by the translation units of libhanabi_b200.so.
// Not part of the ABI; the ABI is include/hanabi_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/hanabi_b200.h"
#include "hb_env.cuh"

#define HB_DECK_STRIDE 64  // bytes per game in the deck / injected-deck arrays (50 used)

// Randomness of the next episode fixed from the host (hb_env_inject); consumed by one reset.
struct HbInject {
  uint8_t flag;
  uint8_t eps_idx[HB_MAX_P];
  uint16_t perm[HB_MAX_P], inv_perm[HB_MAX_P];
  uint8_t deck[HB_DECK];
  uint8_t pad[HB_DECK_STRIDE * 2 - 1 - HB_MAX_P - 4 * HB_MAX_P - HB_DECK];
};
static_assert(sizeof(HbInject) == 128, "HbInject must be 128 bytes");

// Device pointers of the "current observation" in the reference's obs-dict layout (hanabi_env.cc:195-204).
struct HbObsPtrs {
  float* priv_s;     // [G,P,F]
  float* legal_move; // [G,P,A]
  float* own_hand;   // [G,P,3H]
  float* eps;        // [G,P]
  // the same priv_s as the policy's first GEMM operand: bf16 hi / lo split, [rows_pad][KS] (null without a policy)
  __nv_bfloat16* s_hi;
  __nv_bfloat16* s_lo;
  int KS;
};

// The CURRENT half of the policy's recurrent state, so that the kernel that restarts a game can zero its agents'
// rows (R2D2Actor::postAct, rela/r2d2_actor.h:113-126).  All null without a policy.
struct HbHidPtrs {
  __nv_bfloat16* h_hi;  // [L][rows_pad][512]
  __nv_bfloat16* h_lo;
  float* c;             // [L][rows_pad][512]
  int rows_pad;
};

enum { HB_PROF_TICK = 0, HB_PROF_FC = 1, HB_PROF_LSTM0 = 2, HB_PROF_LSTM1 = 3, HB_PROF_HEAD = 4, HB_PROF_N = 5 };

struct hb_engine {
  hb_config cfg;
  HbEnvCfg env;
  int G, P, H, F, A, rows;
  int device;
  int sm_count;
  cudaStream_t stream;
  int64_t launches;
  // per-kernel-class device timing (hb_profile): CUDA events around every launch of a tick, off by default
  int prof_on;
  cudaEvent_t prof_ev[2 * HB_PROF_N];
  double prof_ms[HB_PROF_N];
  int64_t prof_n[HB_PROF_N];
  int pending_actions;   // d_a / d_greedy_a hold a reply the environment has not consumed yet
  int obs_stale;         // obs.priv_s / obs.own_hand are older than the board records (the fused tick does not write them)
  int64_t num_act;       // sum of R2D2Actor::numAct_ (r2d2_actor.h:98): env-steps acted on

  // ---- environment
  HbGame* d_games;       // [G]
  uint8_t* d_decks;      // [G][HB_DECK_STRIDE]
  HbInject* d_inject;    // [G]
  float* d_eps_list;     // [num_eps]
  HbObsPtrs obs;
  float* d_reward;       // [G]
  uint8_t* d_terminal;   // [G]
  int64_t* d_a;          // [G,P]
  int64_t* d_greedy_a;   // [G,P]
  int* d_flags;          // [4]: 0 any_terminated, 1 illegal count of the last launch, 2 invariant violations, 3 illegal count (sticky)
  int* h_flags;          // pinned mirror
  // device-side guards (policy GEMM spin guard, illegal action inside hb_rollout): mirrored to pinned memory at the end of
  // every hb_rollout, examined by the next call / hb_sync / hb_counters
  int* h_status;         // pinned [2]: 0 policy d_error, 1 sticky illegal count
  cudaEvent_t ev_status;
  int status_pending;

  // ---- policy / replay (hb_policy.cu, hb_replay.cu) -- opaque here
  struct HbPolicy* policy;
  struct HbReplay* replay;
};

void hb_set_error(const char* fmt, ...);

// Brackets one kernel launch with events when profiling is on (see hb_profile in include/hanabi_b200.h).
struct HbProfScope {
  hb_engine* e;
  int k;
  HbProfScope(hb_engine* e_, int k_) : e(e_), k(k_) { if (e->prof_on) cudaEventRecord(e->prof_ev[2 * k], e->stream); }
  ~HbProfScope() { if (e->prof_on) cudaEventRecord(e->prof_ev[2 * k + 1], e->stream); }
};

#define HB_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (call);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      hb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e));        \
      return -2;                                                                                        \
    }                                                                                                   \
  } while (0)

// hb_api.cu: queue the mirror copy (hb_status_post), examine it (hb_status_poll: wait = block until the copy has landed;
// returns -4 with the message set if a guard fired, clearing it)
int hb_status_post(hb_engine* e);
int hb_status_poll(hb_engine* e, bool wait);
HbHidPtrs hb_policy_hidden_ptrs(hb_engine* e);  // hb_policy.cu
int hb_launch_tick(hb_engine* e, int do_step, int do_reset, int clear_flags);  // hb_rollout.cu
// hb_env_kernels.cu
int hb_launch_env(hb_engine* e, int do_reset, int do_step, const int64_t* a_dev, const int64_t* greedy_a_dev);
int hb_launch_random_actions(hb_engine* e, uint64_t counter);
int hb_launch_check_invariants(hb_engine* e);
int hb_refresh_obs(hb_engine* e);  // re-encode obs.priv_s / obs.own_hand from the board records if the fused tick left them stale
