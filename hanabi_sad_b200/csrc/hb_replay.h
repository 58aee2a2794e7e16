// hb_replay.h -- device-resident prioritized episode replay: what rela/transition_buffer.h (MultiStepBuffer,
// R2D2Buffer), rela/r2d2_actor.h postAct and rela/prioritized_replay.h keep in host deques of tensor dicts lives here
// as one ring of fixed-size episode slots in HBM.  A slot does NOT hold observations: it holds, per step, the 256-byte
// board record the observation was encoded from ("replay compression": 256 B instead of P*(F+A+3H+1)*4 =
// 7 KB per step at 2 players).  Every feature is a pure function of that record (hb_env.cuh), so the sampling kernel
// re-expands the reference's padded fp32 episode bit-exactly with the very code that produced the actors' observations.
// A game writes the record straight into the slot it claimed when its episode began; at the terminal step its CTA turns the raw rewards into n-step returns, computes the
// per-step priorities from the Q-values the policy kernel left behind, aggregates them and commits the slot.
#pragma once
#include <stdint.h>

#include "hb_types.h"

enum { HB_SLOT_FREE = 0, HB_SLOT_INFLIGHT = 1, HB_SLOT_COMMITTED = 2 };
enum { HB_CNT_HEAD = 0, HB_CNT_COMMIT = 1, HB_CNT_DROPPED = 2, HB_CNT_TICK = 3, HB_CNT_POPPED = 4, HB_CNT_STALLED = 5, HB_CNT_N = 8 };
// game_slot flag: the game's episode is finished but its commit is waiting for room in the ring (replay_block = 1:
// ConcurrentQueue::blockAppend, rela/prioritized_replay.h:44-48); the game does not start its next episode meanwhile
#define HB_SLOT_PENDING 0x40000000

struct HbRing {                 // device pointers + geometry, passed by value to kernels
  int T, P, F, A, OH;           // seq_len, players, feature size, num_action, 3*hand_size
  int NE;                       // replay entries per episode slot: 1 (vdn) or P (iql, one per player)
  int cap_slots, phys_slots;
  int block;                    // 1: reference back-pressure -- hold up to limit_slots committed episodes, evict only what sample() popped
  int limit_slots;              // int(1.25 * capacity) / NE (prioritized_replay.h:183)
  int n_step;
  float gamma, gamma_n, eta, alpha;
  int uniform_priority;
  HbGame* states;               // [phys][T]   board record behind the observation of step t (re-encoded when sampled)
  int64_t* a;                   // [phys][T][P]
  int64_t* greedy_a;            // [phys][T][P]
  float* reward;                // [phys][T]   n-step return (transition_buffer.h:83-90)
  float* bootstrap;             // [phys][T]
  int* seq_len;                 // [phys]
  float* weight;                // [phys][NE]  priority^alpha (prioritized_replay.h:192-197); 0 = not sampleable
  long long* commit_seq;        // [phys]      order of arrival (ConcurrentQueue order), -1 while in flight
  int* state;                   // [phys]      HB_SLOT_*
  int* game_slot;               // [G]         slot under construction
  float* sc_reward;             // [G][T]      raw per-step reward
  float* sc_oq;                 // [G][T][P]   Q_online(s_t, a_t)          per agent
  float* sc_tq;                 // [G][T][P]   Q_target(s_t, greedy_t)     per agent
  unsigned long long* counters; // [HB_CNT_N]
};

#define HB_SAMPLE_SETS 4   // the reference's default prefetch (3) + the batch being trained on

struct HbReplay {
  HbRing ring;
  int capacity;                 // entries
  float beta;
  uint64_t seed;
  unsigned long long sample_count;
  // sampling scratch
  double* prefix;               // [phys*NE]
  double* bsum;                 // [blocks] per-block weight sums, then (in place) their exclusive scan
  int* bcnt;                    // [blocks] per-block sampleable counts
  // Outstanding samples, oldest first (a FIFO of HB_SAMPLE_SETS id sets): one for hb_replay_sample, several when batches are
  // drawn ahead of their use (hb_replay_prefetch = the reference's `prefetch` futures, prioritized_replay.h:219-240).
  // hb_replay_update_priority always applies to the oldest.
  int* sampled_idx;             // [SETS][max_batch] entry index = slot*NE + e
  long long* sampled_seq;       // [SETS][max_batch] commit_seq at sampling time (evicted-since check, prioritized_replay.h:106-120)
  float* sampled_w;             // [SETS][max_batch]
  int* d_entry;                 // scratch of hb_replay_get
  float* d_prio;                // [max_batch] staging for update_priority
  double* d_targets;            // [max_batch] caller-supplied draw positions (hb_replay_sample_ex)
  int *d_max_len, *h_max_len;   // [SETS] longest episode of each outstanding batch (device, pinned mirror)
  cudaEvent_t set_ev[HB_SAMPLE_SETS];   // recorded behind the set's gather kernel
  int set_n[HB_SAMPLE_SETS];    // batch size of the set
  int set_waited[HB_SAMPLE_SETS];
  int set_head, set_count;
  int max_batch;
  unsigned long long* h_counters;  // pinned mirror
};
