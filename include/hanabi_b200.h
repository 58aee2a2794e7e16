/*
 * hanabi_b200.h -- C ABI of libhanabi_b200.so, the B200 (sm_100a) Hanabi actor hot path.
 *
 * One hb_engine owns, on ONE GPU: `num_games` board states, the per-agent LSTM hidden state, the policy
 * weights (online + target), the per-game episode under construction and the prioritized episode replay.
 * It replaces, as one unit, what the reference spreads over
 *     cpp/hanabi_env.{h,cc}  + hanabi-learning-environment/hanabi_lib/       (simulator + encoder)
 *     rela/env.h VectorEnv, cpp/thread_loop.h HanabiThreadLoop                (tick loop)
 *     rela/batcher.h, rela/batch_runner.h + TorchScript R2D2Agent.act         (policy forward)
 *     rela/r2d2_actor.h, rela/transition_buffer.h                             (n-step, episode assembly, priority)
 *     rela/prioritized_replay.h                                               (replay add / sample / update)
 * (paths relative to the reference root).  Every entry point below names the reference interface it replaces.
 *
 * Conventions: plain C types only; every function returns 0 on success and a negative code on failure
 * (hb_last_error() gives the message); no exceptions cross this boundary; the caller owns every buffer it
 * passes, the engine owns all device memory.  "host" pointers are ordinary (ideally pinned) host memory,
 * "dev" pointers are device memory on the engine's GPU.  All calls on one engine must come from one thread
 * at a time (same rule as the reference's sample/update alternation, rela/prioritized_replay.h:209-212).
 */
#ifndef HANABI_B200_H_
#define HANABI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_engine hb_engine;

/* Everything the reference passes to HanabiEnv (cpp/hanabi_env.h:19-48), create_envs/create_threads
 * (pyhanabi/create.py:24-76), R2D2Actor (rela/r2d2_actor.h:25-58), R2D2Agent (pyhanabi/r2d2.py:163-206) and
 * RNNPrioritizedReplay (rela/prioritized_replay.h:176-190) for one act device. */
typedef struct hb_config {
  int32_t device;          /* CUDA ordinal */
  int32_t num_games;       /* num_thread * num_game_per_thread on this device */
  int32_t players;         /* 2..5 */
  int32_t hand_size;       /* 2..5 (selfplay.py passes it explicitly, create.py:40) */
  int32_t bomb;            /* 0 | 1 | -1  (hanabi_state.cc:362-376) */
  int32_t max_len;         /* <=0: no forced termination; also the replay sequence length when > 0 */
  int32_t sad;             /* append the greedy-action block (hanabi_env.cc:154-160) */
  int32_t shuffle_color;   /* Other-Play colour permutation */
  int32_t num_eps;         /* length of eps_list */
  const float* eps_list;   /* host; copied */
  uint64_t seed;           /* production RNG (Philox) seed; game g uses stream (seed, g) */
  /* actor / replay (ignored until hb_policy_* / hb_rollout are used) */
  int32_t vdn;             /* 1: one replay entry per game (players axis kept); 0: IQL, one per player */
  int32_t multi_step;      /* n of the n-step return */
  float gamma;
  float eta;               /* priority = eta*max + (1-eta)*mean (r2d2_actor.h:10-21) */
  int32_t seq_len;         /* episode slots T in the replay (reference: max_len = 80) */
  int32_t replay_capacity; /* episodes; 0 = no replay (eval-style actors) */
  float alpha, beta;       /* prioritized_replay.h:176-190 */
  int32_t hid_dim;         /* 512 */
  int32_t num_lstm_layer;  /* 2 */
  int32_t num_fc_layer;    /* 1 (r2d2.py:42-46) */
  int32_t skip_connect;    /* r2d2.py:74-75 */
  int32_t priority_mode;   /* 0: reference semantics (online + target forward every tick, both fp32-class);
                            * 1: uniform priority (R2D2Agent uniform_priority, r2d2.py:310-311: no target forward);
                            * 2: as 0 but the target forward runs in plain bf16 (priorities are heuristics) */
  int32_t eval_seats;      /* 1: evaluation engine -- one network PER SEAT (hb_policy_set_weights net = seat index), which may
                            * differ in weights and in architecture variant (num_fc_layer 1|2, skip_connect): tools/eval_model.py
                            * cross-play (utils.load_op_model, utils.py:35-84).  No target network, no replay. */
  int32_t replay_block;    /* 0: free-running ring -- once `replay_capacity` episodes are held every new commit evicts the oldest
                            *    (actors never wait; throughput runs without a learner);
                            * 1: the reference's back-pressure (ConcurrentQueue::blockAppend, rela/prioritized_replay.h:44-48,183):
                            *    the ring holds up to int(1.25*capacity) entries, a game whose finished episode finds it full WAITS
                            *    (does not start its next episode) until hb_replay_sample pops down to `capacity` (:326-332). */
  int32_t reserved[5];
} hb_config;

/* Snapshot of one game, for tests / eval (HanabiEnv getters, cpp/pybind.cc:23-38). */
typedef struct hb_game_info {
  int32_t cur_player;      /* -1 while a deal is pending (only visible at terminal states) */
  int32_t score, life, info, deck_size, num_step, terminated, last_score, illegal;
  int32_t fireworks[5];
  int32_t hand_len[5];
  int32_t hand_card[5][5]; /* colour*5+rank, -1 = empty */
  int32_t eps_idx[5];
  int32_t perm[5][5];      /* colour permutation per observer (identity unless shuffle_color) */
  uint32_t episode;        /* episodes started on this seat */
} hb_game_info;

const char* hb_last_error(void);
int hb_version(void);

/* HanabiEnv ctor x num_games + HanabiVecEnv.append (cpp/pybind.cc:15-43, create.py:36-53,63-71). */
int hb_create(const hb_config* cfg, hb_engine** out);
void hb_destroy(hb_engine* e);

int hb_feature_size(const hb_engine* e); /* HanabiEnv::featureSize, hanabi_env.h:52-59 */
int hb_num_action(const hb_engine* e);   /* HanabiEnv::numAction,   hanabi_env.h:61-63 */
int hb_num_games(const hb_engine* e);

/* ---- environment, VectorEnv semantics (rela/env.h:48-104); host buffers ---------------------------- */

/* Parity hook: fix the randomness of the NEXT episode of `game` (deck order, per-player eps index, per-player
 * colour permutation real->shown [P][5], NULL = identity).  Replaces the mt19937 draws of HanabiEnv::reset
 * (hanabi_env.cc:12-44) and ApplyRandomChance (hanabi_state.cc:285-289).  Without it Philox supplies them. */
int hb_env_inject(hb_engine* e, int game, const int8_t* deck50, const int32_t* eps_idx, const int32_t* perms);

/* VectorEnv::reset (env.h:48-60): start a new episode on every terminated game, then encode every game. */
int hb_env_reset(hb_engine* e);

/* VectorEnv::step (env.h:66-87) -> HanabiEnv::step (hanabi_env.cc:49-113).  a / greedy_a: host int64 [G,P]
 * (only the entry of each game's current player is read); reward: host float [G]; terminal: host uint8 [G].
 * An illegal action marks the game (hb_game_info.illegal) and returns -3 where the reference aborts. */
int hb_env_step(hb_engine* e, const int64_t* a, const int64_t* greedy_a, float* reward, uint8_t* terminal);

/* The obs dict of the last reset/step (hanabi_env.cc:195-204): host float priv_s [G,P,F], legal_move [G,P,A],
 * own_hand [G,P,3H], eps [G,P].  Any pointer may be NULL. */
int hb_env_observe(hb_engine* e, float* priv_s, float* legal_move, float* own_hand, float* eps);

/* Same data without the host copy: device pointers valid until hb_destroy (layouts as above).  The fused hb_rollout
 * does not keep the fp32 priv_s / own_hand up to date (nothing on the device reads them: the policy consumes a bf16
 * operand, the replay the board record); asking for them here (or in hb_env_observe) re-encodes them from the current
 * board records first, so call this again after a rollout rather than caching the contents. */
int hb_env_observe_dev(hb_engine* e, const float** priv_s, const float** legal_move, const float** own_hand,
                       const float** eps, const float** reward, const uint8_t** terminal);

/* Device-resident variant of hb_env_step: a / greedy_a are device int64 [G,P]; results stay on the device. */
int hb_env_step_dev(hb_engine* e, const int64_t* a_dev, const int64_t* greedy_a_dev);

/* VectorEnv::anyTerminated (env.h:89-96) and the getters of cpp/pybind.cc:23-38. */
int hb_env_any_terminated(hb_engine* e, int* out);
int hb_env_query(hb_engine* e, int game, hb_game_info* out);

/* HanabiEnv::lastScore of every game (hanabi_env.h:108-110): host int32 [G], -1 before a game's first terminal. */
int hb_env_last_scores(hb_engine* e, int32_t* out);

/* The 50-card deal order of `game`'s current episode (card id colour*5+rank), whether injected or drawn from the
 * engine's Philox stream -- what ApplyRandomChance (hanabi_state.cc:285-289) would have produced one draw at a
 * time.  Together with hb_env_query's eps_idx / perm it lets a test replay the same episode on the CPU oracle. */
int hb_env_get_deck(hb_engine* e, int game, int8_t* deck50);

/* Device-side audit of every board record: card conservation (hands + discards + fireworks + undealt = the 50-card
 * multiset), token / hand-length / turn ranges.  *num_bad = number of games violating an invariant.  Plays the role
 * of the reference's live asserts (hanabi_state.cc:225, hanabi_hand.cc:85-97) for the batched state. */
int hb_env_check_invariants(hb_engine* e, int* num_bad);

/* Uniform-random legal policy on the device (test / benchmark driver, the analogue of random_action in
 * r2d2.py:273): fills the engine's own action buffers for the next hb_env_step_dev(e, NULL, NULL). */
int hb_env_random_actions(hb_engine* e, uint64_t counter);

/* Host copies of the engine's device-resident action buffers (the reply of R2D2Actor::act, r2d2_actor.h:78-96:
 * a / greedy_a int64 [G,P]) and of the last step's result (reward float [G], terminal uint8 [G]).  NULL skips. */
int hb_env_get_actions(hb_engine* e, int64_t* a, int64_t* greedy_a);
/* The other direction: overwrite the pending reply (host int64 [G,P]; greedy_a NULL = a) that the next hb_env_step_dev(e, NULL,
 * NULL) / hb_rollout tick applies -- a test hook (e.g. to drive an illegal action into the fused loop). */
int hb_env_set_actions(hb_engine* e, const int64_t* a, const int64_t* greedy_a);
int hb_env_get_result(hb_engine* e, float* reward, uint8_t* terminal);

/* ---- policy: the R2D2 act forward (pyhanabi/r2d2.py:65-78, 234-303) ---------------------------------- */

/* One network's parameters in the layouts of R2D2Net.state_dict() (fp32, row-major, host OR device pointers):
 * net.0.{weight [512,F], bias}, lstm.{weight_ih, weight_hh [2048,512], bias_ih, bias_hh [2048]}_l{0,1},
 * fc_a.{weight [A,512], bias}, fc_v.{weight [1,512], bias}. */
typedef struct hb_weights {
  const float* fc_w; const float* fc_b;
  const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  const float* fc_a_w; const float* fc_a_b;
  const float* fc_v_w; const float* fc_v_b;
  /* architecture variants of the OP-paper models (r2d2.py:42-46, 74-75); honoured by eval_seats engines only */
  const float* fc2_w; const float* fc2_b;   /* net.2.{weight [512,512], bias}: second fc layer, NULL if num_fc_layer == 1 */
  int32_t skip_connect;                      /* the head consumes lstm_out + fc_out */
} hb_weights;

/* BatchRunner::updateModel (rela/batch_runner.h:74-77) / R2D2Agent.sync_target_with_online: net 0 = online_net,
 * 1 = target_net.  Synchronous: the caller's tensors may be released when it returns. */
int hb_policy_set_weights(hb_engine* e, int net, const hb_weights* w);

/* R2D2Actor::act minus the environment (rela/r2d2_actor.h:61-100 -> R2D2Agent.act): one forward of both networks
 * over the engine's current observation; (a, greedy_a) land in the engine's action buffers, the hidden state
 * advances.  greedy_only != 0 ignores eps (eval actors). */
int hb_policy_act(hb_engine* e, int greedy_only);

/* pyhanabi/eval.py:19-66 / HanabiThreadLoop(eval = true) (cpp/thread_loop.h:74-86) as ONE call: (re)start every game, then
 * act -> step on the device until every game has finished its episode (finished games stay frozen; at most max_ticks ticks,
 * <= 0: 512), without a host round trip per tick; scores: host int32 [G] = HanabiEnv::lastScore of every game (may be NULL),
 * *ticks_run (may be NULL) = ticks queued.  For engines without a replay (evaluation actors never call postAct). */
int hb_eval_rollout(hb_engine* e, int max_ticks, int32_t* scores, int* ticks_run);

/* Host copies of the last forward's outputs (NULL skips): adv float [G*P,A] online advantages; online_q / target_q
 * float [G*P] = the dueling Q-values compute_priority uses (r2d2.py:344-348); h, c float [2, G*P, 512] = the
 * CURRENT hidden state (R2D2Actor::hidden_). */
int hb_policy_get(hb_engine* e, float* adv, float* online_q, float* target_q, float* h, float* c);

/* ---- fused actor loop + replay -------------------------------------------------------------------------- */

/* n iterations of HanabiThreadLoop::mainLoop (cpp/thread_loop.h:42-88) for every game of the engine, on the device:
 * VectorEnv::step with the pending reply, R2D2Actor::postAct (n-step return, priority, episode assembly,
 * PrioritizedReplay::add when an episode ends, hidden-state reset), VectorEnv::reset of finished games, observation
 * encoding, R2D2Actor::act (policy forward, eps-greedy).  Asynchronous: returns once the work is queued. */
int hb_rollout(hb_engine* e, int n_ticks);

/* RNNPrioritizedReplay::size / numAdd (rela/prioritized_replay.h:259-265) and the sum of R2D2Actor::numAct
 * (rela/r2d2_actor.h:57-59) in env-steps.  NULL skips.  Synchronises with the engine stream and, like hb_sync, returns -4
 * if a device-side guard fired since the last report (see hb_sync). */
int hb_counters(hb_engine* e, int64_t* size, int64_t* num_add, int64_t* num_act);

/* Everything the replay knows about itself (one synchronising read).  `size` .. `num_act` as hb_counters; `dropped` =
 * episodes that could not be recorded (no free slot: cannot happen with replay_block = 1); `stalled_ticks` = game-ticks spent
 * waiting for room (replay_block = 1; the reference's actors sit in cvSize_.wait, prioritized_replay.h:48); `popped` = entries
 * evicted so far (ConcurrentQueue::blockPop); `weight_sum` / `sampleable` = sum of priority^alpha and count over the entries
 * a sample() issued now would draw from (safeSize(&sum), :278-279). */
typedef struct hb_replay_info {
  int64_t size, num_add, num_act, dropped, stalled_ticks, popped, capacity, phys_slots, sampleable;
  double weight_sum;
} hb_replay_info;
int hb_replay_stats(hb_engine* e, hb_replay_info* out);

/* Destination of one sampled batch: DEVICE pointers owned by the caller, in the layouts RNNTransition::makeBatch
 * produces (rela/transition.cc:160-202).  With vdn: priv_s [T,B,P,F], legal_move [T,B,P,A], own_hand [T,B,P,3H],
 * eps [T,B,P], a / greedy_a int64 [T,B,P]; without (iql): the same without the P axis.  reward, bootstrap float [T,B];
 * terminal uint8 [T,B]; seq_len float [B]; weight float [B] (importance weights); ids int32 [B] or NULL. */
typedef struct hb_batch {
  float* priv_s; float* legal_move; float* own_hand; float* eps;
  int64_t* a; int64_t* greedy_a;
  float* reward; float* bootstrap; uint8_t* terminal; float* seq_len;
  float* weight; int32_t* ids;
} hb_batch;

/* PrioritizedReplay::sample (rela/prioritized_replay.h:208-240, 274-345).  Returns -3 if the replay holds fewer than
 * `batchsize` entries or the previous batch's priorities were not written back. */
int hb_replay_sample(hb_engine* e, int batchsize, const hb_batch* out);

/* hb_replay_sample with the two hooks a replay SHARDED over several engines / ranks needs to reproduce the reference's
 * importance weights (N * w_i / sum_w)^-beta / max over the UNION of the shards (prioritized_replay.h:334-339):
 *   targets      host double [batchsize] or NULL: the stratified draw positions inside this shard's cumulative weight
 *                [0, weight_sum) chosen by the caller (who splits one global stratified draw, :287-297, over the shards);
 *                NULL = draw here (Philox);
 *   total_weight / total_size   sum_w and N the importance weight is computed with; <= 0 = this shard's own;
 *   normalize    != 0: divide by the batch maximum here (single shard); 0: leave (N*w/sum)^-beta raw -- the caller divides
 *                by the maximum over all shards / ranks. */
typedef struct hb_sample_opts {
  const double* targets;
  double total_weight, total_size;
  int32_t normalize;
} hb_sample_opts;
int hb_replay_sample_ex(hb_engine* e, int batchsize, const hb_batch* out, const hb_sample_opts* opts);

/* The reference's `prefetch` (PrioritizedReplay::sample with futures_, rela/prioritized_replay.h:219-240): draw batches AHEAD of
 * their use.  hb_replay_prefetch queues draw + gather of one more batch on the engine stream and returns at once (up to 4
 * batches outstanding, the reference's default prefetch 3 + the one being trained on; -3 beyond); the outstanding batches form a
 * FIFO.  hb_replay_take waits until the OLDEST outstanding batch is complete in `out` of its prefetch call (*batchsize, may be
 * NULL, receives its size); hb_replay_update_priority and hb_replay_last_max_len refer to that oldest batch, and
 * update_priority removes it from the FIFO.  hb_replay_sample(_ex) = prefetch + take with nothing else outstanding. */
int hb_replay_prefetch(hb_engine* e, int batchsize, const hb_batch* out, const hb_sample_opts* opts /* may be NULL */);
int hb_replay_take(hb_engine* e, int* batchsize);

/* PrioritizedReplay::get (rela/prioritized_replay.h:259-261 -> ConcurrentQueue::get, :125-128; used by
 * pyhanabi/tools/action_matrix.py:90-107): the idx-th OLDEST entry still held (0 <= idx < size), written UNBATCHED into
 * `out` (device pointers; layouts of hb_batch with B = 1, i.e. priv_s [T,(P,)F], ..., reward [T], seq_len [1]).
 * out->weight is not written; out->ids (or NULL) receives the physical entry index.  Does not touch the sampling state. */
int hb_replay_get(hb_engine* e, int64_t idx, const hb_batch* out);

/* PrioritizedReplay::updatePriority (rela/prioritized_replay.h:242-257): `priority` float [n], host or device.  Host memory:
 * synchronous.  Device memory: queued on the engine stream -- the buffer must stay valid (and be complete: hb_stream_wait)
 * until that stream has consumed it.  Applies to the OLDEST outstanding batch; n = 0 forgets it (:243-246). */
int hb_replay_update_priority(hb_engine* e, const float* priority, int n);

/* Measurement hook: device time per kernel class of the fused tick (CUDA events on the engine stream).  Returns the
 * accumulated milliseconds / launch counts [5] = {tick, fc GEMM, LSTM-0 GEMM, LSTM-1 GEMM, head} gathered while
 * profiling was on, then switches profiling on/off (switching on clears the accumulators).  NULL skips. */
int hb_profile(hb_engine* e, int on, double* ms_sum, int64_t* launches);

/* Diagnostic: run the tcgen05 GEMM template alone, C[M,N] = A[M,K] B[N,K]^T + bias on host fp32 buffers
 * (M % 128 == N % 256 == K % 64 == 0; split != 0 selects the bf16x3 fp32-class mode). */
int hb_debug_gemm(int device, const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int split);

/* ---- learner side: the T-step LSTM of R2D2Net.forward and its backward ------------------------ */

/* torch.nn.LSTM(512, 512, num_layers=2) as the reference learner runs it (pyhanabi/r2d2.py:48-52 construction,
 * :99-105 call from R2D2Net.forward, :383-401 td_error: padded [T, rows, 512] sequences, EMPTY hid = zero initial
 * state) -- forward over whole sequences and the matching backward, on device buffers of the caller (fp32, e.g. torch
 * CUDA tensors).  Independent of hb_engine: one hb_lstm per learner process / GPU. */
typedef struct hb_lstm hb_lstm;

/* Parameters in nn.LSTM's own layout: weight_ih_l{k}, weight_hh_l{k} [2048, 512] (gate order i,f,g,o), bias_ih_l{k},
 * bias_hh_l{k} [2048]; DEVICE pointers. */
typedef struct hb_lstm_weights {
  const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
} hb_lstm_weights;
typedef struct hb_lstm_grads {
  float* dw_ih[2]; float* dw_hh[2]; float* db_ih[2]; float* db_hh[2];   /* overwritten, not accumulated */
} hb_lstm_grads;

/* Workspace for sequences up to max_T steps x max_rows rows (rows = batch, or batch * num_player for VDN). */
int hb_lstm_create(int device, int max_T, int max_rows, hb_lstm** out);
void hb_lstm_destroy(hb_lstm* l);

/* nn.LSTM.forward for `nets` (1 or 2) independent networks in one pass -- the reference calls online_net and target_net
 * on the same batch back to back (r2d2.py:398-401).  x[n], y[n]: device float [T, rows, 512] (input / top-layer output
 * sequence of network n), w[n] its parameters; y[n] (and dy of hb_lstm_backward) 32-byte aligned, as any cudaMalloc / torch
 * allocation is.  save != 0 keeps network 0's activations for hb_lstm_backward.
 * `stream`: cudaStream_t the work is queued on (e.g. torch's current stream).  Asynchronous: returns once everything is
 * queued; buffers must stay valid until the stream reaches that point (stream-ordered allocators do that by themselves).
 * A device-side failure (a spin guard of the persistent kernels) is reported by the NEXT hb_lstm_* call or hb_lstm_sync. */
int hb_lstm_forward(hb_lstm* l, int T, int rows, int nets, const float* const* x, const hb_lstm_weights* w, float* const* y,
                    int save, void* stream);

/* Backward of network 0 of the last saving forward: dy = dLoss/dy [T, rows, 512] -> dx = dLoss/dx [T, rows, 512] (may be
 * NULL) and the parameter gradients. */
int hb_lstm_backward(hb_lstm* l, const float* dy, float* dx, const hb_lstm_grads* g, void* stream);

/* The dense layers around the LSTM (nn.Linear(838, 512) of R2D2Net.net over all T*rows steps, pyhanabi/r2d2.py:42-46, 99,
 * and its weight gradient): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) on DEVICE fp32 row-major buffers (row strides lda, ldb,
 * ldc in elements) at fp32-class accuracy on the tensor cores (bf16x3) -- what cuBLAS runs as a SIMT sgemm when TF32 is off
 * (torch's default for matmul).  Any M, N, K >= 1 (padded internally; long-K problems with few output tiles are split over
 * K).  Asynchronous on `stream`; scratch is per device and grow-only, so calls on one device must not overlap. */
int hb_gemm_nt(int device, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
               int M, int N, int K, void* stream);

/* Waits for the last hb_lstm_forward / hb_lstm_backward and reports its device-side status. */
int hb_lstm_sync(hb_lstm* l);

int64_t hb_lstm_launches(const hb_lstm* l); /* kernels launched through this handle so far */

/* ---- learner side: one whole R2D2 update on the device ------------------------------------------------------ */

/* What pyhanabi/selfplay.py:208-244 does per iteration between replay.sample() and replay.update_priority():
 * R2D2Agent.loss (pyhanabi/r2d2.py:461-499: td_error :383-428, smooth-L1, aux task :430-459), (loss * weight).mean().backward(),
 * clip_grad_norm_, Adam.step, rela.aggregate_priority -- as kernels on the caller's stream, no host synchronisation.
 * Parameters / gradients / Adam moments are FLAT fp32 device buffers owned by the caller (hb_trainer_layout: element offsets of
 * net.0.weight, net.0.bias, lstm.{weight_ih, weight_hh, bias_ih, bias_hh}_l0, .._l1, fc_v.weight, fc_v.bias, fc_a.weight,
 * fc_a.bias, pred.weight, pred.bias in that order, then the total length; every tensor starts 16-byte aligned). */
typedef struct hb_trainer hb_trainer;
typedef struct hb_trainer_config {
  int32_t device;
  int32_t in_dim, num_action, hand_size;   /* R2D2Net(in_dim, 512, num_action, 2, hand_size) */
  int32_t num_player;                       /* player axis of one replay entry: P with vdn, 1 with iql */
  int32_t vdn, multi_step, seq_len, max_batch;   /* max_batch * num_player <= 256 rows */
  float gamma, eta;
  float lr, adam_eps, beta1, beta2;        /* torch.optim.Adam(lr, eps) with its default betas (selfplay.py:141) */
  float grad_clip;                          /* clip_grad_norm_ max_norm; <= 0: none */
  int32_t reserved[8];
} hb_trainer_config;
typedef struct hb_train_stats {
  float loss, rl_loss, aux_xent, grad_norm;   /* selfplay.py stat["loss"], ["rl_loss"], ["aux1"], ["grad_norm"] of the last update */
  int64_t num_update, launches;
} hb_train_stats;

int hb_trainer_layout(int in_dim, int num_action, int hand_size, int64_t* offsets /* [17] */);
int hb_trainer_create(const hb_trainer_config* cfg, float* online, float* target, float* grads, float* adam_m, float* adam_v, hb_trainer** out);
void hb_trainer_destroy(hb_trainer* t);
/* loss + backward for one sampled batch (DEVICE pointers in the layouts of hb_batch, batch->weight = importance weights);
 * t_eff = longest episode of the batch (steps beyond it are padding in every row and are skipped; pass seq_len to disable);
 * gradients -> the flat `grads` buffer (overwritten), aggregated priorities -> priority (device float [batchsize]). */
int hb_trainer_backward(hb_trainer* t, const hb_batch* batch, int batchsize, int t_eff, float pred_weight, float* priority, void* stream);
/* The same for ONE MICRO-BATCH of a larger batch (one pass of the LSTM kernels holds 256 rows = 256 / num_player entries; the
 * reference's learner has no such limit, e.g. 5-player VDN at batchsize 128 = 640 rows): `batch` holds `batchsize` contiguous
 * entries, the loss mean runs over `total_batch` entries, accumulate != 0 ADDS this pass's gradients and loss statistics to
 * those of the previous passes (first micro-batch: 0).  Exact: batch rows interact only through the mean. */
int hb_trainer_backward_ex(hb_trainer* t, const hb_batch* batch, int batchsize, int t_eff, float pred_weight, float* priority, int total_batch,
                           int accumulate, void* stream);
/* clip_grad_norm_ + Adam.step on the online network (a data-parallel learner all-reduces `grads` before this call). */
int hb_trainer_optim_step(hb_trainer* t, void* stream);
int hb_trainer_sync_target(hb_trainer* t, void* stream);      /* R2D2Agent.sync_target_with_online (r2d2.py:208-210) */
int hb_trainer_stats(hb_trainer* t, hb_train_stats* out);     /* waits for the last update */
/* Without waiting: statistics of the most recent update that has COMPLETED (out->num_update tells which, 0 = none yet). */
int hb_trainer_stats_nowait(hb_trainer* t, hb_train_stats* out);

/* Longest episode (steps) among the entries of the oldest outstanding batch of this engine (hb_replay_sample*: the last one;
 * complete after hb_replay_take): the t_eff of hb_trainer_backward. */
int hb_replay_last_max_len(hb_engine* e);
/* Make the engine's stream wait for everything queued so far on `stream` (and the other way round): lets a learner on its own
 * stream hand priorities to hb_replay_update_priority, or consume a sampled batch, without a host synchronisation. */
int hb_stream_wait(hb_engine* e, void* stream);
int hb_stream_wait_engine(hb_engine* e, void* stream);

/* Diagnostic: the observation as the policy consumes it -- the bf16 hi / lo operand of the first GEMM written by the fused
 * tick's encoder, as bit patterns [G*P][*ks] (host; hi or lo may be NULL; *ks = row length, F rounded up to 64). */
int hb_debug_operand(hb_engine* e, uint16_t* hi, uint16_t* lo, int* ks);

/* cudaStreamSynchronize on the engine stream, then the device-side guards: returns -4 (hb_last_error says which) if, since
 * the last report, a pipeline barrier of the policy GEMMs ran into its spin guard (the actions of that tick are garbage) or
 * an illegal action reached a game inside hb_rollout (the reference aborts there, hanabi_env.cc:63-80).  hb_rollout itself is
 * asynchronous: it reports such a failure of an EARLIER call as soon as it has seen it. */
int hb_sync(hb_engine* e);
void* hb_stream(hb_engine* e); /* cudaStream_t the engine launches on (for CUDA-event timing) */
int64_t hb_kernel_launches(const hb_engine* e); /* kernels launched by this engine so far */

#ifdef __cplusplus
}
#endif
#endif /* HANABI_B200_H_ */
