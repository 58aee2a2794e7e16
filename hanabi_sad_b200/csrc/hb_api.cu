// hb_api.cu -- the C ABI of libhanabi_b200.so (include/hanabi_b200.h): engine lifetime and the environment
// entry points.  Host code only; kernels live in hb_env_kernels.cu / hb_policy.cu / hb_replay.cu.
#include <stdarg.h>
#include <string.h>
#include <new>

#include "hb_engine.h"
#include "hb_policy.h"

static thread_local char g_err[512] = "";

void hb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int hb_policy_error_ptr(hb_engine* e, int** out);   // hb_policy.cu

int hb_status_post(hb_engine* e) {
  int* perr = nullptr;
  hb_policy_error_ptr(e, &perr);
  if (perr) HB_CUDA(cudaMemcpyAsync(e->h_status, perr, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaMemcpyAsync(e->h_status + 1, e->d_flags + 3, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaEventRecord(e->ev_status, e->stream));
  e->status_pending = 1;
  return 0;
}

int hb_status_poll(hb_engine* e, bool wait) {
  if (!e->status_pending) return 0;
  if (wait) HB_CUDA(cudaEventSynchronize(e->ev_status));
  else if (cudaEventQuery(e->ev_status) != cudaSuccess) { (void)cudaGetLastError(); return 0; }   // still in flight: examined later
  e->status_pending = 0;
  const int gemm = e->h_status[0], illegal = e->h_status[1];
  if (!gemm && !illegal) return 0;
  e->h_status[0] = e->h_status[1] = 0;
  int* perr = nullptr;
  hb_policy_error_ptr(e, &perr);
  if (perr) cudaMemsetAsync(perr, 0, sizeof(int), e->stream);
  cudaMemsetAsync(e->d_flags + 3, 0, sizeof(int), e->stream);
  if (gemm) hb_set_error("policy GEMM: a pipeline barrier timed out (spin guard) during an earlier tick -- the actions since then are not trustworthy");
  else hb_set_error("hb_rollout: %d illegal action(s) reached the environment (the reference aborts here, hanabi_env.cc:63-80); those episodes were dropped", illegal);
  return -4;
}

extern "C" {

const char* hb_last_error(void) { return g_err; }
int hb_version(void) { return 1; }

static int hb_fail(int code, const char* msg) {
  hb_set_error("%s", msg);
  return code;
}

int hb_policy_create(hb_engine* e);   // hb_policy.cu
void hb_policy_destroy(hb_engine* e);
int hb_replay_create(hb_engine* e);   // hb_replay.cu
void hb_replay_destroy(hb_engine* e);

int hb_create(const hb_config* cfg, hb_engine** out) {
  if (!cfg || !out) return hb_fail(-1, "hb_create: null argument");
  *out = nullptr;
  if (cfg->players < 2 || cfg->players > HB_MAX_P) return hb_fail(-1, "hb_create: players must be 2..5");
  if (cfg->hand_size < 2 || cfg->hand_size > HB_MAX_H) return hb_fail(-1, "hb_create: hand_size must be 2..5");
  if (cfg->num_games < 1) return hb_fail(-1, "hb_create: num_games must be >= 1");
  if (cfg->bomb < -1 || cfg->bomb > 1) return hb_fail(-1, "hb_create: bomb must be -1, 0 or 1");
  if (cfg->num_eps < 1 || cfg->num_eps > 255 || !cfg->eps_list) return hb_fail(-1, "hb_create: eps_list must hold 1..255 values");
  if (cfg->players * cfg->hand_size > HB_DECK) return hb_fail(-1, "hb_create: players*hand_size exceeds the deck");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    hb_set_error("hb_create: no CUDA device (%s) -- libhanabi_b200 has no CPU path", cudaGetErrorString(ce));
    return -2;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return hb_fail(-1, "hb_create: bad device ordinal");
  HB_CUDA(cudaSetDevice(cfg->device));
  hb_engine* e = new (std::nothrow) hb_engine();
  if (!e) return hb_fail(-2, "hb_create: out of host memory");
  memset(e, 0, sizeof(*e));
  e->cfg = *cfg;
  e->cfg.eps_list = nullptr;
  e->device = cfg->device;
  e->G = cfg->num_games; e->P = cfg->players; e->H = cfg->hand_size;
  e->env.g = hb_make_geom(e->P, e->H, cfg->sad ? 1 : 0);
  e->env.bomb = cfg->bomb; e->env.max_len = cfg->max_len; e->env.shuffle_color = cfg->shuffle_color ? 1 : 0;
  e->env.n_eps = cfg->num_eps;
  e->F = e->env.g.F; e->A = e->env.g.A; e->rows = e->G * e->P;
  cudaDeviceProp prop;
  HB_CUDA(cudaGetDeviceProperties(&prop, e->device));
  e->sm_count = prop.multiProcessorCount;
  HB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  const size_t G = (size_t)e->G, P = (size_t)e->P;
  HB_CUDA(cudaMalloc(&e->d_games, G * sizeof(HbGame)));
  HB_CUDA(cudaMalloc(&e->d_decks, G * HB_DECK_STRIDE));
  HB_CUDA(cudaMalloc(&e->d_inject, G * sizeof(HbInject)));
  HB_CUDA(cudaMalloc(&e->d_eps_list, cfg->num_eps * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->obs.priv_s, G * P * e->F * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->obs.legal_move, G * P * e->A * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->obs.own_hand, G * P * 3 * e->H * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->obs.eps, G * P * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->d_reward, G * sizeof(float)));
  HB_CUDA(cudaMalloc(&e->d_terminal, G));
  HB_CUDA(cudaMalloc(&e->d_a, G * P * sizeof(int64_t)));
  HB_CUDA(cudaMalloc(&e->d_greedy_a, G * P * sizeof(int64_t)));
  HB_CUDA(cudaMalloc(&e->d_flags, 8 * sizeof(int)));
  HB_CUDA(cudaMallocHost(&e->h_flags, 8 * sizeof(int)));
  HB_CUDA(cudaMallocHost(&e->h_status, 2 * sizeof(int)));
  e->h_status[0] = e->h_status[1] = 0;
  HB_CUDA(cudaEventCreateWithFlags(&e->ev_status, cudaEventDisableTiming));
  HB_CUDA(cudaMemsetAsync(e->d_flags, 0, 8 * sizeof(int), e->stream));
  HB_CUDA(cudaMemcpyAsync(e->d_eps_list, cfg->eps_list, cfg->num_eps * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_decks, 0, G * HB_DECK_STRIDE, e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_inject, 0, G * sizeof(HbInject), e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_reward, 0, G * sizeof(float), e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_terminal, 1, G, e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_a, 0, G * P * sizeof(int64_t), e->stream));
  HB_CUDA(cudaMemsetAsync(e->d_greedy_a, 0, G * P * sizeof(int64_t), e->stream));
  {  // every game starts "terminated" (state_ == nullptr, hanabi_env.h:80-82) with identity colour permutations
    HbGame g0;
    memset(&g0, 0, sizeof(g0));
    g0.terminated = 1;
    g0.last_score = -1;
    for (int p = 0; p < HB_MAX_P; ++p) { g0.perm[p] = hb_perm_identity(); g0.inv_perm[p] = hb_perm_identity(); }
    HbGame* tmp = new HbGame[G];
    for (size_t i = 0; i < G; ++i) tmp[i] = g0;
    cudaError_t c2 = cudaMemcpyAsync(e->d_games, tmp, G * sizeof(HbGame), cudaMemcpyHostToDevice, e->stream);
    cudaStreamSynchronize(e->stream);
    delete[] tmp;
    HB_CUDA(c2);
  }
  int rc = hb_policy_create(e);
  if (rc == 0) rc = hb_replay_create(e);
  if (rc != 0) { hb_destroy(e); return rc; }
  HB_CUDA(cudaStreamSynchronize(e->stream));
  *out = e;
  return 0;
}

void hb_destroy(hb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  hb_replay_destroy(e);
  hb_policy_destroy(e);
  cudaFree(e->d_games); cudaFree(e->d_decks); cudaFree(e->d_inject); cudaFree(e->d_eps_list);
  cudaFree(e->obs.priv_s); cudaFree(e->obs.legal_move); cudaFree(e->obs.own_hand); cudaFree(e->obs.eps);
  cudaFree(e->d_reward); cudaFree(e->d_terminal); cudaFree(e->d_a); cudaFree(e->d_greedy_a); cudaFree(e->d_flags);
  cudaFreeHost(e->h_flags);
  if (e->h_status) cudaFreeHost(e->h_status);
  if (e->ev_status) cudaEventDestroy(e->ev_status);
  for (int k = 0; k < 2 * HB_PROF_N; ++k) if (e->prof_ev[k]) cudaEventDestroy(e->prof_ev[k]);
  cudaStreamDestroy(e->stream);
  delete e;
}

int hb_feature_size(const hb_engine* e) { return e ? e->F : -1; }
int hb_num_action(const hb_engine* e) { return e ? e->A : -1; }
int hb_num_games(const hb_engine* e) { return e ? e->G : -1; }
void* hb_stream(hb_engine* e) { return e ? (void*)e->stream : nullptr; }
int64_t hb_kernel_launches(const hb_engine* e) { return e ? e->launches : -1; }

int hb_sync(hb_engine* e) {
  if (!e) return hb_fail(-1, "hb_sync: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  if (!e->policy) return 0;
  int rc = hb_status_post(e);
  if (rc) return rc;
  return hb_status_poll(e, true);
}

int hb_env_inject(hb_engine* e, int game, const int8_t* deck50, const int32_t* eps_idx, const int32_t* perms) {
  if (!e || !deck50 || !eps_idx) return hb_fail(-1, "hb_env_inject: null argument");
  if (game < 0 || game >= e->G) return hb_fail(-1, "hb_env_inject: game out of range");
  HbInject inj;
  memset(&inj, 0, sizeof(inj));
  int count[HB_NCARD] = {0};
  for (int i = 0; i < HB_DECK; ++i) {
    if (deck50[i] < 0 || deck50[i] >= HB_NCARD) return hb_fail(-1, "hb_env_inject: card id out of range");
    inj.deck[i] = (uint8_t)deck50[i];
    ++count[deck50[i]];
  }
  for (int k = 0; k < HB_NCARD; ++k)
    if (count[k] != hb_card_mult(k % HB_NR)) return hb_fail(-1, "hb_env_inject: not a permutation of the 50-card deck");
  for (int p = 0; p < HB_MAX_P; ++p) { inj.perm[p] = hb_perm_identity(); inj.inv_perm[p] = hb_perm_identity(); }
  for (int p = 0; p < e->P; ++p) {
    if (eps_idx[p] < 0 || eps_idx[p] >= e->cfg.num_eps) return hb_fail(-1, "hb_env_inject: eps index out of range");
    inj.eps_idx[p] = (uint8_t)eps_idx[p];
    if (perms && e->env.shuffle_color) {
      uint16_t fw = 0, inv = 0;
      unsigned seen = 0;
      for (int c = 0; c < HB_NC; ++c) {
        const int v = perms[p * HB_NC + c];
        if (v < 0 || v >= HB_NC) return hb_fail(-1, "hb_env_inject: bad colour permutation");
        seen |= 1u << v;
        fw |= (uint16_t)(v << (3 * c));
        inv |= (uint16_t)(c << (3 * v));
      }
      if (seen != 31u) return hb_fail(-1, "hb_env_inject: bad colour permutation");
      inj.perm[p] = fw; inj.inv_perm[p] = inv;
    }
  }
  inj.flag = 1;
  HB_CUDA(cudaSetDevice(e->device));
  HB_CUDA(cudaMemcpyAsync(e->d_inject + game, &inj, sizeof(inj), cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));  // `inj` is a stack object
  return 0;
}

static int hb_read_flags(hb_engine* e) {
  HB_CUDA(cudaMemcpyAsync(e->h_flags, e->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int hb_env_reset(hb_engine* e) {
  if (!e) return hb_fail(-1, "hb_env_reset: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  return hb_launch_env(e, 1, 0, nullptr, nullptr);
}

int hb_env_step_dev(hb_engine* e, const int64_t* a_dev, const int64_t* greedy_a_dev) {
  if (!e) return hb_fail(-1, "hb_env_step_dev: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  if (!a_dev) { a_dev = e->d_a; greedy_a_dev = e->d_greedy_a; }
  e->pending_actions = 0;
  return hb_launch_env(e, 0, 1, a_dev, greedy_a_dev);
}

int hb_env_step(hb_engine* e, const int64_t* a, const int64_t* greedy_a, float* reward, uint8_t* terminal) {
  if (!e || !a) return hb_fail(-1, "hb_env_step: null argument");
  HB_CUDA(cudaSetDevice(e->device));
  const size_t nb = (size_t)e->rows * sizeof(int64_t);
  HB_CUDA(cudaMemcpyAsync(e->d_a, a, nb, cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaMemcpyAsync(e->d_greedy_a, greedy_a ? greedy_a : a, nb, cudaMemcpyHostToDevice, e->stream));
  e->pending_actions = 0;
  int rc = hb_launch_env(e, 0, 1, e->d_a, e->d_greedy_a);
  if (rc) return rc;
  if (reward) HB_CUDA(cudaMemcpyAsync(reward, e->d_reward, (size_t)e->G * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (terminal) HB_CUDA(cudaMemcpyAsync(terminal, e->d_terminal, (size_t)e->G, cudaMemcpyDeviceToHost, e->stream));
  rc = hb_read_flags(e);
  if (rc) return rc;
  if (e->h_flags[1] > 0) {
    hb_set_error("hb_env_step: %d game(s) received an illegal action (the reference aborts here, hanabi_env.cc:63-80)", e->h_flags[1]);
    return -3;
  }
  return 0;
}

int hb_env_observe(hb_engine* e, float* priv_s, float* legal_move, float* own_hand, float* eps) {
  if (!e) return hb_fail(-1, "hb_env_observe: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  if (priv_s || own_hand) { const int rc = hb_refresh_obs(e); if (rc) return rc; }
  const size_t R = (size_t)e->rows;
  if (priv_s) HB_CUDA(cudaMemcpyAsync(priv_s, e->obs.priv_s, R * e->F * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (legal_move) HB_CUDA(cudaMemcpyAsync(legal_move, e->obs.legal_move, R * e->A * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (own_hand) HB_CUDA(cudaMemcpyAsync(own_hand, e->obs.own_hand, R * 3 * e->H * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (eps) HB_CUDA(cudaMemcpyAsync(eps, e->obs.eps, R * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int hb_env_observe_dev(hb_engine* e, const float** priv_s, const float** legal_move, const float** own_hand,
                       const float** eps, const float** reward, const uint8_t** terminal) {
  if (!e) return hb_fail(-1, "hb_env_observe_dev: null engine");
  if (priv_s || own_hand) { HB_CUDA(cudaSetDevice(e->device)); const int rc = hb_refresh_obs(e); if (rc) return rc; }
  if (priv_s) *priv_s = e->obs.priv_s;
  if (legal_move) *legal_move = e->obs.legal_move;
  if (own_hand) *own_hand = e->obs.own_hand;
  if (eps) *eps = e->obs.eps;
  if (reward) *reward = e->d_reward;
  if (terminal) *terminal = e->d_terminal;
  return 0;
}

int hb_env_any_terminated(hb_engine* e, int* out) {
  if (!e || !out) return hb_fail(-1, "hb_env_any_terminated: null argument");
  HB_CUDA(cudaSetDevice(e->device));
  int rc = hb_read_flags(e);
  if (rc) return rc;
  *out = e->h_flags[0] != 0;
  return 0;
}

int hb_env_query(hb_engine* e, int game, hb_game_info* out) {
  if (!e || !out) return hb_fail(-1, "hb_env_query: null argument");
  if (game < 0 || game >= e->G) return hb_fail(-1, "hb_env_query: game out of range");
  HB_CUDA(cudaSetDevice(e->device));
  HbGame s;
  HB_CUDA(cudaMemcpyAsync(&s, e->d_games + game, sizeof(s), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  memset(out, 0, sizeof(*out));
  out->cur_player = s.cur_player == HB_CHANCE ? -1 : s.cur_player;
  out->score = hb_score(s, e->env.bomb);
  out->life = s.life; out->info = s.info; out->deck_size = HB_DECK - s.deck_pos; out->num_step = s.num_step;
  out->terminated = s.terminated; out->last_score = s.last_score; out->illegal = s.illegal; out->episode = s.episode;
  for (int c = 0; c < HB_NC; ++c) out->fireworks[c] = s.fireworks[c];
  for (int p = 0; p < HB_MAX_P; ++p) {
    out->hand_len[p] = s.hand_len[p];
    out->eps_idx[p] = s.eps_idx[p];
    for (int i = 0; i < HB_MAX_H; ++i) out->hand_card[p][i] = s.hand_card[p][i] == HB_NO_CARD ? -1 : s.hand_card[p][i];
    for (int c = 0; c < HB_NC; ++c) out->perm[p][c] = hb_perm_get(s.perm[p], c);
  }
  return 0;
}

int hb_env_last_scores(hb_engine* e, int32_t* out) {
  if (!e || !out) return hb_fail(-1, "hb_env_last_scores: null argument");
  HB_CUDA(cudaSetDevice(e->device));
  HbGame* tmp = new (std::nothrow) HbGame[e->G];
  if (!tmp) return hb_fail(-2, "hb_env_last_scores: out of host memory");
  cudaError_t c1 = cudaMemcpyAsync(tmp, e->d_games, (size_t)e->G * sizeof(HbGame), cudaMemcpyDeviceToHost, e->stream);
  cudaError_t c2 = cudaStreamSynchronize(e->stream);
  for (int g = 0; g < e->G; ++g) out[g] = tmp[g].last_score;
  delete[] tmp;
  HB_CUDA(c1);
  HB_CUDA(c2);
  return 0;
}

int hb_env_get_deck(hb_engine* e, int game, int8_t* deck50) {
  if (!e || !deck50) return hb_fail(-1, "hb_env_get_deck: null argument");
  if (game < 0 || game >= e->G) return hb_fail(-1, "hb_env_get_deck: game out of range");
  HB_CUDA(cudaSetDevice(e->device));
  HB_CUDA(cudaMemcpyAsync(deck50, e->d_decks + (size_t)game * HB_DECK_STRIDE, HB_DECK, cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int hb_env_check_invariants(hb_engine* e, int* num_bad) {
  if (!e || !num_bad) return hb_fail(-1, "hb_env_check_invariants: null argument");
  HB_CUDA(cudaSetDevice(e->device));
  int rc = hb_launch_check_invariants(e);
  if (rc) return rc;
  HB_CUDA(cudaMemcpyAsync(e->h_flags + 2, e->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  *num_bad = e->h_flags[2];
  return 0;
}

int hb_env_get_actions(hb_engine* e, int64_t* a, int64_t* greedy_a) {
  if (!e) return hb_fail(-1, "hb_env_get_actions: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  const size_t nb = (size_t)e->rows * sizeof(int64_t);
  if (a) HB_CUDA(cudaMemcpyAsync(a, e->d_a, nb, cudaMemcpyDeviceToHost, e->stream));
  if (greedy_a) HB_CUDA(cudaMemcpyAsync(greedy_a, e->d_greedy_a, nb, cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int hb_env_set_actions(hb_engine* e, const int64_t* a, const int64_t* greedy_a) {
  if (!e || !a) return hb_fail(-1, "hb_env_set_actions: null argument");
  HB_CUDA(cudaSetDevice(e->device));
  const size_t nb = (size_t)e->rows * sizeof(int64_t);
  HB_CUDA(cudaMemcpyAsync(e->d_a, a, nb, cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaMemcpyAsync(e->d_greedy_a, greedy_a ? greedy_a : a, nb, cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  e->pending_actions = 1;
  return 0;
}

int hb_env_get_result(hb_engine* e, float* reward, uint8_t* terminal) {
  if (!e) return hb_fail(-1, "hb_env_get_result: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  if (reward) HB_CUDA(cudaMemcpyAsync(reward, e->d_reward, (size_t)e->G * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (terminal) HB_CUDA(cudaMemcpyAsync(terminal, e->d_terminal, (size_t)e->G, cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

// The evaluation loop of HanabiThreadLoop(eval = true) (cpp/thread_loop.h:74-86) / pyhanabi/eval.py:19-66 on the device: every
// game plays ONE episode (act -> step, finished games stay frozen).  Ticks are queued in chunks; after each chunk the number
// of games still running travels to pinned memory asynchronously and is looked at one chunk LATER, so the stream never drains
// and the host never waits inside the loop (a finished evaluation costs at most one surplus chunk of no-op ticks).
int hb_eval_rollout(hb_engine* e, int max_ticks, int32_t* scores, int* ticks_run) {
  if (!e) return hb_fail(-1, "hb_eval_rollout: null engine");
  if (!e->policy || !e->policy->have_weights[0]) return hb_fail(-1, "hb_eval_rollout: no policy weights (hb_policy_set_weights)");
  if (e->replay) return hb_fail(-1, "hb_eval_rollout: evaluation engines have no replay (actors with eval = true never call postAct)");
  if (max_ticks <= 0) max_ticks = 512;
  HB_CUDA(cudaSetDevice(e->device));
  int rc = hb_launch_env(e, 1, 0, nullptr, nullptr);   // VectorEnv::reset: start every game that is not running
  if (rc) return rc;
  const int CH = 8;
  cudaEvent_t ev[2];
  HB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  HB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  int ticks = 0, chunk = 0;
  bool done = false;
  while (!done && ticks < max_ticks) {
    for (int i = 0; i < CH && ticks < max_ticks; ++i, ++ticks) {
      rc = hb_policy_forward(e, 0);
      if (!rc) rc = hb_launch_env(e, 0, 1, e->d_a, e->d_greedy_a);
      if (rc) { cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); return rc; }
    }
    cudaMemcpyAsync(e->h_flags + 4 + (chunk & 1), e->d_flags + 4, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    cudaEventRecord(ev[chunk & 1], e->stream);
    if (chunk > 0) {   // the PREVIOUS chunk's count: by now almost certainly there
      cudaEventSynchronize(ev[(chunk - 1) & 1]);
      if (e->h_flags[4 + ((chunk - 1) & 1)] == 0) done = true;
    }
    ++chunk;
  }
  cudaEventSynchronize(ev[(chunk - 1) & 1]);
  const int live = e->h_flags[4 + ((chunk - 1) & 1)];
  cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
  e->pending_actions = 0;
  if (ticks_run) *ticks_run = ticks;
  if (live != 0) { hb_set_error("hb_eval_rollout: %d game(s) still running after %d ticks", live, ticks); return -3; }
  rc = hb_status_post(e);
  if (!rc) rc = hb_status_poll(e, true);
  if (rc) return rc;
  return scores ? hb_env_last_scores(e, scores) : 0;
}

int hb_env_random_actions(hb_engine* e, uint64_t counter) {
  if (!e) return hb_fail(-1, "hb_env_random_actions: null engine");
  HB_CUDA(cudaSetDevice(e->device));
  return hb_launch_random_actions(e, counter);
}

}  // extern "C"
