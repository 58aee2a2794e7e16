// TEST SCAFFOLDING (CPU-only logic tests, `-m "not gpu"`): compiles the product's device header
// hanabi_sad_b200/csrc/hb_env.cuh as plain C++ so the rules/feature functions the CUDA kernels run can be
// diffed against the oracle inside this GPU-less container.  It is NOT a product path: nothing under
// hanabi_sad_b200/ loads it, and the package refuses to run without the CUDA library.
#include <cstring>
#include <cstdlib>
#include "../../hanabi_sad_b200/csrc/hb_env.cuh"

struct HostEnv {
  HbEnvCfg cfg;
  HbGame s;
  uint8_t deck[64];
};

extern "C" {

void* he_create(int P, int H, int sad, int shuffle_color, int bomb, int max_len) {
  HostEnv* e = (HostEnv*)calloc(1, sizeof(HostEnv));
  e->cfg.g = hb_make_geom(P, H, sad);
  e->cfg.bomb = bomb; e->cfg.max_len = max_len; e->cfg.shuffle_color = shuffle_color; e->cfg.n_eps = 1;
  e->s.terminated = 1;
  for (int p = 0; p < HB_MAX_P; ++p) { e->s.perm[p] = hb_perm_identity(); e->s.inv_perm[p] = hb_perm_identity(); }
  return e;
}
void he_destroy(void* h) { free(h); }
int he_feature_size(void* h) { return ((HostEnv*)h)->cfg.g.F; }
int he_num_action(void* h) { return ((HostEnv*)h)->cfg.g.A; }

void he_reset(void* h, const int8_t* deck50, const int* eps_idx, const int* perms) {
  HostEnv* e = (HostEnv*)h;
  for (int i = 0; i < HB_DECK; ++i) e->deck[i] = (uint8_t)deck50[i];
  for (int p = 0; p < e->cfg.g.P; ++p) {
    e->s.eps_idx[p] = (uint8_t)eps_idx[p];
    uint16_t pm = 0, inv = 0;
    for (int c = 0; c < HB_NC; ++c) {
      int v = (perms && e->cfg.shuffle_color) ? perms[p * HB_NC + c] : c;
      pm |= (uint16_t)(v << (3 * c));
      inv |= (uint16_t)(c << (3 * v));
    }
    e->s.perm[p] = pm; e->s.inv_perm[p] = inv;
  }
  hb_reset_game(e->s, e->cfg.g, e->deck);
}

int he_step(void* h, const int64_t* a, const int64_t* g, float* reward) {
  HostEnv* e = (HostEnv*)h;
  int cur = e->s.cur_player;
  bool t = hb_step_game(e->s, e->cfg, e->deck, (int)a[cur], (int)g[cur]);
  *reward = e->s.reward;
  return e->s.illegal ? -1 : (t ? 1 : 0);
}

void he_observe(void* h, float* priv_s, float* legal, float* own) {
  HostEnv* e = (HostEnv*)h;
  const HbGeom& g = e->cfg.g;
  HbEncTables t;
  for (int k = 0; k < HB_NCARD; ++k) t.pub_count[k] = (uint8_t)hb_pub_count(e->s, k);
  for (int p = 0; p < g.P; ++p)
    for (int i = 0; i < g.H; ++i) t.belief_total[p][i] = i < e->s.hand_len[p] ? hb_belief_total(e->s, t, p, i) : 0.f;
  for (int o = 0; o < g.P; ++o) {
    for (int f = 0; f < g.F; ++f) priv_s[o * g.F + f] = hb_feature(e->s, t, e->cfg, o, f);
    for (int u = 0; u < g.A; ++u) legal[o * g.A + u] = hb_legal_elem(e->s, e->cfg, o, u);
    for (int j = 0; j < 3 * g.H; ++j) own[o * 3 * g.H + j] = hb_own_hand_elem(e->s, o, j);
  }
}

int he_cur_player(void* h) { HostEnv* e = (HostEnv*)h; return e->s.cur_player == HB_CHANCE ? -1 : e->s.cur_player; }
int he_last_score(void* h) { return ((HostEnv*)h)->s.last_score; }
int he_terminated(void* h) { return ((HostEnv*)h)->s.terminated; }
int he_deck_pos(void* h) { return ((HostEnv*)h)->s.deck_pos; }
}
