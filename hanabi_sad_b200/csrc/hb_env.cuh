// hb_env.cuh -- Hanabi rules + observation features as device functions over an HbGame record.
//
// Behavioural contract (bit-exact, checked against oracle/ on the GPU by tests/test_env_parity.py):
//   rules        : hanabi_lib/hanabi_state.cc:107-388, hanabi_hand.cc:84-130, hanabi_game.cc:81-185
//   env adapter  : cpp/hanabi_env.cc:9-205, cpp/hanabi_env.h:79-146
//   features     : hanabi_lib/canonical_encoders.cc:70-727 (+ hanabi_observation.cc:34-95 for the
//                  observer-relative view)
// Design: the reference materialises 3 observation objects and several float vectors per player per step;
// here every feature is a pure function f(game record, observer, feature index), so a CTA simply strides
// its threads over the P*F output floats with coalesced stores and no intermediate objects.
#pragma once
#include "hb_types.h"

// HB_DEV functions are device code in the product.  The same header also compiles as plain C++ for the
// CPU-only logic tests (tests/host_emul), which is test scaffolding: no product entry point runs it.
#ifdef __CUDACC__
#define HB_DEV __host__ __device__ __forceinline__
#else
#define HB_DEV inline
#endif
#if defined(__CUDA_ARCH__)
#define HB_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define HB_FDIV(a, b) ((a) / (b))
#endif

struct HbEnvCfg {
  HbGeom g;
  int bomb;           // 0 | 1 | -1 (hanabi_state.cc:362-376)
  int max_len;        // <=0: no forced termination (hanabi_env.cc:97-101)
  int shuffle_color;
  int n_eps;
};

HB_DEV int hb_card_mult(int rank) { return rank == 0 ? 3 : (rank == HB_NR - 1 ? 1 : 2); }

HB_DEV int hb_score(const HbGame& s, int bomb) {
  int sc = 0;
#pragma unroll
  for (int c = 0; c < HB_NC; ++c) sc += s.fireworks[c];
  if (s.life == 0) {
    if (bomb == 0) return 0;
    if (bomb == -1) return sc > 0 ? sc - 1 : 0;
  }
  return sc;
}

HB_DEV bool hb_state_terminal(const HbGame& s, int bomb) {
  if (s.life < 1) return true;
  if (hb_score(s, bomb) >= HB_NCARD) return true;
  return s.turns_to_play <= 0;
}

// A decoded move (hanabi_game.cc:161-185).  Colour is in REAL colour space after hb_decode_move.
struct HbMove { int type, card_index, target_offset, color, rank; };

HB_DEV HbMove hb_decode_move(const HbGeom& g, int uid) {
  HbMove m = {HB_MV_INVALID, -1, -1, -1, -1};
  if (uid < 0 || uid >= g.A - 1) return m;
  if (uid < g.H) { m.type = HB_MV_DISCARD; m.card_index = uid; return m; }
  uid -= g.H;
  if (uid < g.H) { m.type = HB_MV_PLAY; m.card_index = uid; return m; }
  uid -= g.H;
  const int nrev = (g.P - 1) * HB_NC;
  if (uid < nrev) { m.type = HB_MV_REVEAL_COLOR; m.target_offset = 1 + uid / HB_NC; m.color = uid % HB_NC; return m; }
  uid -= nrev;
  m.type = HB_MV_REVEAL_RANK; m.target_offset = 1 + uid / HB_NR; m.rank = uid % HB_NR;
  return m;
}

// MoveIsLegal for the current player (hanabi_state.cc:169-222).
HB_DEV bool hb_move_legal(const HbGame& s, const HbGeom& g, const HbMove& m) {
  if (s.cur_player >= g.P) return false;
  const int cur = s.cur_player;
  switch (m.type) {
    case HB_MV_DISCARD: return s.info < HB_MAX_INFO && m.card_index < s.hand_len[cur];
    case HB_MV_PLAY: return m.card_index < s.hand_len[cur];
    case HB_MV_REVEAL_COLOR:
    case HB_MV_REVEAL_RANK: {
      if (s.info == 0) return false;
      if (m.target_offset < 1 || m.target_offset >= g.P) return false;
      const int tgt = (cur + m.target_offset) % g.P;
      for (int i = 0; i < s.hand_len[tgt]; ++i) {
        const int card = s.hand_card[tgt][i];
        if (m.type == HB_MV_REVEAL_COLOR ? (card / HB_NR == m.color) : (card % HB_NR == m.rank)) return true;
      }
      return false;
    }
    default: return false;
  }
}

// The history item the reference would record for `m` applied to `s` (hanabi_state.cc:224-278), computed
// WITHOUT mutating the state.  Used for the real move and for the SAD greedy clone alike.
HB_DEV HbLastMove hb_describe_move(const HbGame& s, const HbGeom& g, const HbMove& m) {
  HbLastMove lm;
  lm.valid = 1; lm.player = s.cur_player; lm.type = (uint8_t)m.type;
  lm.target_offset = 0; lm.color = 0; lm.rank = 0; lm.card_index = 0; lm.reveal_mask = 0;
  lm.card_color = 0; lm.card_rank = 0; lm.scored = 0; lm.info_token = 0;
  const int cur = s.cur_player;
  if (m.type == HB_MV_PLAY || m.type == HB_MV_DISCARD) {
    const int card = s.hand_card[cur][m.card_index];
    lm.card_index = (uint8_t)m.card_index;
    lm.card_color = (uint8_t)(card / HB_NR);
    lm.card_rank = (uint8_t)(card % HB_NR);
    if (m.type == HB_MV_PLAY) {
      const bool scored = lm.card_rank == s.fireworks[lm.card_color];
      lm.scored = scored;
      lm.info_token = scored && (s.fireworks[lm.card_color] + 1 == HB_NR) && s.info < HB_MAX_INFO;
    } else {
      lm.info_token = s.info < HB_MAX_INFO;
    }
  } else {
    const int tgt = (cur + m.target_offset) % g.P;
    lm.target_offset = (uint8_t)m.target_offset;
    unsigned mask = 0;
    for (int i = 0; i < s.hand_len[tgt]; ++i) {
      const int card = s.hand_card[tgt][i];
      if (m.type == HB_MV_REVEAL_COLOR ? (card / HB_NR == m.color) : (card % HB_NR == m.rank)) mask |= 1u << i;
    }
    lm.reveal_mask = (uint8_t)mask;
    if (m.type == HB_MV_REVEAL_COLOR) lm.color = (uint8_t)m.color; else lm.rank = (uint8_t)m.rank;
  }
  return lm;
}

HB_DEV void hb_remove_card(HbGame& s, int p, int idx) {  // hanabi_hand.cc:91-98
  const int n = s.hand_len[p];
  for (int i = idx; i + 1 < n; ++i) { s.hand_card[p][i] = s.hand_card[p][i + 1]; s.know[p][i] = s.know[p][i + 1]; }
  s.hand_card[p][n - 1] = HB_NO_CARD;
  s.know[p][n - 1] = HB_KNOW_BLANK;
  s.hand_len[p] = (uint8_t)(n - 1);
}

// Mutating half of ApplyMove for a player move (hanabi_state.cc:224-278); `lm` = hb_describe_move(s, m).
HB_DEV void hb_mutate(HbGame& s, const HbGeom& g, const HbMove& m, const HbLastMove& lm) {
  if (s.deck_pos >= HB_DECK) --s.turns_to_play;  // deck empty (hanabi_state.cc:226-228)
  const int cur = s.cur_player;
  switch (m.type) {
    case HB_MV_DISCARD:
      if (s.info < HB_MAX_INFO) ++s.info;
      ++s.discard_count[lm.card_color * HB_NR + lm.card_rank];
      hb_remove_card(s, cur, m.card_index);
      break;
    case HB_MV_PLAY:
      if (lm.scored) {
        ++s.fireworks[lm.card_color];
        if (lm.info_token) ++s.info;
      } else {
        --s.life;
        ++s.discard_count[lm.card_color * HB_NR + lm.card_rank];
      }
      hb_remove_card(s, cur, m.card_index);
      break;
    case HB_MV_REVEAL_COLOR: {
      --s.info;
      const int tgt = (cur + m.target_offset) % g.P;
      for (int i = 0; i < s.hand_len[tgt]; ++i) {  // hanabi_hand.cc:100-114
        uint16_t k = s.know[tgt][i];
        if (s.hand_card[tgt][i] / HB_NR == m.color) k = (uint16_t)((k & ~0x1C1Fu) | (1u << m.color) | ((unsigned)m.color << 10));
        else k = (uint16_t)(k & ~(1u << m.color));
        s.know[tgt][i] = k;
      }
      break;
    }
    case HB_MV_REVEAL_RANK: {
      --s.info;
      const int tgt = (cur + m.target_offset) % g.P;
      for (int i = 0; i < s.hand_len[tgt]; ++i) {  // hanabi_hand.cc:116-130
        uint16_t k = s.know[tgt][i];
        if (s.hand_card[tgt][i] % HB_NR == m.rank) k = (uint16_t)((k & ~0xE3E0u) | (1u << (5 + m.rank)) | ((unsigned)m.rank << 13));
        else k = (uint16_t)(k & ~(1u << (5 + m.rank)));
        s.know[tgt][i] = k;
      }
      break;
    }
    default: break;
  }
  // AdvanceToNextPlayer (hanabi_state.cc:107-114)
  bool short_hand = false;
  for (int p = 0; p < g.P; ++p) short_hand |= s.hand_len[p] < g.H;
  if (s.deck_pos < HB_DECK && short_hand) {
    s.cur_player = HB_CHANCE;
  } else {
    s.cur_player = s.next_player;
    s.next_player = (uint8_t)((s.cur_player + 1) % g.P);
  }
}

// Chance moves: deal from the pre-arranged deck until every hand is full or the deck is empty
// (hanabi_state.cc:232-244 with PlayerToDeal :160-167), then hand the turn on.
HB_DEV void hb_deal_pending(HbGame& s, const HbGeom& g, const uint8_t* __restrict__ deck) {
  while (s.cur_player == HB_CHANCE) {
    int p = 0;
    while (p < g.P && s.hand_len[p] >= g.H) ++p;
    const int n = s.hand_len[p];
    s.hand_card[p][n] = deck[s.deck_pos];
    s.know[p][n] = HB_KNOW_BLANK;
    s.hand_len[p] = (uint8_t)(n + 1);
    ++s.deck_pos;
    bool short_hand = false;
    for (int q = 0; q < g.P; ++q) short_hand |= s.hand_len[q] < g.H;
    if (!(s.deck_pos < HB_DECK && short_hand)) {
      s.cur_player = s.next_player;
      s.next_player = (uint8_t)((s.cur_player + 1) % g.P);
    }
  }
}

// HanabiEnv::reset minus the randomness (hanabi_env.cc:9-47): the deck order, eps indices and colour
// permutations of the new episode are already in place.
HB_DEV void hb_reset_game(HbGame& s, const HbGeom& g, const uint8_t* __restrict__ deck) {
  for (int p = 0; p < HB_MAX_P; ++p) {
    s.hand_len[p] = 0;
    for (int i = 0; i < HB_MAX_H; ++i) { s.hand_card[p][i] = HB_NO_CARD; s.know[p][i] = HB_KNOW_BLANK; }
  }
  for (int i = 0; i < HB_NCARD; ++i) s.discard_count[i] = 0;
  for (int c = 0; c < HB_NC; ++c) s.fireworks[c] = 0;
  s.info = HB_MAX_INFO; s.life = HB_MAX_LIFE;
  s.cur_player = HB_CHANCE; s.next_player = 0;
  s.turns_to_play = (int8_t)g.P;
  s.deck_pos = 0; s.num_step = 0;
  s.last.valid = 0; s.greedy.valid = 0; s.greedy_valid = 0;
  s.terminated = 0; s.reward = 0.f; s.ep_len = 0;
  hb_deal_pending(s, g, deck);
  ++s.episode;
}

// HanabiEnv::step minus observation encoding (hanabi_env.cc:49-108).  Returns terminal; reward in s.reward.
HB_DEV bool hb_step_game(HbGame& s, const HbEnvCfg& cfg, const uint8_t* __restrict__ deck,
                                             int action_uid, int greedy_uid) {
  const HbGeom& g = cfg.g;
  s.num_step += 1;
  const float prev_score = (float)hb_score(s, cfg.bomb);
  const int cur = s.cur_player;
  HbMove mv = hb_decode_move(g, action_uid);
  if (cfg.shuffle_color && mv.type == HB_MV_REVEAL_COLOR) mv.color = hb_perm_get(s.inv_perm[cur], mv.color);
  if (!hb_move_legal(s, g, mv)) { s.illegal = 1; s.terminated = 1; s.reward = 0.f; return true; }
  if (g.sad) {
    HbMove gm = hb_decode_move(g, greedy_uid);
    if (cfg.shuffle_color && gm.type == HB_MV_REVEAL_COLOR) gm.color = hb_perm_get(s.inv_perm[cur], gm.color);
    if (!hb_move_legal(s, g, gm)) { s.illegal = 1; s.terminated = 1; s.reward = 0.f; return true; }
    s.greedy = hb_describe_move(s, g, gm);
    s.greedy_valid = 1;
  }
  const HbLastMove lm = hb_describe_move(s, g, mv);
  hb_mutate(s, g, mv, lm);
  s.last = lm;
  bool terminal = hb_state_terminal(s, cfg.bomb);
  float reward = (float)hb_score(s, cfg.bomb) - prev_score;
  if (cfg.max_len > 0 && s.num_step == cfg.max_len) { terminal = true; reward = 0.f - prev_score; }
  if (!terminal) hb_deal_pending(s, g, deck);
  s.reward = reward;
  if (terminal) { s.terminated = 1; s.last_score = (int16_t)hb_score(s, cfg.bomb); }  // hanabi_env.h:92-94
  return terminal;
}

// ---------------------------------------------------------------------------------------- features

// Per-game tables shared by all feature evaluations of one game (built once per CTA in shared memory).
struct HbEncTables {
  uint8_t pub_count[HB_NCARD];            // ComputeCardCount(publ=true) in REAL colour space (canonical_encoders.cc:783-823)
  float belief_total[HB_MAX_P][HB_MAX_H]; // sum over plausible card types of pub_count (canonical_encoders.cc:553-563)
};

HB_DEV int hb_pub_count(const HbGame& s, int card) {
  const int c = card / HB_NR, r = card % HB_NR;
  return hb_card_mult(r) - s.discard_count[card] - (r < s.fireworks[c] ? 1 : 0);
}

HB_DEV float hb_belief_total(const HbGame& s, const HbEncTables& t, int p, int slot) {
  const unsigned k = s.know[p][slot];
  int tot = 0;
#pragma unroll
  for (int c = 0; c < HB_NC; ++c)
#pragma unroll
    for (int r = 0; r < HB_NR; ++r)
      if (((k >> c) & 1u) && ((k >> (5 + r)) & 1u)) tot += t.pub_count[c * HB_NR + r];
  return (float)tot;
}

// One entry of the (2P+2H+41)-wide last-action block (canonical_encoders.cc:293-422).
HB_DEV float hb_last_action_elem(const HbLastMove& lm, const HbGeom& g, int observer, uint16_t perm,
                                                     bool shuffle, int j) {
  if (!lm.valid) return 0.f;
  const int P = g.P, H = g.H;
  const int rel = (lm.player - observer + P) % P;  // hanabi_observation.cc:47
  const bool hint = lm.type == HB_MV_REVEAL_COLOR || lm.type == HB_MV_REVEAL_RANK;
  const bool card_move = lm.type == HB_MV_PLAY || lm.type == HB_MV_DISCARD;
  if (j < P) return j == rel ? 1.f : 0.f;
  j -= P;
  if (j < 4) return j == lm.type - 1 ? 1.f : 0.f;  // play, discard, reveal colour, reveal rank
  j -= 4;
  if (j < P) return (hint && j == (rel + lm.target_offset) % P) ? 1.f : 0.f;
  j -= P;
  if (j < HB_NC) {
    if (lm.type != HB_MV_REVEAL_COLOR) return 0.f;
    const int shown = shuffle ? hb_perm_get(perm, lm.color) : lm.color;
    return j == shown ? 1.f : 0.f;
  }
  j -= HB_NC;
  if (j < HB_NR) return (lm.type == HB_MV_REVEAL_RANK && j == lm.rank) ? 1.f : 0.f;
  j -= HB_NR;
  if (j < H) return (hint && ((lm.reveal_mask >> j) & 1)) ? 1.f : 0.f;
  j -= H;
  if (j < H) return (card_move && j == lm.card_index) ? 1.f : 0.f;
  j -= H;
  if (j < HB_NCARD) {
    if (!card_move) return 0.f;
    const int shown = shuffle ? hb_perm_get(perm, lm.card_color) : lm.card_color;
    return j == shown * HB_NR + lm.card_rank ? 1.f : 0.f;
  }
  j -= HB_NCARD;
  if (lm.type != HB_MV_PLAY) return 0.f;
  return j == 0 ? (lm.scored ? 1.f : 0.f) : (lm.info_token ? 1.f : 0.f);
}

// priv_s[observer][f]  (CanonicalObservationEncoder::Encode, canonical_encoders.cc:648-688, plus the SAD
// block appended by hanabi_env.cc:154-160).
HB_DEV float hb_feature(const HbGame& s, const HbEncTables& t, const HbEnvCfg& cfg, int observer, int f) {
  const HbGeom& g = cfg.g;
  const int P = g.P, H = g.H;
  const bool shuffle = cfg.shuffle_color != 0;
  const uint16_t perm = s.perm[observer], inv = s.inv_perm[observer];
  if (f < g.off_board) {  // ---- hands (canonical_encoders.cc:70-142)
    const int ncard_bits = P * H * HB_NCARD;
    if (f >= ncard_bits) return s.hand_len[(observer + f - ncard_bits) % P] < H ? 1.f : 0.f;
    const int rel = f / (H * HB_NCARD);
    if (rel == 0) return 0.f;  // own cards hidden
    const int r2 = f % (H * HB_NCARD);
    const int slot = r2 / HB_NCARD, k = r2 % HB_NCARD;
    const int p = (observer + rel) % P;
    if (slot >= s.hand_len[p]) return 0.f;
    const int shown_c = k / HB_NR;
    const int real_c = shuffle ? hb_perm_get(inv, shown_c) : shown_c;
    return s.hand_card[p][slot] == real_c * HB_NR + k % HB_NR ? 1.f : 0.f;
  }
  if (f < g.off_discard) {  // ---- board (canonical_encoders.cc:160-231)
    int j = f - g.off_board;
    if (j < g.deck_bits) return j < HB_DECK - s.deck_pos ? 1.f : 0.f;
    j -= g.deck_bits;
    if (j < HB_NCARD) {
      const int shown_c = j / HB_NR;
      const int real_c = shuffle ? hb_perm_get(inv, shown_c) : shown_c;
      const int fw = s.fireworks[real_c];
      return (fw > 0 && fw - 1 == j % HB_NR) ? 1.f : 0.f;
    }
    j -= HB_NCARD;
    if (j < HB_MAX_INFO) return j < s.info ? 1.f : 0.f;
    j -= HB_MAX_INFO;
    return j < s.life ? 1.f : 0.f;
  }
  if (f < g.off_last) {  // ---- discards: thermometers of widths 3,2,2,2,1 per shown colour (:252-280)
    const int j = f - g.off_discard;
    const int shown_c = j / 10, w = j % 10;
    const int r = w < 3 ? 0 : (w < 5 ? 1 : (w < 7 ? 2 : (w < 9 ? 3 : 4)));
    const int i = w - (r == 0 ? 0 : 2 * r + 1);
    const int real_c = shuffle ? hb_perm_get(inv, shown_c) : shown_c;
    return i < s.discard_count[real_c * HB_NR + r] ? 1.f : 0.f;
  }
  if (f < g.off_belief) return hb_last_action_elem(s.last, g, observer, perm, shuffle, f - g.off_last);
  if (f < g.off_sad) {  // ---- V0 belief (canonical_encoders.cc:450-581)
    const int j = f - g.off_belief;
    const int rel = j / (H * 35), r2 = j % (H * 35);
    const int slot = r2 / 35, k = r2 % 35;
    const int p = (observer + rel) % P;
    if (slot >= s.hand_len[p]) return 0.f;
    const unsigned kn = s.know[p][slot];
    if (k < HB_NCARD) {
      const int shown_c = k / HB_NR, r = k % HB_NR;
      const int real_c = shuffle ? hb_perm_get(inv, shown_c) : shown_c;
      if (!(((kn >> real_c) & 1u) && ((kn >> (5 + r)) & 1u))) return 0.f;
      const float total = t.belief_total[p][slot];
      // float(count) / float(total), IEEE round-to-nearest: the reference's `enc *= count; enc /= total`
      return total > 0.f ? HB_FDIV((float)t.pub_count[real_c * HB_NR + r], total) : 0.f;
    }
    if (k < HB_NCARD + HB_NC) {
      const int hinted = (kn >> 10) & 7;
      if (hinted == 7) return 0.f;
      const int shown = shuffle ? hb_perm_get(perm, hinted) : hinted;
      return k - HB_NCARD == shown ? 1.f : 0.f;
    }
    const int hinted = (kn >> 13) & 7;
    return (hinted != 7 && k - HB_NCARD - HB_NC == hinted) ? 1.f : 0.f;
  }
  // ---- SAD block: last action of the greedy clone; right after reset the clone IS the state (hanabi_env.cc:46)
  return hb_last_action_elem(s.greedy_valid ? s.greedy : s.last, g, observer, perm, shuffle, f - g.off_sad);
}

// legal_move[observer][uid]  (hanabi_env.cc:171-193 over hanabi_state.cc:291-307)
HB_DEV float hb_legal_elem(const HbGame& s, const HbEnvCfg& cfg, int observer, int uid) {
  const HbGeom& g = cfg.g;
  if (observer != s.cur_player) return uid == g.A - 1 ? 1.f : 0.f;  // others (and everyone at terminal): no-op only
  if (uid == g.A - 1) {  // no-op only when the acting player has no legal move at all (hanabi_env.cc:189-191)
    for (int u = 0; u < g.A - 1; ++u)
      if (hb_move_legal(s, g, hb_decode_move(g, u))) return 0.f;  // legality is colour-permutation invariant as a set
    return 1.f;
  }
  HbMove m = hb_decode_move(g, uid);
  // uid is in SHOWN colour space: shown colour c' is legal iff real colour inv[c'] is hintable
  if (cfg.shuffle_color && m.type == HB_MV_REVEAL_COLOR) m.color = hb_perm_get(s.inv_perm[observer], m.color);
  return hb_move_legal(s, g, m) ? 1.f : 0.f;
}

// own_hand[observer][3*slot + {playable, dead, future}]  (canonical_encoders.cc:690-727)
HB_DEV float hb_own_hand_elem(const HbGame& s, int observer, int j) {
  const int slot = j / 3, k = j % 3;
  if (slot >= s.hand_len[observer]) return 0.f;
  const int card = s.hand_card[observer][slot];
  const int fw = s.fireworks[card / HB_NR], r = card % HB_NR;
  const int cls = r == fw ? 0 : (r < fw ? 1 : 2);
  return k == cls ? 1.f : 0.f;
}
