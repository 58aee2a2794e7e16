/*
 * hanabi_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the product).
 *
 * A plain-C, single-threaded, CPU restatement of the reference's Hanabi hot path
 * (facebookresearch/hanabi_SAD @ 415804b), written to be the bit-exact checker for the CUDA
 * kernels in hanabi_sad_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   (1) the 41 sha256 known-answer hashes of SURVEY.md Appendix B (generated from the unmodified
 *       reference build), and
 *   (2) the unmodified reference itself (oracle/_ref/hanalearn*.so, built by oracle/build_ref.sh)
 *       step by step on random seeds, and the committed fixtures in tests/golden/.
 *
 * Third-party arithmetic restated here (absent from /root/reference): libstdc++ <random> as shipped
 * with GCC 13.3 -- std::mt19937, std::discrete_distribution<unsigned long>::operator()
 * (bits/random.tcc:2657-2714), std::generate_canonical<double,53> (random.tcc, 2 draws),
 * std::shuffle's two-at-a-time path (bits/stl_algo.h:3742-3800) and
 * std::uniform_int_distribution's Lemire downscaling (bits/uniform_int_dist.h _S_nd).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_MAX_PLAYERS 5
#define ORC_MAX_HAND 5
#define ORC_NUM_COLORS 5
#define ORC_NUM_RANKS 5
#define ORC_NUM_CARDS 25
#define ORC_DECK 50
#define ORC_MAX_INFO 8
#define ORC_MAX_LIFE 3
#define ORC_MAX_HISTORY 512
#define ORC_CHANCE_PLAYER (-1)

enum { MV_INVALID = 0, MV_PLAY = 1, MV_DISCARD = 2, MV_REVEAL_COLOR = 3, MV_REVEAL_RANK = 4, MV_DEAL = 5 };

/* ---------------------------------------------------------------- libstdc++ <random> restated */

typedef struct {
  uint32_t mt[624];
  int idx;
  uint64_t draws; /* number of 32-bit outputs produced so far (diagnostics / stream export) */
} orc_mt19937;

static void mt_seed(orc_mt19937* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
  g->draws = 0;
}

static uint32_t mt_next(orc_mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  g->draws++;
  return y;
}

/* std::generate_canonical<double,53>(mt19937): m = 2 draws, low word first (random.tcc). */
static double canonical_double(orc_mt19937* g) {
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; ++k) {
    sum += (double)mt_next(g) * tmp;
    tmp *= 4294967296.0;
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = 0.99999999999999988897769753748; /* nextafter(1.0, 0.0) */
  return ret;
}

/* std::discrete_distribution<unsigned long>(probs)(rng)  (random.tcc:2657-2714).
 * < 2 weights: returns 0 WITHOUT drawing. */
static int discrete_draw(orc_mt19937* g, const double* w, int n) {
  if (n < 2) return 0;
  double prob[ORC_NUM_CARDS], cp[ORC_NUM_CARDS];
  double sum = 0.0;
  for (int i = 0; i < n; ++i) sum += w[i];
  for (int i = 0; i < n; ++i) prob[i] = w[i] / sum;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) { /* std::partial_sum: first element copied, then running + */
    acc = (i == 0) ? prob[0] : acc + prob[i];
    cp[i] = acc;
  }
  cp[n - 1] = 1.0;
  double p = canonical_double(g);
  int pos = 0; /* std::lower_bound: first cp[pos] >= p */
  while (pos < n && cp[pos] < p) ++pos;
  return pos;
}

/* uniform_int_distribution<unsigned long>{0, range-1}(mt19937): Lemire (_S_nd<uint64_t>). */
static uint32_t lemire_below(orc_mt19937* g, uint32_t range) {
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (uint32_t)(0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

/* std::shuffle(v.begin(), v.end(), mt19937) for small n (urngrange/n >= n path, stl_algo.h:3775-3800). */
static void std_shuffle_int(orc_mt19937* g, int* v, int n) {
  if (n == 0) return;
  int i = 1;
#define ORC_SWAP(a, b) do { int _t = v[a]; v[a] = v[b]; v[b] = _t; } while (0)
  if ((n % 2) == 0) {
    uint32_t j = lemire_below(g, 2);
    ORC_SWAP(i, (int)j);
    ++i;
  }
  while (i != n) {
    uint32_t swap_range = (uint32_t)i + 1;
    uint32_t x = lemire_below(g, swap_range * (swap_range + 1));
    uint32_t p0 = x / (swap_range + 1), p1 = x % (swap_range + 1);
    ORC_SWAP(i, (int)p0);
    ++i;
    ORC_SWAP(i, (int)p1);
    ++i;
  }
#undef ORC_SWAP
}

/* ---------------------------------------------------------------- game objects */

typedef struct { int8_t color, rank; } orc_card; /* (-1,-1) == invalid (hanabi_card.h) */

typedef struct { /* HanabiHand::CardKnowledge (hanabi_hand.h:54-94) */
  int8_t color_value;              /* -1 = not hinted */
  int8_t rank_value;
  uint8_t color_plausible[ORC_NUM_COLORS];
  uint8_t rank_plausible[ORC_NUM_RANKS];
} orc_knowledge;

typedef struct {
  int n;
  orc_card cards[ORC_MAX_HAND + 1];
  orc_knowledge know[ORC_MAX_HAND + 1];
} orc_hand;

typedef struct { /* HanabiMove (hanabi_move.h) */
  int8_t type, card_index, target_offset, color, rank;
} orc_move;

typedef struct { /* HanabiHistoryItem (hanabi_history_item.h:27-57) */
  orc_move move;
  int8_t player;
  uint8_t scored, information_token;
  int8_t color, rank;
  uint8_t reveal_bitmask, newly_revealed_bitmask;
  int8_t deal_to_player;
} orc_history_item;

typedef struct { /* HanabiState (hanabi_state.h:192-203) */
  int card_count[ORC_NUM_CARDS]; /* HanabiDeck::card_count_ */
  int total_count;
  orc_hand hands[ORC_MAX_PLAYERS];
  orc_card discard_pile[ORC_DECK];
  int n_discard;
  int cur_player, next_non_chance_player;
  int information_tokens, life_tokens;
  int fireworks[ORC_NUM_COLORS];
  int turns_to_play;
  orc_history_item history[ORC_MAX_HISTORY];
  int n_history;
  int valid; /* state_ != nullptr */
} orc_state;

typedef struct OrcEnv {
  /* HanabiGame (hanabi_game.cc:29-69) */
  int players, hand_size, bomb, seed;
  orc_mt19937 rng;
  /* HanabiEnv (cpp/hanabi_env.h:18-50) */
  float eps_list[256];
  int n_eps;
  int max_len, sad, shuffle_color;
  float player_eps[ORC_MAX_PLAYERS];
  int num_step;
  int color_permute[ORC_MAX_PLAYERS][ORC_NUM_COLORS];
  int inv_color_permute[ORC_MAX_PLAYERS][ORC_NUM_COLORS];
  int last_score;
  orc_state state;
  orc_state clone; /* SAD greedy clone (hanabi_env.cc:82-91) */
  int clone_valid; /* 1 = clone holds the greedy-applied state; 0 = use state (reset) */
  /* injected randomness (RNG-independent parity mode) */
  int inject;
  int8_t inj_deck[ORC_DECK]; /* card index c*5+r in deal order */
  int inj_pos;
  int inj_eps_idx[ORC_MAX_PLAYERS];
  /* exported stream of the current episode (whatever mode produced it) */
  int8_t dealt[ORC_DECK];
  int n_dealt;
  int eps_idx[ORC_MAX_PLAYERS];
} OrcEnv;

/* ---------------------------------------------------------------- HanabiGame */

static int number_card_instances(int rank) { /* hanabi_game.cc:128-138 */
  if (rank == 0) return 3;
  if (rank == ORC_NUM_RANKS - 1) return 1;
  return 2;
}
static int max_discard_moves(const OrcEnv* e) { return e->hand_size; }
static int max_play_moves(const OrcEnv* e) { return e->hand_size; }
static int max_reveal_color_moves(const OrcEnv* e) { return (e->players - 1) * ORC_NUM_COLORS; }
static int max_reveal_rank_moves(const OrcEnv* e) { return (e->players - 1) * ORC_NUM_RANKS; }
static int max_moves(const OrcEnv* e) {
  return max_discard_moves(e) + max_play_moves(e) + max_reveal_color_moves(e) + max_reveal_rank_moves(e);
}

static orc_move construct_move(const OrcEnv* e, int uid) { /* hanabi_game.cc:161-185 */
  orc_move m = {MV_INVALID, -1, -1, -1, -1};
  if (uid < 0 || uid >= max_moves(e)) return m;
  if (uid < max_discard_moves(e)) { m.type = MV_DISCARD; m.card_index = (int8_t)uid; return m; }
  uid -= max_discard_moves(e);
  if (uid < max_play_moves(e)) { m.type = MV_PLAY; m.card_index = (int8_t)uid; return m; }
  uid -= max_play_moves(e);
  if (uid < max_reveal_color_moves(e)) {
    m.type = MV_REVEAL_COLOR; m.target_offset = (int8_t)(1 + uid / ORC_NUM_COLORS); m.color = (int8_t)(uid % ORC_NUM_COLORS);
    return m;
  }
  uid -= max_reveal_color_moves(e);
  m.type = MV_REVEAL_RANK; m.target_offset = (int8_t)(1 + uid / ORC_NUM_RANKS); m.rank = (int8_t)(uid % ORC_NUM_RANKS);
  return m;
}

static int get_move_uid(const OrcEnv* e, orc_move m) { /* hanabi_game.cc:81-97 */
  switch (m.type) {
    case MV_DISCARD: return m.card_index;
    case MV_PLAY: return max_discard_moves(e) + m.card_index;
    case MV_REVEAL_COLOR: return max_discard_moves(e) + max_play_moves(e) + (m.target_offset - 1) * ORC_NUM_COLORS + m.color;
    case MV_REVEAL_RANK:
      return max_discard_moves(e) + max_play_moves(e) + max_reveal_color_moves(e) + (m.target_offset - 1) * ORC_NUM_RANKS + m.rank;
    default: return -1;
  }
}

/* ---------------------------------------------------------------- HanabiState */

static void state_init(const OrcEnv* e, orc_state* s) { /* hanabi_state.cc:54-66,93-105 */
  memset(s, 0, sizeof(*s));
  for (int c = 0; c < ORC_NUM_COLORS; ++c)
    for (int r = 0; r < ORC_NUM_RANKS; ++r) {
      s->card_count[c * ORC_NUM_RANKS + r] = number_card_instances(r);
      s->total_count += number_card_instances(r);
    }
  s->cur_player = ORC_CHANCE_PLAYER;
  s->next_non_chance_player = 0; /* random_start_player=false (hanabi_game.cc:140-147) */
  s->information_tokens = ORC_MAX_INFO;
  s->life_tokens = ORC_MAX_LIFE;
  s->turns_to_play = e->players;
  s->valid = 1;
}

static int deck_empty(const orc_state* s) { return s->total_count == 0; }

static int player_to_deal(const OrcEnv* e, const orc_state* s) { /* hanabi_state.cc:160-167 */
  for (int i = 0; i < e->players; ++i)
    if (s->hands[i].n < e->hand_size) return i;
  return -1;
}

static void advance_to_next_player(const OrcEnv* e, orc_state* s) { /* hanabi_state.cc:107-114 */
  if (!deck_empty(s) && player_to_deal(e, s) >= 0) {
    s->cur_player = ORC_CHANCE_PLAYER;
  } else {
    s->cur_player = s->next_non_chance_player;
    s->next_non_chance_player = (s->cur_player + 1) % e->players;
  }
}

static orc_hand* hand_by_offset(const OrcEnv* e, orc_state* s, int offset) { /* hanabi_state.h:176-181 */
  return &s->hands[(s->cur_player + offset) % e->players];
}

static int move_is_legal(const OrcEnv* e, const orc_state* s, orc_move m) { /* hanabi_state.cc:169-222 */
  switch (m.type) {
    case MV_DEAL:
      if (s->cur_player != ORC_CHANCE_PLAYER) return 0;
      if (s->card_count[m.color * ORC_NUM_RANKS + m.rank] == 0) return 0;
      return 1;
    case MV_DISCARD:
      if (s->information_tokens >= ORC_MAX_INFO) return 0;
      if (m.card_index >= s->hands[s->cur_player].n) return 0;
      return 1;
    case MV_PLAY:
      if (m.card_index >= s->hands[s->cur_player].n) return 0;
      return 1;
    case MV_REVEAL_COLOR:
    case MV_REVEAL_RANK: {
      if (s->information_tokens <= 0) return 0; /* HintingIsLegal :149-158 */
      if (m.target_offset < 1 || m.target_offset >= e->players) return 0;
      const orc_hand* h = &s->hands[(s->cur_player + m.target_offset) % e->players];
      for (int i = 0; i < h->n; ++i) {
        if (m.type == MV_REVEAL_COLOR && h->cards[i].color == m.color) return 1;
        if (m.type == MV_REVEAL_RANK && h->cards[i].rank == m.rank) return 1;
      }
      return 0;
    }
    default: return 0;
  }
}

static void knowledge_blank(orc_knowledge* k) { /* hanabi_hand.cc:25-28,48-49 */
  k->color_value = -1;
  k->rank_value = -1;
  memset(k->color_plausible, 1, sizeof(k->color_plausible));
  memset(k->rank_plausible, 1, sizeof(k->rank_plausible));
}

static void remove_from_hand(orc_state* s, orc_hand* h, int idx, int to_discard) { /* hanabi_hand.cc:91-98 */
  if (to_discard) s->discard_pile[s->n_discard++] = h->cards[idx];
  for (int i = idx; i + 1 < h->n; ++i) { h->cards[i] = h->cards[i + 1]; h->know[i] = h->know[i + 1]; }
  h->n--;
}

static int increment_information_tokens(orc_state* s) { /* hanabi_state.cc:116-123 */
  if (s->information_tokens < ORC_MAX_INFO) { ++s->information_tokens; return 1; }
  return 0;
}

static void apply_move(const OrcEnv* e, orc_state* s, orc_move m) { /* hanabi_state.cc:224-278 */
  if (!move_is_legal(e, s, m)) { fprintf(stderr, "oracle: illegal move in apply_move\n"); abort(); }
  if (deck_empty(s)) --s->turns_to_play;
  orc_history_item h;
  memset(&h, 0, sizeof(h));
  h.move = m; h.player = (int8_t)s->cur_player; h.color = -1; h.rank = -1; h.deal_to_player = -1;
  switch (m.type) {
    case MV_DEAL: {
      h.deal_to_player = (int8_t)player_to_deal(e, s);
      orc_hand* hand = &s->hands[h.deal_to_player];
      int index = m.color * ORC_NUM_RANKS + m.rank; /* HanabiDeck::DealCard(color, rank) :80-91 */
      --s->card_count[index];
      --s->total_count;
      hand->cards[hand->n].color = m.color; hand->cards[hand->n].rank = m.rank;
      knowledge_blank(&hand->know[hand->n]);
      hand->n++;
      break;
    }
    case MV_DISCARD: {
      orc_hand* hand = &s->hands[s->cur_player];
      h.information_token = (uint8_t)increment_information_tokens(s);
      h.color = hand->cards[m.card_index].color; h.rank = hand->cards[m.card_index].rank;
      remove_from_hand(s, hand, m.card_index, 1);
      break;
    }
    case MV_PLAY: {
      orc_hand* hand = &s->hands[s->cur_player];
      orc_card c = hand->cards[m.card_index];
      h.color = c.color; h.rank = c.rank;
      if (c.rank == s->fireworks[c.color]) { /* AddToFireworks :135-147 */
        ++s->fireworks[c.color];
        h.scored = 1;
        if (s->fireworks[c.color] == ORC_NUM_RANKS) h.information_token = (uint8_t)increment_information_tokens(s);
      } else {
        --s->life_tokens;
      }
      remove_from_hand(s, hand, m.card_index, h.scored ? 0 : 1);
      break;
    }
    case MV_REVEAL_COLOR: {
      --s->information_tokens;
      orc_hand* hand = hand_by_offset(e, s, m.target_offset);
      for (int i = 0; i < hand->n; ++i) { /* HandColorBitmask :28-38 + RevealColor hanabi_hand.cc:100-114 */
        if (hand->cards[i].color == m.color) {
          h.reveal_bitmask |= (uint8_t)(1u << i);
          if (hand->know[i].color_value < 0) h.newly_revealed_bitmask |= (uint8_t)(1u << i);
          hand->know[i].color_value = m.color; /* ApplyIsValueHint :30-40 */
          memset(hand->know[i].color_plausible, 0, ORC_NUM_COLORS);
          hand->know[i].color_plausible[m.color] = 1;
        } else {
          hand->know[i].color_plausible[m.color] = 0;
        }
      }
      break;
    }
    case MV_REVEAL_RANK: {
      --s->information_tokens;
      orc_hand* hand = hand_by_offset(e, s, m.target_offset);
      for (int i = 0; i < hand->n; ++i) {
        if (hand->cards[i].rank == m.rank) {
          h.reveal_bitmask |= (uint8_t)(1u << i);
          if (hand->know[i].rank_value < 0) h.newly_revealed_bitmask |= (uint8_t)(1u << i);
          hand->know[i].rank_value = m.rank;
          memset(hand->know[i].rank_plausible, 0, ORC_NUM_RANKS);
          hand->know[i].rank_plausible[m.rank] = 1;
        } else {
          hand->know[i].rank_plausible[m.rank] = 0;
        }
      }
      break;
    }
    default: abort();
  }
  if (s->n_history >= ORC_MAX_HISTORY) { fprintf(stderr, "oracle: history overflow\n"); abort(); }
  s->history[s->n_history++] = h;
  advance_to_next_player(e, s);
}

static int state_score(const OrcEnv* e, const orc_state* s) { /* hanabi_state.cc:362-376 */
  int score = 0;
  for (int c = 0; c < ORC_NUM_COLORS; ++c) score += s->fireworks[c];
  if (s->life_tokens <= 0) {
    if (e->bomb == 0) return 0;
    if (e->bomb == -1) return score - 1 > 0 ? score - 1 : 0;
    if (e->bomb == 1) return score;
  }
  return score;
}

static int state_is_terminal(const OrcEnv* e, const orc_state* s) { /* hanabi_state.cc:378-388 */
  if (s->life_tokens < 1) return 1;
  if (state_score(e, s) >= ORC_NUM_COLORS * ORC_NUM_RANKS) return 1;
  if (s->turns_to_play <= 0) return 1;
  return 0;
}

/* ApplyRandomChance (hanabi_state.cc:285-289) -> ChanceOutcomes (:316-328) -> PickRandomChance
 * (hanabi_game.cc:108-114).  In inject mode the next card of the injected deck is dealt instead. */
static void apply_random_chance(OrcEnv* e, orc_state* s) {
  int idx;
  if (e->inject) {
    if (e->inj_pos >= ORC_DECK) { fprintf(stderr, "oracle: injected deck exhausted\n"); abort(); }
    idx = e->inj_deck[e->inj_pos++];
  } else {
    int uids[ORC_NUM_CARDS];
    double probs[ORC_NUM_CARDS];
    int n = 0;
    for (int uid = 0; uid < ORC_NUM_CARDS; ++uid) {
      if (s->card_count[uid] > 0) {
        uids[n] = uid;
        probs[n] = (double)s->card_count[uid] / (double)s->total_count;
        ++n;
      }
    }
    if (n == 0) abort();
    idx = uids[discrete_draw(&e->rng, probs, n)];
  }
  orc_move m = {MV_DEAL, -1, -1, (int8_t)(idx / ORC_NUM_RANKS), (int8_t)(idx % ORC_NUM_RANKS)};
  e->dealt[e->n_dealt++] = (int8_t)idx;
  apply_move(e, s, m);
}

/* ---------------------------------------------------------------- observation + encoder */

/* The reference builds HanabiObservation(state, observer) (hanabi_observation.cc:52-95): hands rotated so
 * index 0 is the observer; last_moves_ = history walked backwards up to and including the observer's own
 * previous move, players made observer-relative (:34-49).  The encoder then reads only the FIRST non-deal
 * item of last_moves_ (canonical_encoders.cc:34-41,306).  Since the backwards walk only ever stops ON a
 * non-deal item, that is always the most recent non-deal move of the whole history (or none). */
static const orc_history_item* last_non_deal(const orc_state* s) {
  for (int i = s->n_history - 1; i >= 0; --i)
    if (s->history[i].move.type != MV_DEAL) return &s->history[i];
  return NULL;
}

static int card_index(int color, int rank, int shuffle_color, const int* perm) { /* canonical_encoders.cc:48-58 */
  if (shuffle_color) color = perm[color];
  return color * ORC_NUM_RANKS + rank;
}

static int hands_section_len(const OrcEnv* e) { return e->players * e->hand_size * ORC_NUM_CARDS + e->players; }
static int board_section_len(const OrcEnv* e) {
  return ORC_DECK - e->players * e->hand_size + ORC_NUM_CARDS + ORC_MAX_INFO + ORC_MAX_LIFE;
}
static int discard_section_len(const OrcEnv* e) { (void)e; return ORC_DECK; }
static int last_action_section_len(const OrcEnv* e) { /* canonical_encoders.cc:585-595 */
  return e->players + 4 + e->players + ORC_NUM_COLORS + ORC_NUM_RANKS + e->hand_size + e->hand_size + ORC_NUM_CARDS + 2;
}
static int knowledge_section_len(const OrcEnv* e) { return e->players * e->hand_size * (ORC_NUM_CARDS + ORC_NUM_COLORS + ORC_NUM_RANKS); }

static int encoder_shape(const OrcEnv* e) { /* canonical_encoders.cc:597-606 */
  return hands_section_len(e) + board_section_len(e) + discard_section_len(e) + last_action_section_len(e) + knowledge_section_len(e);
}

/* EncodeHands, show_own_cards=false, order={} (canonical_encoders.cc:70-142) */
static int encode_hands(const OrcEnv* e, const orc_state* s, int observer, int start, const int* perm, float* enc) {
  int offset = start;
  for (int rel = 0; rel < e->players; ++rel) {
    const orc_hand* h = &s->hands[(observer + rel) % e->players];
    int num_cards = 0;
    for (int i = 0; i < h->n; ++i) {
      if (rel != 0) enc[offset + card_index(h->cards[i].color, h->cards[i].rank, e->shuffle_color, perm)] = 1;
      ++num_cards;
      offset += ORC_NUM_CARDS;
    }
    if (num_cards < e->hand_size) offset += (e->hand_size - num_cards) * ORC_NUM_CARDS;
  }
  for (int rel = 0; rel < e->players; ++rel)
    if (s->hands[(observer + rel) % e->players].n < e->hand_size) enc[offset + rel] = 1;
  offset += e->players;
  return offset - start;
}

/* EncodeBoard (canonical_encoders.cc:160-231) */
static int encode_board(const OrcEnv* e, const orc_state* s, int start, const int* inv_perm, float* enc) {
  int offset = start;
  for (int i = 0; i < s->total_count; ++i) enc[offset + i] = 1;
  offset += ORC_DECK - e->hand_size * e->players;
  for (int c = 0; c < ORC_NUM_COLORS; ++c) {
    int color = e->shuffle_color ? inv_perm[c] : c;
    if (s->fireworks[color] > 0) enc[offset + s->fireworks[color] - 1] = 1;
    offset += ORC_NUM_RANKS;
  }
  for (int i = 0; i < s->information_tokens; ++i) enc[offset + i] = 1;
  offset += ORC_MAX_INFO;
  for (int i = 0; i < s->life_tokens; ++i) enc[offset + i] = 1;
  offset += ORC_MAX_LIFE;
  return offset - start;
}

/* EncodeDiscards (canonical_encoders.cc:252-280) */
static int encode_discards(const OrcEnv* e, const orc_state* s, int start, const int* perm, float* enc) {
  int offset = start;
  int discard_counts[ORC_NUM_CARDS] = {0};
  for (int i = 0; i < s->n_discard; ++i)
    ++discard_counts[card_index(s->discard_pile[i].color, s->discard_pile[i].rank, e->shuffle_color, perm)];
  for (int c = 0; c < ORC_NUM_COLORS; ++c)
    for (int r = 0; r < ORC_NUM_RANKS; ++r) {
      int nd = discard_counts[c * ORC_NUM_RANKS + r];
      for (int i = 0; i < nd; ++i) enc[offset + i] = 1;
      offset += number_card_instances(r);
    }
  return offset - start;
}

/* EncodeLastAction_ (canonical_encoders.cc:293-422), order={} */
static int encode_last_action(const OrcEnv* e, const orc_state* s, int observer, int start, const int* perm, float* enc) {
  int offset = start;
  const orc_history_item* lm = last_non_deal(s);
  if (lm == NULL) return last_action_section_len(e);
  int t = lm->move.type;
  int rel_player = (lm->player - observer + e->players) % e->players; /* hanabi_observation.cc:47 */
  enc[offset + rel_player] = 1;
  offset += e->players;
  switch (t) {
    case MV_PLAY: enc[offset] = 1; break;
    case MV_DISCARD: enc[offset + 1] = 1; break;
    case MV_REVEAL_COLOR: enc[offset + 2] = 1; break;
    case MV_REVEAL_RANK: enc[offset + 3] = 1; break;
    default: abort();
  }
  offset += 4;
  if (t == MV_REVEAL_COLOR || t == MV_REVEAL_RANK) {
    int target = (rel_player + lm->move.target_offset) % e->players;
    enc[offset + target] = 1;
  }
  offset += e->players;
  if (t == MV_REVEAL_COLOR) {
    int color = lm->move.color;
    if (e->shuffle_color) color = perm[color];
    enc[offset + color] = 1;
  }
  offset += ORC_NUM_COLORS;
  if (t == MV_REVEAL_RANK) enc[offset + lm->move.rank] = 1;
  offset += ORC_NUM_RANKS;
  if (t == MV_REVEAL_COLOR || t == MV_REVEAL_RANK) {
    for (int i = 0, mask = 1; i < e->hand_size; ++i, mask <<= 1)
      if ((lm->reveal_bitmask & mask) > 0) enc[offset + i] = 1;
  }
  offset += e->hand_size;
  if (t == MV_PLAY || t == MV_DISCARD) enc[offset + lm->move.card_index] = 1;
  offset += e->hand_size;
  if (t == MV_PLAY || t == MV_DISCARD) enc[offset + card_index(lm->color, lm->rank, e->shuffle_color, perm)] = 1;
  offset += ORC_NUM_CARDS;
  if (t == MV_PLAY) {
    if (lm->scored) enc[offset] = 1;
    if (lm->information_token) enc[offset + 1] = 1;
  }
  offset += 2;
  return offset - start;
}

/* ComputeCardCount(publ=true) (canonical_encoders.cc:783-823) */
static void compute_card_count(const OrcEnv* e, const orc_state* s, const int* perm, int* card_count) {
  for (int c = 0; c < ORC_NUM_COLORS; ++c)
    for (int r = 0; r < ORC_NUM_RANKS; ++r) card_count[card_index(c, r, e->shuffle_color, perm)] = number_card_instances(r);
  for (int i = 0; i < s->n_discard; ++i)
    --card_count[card_index(s->discard_pile[i].color, s->discard_pile[i].rank, e->shuffle_color, perm)];
  for (int c = 0; c < ORC_NUM_COLORS; ++c)
    for (int r = 0; r < s->fireworks[c]; ++r) --card_count[card_index(c, r, e->shuffle_color, perm)];
}

/* EncodeV0Belief_ = EncodeCardKnowledge (canonical_encoders.cc:450-519) then count-weighting + float
 * normalisation (:521-581).  Observer's own knowledge included (hide_knowledge is false). */
static int encode_v0_belief(const OrcEnv* e, const orc_state* s, int observer, int start, const int* perm, float* enc) {
  int card_count[ORC_NUM_CARDS];
  compute_card_count(e, s, perm, card_count);
  const int per_card = ORC_NUM_CARDS + ORC_NUM_COLORS + ORC_NUM_RANKS;
  int offset = start;
  for (int rel = 0; rel < e->players; ++rel) {
    const orc_hand* h = &s->hands[(observer + rel) % e->players];
    int num_cards = 0;
    for (int i = 0; i < h->n; ++i) {
      const orc_knowledge* k = &h->know[i];
      for (int color = 0; color < ORC_NUM_COLORS; ++color)
        if (k->color_plausible[color])
          for (int rank = 0; rank < ORC_NUM_RANKS; ++rank)
            if (k->rank_plausible[rank]) enc[offset + card_index(color, rank, e->shuffle_color, perm)] = 1;
      offset += ORC_NUM_CARDS;
      if (k->color_value >= 0) {
        int color = k->color_value;
        if (e->shuffle_color) color = perm[color];
        enc[offset + color] = 1;
      }
      offset += ORC_NUM_COLORS;
      if (k->rank_value >= 0) enc[offset + k->rank_value] = 1;
      offset += ORC_NUM_RANKS;
      ++num_cards;
    }
    if (num_cards < e->hand_size) offset += (e->hand_size - num_cards) * per_card;
  }
  const int len = offset - start;
  const int player_offset = len / e->players;
  for (int rel = 0; rel < e->players; ++rel) {
    int num_cards = s->hands[(observer + rel) % e->players].n;
    for (int ci = 0; ci < num_cards; ++ci) {
      float total = 0;
      for (int i = 0; i < ORC_NUM_CARDS; ++i) {
        int o = start + player_offset * rel + ci * per_card + i;
        enc[o] *= card_count[i]; /* float *= int */
        total += enc[o];
      }
      if (total <= 0) { fprintf(stderr, "oracle: belief total = 0\n"); abort(); }
      for (int i = 0; i < ORC_NUM_CARDS; ++i) {
        int o = start + player_offset * rel + ci * per_card + i;
        enc[o] /= total;
      }
    }
  }
  return len;
}

/* CanonicalObservationEncoder::Encode(obs, false, {}, shuffle_color, perm, inv_perm, false)
 * (canonical_encoders.cc:648-688) */
static void encode_obs(const OrcEnv* e, const orc_state* s, int observer, float* enc) {
  const int* perm = e->color_permute[observer];
  const int* inv = e->inv_color_permute[observer];
  int n = encoder_shape(e);
  memset(enc, 0, sizeof(float) * (size_t)n);
  int offset = 0;
  offset += encode_hands(e, s, observer, offset, perm, enc);
  offset += encode_board(e, s, offset, inv, enc);
  offset += encode_discards(e, s, offset, perm, enc);
  offset += encode_last_action(e, s, observer, offset, perm, enc);
  offset += encode_v0_belief(e, s, observer, offset, perm, enc);
  if (offset != n) { fprintf(stderr, "oracle: encode length mismatch %d vs %d\n", offset, n); abort(); }
}

/* EncodeOwnHandTrinary (canonical_encoders.cc:690-727) on the cheat observation (show_cards=true) */
static void encode_own_hand_trinary(const OrcEnv* e, const orc_state* s, int observer, float* enc) {
  memset(enc, 0, sizeof(float) * (size_t)(e->hand_size * 3));
  const orc_hand* h = &s->hands[observer];
  int offset = 0;
  for (int i = 0; i < h->n; ++i) {
    int fw = s->fireworks[h->cards[i].color];
    if (h->cards[i].rank == fw) enc[offset] = 1;
    else if (h->cards[i].rank < fw) enc[offset + 1] = 1;
    else enc[offset + 2] = 1;
    offset += 3;
  }
}

/* ---------------------------------------------------------------- HanabiEnv */

int orc_feature_size(const OrcEnv* e) { /* hanabi_env.h:52-59 */
  int size = encoder_shape(e);
  if (e->sad) size += last_action_section_len(e);
  return size;
}
int orc_num_action(const OrcEnv* e) { return max_moves(e) + 1; } /* hanabi_env.h:61-63 */
int orc_hand_size(const OrcEnv* e) { return e->hand_size; }
int orc_players(const OrcEnv* e) { return e->players; }

OrcEnv* orc_env_create(int players, int hand_size, int seed, int bomb, const float* eps_list, int n_eps, int max_len,
                       int sad, int shuffle_color) {
  if (players < 2 || players > ORC_MAX_PLAYERS || hand_size < 1 || hand_size > ORC_MAX_HAND || n_eps < 1 || n_eps > 256) return NULL;
  OrcEnv* e = (OrcEnv*)calloc(1, sizeof(OrcEnv));
  e->players = players; e->hand_size = hand_size; e->seed = seed; e->bomb = bomb;
  mt_seed(&e->rng, (uint32_t)seed); /* hanabi_game.cc:50-53 */
  memcpy(e->eps_list, eps_list, sizeof(float) * (size_t)n_eps);
  e->n_eps = n_eps; e->max_len = max_len; e->sad = sad; e->shuffle_color = shuffle_color;
  e->last_score = -1;
  for (int p = 0; p < ORC_MAX_PLAYERS; ++p)
    for (int c = 0; c < ORC_NUM_COLORS; ++c) { e->color_permute[p][c] = c; e->inv_color_permute[p][c] = c; }
  return e;
}

void orc_env_destroy(OrcEnv* e) { free(e); }

/* Inject the randomness of the NEXT episode: deck order (card index per deal), eps-list index per player and
 * per-player colour permutations (ignored unless shuffle_color).  Disables the mt19937 stream for that reset. */
void orc_env_inject(OrcEnv* e, const int8_t* deck50, const int* eps_idx, const int* perms /* [P][5] or NULL */) {
  e->inject = 1;
  memcpy(e->inj_deck, deck50, ORC_DECK);
  e->inj_pos = 0;
  for (int p = 0; p < e->players; ++p) e->inj_eps_idx[p] = eps_idx[p];
  if (perms != NULL && e->shuffle_color) {
    for (int p = 0; p < e->players; ++p)
      for (int c = 0; c < ORC_NUM_COLORS; ++c) {
        e->color_permute[p][c] = perms[p * ORC_NUM_COLORS + c];
        e->inv_color_permute[p][perms[p * ORC_NUM_COLORS + c]] = c;
      }
  }
}

int orc_env_terminated(OrcEnv* e) { /* hanabi_env.h:79-95 */
  if (!e->state.valid) return 1;
  int term;
  if (e->max_len <= 0) term = state_is_terminal(e, &e->state);
  else term = state_is_terminal(e, &e->state) || e->num_step >= e->max_len;
  if (term) e->last_score = state_score(e, &e->state);
  return term;
}

void orc_env_reset(OrcEnv* e) { /* hanabi_env.cc:9-47 */
  state_init(e, &e->state);
  e->n_dealt = 0;
  while (e->state.cur_player == ORC_CHANCE_PLAYER) apply_random_chance(e, &e->state);
  e->num_step = 0;
  for (int pid = 0; pid < e->players; ++pid) {
    int idx = e->inject ? e->inj_eps_idx[pid] : (int)(mt_next(&e->rng) % (uint32_t)e->n_eps);
    e->eps_idx[pid] = idx;
    e->player_eps[pid] = e->eps_list[idx];
  }
  if (e->shuffle_color && !e->inject) {
    int fix = (int)(mt_next(&e->rng) % (uint32_t)e->players);
    for (int pid = 0; pid < e->players; ++pid) {
      int* perm = e->color_permute[pid];
      int* inv = e->inv_color_permute[pid];
      for (int i = 0; i < ORC_NUM_COLORS; ++i) perm[i] = i;
      if (pid != fix) std_shuffle_int(&e->rng, perm, ORC_NUM_COLORS);
      /* std::sort(inv, by perm[i] < perm[j]) == inverse permutation (hanabi_env.cc:36-38) */
      for (int i = 0; i < ORC_NUM_COLORS; ++i) inv[perm[i]] = i;
    }
  }
  e->clone_valid = 0;
}

/* maybeInversePermuteColor_ (hanabi_env.h:138-146) */
static void maybe_inverse_permute_color(const OrcEnv* e, orc_move* m, int cur) {
  if (e->shuffle_color && m->type == MV_REVEAL_COLOR) m->color = (int8_t)e->inv_color_permute[cur][m->color];
}

/* HanabiEnv::step (hanabi_env.cc:49-113).  a / greedy_a: [players] int64.  Returns 0 ok, -1 illegal move. */
int orc_env_step(OrcEnv* e, const int64_t* a, const int64_t* greedy_a, float* reward, int* terminal) {
  e->num_step += 1;
  float prev_score = (float)state_score(e, &e->state);
  int cur = e->state.cur_player;
  orc_move move = construct_move(e, (int)a[cur]);
  maybe_inverse_permute_color(e, &move, cur);
  if (!move_is_legal(e, &e->state, move)) return -1;
  if (e->sad) {
    e->clone = e->state;
    orc_move gm = construct_move(e, (int)greedy_a[cur]);
    maybe_inverse_permute_color(e, &gm, cur);
    if (!move_is_legal(e, &e->state, gm)) return -1;
    apply_move(e, &e->clone, gm);
    e->clone_valid = 1;
  }
  apply_move(e, &e->state, move);
  int term = state_is_terminal(e, &e->state);
  float r = (float)state_score(e, &e->state) - prev_score;
  if (e->max_len > 0 && e->num_step == e->max_len) { term = 1; r = 0 - prev_score; }
  if (!term)
    while (e->state.cur_player == ORC_CHANCE_PLAYER) apply_random_chance(e, &e->state);
  *reward = r;
  *terminal = term;
  if (e->inject && term) e->inject = 0; /* injected randomness covers exactly one episode */
  return 0;
}

/* computeFeatureAndLegalMove (hanabi_env.cc:115-205): priv_s [P,F], legal_move [P,A], own_hand [P,3H], eps [P] */
void orc_env_observe(const OrcEnv* e, float* priv_s, float* legal_move, float* own_hand, float* eps) {
  const int F = orc_feature_size(e), A = orc_num_action(e), la = last_action_section_len(e);
  const orc_state* s = &e->state;
  const orc_state* cs = e->clone_valid ? &e->clone : &e->state;
  for (int i = 0; i < e->players; ++i) {
    float* v = priv_s + (size_t)i * F;
    encode_obs(e, s, i, v);
    if (e->sad) { /* EncodeLastAction on the greedy clone (canonical_encoders.cc:608-619) */
      float* g = v + encoder_shape(e);
      memset(g, 0, sizeof(float) * (size_t)la);
      encode_last_action(e, cs, i, 0, e->color_permute[i], g);
    }
    encode_own_hand_trinary(e, s, i, own_hand + (size_t)i * e->hand_size * 3);
    float* lm = legal_move + (size_t)i * A;
    memset(lm, 0, sizeof(float) * (size_t)A);
    int n_legal = 0;
    if (i == s->cur_player) { /* LegalMoves (hanabi_state.cc:291-307) */
      for (int uid = 0; uid < max_moves(e); ++uid) {
        orc_move m = construct_move(e, uid);
        if (!move_is_legal(e, s, m)) continue;
        if (e->shuffle_color && m.type == MV_REVEAL_COLOR) m.color = (int8_t)e->color_permute[i][m.color];
        lm[get_move_uid(e, m)] = 1;
        ++n_legal;
      }
    }
    if (n_legal == 0) lm[A - 1] = 1;
    eps[i] = e->player_eps[i];
  }
}

/* ---------------------------------------------------------------- accessors (pybind.cc:15-38 getters) */
int orc_env_cur_player(const OrcEnv* e) { return e->state.cur_player; }
int orc_env_last_score(const OrcEnv* e) { return e->last_score; }
int orc_env_score(const OrcEnv* e) { return state_score(e, &e->state); }
int orc_env_life(const OrcEnv* e) { return e->state.life_tokens; }
int orc_env_info(const OrcEnv* e) { return e->state.information_tokens; }
void orc_env_fireworks(const OrcEnv* e, int* out) { for (int c = 0; c < ORC_NUM_COLORS; ++c) out[c] = e->state.fireworks[c]; }
int orc_env_num_step(const OrcEnv* e) { return e->num_step; }
int orc_env_deck_size(const OrcEnv* e) { return e->state.total_count; }
int orc_env_move_is_legal(const OrcEnv* e, int uid) { return move_is_legal(e, &e->state, construct_move(e, uid)); }
/* randomness actually consumed by the current episode: cards dealt so far (card index), eps idx, perms */
int orc_env_dealt(const OrcEnv* e, int8_t* out) { memcpy(out, e->dealt, (size_t)e->n_dealt); return e->n_dealt; }
void orc_env_eps_idx(const OrcEnv* e, int* out) { for (int p = 0; p < e->players; ++p) out[p] = e->eps_idx[p]; }
void orc_env_perms(const OrcEnv* e, int* out) {
  for (int p = 0; p < e->players; ++p) for (int c = 0; c < ORC_NUM_COLORS; ++c) out[p * ORC_NUM_COLORS + c] = e->color_permute[p][c];
}
uint64_t orc_env_rng_draws(const OrcEnv* e) { return e->rng.draws; }

/* Full 50-card order of the CURRENT episode: the cards dealt so far followed by the cards the mt19937 stream
 * would deal next.  The chance draws depend only on the remaining counts (hanabi_state.cc:316-328), never on the
 * players' moves, so the order is fixed once the episode's reset has run; computed on a COPY of the rng.
 * (In inject mode it is simply the injected deck.)  Lets a test replay the reference's own randomness on the GPU. */
void orc_env_peek_deck(const OrcEnv* e, int8_t* out50) {
  if (e->inject) { memcpy(out50, e->inj_deck, ORC_DECK); return; }
  memcpy(out50, e->dealt, (size_t)e->n_dealt);
  orc_mt19937 g = e->rng;
  int cnt[ORC_NUM_CARDS], total = e->state.total_count, n = e->n_dealt;
  for (int i = 0; i < ORC_NUM_CARDS; ++i) cnt[i] = e->state.card_count[i];
  while (total > 0) {
    int uids[ORC_NUM_CARDS]; double probs[ORC_NUM_CARDS]; int k = 0;
    for (int uid = 0; uid < ORC_NUM_CARDS; ++uid)
      if (cnt[uid] > 0) { uids[k] = uid; probs[k] = (double)cnt[uid] / (double)total; ++k; }
    int idx = uids[discrete_draw(&g, probs, k)];
    out50[n++] = (int8_t)idx; --cnt[idx]; --total;
  }
}

/* ---------------------------------------------------------------- CPU baseline driver ("port" kind)
 * Runs `num_env` envs for `num_steps` steps each with a uniformly random legal policy (LCG), doing exactly
 * what one HanabiVecEnv thread does per tick minus the network: step + full observation encode + auto-reset.
 * Returns the number of env steps executed.  Used only by bench.py's cpu_baseline leg and tests. */
long orc_bench_random_rollout(int players, int hand_size, int sad, int shuffle_color, int max_len, int num_env,
                              int num_steps, int seed, double* checksum) {
  float eps[1] = {0.0f};
  long steps = 0;
  double sum = 0;
  uint32_t x = 12345u + (uint32_t)seed;
  float *priv_s = NULL, *lm = NULL, *oh = NULL, ep[ORC_MAX_PLAYERS];
  for (int g = 0; g < num_env; ++g) {
    OrcEnv* e = orc_env_create(players, hand_size, seed + g, 0, eps, 1, max_len, sad, shuffle_color);
    const int F = orc_feature_size(e), A = orc_num_action(e);
    if (!priv_s) {
      priv_s = (float*)malloc(sizeof(float) * (size_t)(players * F));
      lm = (float*)malloc(sizeof(float) * (size_t)(players * A));
      oh = (float*)malloc(sizeof(float) * (size_t)(players * hand_size * 3));
    }
    orc_env_reset(e);
    orc_env_observe(e, priv_s, lm, oh, ep);
    for (int t = 0; t < num_steps; ++t) {
      if (orc_env_terminated(e)) { orc_env_reset(e); orc_env_observe(e, priv_s, lm, oh, ep); }
      int cur = e->state.cur_player;
      int legal[64], nl = 0;
      for (int u = 0; u < A; ++u) if (lm[cur * A + u] == 1.0f) legal[nl++] = u;
      int64_t a[ORC_MAX_PLAYERS], ga[ORC_MAX_PLAYERS];
      for (int p = 0; p < players; ++p) { a[p] = A - 1; ga[p] = A - 1; }
      x = (1103515245u * x + 12345u) & 0x7fffffffu; a[cur] = legal[(x >> 8) % (uint32_t)nl];
      x = (1103515245u * x + 12345u) & 0x7fffffffu; ga[cur] = legal[(x >> 8) % (uint32_t)nl];
      float r; int term;
      if (orc_env_step(e, a, ga, &r, &term) != 0) abort();
      orc_env_observe(e, priv_s, lm, oh, ep);
      sum += r + priv_s[(t * 7) % F];
      ++steps;
    }
    orc_env_destroy(e);
  }
  free(priv_s); free(lm); free(oh);
  if (checksum) *checksum = sum;
  return steps;
}
