// hb_types.h -- board-state record, static game geometry and feature layout shared by the host driver
// and the sm_100a kernels.  Nothing here is derived from code: the layouts restate WHAT the reference
// encoder emits (hanabi_lib/canonical_encoders.cc:70-581) as closed-form offsets so that one GPU
// thread can evaluate any single feature directly from a 256-byte game record.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

#define HB_MAX_P 5
#define HB_MAX_H 5
#define HB_NC 5            // colours
#define HB_NR 5            // ranks
#define HB_NCARD 25        // card types (colour-major: c*5+r)
#define HB_DECK 50
#define HB_MAX_INFO 8
#define HB_MAX_LIFE 3
#define HB_CHANCE 0xFF     // cur_player while a deal is pending (reference kChancePlayerId = -1)
#define HB_NO_CARD 0xFF
#define HB_HID 512         // LSTM hidden size served by the fused kernels (reference --rnn_hid_dim 512)
#define HB_LAYERS 2        // reference num_lstm_layer default (r2d2.py / selfplay.py:60)

// move types, numerically equal to the reference enum (hanabi_move.h:35)
enum { HB_MV_INVALID = 0, HB_MV_PLAY = 1, HB_MV_DISCARD = 2, HB_MV_REVEAL_COLOR = 3, HB_MV_REVEAL_RANK = 4 };

// What the encoder needs to know about "the last non-deal move" (canonical_encoders.cc:293-422).
struct HbLastMove {
  uint8_t valid;          // 0: no player move yet in this episode
  uint8_t player;         // absolute seat of the actor
  uint8_t type;           // HB_MV_*
  uint8_t target_offset;  // hints: 1..P-1 relative to the actor
  uint8_t color;          // hinted colour (real, un-permuted)
  uint8_t rank;           // hinted rank
  uint8_t card_index;     // play/discard: hand slot
  uint8_t reveal_mask;    // hints: bit i = target's slot i matched
  uint8_t card_color;     // play/discard: identity of the card
  uint8_t card_rank;
  uint8_t scored;         // play: landed on the fireworks
  uint8_t info_token;     // play: completed a stack and regained a token
};

// One game = one 256-byte record (16 coalesced 16-byte loads by half a warp).  Everything the rules
// (hanabi_state.cc:169-388) and the encoder read; the discard PILE is kept as per-card-type counts because
// its order is never observed (canonical_encoders.cc:262-276, 803-806).
struct __attribute__((aligned(16))) HbGame {
  uint8_t hand_card[HB_MAX_P][HB_MAX_H];  // card id c*5+r, oldest first; HB_NO_CARD beyond hand_len
  uint8_t hand_len[HB_MAX_P];
  uint16_t know[HB_MAX_P][HB_MAX_H];      // bits 0-4 colour-plausible, 5-9 rank-plausible, 10-12 hinted colour (7=none), 13-15 hinted rank (7=none)
  uint8_t discard_count[HB_NCARD];
  uint8_t fireworks[HB_NC];
  uint8_t info, life;
  uint8_t cur_player, next_player;        // cur_player == HB_CHANCE while a deal is pending (only visible at terminal states)
  int8_t turns_to_play;
  uint8_t deck_pos;                       // cards dealt so far (deck size = 50 - deck_pos)
  uint16_t num_step;
  HbLastMove last;                        // real last move
  HbLastMove greedy;                      // SAD: the greedy move "as if applied" to the pre-move state (hanabi_env.cc:82-91)
  uint8_t eps_idx[HB_MAX_P];
  uint8_t terminated;                     // HanabiEnv::terminated() (hanabi_env.h:79-95); 1 before the first reset
  uint16_t perm[HB_MAX_P];                // colour permutation per observer, 3 bits per colour (real -> shown)
  uint16_t inv_perm[HB_MAX_P];            // shown -> real
  int16_t last_score;
  uint8_t greedy_valid;                   // 0 right after reset: SAD block encodes `last` (hanabi_env.cc:46)
  uint8_t illegal;                        // sticky: an illegal action reached hb_step (reference aborts, hanabi_env.cc:63-80)
  uint32_t episode;                       // episodes started on this seat (Philox stream id)
  float reward;                           // reward of the last step
  int16_t ep_len;                         // steps recorded in the staging episode
  uint8_t pad[256 - 182];
};
static_assert(sizeof(HbGame) == 256, "HbGame must be 256 bytes");

#define HB_KNOW_BLANK 0xFFFFu

HB_HD int hb_perm_get(uint16_t p, int c) { return (p >> (3 * c)) & 7; }
HB_HD uint16_t hb_perm_identity() { return (uint16_t)(0 | (1 << 3) | (2 << 6) | (3 << 9) | (4 << 12)); }

// Static geometry for a (players, hand_size, sad) configuration.
struct HbGeom {
  int P, H, sad;
  int A;            // num_action = 2H + 10(P-1) + 1 (hanabi_env.h:61-63); uid A-1 is the no-op
  int F;            // feature_size (hanabi_env.h:52-59)
  int la_len;       // last-action block 2P + 2H + 41 (canonical_encoders.cc:585-595)
  int off_board, off_discard, off_last, off_belief, off_sad;  // section starts inside priv_s
  int deck_bits;    // 50 - P*H
};

HB_HD HbGeom hb_make_geom(int P, int H, int sad) {
  HbGeom g;
  g.P = P; g.H = H; g.sad = sad;
  g.A = 2 * H + 2 * HB_NC * (P - 1) + 1;
  g.la_len = 2 * P + 2 * H + 41;
  g.deck_bits = HB_DECK - P * H;
  g.off_board = P * H * HB_NCARD + P;
  g.off_discard = g.off_board + g.deck_bits + HB_NCARD + HB_MAX_INFO + HB_MAX_LIFE;
  g.off_last = g.off_discard + HB_DECK;
  g.off_belief = g.off_last + g.la_len;
  g.off_sad = g.off_belief + P * H * 35;
  g.F = g.off_sad + (sad ? g.la_len : 0);
  return g;
}
