// hb_policy.h -- device-side policy state shared by hb_policy.cu (forward), hb_env_kernels.cu / hb_rollout.cu (the
// encoder writes the GEMM operand directly, the tick zeroes the hidden state of finished games).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "hb_types.h"
#include "hb_head.cuh"

namespace hbg { struct Params; }

struct HbNetWeights {                   // one network (online or target) in GEMM operand form
  __nv_bfloat16 *w0_hi, *w0_lo;         // [512][KS]        net.0.weight, K zero-padded to a multiple of 64
  float* b0;                            // [512]
  __nv_bfloat16 *wl_hi[HB_LAYERS], *wl_lo[HB_LAYERS];  // [2048][1024] = [W_ih | W_hh], rows in gate-interleaved tile order
  float* bl[HB_LAYERS];                 // [2048]           b_ih + b_hh in the same row order
  float *wa, *ba, *wv, *bv;             // fc_a [A][512], [A]; fc_v [512], [1]  (fp32 upload copies; biases are read by the act kernel)
  float* head_tiles;                    // [8 n-tiles][64][hop] fp32: fc_a rows then fc_v (padded to 8), sliced per LSTM output tile (hb_gemm.cuh)
  __nv_bfloat16 *w1_hi, *w1_lo;         // [512][512]       net.2.weight (second fc layer; eval_seats engines only)
  float* b1;                            // [512]
  int has_fc2, skip;                    // architecture variant of this network
  float *raw, *raw2;                    // upload staging
};

struct HbPolicy {
  int rows, rows_pad, KS;
  int parity;                           // state half holding the CURRENT hidden state
  int target_split;                     // 1: the target network also runs bf16x3
  int seat_mode;                        // eval_seats: net index = seat, one problem per seat over that seat's agents
  int n_fc2;                            // seats whose network has a second fc layer (compact problem list of the fc2 launch)
  int have_weights[HB_MAX_P];
  int64_t act_count;                    // forwards so far (Philox counter of the eps-greedy draw)
  HbNetWeights net[HB_MAX_P];           // training: 0 online, 1 target; seat mode: one per seat
  __nv_bfloat16 *s_hi, *s_lo;           // [rows_pad][KS]   priv_s as the fc GEMM operand (written by the encoder)
  __nv_bfloat16 *x_hi[2], *x_lo[2];     // [rows_pad][512]  per network
  __nv_bfloat16 *h_hi[2], *h_lo[2];     // ping-pong halves: [L][rows_pad][512]
  float* c[2];                          // ping-pong halves: [L][rows_pad][512]
  __nv_bfloat16 *th_hi, *th_lo;         // target network layer-0 output
  float* head_part[2];                  // [8 n-tiles][rows_pad][hop] per network (hop = A + 1 padded to 8): partial head sums written by the LSTM-1 epilogue
  float *adv, *oq, *tq;                 // [rows][A], [rows], [rows]
  hbg::Params* d_params;                // [2 parity][4 launches: fc, fc2, lstm0, lstm1][HB_MAX_P problems]
  int* d_error;
  int head_pending;                     // the last forward's head / act step has not run yet (deferred into the next fused tick)
  HbHeadArgs pending_head;
};

int hb_policy_forward(struct hb_engine* e, int greedy_only, int defer_head = 0);
int hb_policy_flush_head(struct hb_engine* e);
