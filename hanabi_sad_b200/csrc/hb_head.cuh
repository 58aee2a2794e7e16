// hb_head.cuh -- the tail of the R2D2 act forward for ONE agent row, executed by one warp: finish the dueling head (sum of the
// 8 per-tile partials the LSTM-1 epilogue left + bias), masked first-index argmax (r2d2.py:242-243), eps-greedy (:273-277),
// Q_online(s, a) and Q_target(s, greedy) for compute_priority (:344-348).  Shared by the stand-alone kernel hb_k_head_act
// (hb_policy.cu) and the fused tick (hb_rollout.cu), which runs it as the prologue of the NEXT tick.
#pragma once
#include "hb_env_cta.cuh"

struct HbHeadArgs {
  int rows, rows_pad, A, have_target;
  int seat_mode, P;           // seat mode: the network of row r is net r % P (no target network)
  const float* part[2];       // [8][rows_pad][hop] partial head sums of the online / target network (LSTM-1 epilogue); hop = A + 1 padded to 8
  const float* ba[HB_MAX_P];  // fc_a bias [A] per network (training: 0 online, 1 target)
  const float* bv[HB_MAX_P];  // fc_v bias [1]
  const float* legal;         // [rows][A]
  const float* eps;           // [rows]
  int64_t* a;                 // [rows]
  int64_t* greedy_a;          // [rows]
  float* adv;                 // [rows][A] online advantages
  float* oq;                  // [rows] online dueling Q of the chosen action          (r2d2.py:344, :124-131)
  float* tq;                  // [rows] target dueling Q of the online greedy action   (r2d2.py:345-348)
  uint64_t seed;
  const unsigned long long* tick_ctr;  // device tick counter (Philox counter of the eps-greedy draw); may be null
  uint32_t tick;
  int greedy_only;            // eval actors: eps ignored
};


// Lane l owns outputs l and l+32.  All 32 lanes of the calling warp must be active.
__device__ __forceinline__ void hb_head_row(const HbHeadArgs& p, int row, int lane) {
  const int A = p.A, HO = (A + 1 + 7) & ~7;   // row stride of the partials
  const unsigned FULL = 0xffffffffu;
  const int nets = p.have_target ? 2 : 1;
  float out[2][2], vv[2];
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    out[net][0] = out[net][1] = vv[net] = 0.f;
    if (net >= nets) continue;
    float s0 = 0.f, s1 = 0.f, sv = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float* pr = p.part[net] + ((size_t)t * p.rows_pad + row) * HO;
      if (lane <= A) s0 += pr[lane];          // lane A (or A - 32 below) carries the value head's partial sums
      if (lane + 32 <= A) s1 += pr[lane + 32];
    }
    sv = A < 32 ? __shfl_sync(FULL, s0, A) : __shfl_sync(FULL, s1, A - 32);   // same sum, same order, no extra loads
    const int wn = p.seat_mode ? row % p.P : net;  // whose biases
    out[net][0] = lane < A ? s0 + __ldg(p.ba[wn] + lane) : 0.f;
    out[net][1] = lane + 32 < A ? s1 + __ldg(p.ba[wn] + lane + 32) : 0.f;
    vv[net] = sv + __ldg(p.bv[wn]);
  }
  const float* lm = p.legal + (size_t)row * A;
  const float l0 = lane < A ? lm[lane] : 0.f;
  const float l1 = lane + 32 < A ? lm[lane + 32] : 0.f;
  const float o0 = out[0][0], o1 = out[0][1];
  // mean over ALL A entries of adv*legal (r2d2.py:129-130)
  float s = o0 * l0 + o1 * l1;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(FULL, s, k);
  const float mean = s / (float)A;
  if (lane < A) p.adv[(size_t)row * A + lane] = o0;
  if (lane + 32 < A) p.adv[(size_t)row * A + lane + 32] = o1;
  // greedy = first index of the largest advantage among legal moves (r2d2.py:242-243)
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (l0 != 0.f) { best = o0; bi = lane; }
  if (l1 != 0.f && (o1 > best || bi == 0x7fffffff)) { best = o1; bi = lane + 32; }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    const float ob = __shfl_xor_sync(FULL, best, k);
    const int oi = __shfl_xor_sync(FULL, bi, k);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
  }
  const int greedy = bi == 0x7fffffff ? A - 1 : bi;
  // eps-greedy: uniform over legal moves with probability eps (r2d2.py:273-277)
  int action = greedy;
  if (!p.greedy_only) {
    const unsigned m0 = __ballot_sync(FULL, l0 != 0.f), m1 = __ballot_sync(FULL, l1 != 0.f);
    const int n_legal = __popc(m0) + __popc(m1);
    const uint32_t tick = p.tick_ctr ? (uint32_t)*p.tick_ctr : p.tick;
    HbRng rng(p.seed, (uint32_t)row, tick, HB_RNG_ACT);
    const float u = rng.uniform();
    int k = n_legal > 0 ? (int)rng.below((uint32_t)n_legal) : 0;
    if (u < p.eps[row] && n_legal > 0) {
      if (k < __popc(m0)) { unsigned m = m0; for (int i = 0; i < k; ++i) m &= m - 1; action = __ffs(m) - 1; }
      else { k -= __popc(m0); unsigned m = m1; for (int i = 0; i < k; ++i) m &= m - 1; action = 32 + __ffs(m) - 1; }
    }
  }
  // Q_online(s, a) for the action actually taken; Q_target(s, greedy) under the target network
  const float qa = action < 32 ? __shfl_sync(FULL, o0 * l0, action) : __shfl_sync(FULL, o1 * l1, action - 32);
  float tq = 0.f;
  if (nets == 2) {
    const float t0 = out[1][0], t1 = out[1][1];
    float ts = t0 * l0 + t1 * l1;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) ts += __shfl_xor_sync(FULL, ts, k);
    const float tqa = greedy < 32 ? __shfl_sync(FULL, t0 * l0, greedy) : __shfl_sync(FULL, t1 * l1, greedy - 32);
    tq = vv[1] + tqa - ts / (float)A;
  }
  if (lane == 0) {
    p.a[row] = action;
    p.greedy_a[row] = greedy;
    p.oq[row] = vv[0] + qa - mean;
    if (nets == 2) p.tq[row] = tq;
  }
}

