// hb_trainer.cu -- learner side: ONE R2D2 update of pyhanabi/selfplay.py:208-244 on the device, without PyTorch
// in the loop:
//
//   R2D2Agent.loss / td_error (pyhanabi/r2d2.py:383-428, 461-499)   online + target forward over the padded [T, B(, P)] batch,
//                                                                   dueling Q, greedy action, n-step target shift, masked
//                                                                   smooth-L1, |err| priorities, aux cross-entropy (:133-153)
//   loss = (loss * weight).mean(); loss.backward()                  analytic backward of heads, LSTM (hb_lstm.cu), ReLU, fc
//   clip_grad_norm_(online_net.parameters(), grad_clip); Adam.step  fused global-norm + Adam over ONE flat parameter buffer
//   rela.aggregate_priority(priority, seq_len, eta)                 on the device, ready for hb_replay_update_priority
//
// Parameters, gradients and Adam moments are flat fp32 device buffers OWNED BY THE CALLER (so a host framework can view them
// as tensors, all-reduce the gradient bucket between hb_trainer_backward and hb_trainer_optim_step, and save checkpoints);
// hb_trainer_layout gives the offsets of the 16 named tensors of R2D2Net (r2d2.py:22-57).  The dense contractions run on the
// tcgen05 GEMM template (hb_gemm_nt: bf16x3, fp32-class), the recurrences on lstm_fwd_kernel / lstm_bwd_kernel; the kernels
// below are the pointwise / reduction glue.  Everything is queued on the caller's stream; nothing synchronises.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "hb_engine.h"

int hb_gemm_nt_ex(int device, const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, const float* bias, float* C, int64_t ldc,
                  int M, int N, int K, void* stream);
int hb_gemm_nt_same_a(int device, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc, int M, int N, int K,
                      void* stream);   // hb_linear.cu

namespace {

constexpr int HID = 512;
enum { P_FC_W = 0, P_FC_B, P_WIH0, P_WHH0, P_BIH0, P_BHH0, P_WIH1, P_WHH1, P_BIH1, P_BHH1, P_FCV_W, P_FCV_B, P_FCA_W, P_FCA_B, P_PRED_W, P_PRED_B, P_N };

void layout(int F, int A, int H, int64_t* off) {
  const int64_t sz[P_N] = {(int64_t)HID * F, HID, 4LL * HID * HID, 4LL * HID * HID, 4 * HID, 4 * HID, 4LL * HID * HID, 4LL * HID * HID, 4 * HID, 4 * HID,
                           HID, 1, (int64_t)A * HID, A, (int64_t)3 * H * HID, 3 * H};
  int64_t o = 0;
  for (int i = 0; i < P_N; ++i) { off[i] = o; o += (sz[i] + 3) / 4 * 4; }   // 16-byte aligned starts
  off[P_N] = o;
}

// ---- head weights packed as one matrix: rows [0, A) fc_a, row A fc_v, rows [A+1, A+1+3H) pred
__global__ void ht_pack_heads(const float* __restrict__ wa, const float* __restrict__ ba, const float* __restrict__ wv, const float* __restrict__ bv,
                              const float* __restrict__ wp, const float* __restrict__ bp, int A, int HO, float* __restrict__ Wh, float* __restrict__ WhT,
                              float* __restrict__ bh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HO * HID) {
    const int o = i / HID, k = i - o * HID;
    const float v = o < A ? wa[(size_t)o * HID + k] : (o == A ? wv[k] : wp[(size_t)(o - A - 1) * HID + k]);
    Wh[i] = v;
    if (WhT) WhT[(size_t)k * HO + o] = v;
  }
  if (i < HO) bh[i] = i < A ? ba[i] : (i == A ? bv[0] : bp[i - A - 1]);
}

__global__ void ht_relu(float* __restrict__ x, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<float4*>(x)[i];
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  reinterpret_cast<float4*>(x)[i] = v;
}
__global__ void ht_relu_bwd(const float* __restrict__ x, float* __restrict__ dx, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = reinterpret_cast<const float4*>(x)[i];
  float4 g = reinterpret_cast<float4*>(dx)[i];
  g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
  reinterpret_cast<float4*>(dx)[i] = g;
}

// One warp per batch row r = (t, b, p): dueling Q of both networks (r2d2.py:124-131), Q_online(s, a), the online greedy
// action (first-index argmax over legal moves, :109-111) and Q_target(s, greedy) (:398-401).
__global__ void ht_q_rows(const float* __restrict__ y_on, const float* __restrict__ y_tg, int HO, const float* __restrict__ legal, const int64_t* __restrict__ act,
                          int A, long long n_rows, float* __restrict__ qa_on, float* __restrict__ qa_tg) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const unsigned FULL = 0xffffffffu;
  const float* yo = y_on + r * HO;
  const float* yt = y_tg + r * HO;
  const float* lm = legal + r * A;
  const float l0 = lane < A ? lm[lane] : 0.f, l1 = lane + 32 < A ? lm[lane + 32] : 0.f;
  const float a0 = lane < A ? yo[lane] * l0 : 0.f, a1 = lane + 32 < A ? yo[lane + 32] * l1 : 0.f;
  const float t0 = lane < A ? yt[lane] * l0 : 0.f, t1 = lane + 32 < A ? yt[lane + 32] * l1 : 0.f;
  float so = a0 + a1, st = t0 + t1;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { so += __shfl_xor_sync(FULL, so, k); st += __shfl_xor_sync(FULL, st, k); }
  const float vo = yo[A], vt = yt[A];
  // q = v + legal_a - mean(legal_a); greedy = argmax over legal of q (== argmax of legal_a), first index on ties; no legal
  // move (padding rows) -> index 0, like argmax over an all-zero row
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (l0 != 0.f) { best = a0; bi = lane; }
  if (l1 != 0.f && (bi == 0x7fffffff || a1 > best)) { best = a1; bi = lane + 32; }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    const float ob = __shfl_xor_sync(FULL, best, k);
    const int oi = __shfl_xor_sync(FULL, bi, k);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
  }
  const int greedy = bi == 0x7fffffff ? 0 : bi;
  const int a = (int)act[r];
  const float qa = a < 32 ? __shfl_sync(FULL, a0, a) : __shfl_sync(FULL, a1, a - 32);
  const float tq = greedy < 32 ? __shfl_sync(FULL, t0, greedy) : __shfl_sync(FULL, t1, greedy - 32);
  if (lane == 0) {
    qa_on[r] = vo + qa - so / (float)A;
    qa_tg[r] = vt + tq - st / (float)A;
  }
}

struct LossArgs {
  int T, B, P, A, H, HO, n_step;   // T = steps computed (t_eff)
  float gamma_n, eta, pred_weight;
  float Bnorm;                     // entries of the WHOLE batch the mean runs over (>= B when this pass is one micro-batch of it)
  const float *qa_on, *qa_tg;      // [T][B][P]
  const float *y_on;               // [T*B*P][HO]
  const float *legal;              // [T][B][P][A]
  const int64_t* act;              // [T][B][P]
  const float *own_hand;           // [T][B][P][3H]
  const float *reward, *bootstrap; // [T][B]
  const float *seq_len, *weight;   // [B]
  float* dy;                       // [T*B*P][HO]
  float* priority;                 // [B] aggregated (rela.aggregate_priority)
  float* stats;                    // [8]: 0 loss, 1 rl_loss / seq_len mean, 2 aux xent / seq_len mean
};

// One CTA per batch entry b, one thread per step t: TD error (r2d2.py:403-428), smooth-L1 (:473-476), priorities, the aux
// cross-entropy (:133-153, 430-459) and dLoss/d(head outputs) of the online network for loss = mean_b(weight_b * loss_b).
__global__ void ht_loss(LossArgs L) {
  __shared__ float red[3][32];
  const int b = blockIdx.x, t = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const int T = L.T, B = L.B, P = L.P, A = L.A, HO = L.HO;
  const float len = L.seq_len[b], wb = L.weight[b];
  float loss_t = 0.f, prio = 0.f, xent_t = 0.f;
  if (t < T) {
    const size_t tb = (size_t)t * B + b;
    float on = 0.f, tg = 0.f;
    for (int p = 0; p < P; ++p) on += L.qa_on[tb * P + p];
    if (t + L.n_step < T) {
      const size_t tb2 = (size_t)(t + L.n_step) * B + b;
      for (int p = 0; p < P; ++p) tg += L.qa_tg[tb2 * P + p];
    }
    const float mask = (float)t < len ? 1.f : 0.f;
    const float target = L.reward[tb] + L.bootstrap[tb] * L.gamma_n * tg;
    const float err = (target - on) * mask;
    const float ae = fabsf(err);
    loss_t = ae < 1.f ? 0.5f * err * err : ae - 0.5f;
    prio = ae;
    // d loss / d qa_sum = smooth_l1'(err) * d err / d qa = clamp(err, -1, 1) * (-mask), scaled by weight_b / (batch entries)
    const float g = -fminf(fmaxf(err, -1.f), 1.f) * mask * wb / L.Bnorm;
    for (int p = 0; p < P; ++p) {
      const size_t r = tb * P + p;
      float* dy = L.dy + r * HO;
      const float* lm = L.legal + r * A;
      const int a = (int)L.act[r];
      for (int j = 0; j < A; ++j) dy[j] = g * lm[j] * ((j == a ? 1.f : 0.f) - 1.f / (float)A);
      dy[A] = g;
      // aux task: cross entropy of softmax(pred(lstm_o)) against the own-hand trinary target, slots averaged, players averaged
      const float* y = L.y_on + r * HO + A + 1;
      const float* oh = L.own_hand + r * 3 * L.H;
      if (L.pred_weight > 0.f) {
        float msum = 0.f;
        for (int s = 0; s < L.H; ++s) msum += oh[3 * s] + oh[3 * s + 1] + oh[3 * s + 2];
        const float M = fmaxf(msum, 1e-6f);
        const float coef = L.pred_weight * wb / (L.Bnorm * (float)P);
        for (int s = 0; s < L.H; ++s) {
          const float z0 = y[3 * s], z1 = y[3 * s + 1], z2 = y[3 * s + 2];
          const float mx = fmaxf(z0, fmaxf(z1, z2));
          const float e0 = expf(z0 - mx), e1 = expf(z1 - mx), e2 = expf(z2 - mx);
          const float lse = mx + logf(e0 + e1 + e2), inv = 1.f / (e0 + e1 + e2);
          const float p0 = oh[3 * s], p1 = oh[3 * s + 1], p2 = oh[3 * s + 2];
          const float ms = p0 + p1 + p2;                       // slot mask = own_hand.sum(-1)
          const float plogq = p0 * (z0 - lse) + p1 * (z1 - lse) + p2 * (z2 - lse);
          xent_t += -(plogq * ms) / M / (float)P;
          const float c = coef * ms / M;                       // d xent / d z_k = -(ms / M) (p_k - q_k sum_p)
          dy[A + 1 + 3 * s] = -c * (p0 - e0 * inv * ms);
          dy[A + 1 + 3 * s + 1] = -c * (p1 - e1 * inv * ms);
          dy[A + 1 + 3 * s + 2] = -c * (p2 - e2 * inv * ms);
        }
      } else {
        for (int j = A + 1; j < HO; ++j) dy[j] = 0.f;
      }
    }
  }
  float s0 = loss_t, s1 = prio, s2 = xent_t, mx = prio;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, k); s1 += __shfl_xor_sync(0xffffffffu, s1, k); s2 += __shfl_xor_sync(0xffffffffu, s2, k);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
  }
  __shared__ float redm[32];
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; redm[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float rl = 0.f, ps = 0.f, xe = 0.f, pm = 0.f;
    for (int w = 0; w < nw; ++w) { rl += red[0][w]; ps += red[1][w]; xe += red[2][w]; pm = fmaxf(pm, redm[w]); }
    L.priority[b] = L.eta * pm + (1.f - L.eta) * ps / len;     // aggregate_priority (r2d2_actor.h:10-21)
    const float total = rl + L.pred_weight * xe;
    atomicAdd(&L.stats[0], total * wb / L.Bnorm);
    atomicAdd(&L.stats[1], rl / len / L.Bnorm);
    atomicAdd(&L.stats[2], xe / len / L.Bnorm);
  }
}

// out[c] = sum_r m[r][c] in two deterministic passes.  Pass 1: grid (ceil(cols / 32), HT_COLSUM_SLICES), 256 threads: a lane
// owns a COLUMN, a warp reads 128 contiguous bytes of a row (one thread per column walking down the rows would touch a
// different sector with every lane), the CTA's 8 warps take every 8th row of the slice; part[slice][c].  Pass 2 adds the slices.
#define HT_COLSUM_SLICES 64
__global__ void ht_colsum_part(const float* __restrict__ m, long long rows, int cols, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < cols)
    for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) s += m[r * cols + c];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    part[(size_t)blockIdx.y * cols + c] = t;
  }
}
__global__ void ht_colsum_final(const float* __restrict__ part, int slices, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = 0.f;
  for (int s = 0; s < slices; ++s) t += part[(size_t)s * cols + c];
  out[c] = t;
}

__global__ void ht_unpack_head_grads(const float* __restrict__ dWh, const float* __restrict__ dbh, int A, int HO, int use_pred, float* __restrict__ g_wa,
                                     float* __restrict__ g_ba, float* __restrict__ g_wv, float* __restrict__ g_bv, float* __restrict__ g_wp, float* __restrict__ g_bp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HO * HID) {
    const int o = i / HID, k = i - o * HID;
    const float v = dWh[i];
    if (o < A) g_wa[(size_t)o * HID + k] = v;
    else if (o == A) g_wv[k] = v;
    else g_wp[(size_t)(o - A - 1) * HID + k] = use_pred ? v : 0.f;
  }
  if (i < HO) {
    const float v = dbh[i];
    if (i < A) g_ba[i] = v;
    else if (i == A) g_bv[0] = v;
    else g_bp[i - A - 1] = use_pred ? v : 0.f;
  }
}

// sum of squares of the flat gradient -> stats[3] (atomic over blocks)
__global__ void ht_sumsq(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))) + torch.optim.Adam.step (no weight decay, no amsgrad)
__global__ void ht_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n4, const float* __restrict__ sumsq,
                        float grad_clip, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, long long skip_lo4, long long skip_hi4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4 || (i >= skip_lo4 && i < skip_hi4)) return;   // [skip_lo, skip_hi): parameters without a gradient this run (pred head, pred_weight = 0)
  const float norm = sqrtf(*sumsq);
  const float coef = grad_clip > 0.f ? fminf(1.f, grad_clip / (norm + 1e-6f)) : 1.f;
  float4 P4 = reinterpret_cast<float4*>(p)[i], M4 = reinterpret_cast<float4*>(m)[i], V4 = reinterpret_cast<float4*>(v)[i];
  const float4 G4 = reinterpret_cast<const float4*>(g)[i];
  float* pp = reinterpret_cast<float*>(&P4); float* mm = reinterpret_cast<float*>(&M4); float* vv = reinterpret_cast<float*>(&V4);
  const float* gg = reinterpret_cast<const float*>(&G4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gr = gg[k] * coef;
    mm[k] = beta1 * mm[k] + (1.f - beta1) * gr;
    vv[k] = beta2 * vv[k] + (1.f - beta2) * gr * gr;
    const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
    pp[k] -= (lr / bc1) * (mm[k] / denom);
  }
  reinterpret_cast<float4*>(p)[i] = P4; reinterpret_cast<float4*>(m)[i] = M4; reinterpret_cast<float4*>(v)[i] = V4;
}

// g += acc (gradient accumulation over the micro-batches of one update)
__global__ void ht_add(float* __restrict__ g, const float* __restrict__ acc, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(g)[i];
  const float4 b = reinterpret_cast<const float4*>(acc)[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  reinterpret_cast<float4*>(g)[i] = a;
}

}  // namespace

struct hb_trainer {
  hb_trainer_config cfg;
  int device, HO, rows_max;
  int64_t off[P_N + 1];
  float *params[2], *grads, *adam_m, *adam_v;   // caller-owned flat buffers
  hb_lstm* lstm;
  float *x[2], *o[2], *y[2];          // fc output, lstm output, head output of (online, target)
  float *qa[2];                       // [T*rows]
  float *dy, *dO, *dX;
  float *Wh[2], *WhT, *bh[2], *dWh, *dbh;
  float* colsum_part;                 // [HT_COLSUM_SLICES][512] partial column sums
  float* acc;                         // gradients of the earlier micro-batches of this update (allocated on first use)
  float* stats;                       // device [8]
  float* h_stats;                     // pinned [8]
  cudaEvent_t ev_stats;
  int stats_pending;
  hb_train_stats last;                // statistics of the most recent update known to be complete (hb_trainer_stats_nowait)
  int64_t step;                       // Adam step count
  int last_use_pred;
  int64_t launches;
};

#define HT_ALLOC(ptr, n)                                                        \
  do {                                                                          \
    HB_CUDA(cudaMalloc((void**)&(ptr), (size_t)(n) * sizeof(float)));          \
    HB_CUDA(cudaMemset((ptr), 0, (size_t)(n) * sizeof(float)));                \
  } while (0)

extern "C" {

int hb_trainer_layout(int in_dim, int num_action, int hand_size, int64_t* offsets) {
  if (!offsets || in_dim < 1 || num_action < 1 || hand_size < 1) { hb_set_error("hb_trainer_layout: bad argument"); return -1; }
  layout(in_dim, num_action, hand_size, offsets);
  return 0;
}

int hb_trainer_create(const hb_trainer_config* cfg, float* online, float* target, float* grads, float* adam_m, float* adam_v, hb_trainer** out) {
  if (!cfg || !online || !target || !grads || !adam_m || !adam_v || !out) { hb_set_error("hb_trainer_create: null argument"); return -1; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { hb_set_error("hb_trainer_create: no CUDA device -- libhanabi_b200 has no CPU path"); return -2; }
  if (cfg->device < 0 || cfg->device >= ndev) { hb_set_error("hb_trainer_create: bad device ordinal"); return -1; }
  const int P = cfg->num_player, rows = cfg->max_batch * P;
  if (cfg->num_action > 63 || cfg->num_action < 2 || P < 1 || P > 5 || cfg->seq_len < 1 || cfg->max_batch < 1 || cfg->multi_step < 1) {
    hb_set_error("hb_trainer_create: need 2 <= num_action <= 63, 1 <= num_player <= 5, seq_len, max_batch, multi_step >= 1");
    return -1;
  }
  if (rows > 256) {
    hb_set_error("hb_trainer_create: max_batch * num_player = %d rows; one pass of the LSTM kernels holds 256 -- split the batch into micro-batches of <= %d "
                 "entries and accumulate (hb_trainer_backward_ex)", rows, 256 / P);
    return -1;
  }
  if (cfg->max_batch > 1024 || cfg->seq_len > 1024) { hb_set_error("hb_trainer_create: max_batch and seq_len must be <= 1024"); return -1; }
  HB_CUDA(cudaSetDevice(cfg->device));
  hb_trainer* tr = new hb_trainer();
  memset(tr, 0, sizeof(*tr));
  tr->cfg = *cfg; tr->device = cfg->device;
  tr->HO = cfg->num_action + 1 + 3 * cfg->hand_size;
  tr->rows_max = rows;
  layout(cfg->in_dim, cfg->num_action, cfg->hand_size, tr->off);
  tr->params[0] = online; tr->params[1] = target; tr->grads = grads; tr->adam_m = adam_m; tr->adam_v = adam_v;
  int rc = hb_lstm_create(cfg->device, cfg->seq_len, rows, &tr->lstm);
  if (rc) { delete tr; return rc; }
  const size_t N = (size_t)cfg->seq_len * rows, HO = tr->HO;
  for (int n = 0; n < 2; ++n) {
    HT_ALLOC(tr->x[n], N * HID); HT_ALLOC(tr->o[n], N * HID); HT_ALLOC(tr->y[n], N * HO); HT_ALLOC(tr->qa[n], N);
    HT_ALLOC(tr->Wh[n], HO * HID); HT_ALLOC(tr->bh[n], HO);
  }
  HT_ALLOC(tr->dy, N * HO); HT_ALLOC(tr->dO, N * HID); HT_ALLOC(tr->dX, N * HID);
  HT_ALLOC(tr->WhT, HO * HID); HT_ALLOC(tr->dWh, HO * HID); HT_ALLOC(tr->dbh, HO);
  HT_ALLOC(tr->colsum_part, (size_t)HT_COLSUM_SLICES * HID);
  HT_ALLOC(tr->stats, 8);
  HB_CUDA(cudaMallocHost((void**)&tr->h_stats, 8 * sizeof(float)));
  memset(tr->h_stats, 0, 8 * sizeof(float));
  HB_CUDA(cudaEventCreateWithFlags(&tr->ev_stats, cudaEventDisableTiming));
  *out = tr;
  return 0;
}

void hb_trainer_destroy(hb_trainer* tr) {
  if (!tr) return;
  cudaSetDevice(tr->device);
  cudaDeviceSynchronize();
  hb_lstm_destroy(tr->lstm);
  for (int n = 0; n < 2; ++n) { cudaFree(tr->x[n]); cudaFree(tr->o[n]); cudaFree(tr->y[n]); cudaFree(tr->qa[n]); cudaFree(tr->Wh[n]); cudaFree(tr->bh[n]); }
  cudaFree(tr->dy); cudaFree(tr->dO); cudaFree(tr->dX);
  cudaFree(tr->WhT); cudaFree(tr->dWh); cudaFree(tr->dbh); cudaFree(tr->stats); cudaFree(tr->acc); cudaFree(tr->colsum_part);
  cudaFreeHost(tr->h_stats); cudaEventDestroy(tr->ev_stats);
  delete tr;
}

// Forward of both networks, loss, priorities and the full backward of the online network: gradients land in the flat
// `grads` buffer, aggregated priorities in `priority` (device float [batchsize]).  `total_batch` = entries of the whole batch
// the loss mean runs over; `accumulate` != 0 adds this pass's gradients and loss statistics to those already there: a batch
// with more than 256 LSTM rows (3-5 player VDN at the reference's batch sizes) is fed as micro-batches of <= 256 / num_player
// entries, the first with accumulate = 0.  The rows of a batch are independent up to the mean, so this is exact.
int hb_trainer_backward_ex(hb_trainer* tr, const hb_batch* b, int batchsize, int t_eff, float pred_weight, float* priority, int total_batch, int accumulate,
                           void* stream) {
  if (!tr || !b || !priority) { hb_set_error("hb_trainer_backward: null argument"); return -1; }
  const hb_trainer_config& c = tr->cfg;
  if (batchsize < 1 || batchsize > c.max_batch) { hb_set_error("hb_trainer_backward: batchsize must be 1..%d", c.max_batch); return -1; }
  if (total_batch < batchsize) { hb_set_error("hb_trainer_backward: total_batch %d < batchsize %d", total_batch, batchsize); return -1; }
  if (!b->priv_s || !b->legal_move || !b->a || !b->reward || !b->bootstrap || !b->seq_len || !b->weight || (pred_weight > 0.f && !b->own_hand)) {
    hb_set_error("hb_trainer_backward: the batch lacks a tensor the loss needs");
    return -1;
  }
  const int T = t_eff < 1 ? 1 : (t_eff > c.seq_len ? c.seq_len : t_eff);   // steps >= the longest episode are padding in every row
  const int P = c.num_player, A = c.num_action, F = c.in_dim, HO = tr->HO, B = batchsize, rows = B * P;
  const long long N = (long long)T * rows;
  HB_CUDA(cudaSetDevice(tr->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t* off = tr->off;
  auto blocks = [](long long n, int t) { return (unsigned)((n + t - 1) / t); };
  int rc;
  const size_t grad_bytes = (size_t)off[P_N] * sizeof(float);
  if (accumulate) {
    if (!tr->acc) HB_CUDA(cudaMalloc((void**)&tr->acc, grad_bytes));
    HB_CUDA(cudaMemcpyAsync(tr->acc, tr->grads, grad_bytes, cudaMemcpyDeviceToDevice, st));
  } else {
    HB_CUDA(cudaMemsetAsync(tr->stats, 0, 8 * sizeof(float), st));
  }
  // ---- forward: fc + ReLU, LSTM (both networks in one pass), heads
  for (int n = 0; n < 2; ++n) {
    const float* p = tr->params[n];
    // the target network's fc layer reads the same observations: their bf16 split is made once
    rc = n == 0 ? hb_gemm_nt(tr->device, b->priv_s, F, p + off[P_FC_W], F, p + off[P_FC_B], tr->x[n], HID, (int)N, HID, F, st)
                : hb_gemm_nt_same_a(tr->device, b->priv_s, F, p + off[P_FC_W], F, p + off[P_FC_B], tr->x[n], HID, (int)N, HID, F, st);
    if (rc) return rc;
    ht_relu<<<blocks(N * HID / 4, 256), 256, 0, st>>>(tr->x[n], N * HID / 4);
    ht_pack_heads<<<blocks((long long)HO * HID, 256), 256, 0, st>>>(p + off[P_FCA_W], p + off[P_FCA_B], p + off[P_FCV_W], p + off[P_FCV_B], p + off[P_PRED_W],
                                                                   p + off[P_PRED_B], A, HO, tr->Wh[n], n == 0 ? tr->WhT : nullptr, tr->bh[n]);
  }
  hb_lstm_weights lw[2];
  for (int n = 0; n < 2; ++n) {
    const float* p = tr->params[n];
    lw[n].w_ih[0] = p + off[P_WIH0]; lw[n].w_hh[0] = p + off[P_WHH0]; lw[n].b_ih[0] = p + off[P_BIH0]; lw[n].b_hh[0] = p + off[P_BHH0];
    lw[n].w_ih[1] = p + off[P_WIH1]; lw[n].w_hh[1] = p + off[P_WHH1]; lw[n].b_ih[1] = p + off[P_BIH1]; lw[n].b_hh[1] = p + off[P_BHH1];
  }
  const float* xs[2] = {tr->x[0], tr->x[1]};
  float* ys[2] = {tr->o[0], tr->o[1]};
  rc = hb_lstm_forward(tr->lstm, T, rows, 2, xs, lw, ys, 1, st);
  if (rc) return rc;
  for (int n = 0; n < 2; ++n) {
    rc = hb_gemm_nt(tr->device, tr->o[n], HID, tr->Wh[n], HID, tr->bh[n], tr->y[n], HO, (int)N, HO, HID, st);
    if (rc) return rc;
  }
  // ---- TD error, loss, priorities, d loss / d heads
  ht_q_rows<<<blocks(N, 8), 256, 0, st>>>(tr->y[0], tr->y[1], HO, b->legal_move, b->a, A, N, tr->qa[0], tr->qa[1]);
  LossArgs L;
  L.T = T; L.B = B; L.P = P; L.A = A; L.H = c.hand_size; L.HO = HO; L.n_step = c.multi_step;
  double gn = 1.0;
  for (int i = 0; i < c.multi_step; ++i) gn *= (double)c.gamma;
  L.gamma_n = (float)gn; L.eta = c.eta; L.pred_weight = pred_weight; L.Bnorm = (float)total_batch;
  L.qa_on = tr->qa[0]; L.qa_tg = tr->qa[1]; L.y_on = tr->y[0]; L.legal = b->legal_move; L.act = b->a; L.own_hand = b->own_hand;
  L.reward = b->reward; L.bootstrap = b->bootstrap; L.seq_len = b->seq_len; L.weight = b->weight;
  L.dy = tr->dy; L.priority = priority; L.stats = tr->stats;
  ht_loss<<<B, (T + 31) / 32 * 32, 0, st>>>(L);
  // ---- backward: heads
  rc = hb_gemm_nt(tr->device, tr->dy, HO, tr->WhT, HO, nullptr, tr->dO, HID, (int)N, HID, HO, st);                  // dO = dY Wh
  if (rc) return rc;
  // dWh = dY^T O: both operands are given with the contraction axis (batch rows) as their SLOW axis -> transposing splits
  rc = hb_gemm_nt_ex(tr->device, tr->dy, HO, 1, tr->o[0], HID, 1, nullptr, tr->dWh, HID, HO, HID, (int)N, st);
  if (rc) return rc;
  ht_colsum_part<<<dim3((HO + 31) / 32, HT_COLSUM_SLICES), 256, 0, st>>>(tr->dy, N, HO, tr->colsum_part);
  ht_colsum_final<<<(HO + 127) / 128, 128, 0, st>>>(tr->colsum_part, HT_COLSUM_SLICES, HO, tr->dbh);
  float* g = tr->grads;
  const int use_pred = pred_weight > 0.f ? 1 : 0;
  ht_unpack_head_grads<<<blocks((long long)HO * HID, 256), 256, 0, st>>>(tr->dWh, tr->dbh, A, HO, use_pred, g + off[P_FCA_W], g + off[P_FCA_B], g + off[P_FCV_W],
                                                                        g + off[P_FCV_B], g + off[P_PRED_W], g + off[P_PRED_B]);
  // ---- backward: LSTM (recurrences + dX / dW GEMMs), ReLU, fc
  hb_lstm_grads lg;
  lg.dw_ih[0] = g + off[P_WIH0]; lg.dw_hh[0] = g + off[P_WHH0]; lg.db_ih[0] = g + off[P_BIH0]; lg.db_hh[0] = g + off[P_BHH0];
  lg.dw_ih[1] = g + off[P_WIH1]; lg.dw_hh[1] = g + off[P_WHH1]; lg.db_ih[1] = g + off[P_BIH1]; lg.db_hh[1] = g + off[P_BHH1];
  rc = hb_lstm_backward(tr->lstm, tr->dO, tr->dX, &lg, st);
  if (rc) return rc;
  ht_relu_bwd<<<blocks(N * HID / 4, 256), 256, 0, st>>>(tr->x[0], tr->dX, N * HID / 4);
  ht_colsum_part<<<dim3(HID / 32, HT_COLSUM_SLICES), 256, 0, st>>>(tr->dX, N, HID, tr->colsum_part);
  ht_colsum_final<<<(HID + 127) / 128, 128, 0, st>>>(tr->colsum_part, HT_COLSUM_SLICES, HID, g + off[P_FC_B]);
  rc = hb_gemm_nt_ex(tr->device, tr->dX, HID, 1, b->priv_s, F, 1, nullptr, g + off[P_FC_W], F, HID, F, (int)N, st);   // dW0 = dXpre^T S
  if (rc) return rc;
  if (accumulate) ht_add<<<blocks(off[P_N] / 4, 256), 256, 0, st>>>(g, tr->acc, off[P_N] / 4);
  HB_CUDA(cudaGetLastError());
  tr->last_use_pred = use_pred;
  tr->launches += 20 + (accumulate ? 1 : 0);
  return 0;
}

int hb_trainer_backward(hb_trainer* tr, const hb_batch* b, int batchsize, int t_eff, float pred_weight, float* priority, void* stream) {
  return hb_trainer_backward_ex(tr, b, batchsize, t_eff, pred_weight, priority, batchsize, 0, stream);
}

int hb_trainer_optim_step(hb_trainer* tr, void* stream) {
  if (!tr) { hb_set_error("hb_trainer_optim_step: null trainer"); return -1; }
  HB_CUDA(cudaSetDevice(tr->device));
  cudaStream_t st = (cudaStream_t)stream;
  const hb_trainer_config& c = tr->cfg;
  const long long n4 = tr->off[P_N] / 4;
  HB_CUDA(cudaMemsetAsync(tr->stats + 3, 0, sizeof(float), st));
  ht_sumsq<<<296, 256, 0, st>>>(tr->grads, n4, tr->stats + 3);
  tr->step += 1;
  const double bc1 = 1.0 - pow((double)c.beta1, (double)tr->step), bc2 = 1.0 - pow((double)c.beta2, (double)tr->step);
  // without the aux task the pred head has no gradient: torch's Adam skips such parameters entirely
  const long long skip_lo = tr->last_use_pred ? n4 : tr->off[P_PRED_W] / 4, skip_hi = tr->last_use_pred ? n4 : n4;
  ht_adam<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(tr->params[0], tr->grads, tr->adam_m, tr->adam_v, n4, tr->stats + 3, c.grad_clip, c.lr, c.beta1, c.beta2,
                                                        c.adam_eps, (float)bc1, (float)sqrt(bc2), skip_lo, skip_hi);
  HB_CUDA(cudaGetLastError());
  HB_CUDA(cudaMemcpyAsync(tr->h_stats, tr->stats, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaEventRecord(tr->ev_stats, st));
  tr->stats_pending = 1;
  tr->launches += 2;
  return 0;
}

// Statistics of the last completed update (waits for it): loss (weighted mean, what selfplay.py feeds to stat["loss"]),
// rl_loss / seq_len mean (stat["rl_loss"]), aux cross-entropy / seq_len mean (stat["aux1"]), gradient norm before clipping.
int hb_trainer_stats(hb_trainer* tr, hb_train_stats* out) {
  if (!tr || !out) { hb_set_error("hb_trainer_stats: null argument"); return -1; }
  HB_CUDA(cudaSetDevice(tr->device));
  if (tr->stats_pending) { HB_CUDA(cudaEventSynchronize(tr->ev_stats)); tr->stats_pending = 0; }
  out->loss = tr->h_stats[0]; out->rl_loss = tr->h_stats[1]; out->aux_xent = tr->h_stats[2]; out->grad_norm = sqrtf(tr->h_stats[3]);
  out->num_update = tr->step; out->launches = tr->launches + hb_lstm_launches(tr->lstm);
  tr->last = *out;
  return hb_lstm_sync(tr->lstm);   // also reports a spin-guard failure of the recurrence kernels
}

// The same without waiting: statistics of the most recent update that HAS completed (out->num_update says which; 0 = none
// yet).  A training loop that logs every iteration stays asynchronous with this (the values lag by an update or two).
int hb_trainer_stats_nowait(hb_trainer* tr, hb_train_stats* out) {
  if (!tr || !out) { hb_set_error("hb_trainer_stats_nowait: null argument"); return -1; }
  HB_CUDA(cudaSetDevice(tr->device));
  if (tr->stats_pending) {
    const cudaError_t q = cudaEventQuery(tr->ev_stats);
    if (q == cudaSuccess) {
      tr->stats_pending = 0;
      tr->last.loss = tr->h_stats[0]; tr->last.rl_loss = tr->h_stats[1]; tr->last.aux_xent = tr->h_stats[2]; tr->last.grad_norm = sqrtf(tr->h_stats[3]);
      tr->last.num_update = tr->step;
    } else if (q != cudaErrorNotReady) {
      HB_CUDA(q);
    }
  }
  *out = tr->last;
  out->launches = tr->launches + hb_lstm_launches(tr->lstm);
  return 0;
}

// R2D2Agent.sync_target_with_online (r2d2.py:208-210)
int hb_trainer_sync_target(hb_trainer* tr, void* stream) {
  if (!tr) { hb_set_error("hb_trainer_sync_target: null trainer"); return -1; }
  HB_CUDA(cudaSetDevice(tr->device));
  HB_CUDA(cudaMemcpyAsync(tr->params[1], tr->params[0], (size_t)tr->off[P_N] * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
