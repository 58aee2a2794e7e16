// hb_replay.cu -- host side and learner-facing kernels of the device replay (hb_replay.h): allocation of the
// episode ring (board records, not observations: hb_replay.h), stratified priority sampling + importance weights (PrioritizedReplay::sample_,
// rela/prioritized_replay.h:274-345), batch assembly in the learner's [T, B, ...] layout with terminal padding
// (RNNTransition::makeBatch, rela/transition.cc:160-202; FFTransition::padLike, :29-40) and priority write-back
// (ConcurrentQueue::update, prioritized_replay.h:106-120).  The producer side (append / finalize / commit) is part of
// the fused tick in hb_rollout.cu.
#include <string.h>

#include "hb_engine.h"
#include "hb_env_cta.cuh"
#include "hb_replay.h"

#define HB_SCAN_THREADS 1024

// First arrival number still held: replay_block -> what sample() popped so far (ConcurrentQueue::blockPop); free-running
// ring -> everything but the `cap_slots` most recent commits.
__device__ __forceinline__ long long hb_ring_oldest(const HbRing& R) {
  const long long commits = (long long)R.counters[HB_CNT_COMMIT];
  if (R.block) return (long long)R.counters[HB_CNT_POPPED];
  return commits > R.cap_slots ? commits - R.cap_slots : 0;
}

__device__ __forceinline__ float hb_entry_weight(const HbRing& R, int i, long long oldest) {
  const int slot = i / R.NE;
  if (R.state[slot] != HB_SLOT_COMMITTED) return 0.f;
  if (R.commit_seq[slot] < oldest) return 0.f;  // beyond `capacity` most recent episodes: evicted (prioritized_replay.h:329-332)
  return R.weight[i];
}

// Inclusive prefix sums (double, like the reference's running accSum) of the sampleable weights over ALL ring entries, in
// three small launches so that the reference's default 131 072-episode buffer (selfplay.py --replay_buffer_size) is scanned by
// the whole GPU instead of one SM: (1) per-block sums of HB_SCAN_BLOCK entries, (2) scan of the block sums by one CTA (also
// out[0] = total weight, out[1] = number of sampleable entries), (3) per-block scan with the block's offset.
#define HB_SCAN_BLOCK 2048   // entries per CTA of passes 1 and 3 (256 threads x 8)

__global__ void __launch_bounds__(256) hb_k_replay_blocksum(HbRing R, double* __restrict__ bsum, int* __restrict__ bcnt) {
  __shared__ double sw[8];
  __shared__ int sc[8];
  const int n = R.phys_slots * R.NE, base = blockIdx.x * HB_SCAN_BLOCK;
  const long long oldest = hb_ring_oldest(R);
  double s = 0;
  int c = 0;
  for (int i = base + threadIdx.x; i < min(n, base + HB_SCAN_BLOCK); i += 256) { const float w = hb_entry_weight(R, i, oldest); s += w; c += w > 0.f; }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, k); c += __shfl_xor_sync(0xffffffffu, c, k); }
  if ((threadIdx.x & 31) == 0) { sw[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; int tc = 0;
    for (int w = 0; w < 8; ++w) { t += sw[w]; tc += sc[w]; }
    bsum[blockIdx.x] = t; bcnt[blockIdx.x] = tc;
  }
}

// exclusive scan of the block sums in place (bsum[b] = weight before block b); one CTA, sequential per thread chunk
__global__ void __launch_bounds__(HB_SCAN_THREADS) hb_k_replay_blockscan(double* __restrict__ bsum, const int* __restrict__ bcnt, int nb, double* __restrict__ out) {
  __shared__ double part[HB_SCAN_THREADS];
  __shared__ int cnt[HB_SCAN_THREADS];
  const int tid = threadIdx.x, chunk = (nb + HB_SCAN_THREADS - 1) / HB_SCAN_THREADS;
  const int lo = tid * chunk, hi = min(nb, lo + chunk);
  double s = 0; int c = 0;
  for (int i = lo; i < hi; ++i) { s += bsum[i]; c += bcnt[i]; }
  part[tid] = s; cnt[tid] = c;
  __syncthreads();
  for (int off = 1; off < HB_SCAN_THREADS; off <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
    double v = 0; int cv = 0;
    if (tid >= off) { v = part[tid - off]; cv = cnt[tid - off]; }
    __syncthreads();
    part[tid] += v; cnt[tid] += cv;
    __syncthreads();
  }
  double acc = tid > 0 ? part[tid - 1] : 0.0;
  for (int i = lo; i < hi; ++i) { const double v = bsum[i]; bsum[i] = acc; acc += v; }
  if (tid == HB_SCAN_THREADS - 1) { out[0] = part[tid]; out[1] = (double)cnt[tid]; }
}

__global__ void __launch_bounds__(256) hb_k_replay_prefix(HbRing R, const double* __restrict__ bsum, double* __restrict__ prefix) {
  __shared__ double part[256];
  const int n = R.phys_slots * R.NE, base = blockIdx.x * HB_SCAN_BLOCK, tid = threadIdx.x;
  const long long oldest = hb_ring_oldest(R);
  const int lo = base + tid * (HB_SCAN_BLOCK / 256), hi = min(n, lo + HB_SCAN_BLOCK / 256);
  float w[HB_SCAN_BLOCK / 256];
  double s = 0;
#pragma unroll
  for (int j = 0; j < HB_SCAN_BLOCK / 256; ++j) { w[j] = lo + j < hi ? hb_entry_weight(R, lo + j, oldest) : 0.f; s += w[j]; }
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    double v = 0;
    if (tid >= off) v = part[tid - off];
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  double acc = bsum[blockIdx.x] + (tid > 0 ? part[tid - 1] : 0.0);
#pragma unroll
  for (int j = 0; j < HB_SCAN_BLOCK / 256; ++j) if (lo + j < hi) { acc += w[j]; prefix[lo + j] = acc; }
}

// the three passes, queued on the engine stream; `tot` (device double[2]) receives the total weight and the entry count
static int hb_launch_prefix(hb_engine* e, double* tot) {
  HbReplay* Q = e->replay;
  const HbRing& R = Q->ring;
  const int n = R.phys_slots * R.NE, nb = (n + HB_SCAN_BLOCK - 1) / HB_SCAN_BLOCK;
  hb_k_replay_blocksum<<<nb, 256, 0, e->stream>>>(R, Q->bsum, Q->bcnt);
  hb_k_replay_blockscan<<<1, HB_SCAN_THREADS, 0, e->stream>>>(Q->bsum, Q->bcnt, nb, tot);
  hb_k_replay_prefix<<<nb, 256, 0, e->stream>>>(R, Q->bsum, Q->prefix);
  HB_CUDA(cudaGetLastError());
  e->launches += 3;
  return 0;
}

// One thread per batch element: stratified draw, binary search, importance weight.  `targets` (device, may be null):
// draw positions inside [0, sum) supplied by the caller (a replay sharded over engines splits ONE global stratified draw);
// norm_sum / norm_size: the sum_w and N of the importance weight (<= 0: this shard's); normalize: divide by the batch max.
__global__ void hb_k_replay_draw(HbRing R, const double* __restrict__ prefix, const double* __restrict__ tot, int B, float beta, uint64_t seed,
                                 unsigned long long draw, const double* __restrict__ targets, double norm_sum, double norm_size, int normalize,
                                 int* __restrict__ idx_out, long long* __restrict__ seq_out, float* __restrict__ w_out,
                                 float* __restrict__ is_weight, int* __restrict__ max_len_out) {
  __shared__ float red[32];
  __shared__ int max_len;
  if (threadIdx.x == 0) max_len = 0;
  __syncthreads();
  const int b = threadIdx.x, n = R.phys_slots * R.NE;
  const float sum = (float)tot[0];
  const float size = (float)tot[1];
  float isw = 0.f;
  if (b < B) {
    float r;
    if (targets != nullptr) {
      r = (float)targets[b];
    } else {
      const float segment = sum / (float)B;
      HbRng rng(seed, (uint32_t)b, (uint32_t)draw, HB_RNG_SAMPLE);
      r = rng.uniform() * segment + (float)b * segment;       // dist(rng_) + i * segment (prioritized_replay.h:296)
    }
    r = fminf(sum - 0.1f, r);
    int lo = 0, hi = n - 1;                                   // first i with prefix[i] > 0 and prefix[i] >= r
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const double p = prefix[mid];
      if (p > 0.0 && p >= (double)r) hi = mid; else lo = mid + 1;
    }
    const float w = (float)(prefix[lo] - (lo > 0 ? prefix[lo - 1] : 0.0));
    idx_out[b] = lo;
    seq_out[b] = R.commit_seq[lo / R.NE];
    atomicMax(&max_len, R.seq_len[lo / R.NE]);
    w_out[b] = w;
    const float ns = norm_sum > 0.0 ? (float)norm_sum : sum, nn = norm_size > 0.0 ? (float)norm_size : size;
    isw = powf(nn * (w / ns), -beta);                         // prioritized_replay.h:337-338
  }
  float m = isw;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    if (threadIdx.x == 0) red[0] = m;
  }
  __syncthreads();
  if (b < B) is_weight[b] = normalize ? isw / red[0] : isw;   // weights /= weights.max()
  if (b == 0) *max_len_out = max_len;
  if (b == 0 && R.block) {   // "pop storage if full" (prioritized_replay.h:326-332): evict down to `capacity`, AFTER the draw
    const long long commits = (long long)R.counters[HB_CNT_COMMIT], popped = (long long)R.counters[HB_CNT_POPPED];
    if (commits - popped > R.cap_slots) R.counters[HB_CNT_POPPED] = (unsigned long long)(commits - R.cap_slots);
  }
}

struct HbBatchPtrs {
  float* priv_s; float* legal; float* own_hand; float* eps; int64_t* a; int64_t* greedy_a;
  float* reward; float* bootstrap; uint8_t* terminal; float* seq_len;
};

// grid (B, T): one CTA re-encodes step t of sampled entry b from its stored board record into the [T, B, (P,) ...]
// batch (the same hb_feature code the actors' observations came from: bit-exact); steps at or beyond the episode length
// are the reference's padding (zeros, terminal = 1).
__global__ void __launch_bounds__(128) hb_k_replay_gather(HbRing R, const int* __restrict__ idx, int B, HbBatchPtrs out, HbEnvCfg cfg,
                                                          const float* __restrict__ eps_list) {
  __shared__ HbGame s;
  __shared__ HbEncTables tab;
  const int b = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const int entry = idx[b];
  if (entry < 0) return;                       // hb_replay_get: lookup failed, reported by the host
  const int slot = entry / R.NE, e = entry % R.NE;
  const int PP = R.NE == 1 ? R.P : 1;          // players kept in one entry
  const int p0 = R.NE == 1 ? 0 : e;
  const int len = R.seq_len[slot];
  const bool live = t < len;
  const size_t src = ((size_t)slot * R.T + t) * R.P + p0;   // (slot, t, p0) in units of one player's row
  const size_t dst = ((size_t)t * B + b) * PP;
  const int nF = PP * R.F, nA = PP * R.A, nO = PP * R.OH;
  if (live) {
    if (tid < 16) reinterpret_cast<uint4*>(&s)[tid] = reinterpret_cast<const uint4*>(R.states + (size_t)slot * R.T + t)[tid];
    __syncthreads();
    hb_cta_build_tables(s, tab, cfg.g);
    __syncthreads();
    hb_cta_build_totals(s, tab, cfg.g);
    __syncthreads();
    hb_cta_write_obs(s, tab, cfg, out.priv_s + dst * R.F, out.legal + dst * R.A, out.own_hand + dst * R.OH, out.eps + dst, eps_list, nullptr, nullptr, 0,
                     p0, PP);
  } else {
    for (int i = tid; i < nF; i += blockDim.x) out.priv_s[dst * R.F + i] = 0.f;
    for (int i = tid; i < nA; i += blockDim.x) out.legal[dst * R.A + i] = 0.f;
    for (int i = tid; i < nO; i += blockDim.x) out.own_hand[dst * R.OH + i] = 0.f;
    if (tid < PP) out.eps[dst + tid] = 0.f;
  }
  if (tid < PP) {
    out.a[dst + tid] = live ? R.a[src + tid] : 0;
    out.greedy_a[dst + tid] = live ? R.greedy_a[src + tid] : 0;
  }
  if (tid == 0) {
    out.reward[(size_t)t * B + b] = live ? R.reward[(size_t)slot * R.T + t] : 0.f;
    out.bootstrap[(size_t)t * B + b] = live ? R.bootstrap[(size_t)slot * R.T + t] : 0.f;
    out.terminal[(size_t)t * B + b] = t >= len - 1 ? 1 : 0;
    if (t == 0) out.seq_len[b] = (float)len;
  }
}

__global__ void hb_k_replay_update(HbRing R, const int* __restrict__ idx, const long long* __restrict__ seq, const float* __restrict__ prio, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int entry = idx[i], slot = entry / R.NE;
  if (R.state[slot] == HB_SLOT_COMMITTED && R.commit_seq[slot] == seq[i]) R.weight[entry] = powf(prio[i], R.alpha);
}

// ConcurrentQueue::get (prioritized_replay.h:125-128): the idx-th oldest entry still held = arrival number
// `oldest + idx / NE`; one thread per slot looks for it (arrival numbers are unique).
__global__ void hb_k_replay_find(HbRing R, long long idx, int* __restrict__ entry_out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= R.phys_slots) return;
  const long long oldest = hb_ring_oldest(R);
  if (R.state[slot] == HB_SLOT_COMMITTED && R.commit_seq[slot] == oldest + idx / R.NE) entry_out[0] = slot * R.NE + (int)(idx % R.NE);
}

HbRing hb_replay_ring(hb_engine* e) { return e->replay->ring; }

#define HB_RALLOC(ptr, bytes)                                        \
  do {                                                               \
    HB_CUDA(cudaMalloc((void**)&(ptr), (bytes)));                    \
    HB_CUDA(cudaMemsetAsync((ptr), 0, (bytes), e->stream));          \
  } while (0)

extern "C" {

int hb_replay_create(hb_engine* e) {
  e->replay = nullptr;
  const hb_config& c = e->cfg;
  if (c.replay_capacity <= 0) return 0;
  if (!e->policy) { hb_set_error("hb_create: a replay needs the device policy (hid_dim = 512)"); return -1; }
  if (c.seq_len < 1 || c.multi_step < 1) { hb_set_error("hb_create: seq_len and multi_step must be >= 1"); return -1; }
  if (c.max_len <= 0 || c.max_len > c.seq_len) {
    hb_set_error("hb_create: with a replay, 0 < max_len <= seq_len is required (the reference asserts it, transition_buffer.h:143)");
    return -1;
  }
  HbReplay* Q = new HbReplay();
  memset(Q, 0, sizeof(*Q));
  e->replay = Q;
  HbRing& R = Q->ring;
  R.T = c.seq_len; R.P = e->P; R.F = e->F; R.A = e->A; R.OH = 3 * e->H;
  R.NE = c.vdn ? 1 : e->P;
  Q->capacity = c.replay_capacity;
  R.cap_slots = (c.replay_capacity + R.NE - 1) / R.NE;
  R.block = c.replay_block ? 1 : 0;
  R.limit_slots = ((int)(1.25 * c.replay_capacity) + R.NE - 1) / R.NE;   // storage_(int(1.25 * capacity)), prioritized_replay.h:183
  if (R.limit_slots < R.cap_slots + 1) R.limit_slots = R.cap_slots + 1;
  // free-running ring: the `cap_slots` newest commits + one episode in flight per game + slack for commit-order skew;
  // replay_block: up to limit_slots held + one in flight per game (a game claims only after its commit went through)
  R.phys_slots = R.block ? R.limit_slots + e->G + 64 : R.cap_slots + 2 * e->G + 64;
  R.n_step = c.multi_step; R.gamma = c.gamma; R.eta = c.eta; R.alpha = c.alpha;
  double gn = 1.0;
  for (int i = 0; i < c.multi_step; ++i) gn *= (double)c.gamma;  // python: self.gamma ** self.multi_step
  R.gamma_n = (float)gn;
  R.uniform_priority = c.priority_mode == 1;
  Q->beta = c.beta; Q->seed = c.seed ^ 0x5851F42D4C957F2DULL;
  const size_t S = R.phys_slots, T = R.T, P = R.P, G = e->G;
  HB_RALLOC(R.states, S * T * sizeof(HbGame));
  HB_RALLOC(R.a, S * T * P * sizeof(int64_t));
  HB_RALLOC(R.greedy_a, S * T * P * sizeof(int64_t));
  HB_RALLOC(R.reward, S * T * sizeof(float));
  HB_RALLOC(R.bootstrap, S * T * sizeof(float));
  HB_RALLOC(R.seq_len, S * sizeof(int));
  HB_RALLOC(R.weight, S * R.NE * sizeof(float));
  HB_RALLOC(R.commit_seq, S * sizeof(long long));
  HB_RALLOC(R.state, S * sizeof(int));
  HB_CUDA(cudaMalloc((void**)&R.game_slot, G * sizeof(int)));
  HB_CUDA(cudaMemsetAsync(R.game_slot, 0xFF, G * sizeof(int), e->stream));  // -1: no slot yet
  HB_RALLOC(R.sc_reward, G * T * sizeof(float));
  HB_RALLOC(R.sc_oq, G * T * P * sizeof(float));
  HB_RALLOC(R.sc_tq, G * T * P * sizeof(float));
  HB_RALLOC(R.counters, HB_CNT_N * sizeof(unsigned long long));
  Q->max_batch = 1024;
  HB_RALLOC(Q->prefix, (S * R.NE + 2) * sizeof(double));
  HB_RALLOC(Q->bsum, ((S * R.NE + HB_SCAN_BLOCK - 1) / HB_SCAN_BLOCK + 1) * sizeof(double));
  HB_RALLOC(Q->bcnt, ((S * R.NE + HB_SCAN_BLOCK - 1) / HB_SCAN_BLOCK + 1) * sizeof(int));
  HB_RALLOC(Q->sampled_idx, HB_SAMPLE_SETS * Q->max_batch * sizeof(int));
  HB_RALLOC(Q->sampled_seq, HB_SAMPLE_SETS * Q->max_batch * sizeof(long long));
  HB_RALLOC(Q->sampled_w, HB_SAMPLE_SETS * Q->max_batch * sizeof(float));
  HB_RALLOC(Q->d_entry, sizeof(int));
  HB_RALLOC(Q->d_prio, Q->max_batch * sizeof(float));
  HB_RALLOC(Q->d_targets, Q->max_batch * sizeof(double));
  HB_RALLOC(Q->d_max_len, HB_SAMPLE_SETS * sizeof(int));
  HB_CUDA(cudaMallocHost((void**)&Q->h_max_len, HB_SAMPLE_SETS * sizeof(int)));
  for (int i = 0; i < HB_SAMPLE_SETS; ++i) {
    Q->h_max_len[i] = 0;
    HB_CUDA(cudaEventCreateWithFlags(&Q->set_ev[i], cudaEventDisableTiming));
  }
  HB_CUDA(cudaMallocHost((void**)&Q->h_counters, (HB_CNT_N + 2) * sizeof(unsigned long long)));
  return 0;
}

void hb_replay_destroy(hb_engine* e) {
  HbReplay* Q = e->replay;
  if (!Q) return;
  HbRing& R = Q->ring;
  cudaFree(R.states); cudaFree(R.a); cudaFree(R.greedy_a); cudaFree(R.reward);
  cudaFree(R.bootstrap); cudaFree(R.seq_len); cudaFree(R.weight); cudaFree(R.commit_seq); cudaFree(R.state); cudaFree(R.game_slot);
  cudaFree(R.sc_reward); cudaFree(R.sc_oq); cudaFree(R.sc_tq); cudaFree(R.counters);
  cudaFree(Q->prefix); cudaFree(Q->bsum); cudaFree(Q->bcnt); cudaFree(Q->sampled_idx); cudaFree(Q->sampled_seq); cudaFree(Q->sampled_w); cudaFree(Q->d_prio); cudaFree(Q->d_targets); cudaFree(Q->d_max_len); cudaFreeHost(Q->h_max_len); cudaFree(Q->d_entry);
  for (int i = 0; i < HB_SAMPLE_SETS; ++i) if (Q->set_ev[i]) cudaEventDestroy(Q->set_ev[i]);
  cudaFreeHost(Q->h_counters);
  delete Q;
  e->replay = nullptr;
}

static int hb_read_counters(hb_engine* e) {   // synchronising copy of the ring counters into the pinned mirror
  HbReplay* Q = e->replay;
  HB_CUDA(cudaSetDevice(e->device));
  HB_CUDA(cudaMemcpyAsync(Q->h_counters, Q->ring.counters, HB_CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}
static long long hb_held_slots(const HbReplay* Q) {
  const long long commits = (long long)Q->h_counters[HB_CNT_COMMIT];
  if (Q->ring.block) return commits - (long long)Q->h_counters[HB_CNT_POPPED];
  return commits < Q->ring.cap_slots ? commits : Q->ring.cap_slots;
}

int hb_counters(hb_engine* e, int64_t* size, int64_t* num_add, int64_t* num_act) {
  if (!e) { hb_set_error("hb_counters: null engine"); return -1; }
  if (num_act) *num_act = e->num_act;
  if (size) *size = 0;
  if (num_add) *num_add = 0;
  HbReplay* Q = e->replay;
  if (!Q) return 0;
  int rc = hb_read_counters(e);
  if (rc) return rc;
  if (size) *size = hb_held_slots(Q) * Q->ring.NE;
  if (num_add) *num_add = (long long)Q->h_counters[HB_CNT_COMMIT] * Q->ring.NE;
  // a game waiting in blockAppend does not act (the reference's thread sits in cvSize_.wait): its ticks do not count
  if (num_act) *num_act = e->num_act - (int64_t)Q->h_counters[HB_CNT_STALLED];
  if (e->policy) {
    rc = hb_status_post(e);
    if (rc) return rc;
    return hb_status_poll(e, true);
  }
  return 0;
}

int hb_replay_stats(hb_engine* e, hb_replay_info* out) {
  if (!e || !out) { hb_set_error("hb_replay_stats: null argument"); return -1; }
  memset(out, 0, sizeof(*out));
  out->num_act = e->num_act;
  HbReplay* Q = e->replay;
  if (!Q) return 0;
  HbRing& R = Q->ring;
  double* tot = Q->prefix + (size_t)R.phys_slots * R.NE;
  HB_CUDA(cudaSetDevice(e->device));
  { const int prc = hb_launch_prefix(e, tot); if (prc) return prc; }
  double h_tot[2] = {0, 0};
  HB_CUDA(cudaMemcpyAsync(h_tot, tot, sizeof(h_tot), cudaMemcpyDeviceToHost, e->stream));
  int rc = hb_read_counters(e);
  if (rc) return rc;
  out->size = hb_held_slots(Q) * R.NE;
  out->num_add = (long long)Q->h_counters[HB_CNT_COMMIT] * R.NE;
  out->num_act = e->num_act - (int64_t)Q->h_counters[HB_CNT_STALLED];
  out->dropped = (int64_t)Q->h_counters[HB_CNT_DROPPED];
  out->stalled_ticks = (int64_t)Q->h_counters[HB_CNT_STALLED];
  out->popped = R.block ? (int64_t)Q->h_counters[HB_CNT_POPPED] * R.NE : (out->num_add - out->size);
  out->capacity = Q->capacity;
  out->phys_slots = R.phys_slots;
  out->weight_sum = h_tot[0];
  out->sampleable = (int64_t)h_tot[1];
  return 0;
}

// Queue the draw + gather of one batch into the next free id set (no host wait).
// `fresh_counters` = 0: judge the replay's size by the mirror of the last counter read (it only grows, or stays at capacity)
// and read it synchronously only if that says "too small" -- a prefetch must not make the host wait for the engine stream.
static int replay_enqueue(hb_engine* e, int batchsize, const hb_batch* out, const hb_sample_opts* opts, const char* who, bool fresh_counters) {
  HbReplay* Q = e->replay;
  if (batchsize < 1 || batchsize > Q->max_batch) { hb_set_error("%s: batchsize must be 1..%d", who, Q->max_batch); return -1; }
  int rc = 0;
  if (fresh_counters || hb_held_slots(Q) * Q->ring.NE < batchsize) rc = hb_read_counters(e);
  if (rc) return rc;
  const int64_t size = hb_held_slots(Q) * Q->ring.NE;
  if (size < batchsize) { hb_set_error("%s: replay holds %lld entries, fewer than the batch size %d", who, (long long)size, batchsize); return -3; }
  HbRing& R = Q->ring;
  const int set = (Q->set_head + Q->set_count) % HB_SAMPLE_SETS;
  int* s_idx = Q->sampled_idx + (size_t)set * Q->max_batch;
  long long* s_seq = Q->sampled_seq + (size_t)set * Q->max_batch;
  float* s_w = Q->sampled_w + (size_t)set * Q->max_batch;
  double* tot = Q->prefix + (size_t)R.phys_slots * R.NE;
  const double* d_targets = nullptr;
  if (opts && opts->targets) {
    HB_CUDA(cudaMemcpyAsync(Q->d_targets, opts->targets, batchsize * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    d_targets = Q->d_targets;
  }
  rc = hb_launch_prefix(e, tot);
  if (rc) return rc;
  const int threads = (batchsize + 31) / 32 * 32;
  hb_k_replay_draw<<<1, threads, 0, e->stream>>>(R, Q->prefix, tot, batchsize, Q->beta, Q->seed, Q->sample_count, d_targets,
                                                 opts ? opts->total_weight : 0.0, opts ? opts->total_size : 0.0, opts ? opts->normalize : 1,
                                                 s_idx, s_seq, s_w, out->weight, Q->d_max_len + set);
  HB_CUDA(cudaMemcpyAsync(Q->h_max_len + set, Q->d_max_len + set, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HbBatchPtrs bp = {out->priv_s, out->legal_move, out->own_hand, out->eps, out->a, out->greedy_a, out->reward, out->bootstrap, out->terminal, out->seq_len};
  hb_k_replay_gather<<<dim3(batchsize, R.T), 128, 0, e->stream>>>(R, s_idx, batchsize, bp, e->env, e->d_eps_list);
  HB_CUDA(cudaGetLastError());
  if (out->ids) HB_CUDA(cudaMemcpyAsync(out->ids, s_idx, batchsize * sizeof(int), cudaMemcpyDeviceToDevice, e->stream));
  HB_CUDA(cudaEventRecord(Q->set_ev[set], e->stream));
  e->launches += 2;
  Q->sample_count += 1;
  Q->set_n[set] = batchsize;
  Q->set_waited[set] = 0;
  Q->set_count += 1;
  return 0;
}

int hb_replay_sample_ex(hb_engine* e, int batchsize, const hb_batch* out, const hb_sample_opts* opts) {
  if (!e || !out) { hb_set_error("hb_replay_sample: null argument"); return -1; }
  HbReplay* Q = e->replay;
  if (!Q) { hb_set_error("hb_replay_sample: this engine has no replay (replay_capacity = 0)"); return -1; }
  if (Q->set_count != 0) {  // prioritized_replay.h:209-212
    hb_set_error("hb_replay_sample: previous samples' priority has not been updated");
    return -3;
  }
  HB_CUDA(cudaSetDevice(e->device));
  int rc = replay_enqueue(e, batchsize, out, opts, "hb_replay_sample", true);
  if (rc) return rc;
  return hb_replay_take(e, nullptr);   // the batch tensors are consumed on the caller's own stream
}

int hb_replay_sample(hb_engine* e, int batchsize, const hb_batch* out) { return hb_replay_sample_ex(e, batchsize, out, nullptr); }

int hb_replay_prefetch(hb_engine* e, int batchsize, const hb_batch* out, const hb_sample_opts* opts) {
  if (!e || !out) { hb_set_error("hb_replay_prefetch: null argument"); return -1; }
  HbReplay* Q = e->replay;
  if (!Q) { hb_set_error("hb_replay_prefetch: this engine has no replay (replay_capacity = 0)"); return -1; }
  if (Q->set_count >= HB_SAMPLE_SETS) { hb_set_error("hb_replay_prefetch: %d batches are outstanding already (update their priorities first)", Q->set_count); return -3; }
  HB_CUDA(cudaSetDevice(e->device));
  return replay_enqueue(e, batchsize, out, opts, "hb_replay_prefetch", false);
}

int hb_replay_take(hb_engine* e, int* batchsize) {
  if (!e || !e->replay) { hb_set_error("hb_replay_take: no replay"); return -1; }
  HbReplay* Q = e->replay;
  if (Q->set_count == 0) { hb_set_error("hb_replay_take: no batch is outstanding (hb_replay_prefetch)"); return -3; }
  HB_CUDA(cudaSetDevice(e->device));
  const int set = Q->set_head;
  if (!Q->set_waited[set]) { HB_CUDA(cudaEventSynchronize(Q->set_ev[set])); Q->set_waited[set] = 1; }
  if (batchsize) *batchsize = Q->set_n[set];
  return 0;
}

int hb_replay_last_max_len(hb_engine* e) {
  if (!e || !e->replay) return -1;
  HbReplay* Q = e->replay;
  // the oldest outstanding batch = the one being trained on; none outstanding: the last one taken
  const int set = Q->set_count ? Q->set_head : (Q->set_head + HB_SAMPLE_SETS - 1) % HB_SAMPLE_SETS;
  return Q->h_max_len[set];
}

int hb_stream_wait(hb_engine* e, void* stream) {
  if (!e) { hb_set_error("hb_stream_wait: null engine"); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  cudaEvent_t ev;
  HB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  HB_CUDA(cudaEventRecord(ev, (cudaStream_t)stream));
  HB_CUDA(cudaStreamWaitEvent(e->stream, ev, 0));
  HB_CUDA(cudaEventDestroy(ev));   // released once the wait has been satisfied
  return 0;
}
int hb_stream_wait_engine(hb_engine* e, void* stream) {
  if (!e) { hb_set_error("hb_stream_wait_engine: null engine"); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  cudaEvent_t ev;
  HB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  HB_CUDA(cudaEventRecord(ev, e->stream));
  HB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ev, 0));
  HB_CUDA(cudaEventDestroy(ev));
  return 0;
}

int hb_replay_get(hb_engine* e, int64_t idx, const hb_batch* out) {
  if (!e || !out) { hb_set_error("hb_replay_get: null argument"); return -1; }
  HbReplay* Q = e->replay;
  if (!Q) { hb_set_error("hb_replay_get: this engine has no replay (replay_capacity = 0)"); return -1; }
  int64_t size = 0;
  int rc = hb_counters(e, &size, nullptr, nullptr);
  if (rc) return rc;
  if (idx < 0 || idx >= size) { hb_set_error("hb_replay_get: index %lld out of range, the replay holds %lld entries", (long long)idx, (long long)size); return -3; }
  HbRing& R = Q->ring;
  int* d_entry = Q->d_entry;
  HB_CUDA(cudaMemsetAsync(d_entry, 0xFF, sizeof(int), e->stream));
  hb_k_replay_find<<<(R.phys_slots + 255) / 256, 256, 0, e->stream>>>(R, (long long)idx, d_entry);
  HbBatchPtrs bp = {out->priv_s, out->legal_move, out->own_hand, out->eps, out->a, out->greedy_a, out->reward, out->bootstrap, out->terminal, out->seq_len};
  hb_k_replay_gather<<<dim3(1, R.T), 128, 0, e->stream>>>(R, d_entry, 1, bp, e->env, e->d_eps_list);
  HB_CUDA(cudaGetLastError());
  int h_entry = -1;
  HB_CUDA(cudaMemcpyAsync(&h_entry, d_entry, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  e->launches += 2;
  if (h_entry < 0) { hb_set_error("hb_replay_get: entry %lld not found (ring bookkeeping corrupt)", (long long)idx); return -2; }
  if (out->ids) HB_CUDA(cudaMemcpy(out->ids, &h_entry, sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int hb_replay_update_priority(hb_engine* e, const float* priority, int n) {
  if (!e) { hb_set_error("hb_replay_update_priority: null engine"); return -1; }
  HbReplay* Q = e->replay;
  if (!Q) { hb_set_error("hb_replay_update_priority: this engine has no replay"); return -1; }
  const int set = Q->set_head;
  auto pop = [&]() { Q->set_head = (Q->set_head + 1) % HB_SAMPLE_SETS; Q->set_count -= 1; };
  if (n == 0) { if (Q->set_count) pop(); return 0; }  // prioritized_replay.h:243-246: forget the (oldest) outstanding sample
  const int want = Q->set_count ? Q->set_n[set] : 0;
  if (!priority || n != want) { hb_set_error("hb_replay_update_priority: expected %d priorities, got %d", want, n); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  cudaPointerAttributes pa;
  const bool on_device = cudaPointerGetAttributes(&pa, priority) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  (void)cudaGetLastError();
  HB_CUDA(cudaMemcpyAsync(Q->d_prio, priority, n * sizeof(float), cudaMemcpyDefault, e->stream));
  hb_k_replay_update<<<(n + 127) / 128, 128, 0, e->stream>>>(Q->ring, Q->sampled_idx + (size_t)set * Q->max_batch, Q->sampled_seq + (size_t)set * Q->max_batch,
                                                           Q->d_prio, n);
  HB_CUDA(cudaGetLastError());
  // host memory may be released by the caller as soon as this returns; device memory is stream-ordered (the caller keeps it
  // valid until the engine stream has passed this point, see hb_stream_wait), so a learner loop never blocks here
  if (!on_device) HB_CUDA(cudaStreamSynchronize(e->stream));
  e->launches += 1;
  pop();
  return 0;
}

}  // extern "C"
