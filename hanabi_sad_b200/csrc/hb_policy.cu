// hb_policy.cu -- the R2D2 act forward on the device (pyhanabi/r2d2.py:65-78 R2D2Net.act, :234-303 greedy_act / act,
// and the Q-values compute_priority needs, :305-361), for the online and the target network at once:
//
//   fc    : x  = ReLU(W0 s + b0)                     tcgen05 GEMM, epilogue bias+ReLU          (hb_gemm.cuh, EPI_RELU)
//   lstm l: [x | h_l] [W_ih | W_hh]^T + b -> gates   tcgen05 GEMM, epilogue = LSTM cell update  (EPI_LSTM)
//   head  : adv = W_a h + b_a, v = W_v h + b_v, masked argmax, eps-greedy, dueling Q            (hb_k_head_act, fp32 CUDA cores)
//
// State lives in two ping-pong halves (h as bf16 hi/lo pairs = the next tick's GEMM operand, c in fp32): a tick reads
// half `parity` and writes half `parity^1`, so the target network (which is fed the ONLINE hidden state,
// r2d2_actor.h:139-152) can run in the same launches as the online network without a hazard.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <vector>

#include "hb_engine.h"
#include "hb_env_cta.cuh"
#include "hb_gemm.cuh"
#include "hb_gemm_host.h"
#include "hb_policy.h"

using hbg::Params;

// GEMM variant used by the policy forward (hb_gemm.cuh): 1 = one CTA per tile, 2 = CTA pairs with TMA multicast of the
// weight tile, 3 = tcgen05 CTA pairs (cta_group::2, M = 256).
#define HB_GEMM_MODE 3

// ---------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled hb_get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

// [rows][cols] bf16, cols contiguous; box = 64 columns (128 bytes, one swizzle span) x box_rows rows.
// `row_stride` (elements, 0 = cols) lets a map describe every P-th row of a buffer: one seat's agents.
int hb_make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint64_t row_stride) {
  PFN_tmapEncodeTiled enc = hb_get_encode();
  if (!enc) { hb_set_error("cuTensorMapEncodeTiled is not available from the driver"); return -2; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {(row_stride ? row_stride : cols) * 2};
  cuuint32_t box[2] = {(cuuint32_t)hbg::BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { hb_set_error("cuTensorMapEncodeTiled failed with code %d", (int)r); return -2; }
  return 0;
}

// ---------------------------------------------------------------------------------------- weight preparation
// fp32 [n][k] (nn.Linear / nn.LSTM layout) -> bf16 hi/lo [n_out][k_out]; rows optionally re-ordered so that every
// 256-row tile holds [gate i|f|g|o][64 hidden units] (hb_gemm.cuh EPI_LSTM), columns placed at col0, rest zero.
__global__ void hb_k_prep_weight(const float* __restrict__ w, int n, int k, int lstm_order, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int ld_out, int col0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * k) return;
  const int r = idx / k, c = idx - r * k;
  int ro = r;
  if (lstm_order) {  // r = gate*512 + unit  ->  tile (unit/64), inside the tile gate*64 + unit%64
    const int gate = r / HB_HID, unit = r % HB_HID;
    ro = (unit / 64) * 256 + gate * 64 + (unit % 64);
  }
  const float v = w[idx];
  __nv_bfloat16 h, l;
  hbg::split_bf16(v, h, l);
  hi[(size_t)ro * ld_out + col0 + c] = h;
  lo[(size_t)ro * ld_out + col0 + c] = l;
}

__global__ void hb_k_prep_lstm_bias(const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 4 * HB_HID) return;
  const int gate = r / HB_HID, unit = r % HB_HID;
  out[(unit / 64) * 256 + gate * 64 + (unit % 64)] = b_ih[r] + b_hh[r];
}

// fc_a [A][512] and fc_v [512] -> [8 n-tiles][64 units][hop]: the slice of the head every LSTM output tile needs, outputs of
// one hidden unit contiguous (hb_gemm.cuh head_dot); hop = A + 1 rounded up to 8, the padding stays zero (cleared at creation).
__global__ void hb_k_prep_head(const float* __restrict__ wa, const float* __restrict__ wv, int A, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = (A + 1) * HB_HID, hop = (A + 1 + 7) & ~7;
  if (idx >= n) return;
  const int o = idx / HB_HID, k = idx % HB_HID;
  const float v = o < A ? wa[(size_t)o * HB_HID + k] : wv[k];
  out[((size_t)(k / 64) * 64 + (k % 64)) * hop + o] = v;
}

// ---------------------------------------------------------------------------------------- head + action selection
#include "hb_head.cuh"

#define HB_HEAD_WARPS 8

// One warp per agent (hb_head.cuh).
__global__ void __launch_bounds__(HB_HEAD_WARPS * 32) hb_k_head_act(HbHeadArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * HB_HEAD_WARPS + warp;
  if (row >= p.rows) return;
  hb_head_row(p, row, lane);
}

// ---------------------------------------------------------------------------------------- host side
static int hb_alloc_zero(void** p, size_t bytes, cudaStream_t st) {
  HB_CUDA(cudaMalloc(p, bytes));
  HB_CUDA(cudaMemsetAsync(*p, 0, bytes, st));
  return 0;
}
#define HB_ALLOC(ptr, bytes)                                                   \
  do {                                                                         \
    int _rc = hb_alloc_zero((void**)&(ptr), (bytes), e->stream);               \
    if (_rc) return _rc;                                                       \
  } while (0)

#define HB_PARAMS_PER_PARITY (4 * HB_MAX_P)  // launches fc, fc2, lstm0, lstm1 x up to HB_MAX_P problems
enum { HB_L_FC = 0, HB_L_FC2 = 1, HB_L_LSTM0 = 2, HB_L_LSTM1 = 3 };

// (Re)builds the Params records of every launch for both state parities.  Training engines: problems = {online, target}
// over all agents.  eval_seats engines: problems = seats, each a strided view (every P-th row) of the same buffers.
static int hb_build_params(hb_engine* e) {
  HbPolicy* P = e->policy;
  const int rp = P->rows_pad, KS = P->KS;
  const bool seat = P->seat_mode != 0;
  const int np = seat ? e->P : 2;
  const int mul = seat ? e->P : 1;
  const uint64_t vrows = seat ? (uint64_t)e->G : (uint64_t)rp;  // rows of one problem's operand view
  std::vector<Params> hp(2 * HB_PARAMS_PER_PARITY);
  memset(hp.data(), 0, hp.size() * sizeof(Params));
  const size_t lsz = (size_t)rp * HB_HID;  // one layer of one state half
  int n_fc2 = 0;
  for (int par = 0; par < 2; ++par) {
    const int cur = par, nxt = par ^ 1;
    n_fc2 = 0;
    for (int z = 0; z < np; ++z) {
      const HbNetWeights& W = P->net[z];
      const int add = seat ? z : 0;
      int rc = 0;
      auto amap = [&](CUtensorMap* m, const __nv_bfloat16* base, int ld) {  // A-operand view of a [rows_pad][ld] buffer
        return hb_make_tmap(m, base + (size_t)add * ld, vrows, ld, hbg::BM, (uint64_t)mul * ld);
      };
      __nv_bfloat16 *xh = seat ? P->x_hi[0] : P->x_hi[z], *xl = seat ? P->x_lo[0] : P->x_lo[z];
      const bool fc2 = seat && W.has_fc2;
      __nv_bfloat16 *lin_h = fc2 ? P->x_hi[1] : xh, *lin_l = fc2 ? P->x_lo[1] : xl;  // what the LSTM (and a skip connection) consumes
      Params* base = hp.data() + (size_t)par * HB_PARAMS_PER_PARITY;
      // ---- fc
      Params& f = base[HB_L_FC * HB_MAX_P + z];
      rc |= amap(&f.a_hi[0], P->s_hi, KS); rc |= amap(&f.a_lo[0], P->s_lo, KS);
      f.a_hi[1] = f.a_hi[0]; f.a_lo[1] = f.a_lo[0];
      rc |= hb_make_tmap(&f.b_hi, W.w0_hi, HB_HID, KS, hbg::BN / 2);
      rc |= hb_make_tmap(&f.b_lo, W.w0_lo, HB_HID, KS, hbg::BN / 2);
      f.k_chunks = KS / hbg::BK; f.k_chunks_seg0 = f.k_chunks;
      f.lo_first = e->env.g.off_belief / hbg::BK;
      f.lo_last = (e->env.g.off_sad + hbg::BK - 1) / hbg::BK;
      f.bias = W.b0;
      f.out_hi = xh; f.out_lo = xl; f.out_ld = HB_HID; f.out_col0 = 0;
      // ---- second fc layer (r2d2.py:42-46), compact list: only the seats that have one
      Params* f2 = nullptr;
      if (fc2) {
        f2 = &base[HB_L_FC2 * HB_MAX_P + n_fc2++];
        rc |= amap(&f2->a_hi[0], xh, HB_HID); rc |= amap(&f2->a_lo[0], xl, HB_HID);
        f2->a_hi[1] = f2->a_hi[0]; f2->a_lo[1] = f2->a_lo[0];
        rc |= hb_make_tmap(&f2->b_hi, W.w1_hi, HB_HID, HB_HID, hbg::BN / 2);
        rc |= hb_make_tmap(&f2->b_lo, W.w1_lo, HB_HID, HB_HID, hbg::BN / 2);
        f2->k_chunks = HB_HID / hbg::BK; f2->k_chunks_seg0 = f2->k_chunks; f2->lo_first = 0; f2->lo_last = f2->k_chunks;
        f2->bias = W.b1;
        f2->out_hi = P->x_hi[1]; f2->out_lo = P->x_lo[1]; f2->out_ld = HB_HID; f2->out_col0 = 0;
      }
      const bool owner = seat || z == 0;  // this problem advances the recurrent state (the target network only reads it)
      // ---- lstm layer 0
      Params& l0 = base[HB_L_LSTM0 * HB_MAX_P + z];
      rc |= amap(&l0.a_hi[0], lin_h, HB_HID); rc |= amap(&l0.a_lo[0], lin_l, HB_HID);
      rc |= amap(&l0.a_hi[1], P->h_hi[cur], HB_HID); rc |= amap(&l0.a_lo[1], P->h_lo[cur], HB_HID);
      rc |= hb_make_tmap(&l0.b_hi, W.wl_hi[0], 4 * HB_HID, 2 * HB_HID, hbg::BN / 2);
      rc |= hb_make_tmap(&l0.b_lo, W.wl_lo[0], 4 * HB_HID, 2 * HB_HID, hbg::BN / 2);
      l0.k_chunks = 2 * HB_HID / hbg::BK; l0.k_chunks_seg0 = HB_HID / hbg::BK; l0.lo_first = 0; l0.lo_last = l0.k_chunks;
      l0.bias = W.bl[0];
      l0.c_in = P->c[cur];
      if (owner) { l0.c_out = P->c[nxt]; l0.out_hi = P->h_hi[nxt]; l0.out_lo = P->h_lo[nxt]; }
      else { l0.c_out = nullptr; l0.out_hi = P->th_hi; l0.out_lo = P->th_lo; }
      l0.out_ld = HB_HID; l0.out_col0 = 0; l0.h_f32 = nullptr;
      // ---- lstm layer 1 (+ fused dueling head)
      Params& l1 = base[HB_L_LSTM1 * HB_MAX_P + z];
      if (owner) { rc |= amap(&l1.a_hi[0], P->h_hi[nxt], HB_HID); rc |= amap(&l1.a_lo[0], P->h_lo[nxt], HB_HID); }
      else { rc |= amap(&l1.a_hi[0], P->th_hi, HB_HID); rc |= amap(&l1.a_lo[0], P->th_lo, HB_HID); }
      rc |= amap(&l1.a_hi[1], P->h_hi[cur] + lsz, HB_HID); rc |= amap(&l1.a_lo[1], P->h_lo[cur] + lsz, HB_HID);
      rc |= hb_make_tmap(&l1.b_hi, W.wl_hi[1], 4 * HB_HID, 2 * HB_HID, hbg::BN / 2);
      rc |= hb_make_tmap(&l1.b_lo, W.wl_lo[1], 4 * HB_HID, 2 * HB_HID, hbg::BN / 2);
      l1.k_chunks = 2 * HB_HID / hbg::BK; l1.k_chunks_seg0 = HB_HID / hbg::BK; l1.lo_first = 0; l1.lo_last = l1.k_chunks;
      l1.bias = W.bl[1];
      l1.c_in = P->c[cur] + lsz;
      if (owner) { l1.c_out = P->c[nxt] + lsz; l1.out_hi = P->h_hi[nxt] + lsz; l1.out_lo = P->h_lo[nxt] + lsz; }
      else { l1.c_out = nullptr; l1.out_hi = nullptr; l1.out_lo = nullptr; }
      l1.out_ld = HB_HID; l1.out_col0 = 0; l1.h_f32 = nullptr;
      l1.head_w = W.head_tiles; l1.head_part = seat ? P->head_part[0] : P->head_part[z]; l1.head_out = e->A + 1; l1.head_rows = rp;
      if (seat && W.skip) { l1.skip_hi = lin_h; l1.skip_lo = lin_l; }
      if (rc) return -2;
      for (Params* q : {&f, f2, &l0, &l1}) {
        if (!q) continue;
        q->split = (seat || z == 0 || P->target_split) ? 1 : 0;
        q->row_mul = mul; q->row_add = add; q->valid_rows = (int)vrows;
        q->error_flag = P->d_error;
      }
    }
  }
  P->n_fc2 = n_fc2;
  HB_CUDA(cudaMemcpyAsync(P->d_params, hp.data(), hp.size() * sizeof(Params), cudaMemcpyHostToDevice, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

extern "C" {

int hb_policy_create(hb_engine* e) {
  e->policy = nullptr;
  const hb_config& c = e->cfg;
  if (c.hid_dim == 0) return 0;  // environment-only engine
  if (c.hid_dim != HB_HID || c.num_lstm_layer != HB_LAYERS) {
    hb_set_error("hb_create: the device policy serves hid_dim=512, num_lstm_layer=2 (got %d, %d)", c.hid_dim, c.num_lstm_layer);
    return -1;
  }
  if (!c.eval_seats && (c.num_fc_layer != 1 || c.skip_connect != 0)) {
    hb_set_error("hb_create: num_fc_layer=2 / skip_connect are served by eval_seats engines only (selfplay.py never trains them)");
    return -1;
  }
  if (c.eval_seats && c.replay_capacity > 0) { hb_set_error("hb_create: an eval_seats engine has no replay"); return -1; }
  if (e->A > 63 || ((e->A + 1 + 7) & ~7) > hbg::HEAD_MAX_OUT) {   // 5-player Hanabi has 49 moves incl. the no-op: 50 outputs, padded to 56
    hb_set_error("hb_create: num_action = %d; the fused head holds %d outputs (advantages + value, padded to 8)", e->A, hbg::HEAD_MAX_OUT);
    return -1;
  }
  HbPolicy* P = new HbPolicy();
  memset(P, 0, sizeof(*P));
  e->policy = P;
  P->rows = e->rows;
  P->rows_pad = (e->rows + 2 * hbg::BM - 1) / (2 * hbg::BM) * (2 * hbg::BM);  // CTA pairs work on two vertically adjacent tiles
  P->KS = (e->F + hbg::BK - 1) / hbg::BK * hbg::BK;
  P->target_split = c.priority_mode == 2 ? 0 : 1;
  P->seat_mode = c.eval_seats ? 1 : 0;
  const size_t rp = P->rows_pad, KS = P->KS, bf = sizeof(__nv_bfloat16);
  HB_ALLOC(P->s_hi, rp * KS * bf);
  HB_ALLOC(P->s_lo, rp * KS * bf);
  const int n_nets = P->seat_mode ? e->P : 2;
  for (int n = 0; n < 2; ++n) {
    HB_ALLOC(P->x_hi[n], rp * HB_HID * bf);
    HB_ALLOC(P->x_lo[n], rp * HB_HID * bf);
    HB_ALLOC(P->head_part[n], (size_t)8 * rp * ((e->A + 1 + 7) & ~7) * sizeof(float));
    HB_ALLOC(P->h_hi[n], HB_LAYERS * rp * HB_HID * bf);
    HB_ALLOC(P->h_lo[n], HB_LAYERS * rp * HB_HID * bf);
    HB_ALLOC(P->c[n], HB_LAYERS * rp * HB_HID * sizeof(float));
  }
  for (int n = 0; n < n_nets; ++n) {
    HbNetWeights& W = P->net[n];
    HB_ALLOC(W.w0_hi, (size_t)HB_HID * KS * bf);
    HB_ALLOC(W.w0_lo, (size_t)HB_HID * KS * bf);
    HB_ALLOC(W.b0, HB_HID * sizeof(float));
    for (int l = 0; l < HB_LAYERS; ++l) {
      HB_ALLOC(W.wl_hi[l], (size_t)4 * HB_HID * 2 * HB_HID * bf);
      HB_ALLOC(W.wl_lo[l], (size_t)4 * HB_HID * 2 * HB_HID * bf);
      HB_ALLOC(W.bl[l], 4 * HB_HID * sizeof(float));
    }
    HB_ALLOC(W.wa, (size_t)e->A * HB_HID * sizeof(float));
    HB_ALLOC(W.ba, e->A * sizeof(float));
    HB_ALLOC(W.wv, HB_HID * sizeof(float));
    HB_ALLOC(W.head_tiles, (size_t)8 * ((e->A + 1 + 7) & ~7) * 64 * sizeof(float));   // zero-filled: the padding columns are never written
    HB_ALLOC(W.bv, sizeof(float));
    size_t raw = (size_t)HB_HID * e->F;
    if (raw < (size_t)4 * HB_HID * HB_HID) raw = (size_t)4 * HB_HID * HB_HID;
    HB_ALLOC(W.raw, raw * sizeof(float));
    HB_ALLOC(W.raw2, (size_t)4 * HB_HID * sizeof(float) * 2);
    if (P->seat_mode) {
      HB_ALLOC(W.w1_hi, (size_t)HB_HID * HB_HID * bf);
      HB_ALLOC(W.w1_lo, (size_t)HB_HID * HB_HID * bf);
      HB_ALLOC(W.b1, HB_HID * sizeof(float));
    }
  }
  HB_ALLOC(P->th_hi, rp * HB_HID * bf);
  HB_ALLOC(P->th_lo, rp * HB_HID * bf);
  HB_ALLOC(P->adv, (size_t)e->rows * e->A * sizeof(float));
  HB_ALLOC(P->oq, e->rows * sizeof(float));
  HB_ALLOC(P->tq, e->rows * sizeof(float));
  HB_ALLOC(P->d_error, sizeof(int));
  HB_ALLOC(P->d_params, 2 * HB_PARAMS_PER_PARITY * sizeof(Params));
  e->obs.s_hi = P->s_hi; e->obs.s_lo = P->s_lo; e->obs.KS = P->KS;
  HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_RELU, HB_GEMM_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_LSTM, HB_GEMM_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  return hb_build_params(e);
}

void hb_policy_destroy(hb_engine* e) {
  HbPolicy* P = e->policy;
  if (!P) return;
  cudaFree(P->s_hi); cudaFree(P->s_lo); cudaFree(P->th_hi); cudaFree(P->th_lo);
  for (int n = 0; n < 2; ++n) {
    cudaFree(P->x_hi[n]); cudaFree(P->x_lo[n]); cudaFree(P->head_part[n]); cudaFree(P->h_hi[n]); cudaFree(P->h_lo[n]); cudaFree(P->c[n]);
  }
  for (int n = 0; n < HB_MAX_P; ++n) {
    HbNetWeights& W = P->net[n];
    cudaFree(W.w1_hi); cudaFree(W.w1_lo); cudaFree(W.b1);
    cudaFree(W.w0_hi); cudaFree(W.w0_lo); cudaFree(W.b0);
    for (int l = 0; l < HB_LAYERS; ++l) { cudaFree(W.wl_hi[l]); cudaFree(W.wl_lo[l]); cudaFree(W.bl[l]); }
    cudaFree(W.wa); cudaFree(W.ba); cudaFree(W.wv); cudaFree(W.bv); cudaFree(W.head_tiles); cudaFree(W.raw); cudaFree(W.raw2);
  }
  cudaFree(P->adv); cudaFree(P->oq); cudaFree(P->tq); cudaFree(P->d_error); cudaFree(P->d_params);
  delete P;
  e->policy = nullptr;
}

// R2D2Agent state_dict -> device operands (BatchRunner::updateModel, rela/batch_runner.h:74-77, without the
// TorchScript module: the tensors may live on the host or on any device -- cudaMemcpyDefault).
int hb_policy_set_weights(hb_engine* e, int net, const hb_weights* w) {
  if (!e || !w) { hb_set_error("hb_policy_set_weights: null argument"); return -1; }
  HbPolicy* P = e->policy;
  if (!P) { hb_set_error("hb_policy_set_weights: this engine was created without a policy (hid_dim = 0)"); return -1; }
  const int n_nets = P->seat_mode ? e->P : 2;
  if (net < 0 || net >= n_nets) {
    hb_set_error(P->seat_mode ? "hb_policy_set_weights: net must be a seat index 0..%d" : "hb_policy_set_weights: net must be 0 (online) or 1 (target), max %d", n_nets - 1);
    return -1;
  }
  if (!P->seat_mode && (w->fc2_w != nullptr || w->skip_connect != 0)) {
    hb_set_error("hb_policy_set_weights: num_fc_layer=2 / skip_connect need an eval_seats engine");
    return -1;
  }
  if (w->fc2_w != nullptr && w->fc2_b == nullptr) { hb_set_error("hb_policy_set_weights: fc2_b is null"); return -1; }
  const void* need[] = {w->fc_w, w->fc_b, w->w_ih[0], w->w_hh[0], w->b_ih[0], w->b_hh[0], w->w_ih[1], w->w_hh[1], w->b_ih[1], w->b_hh[1],
                        w->fc_a_w, w->fc_a_b, w->fc_v_w, w->fc_v_b};
  for (const void* p : need)
    if (!p) { hb_set_error("hb_policy_set_weights: a weight pointer is null"); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  HbNetWeights& W = P->net[net];
  cudaStream_t st = e->stream;
  const int F = e->F, A = e->A, KS = P->KS;
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  HB_CUDA(cudaMemcpyAsync(W.raw, w->fc_w, (size_t)HB_HID * F * sizeof(float), cudaMemcpyDefault, st));
  hb_k_prep_weight<<<blocks((size_t)HB_HID * F), 256, 0, st>>>(W.raw, HB_HID, F, 0, W.w0_hi, W.w0_lo, KS, 0);
  HB_CUDA(cudaMemcpyAsync(W.b0, w->fc_b, HB_HID * sizeof(float), cudaMemcpyDefault, st));
  W.has_fc2 = w->fc2_w != nullptr;
  W.skip = w->skip_connect != 0;
  if (W.has_fc2) {
    HB_CUDA(cudaMemcpyAsync(W.raw, w->fc2_w, (size_t)HB_HID * HB_HID * sizeof(float), cudaMemcpyDefault, st));
    hb_k_prep_weight<<<blocks((size_t)HB_HID * HB_HID), 256, 0, st>>>(W.raw, HB_HID, HB_HID, 0, W.w1_hi, W.w1_lo, HB_HID, 0);
    HB_CUDA(cudaMemcpyAsync(W.b1, w->fc2_b, HB_HID * sizeof(float), cudaMemcpyDefault, st));
    e->launches += 1;
  }
  for (int l = 0; l < HB_LAYERS; ++l) {
    const size_t n = (size_t)4 * HB_HID * HB_HID;
    HB_CUDA(cudaMemcpyAsync(W.raw, w->w_ih[l], n * sizeof(float), cudaMemcpyDefault, st));
    hb_k_prep_weight<<<blocks(n), 256, 0, st>>>(W.raw, 4 * HB_HID, HB_HID, 1, W.wl_hi[l], W.wl_lo[l], 2 * HB_HID, 0);
    HB_CUDA(cudaMemcpyAsync(W.raw, w->w_hh[l], n * sizeof(float), cudaMemcpyDefault, st));
    hb_k_prep_weight<<<blocks(n), 256, 0, st>>>(W.raw, 4 * HB_HID, HB_HID, 1, W.wl_hi[l], W.wl_lo[l], 2 * HB_HID, HB_HID);
    HB_CUDA(cudaMemcpyAsync(W.raw2, w->b_ih[l], 4 * HB_HID * sizeof(float), cudaMemcpyDefault, st));
    HB_CUDA(cudaMemcpyAsync(W.raw2 + 4 * HB_HID, w->b_hh[l], 4 * HB_HID * sizeof(float), cudaMemcpyDefault, st));
    hb_k_prep_lstm_bias<<<blocks(4 * HB_HID), 256, 0, st>>>(W.raw2, W.raw2 + 4 * HB_HID, W.bl[l]);
    e->launches += 3;
  }
  e->launches += 1;
  HB_CUDA(cudaMemcpyAsync(W.wa, w->fc_a_w, (size_t)A * HB_HID * sizeof(float), cudaMemcpyDefault, st));
  HB_CUDA(cudaMemcpyAsync(W.ba, w->fc_a_b, A * sizeof(float), cudaMemcpyDefault, st));
  HB_CUDA(cudaMemcpyAsync(W.wv, w->fc_v_w, HB_HID * sizeof(float), cudaMemcpyDefault, st));
  HB_CUDA(cudaMemcpyAsync(W.bv, w->fc_v_b, sizeof(float), cudaMemcpyDefault, st));
  hb_k_prep_head<<<blocks((size_t)(A + 1) * HB_HID), 256, 0, st>>>(W.wa, W.wv, A, W.head_tiles);
  e->launches += 1;
  HB_CUDA(cudaGetLastError());
  HB_CUDA(cudaStreamSynchronize(st));  // the caller may free / overwrite its tensors once this returns
  P->have_weights[net] = 1;
  if (P->seat_mode) return hb_build_params(e);  // the launch lists depend on each seat's architecture variant
  return 0;
}

}  // extern "C"

// Persistent launch: one CTA per SM (as many as there are work items if fewer), in clusters of `cl` CTAs.
int hb_launch_gemm(HbGemmKernel k, int cl, int sm_count, cudaStream_t st, const Params* ps, int nt, int mt, int nprob) {
  const int items = nt * (mt / cl) * nprob;
  int clusters = sm_count / cl;
  if (items < clusters) clusters = items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * cl), 1, 1);
  cfg.blockDim = dim3(hbg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = hbg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  HB_CUDA(cudaLaunchKernelEx(&cfg, k, ps, nt, mt, nprob));
  return 0;
}

int hb_policy_error_ptr(hb_engine* e, int** out) { *out = e->policy ? e->policy->d_error : nullptr; return 0; }

HbHidPtrs hb_policy_hidden_ptrs(hb_engine* e) {
  HbHidPtrs h = {nullptr, nullptr, nullptr, 0};
  if (e->policy) { HbPolicy* P = e->policy; h.h_hi = P->h_hi[P->parity]; h.h_lo = P->h_lo[P->parity]; h.c = P->c[P->parity]; h.rows_pad = P->rows_pad; }
  return h;
}

// One act forward over the CURRENT observation operands (s_hi / s_lo, legal_move, eps) and state half `parity`:
// 3 GEMM launches (both networks each) + the head kernel.  Leaves the new state in half parity^1 and flips parity.
// Launches the head / act kernel a deferred forward left pending (no-op otherwise).
int hb_policy_flush_head(hb_engine* e) {
  HbPolicy* P = e->policy;
  if (!P || !P->head_pending) return 0;
  P->head_pending = 0;
  { HbProfScope ps(e, HB_PROF_HEAD);
    hb_k_head_act<<<(e->rows + HB_HEAD_WARPS - 1) / HB_HEAD_WARPS, HB_HEAD_WARPS * 32, 0, e->stream>>>(P->pending_head); }
  HB_CUDA(cudaGetLastError());
  e->launches += 1;
  return 0;
}

// defer_head != 0: the three GEMM launches only; the head / act step of this forward is left pending (HbPolicy::pending_head)
// for the next fused tick to run as its prologue (hb_launch_tick), or for hb_policy_flush_head.
int hb_policy_forward(hb_engine* e, int greedy_only, int defer_head) {
  HbPolicy* P = e->policy;
  if (!P || !P->have_weights[0]) { hb_set_error("policy forward without online weights (call hb_policy_set_weights first)"); return -1; }
  int nets;
  if (P->seat_mode) {
    nets = e->P;
    for (int n = 0; n < nets; ++n)
      if (!P->have_weights[n]) { hb_set_error("policy forward: seat %d has no weights (hb_policy_set_weights)", n); return -1; }
  } else {
    nets = P->have_weights[1] && e->cfg.priority_mode != 1 ? 2 : 1;
  }
  const int view_rows = P->seat_mode ? (e->G + 2 * hbg::BM - 1) / (2 * hbg::BM) * (2 * hbg::BM) : P->rows_pad;
  const int mt = view_rows / hbg::BM;
  const Params* base = P->d_params + (size_t)P->parity * HB_PARAMS_PER_PARITY;
  const int nt_fc = HB_HID / hbg::BN, nt_l = 4 * HB_HID / hbg::BN;
  int rc;
  { HbProfScope ps(e, HB_PROF_FC);
    rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_RELU, HB_GEMM_MODE>, 2, e->sm_count, e->stream, base + HB_L_FC * HB_MAX_P, nt_fc, mt, nets);
    if (!rc && P->seat_mode && P->n_fc2 > 0) {
      rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_RELU, HB_GEMM_MODE>, 2, e->sm_count, e->stream, base + HB_L_FC2 * HB_MAX_P, nt_fc, mt, P->n_fc2);
      e->launches += 1;
    } }
  if (rc) return rc;
  { HbProfScope ps(e, HB_PROF_LSTM0);
    rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_LSTM, HB_GEMM_MODE>, 2, e->sm_count, e->stream, base + HB_L_LSTM0 * HB_MAX_P, nt_l, mt, nets); }
  if (rc) return rc;
  { HbProfScope ps(e, HB_PROF_LSTM1);
    rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_LSTM, HB_GEMM_MODE>, 2, e->sm_count, e->stream, base + HB_L_LSTM1 * HB_MAX_P, nt_l, mt, nets); }
  if (rc) return rc;
  HbHeadArgs a;
  a.rows = e->rows; a.rows_pad = P->rows_pad; a.A = e->A; a.have_target = !P->seat_mode && nets == 2;
  a.seat_mode = P->seat_mode; a.P = e->P;
  for (int n = 0; n < 2; ++n) a.part[n] = P->head_part[n];
  for (int n = 0; n < HB_MAX_P; ++n) { a.ba[n] = P->net[n].ba; a.bv[n] = P->net[n].bv; }
  a.legal = e->obs.legal_move; a.eps = e->obs.eps; a.a = e->d_a; a.greedy_a = e->d_greedy_a;
  a.adv = P->adv; a.oq = P->oq; a.tq = P->tq; a.seed = e->cfg.seed; a.tick = (uint32_t)P->act_count; a.tick_ctr = nullptr;
  a.greedy_only = greedy_only;
  P->pending_head = a;
  P->head_pending = 1;
  e->launches += 3;
  P->parity ^= 1;
  P->act_count += 1;
  if (defer_head) { HB_CUDA(cudaGetLastError()); return 0; }
  return hb_policy_flush_head(e);
}

// ---------------------------------------------------------------------------------------- diagnostics
__global__ void hb_k_split_rows(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo, int ld) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * cols) return;
  const int r = (int)(idx / cols), c = (int)(idx % cols);
  __nv_bfloat16 h, l;
  hbg::split_bf16(src[idx], h, l);
  hi[(size_t)r * ld + c] = h;
  lo[(size_t)r * ld + c] = l;
}

__global__ void hb_k_merge_hidden(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int rows, int rows_pad,
                                  float* __restrict__ out) {  // [L][rows_pad][512] hi/lo -> [L][rows][512] fp32
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)HB_LAYERS * rows * HB_HID) return;
  const int l = (int)(idx / ((size_t)rows * HB_HID));
  const size_t rem = idx % ((size_t)rows * HB_HID);
  const size_t src = (size_t)l * rows_pad * HB_HID + rem;
  out[idx] = __bfloat162float(hi[src]) + __bfloat162float(lo[src]);
}

extern "C" {

// Stand-alone run of the tcgen05 GEMM template: C[M,N] = A[M,K] B[N,K]^T + bias (host fp32 in/out), `split` = bf16x3.
int hb_debug_gemm(int device, const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int split) {
  if (!A || !B || !C) { hb_set_error("hb_debug_gemm: null argument"); return -1; }
  if (M % hbg::BM || N % hbg::BN || K % hbg::BK) { hb_set_error("hb_debug_gemm: M, N, K must be multiples of 128, 256, 64"); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { hb_set_error("hb_debug_gemm: no CUDA device"); return -2; }
  HB_CUDA(cudaSetDevice(device));
  float *dA, *dB, *dC, *dbias;
  __nv_bfloat16 *ah, *al, *bh, *bl;
  Params* dp;
  int* derr;
  const size_t bf = sizeof(__nv_bfloat16);
  HB_CUDA(cudaMalloc(&dA, (size_t)M * K * 4)); HB_CUDA(cudaMalloc(&dB, (size_t)N * K * 4)); HB_CUDA(cudaMalloc(&dC, (size_t)M * N * 4));
  HB_CUDA(cudaMalloc(&dbias, (size_t)N * 4));
  HB_CUDA(cudaMalloc(&ah, (size_t)M * K * bf)); HB_CUDA(cudaMalloc(&al, (size_t)M * K * bf));
  HB_CUDA(cudaMalloc(&bh, (size_t)N * K * bf)); HB_CUDA(cudaMalloc(&bl, (size_t)N * K * bf));
  HB_CUDA(cudaMalloc(&dp, sizeof(Params))); HB_CUDA(cudaMalloc(&derr, sizeof(int)));
  HB_CUDA(cudaMemset(derr, 0, sizeof(int)));
  HB_CUDA(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(dB, B, (size_t)N * K * 4, cudaMemcpyHostToDevice));
  if (bias) HB_CUDA(cudaMemcpy(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  else HB_CUDA(cudaMemset(dbias, 0, (size_t)N * 4));
  hb_k_split_rows<<<(unsigned)(((size_t)M * K + 255) / 256), 256>>>(dA, M, K, ah, al, K);
  hb_k_split_rows<<<(unsigned)(((size_t)N * K + 255) / 256), 256>>>(dB, N, K, bh, bl, K);
  Params hp;
  memset(&hp, 0, sizeof(hp));
  int rc = 0;
  rc |= hb_make_tmap(&hp.a_hi[0], ah, M, K, hbg::BM); rc |= hb_make_tmap(&hp.a_lo[0], al, M, K, hbg::BM);
  hp.a_hi[1] = hp.a_hi[0]; hp.a_lo[1] = hp.a_lo[0];
  // exercises all three variants: M % 512 == 0 -> cta_group::2 pairs, M % 256 == 0 -> multicast pairs, else single CTAs
  const int mode = (M % (4 * hbg::BM) == 0) ? 3 : ((M % (2 * hbg::BM) == 0) ? 2 : 1);
  const int cl = mode == 1 ? 1 : 2;
  rc |= hb_make_tmap(&hp.b_hi, bh, N, K, hbg::BN / cl); rc |= hb_make_tmap(&hp.b_lo, bl, N, K, hbg::BN / cl);
  if (rc) return rc;
  hp.k_chunks = K / hbg::BK; hp.k_chunks_seg0 = hp.k_chunks; hp.lo_first = 0; hp.lo_last = hp.k_chunks; hp.split = split;
  hp.bias = dbias; hp.c_f32 = dC; hp.ldc = N; hp.error_flag = derr;
  hp.row_mul = 1; hp.row_add = 0; hp.valid_rows = M;
  HB_CUDA(cudaMemcpy(dp, &hp, sizeof(hp), cudaMemcpyHostToDevice));
  HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_F32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_F32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_F32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  {
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, device));
    rc = hb_launch_gemm(mode == 3 ? hbg::gemm3_kernel<hbg::EPI_F32, 3> : (mode == 2 ? hbg::gemm3_kernel<hbg::EPI_F32, 2> : hbg::gemm3_kernel<hbg::EPI_F32, 1>), cl,
                        prop.multiProcessorCount, 0, dp,
                        N / hbg::BN, M / hbg::BM, 1);
    if (rc) return rc;
  }
  HB_CUDA(cudaGetLastError());
  HB_CUDA(cudaDeviceSynchronize());
  int herr = 0;
  HB_CUDA(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
  HB_CUDA(cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias); cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl); cudaFree(dp); cudaFree(derr);
  if (herr) { hb_set_error("hb_debug_gemm: a pipeline barrier timed out (spin guard)"); return -4; }
  return 0;
}

// Diagnostic: the fc GEMM's A operand as the encoder left it -- bf16 bit patterns [rows][KS] (hi and lo halves).
int hb_debug_operand(hb_engine* e, uint16_t* hi, uint16_t* lo, int* ks) {
  if (!e || !e->policy) { hb_set_error("hb_debug_operand: no policy"); return -1; }
  HbPolicy* P = e->policy;
  HB_CUDA(cudaSetDevice(e->device));
  if (ks) *ks = P->KS;
  const size_t n = (size_t)e->rows * P->KS * sizeof(uint16_t);
  if (hi) HB_CUDA(cudaMemcpyAsync(hi, P->s_hi, n, cudaMemcpyDeviceToHost, e->stream));
  if (lo) HB_CUDA(cudaMemcpyAsync(lo, P->s_lo, n, cudaMemcpyDeviceToHost, e->stream));
  HB_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

// R2D2Actor::act without the environment (r2d2_actor.h:61-100): one policy forward on the engine's current
// observation; the reply (a, greedy_a) lands in the engine's action buffers, the hidden state advances.
int hb_policy_act(hb_engine* e, int greedy_only) {
  if (!e) { hb_set_error("hb_policy_act: null engine"); return -1; }
  HB_CUDA(cudaSetDevice(e->device));
  e->pending_actions = 1;
  return hb_policy_forward(e, greedy_only, 0);
}

// Host copies of the policy's outputs of the last forward: adv [rows,A] (online advantages), online_q / target_q
// [rows] (dueling Q of the taken action / of the greedy action under the target net), and the CURRENT hidden state
// h, c as [L, rows, 512] fp32 (R2D2Actor::hidden_, r2d2_actor.h:186).  NULL skips.
int hb_policy_get(hb_engine* e, float* adv, float* online_q, float* target_q, float* h, float* c) {
  if (!e || !e->policy) { hb_set_error("hb_policy_get: no policy"); return -1; }
  HbPolicy* P = e->policy;
  HB_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  const size_t R = e->rows;
  if (adv) HB_CUDA(cudaMemcpyAsync(adv, P->adv, R * e->A * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (online_q) HB_CUDA(cudaMemcpyAsync(online_q, P->oq, R * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (target_q) HB_CUDA(cudaMemcpyAsync(target_q, P->tq, R * sizeof(float), cudaMemcpyDeviceToHost, st));
  const int cur = P->parity;
  if (h) {
    float* tmp;
    HB_CUDA(cudaMalloc(&tmp, HB_LAYERS * R * HB_HID * sizeof(float)));
    hb_k_merge_hidden<<<(unsigned)((HB_LAYERS * R * HB_HID + 255) / 256), 256, 0, st>>>(P->h_hi[cur], P->h_lo[cur], (int)R, P->rows_pad, tmp);
    HB_CUDA(cudaMemcpyAsync(h, tmp, HB_LAYERS * R * HB_HID * sizeof(float), cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    e->launches += 1;
  }
  if (c) {
    for (int l = 0; l < HB_LAYERS; ++l)
      HB_CUDA(cudaMemcpyAsync(c + (size_t)l * R * HB_HID, P->c[cur] + (size_t)l * P->rows_pad * HB_HID, R * HB_HID * sizeof(float),
                              cudaMemcpyDeviceToHost, st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  int herr = 0;
  HB_CUDA(cudaMemcpy(&herr, P->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) { hb_set_error("policy GEMM: a pipeline barrier timed out (spin guard)"); return -4; }
  return 0;
}

}  // extern "C"
