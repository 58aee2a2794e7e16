// hb_lstm.cu -- learner side: the T-step, 2-layer, hid=512 LSTM of R2D2Net.forward (pyhanabi/r2d2.py:99-105,
// nn.LSTM over the padded [T, rows, 512] sequence with zero initial state, r2d2.py:383-401) and its backward, as
// sm_100a kernels.  cuDNN runs this as ~1000 small launches per update (one GEMM + one pointwise kernel per step, layer
// and direction); here each layer is
//
//   input projection   GX = X W_ih^T + (b_ih + b_hh) over ALL steps at once      tcgen05 GEMM template (hb_gemm.cuh, EPI_F32)
//   recurrence         ONE persistent kernel per layer (lstm_fwd_kernel): every CTA keeps a 64-column slice
//                      ([i|f|g|o] x 16 hidden units) of W_hh resident in shared memory as bf16 hi/lo, per step pulls
//                      h_{t-1} (bf16 hi/lo, TMA, multicast inside 4-CTA clusters) of its 128-row block, runs 64 tcgen05.mma
//                      (bf16x3 in two instructions per K-step: h_hi x [W_hi | W_lo] as one N = 128 tile + h_lo x W_hi, fp32
//                      accumulate in TMEM), applies the cell update with c kept in registers across all T steps, and
//                      publishes its slice of h_t; the 32 CTAs of a row block meet at a global-memory step counter.
//                      Online and target network run in the same launch.
//   backward           lstm_bwd_kernel, same residency idea with the contraction split over K: a CTA turns dh_t of its 16
//                      units into its 64 dgate columns, multiplies them (as the A operand, written to swizzled shared
//                      memory) with its resident 64 x 512 slice of W_hh and ADDS its [128 x 512] partial of dh_{t-1} into
//                      the row block's accumulator with bulk reductions (cp.reduce.async.bulk .add.f32 from shared-memory
//                      staging, one pipeline per warp: the 32-way sum happens at L2); after the step barrier a consumer
//                      reads its 64 bytes and clears them.
//   layer wavefront    the two layers' recurrences run side by side on internal streams, one 8-step time chunk apart, with
//                      the chunk's input projection / dX GEMM on the SMs they leave free; with one row block the weight
//                      gradients are accumulated chunk by chunk on a fourth stream as well (hb_lstm_forward / _backward).
//   dX, dW, db         dX = dG W_ih and the four weight gradients dG^T X / dG^T H_{t-1} as GEMM-template launches (the four
//                      as problems of ONE launch, or per chunk inside the wavefront); the recurrence kernels leave h and dG
//                      row-major as bf16 hi/lo, a tiled transpose (lstm_transpose_pair) adds the other orientation; db by a
//                      reduction.
//   row accesses       a lane of the cell updates owns a ROW: every global access is 32 bytes wide (ld/st.global.v8), i.e. a
//                      whole sector per lane -- 16-byte accesses cost twice the LSU sector transactions.
//
// Measured (B200, T = 80): forward recurrence 7.0 us / step, backward 12.2 us / step stand-alone (round 1: 10.3 /
// 19) -- bound by moving 256 KB per SM and step (h_{t-1} in, the dh partial out through L2 reductions) and the per-step
// handshake, not by the tensor pipe; whole update of the learner 3.0 ms (IQL, 128 rows) / 4.2 ms (VDN, 256 rows), of which
// the recurrences are 1.9 / 2.4 ms.
//
// Arithmetic: fp32-class everywhere (bf16x3 products accumulated in fp32, pointwise math in fp32); parity target: CPU fp32
// torch.nn.LSTM forward / autograd within 1e-4 (tests/test_lstm_train_parity.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "hb_engine.h"
#include "hb_gemm.cuh"
#include "hb_gemm_host.h"

using hbg::Params;
using namespace hbg;

namespace hbl {

constexpr int HIDN = 512;
constexpr int G4 = 4 * HIDN;          // 2048 gate columns
constexpr int UPC = 16;               // hidden units per CTA
constexpr int NC = 4 * UPC;           // gate columns per CTA: [i|f|g|o][16 units]
constexpr int SLICES = HIDN / UPC;    // 32 CTAs share one 128-row block ("domain")
constexpr int KCH = HIDN / BK;        // 8 K-chunks of 64
constexpr int W_HALF = NC * BK * 2;   // 8 KB: one [64 x 64] bf16 tile
constexpr int W_BYTES = KCH * 2 * W_HALF;   // 128 KB: the CTA's W_hh slice, hi + lo
constexpr int FWD_NST = 3;
constexpr int FWD_STAGE = 2 * A_TILE;       // h hi + lo chunk: 32 KB
constexpr int FWD_SMEM = W_BYTES + FWD_NST * FWD_STAGE + 1024 + 256;
constexpr int BWD_WT_HALF = 256 * BK * 2;   // 32 KB: [256 x 64] bf16 tile
constexpr int BWD_WT_BYTES = 4 * BWD_WT_HALF;  // 128 KB: [512 x 64] hi + lo
constexpr int BWD_STAGE = 64 * BM * 4;       // 32 KB: one [64 columns][128 rows] fp32 piece of the dh partial, staged for a bulk reduction
constexpr int BWD_SMEM = BWD_WT_BYTES + 2 * A_TILE + 2 * BWD_STAGE + 1024 + 256;
constexpr int CTR_STRIDE = 32;        // unsigned ints between two domains' step counters (128 bytes)

struct FwdNet {
  CUtensorMap w_hi, w_lo;             // W_hh in CTA-slice row order [2048][512], box 64 x 64
  CUtensorMap h_hi, h_lo;             // h sequence [(T+1)*R_pad][512], block 0 = h_{-1} = 0; box 128 x 64
  CUtensorMap hq_hi, hq_lo;           // the same with box 32 x 64 (cluster multicast: each CTA fetches a quarter)
  const float* gx;                    // [T*R_pad][2048] input projection + bias, CTA-slice column order
  __nv_bfloat16 *hs_hi, *hs_lo;       // same buffer as h_hi / h_lo
  float* y;                           // fp32 output [T][rows][512] (the caller's tensor) or null
  float* act;                         // [T*R_pad][2048] gate activations i,f,g,o (saved for backward) or null
  float* cs;                          // [T*R_pad][512] cell states (saved for backward) or null
};
struct __align__(64) FwdParams {
  FwdNet net[2];
  int T, rows, R_pad, MB;
  long long ldT;
  unsigned* ctr;
  int* error_flag;
  // layer wavefront (layer 1 only, null otherwise): gx of steps [c*chunk, (c+1)*chunk) is complete once chunk_flags[c] != 0
  const unsigned* chunk_flags;
  int chunk;
  long long* trace;                   // diagnostic (HB_LSTM_TRACE): %globaltimer stamps of CTA 0, [T][16]; null in production
};
struct __align__(64) BwdParams {
  CUtensorMap wt_hi, wt_lo;           // W_hh^T [512][2048 slice-order columns], box 256 x 64
  const float* dh_ext;                // [T][dh_rows][512] gradient w.r.t. this layer's output sequence
  int dh_rows;
  const float* act;
  const float* cs;
  __nv_bfloat16 *dg_hi, *dg_lo;       // [T*R_pad][2048] dgates, slice column order
  float* part;                        // [2][MB*32][128][512] split-K partials of dh_{t-1}
  int T, rows, R_pad, MB;
  unsigned* ctr;
  int* error_flag;
  // layer wavefront (layer 0 only, null otherwise): dh_ext of steps [T - (c+1)*chunk, T - c*chunk) is complete once
  // chunk_flags[c] != 0 (set by the host-side pipeline after the dX GEMM of that time chunk of layer 1)
  const unsigned* chunk_flags;
  int chunk;
  long long* trace;                   // diagnostic (HB_LSTM_TRACE): %globaltimer stamps of CTA 0, [T][16]; null in production
};

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define HBL_STAMP(tr, step, k) do { if ((tr) != nullptr) (tr)[(size_t)(step) * 16 + (k)] = gtimer(); } while (0)

__device__ __forceinline__ constexpr uint32_t idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Bounded waits: a protocol bug must not hang the (shared) GPU.  Once any thread of the grid gives up it raises
// *error_flag, and every other wait in the grid sees the flag within ~1000 polls and gives up too.
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int* error_flag, bool& dead) {
  if (dead) return;
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if ((spin & 1023u) == 1023u) {
      if (*(volatile int*)error_flag != 0) { dead = true; return; }
      if (spin > (1u << 22)) { atomicCAS(error_flag, 0, 1); dead = true; return; }
    }
  }
}
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, int* error_flag, bool& dead, int code = 2) {
  if (dead) return;
#ifdef HBL_POLL_ACQUIRE
  for (uint32_t spin = 0; ld_acquire(ctr) < target; ++spin) {   // every poll is an acquire: no extra round trip once the value is seen
#else
  for (uint32_t spin = 0; ld_relaxed(ctr) < target; ++spin) {
#endif
    if ((spin & 255u) == 255u) {
      if (*(volatile int*)error_flag != 0) { dead = true; return; }
      if (spin > (1u << 20)) { atomicCAS(error_flag, 0, code); dead = true; return; }   // the FIRST failure's code is kept
    }
  }
#ifndef HBL_POLL_ACQUIRE
  // the counter only grows: an acquire load of it now synchronises with every release that contributed to the value seen
  // (cheaper on the step chain than a full fence, which also waits for this thread's own outstanding loads)
  (void)ld_acquire(ctr);
#endif
}
// Step publication (the grid.sync idiom): every thread's stores are ordered before the CTA barrier, ONE thread then
// fences at GPU scope (cumulative over what the barrier ordered) and bumps the domain's step counter.
__device__ __forceinline__ void publish_step(unsigned* ctr) {
  fence_async_global();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 64) red_release_add(ctr, 1u);   // release is cumulative over what the barrier ordered before it
}

// ------------------------------------------------------------------------------------------------ forward recurrence
// grid = nets * MB * 32 CTAs (all co-resident: one per SM), 192 threads: warp 0 TMA producer + step-barrier poller,
// warp 1 MMA issuer / TMEM owner, warps 2-5 cell update (thread = one sequence row of the block).
//
// CL = 4: the grid is launched as clusters of four CTAs of the same row block.  All of them need the same h_{t-1}
// chunks, so each CTA fetches a quarter (32 rows) of every chunk and TMA-multicasts it into the four shared memories:
// L2 -> SM traffic of the recurrence drops 4x.  A ring slot is refilled only after all four CTAs' MMAs released it
// (empty barriers count 4, released by a multicast tcgen05.commit).
template <int CL>
__global__ void __launch_bounds__(192, 1) lstm_fwd_kernel(const FwdParams* __restrict__ pp) {
  extern __shared__ uint8_t smem_raw[];
  const FwdParams& P = *pp;
  const int slice = blockIdx.x % SLICES, dom = blockIdx.x / SLICES, mb = dom % P.MB, net = dom / P.MB;
  const FwdNet& N = P.net[net];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base, ring = base + W_BYTES, bar_base = ring + FWD_NST * FWD_STAGE;
  const uint32_t bar_full = bar_base, bar_empty = bar_full + 8 * FWD_NST, bar_w = bar_empty + 8 * FWD_NST;
  const uint32_t bar_tfull = bar_w + 8, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P.T, R_pad = P.R_pad;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  unsigned* ctr = P.ctr + (size_t)dom * CTR_STRIDE;
  int* ef = P.error_flag;
  bool dead = false;
  long long* tr = blockIdx.x == 0 ? P.trace : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&N.w_hi); tma_prefetch_desc(&N.w_lo); tma_prefetch_desc(&N.h_hi); tma_prefetch_desc(&N.h_lo);
    for (int s = 0; s < FWD_NST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CL); }
    mbar_init(bar_w, 1); mbar_init(bar_tfull, 1); mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the peers' barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_w, W_BYTES);
      for (int kc = 0; kc < KCH; ++kc) {
        tma_load_2d(w_base + kc * 2 * W_HALF, &N.w_hi, bar_w, kc * BK, slice * NC);
        tma_load_2d(w_base + kc * 2 * W_HALF + W_HALF, &N.w_lo, bar_w, kc * BK, slice * NC);
      }
      uint32_t it = 0;
      for (int t = 1; t < T; ++t) {
        wait_counter(ctr, (unsigned)(SLICES * t), ef, dead);   // h_{t-1} of this row block is complete
        HBL_STAMP(tr, t, 0);
        fence_async_global();
        for (int kc = 0; kc < KCH; ++kc, ++it) {
          const uint32_t s = it % FWD_NST, ph = (it / FWD_NST) & 1u;
          wait_bar(bar_empty + 8 * s, ph ^ 1u, ef, dead);
          if (dead) break;
          const uint32_t st = ring + s * FWD_STAGE;
          mbar_expect_tx(bar_full + 8 * s, FWD_STAGE);
          const int row0 = t * R_pad + mb * BM;                // block t of the h sequence = h_{t-1}
          if (CL == 1) {
            tma_load_2d(st, &N.h_hi, bar_full + 8 * s, kc * BK, row0);
            tma_load_2d(st + A_TILE, &N.h_lo, bar_full + 8 * s, kc * BK, row0);
          } else {                                             // this CTA's quarter of the rows, delivered to all four CTAs
            const uint32_t q = (uint32_t)rank * (A_TILE / CL);
            tma_load_2d_mc(st + q, &N.hq_hi, bar_full + 8 * s, kc * BK, row0 + rank * (BM / CL), CMASK);
            tma_load_2d_mc(st + A_TILE + q, &N.hq_lo, bar_full + 8 * s, kc * BK, row0 + rank * (BM / CL), CMASK);
          }
        }
        HBL_STAMP(tr, t, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // Two MMAs per K-step instead of three: the hi and lo halves of the weight chunk lie back to back in shared memory, i.e.
      // they ARE one [128 x 64] B tile -- h_hi x [W_hi | W_lo] lands in accumulator columns 0-63 (hi*hi) and 64-127 (hi*lo),
      // h_lo x W_hi adds to columns 0-63; the cell update adds the two blocks.  The N = 64 MMAs were bound by their operand
      // reads from shared memory (6 KB each): 14 instead of 18 KB per K-step, and a third fewer instructions.
      constexpr uint32_t idesc = idesc_mn(BM, NC), idesc2 = idesc_mn(BM, 2 * NC);
      wait_bar(bar_w, 0, ef, dead);
      uint32_t it = 0;
      for (int t = 1; t < T; ++t) {
        wait_bar(bar_tempty, (uint32_t)(t - 1) & 1u, ef, dead);   // the cell update of step t-1 has drained the accumulator
        tc_fence_after();
        uint32_t acc = 0;
        for (int kc = 0; kc < KCH; ++kc, ++it) {
          const uint32_t s = it % FWD_NST, ph = (it / FWD_NST) & 1u;
          wait_bar(bar_full + 8 * s, ph, ef, dead);
          if (dead) break;
          if (kc == 0) HBL_STAMP(tr, t, 2);
          tc_fence_after();
          const uint32_t st = ring + s * FWD_STAGE;
          const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + A_TILE);
          const uint64_t b_hi = make_desc_sw128(w_base + kc * 2 * W_HALF);   // rows 0-63: hi; as a 128-row tile: hi | lo
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
            umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc2, acc);
            umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
            acc = 1;
          }
          if (CL == 1) umma_commit(bar_empty + 8 * s);
          else umma_commit_mc(bar_empty + 8 * s, CMASK);        // the slot is free in ALL CTAs of the cluster only when all have read it
        }
        if (dead) break;
        umma_commit(bar_tfull);
        HBL_STAMP(tr, t, 3);
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, row = mb * BM + r;
    const bool valid = row < P.rows;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int unit = slice * UPC;
    float c[UPC];
#pragma unroll
    for (int i = 0; i < UPC; ++i) c[i] = 0.f;
    for (int t = 0; t < T; ++t) {
      const size_t grow = (size_t)t * R_pad + row;
      float g[NC];
      if (P.chunk_flags != nullptr && t % P.chunk == 0) {   // entering a new time chunk of the layer below's input projection
        if (lane == 0) wait_counter(P.chunk_flags + t / P.chunk, 1u, ef, dead, 4);
        dead = __shfl_sync(0xffffffffu, dead ? 1 : 0, 0) != 0;
      }
      if (valid) {
        // gx may have been written by a GEMM running concurrently (layer wavefront): L2, not the non-coherent path.  32-byte
        // accesses: a lane owns a row, so narrower ones would move half sectors (hb_gemm.cuh ldg256)
        const float* src = N.gx + grow * G4 + slice * NC;
#pragma unroll
        for (int i = 0; i < NC / 8; ++i) ldg256_cg(src + 8 * i, g + 8 * i);
      } else {
#pragma unroll
        for (int i = 0; i < NC; ++i) g[i] = 0.f;
      }
      if (t > 0) {
        wait_bar(bar_tfull, (uint32_t)(t - 1) & 1u, ef, dead);
        if (threadIdx.x == 64) HBL_STAMP(tr, t, 4);
        tc_fence_after();
        if (!dead) {
#pragma unroll
          for (int gate = 0; gate < 4; ++gate) {
            float v[16], v2[16];
            tmem_ld16(taddr + gate * UPC, v);
            tmem_ld16(taddr + NC + gate * UPC, v2);        // the hi*lo block
#pragma unroll
            for (int i = 0; i < 16; ++i) g[gate * UPC + i] += v[i] + v2[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
      float h[UPC];
#pragma unroll
      for (int i = 0; i < UPC; ++i) {
        const float ig = sigmoid_f(g[i]), fg = sigmoid_f(g[UPC + i]), gg = tanh_f(g[2 * UPC + i]), og = sigmoid_f(g[3 * UPC + i]);
        c[i] = fg * c[i] + ig * gg;
        h[i] = og * tanh_f(c[i]);
        g[i] = ig; g[UPC + i] = fg; g[2 * UPC + i] = gg; g[3 * UPC + i] = og;
      }
      // critical path first: the slice of h_t the other CTAs are waiting for, then the step counter
      if (valid && !dead) {
        const size_t o = ((size_t)(t + 1) * R_pad + row) * HIDN + unit;
        store_split16(h, N.hs_hi + o, N.hs_lo + o);
      }
      if (threadIdx.x == 64) HBL_STAMP(tr, t, 5);
      publish_step(ctr);
      if (threadIdx.x == 64) HBL_STAMP(tr, t, 6);
      // everything only later kernels read overlaps with the other CTAs' next step
      if (valid && !dead) {
        if (N.y) {
          float* dst = N.y + ((size_t)t * P.rows + row) * HIDN + unit;
          stg256(dst, h); stg256(dst + 8, h + 8);
        }
        if (N.act) {
          float* dst = N.act + grow * G4 + slice * NC;
#pragma unroll
          for (int i = 0; i < NC / 8; ++i) stg256(dst + 8 * i, g + 8 * i);
          float* dc = N.cs + grow * HIDN + unit;
          stg256(dc, c); stg256(dc + 8, c + 8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers may still be multicasting into / arriving on this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ backward recurrence
// grid = MB * 32 CTAs.  Warp 0: loads the CTA's slice of W_hh^T; warp 1: MMA issuer; warps 2-5: pointwise backward,
// A-operand staging, partial drain.  Steps run t = T-1 .. 0; the partial of step t is consumed at step t-1.
__global__ void __launch_bounds__(192, 1) lstm_bwd_kernel(const BwdParams* __restrict__ pp) {
  extern __shared__ uint8_t smem_raw[];
  const BwdParams& P = *pp;
  const int slice = blockIdx.x % SLICES, dom = blockIdx.x / SLICES, mb = dom;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wt_base = base, a_base = base + BWD_WT_BYTES, stage_base = a_base + 2 * A_TILE, bar_base = stage_base + 2 * BWD_STAGE;
  float* stage_ptr = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw)));
  const uint32_t bar_w = bar_base, bar_a = bar_w + 8, bar_tfull = bar_a + 8, tmem_slot = bar_tfull + 32;   // bar_tfull: one per accumulator quarter
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* a_ptr = smem_raw + (a_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P.T, R_pad = P.R_pad;
  unsigned* ctr = P.ctr + (size_t)dom * CTR_STRIDE;
  int* ef = P.error_flag;
  bool dead = false;
  long long* tr = (blockIdx.x == 0 && threadIdx.x == 64) ? P.trace : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.wt_hi); tma_prefetch_desc(&P.wt_lo);
    mbar_init(bar_w, 1); mbar_init(bar_a, 1);
    for (int q = 0; q < 4; ++q) mbar_init(bar_tfull + 8 * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_w, BWD_WT_BYTES);
      // B operand [N = 512 outputs][K = this CTA's 64 gate columns]: hi rows 0-255, hi rows 256-511, lo likewise
      tma_load_2d(wt_base + 0 * BWD_WT_HALF, &P.wt_hi, bar_w, slice * NC, 0);
      tma_load_2d(wt_base + 1 * BWD_WT_HALF, &P.wt_hi, bar_w, slice * NC, 256);
      tma_load_2d(wt_base + 2 * BWD_WT_HALF, &P.wt_lo, bar_w, slice * NC, 0);
      tma_load_2d(wt_base + 3 * BWD_WT_HALF, &P.wt_lo, bar_w, slice * NC, 256);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_mn(BM, 128);
      wait_bar(bar_w, 0, ef, dead);
      for (int t = T - 1; t >= 1; --t) {
        wait_bar(bar_a, (uint32_t)(T - 1 - t) & 1u, ef, dead);   // A tile of step t staged, accumulator of step t+1 drained
        if (dead) break;
        tc_fence_after();
        const uint64_t a_hi = make_desc_sw128(a_base), a_lo = make_desc_sw128(a_base + A_TILE);
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {   // four 128-column quarters of dh_{t-1}, committed one by one: the drain starts early
          const uint32_t boff = (uint32_t)(qt >> 1) * BWD_WT_HALF + (uint32_t)(qt & 1) * (128 * 128);   // 128 weight rows = 16 swizzle groups
          const uint64_t b_hi = make_desc_sw128(wt_base + boff), b_lo = make_desc_sw128(wt_base + 2 * BWD_WT_HALF + boff);
          const uint32_t d = tmem_base + qt * 128;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
            umma_bf16(d, a_hi + adv, b_lo + adv, idesc, k > 0 ? 1u : 0u);
            umma_bf16(d, a_lo + adv, b_hi + adv, idesc, 1);
            umma_bf16(d, a_hi + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(bar_tfull + 8 * qt);
        }
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, row = mb * BM + r;
    const bool valid = row < P.rows;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int unit = slice * UPC;
    float dc[UPC];
#pragma unroll
    for (int i = 0; i < UPC; ++i) dc[i] = 0.f;
    for (int t = T - 1; t >= 0; --t) {
      const size_t grow = (size_t)t * R_pad + row;
      float dh[UPC], a[NC], ct[UPC], cp[UPC];
      HBL_STAMP(tr, t, 0);
      if (P.chunk_flags != nullptr && (T - 1 - t) % P.chunk == 0) {   // entering a new time chunk of the layer above's gradient
        if (lane == 0) wait_counter(P.chunk_flags + (T - 1 - t) / P.chunk, 1u, ef, dead, 4);
        dead = __shfl_sync(0xffffffffu, dead ? 1 : 0, 0) != 0;
      }
      if (valid) {
        // 32-byte loads (a lane owns a row).  dh_ext may have been written by a kernel running concurrently (layer wavefront):
        // L2, not the non-coherent path
        const float* s0 = P.dh_ext + ((size_t)t * P.dh_rows + row) * HIDN + unit;
        const float* s1 = P.act + grow * G4 + slice * NC;
        const float* s2 = P.cs + grow * HIDN + unit;
        ldg256_cg(s0, dh); ldg256_cg(s0 + 8, dh + 8);
        ldg256_nc(s2, ct); ldg256_nc(s2 + 8, ct + 8);
        if (t > 0) { ldg256_nc(s2 - (size_t)R_pad * HIDN, cp); ldg256_nc(s2 - (size_t)R_pad * HIDN + 8, cp + 8); }
        else {
#pragma unroll
          for (int i = 0; i < UPC; ++i) cp[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NC / 8; ++i) ldg256_nc(s1 + 8 * i, a + 8 * i);
      } else {
#pragma unroll
        for (int i = 0; i < UPC; ++i) { dh[i] = 0.f; ct[i] = 0.f; cp[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < NC; ++i) a[i] = 0.f;
      }
      if (t < T - 1) {
        // dh_t += sum over the 32 CTAs' partials written at step t+1 (buffer (t+1)&1)
        if (lane == 0) wait_counter(ctr, (unsigned)(SLICES * (T - 1 - t)), ef, dead);
        dead = __shfl_sync(0xffffffffu, dead ? 1 : 0, 0) != 0;
        HBL_STAMP(tr, t, 1);
        if (!dead) {
          // the 32 CTAs of the row block ADDED their partials into one accumulator (bulk reductions at L2, see the drain below),
          // laid out [8 pieces][4 warps][64 columns][32 rows]: 16 coalesced reads, then clear the slice for the step after next
          float* accp = P.part + ((size_t)((t + 1) & 1) * P.MB + dom) * (size_t)(BM * HIDN) + (size_t)(unit / 64) * (64 * BM) + (size_t)q * (64 * 32) +
                        (size_t)(unit % 64) * 32 + lane;
#pragma unroll
          for (int i = 0; i < UPC; ++i) {
            dh[i] += __ldcg(accp + i * 32);
            __stcg(accp + i * 32, 0.f);
          }
        }
      }
      // pointwise backward of the cell (gate activations saved by the forward kernel)
      __align__(16) __nv_bfloat16 ghi[NC], glo[NC];
#pragma unroll
      for (int u = 0; u < UPC; ++u) {
        const float ig = a[u], fg = a[UPC + u], gg = a[2 * UPC + u], og = a[3 * UPC + u];
        const float tc = tanh_f(ct[u]);
        const float dct = dc[u] + dh[u] * og * (1.f - tc * tc);
        float d_i = dct * gg * ig * (1.f - ig);
        float d_f = dct * cp[u] * fg * (1.f - fg);
        float d_g = dct * ig * (1.f - gg * gg);
        float d_o = dh[u] * tc * og * (1.f - og);
        dc[u] = dct * fg;
        if (!valid) { d_i = d_f = d_g = d_o = 0.f; dc[u] = 0.f; }
        split_bf16(d_i, ghi[u], glo[u]);
        split_bf16(d_f, ghi[UPC + u], glo[UPC + u]);
        split_bf16(d_g, ghi[2 * UPC + u], glo[2 * UPC + u]);
        split_bf16(d_o, ghi[3 * UPC + u], glo[3 * UPC + u]);
      }
      if (t > 0) {
        // stage this CTA's dgate slice as the A operand [128 rows][64 K] (128-byte swizzle, what TMA would have written)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(a_ptr + off) = reinterpret_cast<const uint4*>(ghi)[ch];
          *reinterpret_cast<uint4*>(a_ptr + A_TILE + off) = reinterpret_cast<const uint4*>(glo)[ch];
        }
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) mbar_arrive(bar_a);
        HBL_STAMP(tr, t, 2);
      }
      // while the tensor core works: the row-major dgate copy the dX / dW GEMMs read after this kernel
      if (valid && !dead) {
        __nv_bfloat16* d0 = P.dg_hi + grow * G4 + slice * NC;
        __nv_bfloat16* d1 = P.dg_lo + grow * G4 + slice * NC;
#pragma unroll
        for (int i = 0; i < NC / 16; ++i) {
          stg256(d0 + 16 * i, reinterpret_cast<const uint32_t*>(ghi) + 8 * i);
          stg256(d1 + 16 * i, reinterpret_cast<const uint32_t*>(glo) + 8 * i);
        }
      }
      if (t == 0) break;
      // Drain: the [128 x 512] fp32 partial of dh_{t-1} leaves TMEM in eight 64-column pieces.  Every WARP runs its own pipeline
      // over its 32 rows: stage [64 columns][32 rows] in shared memory (conflict-free: consecutive lanes = consecutive rows),
      // ADD it into the row block's accumulator with one bulk reduction (cp.reduce.async.bulk .add.f32: the 32-way sum happens at
      // L2, a warp issues 8 instructions per step instead of 4096 vector atomics), two staging buffers per warp -- piece p+2
      // waits until the reduction of piece p has read its buffer.  No CTA barrier and no single issuing thread inside the drain.
      float* dst = P.part + ((size_t)(t & 1) * P.MB + dom) * (size_t)(BM * HIDN) + (size_t)q * (64 * 32);
#pragma unroll 1
      for (int qt = 0; qt < 4; ++qt) {
        wait_bar(bar_tfull + 8 * qt, (uint32_t)(T - 1 - t) & 1u, ef, dead);
        if ((qt & 1) == 0) HBL_STAMP(tr, t, 3 + qt);
        tc_fence_after();
#pragma unroll 1
        for (int q4 = 0; q4 < 2; ++q4) {
          const int pc = qt * 2 + q4, buf = pc & 1;
          uint32_t v[64];
          if (!dead) {
            tmem_ld32_nowait(taddr + pc * 64, v);
            tmem_ld32_nowait(taddr + pc * 64 + 32, v + 32);
            tmem_wait_ld();
          }
          if (pc >= 2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
          const uint32_t sb_off = (uint32_t)(q * 2 + buf) * (64 * 32 * 4);
          float* sb = stage_ptr + sb_off / 4 + lane;
#pragma unroll
          for (int i = 0; i < 64; ++i) sb[i * 32] = dead ? 0.f : __uint_as_float(v[i]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (size_t)pc * (64 * BM)),
                         "r"(stage_base + sb_off), "r"(64 * 32 * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (qt & 1) HBL_STAMP(tr, t, 3 + qt);
      }
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's eight reductions have been performed
        fence_async_global();
      }
      tc_fence_before();
      fence_async_global();   // the dgate rows of this step are read through TMA by the layer wavefront's GEMM while this kernel runs
      asm volatile("bar.sync 1, 128;" ::: "memory");
      HBL_STAMP(tr, t, 7);
      if (threadIdx.x == 64) {
        HBL_STAMP(tr, t, 8);
        fence_async_global();
        red_release_add(ctr, 1u);
        HBL_STAMP(tr, t, 9);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Layer-wavefront plumbing (one thread each): block a stream until every row block of a recurrence kernel that is still
// RUNNING has completed `target` step publications; raise a flag the other layer's kernel polls.
__global__ void lstm_wait_steps(const unsigned* __restrict__ ctr, int MB, unsigned target, int* error_flag) {
  bool dead = false;
  for (int d = 0; d < MB; ++d) wait_counter(ctr + (size_t)d * CTR_STRIDE, target, error_flag, dead, 3);
}
__global__ void lstm_set_flag(unsigned* flag) {
  __threadfence();
  fence_async_global();
  atomicExch(flag, 1u);
}

// [rows][cols] bf16 pair -> [cols][ld_dst] (column r of the destination = row r of the source); 64 x 64 tiles, 8-byte global
// accesses on both sides (four bf16 per thread and access: the 2-byte version moved 2.2 TB/s, latency bound)
__global__ void __launch_bounds__(256) lstm_transpose_pair(const __nv_bfloat16* __restrict__ s_hi, const __nv_bfloat16* __restrict__ s_lo, int cols,
                                                           __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, long long ld_dst) {
  __shared__ __align__(8) unsigned short th[64][66], tl[64][66];
  const size_t r0 = (size_t)blockIdx.y * 64;
  const int c0 = blockIdx.x * 64, q = threadIdx.x & 15, jb = threadIdx.x >> 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = jb + 16 * k;
    const uint2 a = *reinterpret_cast<const uint2*>(s_hi + (r0 + j) * cols + c0 + 4 * q);
    const uint2 b = *reinterpret_cast<const uint2*>(s_lo + (r0 + j) * cols + c0 + 4 * q);
    uint32_t* ph = reinterpret_cast<uint32_t*>(&th[j][4 * q]);
    uint32_t* pl = reinterpret_cast<uint32_t*>(&tl[j][4 * q]);
    ph[0] = a.x; ph[1] = a.y; pl[0] = b.x; pl[1] = b.y;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = jb + 16 * k;   // destination row = source column
    uint2 a, b;
    a.x = (uint32_t)th[4 * q][c] | ((uint32_t)th[4 * q + 1][c] << 16); a.y = (uint32_t)th[4 * q + 2][c] | ((uint32_t)th[4 * q + 3][c] << 16);
    b.x = (uint32_t)tl[4 * q][c] | ((uint32_t)tl[4 * q + 1][c] << 16); b.y = (uint32_t)tl[4 * q + 2][c] | ((uint32_t)tl[4 * q + 3][c] << 16);
    *reinterpret_cast<uint2*>(d_hi + (size_t)(c0 + c) * ld_dst + r0 + 4 * q) = a;
    *reinterpret_cast<uint2*>(d_lo + (size_t)(c0 + c) * ld_dst + r0 + 4 * q) = b;
  }
}

// ------------------------------------------------------------------------------------------------ operand preparation
// CTA-slice order of the 2048 gate rows / columns: p = slice*64 + gate*16 + u  <->  nn.LSTM row gate*512 + slice*16 + u.
__device__ __forceinline__ int perm_to_orig(int p) { return ((p >> 4) & 3) * HIDN + (p >> 6) * UPC + (p & 15); }

// W [2048][512] fp32 (nn.LSTM layout) -> Wp hi/lo [2048 slice order][512]  and (optional) WT hi/lo [512][2048 slice order]
__global__ void lstm_prep_weight(const float* __restrict__ w, __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo,
                                 __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.y * 32, k0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;  // block (32, 8)
  for (int j = ty; j < 32; j += 8) {
    const float v = w[(size_t)perm_to_orig(p0 + j) * HIDN + k0 + tx];
    tile[j][tx] = v;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    p_hi[(size_t)(p0 + j) * HIDN + k0 + tx] = h;
    p_lo[(size_t)(p0 + j) * HIDN + k0 + tx] = l;
  }
  if (!t_hi) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    __nv_bfloat16 h, l;
    split_bf16(tile[tx][j], h, l);
    t_hi[(size_t)(k0 + j) * G4 + p0 + tx] = h;
    t_lo[(size_t)(k0 + j) * G4 + p0 + tx] = l;
  }
}
__global__ void lstm_prep_bias(const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < G4) { const int o = perm_to_orig(p); out[p] = b_ih[o] + b_hh[o]; }
}

// x [T][rows][512] fp32 -> xs hi/lo [T*R_pad][512] and (optional) xT hi/lo [512][ldT] at column t*R_pad + row
__global__ void lstm_prep_x(const float* __restrict__ x, int T, int rows, int R_pad, __nv_bfloat16* __restrict__ s_hi,
                            __nv_bfloat16* __restrict__ s_lo, __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo, long long ldT) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z, r0 = blockIdx.y * 32, k0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const int row = r0 + j;
    const float v = row < rows ? x[((size_t)t * rows + row) * HIDN + k0 + tx] : 0.f;
    tile[j][tx] = v;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    if (row < R_pad) {
      s_hi[((size_t)t * R_pad + row) * HIDN + k0 + tx] = h;
      s_lo[((size_t)t * R_pad + row) * HIDN + k0 + tx] = l;
    }
  }
  if (!t_hi) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int row = r0 + tx;
    if (row >= R_pad) continue;
    __nv_bfloat16 h, l;
    split_bf16(tile[tx][j], h, l);
    t_hi[(size_t)(k0 + j) * ldT + (size_t)t * R_pad + row] = h;
    t_lo[(size_t)(k0 + j) * ldT + (size_t)t * R_pad + row] = l;
  }
}

// db[orig row] = sum over all (t, row) of dgate column p: one block per slice-order column, reading the transposed copy
__global__ void lstm_bias_grad(const __nv_bfloat16* __restrict__ t_hi, const __nv_bfloat16* __restrict__ t_lo, long long n,
                               float* __restrict__ db_ih, float* __restrict__ db_hh) {
  __shared__ float red[8];
  const int p = blockIdx.x;
  const __nv_bfloat16 *a = t_hi + (size_t)p * n, *b = t_lo + (size_t)p * n;
  float s = 0.f;
  // n = T * R_pad is a multiple of 128: 16-byte loads (eight bf16 per thread and step)
  for (long long i = (long long)threadIdx.x * 8; i < n; i += (long long)blockDim.x * 8) {
    const uint4 va = *reinterpret_cast<const uint4*>(a + i), vb = *reinterpret_cast<const uint4*>(b + i);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&va);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&vb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __bfloat1622float2(pa[j]), fb = __bfloat1622float2(pb[j]);
      s += (fa.x + fb.x) + (fa.y + fb.y);
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const int o = perm_to_orig(p);
    db_ih[o] = tot;
    db_hh[o] = tot;
  }
}

// dWp [2048 slice order][512] -> dW [2048][512] (nn.LSTM row order)
__global__ void lstm_unperm_rows(const float* __restrict__ src, float* __restrict__ dst) {
  const int p = blockIdx.x;
  const int o = perm_to_orig(p);
  reinterpret_cast<float4*>(dst + (size_t)o * HIDN)[threadIdx.x] = reinterpret_cast<const float4*>(src + (size_t)p * HIDN)[threadIdx.x];  // 128 threads
}

// padded [T*R_pad][512] -> caller's [T][rows][512]
__global__ void lstm_unpad(const float* __restrict__ src, float* __restrict__ dst, int rows, int R_pad) {
  const int t = blockIdx.y, row = blockIdx.x;
  reinterpret_cast<float4*>(dst + ((size_t)t * rows + row) * HIDN)[threadIdx.x] =
      reinterpret_cast<const float4*>(src + ((size_t)t * R_pad + row) * HIDN)[threadIdx.x];
}

}  // namespace hbl

// ---------------------------------------------------------------------------------------------------- host side
struct HbLstmNetBuf {                       // per network (0: the one that may be saved for backward, 1: forward only)
  __nv_bfloat16 *wih_hi[2], *wih_lo[2], *whh_hi[2], *whh_lo[2];     // [2048p][512]
  __nv_bfloat16 *wihT_hi[2], *wihT_lo[2], *whhT_hi[2], *whhT_lo[2]; // [512][2048p] (net 0 only)
  float* bias[2];                           // [2048p]
  __nv_bfloat16 *xs_hi, *xs_lo;             // [N][512]
  __nv_bfloat16 *hs_hi[2], *hs_lo[2];       // [(T+1)*R_pad][512] per layer
  float* gx;                                // [N][2048] input projection of layer 0 (and of layer 1 without the layer wavefront)
  float* gx1;                               // [N][2048] input projection of layer 1 (layer wavefront: both are live at once)
};

struct hb_lstm {
  int device, sm_count, max_T, max_rows, max_rpad;
  int T, rows, R_pad, MB;                   // geometry of the last forward
  int saved;                                // 1: net 0's activations of the last forward are valid for backward
  int use_clusters, last_cluster;           // forward recurrence as 4-CTA multicast clusters when they all fit on the device
  int zero_rpad, zero_rows_max;             // row geometry for which the h-sequence buffers are known to be clean
  HbLstmNetBuf nb[2];
  // saved for backward (net 0)
  __nv_bfloat16 *xT_hi, *xT_lo;             // [512][ldT]
  __nv_bfloat16 *hsT_hi[2], *hsT_lo[2];     // [512][ldT]
  float *act[2], *cs[2];
  // backward scratch
  __nv_bfloat16 *dg_hi[2], *dg_lo[2], *dgT_hi[2], *dgT_lo[2];
  float* dh0;                               // [N][512] gradient w.r.t. layer 0's output
  float* dx_pad;                            // [N][512]
  float* part;                              // [2 layers][2][R_pad][512] dh accumulators of the backward recurrences
  float* dwp;                               // [4][2048][512]
  unsigned* ctr;
  unsigned* chunk_flags;                    // [64] layer-wavefront progress flags
  long long* d_trace;                       // diagnostic (HB_LSTM_TRACE=<file>): [4 kernels][max_T][16] time stamps
  HbUploadRing upload;                      // pinned staging of launch-parameter records (hb_gemm_host.h)
  cudaStream_t ws[4];                       // internal streams of the layer wavefront (recurrence above / chunk GEMMs / recurrence below / dW chunks)
  cudaEvent_t wev[6];
  int use_wavefront;
  int* d_error;
  int* h_error;                             // pinned mirror of d_error, filled asynchronously at the end of every call
  cudaEvent_t ev_done;
  int check_pending;
  hbl::FwdParams* d_fwd;                    // [2 layers]
  hbl::BwdParams* d_bwd;                    // [2 layers]
  Params* d_gemm;                           // [16]
  int64_t launches;
};

#define HBL_ALLOC(ptr, bytes)                                             \
  do {                                                                    \
    HB_CUDA(cudaMalloc((void**)&(ptr), (bytes)));                         \
    HB_CUDA(cudaMemset((ptr), 0, (bytes)));                               \
  } while (0)

static void hbl_gemm_problem(Params& p, int& rc, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, uint64_t m, uint64_t lda,
                             const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, uint64_t n, uint64_t ldb, uint64_t k, int cl,
                             const float* bias, float* c, int ldc, int* err) {
  memset(&p, 0, sizeof(p));
  rc |= hb_make_tmap(&p.a_hi[0], a_hi, m, k, BM, lda);
  rc |= hb_make_tmap(&p.a_lo[0], a_lo, m, k, BM, lda);
  p.a_hi[1] = p.a_hi[0]; p.a_lo[1] = p.a_lo[0];
  rc |= hb_make_tmap(&p.b_hi, b_hi, n, k, BN / cl, ldb);
  rc |= hb_make_tmap(&p.b_lo, b_lo, n, k, BN / cl, ldb);
  p.k_chunks = (int)(k / BK); p.k_chunks_seg0 = p.k_chunks; p.lo_first = 0; p.lo_last = p.k_chunks; p.split = 1;
  p.bias = bias; p.c_f32 = c; p.ldc = ldc; p.error_flag = err;
  p.row_mul = 1; p.row_add = 0; p.valid_rows = (int)m;
}

static int hbl_run_gemm(hb_lstm* L, cudaStream_t st, const Params* hp, int nprob, int mt, int nt, int slot, int sm_limit = 0) {
  { const int urc = hb_upload(&L->upload, L->d_gemm + slot, hp, nprob * sizeof(Params), st); if (urc) return urc; }
  const bool pair = (mt % 2) == 0;
  const int sms = sm_limit > 0 ? sm_limit : L->sm_count;   // a limit keeps the persistent grid off the SMs a co-running recurrence needs
  L->launches += 1;
  if (pair) return hb_launch_gemm(gemm3_kernel<EPI_F32, 3>, 2, sms, st, L->d_gemm + slot, nt, mt, nprob);
  return hb_launch_gemm(gemm3_kernel<EPI_F32, 1>, 1, sms, st, L->d_gemm + slot, nt, mt, nprob);
}

extern "C" {

int hb_lstm_create(int device, int max_T, int max_rows, hb_lstm** out) {
  if (!out) { hb_set_error("hb_lstm_create: null argument"); return -1; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { hb_set_error("hb_lstm_create: no CUDA device -- libhanabi_b200 has no CPU path"); return -2; }
  if (device < 0 || device >= ndev) { hb_set_error("hb_lstm_create: bad device ordinal"); return -1; }
  if (max_T < 1 || max_T > 4096 || max_rows < 1 || max_rows > 512) { hb_set_error("hb_lstm_create: need 1 <= max_T <= 4096, 1 <= max_rows <= 512"); return -1; }
  HB_CUDA(cudaSetDevice(device));
  hb_lstm* L = new hb_lstm();
  memset(L, 0, sizeof(*L));
  L->device = device; L->max_T = max_T; L->max_rows = max_rows;
  L->max_rpad = (max_rows + BM - 1) / BM * BM;
  cudaDeviceProp prop;
  HB_CUDA(cudaGetDeviceProperties(&prop, device));
  L->sm_count = prop.multiProcessorCount;
  const size_t bf = sizeof(__nv_bfloat16), N = (size_t)max_T * L->max_rpad, N1 = (size_t)(max_T + 1) * L->max_rpad, WN = (size_t)hbl::G4 * hbl::HIDN;
  for (int n = 0; n < 2; ++n) {
    HbLstmNetBuf& B = L->nb[n];
    for (int l = 0; l < 2; ++l) {
      HBL_ALLOC(B.wih_hi[l], WN * bf); HBL_ALLOC(B.wih_lo[l], WN * bf); HBL_ALLOC(B.whh_hi[l], WN * bf); HBL_ALLOC(B.whh_lo[l], WN * bf);
      if (n == 0) { HBL_ALLOC(B.wihT_hi[l], WN * bf); HBL_ALLOC(B.wihT_lo[l], WN * bf); HBL_ALLOC(B.whhT_hi[l], WN * bf); HBL_ALLOC(B.whhT_lo[l], WN * bf); }
      HBL_ALLOC(B.bias[l], hbl::G4 * sizeof(float));
      HBL_ALLOC(B.hs_hi[l], N1 * hbl::HIDN * bf); HBL_ALLOC(B.hs_lo[l], N1 * hbl::HIDN * bf);
    }
    HBL_ALLOC(B.xs_hi, N * hbl::HIDN * bf); HBL_ALLOC(B.xs_lo, N * hbl::HIDN * bf);
    HBL_ALLOC(B.gx, N * hbl::G4 * sizeof(float));
    HBL_ALLOC(B.gx1, N * hbl::G4 * sizeof(float));
  }
  HBL_ALLOC(L->xT_hi, N1 * hbl::HIDN * bf); HBL_ALLOC(L->xT_lo, N1 * hbl::HIDN * bf);
  for (int l = 0; l < 2; ++l) {
    HBL_ALLOC(L->hsT_hi[l], N1 * hbl::HIDN * bf); HBL_ALLOC(L->hsT_lo[l], N1 * hbl::HIDN * bf);
    HBL_ALLOC(L->act[l], N * hbl::G4 * sizeof(float)); HBL_ALLOC(L->cs[l], N * hbl::HIDN * sizeof(float));
    HBL_ALLOC(L->dg_hi[l], N * hbl::G4 * bf); HBL_ALLOC(L->dg_lo[l], N * hbl::G4 * bf);
    HBL_ALLOC(L->dgT_hi[l], N * hbl::G4 * bf); HBL_ALLOC(L->dgT_lo[l], N * hbl::G4 * bf);
  }
  HBL_ALLOC(L->dh0, N * hbl::HIDN * sizeof(float));
  HBL_ALLOC(L->dx_pad, N * hbl::HIDN * sizeof(float));
  HBL_ALLOC(L->part, (size_t)2 * 2 * L->max_rpad * hbl::HIDN * sizeof(float));
  HBL_ALLOC(L->chunk_flags, 64 * sizeof(unsigned));
  for (int i = 0; i < 4; ++i) HB_CUDA(cudaStreamCreateWithFlags(&L->ws[i], cudaStreamNonBlocking));
  for (int i = 0; i < 6; ++i) HB_CUDA(cudaEventCreateWithFlags(&L->wev[i], cudaEventDisableTiming));
  L->use_wavefront = getenv("HB_LSTM_NO_WAVEFRONT") ? 0 : 1;   // diagnostic switch: the two layers' recurrences one after the other
  if (getenv("HB_LSTM_TRACE")) HBL_ALLOC(L->d_trace, (size_t)4 * max_T * 16 * sizeof(long long));
  {
    // CUDA loads kernels lazily, and loading one may wait for the device to go idle.  Inside the layer wavefront a first-time
    // launch would then wait for a recurrence kernel that is itself waiting for what that launch produces: load everything now.
    cudaFuncAttributes fa;
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_wait_steps));
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_set_flag));
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_transpose_pair));
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_unpad));
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_bias_grad));
    HB_CUDA(cudaFuncGetAttributes(&fa, hbl::lstm_unperm_rows));
  }
  HBL_ALLOC(L->dwp, 4 * WN * sizeof(float));
  HBL_ALLOC(L->ctr, 16 * hbl::CTR_STRIDE * sizeof(unsigned));
  HBL_ALLOC(L->d_error, sizeof(int));
  HB_CUDA(cudaMallocHost((void**)&L->h_error, sizeof(int)));
  *L->h_error = 0;
  HB_CUDA(cudaEventCreateWithFlags(&L->ev_done, cudaEventDisableTiming));
  HBL_ALLOC(L->d_fwd, 2 * sizeof(hbl::FwdParams));
  HBL_ALLOC(L->d_bwd, 2 * sizeof(hbl::BwdParams));
  HBL_ALLOC(L->d_gemm, 64 * sizeof(Params));
  { const int urc = hb_upload_init(&L->upload); if (urc) return urc; }
  HB_CUDA((cudaFuncSetAttribute(hbl::lstm_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbl::FWD_SMEM)));
  HB_CUDA((cudaFuncSetAttribute(hbl::lstm_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbl::FWD_SMEM)));
  HB_CUDA((cudaFuncSetAttribute(hbl::lstm_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbl::FWD_SMEM)));
  L->use_clusters = getenv("HB_LSTM_NO_CLUSTER") ? 0 : (getenv("HB_LSTM_CL") ? atoi(getenv("HB_LSTM_CL")) : 4);   // diagnostic: multicast cluster size of the forward recurrence (0 / 1: none)
  HB_CUDA((cudaFuncSetAttribute(hbl::lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hbl::BWD_SMEM)));
  HB_CUDA((cudaFuncSetAttribute(gemm3_kernel<EPI_F32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  HB_CUDA((cudaFuncSetAttribute(gemm3_kernel<EPI_F32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  *out = L;
  return 0;
}

void hb_lstm_destroy(hb_lstm* L) {
  if (!L) return;
  cudaSetDevice(L->device);
  for (int n = 0; n < 2; ++n) {
    HbLstmNetBuf& B = L->nb[n];
    for (int l = 0; l < 2; ++l) {
      cudaFree(B.wih_hi[l]); cudaFree(B.wih_lo[l]); cudaFree(B.whh_hi[l]); cudaFree(B.whh_lo[l]);
      cudaFree(B.wihT_hi[l]); cudaFree(B.wihT_lo[l]); cudaFree(B.whhT_hi[l]); cudaFree(B.whhT_lo[l]);
      cudaFree(B.bias[l]); cudaFree(B.hs_hi[l]); cudaFree(B.hs_lo[l]);
    }
    cudaFree(B.xs_hi); cudaFree(B.xs_lo); cudaFree(B.gx); cudaFree(B.gx1);
  }
  cudaFree(L->xT_hi); cudaFree(L->xT_lo);
  for (int l = 0; l < 2; ++l) {
    cudaFree(L->hsT_hi[l]); cudaFree(L->hsT_lo[l]); cudaFree(L->act[l]); cudaFree(L->cs[l]);
    cudaFree(L->dg_hi[l]); cudaFree(L->dg_lo[l]); cudaFree(L->dgT_hi[l]); cudaFree(L->dgT_lo[l]);
  }
  cudaFree(L->dh0); cudaFree(L->dx_pad); cudaFree(L->part); cudaFree(L->dwp); cudaFree(L->ctr); cudaFree(L->d_error);
  cudaFree(L->d_fwd); cudaFree(L->d_bwd); cudaFree(L->d_gemm); cudaFree(L->chunk_flags); cudaFree(L->d_trace);
  for (int i = 0; i < 4; ++i) if (L->ws[i]) cudaStreamDestroy(L->ws[i]);
  for (int i = 0; i < 6; ++i) if (L->wev[i]) cudaEventDestroy(L->wev[i]);
  cudaFreeHost(L->h_error); cudaEventDestroy(L->ev_done);
  hb_upload_destroy(&L->upload);
  delete L;
}

// The calls are asynchronous (work is queued on the caller's stream).  The kernels' spin guards raise d_error; its value
// travels to a pinned mirror at the end of every call and is examined at the START of the next one (or by hb_lstm_sync).
// Diagnostic: append the time stamps CTA 0 of each recurrence kernel took (ns, %globaltimer) to the file HB_LSTM_TRACE names.
static void hbl_dump_trace(hb_lstm* L, cudaStream_t st, int first, int T, const char* tag) {
  if (!L->d_trace) return;
  cudaStreamSynchronize(st);
  std::vector<long long> h((size_t)2 * L->max_T * 16);
  cudaMemcpy(h.data(), L->d_trace + (size_t)first * L->max_T * 16, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(L->d_trace + (size_t)first * L->max_T * 16, 0, h.size() * sizeof(long long));
  FILE* f = fopen(getenv("HB_LSTM_TRACE"), "a");
  if (!f) return;
  for (int l = 0; l < 2; ++l)
    for (int t = 0; t < T; ++t) {
      fprintf(f, "%s layer %d t %d", tag, l, t);
      for (int k = 0; k < 12; ++k) fprintf(f, " %lld", h[((size_t)l * L->max_T + t) * 16 + k]);
      fprintf(f, "\n");
    }
  fclose(f);
}

static int hbl_finish_call(hb_lstm* L, cudaStream_t st) {
  HB_CUDA(cudaMemcpyAsync(L->h_error, L->d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaEventRecord(L->ev_done, st));
  L->check_pending = 1;
  return 0;
}
static int hbl_check_previous(hb_lstm* L, bool wait) {
  if (!L->check_pending) return 0;
  if (wait) HB_CUDA(cudaEventSynchronize(L->ev_done));
  else if (cudaEventQuery(L->ev_done) != cudaSuccess) { (void)cudaGetLastError(); return 0; }   // still running: examined later
  L->check_pending = 0;
  if (*L->h_error) {
    const int code = *L->h_error;
    *L->h_error = 0;
    cudaMemset(L->d_error, 0, sizeof(int));
    hb_set_error("hb_lstm: a pipeline / step barrier of an earlier call timed out (spin guard, code %d)", code);
    return -4;
  }
  return 0;
}

int hb_lstm_forward(hb_lstm* L, int T, int rows, int nets, const float* const* x, const hb_lstm_weights* w, float* const* y, int save,
                    void* stream) {
  if (!L || !x || !w || !y) { hb_set_error("hb_lstm_forward: null argument"); return -1; }
  if (T < 1 || T > L->max_T || rows < 1 || rows > L->max_rows) { hb_set_error("hb_lstm_forward: T=%d rows=%d exceed the handle's (%d, %d)", T, rows, L->max_T, L->max_rows); return -1; }
  if (nets < 1 || nets > 2) { hb_set_error("hb_lstm_forward: nets must be 1 or 2"); return -1; }
  for (int n = 0; n < nets; ++n) {
    if (!x[n] || !y[n]) { hb_set_error("hb_lstm_forward: null sequence pointer"); return -1; }
    if (reinterpret_cast<uintptr_t>(y[n]) & 31) { hb_set_error("hb_lstm_forward: y[%d] must be 32-byte aligned (the kernels write rows with 32-byte stores)", n); return -1; }
    for (int l = 0; l < 2; ++l)
      if (!w[n].w_ih[l] || !w[n].w_hh[l] || !w[n].b_ih[l] || !w[n].b_hh[l]) { hb_set_error("hb_lstm_forward: null weight pointer"); return -1; }
  }
  HB_CUDA(cudaSetDevice(L->device));
  { const int prc = hbl_check_previous(L, false); if (prc) return prc; }
  cudaStream_t st = (cudaStream_t)stream;
  const int R_pad = (rows + BM - 1) / BM * BM, MB = R_pad / BM;
  const long long ldT = (long long)(T + 1) * R_pad;
  const size_t bf = sizeof(__nv_bfloat16);
  L->T = T; L->rows = rows; L->R_pad = R_pad; L->MB = MB; L->saved = 0;
  if (nets * MB * hbl::SLICES > L->sm_count) { hb_set_error("hb_lstm_forward: %d networks x %d rows need %d co-resident CTAs, the device has %d SMs", nets, rows, nets * MB * hbl::SLICES, L->sm_count); return -1; }
  // ---- operands
  for (int n = 0; n < nets; ++n) {
    HbLstmNetBuf& B = L->nb[n];
    const bool sv = save && n == 0;
    for (int l = 0; l < 2; ++l) {
      hbl::lstm_prep_weight<<<dim3(hbl::HIDN / 32, hbl::G4 / 32), dim3(32, 8), 0, st>>>(w[n].w_ih[l], B.wih_hi[l], B.wih_lo[l], sv ? B.wihT_hi[l] : nullptr, sv ? B.wihT_lo[l] : nullptr);
      hbl::lstm_prep_weight<<<dim3(hbl::HIDN / 32, hbl::G4 / 32), dim3(32, 8), 0, st>>>(w[n].w_hh[l], B.whh_hi[l], B.whh_lo[l], sv ? B.whhT_hi[l] : nullptr, sv ? B.whhT_lo[l] : nullptr);
      hbl::lstm_prep_bias<<<hbl::G4 / 256, 256, 0, st>>>(w[n].b_ih[l], w[n].b_hh[l], B.bias[l]);
      // block 0 of the h sequence is h_{-1} = 0 (and stays 0: the kernels never write it); padded rows stay 0 as well
    }
    hbl::lstm_prep_x<<<dim3(hbl::HIDN / 32, R_pad / 32, T), dim3(32, 8), 0, st>>>(x[n], T, rows, R_pad, B.xs_hi, B.xs_lo, sv ? L->xT_hi : nullptr, sv ? L->xT_lo : nullptr, ldT);
    L->launches += 7;
  }
  // Block 0 of the h sequences (h_{-1} = 0) and the padded rows must read as zero.  The kernels never write either, so the
  // buffers (zeroed at creation) only need clearing when the row geometry changes or rows were valid in an earlier call.
  if (R_pad != L->zero_rpad || rows < L->zero_rows_max) {
    for (int n = 0; n < 2; ++n)
      for (int l = 0; l < 2; ++l) {
        HB_CUDA(cudaMemsetAsync(L->nb[n].hs_hi[l], 0, (size_t)(L->max_T + 1) * L->max_rpad * hbl::HIDN * bf, st));
        HB_CUDA(cudaMemsetAsync(L->nb[n].hs_lo[l], 0, (size_t)(L->max_T + 1) * L->max_rpad * hbl::HIDN * bf, st));
      }
    L->zero_rpad = R_pad; L->zero_rows_max = rows;
  }
  if (rows > L->zero_rows_max) L->zero_rows_max = rows;
  HB_CUDA(cudaMemsetAsync(L->ctr, 0, 16 * hbl::CTR_STRIDE * sizeof(unsigned), st));
  std::vector<hbl::FwdParams> fp(2);
  memset(fp.data(), 0, 2 * sizeof(hbl::FwdParams));
  int rc = 0;
  const int n_ctas = nets * MB * hbl::SLICES;
  // Layer wavefront (see hb_lstm_backward): layer 1's recurrence runs next to layer 0's, one time chunk behind; its input
  // projection is computed chunk by chunk on the SMs the two recurrences leave free.
  const int gemm_sms = (L->sm_count - 2 * n_ctas) & ~1;
  int chunk = 8;
  while ((T + chunk - 1) / chunk > 32) chunk *= 2;
  const int n_chunks = (T + chunk - 1) / chunk;
  const bool wave = L->use_wavefront && gemm_sms >= 8 && T >= 2 * chunk;
  // input projection of layer l over the steps [t0, t1), both networks as problems of one launch
  auto gx_gemm = [&](int l, int t0, int t1, cudaStream_t s, int slot, int sm_limit) -> int {
    Params gp[2];
    int r2 = 0;
    const size_t r0 = (size_t)t0 * R_pad, nr = (size_t)(t1 - t0) * R_pad;
    for (int n = 0; n < nets; ++n) {
      HbLstmNetBuf& B = L->nb[n];
      const __nv_bfloat16* a_hi = l == 0 ? B.xs_hi : B.hs_hi[0] + (size_t)R_pad * hbl::HIDN;   // layer 1 consumes h^0_t = block t+1
      const __nv_bfloat16* a_lo = l == 0 ? B.xs_lo : B.hs_lo[0] + (size_t)R_pad * hbl::HIDN;
      float* dst = (l == 0 || !wave) ? B.gx : B.gx1;
      hbl_gemm_problem(gp[n], r2, a_hi + r0 * hbl::HIDN, a_lo + r0 * hbl::HIDN, nr, hbl::HIDN, B.wih_hi[l], B.wih_lo[l], hbl::G4, hbl::HIDN, hbl::HIDN,
                       ((nr / BM) % 2 == 0) ? 2 : 1, B.bias[l], dst + r0 * hbl::G4, hbl::G4, L->d_error);
    }
    if (r2) return -2;
    return hbl_run_gemm(L, s, gp, nets, (int)(nr / BM), hbl::G4 / BN, slot, sm_limit);
  };
  for (int l = 0; l < 2; ++l) {
    hbl::FwdParams& F = fp[l];
    F.T = T; F.rows = rows; F.R_pad = R_pad; F.MB = MB; F.ldT = ldT; F.ctr = L->ctr + (size_t)l * 8 * hbl::CTR_STRIDE; F.error_flag = L->d_error;
    F.chunk_flags = (wave && l == 1) ? L->chunk_flags : nullptr; F.chunk = chunk;
    F.trace = L->d_trace ? L->d_trace + (size_t)l * L->max_T * 16 : nullptr;
    for (int n = 0; n < nets; ++n) {
      HbLstmNetBuf& B = L->nb[n];
      hbl::FwdNet& Q = F.net[n];
      const bool sv = save && n == 0;
      rc |= hb_make_tmap(&Q.w_hi, B.whh_hi[l], hbl::G4, hbl::HIDN, hbl::NC);
      rc |= hb_make_tmap(&Q.w_lo, B.whh_lo[l], hbl::G4, hbl::HIDN, hbl::NC);
      rc |= hb_make_tmap(&Q.h_hi, B.hs_hi[l], (uint64_t)(T + 1) * R_pad, hbl::HIDN, BM);
      rc |= hb_make_tmap(&Q.h_lo, B.hs_lo[l], (uint64_t)(T + 1) * R_pad, hbl::HIDN, BM);
      Q.gx = (l == 0 || !wave) ? B.gx : B.gx1; Q.hs_hi = B.hs_hi[l]; Q.hs_lo = B.hs_lo[l];
      const int clq = L->use_clusters == 2 ? 2 : 4;
      rc |= hb_make_tmap(&Q.hq_hi, B.hs_hi[l], (uint64_t)(T + 1) * R_pad, hbl::HIDN, BM / clq);
      rc |= hb_make_tmap(&Q.hq_lo, B.hs_lo[l], (uint64_t)(T + 1) * R_pad, hbl::HIDN, BM / clq);
      Q.y = l == 1 ? y[n] : nullptr;
      Q.act = sv ? L->act[l] : nullptr; Q.cs = sv ? L->cs[l] : nullptr;
    }
  }
  if (rc) return -2;
  rc = hb_upload(&L->upload, L->d_fwd, fp.data(), 2 * sizeof(hbl::FwdParams), st);
  if (rc) return rc;
  auto launch_rec = [&](int l, cudaStream_t s) -> int {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ctas, 1, 1);
    cfg.blockDim = dim3(192, 1, 1);
    cfg.dynamicSmemBytes = hbl::FWD_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    const int clw = L->use_clusters == 2 ? 2 : 4;
    attr[0].val.clusterDim.x = (unsigned)clw; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;   // every CTA must be resident at once (step barrier): use clusters only if they all fit
    const bool cl4 = L->use_clusters > 1 &&
                     cudaOccupancyMaxActiveClusters(&max_clusters, clw == 2 ? hbl::lstm_fwd_kernel<2> : hbl::lstm_fwd_kernel<4>, &cfg) == cudaSuccess &&
                     max_clusters * clw >= (wave ? 2 : 1) * n_ctas;
    const hbl::FwdParams* dp = L->d_fwd + l;
    if (cl4) {
      if (clw == 2) HB_CUDA(cudaLaunchKernelEx(&cfg, hbl::lstm_fwd_kernel<2>, dp));
      else HB_CUDA(cudaLaunchKernelEx(&cfg, hbl::lstm_fwd_kernel<4>, dp));
    } else {
      (void)cudaGetLastError();
      hbl::lstm_fwd_kernel<1><<<n_ctas, 192, hbl::FWD_SMEM, s>>>(dp);
    }
    L->last_cluster = cl4 ? clw : 1;
    HB_CUDA(cudaGetLastError());
    L->launches += 1;
    return 0;
  };
  rc = gx_gemm(0, 0, T, st, 0, 0);
  if (rc) return rc;
  if (wave) {
    HB_CUDA(cudaMemsetAsync(L->chunk_flags, 0, 64 * sizeof(unsigned), st));
    HB_CUDA(cudaEventRecord(L->wev[0], st));
    for (int i = 0; i < 3; ++i) HB_CUDA(cudaStreamWaitEvent(L->ws[i], L->wev[0], 0));
    rc = launch_rec(0, L->ws[0]);
    if (rc) return rc;
    HB_CUDA(cudaEventRecord(L->wev[1], L->ws[0]));
    rc = launch_rec(1, L->ws[2]);
    if (rc) return rc;
    HB_CUDA(cudaEventRecord(L->wev[3], L->ws[2]));
    for (int c = 0; c < n_chunks; ++c) {
      const int t0 = c * chunk, t1 = t0 + chunk < T ? t0 + chunk : T;
      // h^0 of steps < t1 is published once layer 0's step counters show t1 completed steps (32 CTAs per row block)
      hbl::lstm_wait_steps<<<1, 1, 0, L->ws[1]>>>(fp[0].ctr, nets * MB, (unsigned)(hbl::SLICES * t1), L->d_error);
      rc = gx_gemm(1, t0, t1, L->ws[1], 16 + 2 * c, gemm_sms);
      if (rc) return rc;
      hbl::lstm_set_flag<<<1, 1, 0, L->ws[1]>>>(L->chunk_flags + c);
      L->launches += 2;
    }
    HB_CUDA(cudaEventRecord(L->wev[2], L->ws[1]));
    HB_CUDA(cudaGetLastError());
    for (int i = 1; i <= 3; ++i) HB_CUDA(cudaStreamWaitEvent(st, L->wev[i], 0));
  } else {
    rc = launch_rec(0, st);
    if (rc) return rc;
    rc = gx_gemm(1, 0, T, st, 2, 0);
    if (rc) return rc;
    rc = launch_rec(1, st);
    if (rc) return rc;
  }
  if (save) {  // transposed copies of net 0's h sequences (block 0 = zeros included): operands of the weight-gradient GEMMs
    for (int l = 0; l < 2; ++l)
      hbl::lstm_transpose_pair<<<dim3(hbl::HIDN / 64, (unsigned)((size_t)(T + 1) * R_pad / 64)), 256, 0, st>>>(L->nb[0].hs_hi[l], L->nb[0].hs_lo[l], hbl::HIDN,
                                                                                                          L->hsT_hi[l], L->hsT_lo[l], ldT);
    L->launches += 2;
  }
  L->saved = save ? 1 : 0;
  hbl_dump_trace(L, st, 0, T, "fwd");
  return hbl_finish_call(L, st);
}

int hb_lstm_backward(hb_lstm* L, const float* dy, float* dx, const hb_lstm_grads* g, void* stream) {
  if (!L || !dy || !g) { hb_set_error("hb_lstm_backward: null argument"); return -1; }
  if (!L->saved) { hb_set_error("hb_lstm_backward: no saved forward (call hb_lstm_forward with save != 0 first)"); return -1; }
  if (reinterpret_cast<uintptr_t>(dy) & 31) { hb_set_error("hb_lstm_backward: dy must be 32-byte aligned (the kernels read rows with 32-byte loads)"); return -1; }
  for (int l = 0; l < 2; ++l)
    if (!g->dw_ih[l] || !g->dw_hh[l] || !g->db_ih[l] || !g->db_hh[l]) { hb_set_error("hb_lstm_backward: null gradient pointer"); return -1; }
  HB_CUDA(cudaSetDevice(L->device));
  { const int prc = hbl_check_previous(L, false); if (prc) return prc; }
  cudaStream_t st = (cudaStream_t)stream;
  const int T = L->T, rows = L->rows, R_pad = L->R_pad, MB = L->MB;
  const size_t N = (size_t)T * R_pad;
  const long long ldT = (long long)(T + 1) * R_pad;
  HbLstmNetBuf& B = L->nb[0];
  HB_CUDA(cudaMemsetAsync(L->ctr, 0, 16 * hbl::CTR_STRIDE * sizeof(unsigned), st));
  int rc = 0;
  std::vector<hbl::BwdParams> bp(2);
  memset(bp.data(), 0, 2 * sizeof(hbl::BwdParams));
  // Layer wavefront: the recurrence of layer 0 needs dLoss/dh0_t = dG1_t W_ih1, i.e. layer 1's dgates of the SAME step -- so
  // the two recurrences can run side by side, layer 0 one time chunk behind: layer 1 (stream 0) publishes its step counter,
  // the dX GEMM of each finished chunk of steps runs on the SMs both recurrences leave free (stream 1) and raises a flag that
  // layer 0's kernel (stream 2) waits for before it enters that chunk.  2 * 80 dependent steps become 80 + one chunk.
  const int gemm_sms = (L->sm_count - 2 * MB * hbl::SLICES) & ~1;
  int chunk = 8;
  while ((T + chunk - 1) / chunk > 32) chunk *= 2;
  const int n_chunks = (T + chunk - 1) / chunk;
  const bool wave = L->use_wavefront && gemm_sms >= 8 && T >= 2 * chunk;
  const size_t part_layer = (size_t)2 * L->max_rpad * hbl::HIDN;
  for (int l = 1; l >= 0; --l) {
    hbl::BwdParams& Q = bp[l];
    rc |= hb_make_tmap(&Q.wt_hi, B.whhT_hi[l], hbl::HIDN, hbl::G4, 256);
    rc |= hb_make_tmap(&Q.wt_lo, B.whhT_lo[l], hbl::HIDN, hbl::G4, 256);
    if (rc) return -2;
    Q.dh_ext = l == 1 ? dy : L->dh0; Q.dh_rows = l == 1 ? rows : R_pad;
    Q.act = L->act[l]; Q.cs = L->cs[l];
    Q.dg_hi = L->dg_hi[l]; Q.dg_lo = L->dg_lo[l];
    Q.part = L->part + (size_t)l * part_layer; Q.T = T; Q.rows = rows; Q.R_pad = R_pad; Q.MB = MB;
    Q.ctr = L->ctr + (size_t)l * 8 * hbl::CTR_STRIDE; Q.error_flag = L->d_error;
    Q.chunk_flags = (wave && l == 0) ? L->chunk_flags : nullptr; Q.chunk = chunk;
    Q.trace = L->d_trace ? L->d_trace + (size_t)(2 + l) * L->max_T * 16 : nullptr;
    if (R_pad != rows) {  // padded rows of the dgate operands must read as zero in the GEMMs below
      const size_t bf = sizeof(__nv_bfloat16);
      HB_CUDA(cudaMemsetAsync(L->dg_hi[l], 0, N * hbl::G4 * bf, st)); HB_CUDA(cudaMemsetAsync(L->dg_lo[l], 0, N * hbl::G4 * bf, st));
    }
    HB_CUDA(cudaMemsetAsync(Q.part, 0, (size_t)2 * R_pad * hbl::HIDN * sizeof(float), st));   // the two dh accumulators
  }
  rc = hb_upload(&L->upload, L->d_bwd, bp.data(), 2 * sizeof(hbl::BwdParams), st);
  if (rc) return rc;
  // dX of layer l over the rows of steps [t0, t1): [rows, 2048] x [2048, 512]
  auto dx_gemm = [&](int l, int t0, int t1, float* dst, cudaStream_t s, int slot, int sm_limit) -> int {
    Params gp;
    int r2 = 0;
    const size_t r0 = (size_t)t0 * R_pad, nr = (size_t)(t1 - t0) * R_pad;
    hbl_gemm_problem(gp, r2, L->dg_hi[l] + r0 * hbl::G4, L->dg_lo[l] + r0 * hbl::G4, nr, hbl::G4, B.wihT_hi[l], B.wihT_lo[l], hbl::HIDN, hbl::G4, hbl::G4,
                     ((nr / BM) % 2 == 0) ? 2 : 1, nullptr, dst + r0 * hbl::HIDN, hbl::HIDN, L->d_error);
    if (r2) return -2;
    return hbl_run_gemm(L, s, &gp, 1, (int)(nr / BM), hbl::HIDN / BN, slot, sm_limit);
  };
  // Weight gradients inside the wavefront: with one row block (rows <= 128, the IQL learner) the two recurrences leave more than
  // half of the SMs idle for a millisecond.  A fourth stream follows both step counters one time chunk behind: transpose the
  // chunk's dgate rows, then ADD its contribution dG_chunk^T x (x | h)_chunk to the four weight gradients (K = chunk * R_pad per
  // launch, `accumulate` epilogue; fixed chunk order, so the sums are deterministic).  After the recurrences only the bias sums
  // and the row permutation are left.  At 256 rows the 20 free SMs could not keep up, so the GEMMs stay behind the wavefront.
  const size_t WN = (size_t)hbl::G4 * hbl::HIDN;
  const int dw_sms = (gemm_sms - 16) & ~1;
  const bool dw_wave = wave && dw_sms >= 48 && getenv("HB_LSTM_NO_DW_WAVE") == nullptr;
  auto dw_chunk = [&](int l, int t0, int t1, bool first, int slot) -> int {
    const size_t r0 = (size_t)t0 * R_pad, nr = (size_t)(t1 - t0) * R_pad;
    hbl::lstm_transpose_pair<<<dim3(hbl::G4 / 64, (unsigned)(nr / 64)), 256, 0, L->ws[3]>>>(L->dg_hi[l] + r0 * hbl::G4, L->dg_lo[l] + r0 * hbl::G4, hbl::G4,
                                                                                         L->dgT_hi[l] + r0, L->dgT_lo[l] + r0, (long long)N);
    Params wq[2];
    int r2 = 0;
    // layer 0: inputs x_t and h0_{t-1}; layer 1: inputs h0_t (= block t+1 of the h0 sequence) and h1_{t-1}
    const __nv_bfloat16 *in_hi = l == 0 ? L->xT_hi + r0 : L->hsT_hi[0] + R_pad + r0, *in_lo = l == 0 ? L->xT_lo + r0 : L->hsT_lo[0] + R_pad + r0;
    hbl_gemm_problem(wq[0], r2, L->dgT_hi[l] + r0, L->dgT_lo[l] + r0, hbl::G4, N, in_hi, in_lo, hbl::HIDN, ldT, nr, 2, nullptr, L->dwp + (2 * l) * WN, hbl::HIDN, L->d_error);
    hbl_gemm_problem(wq[1], r2, L->dgT_hi[l] + r0, L->dgT_lo[l] + r0, hbl::G4, N, L->hsT_hi[l] + r0, L->hsT_lo[l] + r0, hbl::HIDN, ldT, nr, 2, nullptr,
                     L->dwp + (2 * l + 1) * WN, hbl::HIDN, L->d_error);
    if (r2) return -2;
    wq[0].accumulate = wq[1].accumulate = first ? 0 : 1;
    L->launches += 1;
    return hbl_run_gemm(L, L->ws[3], wq, 2, hbl::G4 / BM, hbl::HIDN / BN, slot, dw_sms);
  };
  if (wave) {
    HB_CUDA(cudaMemsetAsync(L->chunk_flags, 0, 64 * sizeof(unsigned), st));
    HB_CUDA(cudaEventRecord(L->wev[0], st));
    for (int i = 0; i < 4; ++i) HB_CUDA(cudaStreamWaitEvent(L->ws[i], L->wev[0], 0));
    hbl::lstm_bwd_kernel<<<MB * hbl::SLICES, 192, hbl::BWD_SMEM, L->ws[0]>>>(L->d_bwd + 1);
    HB_CUDA(cudaEventRecord(L->wev[1], L->ws[0]));
    hbl::lstm_bwd_kernel<<<MB * hbl::SLICES, 192, hbl::BWD_SMEM, L->ws[2]>>>(L->d_bwd + 0);
    HB_CUDA(cudaEventRecord(L->wev[3], L->ws[2]));
    for (int c = 0; c < n_chunks; ++c) {
      const int t1 = T - c * chunk, t0 = t1 - chunk > 0 ? t1 - chunk : 0;
      if (t0 > 0) {   // steps t1-1 .. t0 of layer 1 are published when its counter shows (c+1)*chunk completed steps
        hbl::lstm_wait_steps<<<1, 1, 0, L->ws[1]>>>(bp[1].ctr, MB, (unsigned)(hbl::SLICES * (c + 1) * chunk), L->d_error);
        L->launches += 1;
      } else {        // step 0 does not bump the counter: the last chunk waits for the kernel itself
        HB_CUDA(cudaStreamWaitEvent(L->ws[1], L->wev[1], 0));
      }
      rc = dx_gemm(1, t0, t1, L->dh0, L->ws[1], 16 + c, gemm_sms);
      if (rc) return rc;
      hbl::lstm_set_flag<<<1, 1, 0, L->ws[1]>>>(L->chunk_flags + c);
      L->launches += 1;
    }
    HB_CUDA(cudaEventRecord(L->wev[2], L->ws[1]));
    if (dw_wave) {
      for (int c = 0; c < n_chunks; ++c) {
        const int t1 = T - c * chunk, t0 = t1 - chunk > 0 ? t1 - chunk : 0;
        for (int l = 1; l >= 0; --l) {   // layer 0 runs one chunk behind layer 1: the stream simply waits for it
          if (t0 > 0) {
            hbl::lstm_wait_steps<<<1, 1, 0, L->ws[3]>>>(bp[l].ctr, MB, (unsigned)(hbl::SLICES * (c + 1) * chunk), L->d_error);
            L->launches += 1;
          } else {
            HB_CUDA(cudaStreamWaitEvent(L->ws[3], L->wev[l == 1 ? 1 : 3], 0));
          }
          rc = dw_chunk(l, t0, t1, c == 0, 48 + 2 * ((2 * c + l) % 8));
          if (rc) return rc;
        }
      }
      // the row permutation and the bias sums stay on that stream too: they run under the dX GEMM below; joined at the end
      float* dsts[4] = {g->dw_ih[0], g->dw_hh[0], g->dw_ih[1], g->dw_hh[1]};
      for (int i = 0; i < 4; ++i) hbl::lstm_unperm_rows<<<hbl::G4, 128, 0, L->ws[3]>>>(L->dwp + i * WN, dsts[i]);
      for (int l = 0; l < 2; ++l) hbl::lstm_bias_grad<<<hbl::G4, 256, 0, L->ws[3]>>>(L->dgT_hi[l], L->dgT_lo[l], (long long)N, g->db_ih[l], g->db_hh[l]);
      L->launches += 6;
      HB_CUDA(cudaEventRecord(L->wev[4], L->ws[3]));
    }
    HB_CUDA(cudaGetLastError());
    for (int i = 1; i <= 3; ++i) HB_CUDA(cudaStreamWaitEvent(st, L->wev[i], 0));
    L->launches += 2;
  } else {
    hbl::lstm_bwd_kernel<<<MB * hbl::SLICES, 192, hbl::BWD_SMEM, st>>>(L->d_bwd + 1);
    rc = dx_gemm(1, 0, T, L->dh0, st, 5, 0);
    if (rc) return rc;
    hbl::lstm_bwd_kernel<<<MB * hbl::SLICES, 192, hbl::BWD_SMEM, st>>>(L->d_bwd + 0);
    HB_CUDA(cudaGetLastError());
    L->launches += 2;
  }
  if (!dw_wave) {
    for (int l = 0; l < 2; ++l)
      hbl::lstm_transpose_pair<<<dim3(hbl::G4 / 64, (unsigned)(N / 64)), 256, 0, st>>>(L->dg_hi[l], L->dg_lo[l], hbl::G4, L->dgT_hi[l], L->dgT_lo[l], (long long)N);
    // the bias sums need nothing but the transposed dgates: on the side stream, under the two GEMMs below (they stream 2 x 84 MB
    // at 256 rows while the GEMMs keep the tensor cores busy); joined at the end
    HB_CUDA(cudaEventRecord(L->wev[5], st));
    HB_CUDA(cudaStreamWaitEvent(L->ws[3], L->wev[5], 0));
    for (int l = 0; l < 2; ++l) hbl::lstm_bias_grad<<<hbl::G4, 256, 0, L->ws[3]>>>(L->dgT_hi[l], L->dgT_lo[l], (long long)N, g->db_ih[l], g->db_hh[l]);
    HB_CUDA(cudaEventRecord(L->wev[4], L->ws[3]));
    L->launches += 4;
  }
  if (dx) {   // gradient w.r.t. the input sequence of layer 0
    float* dst = R_pad == rows ? dx : L->dx_pad;
    // (not chunked into the wavefront: 16 tiles of K = 2048 per chunk on 16 SMs take as long as the chunk itself)
    rc = dx_gemm(0, 0, T, dst, st, 4, 0);
    if (rc) return rc;
    if (dst != dx) { hbl::lstm_unpad<<<dim3(rows, T), 128, 0, st>>>(L->dx_pad, dx, rows, R_pad); L->launches += 1; }
  }
  // ---- weight gradients: four [2048, 512] = dG^T [2048, N] x (operand^T [512, N])^T problems in one launch
  Params wp[4];
  hbl_gemm_problem(wp[0], rc, L->dgT_hi[0], L->dgT_lo[0], hbl::G4, N, L->xT_hi, L->xT_lo, hbl::HIDN, ldT, N, 2, nullptr, L->dwp + 0 * WN, hbl::HIDN, L->d_error);
  hbl_gemm_problem(wp[1], rc, L->dgT_hi[0], L->dgT_lo[0], hbl::G4, N, L->hsT_hi[0], L->hsT_lo[0], hbl::HIDN, ldT, N, 2, nullptr, L->dwp + 1 * WN, hbl::HIDN, L->d_error);
  hbl_gemm_problem(wp[2], rc, L->dgT_hi[1], L->dgT_lo[1], hbl::G4, N, L->hsT_hi[0] + R_pad, L->hsT_lo[0] + R_pad, hbl::HIDN, ldT, N, 2, nullptr, L->dwp + 2 * WN, hbl::HIDN, L->d_error);
  hbl_gemm_problem(wp[3], rc, L->dgT_hi[1], L->dgT_lo[1], hbl::G4, N, L->hsT_hi[1], L->hsT_lo[1], hbl::HIDN, ldT, N, 2, nullptr, L->dwp + 3 * WN, hbl::HIDN, L->d_error);
  if (rc) return -2;
  if (!dw_wave) {
    rc = hbl_run_gemm(L, st, wp, 4, hbl::G4 / BM, hbl::HIDN / BN, 8);
    if (rc) return rc;
    float* dsts[4] = {g->dw_ih[0], g->dw_hh[0], g->dw_ih[1], g->dw_hh[1]};
    for (int i = 0; i < 4; ++i) hbl::lstm_unperm_rows<<<hbl::G4, 128, 0, st>>>(L->dwp + i * WN, dsts[i]);
    L->launches += 4;
  }
  HB_CUDA(cudaStreamWaitEvent(st, L->wev[4], 0));   // the side stream: bias sums (and, in the wavefront, the weight gradients)
  HB_CUDA(cudaGetLastError());
  hbl_dump_trace(L, st, 2, T, "bwd");
  return hbl_finish_call(L, st);
}

int hb_lstm_sync(hb_lstm* L) {
  if (!L) { hb_set_error("hb_lstm_sync: null handle"); return -1; }
  HB_CUDA(cudaSetDevice(L->device));
  return hbl_check_previous(L, true);
}

int64_t hb_lstm_launches(const hb_lstm* L) { return L ? L->launches : 0; }

}  // extern "C"
