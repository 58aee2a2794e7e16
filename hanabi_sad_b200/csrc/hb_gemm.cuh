// hb_gemm.cuh -- the dense contraction of the R2D2 act forward (pyhanabi/r2d2.py:65-78: Linear+ReLU, 2-layer
// LSTM cell) as ONE sm_100a kernel template: TMA-staged 128B-swizzled shared-memory tiles, tcgen05.mma with the
// accumulator in TMEM, and the layer's pointwise tail (bias+ReLU, or the LSTM gate non-linearities and state
// update) fused into the TMEM->register epilogue.
//
// Precision: the reference network is fp32 (contract: 1e-4 against CPU fp32 nn.LSTM).  Tensor
// cores are fed a 2-term bf16 split of both operands, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and the
// product is accumulated in fp32 as  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is < 2^-16 relative), i.e.
// three tcgen05.mma per K-slice ("bf16x3").  Activations are produced already split by the previous layer's
// epilogue; weights are split once when they are uploaded.
//
// Tile: BM=128 rows (agents) x BN=256 output columns, K streamed in 64-wide chunks.  For an LSTM layer the 256
// columns of a tile are [gate i|f|g|o][64 hidden units] (weights are stored gate-interleaved per tile), so one
// epilogue thread (= one TMEM lane = one agent row) holds all four gates of a hidden unit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hbg {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle span
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int A_TILE = BM * BK * 2;                  // 16 KB
constexpr int B_TILE = BN * BK * 2;                  // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE; // hi+lo of both operands: 96 KB
constexpr int HEAD_MAX_OUT = 56;                     // advantage head rows + the value head, padded to a multiple of 8
constexpr int HEAD_SMEM = HEAD_MAX_OUT * 64 * 4;     // one n-tile's slice of the head weights: [64 units][out padded to 8] fp32
constexpr int BIAS_SMEM = 4 * BN * 4;                // EPI_LSTM: the tile's 256 gate biases, one private copy per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/ + HEAD_SMEM + BIAS_SMEM;
constexpr int THREADS = 192;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int TMEM_COLS = 256;
constexpr int HID = 512;

enum { EPI_F32 = 0, EPI_RELU = 1, EPI_LSTM = 2 };

struct __align__(64) Params {
  CUtensorMap a_hi[2], a_lo[2];  // A operand [rows][K] bf16, K contiguous; segment 0 then segment 1 along K
  CUtensorMap b_hi, b_lo;        // B operand [N][K] bf16 (= nn.Linear / nn.LSTM weight layout), K contiguous
  int k_chunks;                  // total K / 64
  int k_chunks_seg0;             // chunks taken from a_*[0]; the rest come from a_*[1] (its own column 0 onward)
  int lo_first, lo_last;         // K-chunk range [first,last) in which the A operand has a non-zero lo part
  int split;                     // 1: bf16x3 (hi*lo + lo*hi + hi*hi); 0: plain bf16 (hi*hi only, lo operands never loaded)
  const float* bias;             // [N]
  // EPI_F32: plain fp32 result (diagnostics / self-test)
  float* c_f32;
  int ldc;
  int accumulate;                // EPI_F32: C += result (weight gradients summed over time chunks, hb_lstm_backward)
  // EPI_RELU / EPI_LSTM: result written as a bf16 hi/lo pair, [rows][out_ld], starting at column out_col0
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int out_ld;
  int out_col0;
  // EPI_LSTM: cell state in / out [rows][512] fp32 (c_out may be null: the target-network pass keeps no state),
  // optional fp32 copy of h' [rows][512] (feeds the advantage head)
  const float* c_in;
  float* c_out;
  float* h_f32;
  // EPI_LSTM, top layer only (null otherwise): the dueling head fused into the epilogue.  head_w = fc_a / fc_v weights
  // re-tiled as [n_tile][64 units][hop] fp32 (head_out = A + 1, last = fc_v; hop = head_out rounded up to 8, zero padded);
  // every epilogue thread multiplies
  // its row's 64 fresh h' values with the tile's slice and writes head_part[n_tile][row][hop] -- the act kernel
  // adds the 8 partial sums.  Keeps h' out of HBM and removes a 512-deep reduction kernel.
  const float* head_w;
  float* head_part;
  int head_out;
  int head_rows;                 // rows_pad (stride of the n_tile axis of head_part)
  // Row mapping of this problem: tile row r (0 .. tiles*128) is row  r*row_mul + row_add  of the caller's buffers and is
  // real only if r < valid_rows.  (1, 0, rows_pad) for a problem over all agents; (P, seat, G) when the problem is ONE
  // seat's network over that seat's agents (evaluation with a different network per seat: the A operands are strided
  // TMA views, everything the epilogue touches goes through this mapping).
  int row_mul, row_add, valid_rows;
  // EPI_LSTM top layer with skip_connect (r2d2.py:74-75): the head consumes h' + x; x as bf16 hi/lo [rows][512]
  const __nv_bfloat16* skip_hi;
  const __nv_bfloat16* skip_lo;
  int* error_flag;               // set to 1 if a barrier wait ran into the spin guard (never in a healthy run)
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (the box is shared); it raises error_flag instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, bool& dead) {
  if (dead) return;
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 21)) {
      if (error_flag) atomicExch(error_flag, 1);
      dead = true;  // stop waiting on anything else in this thread: finish fast, report through error_flag
      return;
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Multicast variant: the box lands at the same shared-memory offset of every CTA in `mask` and completes bytes on the
// mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// cta_group::2 variant: issued by either CTA of a pair, the bytes complete on the mbarrier of the pair's LEADER (the
// address has the CTA-rank bit cleared by the caller).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Arrives on the mbarrier at the same offset in every CTA of `mask` once the MMAs issued so far have completed.
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive TMEM columns of this warp's 32 lanes WITHOUT waiting: lets the next load fly while the previous one is
// being consumed (tmem_wait_ld waits for everything outstanding).
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand tile of [rows][64] bf16, 128-byte swizzle (what TMA's
// CU_TENSOR_MAP_SWIZZLE_128B writes): 8-row groups are 1024 bytes apart (SBO), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset: unused for swizzled K-major layouts
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // version
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: D fp32, A and B bf16, both K-major, M=128, N=BN.
__device__ __forceinline__ constexpr uint32_t make_idesc(int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// LSTM activations on the two MUFU approximations (ex2, rcp; both within 2 ulp).  Written with the bare instructions:
// __expf / __fdividef wrap them in range extensions for denormal results and for |divisor| > 2^126 (a compare and two
// predicated multiplies each, 8.4 instructions per activation against 4-5 here), neither of which can matter -- 1 + e^x
// is never below 1, and a flushed e^x below 2^-126 changes nothing next to that 1.  In the normal range the values are
// bit-identical to the intrinsics' (same ex2 argument rounding: 2x is exact; 1 - 2r rounds once either way).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return fmaf(-2.f, rcp_ftz(1.f + ex2_ftz(2.8853900817779268f * x)), 1.f); }

// acc[j] += h[u] * w[u][j] for 2 * NP outputs j, u = 0..63: the fused dueling head of the top LSTM layer's epilogue.
// w: shared memory, [64][ld] floats (16-byte aligned rows); out: this row's slice of head_part, 16-byte aligned (rows of
// head_part are padded to `hop` floats, so a thread writes whole 32-byte sectors with 16-byte stores -- with the unpadded
// [A + 1] rows every scalar store of a warp touched 32 different sectors, which cost 18 us per launch at 5 players).
template <int NP>
__device__ __forceinline__ void head_dot(const float* h, const float* w, int ld, float* out, bool valid) {
  unsigned long long acc[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = 0ull;
#pragma unroll
  for (int u = 0; u < 64; ++u) {   // fully unrolled: h[] must stay in registers
    unsigned long long hh;
    asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h[u]));
    const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(w + u * ld);
#pragma unroll
    for (int j = 0; j < NP / 2; ++j) {
      const ulonglong2 ww = wr[j];
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[2 * j]) : "l"(hh), "l"(ww.x));
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[2 * j + 1]) : "l"(hh), "l"(ww.y));
    }
  }
  if (!valid) return;
#pragma unroll
  for (int j = 0; j < NP / 2; ++j) reinterpret_cast<ulonglong2*>(out)[j] = make_ulonglong2(acc[2 * j], acc[2 * j + 1]);
}

// 256-bit global accesses (sm_100: LDG / STG.E.ENL2.256).  In the epilogues a lane owns a ROW: its 16-byte accesses land in
// 32 different sectors per warp request, half a sector each, and the other half comes with the next request -- 32-byte
// accesses move whole sectors and halve the LSU's sector transactions (the epilogue's real cost, see head_dot).
// p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_cg(const float* p, float* v) {   // L2 only: data another kernel writes concurrently
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_nc(const float* p, float* v) {   // read-only for the kernel's lifetime
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// 16 values -> bf16 hi / lo halves, 32 bytes each (both pointers 32-byte aligned)
__device__ __forceinline__ void store_split16(const float* v, __nv_bfloat16* hi_ptr, __nv_bfloat16* lo_ptr) {
  // two values per conversion (cvt.rn.bf16x2.f32): same round-to-nearest-even results as split_bf16, 40 % fewer instructions
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 back = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  stg256(hi_ptr, h);
  stg256(lo_ptr, l);
}

// ---------------------------------------------------------------------------------------------- the kernel
// Persistent: grid = min(#tiles, #SMs), every CTA walks the tile list  tile = blockIdx.x, += gridDim.x  with
// tile -> (problem z, m_tile, n_tile), n fastest, so that the CTAs running at the same time share A row-blocks and
// the whole B matrix of one network in L2.  `ps[z]` picks the problem (online / target network): both networks' layer
// run as one launch.  Params records live in global memory (the tensor maps inside them are read by the TMA unit
// through their generic address).
//
// Three concurrent pipelines per CTA:
//   warp 0 (1 thread)  TMA producer: streams K-chunks of A/B into a STAGES-deep shared-memory ring, across tile borders;
//   warp 1 (1 thread)  tcgen05.mma issuer: accumulates a tile into one of TWO 256-column TMEM accumulators;
//   warps 2-5          epilogue: drain the other accumulator (tcgen05.ld), apply the layer's pointwise tail, store.
// so the epilogue of tile i overlaps the main loop of tile i+1.
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int ACC_STAGES = 2;
constexpr int TMEM_ALLOC_COLS = ACC_STAGES * TMEM_COLS;  // 512: the whole tensor memory of the SM

//
// CL = 2: the grid is launched as clusters of two CTAs that work on vertically adjacent tiles (same n_tile, m_tile =
// 2*pair + rank) and therefore need the SAME weight tile: each CTA fetches one half of it (128 of the 256 rows) and
// TMA-multicasts it into both shared memories, halving the L2 -> SMEM traffic of the B operand (which dominates: the
// tile is 128 x 256).  A slot is refilled only after BOTH CTAs' MMAs have released it (empty barriers count 2, the
// release is a multicast tcgen05.commit).  Accumulators, MMAs and epilogues stay per-CTA (cta_group::1).
//
// CL = 3: the same pairs as ONE tcgen05 CTA pair (cta_group::2): the leader CTA issues M = 256 MMAs for both, each CTA
// holds its own 128 rows of A and HALF of the weight tile (the tensor core reads the other half from the peer), its own
// 128 x 256 accumulator in its own TMEM and runs its own epilogue.  Per MMA a CTA's shared memory supplies 8 KB instead
// of 12 KB and receives 64 KB instead of 96 KB per K-chunk (3-deep ring), which lifts the shared-memory bandwidth bound
// of the single-CTA form (TMA fill 62 B/clk + operand reads 96 B/clk > 128 B/clk).
template <int EPI, int CL>
__global__ void __launch_bounds__(THREADS, 1) gemm3_kernel(const Params* __restrict__ ps, int n_tiles_n, int n_tiles_m, int n_problems) {
  constexpr bool PAIR = CL == 3;
  constexpr int CW = CL == 1 ? 1 : 2;                       // CTAs per cluster
  constexpr int NST = PAIR ? 3 : STAGES;                    // ring depth
  constexpr int STB = PAIR ? 2 * A_TILE + B_TILE : STAGE_BYTES;  // bytes per ring slot (PAIR: half weight tiles)
  constexpr int OFF_BLO = PAIR ? 2 * A_TILE + B_TILE / 2 : 2 * A_TILE + B_TILE;
  static_assert(NST * STB <= STAGES * STAGE_BYTES, "ring must fit the shared-memory budget");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers (8 bytes each): full[NST], empty[NST], tmem_full[ACC_STAGES], tmem_empty[ACC_STAGES], then the TMEM base slot
  const uint32_t bar_full = bar_base, bar_empty = bar_full + 8 * NST, bar_tfull = bar_empty + 8 * NST;
  const uint32_t bar_tempty = bar_tfull + 8 * ACC_STAGES, tmem_slot = bar_tempty + 8 * ACC_STAGES;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* head_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 128 - smem_u32(smem_raw)));  // [64][hop] (EPI_RELU: the tile's biases)
  float* bias_smem = head_smem + HEAD_SMEM / 4;   // [4 epilogue warps][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool dead = false;
  // work items: single tiles (CL = 1) or vertical tile pairs (CL = 2), walked with stride = number of CTAs / clusters
  const int rank = CW == 2 ? (int)cluster_ctarank() : 0;
  const int tiles_per_problem = n_tiles_n * (n_tiles_m / CW);
  const int total_tiles = tiles_per_problem * n_problems;
  const int first_tile = blockIdx.x / CW, tile_stride = gridDim.x / CW;

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < n_problems; ++z) {
      tma_prefetch_desc(&ps[z].a_hi[0]); tma_prefetch_desc(&ps[z].a_lo[0]); tma_prefetch_desc(&ps[z].a_hi[1]); tma_prefetch_desc(&ps[z].a_lo[1]);
      tma_prefetch_desc(&ps[z].b_hi); tma_prefetch_desc(&ps[z].b_lo);
    }
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CL == 2 ? 2 : 1); }
    for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_ALLOC_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_ALLOC_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CW == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int z = tile / tiles_per_problem, r = tile - z * tiles_per_problem;
        const int m_tile = (r / n_tiles_n) * CW + rank, n_tile = r % n_tiles_n;
        const Params& p = ps[z];
        const int k_chunks = p.k_chunks, lo_first = p.lo_first, lo_last = p.lo_last, seg0 = p.k_chunks_seg0;
        const bool split = p.split != 0;
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const uint32_t s = it % NST, ph = (it / NST) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.error_flag, dead);  // slot free (first pass: passes immediately)
          const uint32_t st = smem_base + s * STB;
          const bool need_lo = split && kc >= lo_first && kc < lo_last;
          const int seg = kc >= seg0 ? 1 : 0;
          const int kx = (seg ? kc - seg0 : kc) * BK;
          if (PAIR) {
            // the leader's barrier collects the bytes of BOTH CTAs; only the leader arms it
            const uint32_t fb = (bar_full + 8 * s) & PEER_MASK;
            const uint32_t bytes = (need_lo ? 2 : 1) * A_TILE + (split ? 2 : 1) * (B_TILE / 2);
            if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * bytes);
            tma_load_2d_pair(st, &p.a_hi[seg], fb, kx, m_tile * BM);
            if (need_lo) tma_load_2d_pair(st + A_TILE, &p.a_lo[seg], fb, kx, m_tile * BM);
            tma_load_2d_pair(st + 2 * A_TILE, &p.b_hi, fb, kc * BK, n_tile * BN + rank * (BN / 2));
            if (split) tma_load_2d_pair(st + OFF_BLO, &p.b_lo, fb, kc * BK, n_tile * BN + rank * (BN / 2));
            continue;
          }
          mbar_expect_tx(bar_full + 8 * s, (need_lo ? 2 : 1) * A_TILE + (split ? 2 : 1) * B_TILE);
          tma_load_2d(st, &p.a_hi[seg], bar_full + 8 * s, kx, m_tile * BM);
          if (need_lo) tma_load_2d(st + A_TILE, &p.a_lo[seg], bar_full + 8 * s, kx, m_tile * BM);
          if (CL == 1) {
            tma_load_2d(st + 2 * A_TILE, &p.b_hi, bar_full + 8 * s, kc * BK, n_tile * BN);
            if (split) tma_load_2d(st + 2 * A_TILE + B_TILE, &p.b_lo, bar_full + 8 * s, kc * BK, n_tile * BN);
          } else {  // this CTA's half of the weight tile, delivered to both CTAs of the pair
            const uint32_t half = (uint32_t)rank * (B_TILE / 2);
            tma_load_2d_mc(st + 2 * A_TILE + half, &p.b_hi, bar_full + 8 * s, kc * BK, n_tile * BN + rank * (BN / 2), 3);
            if (split) tma_load_2d_mc(st + 2 * A_TILE + B_TILE + half, &p.b_lo, bar_full + 8 * s, kc * BK, n_tile * BN + rank * (BN / 2), 3);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && (!PAIR || rank == 0)) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM);
      uint32_t it = 0, lt = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++lt) {
        const int z = tile / tiles_per_problem;
        const Params& p = ps[z];
        const int k_chunks = p.k_chunks, lo_first = p.lo_first, lo_last = p.lo_last;
        const bool split = p.split != 0;
        const uint32_t acc_stage = lt % ACC_STAGES, aph = (lt / ACC_STAGES) & 1u;
        mbar_wait(bar_tempty + 8 * acc_stage, aph ^ 1u, p.error_flag, dead);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * TMEM_COLS;
        uint32_t acc = 0;
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const uint32_t s = it % NST, ph = (it / NST) & 1u;
          mbar_wait(bar_full + 8 * s, ph, p.error_flag, dead);
          tc_fence_after();
          const uint32_t st = smem_base + s * STB;
          const bool need_lo = split && kc >= lo_first && kc < lo_last;
          const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + A_TILE);
          const uint64_t b_hi = make_desc_sw128(st + 2 * A_TILE), b_lo = make_desc_sw128(st + OFF_BLO);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);  // advance the start address inside the swizzle span
            if (PAIR) {
              if (split) { umma_bf16_pair(d_tmem, a_hi + adv, b_lo + adv, idesc, acc); acc = 1; }
              if (need_lo) umma_bf16_pair(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
              umma_bf16_pair(d_tmem, a_hi + adv, b_hi + adv, idesc, acc);
            } else {
              if (split) { umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, acc); acc = 1; }  // small terms first
              if (need_lo) umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
              umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, acc);
            }
            acc = 1;
          }
          if (CL == 1) umma_commit(bar_empty + 8 * s);  // the smem slot is free once these MMAs have read it
          else if (CL == 2) umma_commit_mc(bar_empty + 8 * s, 3);  // ... in BOTH CTAs: the peer's multicast writes into this slot too
          else umma_commit_pair(bar_empty + 8 * s, 3);
        }
        if (PAIR) umma_commit_pair(bar_tfull + 8 * acc_stage, 3);  // both CTAs' accumulators are complete
        else umma_commit(bar_tfull + 8 * acc_stage);               // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> pointwise tail -> global =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    uint32_t lt = 0;
    int staged_head = -1;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++lt) {
      const int z = tile / tiles_per_problem, r = tile - z * tiles_per_problem;
      const int m_tile = (r / n_tiles_n) * CW + rank, n_tile = r % n_tiles_n;
      const Params& p = ps[z];
      const int row_local = m_tile * BM + row_in_tile;
      const bool valid = row_local < p.valid_rows;
      const size_t row = (size_t)row_local * p.row_mul + p.row_add;
      const uint32_t acc_stage = lt % ACC_STAGES, aph = (lt / ACC_STAGES) & 1u;
      if (EPI == EPI_LSTM && p.head_w != nullptr && staged_head != z * 64 + n_tile) {
        // stage this (network, n-tile)'s slice of the head weights (the 4 epilogue warps only: named barrier 1) BEFORE waiting
        // for the accumulator, so that the copy runs under the tile's MMAs
        const int hop_s = (p.head_out + 7) & ~7;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with the previous tile's slice
        const float4* src = reinterpret_cast<const float4*>(p.head_w + (size_t)n_tile * hop_s * 64);
        float4* dst = reinterpret_cast<float4*>(head_smem);
        const int et = threadIdx.x - 64;
        for (int i = et; i < hop_s * 16; i += 128) dst[i] = __ldg(src + i);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        staged_head = z * 64 + n_tile;
      }
      if (EPI == EPI_LSTM || EPI == EPI_RELU || (EPI == EPI_F32 && p.bias != nullptr)) {
        // the tile's 256 (gate) biases into this warp's private shared-memory copy (two coalesced 16-byte loads per lane
        // instead of 256 uniform global loads per lane and tile), also ahead of the accumulator wait
        float* bw = bias_smem + q * BN;
        __syncwarp();
        const float4* bsrc = reinterpret_cast<const float4*>(p.bias + n_tile * BN);
        reinterpret_cast<float4*>(bw)[lane] = __ldg(bsrc + lane);
        reinterpret_cast<float4*>(bw)[lane + 32] = __ldg(bsrc + lane + 32);
        __syncwarp();
      }
      mbar_wait(bar_tfull + 8 * acc_stage, aph, p.error_flag, dead);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_stage * TMEM_COLS;
      if (EPI == EPI_F32) {
        // a lane owns a row of C: 32-byte stores (whole sectors) whenever the rows are 32-byte aligned
        const bool wide = ((reinterpret_cast<uintptr_t>(p.c_f32) | ((uintptr_t)p.ldc * 4u)) & 31u) == 0;
        const float* bw = bias_smem + q * BN;
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          float* dst = p.c_f32 + row * p.ldc + (size_t)n_tile * BN + c0;
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += bw[c0 + i];
          }
          if (valid && p.accumulate) {
            float old[16];
            if (wide) { ldg256(dst, old); ldg256(dst + 8, old + 8); }
            else {
#pragma unroll
              for (int i = 0; i < 16; ++i) old[i] = dst[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += old[i];
          }
          if (valid) {
            if (wide) { stg256(dst, v); stg256(dst + 8, v + 8); }
            else {
#pragma unroll
              for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
        }
      } else if (EPI == EPI_RELU) {
        // The mainloop of this layer is short (K = 896), so the epilogue must not be the longer of the two: the tile's 256
        // biases come from shared memory (instead of 256 global loads per thread), and the accumulator is read 32
        // columns at a time with the next tcgen05.ld in flight while the previous 32 values are rectified, split and stored.
        const float* bw = bias_smem + q * BN;   // staged per warp ahead of the accumulator wait (above)
        uint32_t buf[2][32];
        tmem_ld32_nowait(taddr, buf[0]);
#pragma unroll
        for (int it = 0; it < BN / 32; ++it) {
          const int c0 = it * 32;
          tmem_wait_ld();
          if (it + 1 < BN / 32) tmem_ld32_nowait(taddr + c0 + 32, buf[(it + 1) & 1]);
          const uint32_t* cur = buf[it & 1];
          const size_t o = row * p.out_ld + p.out_col0 + (size_t)n_tile * BN + c0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[16];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 b = *reinterpret_cast<const float4*>(bw + c0 + 16 * h + 4 * i4);
              v[4 * i4] = fmaxf(__uint_as_float(cur[16 * h + 4 * i4]) + b.x, 0.f);
              v[4 * i4 + 1] = fmaxf(__uint_as_float(cur[16 * h + 4 * i4 + 1]) + b.y, 0.f);
              v[4 * i4 + 2] = fmaxf(__uint_as_float(cur[16 * h + 4 * i4 + 2]) + b.z, 0.f);
              v[4 * i4 + 3] = fmaxf(__uint_as_float(cur[16 * h + 4 * i4 + 3]) + b.w, 0.f);
            }
            if (valid) store_split16(v, p.out_hi + o + 16 * h, p.out_lo + o + 16 * h);
          }
        }
      } else {
        // tile columns: [gate i | f | g | o][64 hidden units]; hidden unit = n_tile*64 + u
        const bool do_head = p.head_w != nullptr;
        const int hop = (p.head_out + 7) & ~7;
        float h_all[64];
#pragma unroll
        for (int u0 = 0; u0 < 64; u0 += 16) {
          float gi[16], gf[16], gg[16], go[16], c[16];
          float* h = h_all + u0;
          tmem_ld16(taddr + 0 * 64 + u0, gi);
          tmem_ld16(taddr + 1 * 64 + u0, gf);
          tmem_ld16(taddr + 2 * 64 + u0, gg);
          tmem_ld16(taddr + 3 * 64 + u0, go);
          const int unit = n_tile * 64 + u0;
          const float* bias = bias_smem + q * BN + u0;
          if (valid) { ldg256(p.c_in + row * HID + unit, c); ldg256(p.c_in + row * HID + unit + 8, c + 8); }
          else {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float ig = sigmoid_f(gi[i] + bias[i]);
            const float fg = sigmoid_f(gf[i] + bias[64 + i]);
            const float g_ = tanh_f(gg[i] + bias[128 + i]);
            const float og = sigmoid_f(go[i] + bias[192 + i]);
            c[i] = fg * c[i] + ig * g_;
            h[i] = og * tanh_f(c[i]);
          }
          if (p.c_out && valid) { stg256(p.c_out + row * HID + unit, c); stg256(p.c_out + row * HID + unit + 8, c + 8); }
          if (p.h_f32 && valid) {
            float4* ho = reinterpret_cast<float4*>(p.h_f32 + row * HID + unit);
#pragma unroll
            for (int i = 0; i < 4; ++i) ho[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
          }
          if (p.out_hi && valid) {
            const size_t o = row * p.out_ld + p.out_col0 + unit;
            store_split16(h, p.out_hi + o, p.out_lo + o);
          }
        }
        if (do_head) {
          // the TMEM reads of this tile are done: let the MMA thread start refilling the accumulator right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_cluster((bar_tempty + 8 * acc_stage) & PEER_MASK); else mbar_arrive(bar_tempty + 8 * acc_stage); }
          if (p.skip_hi != nullptr && valid) {  // skip_connect: the head sees h' + x
            const __nv_bfloat16* xh = p.skip_hi + row * HID + n_tile * 64;
            const __nv_bfloat16* xl = p.skip_lo + row * HID + n_tile * 64;
#pragma unroll
            for (int u = 0; u < 64; u += 8) {
              const uint4 a = *reinterpret_cast<const uint4*>(xh + u), b = *reinterpret_cast<const uint4*>(xl + u);
              const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(&a);
              const __nv_bfloat16* bl = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
              for (int i = 0; i < 8; ++i) h_all[u + i] += __bfloat162float(ah[i]) + __bfloat162float(bl[i]);
            }
          }
          // 16 (then 8) outputs at a time on packed fp32 FMAs (FFMA2: two accumulators per instruction, each one the same
          // rn FMA chain over u = 0..63 as a scalar loop would run); the weights of one unit are contiguous in shared
          // memory, so one broadcast LDS.128 feeds two FFMA2
          float* part = p.head_part + ((size_t)n_tile * p.head_rows + row) * hop;
          int o0 = 0;
          for (; o0 + 16 <= hop; o0 += 16) head_dot<8>(h_all, head_smem + o0, hop, part + o0, valid);
          if (o0 < hop) head_dot<4>(h_all, head_smem + o0, hop, part + o0, valid);
          continue;  // the accumulator was already released above
        }
      }
      // all of this warp's tcgen05.ld have completed (tmem_ld16 waits): hand the accumulator back to the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster((bar_tempty + 8 * acc_stage) & PEER_MASK); else mbar_arrive(bar_tempty + 8 * acc_stage); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CW == 2) cluster_sync_all();  // the peer may still be arriving on this CTA's barriers / reading its shared memory
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC_COLS) : "memory");
  }
}

}  // namespace hbg
