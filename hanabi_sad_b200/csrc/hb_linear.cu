// hb_linear.cu -- learner side: the dense layers around the LSTM.  The reference learner runs
// nn.Linear(838 -> 512) + ReLU over all T*rows steps of a batch (pyhanabi/r2d2.py:42-46, 99) and its weight gradient in
// fp32 on the CUDA cores (cuBLAS SIMT sgemm, 0.7 + 0.45 ms per update at rows = 256).  hb_gemm_nt is the same contraction
// at fp32-class accuracy on the tensor cores: C[M,N] = A[M,K] B[N,K]^T (+ bias) with both operands split into bf16
// hi/lo pairs and multiplied as hi*lo + lo*hi + hi*hi in the tcgen05 GEMM template of hb_gemm.cuh (EPI_F32).  Shapes are
// padded internally to the template's 128 x 256 x 64 tiles; when a problem has few output tiles and a long K (the weight
// gradient: [512, T*rows] x [T*rows, 838]) it is split over K into partial problems of ONE launch and summed afterwards,
// so that all SMs work.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <vector>

#include "hb_engine.h"
#include "hb_gemm.cuh"
#include "hb_gemm_host.h"

using hbg::Params;

int hb_upload_init(HbUploadRing* r) {
  if (r->base) return 0;
  HB_CUDA(cudaMallocHost((void**)&r->base, HbUploadRing::SLOTS * HbUploadRing::SLOT_BYTES));
  for (int i = 0; i < HbUploadRing::SLOTS; ++i) { HB_CUDA(cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming)); r->used[i] = false; }
  r->next = 0;
  return 0;
}
void hb_upload_destroy(HbUploadRing* r) {
  if (!r->base) return;
  for (int i = 0; i < HbUploadRing::SLOTS; ++i) cudaEventDestroy(r->ev[i]);
  cudaFreeHost(r->base);
  r->base = nullptr;
}
int hb_upload(HbUploadRing* r, void* dst_device, const void* src_host, size_t bytes, cudaStream_t st) {
  if (bytes > HbUploadRing::SLOT_BYTES || !r->base) {   // oversized (or no ring): the synchronous path is still correct
    HB_CUDA(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  const int slot = r->next;
  r->next = (r->next + 1) % HbUploadRing::SLOTS;
  if (r->used[slot]) HB_CUDA(cudaEventSynchronize(r->ev[slot]));   // the copy that last read this slot has executed
  unsigned char* p = r->base + (size_t)slot * HbUploadRing::SLOT_BYTES;
  memcpy(p, src_host, bytes);
  HB_CUDA(cudaMemcpyAsync(dst_device, p, bytes, cudaMemcpyHostToDevice, st));
  HB_CUDA(cudaEventRecord(r->ev[slot], st));
  r->used[slot] = true;
  return 0;
}

namespace {

constexpr int MAX_SPLIT = 16;

struct GemmScratch {            // grow-only device scratch per GPU (one learner process per GPU: calls are serialized by the caller)
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  size_t a_cap = 0, b_cap = 0;  // elements
  float* c_part = nullptr;      // [split][M_pad][N_pad]
  size_t c_cap = 0;
  Params* d_params = nullptr;   // [MAX_SPLIT]
  int* d_error = nullptr;
  int sm_count = 0;
  bool attrs = false;
  HbUploadRing ring;
  // the A operand the split buffers hold (hb_gemm_nt_same_a: the two networks' fc layers read the same observations)
  const float* last_a = nullptr; long long last_lda = 0; int last_m = 0, last_k = 0, last_kp = 0, last_ta = 0;
};
GemmScratch g_scratch[16];

// fp32 [rows][cols] (row stride ld) -> bf16 hi/lo [rows_pad][cols_pad], zero padded.  A thread converts FOUR consecutive
// columns: one 16-byte (or two 8-byte, rows only 8-byte aligned: ld even) load, two 8-byte stores -- a third of the memory
// instructions of the one-element-per-thread form, which ran at 40 % of the copy bandwidth (cols_pad is a multiple of 64).
__global__ void split_pad(const float* __restrict__ src, long long ld, int rows, int cols, int rows_pad, int cols_pad,
                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int align) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q_per_row = cols_pad >> 2;
  if (i4 >= (long long)rows_pad * q_per_row) return;
  const int r = (int)(i4 / q_per_row), c = (int)(i4 - (long long)r * q_per_row) << 2;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (r < rows) {
    const float* p = src + (long long)r * ld + c;
    if (c + 3 < cols && align == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if (c + 3 < cols && align == 2) {
      const float2 t0 = *reinterpret_cast<const float2*>(p), t1 = *reinterpret_cast<const float2*>(p + 2);
      v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (c + j < cols) v[j] = p[j];
    }
  }
  __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) hbg::split_bf16(v[j], h[j], l[j]);
  const long long o = (long long)r * cols_pad + c;
  *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l);
}

// The same for an operand given TRANSPOSED: src is [cols][rows] row-major (row stride ld), i.e. element (r, c) of the operand is
// src[c * ld + r] -- the weight-gradient GEMMs contract over the batch rows, which are the SLOW axis of the caller's tensors.
// 32 x 32 tiles through shared memory: coalesced reads along r, coalesced writes along c; no fp32 transposed copy in between.
__global__ void split_pad_t(const float* __restrict__ src, long long ld, int rows, int cols, int rows_pad, int cols_pad,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
  for (int j = ty; j < 32; j += 8) {   // tile[j][tx] = operand(r0 + tx, c0 + j) = src[(c0 + j) * ld + r0 + tx]
    const int r = r0 + tx, c = c0 + j;
    tile[j][tx] = (r < rows && c < cols) ? src[(long long)c * ld + r] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {   // write operand(r0 + j, c0 + tx)
    const int r = r0 + j, c = c0 + tx;
    if (r < rows_pad && c < cols_pad) {
      __nv_bfloat16 h, l;
      hbg::split_bf16(tile[tx][j], h, l);
      hi[(long long)r * cols_pad + c] = h;
      lo[(long long)r * cols_pad + c] = l;
    }
  }
}

// C[r][c] = sum_s part[s][r][c] (+ bias[c]) for r < M, c < N
__global__ void sum_parts(const float* __restrict__ part, int split, long long part_stride, int n_pad, const float* __restrict__ bias,
                          float* __restrict__ C, long long ldc, int M, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int r = (int)(i / N), c = (int)(i - (long long)r * N);
  float acc = bias ? bias[c] : 0.f;
  for (int s = 0; s < split; ++s) acc += part[(long long)s * part_stride + (long long)r * n_pad + c];
  C[(long long)r * ldc + c] = acc;
}

// widest vector load every row of `src` allows: 4 floats (16-byte aligned rows), 2, or 1
int split_align(const float* src, long long ld) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  if ((a & 15) == 0 && (ld & 3) == 0) return 4;
  if ((a & 7) == 0 && (ld & 1) == 0) return 2;
  return 1;
}

int grow(void** p, size_t* cap, size_t need, size_t elem) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  if (cudaMalloc(p, need * elem) != cudaSuccess) { hb_set_error("hb_gemm_nt: out of device memory (%zu bytes)", need * elem); return -2; }
  *cap = need;
  return 0;
}

}  // namespace

// C[M,N] = op(A) op(B)^T (+ bias): op(X) = X given [rows][K] (trans = 0, row stride ld >= K) or X given [K][rows] (trans = 1,
// row stride ld >= rows).  hb_gemm_nt (the public entry point) is the trans = 0 case.
static int gemm_nt_impl(int device, const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, const float* bias, float* C, int64_t ldc,
                        int M, int N, int K, void* stream, bool same_a) {
  if (!A || !B || !C) { hb_set_error("hb_gemm_nt: null argument"); return -1; }
  if (M < 1 || N < 1 || K < 1 || lda < (transA ? M : K) || ldb < (transB ? N : K) || ldc < N) { hb_set_error("hb_gemm_nt: bad shape M=%d N=%d K=%d lda=%lld ldb=%lld ldc=%lld", M, N, K, (long long)lda, (long long)ldb, (long long)ldc); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { hb_set_error("hb_gemm_nt: no CUDA device -- libhanabi_b200 has no CPU path"); return -2; }
  if (device < 0 || device >= ndev || device >= 16) { hb_set_error("hb_gemm_nt: bad device ordinal"); return -1; }
  HB_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  GemmScratch& S = g_scratch[device];
  if (!S.sm_count) {
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, device));
    S.sm_count = prop.multiProcessorCount;
    HB_CUDA(cudaMalloc((void**)&S.d_params, MAX_SPLIT * sizeof(Params)));
    HB_CUDA(cudaMalloc((void**)&S.d_error, sizeof(int)));
    HB_CUDA(cudaMemset(S.d_error, 0, sizeof(int)));
    { const int urc = hb_upload_init(&S.ring); if (urc) return urc; }
    HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_F32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
    HB_CUDA((cudaFuncSetAttribute(hbg::gemm3_kernel<hbg::EPI_F32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hbg::SMEM_BYTES)));
  }
  const int Mp = (M + hbg::BM - 1) / hbg::BM * hbg::BM, Np = (N + hbg::BN - 1) / hbg::BN * hbg::BN;
  const int mt = Mp / hbg::BM, nt = Np / hbg::BN;
  // split K so that the launch has about one tile per SM (each part at least 8 K-chunks)
  const int kc_total = (K + hbg::BK - 1) / hbg::BK;
  int split = 1;
  while (split < MAX_SPLIT && mt * nt * split * 2 <= S.sm_count && kc_total / (split * 2) >= 8) split *= 2;
  const int kc_part = (kc_total + split - 1) / split;
  const int Kp = kc_part * split * hbg::BK;
  const bool direct = split == 1 && Mp == M && Np == N && ldc == N;   // the template can write the caller's C itself
  int rc = 0;
  // Grow-only scratch, grown GEOMETRICALLY: a learner's batches get longer as its agent learns to survive (t_eff 10 -> 80 steps),
  // and every regrowth is a cudaFree + cudaMalloc, i.e. a device-wide synchronisation in the middle of an update.
  auto roomy = [](size_t need, size_t cap) { return need > cap + cap / 2 ? need : cap + cap / 2; };
  if (S.a_cap < (size_t)Mp * Kp) {   // hi and lo grow together (cudaFree synchronises the device: nothing queued still reads them)
    size_t c1 = 0, c2 = 0;
    const size_t want = roomy((size_t)Mp * Kp, S.a_cap);
    if (S.a_hi) cudaFree(S.a_hi);
    if (S.a_lo) cudaFree(S.a_lo);
    S.a_hi = S.a_lo = nullptr; S.a_cap = 0;
    rc |= grow((void**)&S.a_hi, &c1, want, sizeof(__nv_bfloat16));
    rc |= grow((void**)&S.a_lo, &c2, want, sizeof(__nv_bfloat16));
    if (rc) return rc;
    S.a_cap = want;
    S.last_a = nullptr;
  }
  if (S.b_cap < (size_t)Np * Kp) {
    size_t c1 = 0, c2 = 0;
    const size_t want = roomy((size_t)Np * Kp, S.b_cap);
    if (S.b_hi) cudaFree(S.b_hi);
    if (S.b_lo) cudaFree(S.b_lo);
    S.b_hi = S.b_lo = nullptr; S.b_cap = 0;
    rc |= grow((void**)&S.b_hi, &c1, want, sizeof(__nv_bfloat16));
    rc |= grow((void**)&S.b_lo, &c2, want, sizeof(__nv_bfloat16));
    if (rc) return rc;
    S.b_cap = want;
  }
  if (!direct && S.c_cap < (size_t)split * Mp * Np) { rc = grow((void**)&S.c_part, &S.c_cap, roomy((size_t)split * Mp * Np, S.c_cap), sizeof(float)); if (rc) return rc; }
  const long long na = (long long)Mp * Kp, nb = (long long)Np * Kp;
  if (same_a) {   // the split buffers still hold exactly this operand (no other call of this GPU's GEMM in between)
    if (S.last_a != A || S.last_lda != lda || S.last_m != M || S.last_k != K || S.last_kp != Kp || S.last_ta != transA) {
      hb_set_error("hb_gemm_nt_same_a: the previous call used a different A operand");
      return -1;
    }
  } else if (transA) split_pad_t<<<dim3((unsigned)((Kp + 31) / 32), (unsigned)((Mp + 31) / 32)), dim3(32, 8), 0, st>>>(A, lda, M, K, Mp, Kp, S.a_hi, S.a_lo);
  else split_pad<<<(unsigned)((na / 4 + 255) / 256), 256, 0, st>>>(A, lda, M, K, Mp, Kp, S.a_hi, S.a_lo, split_align(A, lda));
  S.last_a = A; S.last_lda = lda; S.last_m = M; S.last_k = K; S.last_kp = Kp; S.last_ta = transA;
  if (transB) split_pad_t<<<dim3((unsigned)((Kp + 31) / 32), (unsigned)((Np + 31) / 32)), dim3(32, 8), 0, st>>>(B, ldb, N, K, Np, Kp, S.b_hi, S.b_lo);
  else split_pad<<<(unsigned)((nb / 4 + 255) / 256), 256, 0, st>>>(B, ldb, N, K, Np, Kp, S.b_hi, S.b_lo, split_align(B, ldb));
  const int cl = (mt % 2 == 0) ? 2 : 1;
  std::vector<Params> hp(split);
  for (int s = 0; s < split; ++s) {
    Params& p = hp[s];
    memset(&p, 0, sizeof(p));
    const size_t koff = (size_t)s * kc_part * hbg::BK;
    rc |= hb_make_tmap(&p.a_hi[0], S.a_hi + koff, Mp, (uint64_t)kc_part * hbg::BK, hbg::BM, Kp);
    rc |= hb_make_tmap(&p.a_lo[0], S.a_lo + koff, Mp, (uint64_t)kc_part * hbg::BK, hbg::BM, Kp);
    p.a_hi[1] = p.a_hi[0]; p.a_lo[1] = p.a_lo[0];
    rc |= hb_make_tmap(&p.b_hi, S.b_hi + koff, Np, (uint64_t)kc_part * hbg::BK, hbg::BN / cl, Kp);
    rc |= hb_make_tmap(&p.b_lo, S.b_lo + koff, Np, (uint64_t)kc_part * hbg::BK, hbg::BN / cl, Kp);
    p.k_chunks = kc_part; p.k_chunks_seg0 = kc_part; p.lo_first = 0; p.lo_last = kc_part; p.split = 1;
    p.bias = direct ? bias : nullptr;
    p.c_f32 = direct ? C : S.c_part + (size_t)s * Mp * Np;
    p.ldc = direct ? (int)ldc : Np;
    p.error_flag = S.d_error; p.row_mul = 1; p.row_add = 0; p.valid_rows = Mp;
  }
  if (rc) return -2;
  rc = hb_upload(&S.ring, S.d_params, hp.data(), split * sizeof(Params), st);
  if (rc) return rc;
  if (cl == 2) rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_F32, 3>, 2, S.sm_count, st, S.d_params, nt, mt, split);
  else rc = hb_launch_gemm(hbg::gemm3_kernel<hbg::EPI_F32, 1>, 1, S.sm_count, st, S.d_params, nt, mt, split);
  if (rc) return rc;
  if (!direct) {
    const long long n = (long long)M * N;
    sum_parts<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.c_part, split, (long long)Mp * Np, Np, bias, C, ldc, M, N);
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}

int hb_gemm_nt_ex(int device, const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, const float* bias, float* C, int64_t ldc,
                  int M, int N, int K, void* stream) {
  return gemm_nt_impl(device, A, lda, transA, B, ldb, transB, bias, C, ldc, M, N, K, stream, false);
}

// hb_gemm_nt for a call whose A operand (pointer, shape AND contents) is the one of the immediately preceding GEMM call on this
// GPU: its bf16 hi / lo split is reused instead of recomputed (the online and the target network read the same observations).
int hb_gemm_nt_same_a(int device, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc, int M, int N, int K,
                      void* stream) {
  return gemm_nt_impl(device, A, lda, 0, B, ldb, 0, bias, C, ldc, M, N, K, stream, true);
}

extern "C" int hb_gemm_nt(int device, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                          int M, int N, int K, void* stream) {
  return hb_gemm_nt_ex(device, A, lda, 0, B, ldb, 0, bias, C, ldc, M, N, K, stream);
}
