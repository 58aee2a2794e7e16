// hb_gemm_host.h -- host-side helpers of the tcgen05 GEMM template (hb_gemm.cuh) shared by hb_policy.cu (act forward)
// and hb_lstm.cu (learner-side LSTM training kernels): TMA tensor-map construction and the persistent launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hbg { struct Params; }

// [rows][cols] bf16, cols contiguous; box = 64 columns (128 bytes, one swizzle span) x box_rows rows.  `row_stride`
// (elements, 0 = cols) is the distance between consecutive rows of the view.
int hb_make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint64_t row_stride = 0);

// Persistent launch of a gemm3_kernel instantiation: one CTA per SM (as many as there are work items if fewer), in
// clusters of `cl` CTAs; nt / mt = number of 256-column / 128-row tiles, nprob problems (ps[0..nprob)).
typedef void (*HbGemmKernel)(const hbg::Params*, int, int, int);
int hb_launch_gemm(HbGemmKernel k, int cl, int sm_count, cudaStream_t st, const hbg::Params* ps, int nt, int mt, int nprob);
