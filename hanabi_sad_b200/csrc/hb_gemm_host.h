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

// Asynchronous upload of small launch-parameter records (GEMM Params with their tensor maps, recurrence parameters): a
// cudaMemcpyAsync from PAGEABLE memory makes the host wait until the stream has drained ("a stream sync is performed before the
// copy is initiated"), i.e. the host would run in lock-step with the device at every such launch and the GPU would idle while
// the host prepares the next one.  The source is therefore first copied into a slot of a pinned ring; a slot is reused only
// after the copy that read it has executed (one event per slot, normally long complete).
struct HbUploadRing {
  static constexpr int SLOTS = 64;
  static constexpr size_t SLOT_BYTES = 24 * 1024;
  unsigned char* base;      // zero-initialised by the owner (a global, or a memset struct): null = not set up
  cudaEvent_t ev[SLOTS];
  bool used[SLOTS];
  int next;
};
int hb_upload_init(HbUploadRing* r);
void hb_upload_destroy(HbUploadRing* r);
int hb_upload(HbUploadRing* r, void* dst_device, const void* src_host, size_t bytes, cudaStream_t st);
