// Host-side plumbing shared by every translation unit of libstabletriton_b200: error reporting for
// the C ABI, the driver entry point for cuTensorMapEncodeTiled (resolved at run time so the library
// has no link-time dependency on libcuda), and small tensor-map builders.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stabletriton_b200.h"

namespace st {

void set_error(const char* fmt, ...);
int device_sm_count();
unsigned long long* launch_counter();

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row pitch ld (elements);
// box = [box_rows, 64 cols], 128-byte swizzle.  Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols = 64);
// 4-D bf16 tensor map over an NHWC activation [N, H, W, C] (C contiguous); box = [Nt, Ht, Wt, 64].
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t Ht,
                   uint32_t Wt, uint32_t Nt = 1);

#define ST_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      st::set_error(__VA_ARGS__);      \
      return ST_ERR_INVALID_ARGUMENT;  \
    }                                  \
  } while (0)

#define ST_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      st::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));         \
      return ST_ERR_CUDA;                                                            \
    }                                                                                \
    ++*st::launch_counter();                                                         \
  } while (0)

}  // namespace st
