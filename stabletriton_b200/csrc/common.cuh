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
void count_launch();  // atomic: entry points may be called from several host threads (one per device)

// Ordinal of the calling thread's current device, or -1 if it is outside [0, kMaxDevices) -- per-device state is
// never aliased onto slot 0.
constexpr int kMaxDevices = 64;
int current_device();

// One-time per-DEVICE set-up (cudaFuncSetAttribute applies to the current context only): `done(dev)` is false until
// `mark(dev)`; the guarded work is idempotent, so two threads racing on the same device merely repeat it.
struct PerDeviceOnce {
  bool done(int dev) const;
  void mark(int dev);
  unsigned long long mask_ = 0;  // accessed with atomic builtins
};

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row pitch ld (elements);
// box = [box_rows, 64 cols], 128-byte swizzle.  Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols = 64, bool swizzle128 = true);
// 4-D bf16 tensor map over an NHWC activation [N, H, W, C] (C contiguous); box = [Nt, Ht, Wt, 64].
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t Ht,
                   uint32_t Wt, uint32_t Nt = 1);

// Programmatic dependent launch (PDL): every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and executes `griddepcontrol.wait` before it touches
// global memory, so the launch latency and the prologue (barrier init, TMEM allocation, tensor-map
// prefetch) of kernel N+1 overlap the tail of kernel N -- inside CUDA graphs too (captured as programmatic
// edges).  ST_DISABLE_PDL=1 falls back to plain stream order.
bool pdl_enabled();

// cluster_x > 1 launches thread-block clusters of (cluster_x, 1, 1); grid.x must be a multiple of it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  return launch_kernel_cluster(kernel, grid, block, smem, stream, 1, args...);
}

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
    st::count_launch();                                                              \
  } while (0)

}  // namespace st
