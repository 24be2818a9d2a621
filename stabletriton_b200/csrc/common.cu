// Library-level state of the C ABI: last-error string, launch accounting, tensor-map encoding.
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace st {

static thread_local char g_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return (dev >= 0 && dev < kMaxDevices) ? dev : -1;
}

bool PerDeviceOnce::done(int dev) const {
  return dev >= 0 && ((__atomic_load_n(&mask_, __ATOMIC_ACQUIRE) >> dev) & 1ull) != 0;
}
void PerDeviceOnce::mark(int dev) {
  if (dev >= 0) __atomic_fetch_or(&mask_, 1ull << dev, __ATOMIC_RELEASE);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ST_DISABLE_PDL");
    return !(e && e[0] && e[0] != '0');
  }();
  return on;
}

int device_sm_count() {
  static int cached[kMaxDevices] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool cacheable = dev >= 0 && dev < kMaxDevices;
  if (cacheable && cached[dev] != 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (cacheable) cached[dev] = n;  // benign race: every writer stores the same value
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ST_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u base=%p) failed: %d",
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, base,
              (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t Ht,
                   uint32_t Wt, uint32_t Nt) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ST_ERR_CUDA;
  }
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, Wt, Ht, Nt};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(nhwc N=%llu H=%llu W=%llu C=%llu box=%ux%u) failed: %d", (unsigned long long)N,
              (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, Ht, Wt, (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

}  // namespace st

extern "C" {

int st_version(void) { return ST_VERSION; }
const char* st_last_error_string(void) { return st::g_error; }
unsigned long long st_launch_count(void) { return st::g_launches.load(std::memory_order_relaxed); }
void st_reset_launch_count(void) { st::g_launches.store(0, std::memory_order_relaxed); }

}  // extern "C"
