// Standalone bring-up / regression harness for libstabletriton_b200 (no torch, no Python): each
// sub-test runs one C-ABI entry point against a naive CUDA-core reference on seeded inputs and
// prints max-abs error (normalised by max |ref|) plus a CUDA-event timing.  Usage:
//   selftest gemm|conv|attn|norm|misc|all
// Run each group under `timeout` on a GPU box: a wrong mbarrier phase hangs instead of failing.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/stabletriton_b200.h"

extern "C" void st_debug_set_gemm_trace(void* buf);
extern "C" int st_debug_attention_occupancy(void);
extern "C" void st_debug_set_attention_trace(void* buf);
extern "C" int st_conv3x3_direct_bf16(const void*, long long, long long, long long, long long, const void*,
                                      const void*, void*, long long, long long, long long, long long, int, int, int,
                                      int, int, st_stream_t);

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
#define ST(x)                                                                   \
  do {                                                                          \
    int r_ = (x);                                                               \
    if (r_ != 0) {                                                              \
      printf("ST error %d (%s) at line %d\n", r_, st_last_error_string(), __LINE__); \
      exit(3);                                                                  \
    }                                                                           \
  } while (0)

static uint32_t g_seed = 12345;
static float frand() {  // uniform in [-1, 1)
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}
static __nv_bfloat16* dev_bf16(size_t n, float scale, std::vector<float>* host = nullptr) {
  std::vector<__nv_bfloat16> h(n);
  if (host) host->resize(n);
  for (size_t i = 0; i < n; ++i) {
    h[i] = __float2bfloat16(frand() * scale);
    if (host) (*host)[i] = __bfloat162float(h[i]);
  }
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  return d;
}
static std::vector<float> to_host(const __nv_bfloat16* d, size_t n) {
  std::vector<__nv_bfloat16> h(n);
  CK(cudaMemcpy(h.data(), d, n * 2, cudaMemcpyDeviceToHost));
  std::vector<float> f(n);
  for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(h[i]);
  return f;
}
static std::vector<float> to_host_f(const float* d, size_t n) {
  std::vector<float> f(n);
  CK(cudaMemcpy(f.data(), d, n * 4, cudaMemcpyDeviceToHost));
  return f;
}
static int g_fail = 0;
static void report(const char* name, const std::vector<float>& got, const std::vector<float>& ref, float tol,
                   float ms = -1.f, double work = 0, const char* unit = "") {
  double maxref = 0, maxerr = 0, dot = 0, ng = 0, nr = 0;
  size_t bad = 0, worst = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    maxref = fmax(maxref, fabs((double)ref[i]));
    const double e = fabs((double)got[i] - ref[i]);
    if (!(e <= maxerr)) {
      maxerr = e;
      worst = i;
    }
    if (got[i] != got[i]) ++bad;
    dot += (double)got[i] * ref[i];
    ng += (double)got[i] * got[i];
    nr += (double)ref[i] * ref[i];
  }
  const double rel = maxerr / (maxref > 0 ? maxref : 1);
  const double cosv = dot / (sqrt(ng) * sqrt(nr) + 1e-30);
  const bool ok = rel <= tol && bad == 0 && cosv > 0.999;
  printf("%-58s %s rel_err=%.3e cos=%.6f nan=%zu worst@%zu(got %.4f ref %.4f)", name, ok ? "PASS" : "FAIL", rel, cosv,
         bad, worst, got.empty() ? 0.f : got[worst], ref.empty() ? 0.f : ref[worst]);
  if (ms > 0) printf("  %.3f ms  %.1f %s", ms, work / (ms * 1e-3), unit);
  printf("\n");
  fflush(stdout);
  if (!ok) ++g_fail;
}

// Device time per call: `iters` calls captured into one CUDA graph and replayed, so host-side launch
// cost (tensor-map encoding, driver submit) is excluded -- the product runs inside a graph too.
// Note: back-to-back replays reuse the same buffers, so small cases are L2-warm.
static cudaStream_t g_stream = nullptr;
template <typename F>
static float time_ms(F f, int warm = 2, int iters = 10) {
  if (!g_stream) CK(cudaStreamCreate(&g_stream));
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < iters; ++i) f(g_stream);
  CK(cudaStreamEndCapture(g_stream, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  for (int i = 0; i < warm; ++i) CK(cudaGraphLaunch(exec, g_stream));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, g_stream));
  CK(cudaGraphLaunch(exec, g_stream));
  CK(cudaEventRecord(b, g_stream));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGraphExecDestroy(exec));
  CK(cudaGraphDestroy(graph));
  return ms / iters;
}

// ------------------------------------------------------------------------------------------------
// naive references
// ------------------------------------------------------------------------------------------------
__device__ float ref_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678f)); }
__device__ float ref_silu(float x) { return x / (1.f + expf(-x)); }

__global__ void ref_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, float* D, int M, int N,
                         int K, const __nv_bfloat16* bias, const __nv_bfloat16* res, int ldr, unsigned flags) {
  const int n_out = (flags & ST_EPI_GEGLU) ? N / 2 : N;
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= n_out || m >= M) return;
  float acc = 0.f, accg = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(A[(size_t)m * lda + k]);
    acc += a * __bfloat162float(W[(size_t)n * ldw + k]);
    if (flags & ST_EPI_GEGLU) accg += a * __bfloat162float(W[(size_t)(n + n_out) * ldw + k]);
  }
  if (bias) acc += __bfloat162float(bias[n]);
  if (flags & ST_EPI_GEGLU) {
    if (bias) accg += __bfloat162float(bias[n + n_out]);
    acc *= ref_gelu(accg);
  }
  if (flags & ST_EPI_SILU) acc = ref_silu(acc);
  if (res) acc += __bfloat162float(res[(size_t)m * ldr + n]);
  D[(size_t)m * n_out + n] = acc;
}

__global__ void ref_conv3x3(const __nv_bfloat16* x, const __nv_bfloat16* w, const __nv_bfloat16* bias, float* y,
                            int N, int H, int W, int C, int K, const __nv_bfloat16* temb, const __nv_bfloat16* res) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int pix = blockIdx.y;
  if (k >= K) return;
  const int q = pix % W, p = (pix / W) % H, n = pix / (W * H);
  float acc = bias ? __bfloat162float(bias[k]) : 0.f;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int ih = p + r - 1, iw = q + s - 1;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const __nv_bfloat16* xp = x + (((size_t)n * H + ih) * W + iw) * C;
      const __nv_bfloat16* wp = w + (((size_t)k * 3 + r) * 3 + s) * C;
      for (int c = 0; c < C; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
    }
  if (temb) acc += __bfloat162float(temb[(size_t)n * K + k]);
  if (res) acc += __bfloat162float(res[(size_t)pix * K + k]);
  y[(size_t)pix * K + k] = acc;
}

// one thread per (b, h, query): two-pass softmax in fp32
__global__ void ref_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk,
                              const __nv_bfloat16* v, int ldv, float* o, int B, int H, int Tq, int Tk, float scale) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int bh = blockIdx.y;
  if (t >= Tq) return;
  const int b = bh / H, h = bh % H;
  const __nv_bfloat16* qr = q + ((size_t)b * Tq + t) * ldq + h * 64;
  float qf[64];
  for (int d = 0; d < 64; ++d) qf[d] = __bfloat162float(qr[d]);
  float mx = -INFINITY;
  for (int j = 0; j < Tk; ++j) {
    const __nv_bfloat16* kr = k + ((size_t)b * Tk + j) * ldk + h * 64;
    float s = 0.f;
    for (int d = 0; d < 64; ++d) s += qf[d] * __bfloat162float(kr[d]);
    mx = fmaxf(mx, s * scale);
  }
  float acc[64];
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  float l = 0.f;
  for (int j = 0; j < Tk; ++j) {
    const __nv_bfloat16* kr = k + ((size_t)b * Tk + j) * ldk + h * 64;
    const __nv_bfloat16* vr = v + ((size_t)b * Tk + j) * ldv + h * 64;
    float s = 0.f;
    for (int d = 0; d < 64; ++d) s += qf[d] * __bfloat162float(kr[d]);
    const float pj = expf(s * scale - mx);
    l += pj;
    for (int d = 0; d < 64; ++d) acc[d] += pj * __bfloat162float(vr[d]);
  }
  float* orow = o + ((size_t)b * Tq + t) * (H * 64) + h * 64;
  for (int d = 0; d < 64; ++d) orow[d] = acc[d] / l;
}

// ------------------------------------------------------------------------------------------------
static void test_gemm_case(int M, int N, int K, unsigned flags, bool bias, bool res, int block_n, bool timeit) {
  const int n_out = (flags & ST_EPI_GEGLU) ? N / 2 : N;
  __nv_bfloat16* A = dev_bf16((size_t)M * K, 1.0f);
  __nv_bfloat16* W = dev_bf16((size_t)N * K, 0.05f);
  __nv_bfloat16* b = bias ? dev_bf16(N, 0.5f) : nullptr;
  __nv_bfloat16* r = res ? dev_bf16((size_t)M * n_out, 1.0f) : nullptr;
  __nv_bfloat16* D;
  float* Dref;
  CK(cudaMalloc(&D, (size_t)M * n_out * 2));
  CK(cudaMemset(D, 0xff, (size_t)M * n_out * 2));
  CK(cudaMalloc(&Dref, (size_t)M * n_out * 4));
  ST(st_gemm_bf16(A, K, W, K, D, n_out, M, N, K, b, r, n_out, flags, block_n, nullptr, 0));
  CK(cudaDeviceSynchronize());
  ref_gemm<<<dim3((n_out + 127) / 128, M), 128>>>(A, K, W, K, Dref, M, N, K, b, r, n_out, flags);
  CK(cudaDeviceSynchronize());
  float ms = -1;
  if (timeit) ms = time_ms([&](cudaStream_t s) { st_gemm_bf16(A, K, W, K, D, n_out, M, N, K, b, r, n_out, flags, block_n, nullptr, s); });
  char name[128];
  snprintf(name, sizeof name, "gemm M=%d N=%d K=%d flags=%u bias=%d res=%d bn=%d", M, N, K, flags, bias, res, block_n);
  report(name, to_host(D, (size_t)M * n_out), to_host_f(Dref, (size_t)M * n_out), 1e-2f, ms, 2.0 * M * N * K * 1e-12,
         "TFLOP/s");
  cudaFree(A);
  cudaFree(W);
  cudaFree(D);
  cudaFree(Dref);
  if (b) cudaFree(b);
  if (r) cudaFree(r);
}

static void test_gemm() {
  // bring-up: smallest case first, one tile, one k-block
  test_gemm_case(128, 128, 64, 0, false, false, 128, false);
  test_gemm_case(128, 128, 256, 0, false, false, 128, false);
  test_gemm_case(128, 64, 128, 0, true, false, 64, false);
  test_gemm_case(256, 256, 128, 0, true, false, 256, false);
  test_gemm_case(154, 640, 2048, 0, false, false, 0, false);   // cross-attn K/V projection, ragged M
  test_gemm_case(384, 320, 320, ST_EPI_SILU, true, false, 0, false);
  test_gemm_case(300, 200, 192, 0, true, true, 128, false);    // ragged M and N, residual
  test_gemm_case(256, 512, 128, ST_EPI_GEGLU, true, false, 128, false);
  test_gemm_case(300, 328, 192, 0, true, true, 192, false);
  test_gemm_case(130, 72, 64, ST_EPI_SILU, true, true, 64, false);
  test_gemm_case(256, 1280, 320, ST_EPI_GEGLU, true, false, 256, false);
  // SDXL shapes, timed
  test_gemm_case(128, 128, 64, ST_W_STATIC, false, false, 128, false);      // K shorter than the prefetch depth
  test_gemm_case(300, 328, 192, ST_W_STATIC, true, true, 192, false);
  test_gemm_case(256, 1280, 320, ST_EPI_GEGLU | ST_W_STATIC, true, false, 256, false);
  test_gemm_case(2048, 1280, 1280, 0, true, true, 0, true);
  test_gemm_case(2048, 1280, 1280, ST_W_STATIC, true, true, 0, true);
  test_gemm_case(2048, 3840, 1280, ST_W_STATIC, false, false, 0, true);
  test_gemm_case(2048, 10240, 1280, ST_EPI_GEGLU | ST_W_STATIC, true, false, 0, true);
  test_gemm_case(2048, 1280, 5120, ST_W_STATIC, true, true, 0, true);
  test_gemm_case(2048, 1280, 1280, 0, true, true, 64, true);
  test_gemm_case(2048, 1280, 1280, 0, true, true, 128, true);
  test_gemm_case(2048, 1280, 1280, 0, true, true, 256, true);
  test_gemm_case(1000, 1280, 5120, 0, true, true, 0, true);    // stream-K with ragged M
  test_gemm_case(2048, 640, 4096, ST_EPI_SILU, true, false, 0, true);  // stream-K, 48 tiles
  test_gemm_case(2048, 3840, 1280, 0, false, false, 0, true);
  test_gemm_case(2048, 10240, 1280, ST_EPI_GEGLU, true, false, 0, true);
  test_gemm_case(2048, 10240, 1280, ST_EPI_GEGLU, true, false, 128, true);
  test_gemm_case(2048, 1280, 5120, 0, true, true, 0, true);
  test_gemm_case(8192, 640, 640, 0, true, true, 0, true);
  test_gemm_case(8192, 5120, 640, ST_EPI_GEGLU, true, false, 0, true);
  test_gemm_case(8192, 640, 2560, 0, true, true, 0, true);
  test_gemm_case(8192, 8192, 8192, 0, false, false, 256, true);
}

static void test_conv_case(int N, int H, int W, int C, int K, bool temb, bool res, int block_n, bool timeit) {
  __nv_bfloat16* x = dev_bf16((size_t)N * H * W * C, 1.0f);
  __nv_bfloat16* w = dev_bf16((size_t)K * 9 * C, 0.03f);
  __nv_bfloat16* b = dev_bf16(K, 0.5f);
  __nv_bfloat16* t = temb ? dev_bf16((size_t)N * K, 1.0f) : nullptr;
  __nv_bfloat16* r = res ? dev_bf16((size_t)N * H * W * K, 1.0f) : nullptr;
  __nv_bfloat16* y;
  float* yref;
  const size_t out = (size_t)N * H * W * K;
  CK(cudaMalloc(&y, out * 2));
  CK(cudaMemset(y, 0xff, out * 2));
  CK(cudaMalloc(&yref, out * 4));
  ST(st_conv3x3_nhwc_bf16(x, w, b, y, N, H, W, C, K, t, K, r, ST_W_STATIC, block_n, nullptr, 0));
  CK(cudaDeviceSynchronize());
  ref_conv3x3<<<dim3((K + 127) / 128, N * H * W), 128>>>(x, w, b, yref, N, H, W, C, K, t, r);
  CK(cudaDeviceSynchronize());
  float ms = -1;
  if (timeit) ms = time_ms([&](cudaStream_t s) { st_conv3x3_nhwc_bf16(x, w, b, y, N, H, W, C, K, t, K, r, ST_W_STATIC, block_n, nullptr, s); });
  char name[128];
  snprintf(name, sizeof name, "conv3x3 N=%d H=%d W=%d C=%d K=%d temb=%d res=%d bn=%d", N, H, W, C, K, temb, res,
           block_n);
  report(name, to_host(y, out), to_host_f(yref, out), 1e-2f, ms, 2.0 * N * H * W * (double)K * C * 9 * 1e-12,
         "TFLOP/s");
  cudaFree(x);
  cudaFree(w);
  cudaFree(b);
  cudaFree(y);
  cudaFree(yref);
  if (t) cudaFree(t);
  if (r) cudaFree(r);
}

static void test_conv() {
  test_conv_case(1, 16, 16, 64, 64, false, false, 64, false);
  test_conv_case(2, 32, 32, 128, 128, true, false, 128, false);
  test_conv_case(1, 64, 64, 64, 320, false, true, 0, false);
  test_conv_case(1, 8, 128, 64, 64, true, true, 64, false);
  test_conv_case(1, 4, 256, 64, 64, false, false, 64, false);
  test_conv_case(3, 8, 8, 64, 64, true, true, 64, false);
  test_conv_case(1, 4, 4, 128, 64, true, false, 64, false);
  test_conv_case(2, 128, 128, 320, 320, true, false, 0, true);
  test_conv_case(2, 64, 64, 640, 640, false, true, 0, true);
  test_conv_case(2, 32, 32, 1280, 1280, true, false, 0, true);
  test_conv_case(2, 32, 32, 2560, 1280, true, false, 0, true);
}

// k[b][t][:] *= 1 + growth * t / Tk: later K/V blocks carry much larger scores, which forces the running
// softmax reference to move many times (the lazy-rescale path of the pipelined kernel)
__global__ void grow_rows(__nv_bfloat16* k, int T, int C, float growth) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int t = (int)((i / C) % T);
  k[i] = __float2bfloat16(__bfloat162float(k[i]) * (1.f + growth * t / T));
}

static void test_attn_case(int B, int H, int Tq, int Tk, bool fused_qkv, bool timeit, float growth = 0.f) {
  const int C = H * 64;
  const int ld = fused_qkv ? 3 * C : C;
  __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* base = nullptr;
  if (fused_qkv) {
    base = dev_bf16((size_t)B * Tq * ld, 1.5f);
    q = base;
    k = base + C;
    v = base + 2 * C;
  } else {
    q = dev_bf16((size_t)B * Tq * C, 1.5f);
    k = dev_bf16((size_t)B * Tk * C, 1.5f);
    v = dev_bf16((size_t)B * Tk * C, 1.0f);
    if (growth != 0.f) grow_rows<<<(unsigned)((size_t)B * Tk * C / 64), 64>>>(k, Tk, C, growth);
  }
  __nv_bfloat16* o;
  float* oref;
  const size_t out = (size_t)B * Tq * C;
  CK(cudaMalloc(&o, out * 2));
  CK(cudaMemset(o, 0xff, out * 2));
  CK(cudaMalloc(&oref, out * 4));
  const float scale = 0.125f;
  ST(st_attention_bf16(q, (long long)Tq * ld, 64, ld, k, (long long)Tk * ld, 64, ld, v, (long long)Tk * ld, 64, ld, o,
                       (long long)Tq * C, 64, C, B, H, Tq, Tk, scale, 0));
  CK(cudaDeviceSynchronize());
  ref_attention<<<dim3((Tq + 63) / 64, B * H), 64>>>(q, ld, k, ld, v, ld, oref, B, H, Tq, Tk, scale);
  CK(cudaDeviceSynchronize());
  float ms = -1;
  if (timeit) ms = time_ms([&](cudaStream_t s) { st_attention_bf16(q, (long long)Tq * ld, 64, ld, k, (long long)Tk * ld, 64, ld, v, (long long)Tk * ld, 64, ld, o,
                        (long long)Tq * C, 64, C, B, H, Tq, Tk, scale, s);
    });
  char name[128];
  snprintf(name, sizeof name, "attention B=%d H=%d Tq=%d Tk=%d fusedqkv=%d growth=%g", B, H, Tq, Tk, fused_qkv, growth);
  report(name, to_host(o, out), to_host_f(oref, out), 2e-2f, ms, 4.0 * B * H * (double)Tq * Tk * 64 * 1e-12,
         "TFLOP/s");
  if (fused_qkv)
    cudaFree(base);
  else {
    cudaFree(q);
    cudaFree(k);
    cudaFree(v);
  }
  cudaFree(o);
  cudaFree(oref);
}

static void test_attn() {
  test_attn_case(1, 1, 128, 128, false, false);
  test_attn_case(1, 2, 128, 256, false, false);
  test_attn_case(1, 2, 256, 512, false, false);
  test_attn_case(2, 3, 200, 77, false, false);
  test_attn_case(1, 2, 384, 333, false, false);
  test_attn_case(2, 2, 256, 256, true, false);
  test_attn_case(1, 2, 256, 1024, false, false, 24.f);   // reference max moves in (almost) every block
  test_attn_case(1, 2, 256, 1000, false, false, -0.9f);  // scores shrink: the reference never moves; ragged tail
  test_attn_case(2, 3, 130, 129, false, false, 6.f);     // second block holds a single key
  test_attn_case(2, 20, 1024, 1024, true, true);
  test_attn_case(2, 10, 4096, 4096, true, true);
  test_attn_case(2, 10, 4096, 77, false, true);
  test_attn_case(2, 20, 1024, 77, false, true);
}

// host references for the norms
static void test_groupnorm_case(int N, int HW, int C, int G, bool silu, float offset, bool timeit) {
  std::vector<float> hx, hg, hb;
  __nv_bfloat16* x = dev_bf16((size_t)N * HW * C, 1.0f, &hx);
  if (offset != 0.f) {  // shift the data to stress the variance computation
    std::vector<__nv_bfloat16> tmp(hx.size());
    for (size_t i = 0; i < hx.size(); ++i) {
      tmp[i] = __float2bfloat16(hx[i] + offset);
      hx[i] = __bfloat162float(tmp[i]);
    }
    CK(cudaMemcpy(x, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
  }
  __nv_bfloat16* g = dev_bf16(C, 1.0f, &hg);
  __nv_bfloat16* b = dev_bf16(C, 0.5f, &hb);
  __nv_bfloat16* y;
  CK(cudaMalloc(&y, (size_t)N * HW * C * 2));
  void* ws;
  CK(cudaMalloc(&ws, st_groupnorm_workspace_bytes(N, HW, C, G)));
  ST(st_groupnorm_nhwc_bf16(x, y, g, b, ws, N, HW, C, G, 1e-5f, silu, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> ref((size_t)N * HW * C);
  const int cpg = C / G;
  for (int n = 0; n < N; ++n)
    for (int gi = 0; gi < G; ++gi) {
      double s = 0, ss = 0;
      for (int p = 0; p < HW; ++p)
        for (int c = 0; c < cpg; ++c) s += hx[((size_t)n * HW + p) * C + gi * cpg + c];
      const double mean = s / ((double)HW * cpg);
      for (int p = 0; p < HW; ++p)
        for (int c = 0; c < cpg; ++c) {
          const double d = hx[((size_t)n * HW + p) * C + gi * cpg + c] - mean;
          ss += d * d;
        }
      const double rstd = 1.0 / sqrt(ss / ((double)HW * cpg) + 1e-5);
      for (int p = 0; p < HW; ++p)
        for (int c = 0; c < cpg; ++c) {
          const size_t i = ((size_t)n * HW + p) * C + gi * cpg + c;
          double v = (hx[i] - mean) * rstd * hg[gi * cpg + c] + hb[gi * cpg + c];
          if (silu) v = v / (1.0 + exp(-v));
          ref[i] = (float)v;
        }
    }
  float ms = -1;
  if (timeit) ms = time_ms([&](cudaStream_t s) { st_groupnorm_nhwc_bf16(x, y, g, b, ws, N, HW, C, G, 1e-5f, silu, s); }, 2, 20);
  char name[128];
  snprintf(name, sizeof name, "groupnorm N=%d HW=%d C=%d G=%d silu=%d off=%.0f", N, HW, C, G, silu, offset);
  report(name, to_host(y, ref.size()), ref, 1.5e-2f, ms, 4.0 * N * HW * (double)C * 1e-9, "GB/s");
  cudaFree(x);
  cudaFree(g);
  cudaFree(b);
  cudaFree(y);
  cudaFree(ws);
}

static void test_layernorm_case(int M, int N, bool timeit) {
  std::vector<float> hx, hg, hb;
  __nv_bfloat16* x = dev_bf16((size_t)M * N, 2.0f, &hx);
  __nv_bfloat16* g = dev_bf16(N, 1.0f, &hg);
  __nv_bfloat16* b = dev_bf16(N, 0.5f, &hb);
  __nv_bfloat16* y;
  CK(cudaMalloc(&y, (size_t)M * N * 2));
  ST(st_layernorm_bf16(x, N, y, N, g, b, M, N, 1e-5f, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> ref((size_t)M * N);
  for (int m = 0; m < M; ++m) {
    double s = 0, ss = 0;
    for (int n = 0; n < N; ++n) s += hx[(size_t)m * N + n];
    const double mean = s / N;
    for (int n = 0; n < N; ++n) {
      const double d = hx[(size_t)m * N + n] - mean;
      ss += d * d;
    }
    const double rstd = 1.0 / sqrt(ss / N + 1e-5);
    for (int n = 0; n < N; ++n) ref[(size_t)m * N + n] = (float)((hx[(size_t)m * N + n] - mean) * rstd * hg[n] + hb[n]);
  }
  float ms = -1;
  if (timeit) ms = time_ms([&](cudaStream_t s) { st_layernorm_bf16(x, N, y, N, g, b, M, N, 1e-5f, s); }, 2, 20);
  char name[128];
  snprintf(name, sizeof name, "layernorm M=%d N=%d", M, N);
  report(name, to_host(y, ref.size()), ref, 1e-2f, ms, 4.0 * M * (double)N * 1e-9, "GB/s");
  cudaFree(x);
  cudaFree(g);
  cudaFree(b);
  cudaFree(y);
}

static void test_norm() {
  test_groupnorm_case(1, 64, 64, 8, false, 0.f, false);
  test_groupnorm_case(2, 256, 320, 32, true, 0.f, false);
  test_groupnorm_case(2, 1024, 960, 32, true, 50.f, false);
  test_groupnorm_case(1, 100, 2560, 32, false, 0.f, false);
  test_groupnorm_case(2, 16384, 320, 32, true, 0.f, true);
  test_groupnorm_case(2, 4096, 640, 32, true, 0.f, true);
  test_groupnorm_case(2, 1024, 1280, 32, true, 0.f, true);
  test_groupnorm_case(16, 16384, 320, 32, true, 0.f, true);
  test_layernorm_case(77, 640, false);
  test_layernorm_case(8192, 640, true);
  test_layernorm_case(2048, 1280, true);
  test_layernorm_case(65536, 1280, true);
}

static void test_misc() {
  // tiny-M linear
  {
    const int M = 2, N = 1280, K = 2816;
    std::vector<float> hx, hw, hb;
    __nv_bfloat16* x = dev_bf16((size_t)M * K, 1.0f, &hx);
    __nv_bfloat16* w = dev_bf16((size_t)N * K, 0.03f, &hw);
    __nv_bfloat16* b = dev_bf16(N, 0.5f, &hb);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)M * N * 2));
    ST(st_linear_small_m_bf16(x, K, w, K, b, y, N, M, N, K, 1, 1, ST_W_STATIC, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)M * N);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = hb[n];
        for (int k = 0; k < K; ++k) {
          const double xv = hx[(size_t)m * K + k];
          acc += xv / (1.0 + exp(-xv)) * hw[(size_t)n * K + k];
        }
        ref[(size_t)m * N + n] = (float)(acc / (1.0 + exp(-acc)));
      }
    report("linear_small_m M=2 N=1280 K=2816 silu_in silu_out", to_host(y, ref.size()), ref, 1e-2f);
  }
  // conv_in style (C=4, NCHW input) and conv_out style (K=4, NCHW output)
  {
    const int N = 2, H = 16, W = 16, C = 4, K = 320;
    std::vector<float> hx, hw, hb;
    __nv_bfloat16* x = dev_bf16((size_t)N * C * H * W, 1.0f, &hx);  // NCHW
    __nv_bfloat16* w = dev_bf16((size_t)K * 9 * C, 0.1f, &hw);
    __nv_bfloat16* b = dev_bf16(K, 0.5f, &hb);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)N * H * W * K * 2));
    ST(st_conv3x3_direct_bf16(x, (long long)C * H * W, W, 1, (long long)H * W, w, b, y, (long long)H * W * K,
                              (long long)W * K, K, 1, N, H, W, C, K, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)N * H * W * K);
    for (int n = 0; n < N; ++n)
      for (int p = 0; p < H; ++p)
        for (int q = 0; q < W; ++q)
          for (int k = 0; k < K; ++k) {
            double acc = hb[k];
            for (int r = 0; r < 3; ++r)
              for (int s = 0; s < 3; ++s) {
                const int ih = p + r - 1, iw = q + s - 1;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                for (int c = 0; c < C; ++c)
                  acc += hx[(((size_t)n * C + c) * H + ih) * W + iw] * hw[(((size_t)k * 3 + r) * 3 + s) * C + c];
              }
            ref[(((size_t)n * H + p) * W + q) * K + k] = (float)acc;
          }
    report("conv3x3_direct C=4 (NCHW in) K=320", to_host(y, ref.size()), ref, 1e-2f);
  }
  {
    const int N = 2, H = 16, W = 16, C = 320, K = 4;
    std::vector<float> hx, hw, hb;
    __nv_bfloat16* x = dev_bf16((size_t)N * H * W * C, 1.0f, &hx);  // NHWC
    __nv_bfloat16* w = dev_bf16((size_t)K * 9 * C, 0.02f, &hw);
    __nv_bfloat16* b = dev_bf16(8, 0.5f, &hb);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)N * K * H * W * 2));
    ST(st_conv3x3_direct_bf16(x, (long long)H * W * C, (long long)W * C, C, 1, w, b, y, (long long)K * H * W, W, 1,
                              (long long)H * W, N, H, W, C, K, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)N * K * H * W);
    for (int n = 0; n < N; ++n)
      for (int p = 0; p < H; ++p)
        for (int q = 0; q < W; ++q)
          for (int k = 0; k < K; ++k) {
            double acc = hb[k];
            for (int r = 0; r < 3; ++r)
              for (int s = 0; s < 3; ++s) {
                const int ih = p + r - 1, iw = q + s - 1;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                for (int c = 0; c < C; ++c)
                  acc += hx[(((size_t)n * H + ih) * W + iw) * C + c] * hw[(((size_t)k * 3 + r) * 3 + s) * C + c];
              }
            ref[(((size_t)n * K + k) * H + p) * W + q] = (float)acc;
          }
    report("conv3x3_direct C=320 K=4 (NCHW out)", to_host(y, ref.size()), ref, 1e-2f);
  }
  // stride-2 conv through im2col + gemm
  {
    const int N = 2, H = 32, W = 32, C = 64, K = 64, Ho = 16, Wo = 16;
    std::vector<float> hx, hw;
    __nv_bfloat16* x = dev_bf16((size_t)N * H * W * C, 1.0f, &hx);
    __nv_bfloat16* w = dev_bf16((size_t)K * 9 * C, 0.05f, &hw);
    __nv_bfloat16 *col, *y;
    CK(cudaMalloc(&col, (size_t)N * Ho * Wo * 9 * C * 2));
    CK(cudaMalloc(&y, (size_t)N * Ho * Wo * K * 2));
    ST(st_im2col3x3_nhwc_bf16(x, col, N, H, W, C, 2, 0));
    ST(st_gemm_bf16(col, 9 * C, w, 9 * C, y, K, N * Ho * Wo, K, 9 * C, nullptr, nullptr, 0, 0, 0, nullptr, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)N * Ho * Wo * K);
    for (int n = 0; n < N; ++n)
      for (int p = 0; p < Ho; ++p)
        for (int q = 0; q < Wo; ++q)
          for (int k = 0; k < K; ++k) {
            double acc = 0;
            for (int r = 0; r < 3; ++r)
              for (int s = 0; s < 3; ++s) {
                const int ih = 2 * p + r - 1, iw = 2 * q + s - 1;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                for (int c = 0; c < C; ++c)
                  acc += hx[(((size_t)n * H + ih) * W + iw) * C + c] * hw[(((size_t)k * 3 + r) * 3 + s) * C + c];
              }
            ref[(((size_t)n * Ho + p) * Wo + q) * K + k] = (float)acc;
          }
    report("conv3x3 stride 2 via im2col + gemm", to_host(y, ref.size()), ref, 1e-2f);
  }
  // geglu, upsample, concat, timestep embedding
  {
    const int R = 64, Cc = 256;
    std::vector<float> hs, hg;
    __nv_bfloat16* s = dev_bf16((size_t)R * Cc, 2.0f, &hs);
    __nv_bfloat16* g = dev_bf16((size_t)R * Cc, 2.0f, &hg);
    __nv_bfloat16* o;
    CK(cudaMalloc(&o, (size_t)R * Cc * 2));
    ST(st_geglu_bf16(s, Cc, g, Cc, o, Cc, R, Cc, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)R * Cc);
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = (float)(hs[i] * 0.5 * hg[i] * (1.0 + erf(hg[i] / sqrt(2.0))));
    report("geglu standalone", to_host(o, ref.size()), ref, 1e-2f);
  }
  {
    const int N = 2, H = 8, W = 8, C = 64;
    std::vector<float> hx;
    __nv_bfloat16* x = dev_bf16((size_t)N * H * W * C, 1.0f, &hx);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)N * 4 * H * W * C * 2));
    ST(st_upsample_nearest2x_nhwc_bf16(x, y, N, H, W, C, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)N * 4 * H * W * C);
    for (int n = 0; n < N; ++n)
      for (int p = 0; p < 2 * H; ++p)
        for (int q = 0; q < 2 * W; ++q)
          for (int c = 0; c < C; ++c)
            ref[(((size_t)n * 2 * H + p) * 2 * W + q) * C + c] = hx[(((size_t)n * H + p / 2) * W + q / 2) * C + c];
    report("upsample nearest 2x", to_host(y, ref.size()), ref, 0.f);
  }
  {
    const int P = 100, Ca = 64, Cb = 128;
    std::vector<float> ha, hb;
    __nv_bfloat16* a = dev_bf16((size_t)P * Ca, 1.0f, &ha);
    __nv_bfloat16* b = dev_bf16((size_t)P * Cb, 1.0f, &hb);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)P * (Ca + Cb) * 2));
    ST(st_concat_channels_bf16(a, Ca, b, Cb, y, P, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)P * (Ca + Cb));
    for (int p = 0; p < P; ++p) {
      for (int c = 0; c < Ca; ++c) ref[(size_t)p * (Ca + Cb) + c] = ha[(size_t)p * Ca + c];
      for (int c = 0; c < Cb; ++c) ref[(size_t)p * (Ca + Cb) + Ca + c] = hb[(size_t)p * Cb + c];
    }
    report("concat channels", to_host(y, ref.size()), ref, 0.f);
  }
  {
    const int B = 2, half = 160;
    float ht[2] = {999.f, 500.f};
    float* t;
    CK(cudaMalloc(&t, 8));
    CK(cudaMemcpy(t, ht, 8, cudaMemcpyHostToDevice));
    __nv_bfloat16* o;
    CK(cudaMalloc(&o, B * 2 * half * 2));
    ST(st_timestep_embedding_bf16(t, o, 2 * half, B, half, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> ref(B * 2 * half);
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < half; ++j) {
        const float f = expf(-logf(10000.f) * j / half);
        ref[b * 2 * half + j] = cosf(ht[b] * f);
        ref[b * 2 * half + half + j] = sinf(ht[b] * f);
      }
    report("timestep embedding", to_host(o, ref.size()), ref, 1e-2f);
  }
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, sm_%d%d, %d SMs, lib version %d\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, st_version());
  {
    void* ws = nullptr;  // stream-K scratch (leaked on purpose: lives for the whole process)
    CK(cudaMalloc(&ws, st_workspace_bytes()));
    ST(st_set_workspace(ws, st_workspace_bytes()));
  }
  if (!strcmp(what, "trace") && argc >= 7) {  // selftest trace M N K flags block_n : per-CTA phase timeline
    const int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]);
    const unsigned flags = (unsigned)atoi(argv[5]);
    const int bn = atoi(argv[6]);
    const int n_out = (flags & ST_EPI_GEGLU) ? N / 2 : N;
    __nv_bfloat16* A = dev_bf16((size_t)M * K, 1.0f);
    __nv_bfloat16* W = dev_bf16((size_t)N * K, 0.05f);
    __nv_bfloat16* b = dev_bf16(N, 0.5f);
    __nv_bfloat16* r = (argc > 7 && !atoi(argv[7])) ? nullptr : dev_bf16((size_t)M * n_out, 1.0f);  // [res: 1 (default) / 0]
    __nv_bfloat16* D;
    CK(cudaMalloc(&D, (size_t)M * n_out * 2));
    unsigned long long* tr;
    CK(cudaMalloc(&tr, 148 * 12 * 8));
    for (int it = 0; it < 3; ++it) ST(st_gemm_bf16(A, K, W, K, D, n_out, M, N, K, b, r, n_out, flags, bn, nullptr, 0));
    CK(cudaMemset(tr, 0, 148 * 12 * 8));
    st_debug_set_gemm_trace(tr);
    ST(st_gemm_bf16(A, K, W, K, D, n_out, M, N, K, b, r, n_out, flags, bn, nullptr, 0));
    CK(cudaDeviceSynchronize());
    st_debug_set_gemm_trace(nullptr);
    std::vector<unsigned long long> h(148 * 12);
    CK(cudaMemcpy(h.data(), tr, 148 * 12 * 8, cudaMemcpyDeviceToHost));
    const char* names[11] = {"setup", "first operands landed", "MMA issue done (tile 0)", "accumulator ready",
                             "epilogue tile 0 done", "exit", "epi: chunk0 tmem ld done", "epi: chunk0 staging free",
                             "epi: chunk0 math+sts done", "epi: chunk1 math+sts done", "epi: group0 store issued"};
    for (int e = 1; e <= 11; ++e) {
      double sum = 0, mn = 1e18, mx = 0;
      int cnt = 0;
      for (int c = 0; c < 148; ++c) {
        if (!h[c * 12] || !h[c * 12 + e]) continue;
        const double d = (double)(h[c * 12 + e] - h[c * 12]);
        sum += d;
        mn = d < mn ? d : mn;
        mx = d > mx ? d : mx;
        ++cnt;
      }
      printf("  t[%d] %-26s cycles since CTA start: min %8.0f avg %8.0f max %8.0f (n=%d)\n", e, names[e - 1], mn,
             cnt ? sum / cnt : 0, mx, cnt);
    }
    return 0;
  }
  if (!strcmp(what, "ctrace") && argc >= 8) {  // selftest ctrace N H W C K block_n : per-CTA phase timeline of a 3x3 conv
    const int N = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]), C = atoi(argv[5]), K = atoi(argv[6]);
    const int bn = atoi(argv[7]);
    __nv_bfloat16* x = dev_bf16((size_t)N * H * W * C, 1.0f);
    __nv_bfloat16* w = dev_bf16((size_t)K * 9 * C, 0.03f);
    __nv_bfloat16* b = dev_bf16(K, 0.5f);
    __nv_bfloat16* t = dev_bf16((size_t)N * K, 1.0f);
    __nv_bfloat16* y;
    CK(cudaMalloc(&y, (size_t)N * H * W * K * 2));
    unsigned long long* tr;
    CK(cudaMalloc(&tr, 148 * 12 * 8));
    for (int it = 0; it < 3; ++it) ST(st_conv3x3_nhwc_bf16(x, w, b, y, N, H, W, C, K, t, K, nullptr, ST_W_STATIC, bn, nullptr, 0));
    CK(cudaMemset(tr, 0, 148 * 12 * 8));
    st_debug_set_gemm_trace(tr);
    ST(st_conv3x3_nhwc_bf16(x, w, b, y, N, H, W, C, K, t, K, nullptr, ST_W_STATIC, bn, nullptr, 0));
    CK(cudaDeviceSynchronize());
    st_debug_set_gemm_trace(nullptr);
    std::vector<unsigned long long> h(148 * 12);
    CK(cudaMemcpy(h.data(), tr, 148 * 12 * 8, cudaMemcpyDeviceToHost));
    const char* names[6] = {"setup", "first operands landed", "MMA issue done (tile 0)", "accumulator ready",
                            "epilogue tile 0 done", "exit"};
    printf("conv3x3 N=%d %dx%d C=%d K=%d bn=%d: %d k-blocks per tile\n", N, H, W, C, K, bn, 9 * C / 64);
    for (int e = 1; e <= 6; ++e) {
      double sum = 0, mn = 1e18, mx = 0;
      int cnt = 0;
      for (int c = 0; c < 148; ++c) {
        if (!h[c * 12] || !h[c * 12 + e]) continue;
        const double d = (double)(h[c * 12 + e] - h[c * 12]);
        sum += d;
        mn = d < mn ? d : mn;
        mx = d > mx ? d : mx;
        ++cnt;
      }
      printf("  t[%d] %-26s cycles since CTA start: min %8.0f avg %8.0f max %8.0f (n=%d)\n", e, names[e - 1], mn,
             cnt ? sum / cnt : 0, mx, cnt);
    }
    return 0;
  }
  if (!strcmp(what, "gemm1") && argc >= 7) {  // selftest gemm1 M N K flags block_n [bias res]: one timed case (ncu target)
    test_gemm_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), (unsigned)atoi(argv[5]), argc > 7 ? atoi(argv[7]) : 1,
                   argc > 8 ? atoi(argv[8]) : 1, atoi(argv[6]), true);
    return g_fail ? 1 : 0;
  }
  if (!strcmp(what, "conv1") && argc >= 8) {  // selftest conv1 N H W C K block_n [mode: 0 temb (default), 1 residual, 2 plain]
    const int mode = argc >= 9 ? atoi(argv[8]) : 0;
    test_conv_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), mode == 0, mode == 1,
                   atoi(argv[7]), true);
    return g_fail ? 1 : 0;
  }
  if (!strcmp(what, "occ")) {  // registers / occupancy of the attention kernels
    st_debug_attention_occupancy();
    return 0;
  }
  if (!strcmp(what, "attn1") && argc >= 6) {  // selftest attn1 B H Tq Tk
    unsigned long long* tr;
    CK(cudaMalloc(&tr, 64 * 8));
    CK(cudaMemset(tr, 0, 64 * 8));
    st_debug_set_attention_trace(tr);
    test_attn_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), true, false);
    st_debug_set_attention_trace(nullptr);
    unsigned long long h[64];
    CK(cudaMemcpy(h, tr, sizeof h, cudaMemcpyDeviceToHost));
    // slot meanings for the pipelined kernel (Tk > 128); the two-CTA kernel fills 1..9 with its own phases
    const char* nm[14] = {"cta start", "exp warp, block 2: before m_ready wait", "m_ready seen, first S chunk requested",
                          "reference read", "-", "-", "exp stream done (64 columns)", "p_full arrived",
                          "mma: p_full(2) seen", "mma: P(2).V(2) issued", "exp warp, last block: before wait",
                          "cta exit", "max warp: m(2) published", "max warp: m(3) published"};
    if (getenv("ST_ATTN_IMPL") && !strcmp(getenv("ST_ATTN_IMPL"), "resident")) {
      // resident kernel, block 4 of CTA (0,0): 1 softmax warp before the s_full wait, 2 S seen, 3 S in registers, 4 reference
      // known, 5 exponentials + P stores issued, 6 stores complete, 7 p_full arrived; issuer: 8 before the p_full wait,
      // 9 P seen, 10 P.V + next S issued; 12 softmax warp sees S(5); 11 CTA exit
      for (int i = 1; i < 13; ++i)
        if (h[i]) printf("  resident slot %2d  %8lld  (+%lld since slot 1)\n", i, (long long)(h[i] - h[0]), (long long)(h[i] - h[1]));
    } else
    for (int i = 1; i < 14; ++i)
      if (h[i]) printf("  %-42s %8lld\n", nm[i], (long long)(h[i] - h[0]));
    if (h[16]) {
      printf("  block 16 / 17 start per exp warp (2..9), m(16) / m(17) published per max warp (10..13):\n   ");
      for (int i = 0; i < 12; ++i) printf(" %lld/%lld", (long long)(h[16 + i] - h[0]), (long long)(h[32 + i] - h[0]));
      printf("\n");
      const int order[9] = {48, 49, 50, 51, 57, 54, 55, 56, 52};
      const char* dn[9] = {"block 16 start (reference known)", "first chunk in registers", "chunk 0: 32 ex2 done",
                           "chunk 0: sums + packs done", "chunk 0: tcgen05.st issued", "chunk 1 in registers",
                           "chunk 1: 32 ex2 done", "chunk 1 stored", "P(16) published"};
      for (int i = 0; i < 9; ++i) printf("    warp 2  %-40s %8lld\n", dn[i], (long long)(h[order[i]] - h[48]));
      if (h[28])
        printf("    warp 2  tail: tcgen05.wait::st done %lld, fence done %lld, arrived %lld; block 17: loop top %lld, reference known %lld (prefetched: %s)\n",
               (long long)(h[28] - h[48]), (long long)(h[29] - h[48]), (long long)(h[52] - h[48]), (long long)(h[32] - h[48]),
               (long long)(h[31] - h[48]), h[15] == 1 ? "yes" : "no");
      printf("    warp 2  block 16 loop top -> reference known: %lld cycles\n", (long long)(h[48] - h[16]));
    }
    {
      const int extra[9] = {58, 59, 60, 61, 62, 46, 47, 44, 45};
      const char* en[9] = {"S issuer: s_free(16) seen", "S issuer: S(18) issued + committed", "PV issuer: p_full(16) seen",
                           "PV issuer: v_full(16) seen", "PV issuer: P(16).V(16) issued + committed",
                           "max warp: s_full(18) seen", "max warp: p_free for m(18) seen", "TMA: k_empty for K(20) seen",
                           "TMA: v_empty for V(20) seen"};
      for (int i = 0; i < 9; ++i)
        if (h[extra[i]]) printf("  %-42s %8lld\n", en[i], (long long)(h[extra[i]] - h[0]));
    }
    test_attn_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), true, true);
    return g_fail ? 1 : 0;
  }
  if (!strcmp(what, "sanitize")) {
    // one small, untimed case per kernel family: the compute-sanitizer targets (tools/sanitize.sh; memcheck, racecheck
    // and synccheck multiply the run time by 10-100x, so the shapes are the smallest that exercise every pipeline:
    // multi-tile persistent GEMM with a ragged edge, every epilogue, multi-block attention sweeps, both norm kernels)
    const char* fam = argc > 2 ? argv[2] : "all";
    const bool every = !strcmp(fam, "all");
    if (every || !strcmp(fam, "gemm")) {
      test_gemm_case(300, 328, 192, ST_W_STATIC, true, true, 192, false);
      test_gemm_case(256, 512, 128, ST_EPI_GEGLU | ST_W_STATIC, true, false, 128, false);
      test_gemm_case(384, 320, 320, ST_EPI_SILU, true, false, 0, false);
      test_gemm_case(130, 72, 64, ST_EPI_SILU, true, true, 64, false);
    }
    if (every || !strcmp(fam, "conv")) {
      test_conv_case(2, 16, 16, 64, 64, true, false, 0, false);
      test_conv_case(1, 8, 8, 128, 192, false, true, 0, false);
    }
    if (every || !strcmp(fam, "attn")) {
      test_attn_case(1, 2, 384, 384, true, false);   // pipelined kernel, 3 K/V blocks
      test_attn_case(1, 2, 200, 77, false, false);   // one-block kernel, ragged Tq and Tk
    }
    if (every || !strcmp(fam, "norm")) {
      test_groupnorm_case(2, 256, 320, 32, true, 0.f, false);
      test_groupnorm_case(1, 100, 2560, 32, false, 0.f, false);
      test_layernorm_case(77, 640, false);
      test_layernorm_case(300, 1280, false);
    }
    if (every || !strcmp(fam, "misc")) test_misc();
    printf("sanitize %s: %d failure(s), %llu kernel launches\n", fam, g_fail, st_launch_count());
    return g_fail ? 1 : 0;
  }
  const bool all = !strcmp(what, "all");
  if (all || !strcmp(what, "gemm")) test_gemm();
  if (all || !strcmp(what, "conv")) test_conv();
  if (all || !strcmp(what, "attn")) test_attn();
  if (all || !strcmp(what, "norm")) test_norm();
  if (all || !strcmp(what, "misc")) test_misc();
  printf("%s: %d failure(s), %llu kernel launches\n", what, g_fail, st_launch_count());
  return g_fail ? 1 : 0;
}
