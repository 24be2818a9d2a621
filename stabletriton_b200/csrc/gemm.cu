// C-ABI launchers for the tcgen05 GEMM / implicit-GEMM conv kernel (gemm_tc.cuh).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <stdlib.h>

namespace st {

// Tile shape from the measured cost model (profiles/r01_gemm_trace.txt, profiles/r02_gemm_pairs.txt, tools/mmabench*.cu).
// Per 64-wide k-block a CTA needs max(tensor pipe, operand fill) cycles: the tensor pipe takes 4 x BLOCK_N/2 cycles
// (a 128 x N x 16 -- or, for a CTA pair, 256 x N x 16 -- tcgen05.mma runs at its N/2 floor with the warp-convergent
// issue loop); the fill moves (128 + BLOCK_N) x 128 bytes per CTA (pair: 128 + BLOCK_N/2 rows, each CTA stages half
// of the weight tile) at min(88 B/clk into one SM, ~9.6 KB/clk out of L2 for all CTAs together).  At full occupancy
// the one-CTA 128 x 256 loop is L2-fill bound (664-760 cycles against 512), which CTA pairs (cta_group::2) remove:
// 8192^3 1351 -> 1465 TFLOP/s, GEGLU 2048 x 10240 x 1280 1242 -> 1331.  A pair launch costs ~1.5 k cycles more set-up
// (cluster scheduling, pair TMEM allocation, two cluster barriers), so short single-wave GEMMs stay unpaired.
// ST_GEMM_CLUSTER=0 / 1 forces pairs off / on wherever they are possible; default: the model decides per shape.
static int cluster_policy() {
  static const int v = [] {
    const char* e = getenv("ST_GEMM_CLUSTER");
    if (!e || !e[0]) return 2;
    return e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2);
  }();
  return v;
}

constexpr int kGemmDefaultSpin = 0;
static int drain_all_policy() {
  static const int v = [] {
    const char* e = getenv("ST_GEMM_DRAIN");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v;
}

static int spin_wait_policy() {
  static const int v = [] {
    const char* e = getenv("ST_GEMM_SPIN");
    return e ? atoi(e) & 3 : kGemmDefaultSpin;
  }();
  return v;
}

static bool pair_possible(int M, int block_n, bool geglu) {
  const int mb = (M + kGemmBlockM - 1) / kGemmBlockM;
  if (mb % 2 != 0) return false;  // vertically adjacent tiles pair up
  return geglu ? block_n == 256 : (block_n == 256 || block_n == 192 || block_n == 160);
}

struct TileChoice {
  int block_n;
  bool pair;
};

// Cycle model of one launch, calibrated on B200 (profiles/r02_gemm_tiles.txt):
//   k-block, one CTA per tile : max(tensor pipe 2*bn, fill/92 B/clk into one SM, fill * CTAs / 11.2 KB/clk out of L2)
//        (128x256 @148 CTAs 627-664 measured / 634 model, 128x256 @80 520-554 / 521, 128x192 @112 406-470 / 435)
//   k-block, CTA pair         : max(tensor pipe 2*bn, 2850 / stages) -- a pair stages half the weight tile per CTA, so
//        the fill never binds, but one trip round its operand ring (MMA commit -> both producers -> TMA -> leader's
//        barrier) takes ~2850 cycles: 256x256 (6 stages) 536 measured / 512 model, 256x160 and 256x192 (7) 407 / 407
//   per launch                : set-up until the first operands have landed 2.9 k (pair 4.4 k: cluster scheduling, pair
//        TMEM allocation, two cluster barriers), the last tile's epilogue ~0.9 k + 1.1 k per 64-column group (GEGLU 2.5 k),
//        and per further tile max(main loop, epilogue): short-K tiles are epilogue-bound.
static double tile_cost(int M, int n_cols, bool geglu, int K, int bn, bool pair, int sms) {
  const int mb = (M + kGemmBlockM - 1) / kGemmBlockM;
  const int out_cols = geglu ? bn / 2 : bn;
  const int nb = (n_cols + out_cols - 1) / out_cols;
  const long tiles = (long)mb * nb;
  const long slots = pair ? (sms / 2) * 2 : sms;
  const long waves = (tiles + slots - 1) / slots;
  const double ctas = tiles < slots ? (double)tiles : (double)slots;
  double kb;
  if (pair) {
    const int stages = bn == 256 ? 6 : 7;
    kb = 2850.0 / stages;
  } else {
    const double bytes = (128.0 + bn) * 128.0;
    const double fill_sm = bytes / 92.0, fill_l2 = bytes * ctas / 11200.0;
    kb = fill_sm > fill_l2 ? fill_sm : fill_l2;
  }
  if (kb < 2.0 * bn) kb = 2.0 * bn;
  const double main_loop = kb * (K / 64);
  const double groups = (out_cols + 63) / 64;
  const double epilogue = 900.0 + (geglu ? 2500.0 : 1100.0) * groups;
  const double steady = main_loop > epilogue ? main_loop : epilogue;
  return (pair ? 4400.0 : 2900.0) + main_loop + (waves - 1) * steady + epilogue;
}

static TileChoice choose_tile(int M, int n_cols, bool geglu, int K, bool allow_pair) {
  const int sms = device_sm_count();
  const int cands_plain[5] = {64, 128, 160, 192, 256};
  const int cands_geglu[2] = {128, 256};
  const int* cands = geglu ? cands_geglu : cands_plain;
  const int ncand = geglu ? 2 : 5;
  const int policy = allow_pair ? cluster_policy() : 0;
  TileChoice best{256, false};
  double best_cost = 1e30;
  for (int i = 0; i < ncand; ++i) {
    const int bn = cands[i];
    for (int pair = 0; pair < 2; ++pair) {
      if (pair && (policy == 0 || !pair_possible(M, bn, geglu))) continue;
      if (!pair && policy == 1 && pair_possible(M, bn, geglu)) continue;
      const double cost = tile_cost(M, n_cols, geglu, K, bn, pair != 0, sms);
      if (cost < best_cost * 0.97) {  // ties go to the narrower / un-paired tile
        best_cost = cost;
        best = TileChoice{bn, pair != 0};
      }
    }
  }
  return best;
}

template <int BLOCK_N, int STAGES, bool kConvA, bool kGeglu, bool kStreamK = false, bool kCluster = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tdt,
                       const CUtensorMap& tr, const CUtensorMap& trt, const GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N, STAGES, kCluster>;
  auto kernel = gemm_bf16_tc_kernel<BLOCK_N, STAGES, kConvA, kGeglu, kStreamK, kCluster>;
  static PerDeviceOnce configured;  // per instantiation AND per device (function attributes live in the context)
  const int dev = current_device();
  if (dev < 0) {
    set_error("gemm: device ordinal outside [0, %d)", kMaxDevices);
    return ST_ERR_INVALID_ARGUMENT;
  }
  if (!configured.done(dev)) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%d bytes) failed: %s", S::kTotal, cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    configured.mark(dev);
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const long units = static_cast<long>(tiles) * (p.K / kGemmBlockK);
  const int sms = device_sm_count();
  // stream-K: every CTA must own at least one (tile, k-block) unit -- an owner waits for its successors
  int grid = p.stream_k ? static_cast<int>(units < sms ? units : sms) : (tiles < sms ? tiles : sms);
  if (kCluster) {  // one cluster per pair of vertically adjacent tiles, at most sms / 2 clusters
    const int pair_tiles = tiles / 2;
    grid = 2 * (pair_tiles < sms / 2 ? pair_tiles : sms / 2);
  }
  launch_kernel_cluster(kernel, dim3(grid), dim3(kGemmThreads), S::kTotal, stream, kCluster ? 2 : 1, ta, tb, td, tdt, tr, trt, p);
  ST_CHECK_LAUNCH("gemm_bf16_tc_kernel");
  return ST_OK;
}

template <bool kConvA>
static int dispatch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tdt,
                         const CUtensorMap& tr, const CUtensorMap& trt, const GemmParams& p, int block_n, bool geglu,
                         cudaStream_t stream) {
  if (p.cluster) {
    if (geglu && block_n == 256) return launch_gemm<256, 6, kConvA, true, false, true>(ta, tb, td, tdt, tr, trt, p, stream);
    if (!geglu && block_n == 256) return launch_gemm<256, 6, kConvA, false, false, true>(ta, tb, td, tdt, tr, trt, p, stream);
    if (!geglu && block_n == 192) return launch_gemm<192, 7, kConvA, false, false, true>(ta, tb, td, tdt, tr, trt, p, stream);
    if (!geglu && block_n == 160) return launch_gemm<160, 7, kConvA, false, false, true>(ta, tb, td, tdt, tr, trt, p, stream);
    set_error("gemm: no cluster instantiation for block_n %d", block_n);
    return ST_ERR_INVALID_ARGUMENT;
  }
  if (geglu) {
    switch (block_n) {
      case 256: return launch_gemm<256, 4, kConvA, true>(ta, tb, td, tdt, tr, trt, p, stream);
      case 128: return launch_gemm<128, 6, kConvA, true>(ta, tb, td, tdt, tr, trt, p, stream);
    }
  } else {
    switch (block_n) {
      case 256:
        if (p.stream_k) return launch_gemm<256, 4, kConvA, false, true>(ta, tb, td, tdt, tr, trt, p, stream);
        return launch_gemm<256, 4, kConvA, false>(ta, tb, td, tdt, tr, trt, p, stream);
      case 192: return launch_gemm<192, 5, kConvA, false>(ta, tb, td, tdt, tr, trt, p, stream);
      case 160: return launch_gemm<160, 5, kConvA, false>(ta, tb, td, tdt, tr, trt, p, stream);
      case 128: return launch_gemm<128, 6, kConvA, false>(ta, tb, td, tdt, tr, trt, p, stream);
      case 64: return launch_gemm<64, 8, kConvA, false>(ta, tb, td, tdt, tr, trt, p, stream);
    }
  }
  set_error("gemm: unsupported block_n %d (0, 64, 128, 160, 192 or 256; GEGLU: 128 or 256)", block_n);
  return ST_ERR_INVALID_ARGUMENT;
}

static unsigned long long* g_gemm_trace = nullptr;  // debug only, see st_debug_set_gemm_trace

// stream-K state: caller-provided fp32 scratch (st_set_workspace) and self-resetting arrival flags
static void* g_ws_ptr[kMaxDevices] = {nullptr};
static size_t g_ws_bytes[kMaxDevices] = {0};
static unsigned* g_sk_flag_ptr[kMaxDevices] = {nullptr};  // address of g_sk_flags in each device's context
__device__ unsigned g_sk_flags[256 * 32];  // one flag per 128-byte line

// Decide whether a GEMM with `tiles` 128x256 tiles and `nkb` k-blocks should run stream-K: only when the
// tiles do not fill the machine and the K loop is long enough to pay for the fix-up (cycle model as above).
static bool want_stream_k(long tiles256, int nkb, long tiles_chosen, float** ws, unsigned** flags) {
  // Experimental, opt-in (ST_ENABLE_STREAMK=1): measured on B200 (profiles/r01_gemm_trace.txt) the fp32
  // fix-up through L2 costs more than the shorter K loop saves (2048x1280x5120: 55 us vs 33 us plain), and at
  // 148 active CTAs the 128x256 tiles become L2-fill bound (~730 instead of 554 cycles per k-block).
  static const bool disabled = [] {
    const char* e = getenv("ST_ENABLE_STREAMK");
    return !(e && e[0] && e[0] != '0');
  }();
  const int sms = device_sm_count();
  const int dev = current_device();
  if (disabled || dev < 0 || !g_ws_ptr[dev] || tiles256 >= sms || tiles_chosen > sms || nkb < 16) return false;
  if (g_ws_bytes[dev] < static_cast<size_t>(sms) * kGemmBlockM * 256 * sizeof(float)) return false;
  const double cost_plain = 554.0 * nkb + 6000.0;
  const double cost_sk = 554.0 * ((tiles256 * nkb + sms - 1) / sms) + 10000.0;
  if (cost_sk > 0.85 * cost_plain) return false;
  if (!g_sk_flag_ptr[dev]) cudaGetSymbolAddress(reinterpret_cast<void**>(&g_sk_flag_ptr[dev]), g_sk_flags);
  *ws = static_cast<float*>(g_ws_ptr[dev]);
  *flags = g_sk_flag_ptr[dev];
  return g_sk_flag_ptr[dev] != nullptr;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ST_GEMM_RES_TMA=0: residual through per-thread row loads everywhere (A/B runs; the default is the TMA path)
static bool res_tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("ST_GEMM_RES_TMA");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace st

extern "C" {

int st_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* D, int ldd, int M, int N, int K,
                 const void* bias, const void* residual, int ldr, unsigned flags, int block_n, void* gn_partial,
                 st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(A && W && D, "gemm: null pointer");
  ST_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: M, N, K must be positive (got %d, %d, %d)", M, N, K);
  ST_CHECK_ARG(K % kGemmBlockK == 0, "gemm: K (%d) must be a multiple of 64", K);
  ST_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0, "gemm: row pitches must be multiples of 8 elements");
  ST_CHECK_ARG(lda >= K && ldw >= K, "gemm: lda/ldw smaller than K");
  ST_CHECK_ARG(aligned16(A) && aligned16(W) && aligned16(D), "gemm: pointers must be 16-byte aligned");
  const bool geglu = (flags & ST_EPI_GEGLU) != 0;
  const bool out_f32 = (flags & ST_EPI_F32OUT) != 0;
  ST_CHECK_ARG(!out_f32 || (!geglu && !gn_partial), "gemm: the fp32 output mode excludes GEGLU and GroupNorm partials");
  ST_CHECK_ARG(!geglu || N % 2 == 0, "gemm: GEGLU needs an even N");
  const int n_out = geglu ? N / 2 : N;
  ST_CHECK_ARG(n_out % 8 == 0, "gemm: output width (%d) must be a multiple of 8", n_out);
  ST_CHECK_ARG(ldd >= n_out, "gemm: ldd smaller than the output width");
  ST_CHECK_ARG(!bias || aligned16(bias), "gemm: bias must be 16-byte aligned");
  ST_CHECK_ARG(!residual || (aligned16(residual) && ldr % 8 == 0 && ldr >= n_out), "gemm: bad residual pitch/alignment");
  ST_CHECK_ARG(!(geglu && (flags & ST_EPI_SILU)), "gemm: GEGLU and SiLU epilogues are exclusive");
  ST_CHECK_ARG(!gn_partial || (!geglu && M % kGemmBlockM == 0 && aligned16(gn_partial)),
               "gemm: GroupNorm partials need M %% 128 == 0 (got %d), no GEGLU and a 16-byte aligned buffer", M);

  float* sk_ws = nullptr;
  unsigned* sk_flags = nullptr;
  bool stream_k = false;
  bool pair = false;
  if (block_n == 0) {
    const TileChoice tc = choose_tile(M, n_out, geglu, K, /*allow_pair=*/true);
    block_n = tc.block_n;
    pair = tc.pair;
    const long mb = (M + kGemmBlockM - 1) / kGemmBlockM;
    if (!geglu && !gn_partial && !out_f32 && !pair && want_stream_k(mb * ((n_out + 255) / 256), K / kGemmBlockK, mb * ((n_out + block_n - 1) / block_n),
                                &sk_ws, &sk_flags)) {
      stream_k = true;
      block_n = 256;
    }
  } else if (block_n < 0) {  // negative: force a CTA-pair launch with tile width -block_n (tests, tuning)
    block_n = -block_n;
    ST_CHECK_ARG(pair_possible(M, block_n, geglu), "gemm: a pair launch needs an even number of row "
                 "blocks and block_n 256 / 192 / 160 (GEGLU: 256); got M = %d, block_n = %d", M, block_n);
    pair = true;
  } else if (cluster_policy() == 1 && pair_possible(M, block_n, geglu)) {
    pair = true;
  }
  const int out_cols = geglu ? block_n / 2 : block_n;

  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.n_out = n_out;
  p.ldd = ldd;
  p.num_m_blocks = (M + kGemmBlockM - 1) / kGemmBlockM;
  p.num_n_blocks = (n_out + out_cols - 1) / out_cols;
  p.D = static_cast<__nv_bfloat16*>(D);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.rowbias = nullptr;
  p.rows_per_batch = 1;
  p.act_silu = (flags & ST_EPI_SILU) ? 1 : 0;
  p.trace = g_gemm_trace;
  p.stream_k = stream_k ? 1 : 0;
  p.w_static = (flags & ST_W_STATIC) ? 1 : 0;
  p.ws = sk_ws;
  p.flags = sk_flags;
  p.cluster = pair ? 1 : 0;
  p.drain_all = drain_all_policy();
  p.spin_wait = spin_wait_policy();
  p.gn_part = static_cast<float*>(gn_partial);
  p.out_f32 = out_f32 ? 1 : 0;
  const int ldd_map = out_f32 ? 2 * ldd : ldd;  // the (unused) store maps describe the fp32 rows as twice as many bf16

  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, M, K, lda, kGemmBlockM);
  if (rc != ST_OK) return rc;
  rc = make_tmap_2d(&tb, W, N, K, ldw, (geglu || p.cluster) ? block_n / 2 : block_n);  // one box per B load
  if (rc != ST_OK) return rc;
  CUtensorMap td, tdt;
  rc = make_tmap_2d(&td, D, M, n_out, ldd_map, kGemmBlockM);
  if (rc != ST_OK) return rc;
  rc = make_tmap_2d(&tdt, D, M, n_out, ldd_map, kGemmBlockM, 32, /*swizzle128=*/false);  // 32-column tail group of a 160-wide tile
  if (rc != ST_OK) return rc;
  CUtensorMap tr = td, trt = tdt;  // residual tile of a CTA's last output tile: same boxes as the store maps
  p.res_tma = (residual && !out_f32 && !geglu && !p.stream_k && res_tma_enabled()) ? 1 : 0;
  if (p.res_tma) {
    rc = make_tmap_2d(&tr, residual, M, n_out, ldr, kGemmBlockM);
    if (rc != ST_OK) return rc;
    rc = make_tmap_2d(&trt, residual, M, n_out, ldr, kGemmBlockM, 32, /*swizzle128=*/false);
    if (rc != ST_OK) return rc;
  }
  return dispatch_gemm<false>(ta, tb, td, tdt, tr, trt, p, block_n, geglu, static_cast<cudaStream_t>(stream));
}

int st_conv3x3_nhwc_bf16(const void* x, const void* w, const void* bias, void* y, int N, int H, int W, int C, int K,
                         const void* temb, int ld_temb, const void* residual, unsigned flags, int block_n,
                         void* gn_partial, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && w && y, "conv3x3: null pointer");
  ST_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && K > 0, "conv3x3: sizes must be positive");
  ST_CHECK_ARG(C % 64 == 0, "conv3x3: C (%d) must be a multiple of 64 (use st_conv3x3_direct_nhwc_bf16)", C);
  ST_CHECK_ARG(K % 8 == 0, "conv3x3: K (%d) must be a multiple of 8", K);
  // A tile = 128 output pixels = a Wt x Ht rectangle of one image, or Nt whole (small) images.
  int Wt, Ht, Nt;
  if (H * W >= kGemmBlockM) {
    ST_CHECK_ARG((H * W) % kGemmBlockM == 0, "conv3x3: H*W (%d) must be a multiple of 128 (or divide 128)", H * W);
    Wt = W < 128 ? W : 128;
    ST_CHECK_ARG(W % Wt == 0 && 128 % Wt == 0, "conv3x3: unsupported width %d", W);
    Ht = 128 / Wt;
    Nt = 1;
    ST_CHECK_ARG(H % Ht == 0, "conv3x3: unsupported height %d for width %d", H, W);
  } else {
    ST_CHECK_ARG(kGemmBlockM % (H * W) == 0, "conv3x3: H*W (%d) must divide 128 (or be a multiple of it)", H * W);
    Wt = W;
    Ht = H;
    Nt = kGemmBlockM / (H * W);
  }
  ST_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y), "conv3x3: pointers must be 16-byte aligned");
  ST_CHECK_ARG(!(flags & ST_EPI_GEGLU), "conv3x3: GEGLU epilogue not supported");
  ST_CHECK_ARG(!temb || (aligned16(temb) && ld_temb % 8 == 0), "conv3x3: bad temb pitch/alignment");

  const int M = N * H * W;
  ST_CHECK_ARG(!gn_partial || (M % kGemmBlockM == 0 && aligned16(gn_partial)),
               "conv3x3: GroupNorm partials need N*H*W %% 128 == 0 and a 16-byte aligned buffer");
  float* sk_ws = nullptr;
  unsigned* sk_flags = nullptr;
  bool stream_k = false;
  bool pair = false;
  if (block_n == 0) {
    static const bool conv_pairs = [] {  // ST_CONV_CLUSTER=0: no CTA pairs for the implicit-GEMM convolutions (A/B runs)
      const char* e = getenv("ST_CONV_CLUSTER");
      return !(e && e[0] == '0');
    }();
    // Measured inside the launch sequence of a step (profiles/r02_gemm_experiments.txt, 9): the pair kernel's k-block
    // costs a convolution ~590 cycles at 256 x 160 (4-D A loads: 26 KB per CTA) against ~525 for the one-CTA kernel, so
    // the GEMM-calibrated model over-rates pairs here; only the 256 x 256 pair tile of the 1280-filter levels pays.
    const TileChoice tc = choose_tile(M, K, false, 9 * C, /*allow_pair=*/conv_pairs && K >= 1280);
    block_n = tc.block_n;
    pair = tc.pair;
    const long mb = (M + kGemmBlockM - 1) / kGemmBlockM;
    if (!gn_partial && !pair && want_stream_k(mb * ((K + 255) / 256), 9 * C / kGemmBlockK, mb * ((K + block_n - 1) / block_n), &sk_ws,
                      &sk_flags)) {
      stream_k = true;
      block_n = 256;
    }
  } else if (block_n < 0) {
    block_n = -block_n;
    ST_CHECK_ARG(pair_possible(M, block_n, false), "conv3x3: a pair launch needs an even number of "
                 "row blocks and block_n 256 / 192 / 160; got N*H*W = %d, block_n = %d", M, block_n);
    pair = true;
  } else if (cluster_policy() == 1 && pair_possible(M, block_n, false)) {
    pair = true;
  }

  GemmParams p{};
  p.M = M;
  p.N = K;
  p.K = 9 * C;
  p.n_out = K;
  p.ldd = K;
  p.num_m_blocks = (M + kGemmBlockM - 1) / kGemmBlockM;
  p.num_n_blocks = (K + block_n - 1) / block_n;
  p.D = static_cast<__nv_bfloat16*>(y);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = K;
  p.rowbias = static_cast<const __nv_bfloat16*>(temb);
  p.ld_rowbias = ld_temb;
  p.rows_per_batch = H * W;
  p.act_silu = (flags & ST_EPI_SILU) ? 1 : 0;
  p.trace = g_gemm_trace;
  p.stream_k = stream_k ? 1 : 0;
  p.w_static = (flags & ST_W_STATIC) ? 1 : 0;
  p.ws = sk_ws;
  p.flags = sk_flags;
  p.cluster = pair ? 1 : 0;
  p.drain_all = drain_all_policy();
  p.spin_wait = spin_wait_policy();
  p.gn_part = static_cast<float*>(gn_partial);
  p.conv_H = H;
  p.conv_W = W;
  p.conv_C = C;
  p.conv_Wt = Wt;
  p.conv_Ht = Ht;

  CUtensorMap ta, tb;
  int rc = make_tmap_nhwc(&ta, x, N, H, W, C, Ht, Wt, Nt);
  if (rc != ST_OK) return rc;
  rc = make_tmap_2d(&tb, w, K, 9 * (uint64_t)C, 9 * (uint64_t)C, p.cluster ? block_n / 2 : block_n);
  if (rc != ST_OK) return rc;
  CUtensorMap td, tdt;
  rc = make_tmap_2d(&td, y, M, K, K, kGemmBlockM);
  if (rc != ST_OK) return rc;
  rc = make_tmap_2d(&tdt, y, M, K, K, kGemmBlockM, 32, /*swizzle128=*/false);
  if (rc != ST_OK) return rc;
  CUtensorMap tr = td, trt = tdt;
  p.res_tma = (residual && !p.stream_k && res_tma_enabled()) ? 1 : 0;
  if (p.res_tma) {
    rc = make_tmap_2d(&tr, residual, M, K, K, kGemmBlockM);
    if (rc != ST_OK) return rc;
    rc = make_tmap_2d(&trt, residual, M, K, K, kGemmBlockM, 32, /*swizzle128=*/false);
    if (rc != ST_OK) return rc;
  }
  return dispatch_gemm<true>(ta, tb, td, tdt, tr, trt, p, block_n, false, static_cast<cudaStream_t>(stream));
}

int st_set_workspace(void* ptr, size_t bytes) {
  const int dev = st::current_device();
  ST_CHECK_ARG(dev >= 0, "set_workspace: device ordinal outside [0, %d)", st::kMaxDevices);
  st::g_ws_ptr[dev] = ptr;
  st::g_ws_bytes[dev] = ptr ? bytes : 0;
  return ST_OK;
}

size_t st_workspace_bytes(void) {
  return static_cast<size_t>(st::device_sm_count()) * st::kGemmBlockM * 256 * sizeof(float);
}

// Debug hook: the launcher's tile choice for a shape -- block_n, or -block_n for a CTA-pair launch.
int st_debug_choose_tile(int M, int n_cols, int geglu, int K) {
  const st::TileChoice tc = st::choose_tile(M, n_cols, geglu != 0, K, true);
  return tc.pair ? -tc.block_n : tc.block_n;
}

// Debug hook (not part of the product surface): when set, every GEMM/conv CTA writes 8 clock64 stamps
// (start, setup done, first operands landed, MMA issue done, accumulator ready, epilogue done, exit) to buf.
void st_debug_set_gemm_trace(void* buf) { st::g_gemm_trace = static_cast<unsigned long long*>(buf); }

}  // extern "C"
