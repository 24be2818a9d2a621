// Flash attention forward for head_dim 64 on sm_100a: TMA-fed tcgen05 tiles, S, P and O in TMEM, softmax in
// fp32 registers.  Two kernels behind st_attention_bf16:
//
//   attn_fwd_kernel            Tk <= 128 (the 77-token cross attention: one K/V block, nothing to pipeline).
//   attn_fwd_pipelined_kernel  Tk  > 128 (self attention): one CTA per SM, double-buffered S / P, O accumulated in
//                              TMEM, max / exp warp specialisation, lazy rescale -- see the comment above it.
//
// attn_fwd_kernel: grid = (ceil(Tq / 128), B * H); 320 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM
//   owner, warps 2..9 softmax / accumulate / store (two warps per TMEM lane quadrant, each owning half of every
//   row).  Two CTAs are resident per SM (<= 84 KB smem, 256 TMEM columns, <= 85 registers).
//
//   S = Q K^T        : tcgen05.mma SS (Q, K 128B-swizzled K-major tiles from TMA), 128 fp32 columns
//   P = 2^(S*c - m)  : written back to TMEM as packed bf16 (64 columns) with tcgen05.st
//   O_j = P V        : tcgen05.mma TS -- A operand straight from TMEM, V as an MN-major smem operand (32 cycles per
//                      128 x 64 x 16 instruction against 48 for the SS form; tools/mmabench.cu)
//   O += O_j         : folded into fp32 registers one block late, rescaled by 2^(m_old - m_new)
//
// Semantics = the *pattern* of fuse_attention (reference: optimizers/replace_attention.py:76-86;
// fp32 restatement ref_attention :109-124): per head softmax(Q K^T * scale) V, no mask, separate
// Tq / Tk with a masked K tail (Tk = 77 cross-attention).  Replaces
// kernels/attention_fa2.py:17-140.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace st {

// One mbarrier arrival per WARP: every lane has finished (and fenced) its own tensor-memory / shared-memory accesses,
// __syncwarp orders them before lane 0's releasing arrive.  Per-thread arrivals looked harmless but are serialised
// read-modify-writes on one shared-memory word -- 640 of them per K/V block in the pipelined kernel (s_free, p_full,
// m_ready), which turned out to BE its ~1.45 k-cycle block period: switching whole stages of the pipeline off left
// the time unchanged (profiles/r02_attention_experiments.txt, section 3).
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

constexpr int kAttnBlockQ = 128;
constexpr int kAttnBlockKV = 128;
constexpr int kAttnD = 64;
constexpr int kAttnStages = 2;
constexpr int kAttnThreads = 320;  // TMA warp + MMA warp + 8 softmax warps
constexpr int kAttnTileBytes = 128 * 64 * 2;  // any [128 x 64] bf16 tile
constexpr int kAttnSmemBytes = kAttnTileBytes * (1 + 2 * kAttnStages) + 256 + 2048 + 1024;  // + barriers + row exchange
constexpr int kAttnTmemCols = 256;            // S: [0,128)  P: [128,192)  O_j: [192,256)

struct AttnParams {
  __nv_bfloat16* O;
  long long o_sb, o_sh, o_st;
  int H, Tq, Tk;
  float scale_log2;
  unsigned long long* trace;  // debug: 16 clock64 stamps for CTA (0,0), or nullptr
  int ablate;                 // debug (ST_ATTN_ABLATE, timing only -- results are wrong): see attn_fwd_pipelined_kernel<.., kAblate>
  unsigned wait_hint_ns, wait_sleep_ns;  // TMA / issuer / max warps of the pipelined kernel: mbar_wait_relaxed parameters
};

// Registers are allocated per group of 4 warps: the 10 warps of this CTA cost as much as 12, so two resident
// CTAs need <= 85 registers per thread -- hence the bound of 384 threads although 320 are launched.
template <bool kTrace>
__global__ void __launch_bounds__(384, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTileBytes;
  uint8_t* sV = sK + kAttnStages * kAttnTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnStages * kAttnTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* v_full = k_full + kAttnStages;
  uint64_t* kv_empty = v_full + kAttnStages;
  uint64_t* s_full = kv_empty + kAttnStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* s_xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y - b * p.H;
  const int nkv = (p.Tk + kAttnBlockKV - 1) / kAttnBlockKV;
#define AT_TRACE(slot)                                                                        \
  do {                                                                                        \
    if (kTrace && p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[slot] = clock64();   \
  } while (0)
  if (threadIdx.x == 0) AT_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);   // one arrival per softmax warp (warp_arrive)
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kAttnTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // prologue above touched only shared / tensor memory
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_P = tmem_base + 128;
  const uint32_t tmem_O = tmem_base + 192;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, kAttnTileBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], kAttnTileBytes);
        tma_load_4d(sK + st * kAttnTileBytes, &tmap_k, &k_full[st], 0, j * kAttnBlockKV, h, b);
        mbar_expect_tx(&v_full[st], kAttnTileBytes);
        tma_load_4d(sV + st * kAttnTileBytes, &tmap_v, &v_full[st], 0, j * kAttnBlockKV, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // whole warp in the loop, one elected lane issues (umma_bf16_ss_elect, ptx.cuh)
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kAttnBlockKV, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);        // P (TMEM)    x V (MN-major)
      const uint64_t desc_q = umma_smem_desc_sw128(smem_u32(sQ), 0, 1024);
      const uint64_t desc_k0 = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
      const uint64_t desc_v0 = umma_smem_desc_sw128(smem_u32(sV), 8192, 1024);
      constexpr uint32_t kTileStep = kAttnTileBytes >> 4;  // descriptor address units (16 bytes)
      auto issue_s = [&](int j) {
        const int st = j % kAttnStages;
        mbar_wait(&k_full[st], (j / kAttnStages) & 1);
        tc_fence_after();
        const uint64_t dk = desc_k0 + static_cast<uint64_t>(st * kTileStep);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_bf16_ss_elect(tmem_S, desc_q + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit_elect(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAttnStages;
        mbar_wait(p_full, j & 1);  // softmax(j) has consumed S(j) and published P(j)
        tc_fence_after();
        if (j == 2 && lane == 0) AT_TRACE(8);
        if (j + 1 < nkv) issue_s(j + 1);
        mbar_wait(&v_full[st], (j / kAttnStages) & 1);
        if (j > 0) mbar_wait(o_empty, (j - 1) & 1);  // O_{j-1} has been folded into registers
        tc_fence_after();
        const uint64_t dv = desc_v0 + static_cast<uint64_t>(st * kTileStep);
#pragma unroll
        for (int kk = 0; kk < kAttnBlockKV / 16; ++kk)  // 16 V rows (2 KB) per instruction
          umma_bf16_ts_elect(tmem_O, tmem_P + kk * 8, dv + 128 * kk, idesc_o, kk != 0);
        umma_commit_elect(&kv_empty[st]);
        umma_commit_elect(o_full);
        if (j == 2 && lane == 0) AT_TRACE(9);
      }
    }
  } else {
    // ===================================== softmax / accumulate ===============================
    // 8 warps: the two warps that share a TMEM lane quadrant split every row -- warp `half` owns score
    // columns [64*half, 64*half+64) of the block and output columns [32*half, 32*half+32) -- and exchange
    // the row maximum through shared memory.  Halving the per-thread serial chain (max pass -> fold ->
    // exp pass) and doubling the warps per scheduler is what hides the MUFU / TMEM latencies.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t pair_bar = 1 + quad;  // named barrier of this quadrant's two warps
    float m = -INFINITY, l = 0.f;        // m: running max of the *scaled* scores (log2 domain); l: my half of the row sum
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int valid = p.Tk - j * kAttnBlockKV - half * 64;  // my columns >= valid are padding
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(1);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(2);
      // pass 1: row maximum of the raw scores (scale > 0, so scaling commutes with max)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + half * 64 + c, v);
        tmem_ld_wait();
        if (c + 32 > valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= valid) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(v[i + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
        }
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      float* xrow = s_xch + (j & 1) * 256 + row;
      xrow[half * 128] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, xrow[(half ^ 1) * 128]);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float alpha = ex2_approx(m - m_new);  // m = -inf on the first block -> 0
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(3);

      // fold in my half of the P.V product of the previous block (computed relative to the old max); this
      // also proves that P(j-1) has been consumed, so the P buffer may be overwritten below
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (j == 2 && warp == 2 && lane == 0) AT_TRACE(4);
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_O + lane_off + half * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive(o_empty, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (acc[i] + __uint_as_float(o[i])) * alpha;
      }
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(5);

      // pass 2: p = 2^(s*c - max), partial row sum, packed bf16 P into TMEM
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + half * 64 + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float e0 = ex2_approx(fmaf(__uint_as_float(v[i + 0]), p.scale_log2, -m_new));
          float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
          float e2 = ex2_approx(fmaf(__uint_as_float(v[i + 2]), p.scale_log2, -m_new));
          float e3 = ex2_approx(fmaf(__uint_as_float(v[i + 3]), p.scale_log2, -m_new));
          if (c + 32 > valid) {
            e0 = (c + i + 0 < valid) ? e0 : 0.f;
            e1 = (c + i + 1 < valid) ? e1 : 0.f;
            e2 = (c + i + 2 < valid) ? e2 : 0.f;
            e3 = (c + i + 3 < valid) ? e3 : 0.f;
          }
          rs0 += e0;
          rs1 += e1;
          rs2 += e2;
          rs3 += e3;
          pk[(i >> 1) + 0] = pack_bf16x2(e0, e1);
          pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
        }
        tmem_st_32x32b_x16(tmem_P + lane_off + half * 32 + (c >> 1), pk);
      }
      l = fmaf(l, alpha, (rs0 + rs1) + (rs2 + rs3));
      m = m_new;
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(6);
      tmem_st_wait();
      tc_fence_before();
      warp_arrive(p_full, lane);
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(7);
    }
    {
      mbar_wait(o_full, (nkv - 1) & 1);
      tc_fence_after();
      // total row sum = my half + the partner's
      float* xrow = s_xch + (nkv & 1) * 256 + row;
      xrow[half * 128] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float inv = 1.f / (l + xrow[(half ^ 1) * 128]);
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_off + half * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = (acc[i] + __uint_as_float(o[i])) * inv;
    }
    if (q0 + row < p.Tq) {
      __nv_bfloat16* orow =
          p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        o.x = pack_bf16x2(acc[i + 0], acc[i + 1]);
        o.y = pack_bf16x2(acc[i + 2], acc[i + 3]);
        o.z = pack_bf16x2(acc[i + 4], acc[i + 5]);
        o.w = pack_bf16x2(acc[i + 6], acc[i + 7]);
        *reinterpret_cast<uint4*>(orow + i) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<kAttnTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Short-context variant for Tk <= 80 (the 77-token cross attention): one K/V block of 80 rows, everything in one
// pass, small enough for FOUR resident CTAs per SM -- 37 KB of shared memory, 128 TMEM columns, 160 threads.  The
// launches of the step are latency chains (TMA -> S -> softmax -> P.V -> store, ~5 us) over 320 CTAs (Tq = 1024) or
// 640 (Tq = 4096): with the two-CTA-per-SM kernel above (296 slots) the 24-CTA second wave doubled the chain.
//
//   warp 0      TMA loads (Q + K on one barrier, V on its own), then the two MMA batches
//   warps 1-4   one query row per thread: the whole 80-column score row in registers (one TMEM round trip), exact
//               max, exponentials, bf16 P written over the dead S columns, then O / l -> HBM
//
//   TMEM   S [0,80) fp32;  P [0,40) packed bf16, aliasing S (a thread overwrites only columns of its own lane, after
//          it has read them);  O [64,128) (S is dead by the time P.V is issued)
constexpr int kShortKV = 80;
constexpr int kShortThreads = 160;
constexpr int kShortKVBytes = kShortKV * kAttnD * 2;
constexpr int kShortSmemBytes = kAttnTileBytes + 2 * kShortKVBytes + 128 + 1024;
constexpr int kShortTmemCols = 128;

__global__ void __launch_bounds__(kShortThreads, 3)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTileBytes;
  uint8_t* sV = sK + kShortKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kShortKVBytes);
  uint64_t* qk_full = bars;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y - b * p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);  // one arrival per softmax warp
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<kShortTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // prologue above touched only shared / tensor memory
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_P = tmem_base;
  const uint32_t tmem_O = tmem_base + 64;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qk_full, kAttnTileBytes + kShortKVBytes);
      tma_load_4d(sQ, &tmap_q, qk_full, 0, q0, h, b);
      tma_load_4d(sK, &tmap_k, qk_full, 0, 0, h, b);
      mbar_expect_tx(v_full, kShortKVBytes);
      tma_load_4d(sV, &tmap_v, v_full, 0, 0, h, b);
    }
    __syncwarp();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kShortKV, 0, 0);  // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);    // P (TMEM)    x V (MN-major)
    const uint64_t desc_q = umma_smem_desc_sw128(smem_u32(sQ), 0, 1024);
    const uint64_t desc_k = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
    const uint64_t desc_v = umma_smem_desc_sw128(smem_u32(sV), 8192, 1024);
    mbar_wait(qk_full, 0);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < kAttnD / 16; ++k) umma_bf16_ss_elect(tmem_S, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k != 0);
    umma_commit_elect(s_full);
    mbar_wait(v_full, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
#pragma unroll
    for (int kk = 0; kk < kShortKV / 16; ++kk)  // 16 V rows (2 KB) per instruction
      umma_bf16_ts_elect(tmem_O, tmem_P + kk * 8, desc_v + 128 * kk, idesc_o, kk != 0);
    umma_commit_elect(o_full);
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int valid = p.Tk;     // score columns >= valid are padding (K rows zero-filled by TMA)
    mbar_wait(s_full, 0);
    tc_fence_after();
    uint32_t s0[32], s1[32], s2[16];
    tmem_ld_32x32b_x32(tmem_S + lane_off, s0);
    tmem_ld_32x32b_x32(tmem_S + lane_off + 32, s1);
    tmem_ld_32x32b_x16(tmem_S + lane_off + 64, s2);
    tmem_ld_wait();
    if (valid < 64) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= valid) s0[i] = 0xff800000u;  // -inf
        if (32 + i >= valid) s1[i] = 0xff800000u;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (64 + i >= valid) s2[i] = 0xff800000u;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 2) mx = max3f(mx, __uint_as_float(s0[i]), __uint_as_float(s0[i + 1]));
#pragma unroll
    for (int i = 0; i < 32; i += 2) mx = max3f(mx, __uint_as_float(s1[i]), __uint_as_float(s1[i + 1]));
#pragma unroll
    for (int i = 0; i < 16; i += 2) mx = max3f(mx, __uint_as_float(s2[i]), __uint_as_float(s2[i + 1]));
    const float neg_m = -mx * p.scale_log2;  // scale > 0 commutes with max; column 0 is always valid, so mx is finite
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    auto exp_pack = [&](const uint32_t* v, uint32_t* pk, int n) {
#pragma unroll
      for (int i = 0; i < n; i += 4) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(v[i + 0]), p.scale_log2, neg_m));
        const float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, neg_m));
        const float e2 = ex2_approx(fmaf(__uint_as_float(v[i + 2]), p.scale_log2, neg_m));
        const float e3 = ex2_approx(fmaf(__uint_as_float(v[i + 3]), p.scale_log2, neg_m));
        l0 += e0;
        l1 += e1;
        l2 += e2;
        l3 += e3;
        pk[(i >> 1) + 0] = pack_bf16x2(e0, e1);
        pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
      }
    };
    {
      uint32_t pk[16];
      exp_pack(s0, pk, 32);
      tmem_st_32x32b_x16(tmem_P + lane_off, pk);
    }
    {
      uint32_t pk[16];
      exp_pack(s1, pk, 32);
      tmem_st_32x32b_x16(tmem_P + lane_off + 16, pk);
    }
    {
      uint32_t pk[8];
      exp_pack(s2, pk, 16);
      tmem_st_32x32b_x8(tmem_P + lane_off + 32, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    warp_arrive(p_full, lane);
    const float inv = 1.f / ((l0 + l1) + (l2 + l3));
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(tmem_O + lane_off, o0);
    tmem_ld_32x32b_x32(tmem_O + lane_off + 32, o1);
    tmem_ld_wait();
    if (q0 + row < p.Tq) {
      __nv_bfloat16* orow = p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o0[i + 0]) * inv, __uint_as_float(o0[i + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + i) = o;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o1[i + 0]) * inv, __uint_as_float(o1[i + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + 32 + i) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<kShortTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Resident variant for K/V sweeps whose tile count is a few per SM (self-attention at T = 1024, CFG batch 2: 320 query
// tiles on 148 SMs): the short kernel's footprint (128 TMEM columns, 160 threads) with a loop over 64-key blocks, THREE
// CTAs resident per SM (measured: the launch time steps between 444 and 448 tiles).  Each CTA is a plain serial chain
// (S -> softmax -> P.V per block, nothing overlapped inside the CTA); the overlap comes from the sibling CTAs, and --
// the point -- every tile of such a launch is resident at once: the pipelined kernel below owns an SM, so 320 tiles
// are 148 + 148 + 24, three rounds for 2.16 rounds of work (28 - 29 us against 25 - 26 us here).  Beyond 3 tiles per SM
// the pipelined kernel wins (~725 against ~940 cycles per 128 x 64 block and SM); the launcher picks by tile count.
//
//   warp 0      TMA loads (Q once; K / V rings of kStages 64-row blocks) and both MMA batches
//   warps 1-4   one query row per thread: the 64 scores of the block in registers, exact block max, lazy reference
//               (moves when the max grows by more than 2^8; then O in TMEM and l are rescaled by this thread -- S(j)
//               complete implies P(j-1).V(j-1) complete, both were issued by the same thread and one commit covers
//               them), exponentials, bf16 P over the dead S columns
//
//   TMEM   S [0,64) fp32;  P [0,32) packed bf16, aliasing S;  O [64,128), accumulated over the whole sweep.
//          P(j).V(j) and S(j+1) are issued back to back: tcgen05.mma executes in issue order, so S(j+1) overwrites
//          the P(j) columns only after P(j).V(j) has read them.
//
// Block period of one CTA (clock64 stamps, selftest attn1 with ST_ATTN_IMPL=resident): ~1.75 k cycles = S in registers
// 90 + max / reference 200 + 64 exponentials, sums, packs, P stores 730 + store wait / fence / arrive 90 + issuer wake-up
// 45 + 8 MMAs and 2 commits issued 360 + execution, commit, wake-up 210.  Variants measured and dropped
// (profiles/r02_attention_experiments.txt, section 8): two softmax warps per row, 32-key ping-pong S buffers, a
// four-warp CTA whose warp 0 also issues, half of the exponentials on the FMA pipe, busy-polling waits.
constexpr int kResKV = 64;
constexpr int kResThreads = 160;
constexpr int kResKVBytes = kResKV * kAttnD * 2;
constexpr int kResTmemCols = 128;
constexpr int kResStages = 2;
constexpr int kResSmemBytes = kAttnTileBytes + 2 * kResStages * kResKVBytes + 256 + 1024;
constexpr int kResCtasPerSm = 3;
constexpr float kResTau = 8.f;  // log2 units: the row reference moves when the block max exceeds it by more than 2^8

// kCtas = 4: sized for four CTAs per SM (<= 96 registers): a thread keeps only 32 of its 64 scores -- the upper half is read
// for the maximum, dropped, and read again from TMEM (P overwrites only the lower 32 columns) for its exponentials.  Chosen
// for saturated launches (resident_four_ctas): more blocks per SM and second, at the price of a longer chain per CTA.
template <int kCtas>
__global__ void __launch_bounds__(kResThreads, kCtas)
attn_fwd_resident_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kStages = kResStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTileBytes;
  uint8_t* sV = sK + kStages * kResKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * kResKVBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* v_full = k_full + kStages;
  uint64_t* kv_empty = v_full + kStages;
  uint64_t* s_full = kv_empty + kStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y - b * p.H;
  const int nkv = (p.Tk + kResKV - 1) / kResKV;
#define RES_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) p.trace[slot] = clock64();            \
  } while (0)
  if (threadIdx.x == 0) RES_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);  // one arrival per softmax warp
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<kResTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // prologue above touched only shared / tensor memory
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_P = tmem_base;
  const uint32_t tmem_O = tmem_base + 64;

  if (warp == 0) {
    // ===================================== TMA + MMA issue ==================================
    auto load_kv = [&](int j) {  // lane 0 only
      const int st = j % kStages;
      mbar_expect_tx(&k_full[st], kResKVBytes);
      tma_load_4d(sK + st * kResKVBytes, &tmap_k, &k_full[st], 0, j * kResKV, h, b);
      mbar_expect_tx(&v_full[st], kResKVBytes);
      tma_load_4d(sV + st * kResKVBytes, &tmap_v, &v_full[st], 0, j * kResKV, h, b);
    };
    if (lane == 0) {
      mbar_expect_tx(q_full, kAttnTileBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, q0, h, b);
      for (int j = 0; j < kStages && j < nkv; ++j) load_kv(j);
    }
    __syncwarp();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kResKV, 0, 0);  // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);  // P (TMEM)    x V (MN-major)
    constexpr uint32_t kStep = kResKVBytes >> 4;                      // descriptor address units (16 bytes)
    const uint64_t desc_q = umma_smem_desc_sw128(smem_u32(sQ), 0, 1024);
    const uint64_t desc_k0 = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
    const uint64_t desc_v0 = umma_smem_desc_sw128(smem_u32(sV), 8192, 1024);
    auto issue_s = [&](int j) {  // K(j) has landed (waited for by the caller, off the P -> S critical path)
      const uint64_t dk = desc_k0 + static_cast<uint64_t>((j % kStages) * kStep);
#pragma unroll
      for (int k = 0; k < kAttnD / 16; ++k) umma_bf16_ss_elect(tmem_S, desc_q + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit_elect(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int st = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      mbar_wait(&v_full[st], ph);
      if (j + 1 < nkv) mbar_wait(&k_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
      if (j == 4) RES_TRACE(8);
      mbar_wait(p_full, j & 1);  // P(j) is in TMEM (and O has been rescaled if the row references moved)
      tc_fence_after();
      if (j == 4) RES_TRACE(9);
      const uint64_t dv = desc_v0 + static_cast<uint64_t>(st * kStep);
#pragma unroll
      for (int kk = 0; kk < kResKV / 16; ++kk)  // 16 V rows (2 KB) per instruction
        umma_bf16_ts_elect(tmem_O, tmem_P + kk * 8, dv + 128 * kk, idesc_o, (j > 0) || (kk != 0));
      umma_commit_elect(&kv_empty[st]);
      if (j + 1 < nkv)
        issue_s(j + 1);
      else
        umma_commit_elect(o_full);
      if (j == 4) RES_TRACE(10);
      if (j + kStages < nkv) {  // refill the stage block j has just released
        mbar_wait(&kv_empty[st], ph);
        if (lane == 0) load_kv(j + kStages);
        __syncwarp();
      }
    }
  } else {
    // ===================================== softmax ==========================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_S + lane_off;
    const uint32_t t_p = tmem_P + lane_off;
    const uint32_t t_o = tmem_O + lane_off;
    float m = -INFINITY, l = 0.f;  // m: reference of my row (scaled scores, log2 domain)
    for (int j = 0; j < nkv; ++j) {
      const int valid = p.Tk - j * kResKV;  // columns >= valid are padding (K rows zero-filled by TMA)
      if (j == 4 && warp == 1) RES_TRACE(1);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j == 4 && warp == 1) RES_TRACE(2);
      if (j == 5 && warp == 1) RES_TRACE(12);
      constexpr bool kReload = kCtas >= 4;
      uint32_t s[kReload ? 32 : 64];  // kReload: the lower 32 columns; the upper 32 live in `hi` only while they are needed
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (kReload) {
        uint32_t hi[32];
        tmem_ld_32x32b_x32(t_s + 32, hi);
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(s));
        tmem_ld_wait();
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= valid) s[i] = 0xff800000u;
            if (32 + i >= valid) hi[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          mx0 = max3f(mx0, __uint_as_float(hi[i]), __uint_as_float(hi[i + 1]));
          mx1 = max3f(mx1, __uint_as_float(hi[i + 2]), __uint_as_float(hi[i + 3]));
          mx2 = max3f(mx2, __uint_as_float(hi[i + 4]), __uint_as_float(hi[i + 5]));
          mx3 = max3f(mx3, __uint_as_float(hi[i + 6]), __uint_as_float(hi[i + 7]));
        }
      } else {
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(s));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(s + 32));
        tmem_ld_wait();
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf -> 2^(-inf) = 0
        }
      }
      if (j == 4 && warp == 1) RES_TRACE(3);
#pragma unroll
      for (int i = 0; i < (kReload ? 32 : 64); i += 8) {
        mx0 = max3f(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mx1 = max3f(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = max3f(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = max3f(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      const float mj = max3f(mx0, mx1, fmaxf(mx2, mx3)) * p.scale_log2;  // scale > 0 commutes with max; column 0 is valid
      const bool move = mj > m + kResTau;                                // always true on the first block (m = -inf)
      const float m_new = move ? mj : m;
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        const float f = ex2_approx(m - m_new);  // exactly 1 for rows that keep their reference
        l *= f;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(t_o + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st_32x32b_x16(t_o + c, *reinterpret_cast<const uint32_t(*)[16]>(o));
          tmem_st_32x32b_x16(t_o + c + 16, *reinterpret_cast<const uint32_t(*)[16]>(o + 16));
        }
      }
      m = m_new;
      const float neg_m = -m;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      if (j == 4 && warp == 1) RES_TRACE(4);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        if (kReload && c == 32) {  // S columns [32, 64) again: P(j) so far covers columns [0, 16) only
          tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(s));
          tmem_ld_wait();
          if (valid < 64) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (32 + i >= valid) s[i] = 0xff800000u;
          }
        }
        const int sc = kReload ? 0 : c;
        float e[32];
#pragma unroll
        for (int q = 0; q < 16; ++q)
          fma2_bcast(__uint_as_float(s[sc + 2 * q]), __uint_as_float(s[sc + 2 * q + 1]), p.scale_log2, neg_m, e[2 * q], e[2 * q + 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = ex2_approx_ordered(e[i]);
        ready16(e, 0);
        ready16(e, 16);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          add2_acc(rs0, rs1, e[i + 0], e[i + 1]);
          add2_acc(rs2, rs3, e[i + 2], e[i + 3]);
          pk[(i >> 1) + 0] = pack_bf16x2(e[i + 0], e[i + 1]);
          pk[(i >> 1) + 1] = pack_bf16x2(e[i + 2], e[i + 3]);
        }
        tmem_st_32x32b_x16(t_p + (c >> 1), pk);
      }
      l += (rs0 + rs1) + (rs2 + rs3);
      if (j == 4 && warp == 1) RES_TRACE(5);
      tmem_st_wait();
      if (j == 4 && warp == 1) RES_TRACE(6);
      tc_fence_before();
      warp_arrive(p_full, lane);
      if (j == 4 && warp == 1) RES_TRACE(7);
    }
    const float inv = 1.f / l;
    mbar_wait(o_full, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_o + c, o);  // warp-collective: outside the row guard
      tmem_ld_wait();
      if (q0 + row < p.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 ov;
          ov.x = pack_bf16x2(__uint_as_float(o[i + 0]) * inv, __uint_as_float(o[i + 1]) * inv);
          ov.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          ov.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          ov.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + i) = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<kResTmemCols>(tmem_base);
  if (threadIdx.x == 0) RES_TRACE(11);
#undef RES_TRACE
}

// ------------------------------------------------------------------------------------------------
// Pipelined variant for Tk > 128 (self-attention): ONE CTA per SM that owns all 512 TMEM columns, so S and P
// are double-buffered and O accumulates in TMEM across the whole K/V sweep; warp-specialised (warp numbers for the
// 8-exp-warp layout, kParts = 2; with kParts = 4 the exp warps are 2-17 and the rest follow):
//
//   warp 0      TMA producer: Q once, K and V rings (4 stages each)
//   warp 14     S issuer:     S(0), S(1); for j: wait s_free(j) -> S(j+2)          (S runs two blocks ahead)
//   warp 1      P.V issuer:   for j: wait P(j) -> O += P(j) V(j)
//   warps 2-9   exp warps:    kParts per TMEM lane quadrant, each owning a column slice of every row: stream S(j) out of
//               TMEM in 32-column chunks, p = 2^(s*c - m), packed bf16 P(j) -> TMEM, partial row sums.
//   warps 10-13 max warps:    one thread per row, run one block AHEAD of the exp warps: exact row max of S(j)
//               (FMNMX3), decide the row's reference m, publish it through shared memory.  m only moves when
//               the row max grows by more than 2^8 (lazy rescale: P <= 256, fp32 sums are safe); moving it
//               multiplies O (in TMEM) by 2^(m_old - m_new) -- rare after the first blocks -- so there is no
//               per-block read-modify-write of O at all.
//
//   TMEM   S0 [0,128)  S1 [128,256)  P0 [256,320)  P1 [320,384)  P2 [384,448)  O [448,512)
//
// Per 128x128 block the tensor pipe needs 512 cycles (4 x 64 + 8 x 32), the MUFU pipe 1024 (16 ex2/clk/SM);
// the two-CTA kernel above spends ~2.7 k cycles per block and SM because each CTA's MMA -> max -> exp -> MMA
// chain is serial (single S / P buffers in 256 columns) and both CTAs hit the MUFU phase together.
constexpr int kA3Stages = 4;
// kParts exp warps per TMEM lane quadrant, each owning 128 / kParts score columns of every row.  4 (the default): 16 exp
// warps (736 threads, 78 registers), one 32-column chunk per warp and block -- four instruction streams per scheduler to
// cover the TMEM / barrier latencies of each warp's serial chain (ncu: pipe_xu 64 % busy against 49 % with two).  2: the
// round-1 layout (480 threads, two 32-column chunks per warp and block, next block's first chunk prefetched); the trace,
// polynomial and ablation instantiations exist in this layout only.
template <int kParts>
struct A3Layout {
  static constexpr int kExpWarps = 4 * kParts;
  static constexpr int kMaxWarp0 = 2 + kExpWarps;   // 4 max warps
  static constexpr int kSWarp = kMaxWarp0 + 4;      // S = Q K^T issuer
  static constexpr int kThreads = 32 * (kSWarp + 1);
};
constexpr int kA3Threads = A3Layout<2>::kThreads;   // 480
constexpr int kA3SmemBytes = kAttnTileBytes * (1 + 2 * kA3Stages) + 256 + 4096 + 1024;
constexpr int kA3TmemCols = 512;
constexpr float kA3Tau = 8.f;  // log2 units

// kPoly: of every 8 element pairs an exp warp (kParts == 2 layout) processes, kPoly take the FMA-pipe polynomial
// (ex2_poly2, ptx.cuh) instead of MUFU.EX2; the pairs are spread evenly (Bresenham) so the polynomial's FMA-pipe work
// issues in the shadow of the MUFU stream.  Measured (profiles/r02_attention_experiments.txt): moving 12.5 - 37.5 % of
// the exponentials off the MUFU pipe changes nothing (T = 4096: 144 us either way), 50 % is slower -- the block period
// (~1.45 k cycles) is not set by MUFU throughput.  Opt-in (ST_ATTN_POLY=2); the default instantiation has kPoly = 0.
__host__ __device__ constexpr bool a3_is_poly_pair(int q, int poly) { return ((q % 8 + 1) * poly) / 8 != ((q % 8) * poly) / 8; }

// kAblate: timing-only instantiation for bottleneck hunting -- p.ablate bits switch pieces of the pipeline off (1: no MUFU
// in the exp warps, 2: no tcgen05.st of P, 4: P.V MMAs not issued, 8: max warps do not read S, 16: no row sums / bf16
// packs, 32: S MMAs not issued, 64: no K/V TMA loads after the first ring pass, 128: exp warps do not read S).  Results
// are wrong by construction; never selected unless ST_ATTN_ABLATE is set.
template <bool kTrace, int kParts = 2, int kPoly = 0, bool kAblate = false>
__global__ void __launch_bounds__(A3Layout<kParts>::kThreads, 1)
attn_fwd_pipelined_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using L = A3Layout<kParts>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTileBytes;
  uint8_t* sV = sK + kA3Stages * kAttnTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA3Stages * kAttnTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + kA3Stages;
  uint64_t* v_full = k_empty + kA3Stages;
  uint64_t* v_empty = v_full + kA3Stages;
  uint64_t* s_full = v_empty + kA3Stages;  // [2]
  uint64_t* m_ready = s_full + 2;          // [2]
  uint64_t* p_full = m_ready + 2;          // [3]
  uint64_t* pv_done = p_full + 3;          // [2]
  uint64_t* s_free = pv_done + 2;          // [2]
  uint64_t* p_free = s_free + 2;           // [3]  P(i).V(i) has drained P buffer i % 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 3);
  float* s_m = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 buffers][128 rows] references
  float* s_l = s_m + 256;                                                         // [kParts][128 rows] row sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y - b * p.H;
  const int nkv = (p.Tk + kAttnBlockKV - 1) / kAttnBlockKV;
  if (threadIdx.x == 0) AT_TRACE(0);
#define A3_WAIT(bar, parity) mbar_wait_relaxed(bar, parity, p.wait_hint_ns, p.wait_sleep_ns)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kA3Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&m_ready[i], 4);  // one arrival per warp (warp_arrive)
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 4 * kParts);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&p_full[i], 4 * kParts);
      mbar_init(&p_free[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kA3TmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;        // + 128 * (j & 1)
  const uint32_t tmem_P = tmem_base + 256;  // + 64 * (j % 3)
  const uint32_t tmem_O = tmem_base + 448;
  const int quad = warp & 3;                // TMEM lane quadrant this warp may touch
  const int row = quad * 32 + lane;
  const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, kAttnTileBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA3Stages;
        const uint32_t ph = (j / kA3Stages) & 1;
        if (kAblate && (p.ablate & 64) && j >= kA3Stages) {  // no K/V traffic after the first ring pass
          A3_WAIT(&k_empty[st], ph ^ 1);
          mbar_arrive(&k_full[st]);
          A3_WAIT(&v_empty[st], ph ^ 1);
          mbar_arrive(&v_full[st]);
          continue;
        }
        A3_WAIT(&k_empty[st], ph ^ 1);
        if (j == 20) AT_TRACE(44);
        mbar_expect_tx(&k_full[st], kAttnTileBytes);
        tma_load_4d(sK + st * kAttnTileBytes, &tmap_k, &k_full[st], 0, j * kAttnBlockKV, h, b);
        A3_WAIT(&v_empty[st], ph ^ 1);
        if (j == 20) AT_TRACE(45);
        mbar_expect_tx(&v_full[st], kAttnTileBytes);
        tma_load_4d(sV + st * kAttnTileBytes, &tmap_v, &v_full[st], 0, j * kAttnBlockKV, h, b);
      }
    }
  } else if (warp == 1 || warp == L::kSWarp) {
    // ===================================== MMA issuers ======================================
    // Two issuing warps on different SM sub-partitions: warp 14 the S = Q K^T products (4 MMAs per block), warp 1 the
    // O += P V products (8 per block).  A single issuer (~250 instructions per block, all on one scheduler) made the
    // two exp warps that share its sub-partition the slowest of the eight, and the slowest quadrant sets the pace.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kAttnBlockKV, 0, 0);  // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);        // P (TMEM)    x V (MN-major)
    constexpr uint32_t kTileStep = kAttnTileBytes >> 4;
    if (warp == L::kSWarp) {
      const uint64_t desc_q = umma_smem_desc_sw128(smem_u32(sQ), 0, 1024);
      const uint64_t desc_k0 = umma_smem_desc_sw128(smem_u32(sK), 0, 1024);
      auto issue_s = [&](int j) {
        const int st = j % kA3Stages;
        A3_WAIT(&k_full[st], (j / kA3Stages) & 1);
        tc_fence_after();
        const uint64_t dk = desc_k0 + static_cast<uint64_t>(st * kTileStep);
        const uint32_t d = tmem_S + (j & 1) * 128;
        if (!(kAblate && (p.ablate & 32))) {
#pragma unroll
          for (int k = 0; k < kAttnD / 16; ++k) umma_bf16_ss_elect(d, desc_q + 2 * k, dk + 2 * k, idesc_s, k != 0);
        }
        umma_commit_elect(&k_empty[st]);
        umma_commit_elect(&s_full[j & 1]);
      };
      A3_WAIT(q_full, 0);
      issue_s(0);
      if (nkv > 1) issue_s(1);
      // S(j+2) goes out as soon as the exp warps have pulled S(j) out of TMEM (s_free, a quarter into block j); the
      // max warps read S(j) before that (the exp warps wait for m_ready(j)).  So the S -> max -> reference chain of
      // block j+2 starts well over a block ahead of its use.
      for (int j = 0; j + 2 < nkv; ++j) {
        A3_WAIT(&s_free[j & 1], (j >> 1) & 1);
        tc_fence_after();
        if (j == 16 && lane == 0) AT_TRACE(58);
        issue_s(j + 2);
        if (j == 16 && lane == 0) AT_TRACE(59);
      }
    } else {
      const uint64_t desc_v0 = umma_smem_desc_sw128(smem_u32(sV), 8192, 1024);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA3Stages;
        A3_WAIT(&p_full[j % 3], (j / 3) & 1);
        if (j == 2 && lane == 0) AT_TRACE(8);
        if (j == 16 && lane == 0) AT_TRACE(60);
        A3_WAIT(&v_full[st], (j / kA3Stages) & 1);
        tc_fence_after();
        if (j == 16 && lane == 0) AT_TRACE(61);
        const uint64_t dv = desc_v0 + static_cast<uint64_t>(st * kTileStep);
        const uint32_t a = tmem_P + (j % 3) * 64;
        if (!(kAblate && (p.ablate & 4))) {
#pragma unroll
          for (int kk = 0; kk < kAttnBlockKV / 16; ++kk)
            umma_bf16_ts_elect(tmem_O, a + kk * 8, dv + 128 * kk, idesc_o, (j > 0) || (kk != 0));
        }
        umma_commit_elect(&v_empty[st]);
        umma_commit_elect(&p_free[j % 3]);
        umma_commit_elect(&pv_done[j & 1]);
        if (j == 2 && lane == 0) AT_TRACE(9);
        if (j == 16 && lane == 0) AT_TRACE(62);
      }
    }
  } else if (warp >= L::kMaxWarp0) {  // four warps, one per TMEM lane quadrant
    // ===================================== max warps ========================================
    float m = -INFINITY;  // reference of my row (scaled scores, log2 domain)
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      const int valid = p.Tk - j * kAttnBlockKV;  // columns >= valid are padding
      A3_WAIT(&s_full[buf], (j >> 1) & 1);
      tc_fence_after();
      if (j == 18 && warp == L::kMaxWarp0 && lane == 0) AT_TRACE(46);
      const uint32_t t_s = tmem_S + buf * 128 + lane_off;
      // two 64-column loads, one TMEM round trip each (four 32-column loads took ~900 cycles per block, and this
      // warp's S -> reference latency is on the per-block critical chain s_free -> S -> max -> m_ready)
      float mx = -INFINITY;
      uint32_t v[64];
      if (kAblate && (p.ablate & 8)) mx = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 64) {
        if (kAblate && (p.ablate & 8)) break;
        if (kParts == 4) {  // 736 threads leave 80 registers: a single 64-register load does not fit next to its context
          tmem_ld_32x32b_x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(v));
          tmem_ld_32x32b_x32(t_s + c + 32, *reinterpret_cast<uint32_t(*)[32]>(v + 32));
        } else {
          tmem_ld_32x32b_x64(t_s + c, v);
        }
        tmem_ld_wait();
        if (c + 64 > valid) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (c + i >= valid) v[i] = 0xff800000u;  // -inf
        }
        float g[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float a0 = max3f(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]), __uint_as_float(v[8 * q + 2]));
          const float a1 = max3f(__uint_as_float(v[8 * q + 3]), __uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
          g[q] = max3f(a0, a1, fmaxf(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
        }
        mx = max3f(mx, max3f(g[0], g[1], g[2]), max3f(g[3], g[4], g[5]));
        mx = max3f(mx, g[6], g[7]);
      }
      const float mj = mx * p.scale_log2;  // scale > 0 commutes with max
      const bool move = mj > m + kA3Tau;   // always true on the first block (m = -inf)
      const float m_new = move ? mj : m;
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        const float f = ex2_approx(m - m_new);  // exactly 1 for rows that keep their reference
        A3_WAIT(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // every P.V issued so far has landed in O
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tmem_O + lane_off + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st_32x32b_x16(tmem_O + lane_off + c, *reinterpret_cast<const uint32_t(*)[16]>(o));
          tmem_st_32x32b_x16(tmem_O + lane_off + c + 16, *reinterpret_cast<const uint32_t(*)[16]>(o + 16));
        }
        tmem_st_wait();
      }
      m = m_new;
      s_m[buf * 128 + row] = m;
      // P is triple-buffered and P(j) overwrites P(j-3): publishing m(j) also vouches that P(j-3).V(j-3) has drained
      // (this warp runs a block ahead and has the time to look; the exp warps do not)
      // (its own barrier per buffer: the next arrival on p_free[j % 3] is P(j).V(j), which needs this very m(j), so
      // the parity wait cannot alias)
      if (j >= 3) mbar_wait(&p_free[j % 3], (j / 3 - 1) & 1);
      if (j == 18 && warp == L::kMaxWarp0 && lane == 0) AT_TRACE(47);
      tc_fence_before();
      if (j == 2 && warp == L::kMaxWarp0 && lane == 0) AT_TRACE(12);
      if (kParts == 2 && (j == 16 || j == 17) && lane == 0) AT_TRACE(16 * (j - 15) + 8 + warp - L::kMaxWarp0);
      if (j == 3 && warp == L::kMaxWarp0 && lane == 0) AT_TRACE(13);
      warp_arrive(&m_ready[buf], lane);  // release: the exp warps' wait acquires s_m and orders the O rescale before P(j)
    }
  } else if (kParts == 4) {
    // ===================================== exp warps, 4 per quadrant ========================
    // One 32-column chunk of S(j) per warp and block: wait for the row references, pull the chunk out of TMEM, release
    // the S buffer, 32 exponentials back to back, sums + bf16 packs, P back into TMEM, publish.  No software prefetch
    // of the next block (<= 88 registers at 736 threads): the three sibling warps on the scheduler cover the latencies.
    const int part = (warp - 2) >> 2;            // columns [32 * part, 32 * part + 32) of every block
    const uint32_t quad_bar = 1 + quad;          // named barrier of this quadrant's four exp warps
    float m_prev = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      const int valid = p.Tk - j * kAttnBlockKV - part * 32;  // my columns >= valid are padding
      uint32_t cur[32];
      mbar_wait(&m_ready[buf], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(tmem_S + buf * 128 + lane_off + part * 32, cur);
      const float m = s_m[buf * 128 + row];
      if (m != m_prev) {
        l *= ex2_approx(m_prev - m);  // first block: l = 0 and 2^(-inf) = 0
        m_prev = m;
      }
      tmem_ld_wait();
      tc_fence_before();
      warp_arrive(&s_free[buf], lane);  // my share of S(j) is in registers: S(j+2) may overwrite the buffer
      if (valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) cur[i] = 0xff800000u;  // -inf -> 2^(-inf) = 0
      }
      float e[32];
      {
        // t = s * c - m as packed pairs (FFMA2), then the 32 exponentials back to back
        float t[32];
        const float neg_m = -m;
#pragma unroll
        for (int q = 0; q < 16; ++q)
          fma2_bcast(__uint_as_float(cur[2 * q]), __uint_as_float(cur[2 * q + 1]), p.scale_log2, neg_m, t[2 * q], t[2 * q + 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = ex2_approx_ordered(t[i]);
      }
      ready16(e, 0);
      ready16(e, 16);
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        add2_acc(rs0, rs1, e[i + 0], e[i + 1]);
        add2_acc(rs2, rs3, e[i + 2], e[i + 3]);
        pk[(i >> 1) + 0] = pack_bf16x2(e[i + 0], e[i + 1]);
        pk[(i >> 1) + 1] = pack_bf16x2(e[i + 2], e[i + 3]);
      }
      l += (rs0 + rs1) + (rs2 + rs3);
      tmem_st_32x32b_x16(tmem_P + (j % 3) * 64 + lane_off + part * 16, pk);
      tmem_st_wait();
      tc_fence_before();
      warp_arrive(&p_full[j % 3], lane);
    }
    {
      mbar_wait(&pv_done[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
      tc_fence_after();
      s_l[part * 128 + row] = l;
      asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");
      const float inv = 1.f / ((s_l[row] + s_l[128 + row]) + (s_l[256 + row] + s_l[384 + row]));
      uint32_t o[16];
      tmem_ld_32x32b_x16(tmem_O + lane_off + part * 16, o);
      tmem_ld_wait();
      if (q0 + row < p.Tq) {
        __nv_bfloat16* orow =
            p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st + part * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 ov;
          ov.x = pack_bf16x2(__uint_as_float(o[i + 0]) * inv, __uint_as_float(o[i + 1]) * inv);
          ov.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          ov.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          ov.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + i) = ov;
        }
      }
    }
  } else {
    // ===================================== exp warps ========================================
    const int half = (warp - 2) >> 2;
    const uint32_t pair_bar = 1 + quad;
    float m_prev = -INFINITY, l = 0.f;  // l: my half of the row sum, relative to m_prev
    // S(j) arrives in two 32-column chunks per thread (va, vb).  While the second chunk of block j is being
    // exponentiated, the first chunk of block j+1 is already requested (its reference is normally published a
    // block ahead), so the barrier / shared-memory / TMEM latencies at a block boundary hide under the MUFU stream.
    uint32_t va[32], vb[32];
    bool pre = false;   // va already holds (a request for) the first chunk of the next block; warp-uniform
    float m_pre = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      const int valid = p.Tk - j * kAttnBlockKV - half * 64;  // my columns >= valid are padding
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(1);
      if (j == nkv - 1 && warp == 2 && lane == 0) AT_TRACE(10);
      if ((j == 16 || j == 17) && lane == 0) AT_TRACE(16 * (j - 15) + warp - 2);  // per-warp block period / skew
      const uint32_t t_s = tmem_S + buf * 128 + lane_off + half * 64;
      // P is triple-buffered: P(j) overwrites P(j-3), and the max warps only publish m(j) once P(j-3).V(j-3) has
      // drained -- so knowing m(j) proves the buffer is free, no barrier to poll here
      const uint32_t t_p = tmem_P + (j % 3) * 64 + lane_off + half * 32;
      float m;
      if (!pre) {
        // the max warps arrive on m_ready(j) after they have seen s_full(j): S(j) is complete
        mbar_wait(&m_ready[buf], (j >> 1) & 1);
        tc_fence_after();
        if (!(kAblate && (p.ablate & 128))) tmem_ld_32x32b_x32(t_s, va);
        m = s_m[buf * 128 + row];
      } else {
        m = m_pre;
      }
      if (kTrace && j == 17 && warp == 2 && lane == 0) {
        AT_TRACE(31);
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[15] = pre ? 1 : 2;
      }
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(2);
      if (m != m_prev) {
        l *= ex2_approx(m_prev - m);  // first block: l = 0 and 2^(-inf) = 0
        m_prev = m;
      }
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(3);
      const bool tr = kTrace && j == 16 && warp == 2 && lane == 0;  // detailed timeline of one steady-state block
      if (tr) AT_TRACE(48);
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      tmem_ld_wait();
      if (tr) AT_TRACE(49);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t(&cur)[32] = ((c >> 5) & 1) ? vb : va;
        if (c == 0) {
          if (!(kAblate && (p.ablate & 128))) tmem_ld_32x32b_x32(t_s + 32, vb);
        } else {
          pre = false;
          if (j + 1 < nkv) {
            const bool ok = mbar_test_wait(&m_ready[buf ^ 1], ((j + 1) >> 1) & 1);  // a poll, never a suspend
            if (__all_sync(0xffffffffu, ok)) {  // tcgen05.ld is warp-collective: decide as a warp
              tc_fence_after();
              if (!(kAblate && (p.ablate & 128))) tmem_ld_32x32b_x32(tmem_S + (buf ^ 1) * 128 + lane_off + half * 64, va);
              m_pre = s_m[(buf ^ 1) * 128 + row];
              pre = true;
            }
          }
        }
        if (c + 32 > valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= valid) cur[i] = 0xff800000u;  // -inf -> fma(-inf, c, -m) = -inf -> 2^(-inf) = 0
        }
        // All 32 exponentials first, back to back, THEN the sums and the bf16 packs: with F2FP / FADD placed right
        // behind their MUFUs (what the compiler does on its own) the in-order warp waits out the MUFU latency for
        // every pair -- 22 cycles per element, independent of what the sibling warp does (measured).
        float e[32];
        {
          // t = s * c - m for all 32 columns (FFMA2: one issue slot per pair)
          const float neg_m = -m;
          float t[32];
#pragma unroll
          for (int q = 0; q < 16; ++q)
            fma2_bcast(__uint_as_float(cur[2 * q]), __uint_as_float(cur[2 * q + 1]), p.scale_log2, neg_m, t[2 * q], t[2 * q + 1]);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (!a3_is_poly_pair(q, kPoly)) {
              if (kAblate && (p.ablate & 1)) {
                e[2 * q] = t[2 * q];
                e[2 * q + 1] = t[2 * q + 1];
              } else {
                e[2 * q] = ex2_approx_ordered(t[2 * q]);
                e[2 * q + 1] = ex2_approx_ordered(t[2 * q + 1]);
              }
            }
          if (c == 0) {
            // the second chunk has landed by now: all of my S(j) is in registers, and the sooner the MMA warp may
            // overwrite this S buffer with S(j+2), the sooner the max warps get to publish m(j+2)
            tmem_ld_wait();
            tc_fence_before();
            warp_arrive(&s_free[buf], lane);
          }
#pragma unroll
          for (int q = 8; q < 16; ++q)
            if (!a3_is_poly_pair(q, kPoly)) {
              if (kAblate && (p.ablate & 1)) {
                e[2 * q] = t[2 * q];
                e[2 * q + 1] = t[2 * q + 1];
              } else {
                e[2 * q] = ex2_approx_ordered(t[2 * q]);
                e[2 * q + 1] = ex2_approx_ordered(t[2 * q + 1]);
              }
            }
          // the polynomial pairs: plain (non-volatile) FMA-pipe code, free to be scheduled between the MUFUs above
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (a3_is_poly_pair(q, kPoly)) ex2_poly2(t[2 * q], t[2 * q + 1], e[2 * q], e[2 * q + 1]);
        }
        ready16(e, 0);
        ready16(e, 16);
        if (tr) AT_TRACE(c == 0 ? 50 : 55);
        uint32_t pk[16];
        if (kAblate && (p.ablate & 16)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = __float_as_uint(e[2 * i]) ^ __float_as_uint(e[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            add2_acc(rs0, rs1, e[i + 0], e[i + 1]);
            add2_acc(rs2, rs3, e[i + 2], e[i + 3]);
            pk[(i >> 1) + 0] = pack_bf16x2(e[i + 0], e[i + 1]);
            pk[(i >> 1) + 1] = pack_bf16x2(e[i + 2], e[i + 3]);
          }
        }
        if (tr && c == 0) AT_TRACE(51);
        if (!(kAblate && (p.ablate & 2))) tmem_st_32x32b_x16(t_p + (c >> 1), pk);
        else asm volatile("" ::"r"(pk[0] ^ pk[5] ^ pk[10] ^ pk[15]));
        if (tr && c == 0) AT_TRACE(57);
        if (tr) AT_TRACE(c == 0 ? 54 : 56);
      }
      l += (rs0 + rs1) + (rs2 + rs3);
      if (j == 2 && warp == 2 && lane == 0) AT_TRACE(6);
      // Publish right away: P(j).V(j) then runs under the next block's first ex2 phase, when these warps do not
      // touch TMEM.  (Publishing later -- after that phase -- was measured: the tcgen05.st of the next chunk then
      // queues behind the running MMA for ~250 cycles.)
      tmem_st_wait();
      if (tr) AT_TRACE(28);
      tc_fence_before();
      if (tr) AT_TRACE(29);
      warp_arrive(&p_full[j % 3], lane);
      if (tr) AT_TRACE(52);
    }
    float acc[32];
    {
      mbar_wait(&pv_done[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
      tc_fence_after();
      // total row sum = my half + the partner's (both relative to the row's final reference)
      s_l[half * 128 + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float inv = 1.f / (l + s_l[(half ^ 1) * 128 + row]);
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_off + half * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(o[i]) * inv;
    }
    if (q0 + row < p.Tq) {
      __nv_bfloat16* orow =
          p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        o.x = pack_bf16x2(acc[i + 0], acc[i + 1]);
        o.y = pack_bf16x2(acc[i + 2], acc[i + 3]);
        o.z = pack_bf16x2(acc[i + 4], acc[i + 5]);
        o.w = pack_bf16x2(acc[i + 6], acc[i + 7]);
        *reinterpret_cast<uint4*>(orow + i) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<kA3TmemCols>(tmem_base);
  if (threadIdx.x == 0) AT_TRACE(11);
}

// 4-D map over a bf16 tensor addressed as [b][h][t][d] with element strides (sb, sh, st, 1) and d = 64;
// box = [1, 1, 128, 64].  Covers both (B, T, H*64) activations (sh = 64, st = row pitch) and the
// reference's (B, H, T, D) layout (kernels/attention_fa2.py:113-140).  OOB rows (t >= T) read as zero.
static int make_tmap_bhtd(CUtensorMap* out, const void* base, int B, int H, int T, long long sb, long long sh,
                          long long st_, int box_rows = 128) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return ST_ERR_CUDA;
    }
    fn = reinterpret_cast<Fn>(ptr);
  }
  cuuint64_t dims[4] = {64, (cuuint64_t)T, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)st_ * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(bhtd B=%d H=%d T=%d sb=%lld sh=%lld st=%lld) failed: %d", B, H, T, sb, sh, st_,
              (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

// Which kernel for a K/V sweep (Tk > 128)?  Cycle model calibrated on B200 (profiles/r02_attention_experiments.txt, 8):
// the pipelined kernel owns an SM -- rounds of `sms` tiles, ~1.45 k cycles per 128-key block + ~7 k of prologue / tail per
// round; the resident kernel holds up to three tiles per SM at once -- a 64-key block costs a CTA ~1.75 k cycles alone and
// ~535 more per sibling CTA on its SM, an SM finishes one every ~940 cycles when it is full, + ~8.6 k once.  So the
// resident kernel takes the short sweeps of more than two tiles per SM (T = 1024 self-attention: 320 tiles 28.4 -> 25.8 us,
// 640 tiles 48 -> 42 us), the pipelined one the long sweeps (T >= 4096) and the launches of one or two rounds.
// The resident kernel exists in two register budgets: three CTAs per SM (a thread keeps its 64 scores; the shorter chain --
// launches of a few tiles per SM) and four (<= 96 registers: the upper 32 scores are re-read from TMEM; ~8 % more blocks per
// SM and second once the machine is full: B16 T1024 138 -> 128 us, but B2 T1024 26 -> 33 us).
static bool resident_four_ctas(long long tiles, int sms) { return tiles >= 8LL * sms; }

static bool sweep_prefers_resident(long long tiles, int Tk, int sms) {
  const long long rounds = (tiles + sms - 1) / sms;
  const long long blocks128 = (Tk + kAttnBlockKV - 1) / kAttnBlockKV, blocks64 = (Tk + kResKV - 1) / kResKV;
  const long long pipelined = rounds * (blocks128 * 1450 + 7000);
  const long long siblings = (rounds < kResCtasPerSm ? rounds : kResCtasPerSm) - 1;
  const long long chain = blocks64 * (1750 + 535 * siblings), throughput = tiles * blocks64 * 940 / sms;
  if ((chain > throughput ? chain : throughput) + 8600 < pipelined) return true;
  // saturated launches (>= 8 tiles per SM) of sweeps up to 4096 keys: the four-CTA form of the resident kernel finishes a
  // 64-key block every ~715 cycles and SM + 2.8 k per tile, the pipelined kernel 2 x 702 per 128 keys + 8.6 k per tile
  // (B4 T4096 248 -> 241 us, B16 T4096 961 -> 872 us; profiles/r02_attention_experiments.txt, 10)
  return resident_four_ctas(tiles, sms) && blocks64 <= 64;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static unsigned long long* g_attn_trace = nullptr;
constexpr int kAttnDefaultParts = 4;  // 16 exp warps: -0.16 ms on the attention launches of a step (gpu_call19)
static int g_attn_parts = 0;  // 0: not decided yet (ST_ATTN_PARTS or the default)
constexpr int kAttnDefaultPoly = 0;
constexpr unsigned kAttnDefaultWaitHint = 1000, kAttnDefaultWaitSleep = 0;  // hint: T = 4096 622 -> 632 TFLOP/s (gpu_call25)
static int g_attn_poly = -1;  // -1: not decided yet (ST_ATTN_POLY or the default)
static int g_attn_impl = -1;  // -1: not decided yet (ST_ATTN_IMPL); 0: by shape; 1 two-CTA, 2 pipelined, 3 short, 4 resident

}  // namespace st

extern "C" {

int st_attention_bf16(const void* q, long long q_sb, long long q_sh, long long q_st, const void* k, long long k_sb,
                      long long k_sh, long long k_st, const void* v, long long v_sb, long long v_sh, long long v_st,
                      void* o, long long o_sb, long long o_sh, long long o_st, int B, int H, int Tq, int Tk,
                      float scale, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(q && k && v && o, "attention: null pointer");
  ST_CHECK_ARG(B > 0 && H > 0 && Tq > 0 && Tk > 0, "attention: sizes must be positive");
  const long long strides[12] = {q_sb, q_sh, q_st, k_sb, k_sh, k_st, v_sb, v_sh, v_st, o_sb, o_sh, o_st};
  for (int i = 0; i < 12; ++i)
    ST_CHECK_ARG(strides[i] > 0 && strides[i] % 8 == 0, "attention: stride %d (%lld) must be a positive multiple of 8",
                 i, strides[i]);
  ST_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "attention: pointers must be 16-byte aligned");
  ST_CHECK_ARG((long long)B * H <= 65535, "attention: B*H too large");
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bhtd(&tq, q, B, H, Tq, q_sb, q_sh, q_st);
  if (rc != ST_OK) return rc;
  // one K/V block (cross-attention, Tk = 77): the short-context kernel (Tk <= 80) or the two-CTA-per-SM kernel
  // (Tk <= 128); longer sweeps: the resident kernel while every query tile of the launch fits on the machine at once
  // (3 CTAs per SM), the pipelined one beyond that
  if (g_attn_impl < 0) {
    const char* e = getenv("ST_ATTN_IMPL");  // debug: "2cta" / "pipelined" / "short" / "resident"
    g_attn_impl = !e ? 0 : (!strcmp(e, "2cta") ? 1 : (!strcmp(e, "pipelined") ? 2 : (!strcmp(e, "short") ? 3 : (!strcmp(e, "resident") ? 4 : (!strcmp(e, "noresident") ? 5 : 0)))));
  }
  const bool no_resident = g_attn_impl == 5;  // A/B runs: choose by shape as before the resident kernel existed
  const int force = no_resident ? 0 : g_attn_impl;
  const long long tiles = static_cast<long long>((Tq + kAttnBlockQ - 1) / kAttnBlockQ) * B * H;
  const bool resident = force == 4 || (force == 0 && !no_resident && Tk > kAttnBlockKV && sweep_prefers_resident(tiles, Tk, device_sm_count()));
  const bool pipelined = !resident && ((force == 1 || force == 2) ? force == 2 : Tk > kAttnBlockKV);
  const bool short_kv = !resident && !pipelined && Tk <= kShortKV && force != 1;
  const int kv_box = resident ? kResKV : (short_kv ? kShortKV : kAttnBlockKV);
  rc = make_tmap_bhtd(&tk, k, B, H, Tk, k_sb, k_sh, k_st, kv_box);
  if (rc != ST_OK) return rc;
  rc = make_tmap_bhtd(&tv, v, B, H, Tk, v_sb, v_sh, v_st, kv_box);
  if (rc != ST_OK) return rc;
  static PerDeviceOnce configured;  // function attributes are per context: once per device, not per process
  const int dev = current_device();
  ST_CHECK_ARG(dev >= 0, "attention: device ordinal outside [0, %d)", kMaxDevices);
  if (!configured.done(dev)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    // two CTAs per SM need 2 x 82 KB: ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(attn_fwd_short_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(attn_fwd_resident_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(attn_fwd_resident_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    e = cudaFuncSetAttribute(attn_fwd_resident_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_resident_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResSmemBytes);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<false, 2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_pipelined_kernel<true, 2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    configured.mark(dev);
  }
  AttnParams p;
  p.O = static_cast<__nv_bfloat16*>(o);
  p.o_sb = o_sb;
  p.o_sh = o_sh;
  p.o_st = o_st;
  p.H = H;
  p.Tq = Tq;
  p.Tk = Tk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  p.ablate = 0;
  static const unsigned wait_hint = [] { const char* e = getenv("ST_ATTN_WAIT_HINT"); return e ? (unsigned)atoi(e) : kAttnDefaultWaitHint; }();
  static const unsigned wait_sleep = [] { const char* e = getenv("ST_ATTN_WAIT_SLEEP"); return e ? (unsigned)atoi(e) : kAttnDefaultWaitSleep; }();
  p.wait_hint_ns = wait_hint;
  p.wait_sleep_ns = wait_sleep;
  const dim3 grid((Tq + kAttnBlockQ - 1) / kAttnBlockQ, B * H);
  const bool trace = p.trace != nullptr;  // the phase stamps are compiled out of the production instantiations
  if (g_attn_parts == 0) {
    const char* e = getenv("ST_ATTN_PARTS");  // exp warps per TMEM lane quadrant: 2 or 4
    g_attn_parts = (e && e[0] == '4') ? 4 : ((e && e[0] == '2') ? 2 : kAttnDefaultParts);
  }
  if (g_attn_poly < 0) {
    const char* e = getenv("ST_ATTN_POLY");  // polynomial pairs out of every 8: 0 or 2 (a quarter of the exponentials)
    g_attn_poly = (e && e[0] == '2') ? 2 : ((e && e[0] == '0') ? 0 : kAttnDefaultPoly);
  }
  static const int ablate = [] {
    const char* e = getenv("ST_ATTN_ABLATE");
    return e ? atoi(e) : 0;
  }();
  // the phase stamps, the polynomial share and the ablation switches exist in the 8-exp-warp layout only
  const int parts = (trace || g_attn_poly == 2 || ablate) ? 2 : g_attn_parts;
  if (resident) {
    static const int ctas_env = [] { const char* e = getenv("ST_ATTN_RES_CTAS"); return e ? atoi(e) : 0; }();  // 3 / 4: force
    const bool four = ctas_env == 4 || (ctas_env != 3 && resident_four_ctas(tiles, device_sm_count()));
    if (four)
      launch_kernel(attn_fwd_resident_kernel<4>, dim3(grid), dim3(kResThreads), kResSmemBytes, static_cast<cudaStream_t>(stream),
                    tq, tk, tv, p);
    else
      launch_kernel(attn_fwd_resident_kernel<3>, dim3(grid), dim3(kResThreads), kResSmemBytes, static_cast<cudaStream_t>(stream),
                    tq, tk, tv, p);
    ST_CHECK_LAUNCH("attn_fwd_resident_kernel");
  } else if (pipelined && parts == 4) {
    launch_kernel(attn_fwd_pipelined_kernel<false, 4>, dim3(grid), dim3(A3Layout<4>::kThreads), kA3SmemBytes,
                  static_cast<cudaStream_t>(stream), tq, tk, tv, p);
    ST_CHECK_LAUNCH("attn_fwd_pipelined_kernel");
  } else if (pipelined) {
    auto fn = attn_fwd_pipelined_kernel<false, 2, 0>;
    switch (trace ? -1 : g_attn_poly) {
      case -1: fn = attn_fwd_pipelined_kernel<true>; break;
      case 2: fn = attn_fwd_pipelined_kernel<false, 2, 2>; break;
      default: break;
    }
    if (ablate) {
      p.ablate = ablate;
      fn = trace ? attn_fwd_pipelined_kernel<true, 2, 0, true> : attn_fwd_pipelined_kernel<false, 2, 0, true>;
    }
    launch_kernel(fn, dim3(grid), dim3(kA3Threads), kA3SmemBytes, static_cast<cudaStream_t>(stream), tq, tk, tv, p);
    ST_CHECK_LAUNCH("attn_fwd_pipelined_kernel");
  } else if (short_kv) {
    launch_kernel(attn_fwd_short_kernel, dim3(grid), dim3(kShortThreads), kShortSmemBytes, static_cast<cudaStream_t>(stream),
                  tq, tk, tv, p);
    ST_CHECK_LAUNCH("attn_fwd_short_kernel");
  } else {
    launch_kernel(trace ? attn_fwd_kernel<true> : attn_fwd_kernel<false>, dim3(grid), dim3(kAttnThreads), kAttnSmemBytes,
                  static_cast<cudaStream_t>(stream), tq, tk, tv, p);
    ST_CHECK_LAUNCH("attn_fwd_kernel");
  }
  return ST_OK;
}

void st_debug_set_attention_trace(void* buf) { st::g_attn_trace = static_cast<unsigned long long*>(buf); }

// Debug / tuning hook: exp warps per TMEM lane quadrant of the pipelined kernel (2 or 4; 0 = back to the default).
void st_debug_set_attention_parts(int parts) { st::g_attn_parts = (parts == 2 || parts == 4) ? parts : 0; }

// Debug / test hook (no device needed): the launcher's choice for a K/V sweep of `tiles` query tiles over Tk keys on a
// machine of `sms` SMs -- 1: resident kernel, 0: pipelined kernel.
int st_debug_attention_prefers_resident(long long tiles, int Tk, int sms) {
  if (!(Tk > st::kAttnBlockKV && st::sweep_prefers_resident(tiles, Tk, sms))) return 0;
  return st::resident_four_ctas(tiles, sms) ? 4 : 1;  // 4: the four-CTA form
}

// Debug / test hook: force one of the kernels behind st_attention_bf16 (0 = by shape, 1 two-CTA, 2 pipelined, 3 short,
// 4 resident, 5 = by shape without the resident kernel; -1 = re-read ST_ATTN_IMPL).  Forcing a one-block kernel onto a longer K/V sweep is the caller's mistake.
void st_debug_set_attention_impl(int impl) { st::g_attn_impl = (impl >= 0 && impl <= 5) ? impl : -1; }

// Debug / tuning hook: element pairs out of every 8 whose exponential takes the FMA-pipe polynomial (0 or 2; -1 = default).
void st_debug_set_attention_poly(int pairs) { st::g_attn_poly = (pairs == 0 || pairs == 2) ? pairs : -1; }

// Debug hook: resident CTAs per SM the driver grants the attention kernel (2 expected).
int st_debug_attention_occupancy(void) {
  int n = -1;
  cudaFuncSetAttribute(st::attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st::kAttnSmemBytes);
  cudaFuncSetAttribute(st::attn_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, st::attn_fwd_kernel<false>, st::kAttnThreads, st::kAttnSmemBytes);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, st::attn_fwd_kernel<false>);
  int n48 = -1, n0 = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n48, st::attn_fwd_kernel<false>, st::kAttnThreads, 48 * 1024);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, st::attn_fwd_kernel<false>, st::kAttnThreads, 0);
  printf("attn kernel: regs %d, static smem %zu, max dyn smem %d, local %zu, occupancy @%d B: %d, @48K: %d, @0: %d\n",
         fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes, st::kAttnSmemBytes, n, n48, n0);
  cudaFuncAttributes ra;
  cudaFuncGetAttributes(&ra, st::attn_fwd_resident_kernel<3>);
  printf("resident kernel: regs %d, local %zu, dyn smem %d\n", ra.numRegs, ra.localSizeBytes, st::kResSmemBytes);
  return n;
}

}  // extern "C"
