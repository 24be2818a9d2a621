// Flash attention forward for head_dim 64 on sm_100a: TMA-fed tcgen05 tiles, S, P and the per-block
// P.V product all in TMEM, online softmax in fp32 registers (one thread per query row, so row max /
// sum need no shuffles).
//
//   grid = (ceil(Tq / 128), B * H); 320 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
//   warps 2..9 softmax / accumulate / store (two warps per TMEM lane quadrant, each owning half of every
//   row).  Two CTAs are resident per SM (<= 84 KB smem, 256 TMEM columns, <= 85 registers), so while one
//   CTA's softmax warps sit on the MUFU pipe the other CTA's MMAs keep the tensor pipe busy.
//
//   S = Q K^T        : tcgen05.mma SS (Q, K 128B-swizzled K-major tiles from TMA), 128 fp32 columns
//   P = 2^(S*c - m)  : written back to TMEM as packed bf16 (64 columns) with tcgen05.st
//   O_j = P V        : tcgen05.mma TS -- A operand straight from TMEM, V as an MN-major smem operand --
//                      ~4x cheaper than the SS form (a 128x64x16 SS MMA is bound by the smem A fetch)
//   O += O_j         : folded into fp32 registers one block late, rescaled by 2^(m_old - m_new)
//
// Semantics = the *pattern* of fuse_attention (reference: optimizers/replace_attention.py:76-86;
// fp32 restatement ref_attention :109-124): per head softmax(Q K^T * scale) V, no mask, separate
// Tq / Tk with a masked K tail (Tk = 77 cross-attention).  Replaces
// kernels/attention_fa2.py:17-140.
#include "common.cuh"
#include "ptx.cuh"

namespace st {

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
};

// Registers are allocated per group of 4 warps: the 10 warps of this CTA cost as much as 12, so two resident
// CTAs need <= 85 registers per thread -- hence the bound of 384 threads although 320 are launched.
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
#define AT_TRACE(slot)                                                                  \
  do {                                                                                  \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[slot] = clock64();       \
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
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 256);
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
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kAttnBlockKV, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);        // P (TMEM)    x V (MN-major)
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int st = j % kAttnStages;
        mbar_wait(&k_full[st], (j / kAttnStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * kAttnTileBytes);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_bf16_ss(tmem_S, umma_smem_desc_sw128(q_addr + k * 32, 0, 1024),
                       umma_smem_desc_sw128(k_addr + k * 32, 0, 1024), idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAttnStages;
        mbar_wait(p_full, j & 1);  // softmax(j) has consumed S(j) and published P(j)
        tc_fence_after();
        if (j == 2) AT_TRACE(8);
        if (j + 1 < nkv) issue_s(j + 1);
        mbar_wait(&v_full[st], (j / kAttnStages) & 1);
        if (j > 0) mbar_wait(o_empty, (j - 1) & 1);  // O_{j-1} has been folded into registers
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + st * kAttnTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnBlockKV / 16; ++kk)
          umma_bf16_ts(tmem_O, tmem_P + kk * 8, umma_smem_desc_sw128(v_addr + kk * 16 * 128, 8192, 1024), idesc_o,
                       kk != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(o_full);
        if (j == 2) AT_TRACE(9);
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
        mbar_arrive(o_empty);
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
      mbar_arrive(p_full);
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

// 4-D map over a bf16 tensor addressed as [b][h][t][d] with element strides (sb, sh, st, 1) and d = 64;
// box = [1, 1, 128, 64].  Covers both (B, T, H*64) activations (sh = 64, st = row pitch) and the
// reference's (B, H, T, D) layout (kernels/attention_fa2.py:113-140).  OOB rows (t >= T) read as zero.
static int make_tmap_bhtd(CUtensorMap* out, const void* base, int B, int H, int T, long long sb, long long sh,
                          long long st_) {
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
  cuuint32_t box[4] = {64, 128, 1, 1};
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

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static unsigned long long* g_attn_trace = nullptr;

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
  rc = make_tmap_bhtd(&tk, k, B, H, Tk, k_sb, k_sh, k_st);
  if (rc != ST_OK) return rc;
  rc = make_tmap_bhtd(&tv, v, B, H, Tk, v_sb, v_sh, v_st);
  if (rc != ST_OK) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ST_ERR_CUDA;
    }
    // two CTAs per SM need 2 x 82 KB: ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    configured = true;
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
  const dim3 grid((Tq + kAttnBlockQ - 1) / kAttnBlockQ, B * H);
  launch_kernel(attn_fwd_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmemBytes, static_cast<cudaStream_t>(stream), tq, tk, tv, p);
  ST_CHECK_LAUNCH("attn_fwd_kernel");
  return ST_OK;
}

void st_debug_set_attention_trace(void* buf) { st::g_attn_trace = static_cast<unsigned long long*>(buf); }

// Debug hook: resident CTAs per SM the driver grants the attention kernel (2 expected).
int st_debug_attention_occupancy(void) {
  int n = -1;
  cudaFuncSetAttribute(st::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st::kAttnSmemBytes);
  cudaFuncSetAttribute(st::attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, st::attn_fwd_kernel, st::kAttnThreads, st::kAttnSmemBytes);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, st::attn_fwd_kernel);
  int n48 = -1, n0 = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n48, st::attn_fwd_kernel, st::kAttnThreads, 48 * 1024);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, st::attn_fwd_kernel, st::kAttnThreads, 0);
  printf("attn kernel: regs %d, static smem %zu, max dyn smem %d, local %zu, occupancy @%d B: %d, @48K: %d, @0: %d\n",
         fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes, st::kAttnSmemBytes, n, n48, n0);
  return n;
}

}  // extern "C"
