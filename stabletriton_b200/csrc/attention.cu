// Flash attention forward for head_dim 64 on sm_100a: TMA-fed tcgen05 tiles, S and the per-block
// P.V product in TMEM, online softmax in fp32 registers (one thread per query row, so row max / sum
// need no shuffles), P re-staged to 128B-swizzled shared memory as the A operand of the second MMA.
//
//   grid = (ceil(Tq / 128), B * H); 192 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
//   warps 2..5 softmax / accumulate / store.
//   Pipelines: K/V smem ring (k_full, v_full, kv_empty), two S buffers (s_full) and two P buffers
//   (p_full) so QK^T of block j+1 runs under the softmax of block j, two PV buffers (o_full, o_empty)
//   whose result is folded into the register accumulator one block late.
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
constexpr int kAttnStages = 3;
constexpr int kAttnThreads = 192;
constexpr int kAttnTileBytes = 128 * 64 * 2;  // any [128 x 64] bf16 tile
constexpr int kAttnSmemBytes = kAttnTileBytes * (1 + 2 * kAttnStages + 4) + 1024 + 1024;

struct AttnParams {
  __nv_bfloat16* O;
  long long o_sb, o_sh, o_st;
  int H, Tq, Tk;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTileBytes;
  uint8_t* sV = sK + kAttnStages * kAttnTileBytes;
  uint8_t* sP = sV + kAttnStages * kAttnTileBytes;  // 2 buffers x [2 sub-tiles of 128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kAttnTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* v_full = k_full + kAttnStages;
  uint64_t* kv_empty = v_full + kAttnStages;
  uint64_t* s_full = kv_empty + kAttnStages;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y - b * p.H;
  const int nkv = (p.Tk + kAttnBlockKV - 1) / kAttnBlockKV;

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 64 columns

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
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kAttnD, 0, 1);        // P (K-major) x V (MN-major)
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int st = j % kAttnStages;
        mbar_wait(&k_full[st], (j / kAttnStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * kAttnTileBytes);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_bf16_ss(tmem_S + (j & 1) * kAttnBlockKV, umma_smem_desc_sw128(q_addr + k * 32, 0, 1024),
                       umma_smem_desc_sw128(k_addr + k * 32, 0, 1024), idesc_s, k != 0);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_s(j + 1);
        const int bf = j & 1;
        const uint32_t u = (j >> 1) & 1;
        const int st = j % kAttnStages;
        mbar_wait(&p_full[bf], u);
        mbar_wait(&o_empty[bf], u ^ 1);
        mbar_wait(&v_full[st], (j / kAttnStages) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + bf * 2 * kAttnTileBytes);
        const uint32_t v_addr = smem_u32(sV + st * kAttnTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnBlockKV / 16; ++kk) {
          const uint64_t da = umma_smem_desc_sw128(p_addr + (kk >> 2) * kAttnTileBytes + (kk & 3) * 32, 0, 1024);
          const uint64_t db = umma_smem_desc_sw128(v_addr + kk * 16 * 128, 8192, 1024);
          umma_bf16_ss(tmem_O + bf * kAttnD, da, db, idesc_o, kk != 0);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&o_full[bf]);
      }
    }
  } else {
    // ===================================== softmax / accumulate ===============================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float acc[kAttnD];
#pragma unroll
    for (int i = 0; i < kAttnD; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int bf = j & 1;
      const uint32_t u = (j >> 1) & 1;
      const int valid = p.Tk - j * kAttnBlockKV;  // columns >= valid are padding
      mbar_wait(&s_full[bf], u);
      tc_fence_after();
      const uint32_t s_addr = tmem_S + lane_off + bf * kAttnBlockKV;
      // pass 1: running maximum of the scaled scores
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < kAttnBlockKV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(v[i]) * p.scale_log2;
          mx = fmaxf(mx, (c + i < valid) ? s : -INFINITY);
        }
      }
      const float alpha = ex2_approx(m - mx);  // m = -inf on the first block -> 0
      // pass 2: p = 2^(s - max), row sum, bf16 P tile into swizzled smem
      float rowsum = 0.f;
      uint8_t* p_row = sP + bf * 2 * kAttnTileBytes + row * 128;
#pragma unroll 1
      for (int c = 0; c < kAttnBlockKV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_addr + c, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = fmaf(__uint_as_float(v[i]), p.scale_log2, -mx);
          e[i] = (c + i < valid) ? ex2_approx(s) : 0.f;
          rowsum += e[i];
        }
        uint8_t* sub = p_row + (c >> 6) * kAttnTileBytes;  // sub-tile of 64 keys
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c & 63) >> 3) + g;  // 16-byte chunk index within the 128-byte row
          uint4 o;
          o.x = pack_bf16x2(e[8 * g + 0], e[8 * g + 1]);
          o.y = pack_bf16x2(e[8 * g + 2], e[8 * g + 3]);
          o.z = pack_bf16x2(e[8 * g + 4], e[8 * g + 5]);
          o.w = pack_bf16x2(e[8 * g + 6], e[8 * g + 7]);
          *reinterpret_cast<uint4*>(sub + ((chunk ^ (row & 7)) << 4)) = o;
        }
      }
      l = fmaf(l, alpha, rowsum);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[bf]);

      // fold in the P.V product of the previous block (it was computed relative to the old max)
      if (j > 0) {
        const int pb = (j - 1) & 1;
        mbar_wait(&o_full[pb], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(tmem_O + lane_off + pb * kAttnD, o0);
        tmem_ld_32x32b_x32(tmem_O + lane_off + pb * kAttnD + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_empty[pb]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[i] = (acc[i] + __uint_as_float(o0[i])) * alpha;
          acc[32 + i] = (acc[32 + i] + __uint_as_float(o1[i])) * alpha;
        }
      }
      m = mx;
    }
    {
      const int pb = (nkv - 1) & 1;
      mbar_wait(&o_full[pb], ((nkv - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tmem_O + lane_off + pb * kAttnD, o0);
      tmem_ld_32x32b_x32(tmem_O + lane_off + pb * kAttnD + 32, o1);
      tmem_ld_wait();
      const float inv = 1.f / l;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        acc[i] = (acc[i] + __uint_as_float(o0[i])) * inv;
        acc[32 + i] = (acc[32 + i] + __uint_as_float(o1[i])) * inv;
      }
    }
    if (q0 + row < p.Tq) {
      __nv_bfloat16* orow = p.O + b * p.o_sb + h * p.o_sh + static_cast<long long>(q0 + row) * p.o_st;
#pragma unroll
      for (int i = 0; i < kAttnD; i += 8) {
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
  if (warp == 1) tmem_dealloc<512>(tmem_base);
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
  const dim3 grid((Tq + kAttnBlockQ - 1) / kAttnBlockQ, B * H);
  attn_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
  ST_CHECK_LAUNCH("attn_fwd_kernel");
  return ST_OK;
}

}  // extern "C"
