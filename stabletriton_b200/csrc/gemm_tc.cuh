// Persistent, warp-specialised bf16 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[m, n] = epilogue( sum_k A[m, k] * B[n, k] )        A, B bf16 (K contiguous), fp32 accumulate in TMEM
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer (one elected
// lane), warps 2..5 = epilogue (TMEM -> registers -> bias / SiLU / GEGLU / residual -> bf16 -> HBM).
// Three pipelines: smem ring (full/empty mbarriers, TMA <-> MMA), two TMEM accumulator stages
// (tmem_full/tmem_empty, MMA <-> epilogue) and a static persistent tile schedule (tile = blockIdx.x +
// i * gridDim.x), so the epilogue of tile i overlaps the main loop of tile i+1.
//
// A is fetched either through a 2-D tensor map (plain GEMM: Linear, 1x1 conv, im2col'ed conv) or a
// 4-D NHWC tensor map (3x3 / pad 1 / stride 1 convolution): the 128 output pixels of a tile form a
// Wt x Ht rectangle of one image, and filter tap (r, s) is the same rectangle shifted by (r-1, s-1);
// TMA zero-fills the out-of-bounds halo, so no im2col buffer and no border branches exist.
//
// Replaces (reference): kernels/linear.py:69-222 (kernel_fma / sdxl_forward), kernels/geglu.py:17-35
// (fused here as an epilogue), kernels/Conv_Kernels/conv_implicit_gemm.py:12-182.
#pragma once
#include "ptx.cuh"

namespace st {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kGemmThreads = 192;

struct GemmParams {
  int M, N, K;          // N = number of B rows consumed (for GEGLU: 2 * n_out), K = reduction length
  int n_out;            // output columns (N, or N/2 for GEGLU)
  int ldd;              // output row stride (elements)
  int num_m_blocks, num_n_blocks;
  __nv_bfloat16* D;
  const __nv_bfloat16* bias;      // [N] or nullptr
  const __nv_bfloat16* residual;  // [M, ldr] or nullptr, added in fp32 before the single rounding
  int ldr;
  const __nv_bfloat16* rowbias;   // [M / rows_per_batch, ld_rowbias] or nullptr (time-embedding broadcast)
  int ld_rowbias;
  int rows_per_batch;
  int act_silu;  // apply SiLU after bias
  // 4-D (conv) A addressing
  int conv_H, conv_W, conv_C;  // input == output spatial size (3x3, pad 1, stride 1)
  int conv_Wt, conv_Ht;        // tile rectangle, Wt * Ht == 128
};

template <int BLOCK_N, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kGemmBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kTotal = STAGES * kStageBytes + kBarrierBytes + 1024;  // +1024: manual alignment slack
};

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// kConvA: A through the 4-D NHWC map.  kGeglu: B tile = [BLOCK_N/2 "state" rows | BLOCK_N/2 "gate" rows].
template <int BLOCK_N, int STAGES, bool kConvA, bool kGeglu>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using S = GemmSmem<BLOCK_N, STAGES>;
  constexpr int kAccCols = BLOCK_N;                       // fp32 accumulator columns per stage
  constexpr int kTmemCols = tmem_cols_for(2 * kAccCols);  // two accumulator stages
  static_assert(2 * kAccCols <= 512, "accumulator stages exceed TMEM");
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ab = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_k_blocks = p.K / kGemmBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = kConvA ? p.conv_C / kGemmBlockK : 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blocks;
        const int n_blk = tile / p.num_m_blocks;
        const int m0 = m_blk * kGemmBlockM;
        int img = 0, p0 = 0, q0 = 0;
        if (kConvA) {
          const int hw = p.conv_H * p.conv_W;
          img = m0 / hw;
          const int rem = m0 - img * hw;
          p0 = rem / p.conv_W;
          q0 = rem - p0 * p.conv_W;
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_ab + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          if (kConvA) {
            const int tap = kb / cblocks;
            const int cb = kb - tap * cblocks;
            const int r = tap / 3, s = tap - r * 3;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], cb * kGemmBlockK, q0 + s - 1, p0 + r - 1, img);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kGemmBlockK, m0);
          }
          if (kGeglu) {
            const int h0 = n_blk * (BLOCK_N / 2);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBlockK, h0);
            tma_load_2d(sb + S::kBBytes / 2, &tmap_b, &full_bar[stage], kb * kGemmBlockK, p.n_out + h0);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n_blk * BLOCK_N);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_ab + stage * S::kStageBytes);
          const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
          for (int k = 0; k < kGemmBlockK / 16; ++k) {
            const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int kOutCols = kGeglu ? BLOCK_N / 2 : BLOCK_N;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const int row = m_blk * kGemmBlockM + quad * 32 + lane;
      const int n0 = n_blk * kOutCols;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols;
      const bool row_ok = row < p.M;
      const __nv_bfloat16* res_row = p.residual ? p.residual + static_cast<size_t>(row) * p.ldr : nullptr;
      const __nv_bfloat16* rb_row =
          p.rowbias ? p.rowbias + static_cast<size_t>(row_ok ? row / p.rows_per_batch : 0) * p.ld_rowbias : nullptr;
      __nv_bfloat16* d_row = p.D + static_cast<size_t>(row) * p.ldd;
#pragma unroll 1
      for (int c = 0; c < kOutCols; c += 32) {
        uint32_t v[32];
        uint32_t g[32];
        tmem_ld_32x32b_x32(t_row + c, v);
        if (kGeglu) tmem_ld_32x32b_x32(t_row + BLOCK_N / 2 + c, g);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = n0 + c + j;
            if (col < p.n_out) {  // n_out is a multiple of 8 (checked on the host)
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[j + e]);
              if (p.bias) {
                const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + col);
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(bw[e]);
                  x[2 * e] += f.x;
                  x[2 * e + 1] += f.y;
                }
              }
              if (kGeglu) {
                float gt[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) gt[e] = __uint_as_float(g[j + e]);
                if (p.bias) {
                  const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + p.n_out + col);
                  const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16x2(bw[e]);
                    gt[2 * e] += f.x;
                    gt[2 * e + 1] += f.y;
                  }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] *= gelu_erf_f(gt[e]);
              }
              if (rb_row) {
                const uint4 bv = *reinterpret_cast<const uint4*>(rb_row + col);
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(bw[e]);
                  x[2 * e] += f.x;
                  x[2 * e + 1] += f.y;
                }
              }
              if (p.act_silu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = silu_f(x[e]);
              }
              if (res_row) {
                const uint4 rv = *reinterpret_cast<const uint4*>(res_row + col);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(rw[e]);
                  x[2 * e] += f.x;
                  x[2 * e + 1] += f.y;
                }
              }
              uint4 o;
              o.x = pack_bf16x2(x[0], x[1]);
              o.y = pack_bf16x2(x[2], x[3]);
              o.z = pack_bf16x2(x[4], x[5]);
              o.w = pack_bf16x2(x[6], x[7]);
              *reinterpret_cast<uint4*>(d_row + col) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace st
