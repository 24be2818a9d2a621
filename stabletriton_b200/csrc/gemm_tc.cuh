// Persistent, warp-specialised bf16 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[m, n] = epilogue( sum_k A[m, k] * B[n, k] )        A, B bf16 (K contiguous), fp32 accumulate in TMEM
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer (one elected
// lane), warps 2..9 = epilogue (TMEM -> registers -> bias / SiLU / GEGLU / residual -> bf16 -> HBM).
// Three pipelines: smem ring (full/empty mbarriers, TMA <-> MMA), two TMEM accumulator stages
// (tmem_full/tmem_empty, MMA <-> epilogue) and a static persistent tile schedule (tile = blockIdx.x +
// i * gridDim.x), so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Measured on B200 (profiles/r01_gemm_trace.txt, tools/mmabench.cu): issued from a warp-convergent loop (see
// umma_bf16_ss_elect) a 128 x N x 16 tcgen05.mma runs at N/2 cycles, so a 128 x 256 k-block needs 512 cycles of
// tensor pipe and 48 KB of operand fill; at 148 resident CTAs the fill (~9.7 KB/clk out of L2 for all SMs) is the
// bound (664 cycles).  The epilogue therefore has to stay off the critical path: the per-tile bias (+ per-image
// time-embedding row) is staged once in shared memory, the residual rows are prefetched into registers *before*
// the accumulator is waited for, and the bf16 result leaves through a swizzled staging tile and TMA bulk stores
// (64 columns x 128 rows each, clipped at the matrix edge by the tensor map), so no global-memory round trip and
// no uncoalesced store sits between TMEM and HBM.
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
constexpr int kGemmEpiThreads = 256;              // 8 epilogue warps
constexpr int kGemmThreads = 64 + kGemmEpiThreads;  // + TMA producer warp + MMA issuer warp

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
  unsigned long long* trace;  // debug: 12 clock64 stamps per CTA (see ST_TRACE), or nullptr
  // stream-K (small grids with a long K loop): the tiles x k-blocks space is cut into gridDim.x equal
  // contiguous ranges; a CTA that starts in the middle of a tile dumps its fp32 partial into
  // ws[blockIdx.x] and raises flags[blockIdx.x]; the CTA that owns the head of the tile merges them.
  int stream_k;
  int w_static;  // W is not written by the preceding kernel: prefetch it ahead of the PDL wait
  int cluster;   // host-side choice: launch the CTA-pair instantiation
  int out_f32;   // D is fp32 [M, ldd] (ST_EPI_F32OUT): written with plain 16-byte stores, one 128-byte line per thread and
                 // 32-column chunk -- for results whose bf16 rounding would be amplified downstream (attention scores
                 // of the 512-wide VAE head ahead of a row softmax)
  int spin_wait; // bit 0: the MMA warp polls full_bar (mbarrier.test_wait) instead of try_wait, bit 1: the producer polls
                 // empty_bar -- a suspended waiter is woken ~300 cycles late, and in a ring that is latency-bound
                 // (CTA pairs: 7 stages against a ~2850-cycle round trip) both waits block on every k-block
  int res_tma;   // the residual tile of a CTA's LAST output tile arrives by TMA in the idle operand ring (host-side choice:
                 // bf16 output, no GEGLU, no stream-K) instead of through per-thread row loads -- see the producer warp
  int drain_all; // debug (ST_GEMM_DRAIN=1): wait for the bulk stores' global writes before exit, not just their smem reads
  float* ws;            // [gridDim.x][128][BLOCK_N] fp32
  unsigned* flags;      // [gridDim.x], zero between launches (self-resetting)
  // GroupNorm statistics of the OUTPUT, emitted by the epilogue (the consumer GroupNorm then needs no statistics pass
  // over the tensor): gn_part[m_blk][n_out][2] = (mean, M2) of each output column over the 128 rows of tile row-block
  // m_blk, computed from the bf16-rounded values that are stored.  Requires M % 128 == 0; plain mode only.
  float* gn_part;
  // 4-D (conv) A addressing
  int conv_H, conv_W, conv_C;  // input == output spatial size (3x3, pad 1, stride 1)
  int conv_Wt, conv_Ht;        // tile rectangle, Wt * Ht == 128
};

// kHalfB: CTA-pair mode, each CTA stages only BLOCK_N / 2 rows of the weight tile
template <int BLOCK_N, int STAGES, bool kHalfB = false>
struct GemmSmem {
  static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
  static constexpr int kBBytes = (kHalfB ? BLOCK_N / 2 : BLOCK_N) * kGemmBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutStageBytes = kGemmBlockM * 64 * 2;  // epilogue staging tile for the TMA store
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kBiasBytes = 2 * BLOCK_N * 4;  // two accumulator stages x BLOCK_N fp32
  static constexpr int kGnBytes = 2 * 4 * 64 * 2 * 4;  // [2 group parities][4 lane quadrants][64 columns][mean, M2]
  static constexpr int kTotal =
      STAGES * kStageBytes + kOutStageBytes + kBarrierBytes + kBiasBytes + kGnBytes + 1024;  // +1024: alignment slack
};

__host__ __device__ constexpr int tmem_cols_for(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

__device__ __forceinline__ void add_bf16x8(float (&x)[8], const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16x2(w[e]);
    x[2 * e] += f.x;
    x[2 * e + 1] += f.y;
  }
}

// Column sums over the 32 rows a warp holds (one row per lane, 32 columns per lane): a transposing butterfly -- at
// every step a lane keeps the half of its values whose column bit matches its lane bit and hands the other half to
// its partner -- 31 shuffles instead of the 160 of a per-column butterfly.  On return lane l holds in v[0] the sum of
// column l over all 32 lanes, accumulated in a fixed order (bit-reproducible).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// kConvA: A through the 4-D NHWC map.  kGeglu: B tile = [BLOCK_N/2 "state" rows | BLOCK_N/2 "gate" rows].
// kCluster: CTA pairs (2-CTA clusters, tcgen05 cta_group::2) on vertically adjacent tiles -- same n-block, m-blocks
// 2i and 2i+1.  One MMA instruction, issued by the even ("leader") CTA, drives both tensor cores on a 256 x
// BLOCK_N tile; every CTA stages its own 128 rows of A and only HALF of the weight tile (BLOCK_N/2 rows), and
// each accumulates its 128 output rows in its own TMEM.  Operand fill per CTA and k-block drops from 48 KB to
// 32 KB (BLOCK_N = 256): at 148 resident CTAs the one-CTA main loop is bound by the ~9.7 KB/clk the L2 can feed
// all SMs together (757 cycles per k-block against 512 of tensor pipe; profiles/r01_gemm_trace.txt).  A plain
// TMA multicast of the weight tile (both CTAs still ingest all of it) was measured slower than no cluster.
template <int BLOCK_N, int STAGES, bool kConvA, bool kGeglu, bool kStreamK = false, bool kCluster = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d_tail,
                    const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_r_tail,
                    const GemmParams p) {
  using S = GemmSmem<BLOCK_N, STAGES, kCluster>;
  constexpr int kAccCols = BLOCK_N;                       // fp32 accumulator columns per stage
  constexpr int kTmemCols = tmem_cols_for(2 * kAccCols);  // two accumulator stages
  static_assert(2 * kAccCols <= 512, "accumulator stages exceed TMEM");
  static_assert(!(kStreamK && kCluster), "stream-K and cluster multicast are exclusive");
  // Tile widths that are not a multiple of 64 (160: eight tiles across N = 1280, so 128 CTAs instead of 112 carry the
  // 2048 x 1280 outputs) end in a 32-column group, stored through tmap_d_tail (un-swizzled 32-column boxes).
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(!kStreamK || BLOCK_N % 64 == 0, "stream-K assumes whole 64-column groups");
  static_assert(!kGeglu || BLOCK_N % 128 == 0, "the TMA store works on 64-column groups of the output tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_ab = smem;
  uint8_t* s_out = smem + STAGES * S::kStageBytes;  // 1024-byte aligned: stages are multiples of 8 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_out + S::kOutStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;  // residual tile of the last output tile has landed in the ring
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  float* s_bias = reinterpret_cast<float*>(s_out + S::kOutStageBytes + S::kBarrierBytes);  // [2][BLOCK_N]
  float* s_gn = s_bias + 2 * BLOCK_N;                                                      // [2][4][64][2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_k_blocks = p.K / kGemmBlockK;
  // Work iterator: plain mode = whole tiles blockIdx.x, +gridDim.x, ...; stream-K = one contiguous range of
  // (tile, k-block) units.  A segment is (tile, [kb0, kb1)).
  const uint32_t cta_rank = kCluster ? cluster_ctarank() : 0;
  const long long sk_units = static_cast<long long>(num_tiles) * num_k_blocks;
  const int sk_u0 = kStreamK ? static_cast<int>(sk_units * blockIdx.x / gridDim.x) : 0;
  const int sk_u1 = kStreamK ? static_cast<int>(sk_units * (blockIdx.x + 1) / gridDim.x) : 0;
  auto next_segment = [&](int& cursor, int& tile, int& kb0, int& kb1) -> bool {
    if (kStreamK) {
      if (cursor >= sk_u1) return false;
      tile = cursor / num_k_blocks;
      kb0 = cursor - tile * num_k_blocks;
      kb1 = min(num_k_blocks, kb0 + (sk_u1 - cursor));
      cursor += kb1 - kb0;
    } else if (kCluster) {
      // the cluster walks pair-tiles (two m-blocks x one n-block); this CTA takes m-block 2*mp + rank
      const int half_mb = p.num_m_blocks >> 1;
      if (cursor >= half_mb * p.num_n_blocks) return false;
      const int nb = cursor / half_mb;
      tile = nb * p.num_m_blocks + 2 * (cursor - nb * half_mb) + static_cast<int>(cta_rank);
      kb0 = 0;
      kb1 = num_k_blocks;
      cursor += gridDim.x >> 1;
    } else {
      if (cursor >= num_tiles) return false;
      tile = cursor;
      kb0 = 0;
      kb1 = num_k_blocks;
      cursor += gridDim.x;
    }
    return true;
  };
  const int cursor0 = kStreamK ? sk_u0 : static_cast<int>(kCluster ? blockIdx.x >> 1 : blockIdx.x);
#define ST_TRACE(slot)                                                      \
  do {                                                                      \
    if (p.trace) p.trace[blockIdx.x * 12 + (slot)] = clock64();              \
  } while (0)
  if (threadIdx.x == 0) ST_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    if (BLOCK_N % 64 != 0) tma_prefetch_desc(&tmap_d_tail);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (kCluster ? 2 : 1) * kGemmEpiThreads / 32);  // pair: both CTAs' epilogue warps
    }
    mbar_init(res_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (kCluster)
      tmem_alloc_pair<kTmemCols>(tmem_slot);  // same warp in both CTAs
    else
      tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (kCluster) cluster_sync();  // the leader's barriers exist before the peer's first TMA load signals them
  // everything above touched only shared / tensor memory: it may overlap the previous kernel's tail
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();
  if (threadIdx.x == 32) ST_TRACE(1);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      // B tile of k-block kb / n-block n_blk into ring slot `sb`.  Pair mode: this CTA fetches only its half of the
      // rows (GEGLU: rank 0 the "state" rows, rank 1 the "gate" rows); completion is signalled to the leader.
      constexpr uint32_t kMyStageBytes = S::kStageBytes;
      auto arm = [&](uint64_t* bar) {  // pair: the leader arms its barrier for both CTAs' bytes
        if (!kCluster)
          mbar_expect_tx(bar, kMyStageBytes);
        else if (cta_rank == 0)
          mbar_expect_tx(bar, 2 * kMyStageBytes);
      };
      auto load_b = [&](uint8_t* sb, uint64_t* bar, int kb, int n_blk) {
        if (kCluster) {
          const int r0 = kGeglu ? (cta_rank ? p.n_out : 0) + n_blk * (BLOCK_N / 2)
                                : n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / 2);
          tma_load_2d_pair(sb, &tmap_b, leader_addr(bar), kb * kGemmBlockK, r0);
        } else if (kGeglu) {
          const int h0 = n_blk * (BLOCK_N / 2);
          tma_load_2d(sb, &tmap_b, bar, kb * kGemmBlockK, h0);
          tma_load_2d(sb + S::kBBytes / 2, &tmap_b, bar, kb * kGemmBlockK, p.n_out + h0);
        } else {
          tma_load_2d(sb, &tmap_b, bar, kb * kGemmBlockK, n_blk * BLOCK_N);
        }
      };
      // The weight operand is not produced by the preceding kernel (p.w_static): start its first tiles
      // towards shared memory BEFORE the programmatic dependency resolves, so the HBM latency of the
      // weights (never L2-resident: 5 GB stream per step) hides under the previous kernel's tail.
      int prefetched = 0;
      if (p.w_static) {
        int cursor = cursor0, tile, kb0, kb1;
        if (next_segment(cursor, tile, kb0, kb1)) {
          const int n_blk = tile / p.num_m_blocks;
          prefetched = min(STAGES, kb1 - kb0);
          for (int i = 0; i < prefetched; ++i) {
            uint8_t* sb = smem_ab + i * S::kStageBytes + S::kABytes;
            const int kb = kb0 + i;
            arm(&full_bar[i]);
            load_b(sb, &full_bar[i], kb, n_blk);
          }
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = kConvA ? p.conv_C / kGemmBlockK : 1;
      int cursor = cursor0, tile, kb0, kb1;
      int last_tile = -1;
      while (next_segment(cursor, tile, kb0, kb1)) {
        last_tile = tile;
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
        for (int kb = kb0; kb < kb1; ++kb) {
          uint8_t* sa = smem_ab + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          const bool b_in_flight = prefetched > 0;  // first ring pass of the first segment
          if (b_in_flight) {
            --prefetched;
          } else {
            if (p.spin_wait & 2)
              mbar_wait_spin(&empty_bar[stage], phase ^ 1);
            else
              mbar_wait(&empty_bar[stage], phase ^ 1);
            arm(&full_bar[stage]);
          }
          if (kConvA) {
            const int tap = kb / cblocks;
            const int cb = kb - tap * cblocks;
            const int r = tap / 3, s = tap - r * 3;
            if (kCluster)
              tma_load_4d_pair(sa, &tmap_a, leader_addr(&full_bar[stage]), cb * kGemmBlockK, q0 + s - 1, p0 + r - 1, img);
            else
              tma_load_4d(sa, &tmap_a, &full_bar[stage], cb * kGemmBlockK, q0 + s - 1, p0 + r - 1, img);
          } else if (kCluster) {
            tma_load_2d_pair(sa, &tmap_a, leader_addr(&full_bar[stage]), kb * kGemmBlockK, m0);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kGemmBlockK, m0);
          }
          if (!b_in_flight) load_b(sb, &full_bar[stage], kb, n_blk);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      // Residual operand of this CTA's LAST tile (for the single-wave GEMMs -- out-projections and FF2 of every
      // transformer block -- the only tile): no operand load follows, so the ring slots fall idle one by one while the
      // last k-blocks are multiplied.  The residual tile goes, 64-column group by group, exactly where the epilogue will
      // stage the output of that group (same 128-byte swizzle), as soon as the slots underneath have been released; the
      // epilogue reads a chunk from there, adds, and writes the rounded result back in place.  The per-thread row loads
      // this replaces (16 bytes per lane and request, 32 sectors per request) cost the main loop of the 2048 x 1280 x
      // 1280 GEMM 2.8 k of its 9.7 k cycles: they compete with the operand stream for the SM's ingest.
      constexpr int kOutColsP = kGeglu ? BLOCK_N / 2 : BLOCK_N;
      constexpr int kGroupsP = (kOutColsP + 63) / 64;
      constexpr bool kTailP = kOutColsP % 64 != 0;
      if (!kStreamK && !kGeglu && p.res_tma && last_tile >= 0 &&
          kGroupsP * S::kOutStageBytes <= STAGES * S::kStageBytes) {
        constexpr int kResSlots = (kGroupsP * S::kOutStageBytes + S::kStageBytes - 1) / S::kStageBytes;
        for (int sl = 0; sl < kResSlots; ++sl)  // as if the slot were loaded again: its last MMAs have retired
          mbar_wait(&empty_bar[sl], (sl >= stage ? phase : phase ^ 1) ^ 1);
        const int m0 = (last_tile % p.num_m_blocks) * kGemmBlockM;
        const int n0 = (last_tile / p.num_m_blocks) * kOutColsP;
        uint32_t bytes = 0;
        for (int g = 0; g < kGroupsP; ++g)
          if (n0 + g * 64 < p.n_out) bytes += (kTailP && g == kGroupsP - 1) ? kGemmBlockM * 32 * 2 : S::kOutStageBytes;
        mbar_expect_tx(res_bar, bytes);
        for (int g = 0; g < kGroupsP; ++g)
          if (n0 + g * 64 < p.n_out)
            tma_load_2d(smem_ab + g * S::kOutStageBytes, (kTailP && g == kGroupsP - 1) ? &tmap_r_tail : &tmap_r, res_bar,
                        n0 + g * 64, m0);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The whole warp runs this loop (see umma_bf16_ss_elect in ptx.cuh): one elected lane issues, operand
    // descriptors advance by 64-bit adds from a per-kernel base.
    if (!kCluster || cta_rank == 0) {  // pair: the leader issues for both tensor cores
      constexpr uint32_t idesc = umma_idesc_bf16(kCluster ? 2 * kGemmBlockM : kGemmBlockM, BLOCK_N, 0, 0);
      const uint64_t desc_a0 = umma_smem_desc_sw128(smem_u32(smem_ab), 0, 1024);
      const uint64_t desc_b0 = umma_smem_desc_sw128(smem_u32(smem_ab) + S::kABytes, 0, 1024);
      constexpr uint32_t kStageStep = S::kStageBytes >> 4;  // descriptor address field counts 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int cursor = cursor0, tile, kb0, kb1;
      bool first_seg = true;
      while (next_segment(cursor, tile, kb0, kb1)) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.spin_wait & 1)
            mbar_wait_spin(&full_bar[stage], phase);
          else
            mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (first_seg && kb == kb0 && lane == 0) ST_TRACE(2);
          const uint64_t da = desc_a0 + static_cast<uint64_t>(stage * kStageStep);
          const uint64_t db = desc_b0 + static_cast<uint64_t>(stage * kStageStep);
#pragma unroll
          for (int k = 0; k < kGemmBlockK / 16; ++k)  // +32 bytes of K per instruction = +2 address units
            if (kCluster)
              umma_bf16_ss_pair_elect(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
            else
              umma_bf16_ss_elect(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
          // smem slot reusable once these MMAs retire (pair: in both CTAs' rings)
          if (kCluster)
            umma_commit_pair_elect(&empty_bar[stage], 0x3);
          else
            umma_commit_elect(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (kCluster)  // accumulator complete -> both CTAs' epilogues
          umma_commit_pair_elect(&tmem_full[acc], 0x3);
        else
          umma_commit_elect(&tmem_full[acc]);
        if (first_seg && lane == 0) ST_TRACE(3);
        first_seg = false;
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    // 8 warps: warp w may touch TMEM lane quadrant (w & 3); the two warps of a quadrant split every
    // 64-column group of the tile (columns 0-31 / 32-63), which doubles the number of independent
    // instruction streams per scheduler -- one warp per scheduler leaves the dependent FADD/convert/
    // pack chain latency-bound (measured: ~1000 cycles per 32-column chunk).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;        // 0: columns [0,32) of each group, 1: columns [32,64)
    const int etid = threadIdx.x - 64;       // 0..255 among the epilogue threads
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int kOutCols = kGeglu ? BLOCK_N / 2 : BLOCK_N;
    constexpr int kGroups = (kOutCols + 63) / 64;   // 64-column groups = TMA store boxes
    constexpr bool kTail = kOutCols % 64 != 0;       // the last group is only 32 columns wide
    // the time-embedding row can be folded into the staged bias when a tile lies inside one image
    const bool rowbias_per_tile = p.rowbias != nullptr && p.rows_per_batch >= kGemmBlockM;
    const bool rowbias_per_row = p.rowbias != nullptr && !rowbias_per_tile;
    const int tile_row = quad * 32 + lane;
    int cursor = cursor0, tile, kb0, kb1;
    bool first_seg = true;
    while (next_segment(cursor, tile, kb0, kb1)) {
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const int row = m_blk * kGemmBlockM + tile_row;
      const int n0 = n_blk * kOutCols;
      const bool row_ok = row < p.M;
      // Last tile of this CTA (for the single-wave GEMMs: the only one): no operand load is in flight or will be
      // issued any more and every MMA has retired once the accumulator barrier fires, so the smem ring is free --
      // each 64-column group gets its own 16 KB staging tile there instead of sharing one.  That removes, per
      // group, the wait for the previous bulk store to finish reading the tile and one of the two 256-thread
      // barriers, on the only epilogue that is not hidden behind a following main loop.
      // (CTA pairs: the accumulator barrier is the pair-wide commit of the last MMA, so BOTH rings are idle by then.)
      const bool last_tile = kCluster ? cursor >= (p.num_m_blocks >> 1) * p.num_n_blocks : cursor >= num_tiles;
      const bool ring_staging = !kStreamK && last_tile && kGroups * S::kOutStageBytes <= STAGES * S::kStageBytes;
      float* sb = s_bias + acc * BLOCK_N;
      const uint32_t t_row0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols;

      if (kStreamK && kb0 > 0) {
        // ---- stream-K contributor: this CTA entered the tile in the middle of its K range.  Dump the raw
        // fp32 partial into its workspace slot and publish it; the tile's owner applies the epilogue.
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        float* wrow = p.ws + (static_cast<size_t>(blockIdx.x) * kGemmBlockM + tile_row) * BLOCK_N;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const int c = g * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row0 + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<uint4*>(wrow + c + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        tc_fence_before();
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (etid == 0) {  // one flag per 128-byte line: spinners of other tiles must not slow this store down
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + blockIdx.x * 32), "r"(1u) : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        first_seg = false;
        continue;
      }
      // stream-K owner of a tile whose tail was computed by the following CTAs (their first segment)
      const bool sk_merge = kStreamK && kb1 < num_k_blocks;

      // (1) stage this tile's bias (+ per-image row bias) as fp32: sb[0..kOutCols) state / plain,
      //     sb[kOutCols..2*kOutCols) gate.  Buffer `acc` was last read two tiles ago by these same threads.
      {
        const __nv_bfloat16* rb =
            rowbias_per_tile ? p.rowbias + static_cast<size_t>((m_blk * kGemmBlockM) / p.rows_per_batch) * p.ld_rowbias
                             : nullptr;
        for (int c = etid; c < BLOCK_N; c += kGemmEpiThreads) {
          const int hb = kGeglu ? c / kOutCols : 0;
          const int col = n0 + (kGeglu ? c - hb * kOutCols : c);
          float b = 0.f;
          if (col < p.n_out) {
            if (p.bias) b = __bfloat162float(p.bias[hb * p.n_out + col]);
            if (rb) b += __bfloat162float(rb[col]);
          }
          sb[c] = b;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }

      // (2) prefetch this thread's residual columns into registers while the main loop is still running
      uint4 res[kGroups * 4];
      const bool has_res = p.residual != nullptr && row_ok;
      const bool res_smem = !kGeglu && ring_staging && p.res_tma != 0;  // the producer warp fetches the tile into the ring
      if (has_res && !res_smem) {
        const __nv_bfloat16* res_row = p.residual + static_cast<size_t>(row) * p.ldr + n0 + half * 32;
#pragma unroll
        for (int g = 0; g < kGroups; ++g)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (n0 + g * 64 + half * 32 + v * 8 < p.n_out) res[g * 4 + v] = ld_global_nc_v4_early(res_row + g * 64 + v * 8);
      }
      const __nv_bfloat16* rb_row =
          rowbias_per_row ? p.rowbias + static_cast<size_t>(row_ok ? row / p.rows_per_batch : 0) * p.ld_rowbias : nullptr;

      // (3) accumulator ready
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (res_smem && p.residual != nullptr) mbar_wait(res_bar, 0);
      if (first_seg && warp == 2 && lane == 0) ST_TRACE(4);
      const uint32_t t_row = t_row0;
      int sk_last = blockIdx.x;  // contributors are CTAs blockIdx.x + 1 .. sk_last
      if (sk_merge) {
        const long long tile_end = static_cast<long long>(tile + 1) * num_k_blocks;
        while (sk_last + 1 < static_cast<int>(gridDim.x) &&
               sk_units * (sk_last + 1) / gridDim.x < tile_end)
          ++sk_last;
        if (etid == 0) {
          for (int cc = blockIdx.x + 1; cc <= sk_last; ++cc) {
            unsigned seen = 0;
            while (true) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + cc * 32) : "memory");
              if (seen) break;
              __nanosleep(64);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (n0 + g * 64 >= p.n_out) break;  // this 64-column group lies entirely past the matrix edge
        // 32-column tail group: the warps of the upper half have no columns of their own; they shadow the lower
        // half's (same barriers, no stores), which keeps the group loop free of divergent synchronisation
        const bool tail = kTail && g == kGroups - 1;
        const bool dead = tail && half == 1;
        const int c = g * 64 + (dead ? 0 : half * 32);   // first tile column of this thread's chunk
        uint32_t v[32];
        uint32_t gt[32];
        tmem_ld_32x32b_x32(t_row + c, v);
        if (kGeglu) tmem_ld_32x32b_x32(t_row + kOutCols + c, gt);
        tmem_ld_wait();
        if (sk_merge) {
          for (int cc = blockIdx.x + 1; cc <= sk_last; ++cc) {
            const float* wrow = p.ws + (static_cast<size_t>(cc) * kGemmBlockM + tile_row) * BLOCK_N + c;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 w4 = __ldcg(reinterpret_cast<const float4*>(wrow + i));
              v[i + 0] = __float_as_uint(__uint_as_float(v[i + 0]) + w4.x);
              v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + w4.y);
              v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + w4.z);
              v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + w4.w);
            }
          }
        }
        if (g == 0 && first_seg && warp == 2 && lane == 0) ST_TRACE(7);
        // the staging tile is reused per group: wait until the previous TMA store has read it
        uint8_t* s_stage = ring_staging ? smem_ab + g * S::kOutStageBytes : s_out;
        if (!ring_staging) {
          if (etid == 0) tma_store_wait_read();
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        if (g == 0 && first_seg && warp == 2 && lane == 0) ST_TRACE(8);
        // All shared-memory reads of this chunk first (the compiler cannot hoist them over the staging
        // stores below: both live in shared memory), then straight-line register math on 32 columns.
        uint4 rs[4];  // residual chunk from the staging tile the producer warp filled (columns past n_out hold TMA's zeros)
        if (res_smem && has_res) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4)
            rs[v4] = *reinterpret_cast<const uint4*>(
                tail ? s_stage + tile_row * 64 + (v4 << 4)
                     : s_stage + tile_row * 128 + (((half * 4 + v4) ^ (tile_row & 7)) << 4));
        }
        float x[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(sb + c + 4 * q);
          x[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b.x;
          x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b.y;
          x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b.z;
          x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b.w;
        }
        if (kGeglu) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(sb + kOutCols + c + 4 * q);
            x[4 * q + 0] *= gelu_erf_f(__uint_as_float(gt[4 * q + 0]) + b.x);
            x[4 * q + 1] *= gelu_erf_f(__uint_as_float(gt[4 * q + 1]) + b.y);
            x[4 * q + 2] *= gelu_erf_f(__uint_as_float(gt[4 * q + 2]) + b.z);
            x[4 * q + 3] *= gelu_erf_f(__uint_as_float(gt[4 * q + 3]) + b.w);
          }
        }
        if (rb_row) {
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (n0 + c + j < p.n_out) {
              float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              add_bf16x8(t, *reinterpret_cast<const uint4*>(rb_row + n0 + c + j));
#pragma unroll
              for (int e = 0; e < 8; ++e) x[j + e] += t[e];
            }
        }
        if (p.act_silu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) x[e] = silu_f(x[e]);
        }
        if (has_res && res_smem) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint32_t w[4] = {rs[j / 8].x, rs[j / 8].y, rs[j / 8].z, rs[j / 8].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[j + 2 * e] += __uint_as_float(w[e] << 16);
              x[j + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
            }
          }
        } else if (has_res) {
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (n0 + c + j < p.n_out) {
              const uint4 r4 = res[g * 4 + j / 8];
              const uint32_t w[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w[e]);
                x[j + 2 * e] += f.x;
                x[j + 2 * e + 1] += f.y;
              }
            }
        }
        if (p.out_f32) {  // uniform over the CTA: nobody stages, nobody waits on the staging barriers below
          if (row_ok && !dead) {
            float* drow = reinterpret_cast<float*>(p.D) + static_cast<size_t>(row) * p.ldd + n0 + c;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + c + j < p.n_out) *reinterpret_cast<float4*>(drow + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          }
          continue;
        }
        // staging tile: [128 rows][64 cols] bf16, 128-byte swizzle (16-byte chunk ^= row & 7)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o;
          o.x = pack_bf16x2(x[j + 0], x[j + 1]);
          o.y = pack_bf16x2(x[j + 2], x[j + 3]);
          o.z = pack_bf16x2(x[j + 4], x[j + 5]);
          o.w = pack_bf16x2(x[j + 6], x[j + 7]);
          const int chunk = half * 4 + (j >> 3);
          if (tail) {  // un-swizzled [128 rows][32 cols]
            if (!dead) *reinterpret_cast<uint4*>(s_stage + tile_row * 64 + ((j >> 3) << 4)) = o;
          } else {
            *reinterpret_cast<uint4*>(s_stage + tile_row * 128 + ((chunk ^ (tile_row & 7)) << 4)) = o;
          }
          if (!kGeglu && !kStreamK && p.gn_part) {  // keep the ROUNDED values: the statistics are those of the stored tensor
            const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[j + 2 * e] = __uint_as_float(w4[e] << 16);
              x[j + 2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
            }
          }
        }
        float* gn_slot = s_gn + (g & 1) * (4 * 64 * 2);
        if (!kGeglu && !kStreamK && p.gn_part) {
          // (mean, M2) of my 32 columns over this warp's 32 rows -> shared memory; merged over the 4 lane quadrants
          // after the barrier below (no extra synchronisation: see the buffer parity)
          float sq[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) sq[e] = x[e] * x[e];
          const float s1 = warp_column_sums(x, lane);
          const float s2 = warp_column_sums(sq, lane);
          const float mean = s1 * (1.f / 32.f);
          const float m2 = fmaxf(s2 - s1 * mean, 0.f);
          if (!dead) *reinterpret_cast<float2*>(gn_slot + (quad * 64 + half * 32 + lane) * 2) = make_float2(mean, m2);
        }
        if (g == 0 && first_seg && warp == 2 && lane == 0) ST_TRACE(9);
        // 64 columns staged: hand them to the TMA store engine (clips rows >= M and columns >= n_out)
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (etid == 0) {
          tma_store_2d(tail ? &tmap_d_tail : &tmap_d, s_stage, n0 + g * 64, m_blk * kGemmBlockM);
          tma_store_commit();
        }
        if (!kGeglu && !kStreamK && p.gn_part && etid >= 64 && etid < 128) {
          // one thread per column: merge the four 32-row partials (equal counts), fixed order
          const int col = etid - 64;
          const int gcol = n0 + g * 64 + col;
          if (gcol < p.n_out && !(tail && col >= 32)) {
            float2 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const float2*>(gn_slot + (i * 64 + col) * 2);
            const float mean = ((q[0].x + q[1].x) + (q[2].x + q[3].x)) * 0.25f;
            float m2 = (q[0].y + q[1].y) + (q[2].y + q[3].y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float d = q[i].x - mean;
              m2 = fmaf(32.f * d, d, m2);
            }
            *reinterpret_cast<float2*>(p.gn_part + (static_cast<size_t>(m_blk) * p.n_out + gcol) * 2) =
                make_float2(mean, m2);
          }
        }
        if (g == 0 && first_seg && warp == 2 && lane == 0) ST_TRACE(11);
      }
      if (sk_merge) {  // every partial has been consumed: re-arm the contributors' flags for the next launch
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (etid == 0)
          for (int cc = blockIdx.x + 1; cc <= sk_last; ++cc) p.flags[cc * 32] = 0u;
      }
      tc_fence_before();
      __syncwarp();
      if (first_seg && warp == 2 && lane == 0) ST_TRACE(5);
      first_seg = false;
      if (lane == 0) {
        if (kCluster)
          mbar_arrive_cluster(leader_addr(&tmem_empty[acc]));  // the leader's MMA warp owns the accumulator hand-off
        else
          mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // Shared memory must outlive the last bulk store's READ of the staging tile; the global writes themselves are ordinary
    // in-flight stores that grid completion (and the dependent kernel's griddepcontrol.wait) covers -- waiting for them
    // here kept every CTA ~1.5 k cycles longer on the exit path of the single-wave launches (p.drain_all: A/B switch).
    if (etid == 0) {
      if (p.drain_all)
        tma_store_wait_all();
      else
        tma_store_wait_read();
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (kCluster) {
    // the leader's MMAs read the peer's ring and its commits signal the peer's barriers: nobody leaves (or frees
    // tensor memory) before both CTAs are done
    cluster_sync();
    if (warp == 1) tmem_dealloc_pair<kTmemCols>(tmem_base);
  } else if (warp == 1) {
    tmem_dealloc<kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) ST_TRACE(6);
#undef ST_TRACE
}

}  // namespace st
