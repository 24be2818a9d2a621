// Bandwidth-bound normalisation kernels: GroupNorm(+SiLU) on NHWC and LayerNorm over the last dim.
// 128-bit coalesced loads/stores, fp32 statistics (per-CTA shifted sums in registers and shared
// memory, warp-shuffle reductions, Chan-style merge of the per-CTA (n, mean, M2) triples), one HBM
// read + one write of the activation (the second GroupNorm read is served by the 126 MB L2).
//
// Replaces (reference): kernels/groupnorm.py:24-161 (semantics fixed to torch.nn.GroupNorm on 4-D
// input) and kernels/layer_norm.py:114-346.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <atomic>

namespace st {

__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm
//   workspace layout (floats):  partial[N][G][3][chunks] (n, mean, M2)  |  scale_shift[N][C][2]
//   kernel 1 (stats): every CTA reduces a pixel range of one image for all channels -- fully coalesced
//     16-byte loads -- into per-channel shifted sums (shift = the CTA's first pixel, so |mean| >> std
//     cannot cancel), folds them into per-group (n, mean, M2) and publishes them; the last CTA of an
//     image to arrive (atomic ticket) merges the chunks and writes per-channel scale/shift.
//   kernel 2 (apply): y = silu?(x * scale_c + shift_c); the second read of x is served by the L2.
// ------------------------------------------------------------------------------------------------
constexpr int kGnMaxThreads = 320;   // >= C/8 for C <= 2560; small CTAs so that 3 fit on an SM (<= 68 registers)
constexpr int kGnMaxImages = 4096;
constexpr int kGnStatsPixPerThread = 8;  // 8 x 16 B in flight per thread
constexpr int kGnApplyPixPerThread = 6;
// Per-image arrival tickets live in the CALLER's workspace (one word per image, zeroed by gn_ticket_zero_kernel at the
// head of every call): two GroupNorm calls that run concurrently -- different streams, parallel graph branches, two
// captured graphs replayed at once -- necessarily own different workspaces, so they can never see each other's
// tickets (a library-global table, however it is indexed, cannot promise that), and a faulted launch leaves nothing
// behind.
// (Measured and rejected, r02: processing a large batch in L2-sized image groups -- statistics, then apply, per 48 MB
// group, so that the apply pass finds its input in L2 -- makes every launch latency-bound again: N = 16 at
// (320, 128^2) 122 -> 164 us.  One round whatever N is.)
constexpr size_t kGnL2WindowBytes = ~static_cast<size_t>(0);

struct GnGeom {
  int N, HW, C, G, cpg;
  int vecs;             // C / 8 : 16-byte vectors per pixel
  int threads;          // multiple of vecs, <= 512
  int pix_lanes;        // threads / vecs
  int chunks;           // stats CTAs per image
  int pix_per_chunk;    // ceil(HW / chunks)
  int a_chunks;         // apply CTAs per image
  int a_pix_per_chunk;
};

static GnGeom gn_geometry(int N, int HW, int C, int G) {
  GnGeom g;
  g.N = N;
  g.HW = HW;
  g.C = C;
  g.G = G;
  g.cpg = C / G;
  g.vecs = C / 8;
  g.pix_lanes = kGnMaxThreads / g.vecs;
  if (g.pix_lanes < 1) g.pix_lanes = 1;
  g.threads = g.pix_lanes * g.vecs;
  const int sms = device_sm_count();
  // Every thread issues ALL of its loads before it consumes any (one memory round trip per CTA instead
  // of one per loop iteration), so a CTA covers at most kGn*PixPerThread x pix_lanes pixels; subject to
  // that bound the pixel range is split into ~2 (stats) / ~4 (apply) CTAs per SM.
  auto split = [&](int ctas_total, int max_per_thread, int* chunks, int* per) {
    int c = (ctas_total + N - 1) / N;
    const int min_c = (HW + max_per_thread * g.pix_lanes - 1) / (max_per_thread * g.pix_lanes);
    const int max_c = (HW + g.pix_lanes - 1) / g.pix_lanes;  // at least one pixel per thread
    if (c > max_c) c = max_c;
    if (c < min_c) c = min_c;
    if (c < 1) c = 1;
    *per = (HW + c - 1) / c;
    *chunks = (HW + *per - 1) / *per;
  };
  split(3 * sms, kGnStatsPixPerThread, &g.chunks, &g.pix_per_chunk);
  split(3 * sms, kGnApplyPixPerThread, &g.a_chunks, &g.a_pix_per_chunk);
  return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kGnMaxThreads, 3)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial,
                const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                float* __restrict__ scale_shift, int HW, int C, int G, int cpg, int vecs, int pix_lanes,
                int pix_per_chunk, float eps, unsigned int* __restrict__ tickets) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_gn[];  // [C] shift | [C] sum(x - shift) | [C] sum((x - shift)^2) | per-thread partials
  float* s_shift = s_gn;
  float* s_sum = s_gn + C;
  float* s_sq = s_gn + 2 * C;
  float* s_part = s_gn + 3 * C;    // [pix_lanes][2][C]: fixed-order (deterministic) cross-thread reduction
  __shared__ int s_last;
  const int n = blockIdx.y;
  const int chunk = blockIdx.x;
  const int chunks = gridDim.x;
  const int cv = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int pb = chunk * pix_per_chunk;
  const int pe = min(pb + pix_per_chunk, HW);
  const __nv_bfloat16* base = x + (static_cast<size_t>(n) * HW) * C + cv * 8;

  uint4 u[kGnStatsPixPerThread];
  const uint4 u_shift = ld_nc_16(base + static_cast<size_t>(pb) * C);  // same address for every pixel lane
#pragma unroll
  for (int i = 0; i < kGnStatsPixPerThread; ++i) {
    const int pp = pb + pl + i * pix_lanes;
    if (pp < pe) u[i] = ld_nc_16(base + static_cast<size_t>(pp) * C);
  }
  float shift[8], s1[8], s2[8];
  unpack8(u_shift, shift);
  if (pl == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_shift[cv * 8 + e] = shift[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s1[e] = 0.f;
    s2[e] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < kGnStatsPixPerThread; ++i) {
    const int pp = pb + pl + i * pix_lanes;
    if (pp < pe) {
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[e] - shift[e];
        s1[e] += d;
        s2[e] = fmaf(d, d, s2[e]);
      }
    }
  }
  {
    float* mine = s_part + static_cast<size_t>(pl) * 2 * C + cv * 8;
    *reinterpret_cast<float4*>(mine) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    *reinterpret_cast<float4*>(mine + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
    *reinterpret_cast<float4*>(mine + C) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    *reinterpret_cast<float4*>(mine + C + 4) = make_float4(s2[4], s2[5], s2[6], s2[7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < pix_lanes; ++l) {
      a += s_part[static_cast<size_t>(l) * 2 * C + c];
      q += s_part[static_cast<size_t>(l) * 2 * C + C + c];
    }
    s_sum[c] = a;
    s_sq[c] = q;
  }
  __syncthreads();

  // per-group (n, mean, M2) of this chunk: one warp per group, lanes over the group's channels
  // blockDim.x = pix_lanes * C/8 need not be a multiple of 32: only full warps take part in the shuffles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const float cnt = static_cast<float>(pe - pb);
  for (int g = warp < warps ? warp : G; g < G; g += warps) {
    float tot = 0.f;
    for (int c = lane; c < cpg; c += 32) tot += s_sum[g * cpg + c] + cnt * s_shift[g * cpg + c];
    tot = warp_sum(tot);
    const float mean = tot / (cnt * cpg);
    float m2 = 0.f;
    for (int c = lane; c < cpg; c += 32) {
      const float d = mean - s_shift[g * cpg + c];
      m2 += s_sq[g * cpg + c] - 2.f * d * s_sum[g * cpg + c] + cnt * d * d;
    }
    m2 = warp_sum(m2);
    if (lane == 0) {
      // layout [n][g][3][chunks]: the finalising warp reads each quantity of a group as one coalesced run
      float* out = partial + ((static_cast<size_t>(n) * G + g) * 3) * chunks + chunk;
      out[0] = cnt * cpg;
      out[chunks] = mean;
      out[2 * chunks] = fmaxf(m2, 0.f);
    }
  }

  // ticket: the last CTA of image n merges all chunks
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[n], 1u) == static_cast<unsigned>(chunks - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int g = warp < warps ? warp : G; g < G; g += warps) {
    const float* q = partial + ((static_cast<size_t>(n) * G + g) * 3) * chunks;
    // every lane fetches ALL of its chunk triples up front and keeps them: one L2 round trip per group (the
    // first version made two -- means, then M2 -- and this serial tail was the longest part of the kernel)
    constexpr int kPer = 8;  // 256 chunks per pass; gn_geometry never makes more than ~3 x SMs / N
    float nsum = 0.f, wsum = 0.f, m2 = 0.f;
    if (chunks <= 32 * kPer) {
      float nk[kPer], mk[kPer], qk[kPer];
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const int k = lane + 32 * i;
        nk[i] = k < chunks ? __ldcg(q + k) : 0.f;
        mk[i] = k < chunks ? __ldcg(q + chunks + k) : 0.f;
        qk[i] = k < chunks ? __ldcg(q + 2 * chunks + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        nsum += nk[i];
        wsum = fmaf(nk[i], mk[i], wsum);
      }
      nsum = warp_sum(nsum);
      wsum = warp_sum(wsum);
      const float mean0 = wsum / nsum;
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const float d = mk[i] - mean0;  // padded lanes: nk = qk = 0, so d does not matter
        m2 += qk[i] + nk[i] * d * d;
      }
    } else {
      for (int k = lane; k < chunks; k += 32) {
        const float nk = __ldcg(q + k);
        nsum += nk;
        wsum = fmaf(nk, __ldcg(q + chunks + k), wsum);
      }
      nsum = warp_sum(nsum);
      wsum = warp_sum(wsum);
      const float mean0 = wsum / nsum;
      for (int k = lane; k < chunks; k += 32) {
        const float d = __ldcg(q + chunks + k) - mean0;
        m2 += __ldcg(q + 2 * chunks + k) + __ldcg(q + k) * d * d;
      }
    }
    const float mean = wsum / nsum;
    m2 = warp_sum(m2);
    const float rstd = rsqrtf(m2 / nsum + eps);  // biased variance, as torch.nn.GroupNorm
    for (int c = lane; c < cpg; c += 32) {
      const int ch = g * cpg + c;
      const float ga = gamma ? __bfloat162float(gamma[ch]) : 1.f;
      const float be = beta ? __bfloat162float(beta[ch]) : 0.f;
      const float sc = ga * rstd;
      float* o = scale_shift + (static_cast<size_t>(n) * C + ch) * 2;
      o[0] = sc;
      o[1] = be - mean * sc;
    }
  }
  if (threadIdx.x == 0) tickets[n] = 0u;
}

__global__ void __launch_bounds__(256)
gn_ticket_zero_kernel(unsigned int* __restrict__ tickets, int n) {
  pdl_launch_dependents();
  pdl_wait();  // the workspace may be recycled memory that the preceding kernel still reads
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) tickets[i] = 0u;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm from producer-emitted partial statistics.  The GEMM / implicit-GEMM conv epilogue that WRITES the
// activation also writes, per 128-row tile and output column, (mean, M2) of what it stored (gemm_tc.cuh: gn_part);
// a tile lies inside one image (HW % 128 == 0), so image n owns tiles [n * HW/128, (n+1) * HW/128).  This kernel
// merges them per (image, group) -- every partial has the same count (128), so the merge is mean = avg(mean_i),
// M2 = sum(M2_i) + 128 * sum((mean_i - mean)^2), evaluated in a fixed order -- and writes per-channel scale / shift;
// gn_apply_kernel then streams the activation ONCE.  A channel concatenation (up-block skip connections) never has
// to be reduced again: channels [0, C_a) come from the first producer's partials, [C_a, C) from the second's.
// One CTA of 128 threads per (image, group); a single pass over the partials with every load independent of the
// previous one (the first version walked them with one dependent L2 round trip per entry and took 20 us):
// shift = the first partial mean of the group, d_i = mean_i - shift, then
//     mean = shift + sum(d_i) / k,      M2 = sum(M2_i) + 128 * (sum(d_i^2) - sum(d_i)^2 / k)
// which cannot cancel (|d_i| is of the order of the spread of the tile means, not of |mean|).
// ------------------------------------------------------------------------------------------------
constexpr int kGnFinThreads = 128;
__global__ void __launch_bounds__(kGnFinThreads)
gn_finalize_kernel(const float* __restrict__ part_a, int C_a, const float* __restrict__ part_b, int C_b,
                   const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                   float* __restrict__ scale_shift, int tiles, int C, int G, int cpg, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_red[3][kGnFinThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int n = blockIdx.y;
  const int entries = tiles * cpg;
  auto entry = [&](int idx) -> float2 {
    const int t = idx / cpg;
    const int ch = g * cpg + (idx - t * cpg);
    const size_t tile = static_cast<size_t>(n) * tiles + t;
    const float* src = ch < C_a ? part_a + (tile * C_a + ch) * 2 : part_b + (tile * C_b + (ch - C_a)) * 2;
    return __ldcg(reinterpret_cast<const float2*>(src));
  };
  const float shift = entry(0).x;
  float sd = 0.f, sdd = 0.f, sm2 = 0.f;
  constexpr int kBatch = 8;  // independent loads in flight per thread
  for (int base = threadIdx.x; base < entries; base += kBatch * kGnFinThreads) {
    float2 e[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int i = base + j * kGnFinThreads;
      e[j] = i < entries ? entry(i) : make_float2(shift, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const float d = e[j].x - shift;
      sd += d;
      sdd = fmaf(d, d, sdd);
      sm2 += e[j].y;
    }
  }
  sd = warp_sum(sd);
  sdd = warp_sum(sdd);
  sm2 = warp_sum(sm2);
  if (lane == 0) {
    s_red[0][warp] = sd;
    s_red[1][warp] = sdd;
    s_red[2][warp] = sm2;
  }
  __syncthreads();
  float tsd = 0.f, tsdd = 0.f, tsm2 = 0.f;
#pragma unroll
  for (int w = 0; w < kGnFinThreads / 32; ++w) {  // fixed order: bit-reproducible
    tsd += s_red[0][w];
    tsdd += s_red[1][w];
    tsm2 += s_red[2][w];
  }
  const float k = static_cast<float>(entries);
  const float mean = shift + tsd / k;
  const float m2 = tsm2 + 128.f * fmaxf(tsdd - tsd * tsd / k, 0.f);
  const float rstd = rsqrtf(m2 / (128.f * k) + eps);  // biased variance, as torch.nn.GroupNorm
  for (int c = threadIdx.x; c < cpg; c += kGnFinThreads) {
    const int ch = g * cpg + c;
    const float ga = gamma ? __bfloat162float(gamma[ch]) : 1.f;
    const float be = beta ? __bfloat162float(beta[ch]) : 0.f;
    const float sc = ga * rstd;
    float* o = scale_shift + (static_cast<size_t>(n) * C + ch) * 2;
    o[0] = sc;
    o[1] = be - mean * sc;
  }
}

// Pass 3: y = silu?(x * scale + shift), streaming.
template <bool kSilu>
__global__ void __launch_bounds__(kGnMaxThreads, 3)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                const float* __restrict__ scale_shift, int HW, int C, int vecs, int pix_lanes, int pix_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const int cv = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int pb = blockIdx.x * pix_per_chunk;
  const int pe = min(pb + pix_per_chunk, HW);
  float sc[8], sh[8];
  {
    const float4* q = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(n) * C + cv * 8) * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = q[i];
      sc[2 * i] = t.x;
      sh[2 * i] = t.y;
      sc[2 * i + 1] = t.z;
      sh[2 * i + 1] = t.w;
    }
  }
  const size_t img = (static_cast<size_t>(n) * HW) * C + cv * 8;
  const __nv_bfloat16* xb = x + img;
  __nv_bfloat16* yb = y + img;
  uint4 u[kGnApplyPixPerThread];
#pragma unroll
  for (int i = 0; i < kGnApplyPixPerThread; ++i) {
    const int pp = pb + pl + i * pix_lanes;
    if (pp < pe) u[i] = ld_nc_16(xb + static_cast<size_t>(pp) * C);
  }
#pragma unroll
  for (int i = 0; i < kGnApplyPixPerThread; ++i) {
    const int pp = pb + pl + i * pix_lanes;
    if (pp < pe) {
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = fmaf(f[e], sc[e], sh[e]);
        if (kSilu) v = silu_f(v);
        f[e] = v;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(yb + static_cast<size_t>(pp) * C) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (two-pass statistics, exact).
// ------------------------------------------------------------------------------------------------
template <int kVecsPerLane, bool kParamsFirst>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                 const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, int M, int N,
                 float eps) {
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = N >> 3;
  // gamma / beta are parameters, not outputs of the preceding kernel: fetch them BEFORE the programmatic dependency
  // resolves, so that only one memory round trip (the row itself) is left on the exposed path of this latency-bound
  // kernel (in sequence a LayerNorm launch costs ~5 us, profiles/r01_timeline_v1.txt).  Kept packed: 8 registers / vector.
  // (wide rows: not enough registers; many rows: the kernel is bandwidth-bound and the 40 extra registers would cost a
  // third of the resident warps -- 65536 x 1280: 5.2 TB/s without, 4.5 TB/s with -- so gamma / beta are loaded at their use)
  constexpr bool kPrefetch = kParamsFirst && kVecsPerLane <= 5;
  uint4 gv[kPrefetch ? kVecsPerLane : 1], bv[kPrefetch ? kVecsPerLane : 1];
#pragma unroll
  for (int i = 0; i < (kPrefetch ? kVecsPerLane : 0); ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      gv[i] = __ldg(reinterpret_cast<const uint4*>(gamma + v * 8));
      bv[i] = beta ? __ldg(reinterpret_cast<const uint4*>(beta + v * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * ldx;
  float f[kVecsPerLane][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVecsPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      unpack8(ld_nc_16(xr + v * 8), f[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += f[i][e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / N;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kVecsPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[i][e] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / N + eps);
  __nv_bfloat16* yr = y + static_cast<size_t>(row) * ldy;
#pragma unroll
  for (int i = 0; i < kVecsPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float g[8], b[8];
      if (kPrefetch) {
        unpack8(gv[kPrefetch ? i : 0], g);
        unpack8(bv[kPrefetch ? i : 0], b);
      } else {
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + v * 8)), g);
        unpack8(beta ? __ldg(reinterpret_cast<const uint4*>(beta + v * 8)) : make_uint4(0u, 0u, 0u, 0u), b);
      }
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((f[i][e] - mean) * rstd, g[e], b[e]);
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]);
      u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + v * 8) = u;
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------
// Row softmax of fp32 scores into bf16 probabilities: P[i, :] = softmax(scale * S[i, :]).  One CTA per row; rows of up to
// 16 384 columns are held in registers (one read, one write), longer ones are streamed twice (online max / sum, then the
// write pass; the second read is served by L2).  Used by the single-head, 512-wide attention of the VAE mid block, whose
// scores come out of st_gemm_bf16(..., ST_EPI_F32OUT): no flash kernel covers head_dim 512, and a bf16 score matrix would
// put 2^-8 relative rounding under an exponential.
// ------------------------------------------------------------------------------------------------
constexpr int kSmThreads = 256;

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kSmThreads / 32; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kSmThreads / 32; ++w) r += red[w];  // fixed order: bit-reproducible
  __syncthreads();
  return r;
}

template <int kVec>  // float4 vectors per thread held in registers; 0: streaming
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_kernel(const float* __restrict__ S, long long lds, __nv_bfloat16* __restrict__ P, long long ldp, int N,
                    float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[kSmThreads / 32];
  const float4* srow = reinterpret_cast<const float4*>(S + static_cast<size_t>(blockIdx.x) * lds);
  uint2* prow = reinterpret_cast<uint2*>(P + static_cast<size_t>(blockIdx.x) * ldp);
  const int nvec = N >> 2;
  if (kVec > 0) {
    float4 v[kVec > 0 ? kVec : 1];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int i = threadIdx.x + k * kSmThreads;
      v[k] = i < nvec ? __ldcs(srow + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      mx = fmaxf(fmaxf(mx, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
    mx = block_max(mx, red);
    const float off = mx * scale_log2;  // scale > 0: max commutes with it
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      v[k].x = ex2_approx(fmaf(v[k].x, scale_log2, -off));
      v[k].y = ex2_approx(fmaf(v[k].y, scale_log2, -off));
      v[k].z = ex2_approx(fmaf(v[k].z, scale_log2, -off));
      v[k].w = ex2_approx(fmaf(v[k].w, scale_log2, -off));
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float inv = 1.f / block_sum(sum, red);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int i = threadIdx.x + k * kSmThreads;
      if (i < nvec) prow[i] = make_uint2(pack_bf16x2(v[k].x * inv, v[k].y * inv), pack_bf16x2(v[k].z * inv, v[k].w * inv));
    }
  } else {
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nvec; i += kSmThreads) {
      const float4 q = srow[i];
      mx = fmaxf(fmaxf(mx, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
    }
    mx = block_max(mx, red);
    const float off = mx * scale_log2;
    float sum = 0.f;
    for (int i = threadIdx.x; i < nvec; i += kSmThreads) {
      const float4 q = srow[i];
      sum += (ex2_approx(fmaf(q.x, scale_log2, -off)) + ex2_approx(fmaf(q.y, scale_log2, -off))) +
             (ex2_approx(fmaf(q.z, scale_log2, -off)) + ex2_approx(fmaf(q.w, scale_log2, -off)));
    }
    const float inv = 1.f / block_sum(sum, red);
    for (int i = threadIdx.x; i < nvec; i += kSmThreads) {
      const float4 q = srow[i];
      prow[i] = make_uint2(pack_bf16x2(ex2_approx(fmaf(q.x, scale_log2, -off)) * inv, ex2_approx(fmaf(q.y, scale_log2, -off)) * inv),
                           pack_bf16x2(ex2_approx(fmaf(q.z, scale_log2, -off)) * inv, ex2_approx(fmaf(q.w, scale_log2, -off)) * inv));
    }
  }
}

}  // namespace st

extern "C" {

// workspace layout: [tickets: N words, padded to 16 B] [partial: per image group] [scale_shift: N * C * 2 floats]
static size_t gn_ticket_floats(int N) { return (static_cast<size_t>(N) + 3) / 4 * 4; }

static int gn_images_per_round(int N, int HW, int C) {
  const size_t per_image = static_cast<size_t>(HW) * C * 2;
  size_t n = st::kGnL2WindowBytes / (per_image ? per_image : 1);
  if (n < 1) n = 1;
  return n < static_cast<size_t>(N) ? static_cast<int>(n) : N;
}

size_t st_groupnorm_workspace_bytes(int N, int HW, int C, int groups) {
  if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups != 0 || C % 8 != 0) return 0;
  const int per_round = gn_images_per_round(N, HW, C);
  const st::GnGeom g = st::gn_geometry(per_round, HW, C, groups);
  const size_t partial = (static_cast<size_t>(N) * g.chunks * groups * 3 + 3) / 4 * 4;
  const size_t ss = static_cast<size_t>(N) * C * 2;
  return (gn_ticket_floats(N) + partial + ss) * sizeof(float);
}

static int gn_check_common(const void* x, const void* y, const void* workspace, int N, int HW, int C, int groups) {
  using namespace st;
  ST_CHECK_ARG(x && y && workspace, "groupnorm: null pointer");
  ST_CHECK_ARG(N > 0 && HW > 0 && C > 0 && groups > 0, "groupnorm: sizes must be positive");
  ST_CHECK_ARG(N <= kGnMaxImages, "groupnorm: at most %d images per call", kGnMaxImages);
  ST_CHECK_ARG(C % groups == 0, "groupnorm: C (%d) not divisible by groups (%d)", C, groups);
  ST_CHECK_ARG(C % 8 == 0, "groupnorm: C (%d) must be a multiple of 8", C);
  ST_CHECK_ARG(C / 8 <= kGnMaxThreads, "groupnorm: C (%d) too large (max %d)", C, 8 * kGnMaxThreads);
  ST_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(workspace), "groupnorm: pointers must be 16-byte aligned");
  return ST_OK;
}

static int gn_launch_apply(const void* x, void* y, const float* scale_shift, int n_images, int HW, int C,
                           const st::GnGeom& g, int apply_silu, cudaStream_t s) {
  using namespace st;
  const dim3 agrid(g.a_chunks, n_images);
  if (apply_silu)
    launch_kernel(gn_apply_kernel<true>, agrid, dim3(g.threads), 0, s, static_cast<const __nv_bfloat16*>(x),
                  static_cast<__nv_bfloat16*>(y), scale_shift, HW, C, g.vecs, g.pix_lanes, g.a_pix_per_chunk);
  else
    launch_kernel(gn_apply_kernel<false>, agrid, dim3(g.threads), 0, s, static_cast<const __nv_bfloat16*>(x),
                  static_cast<__nv_bfloat16*>(y), scale_shift, HW, C, g.vecs, g.pix_lanes, g.a_pix_per_chunk);
  ST_CHECK_LAUNCH("gn_apply_kernel");
  return ST_OK;
}

int st_groupnorm_nhwc_bf16(const void* x, void* y, const void* gamma, const void* beta, void* workspace, int N,
                           int HW, int C, int groups, float eps, int apply_silu, st_stream_t stream) {
  using namespace st;
  int rc = gn_check_common(x, y, workspace, N, HW, C, groups);
  if (rc != ST_OK) return rc;
  ST_CHECK_ARG(C / groups >= 7 || C / groups == 4,
               "groupnorm: %d channels per group unsupported (a 16-byte vector may span at most two groups)", C / groups);
  const int per_round = gn_images_per_round(N, HW, C);
  const GnGeom g = gn_geometry(per_round, HW, C, groups);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* tickets = static_cast<unsigned int*>(workspace);
  float* partial = static_cast<float*>(workspace) + gn_ticket_floats(N);
  const size_t partial_elems = (static_cast<size_t>(N) * g.chunks * groups * 3 + 3) / 4 * 4;
  float* scale_shift = partial + partial_elems;

  const size_t smem = (static_cast<size_t>(3) * C + static_cast<size_t>(g.pix_lanes) * 2 * C) * sizeof(float);
  static PerDeviceOnce configured;  // function attributes are per context
  const int dev = current_device();
  ST_CHECK_ARG(dev >= 0, "groupnorm: device ordinal outside [0, %d)", kMaxDevices);
  if (!configured.done(dev)) {
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    configured.mark(dev);
  }
  ST_CHECK_ARG(smem <= 96 * 1024, "groupnorm: C (%d) needs too much shared memory", C);
  launch_kernel(gn_ticket_zero_kernel, dim3((N + 255) / 256), dim3(256), 0, s, tickets, N);
  ST_CHECK_LAUNCH("gn_ticket_zero_kernel");
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  for (int n0 = 0; n0 < N; n0 += per_round) {  // one round unless the activation exceeds the L2 window
    const int nr = N - n0 < per_round ? N - n0 : per_round;
    const size_t off = static_cast<size_t>(n0) * HW * C;
    launch_kernel(gn_stats_kernel, dim3(g.chunks, nr), dim3(g.threads), smem, s, xp + off,
                  partial + static_cast<size_t>(n0) * g.chunks * groups * 3, static_cast<const __nv_bfloat16*>(gamma),
                  static_cast<const __nv_bfloat16*>(beta), scale_shift + static_cast<size_t>(n0) * C * 2, HW, C, groups,
                  g.cpg, g.vecs, g.pix_lanes, g.pix_per_chunk, eps, tickets + n0);
    ST_CHECK_LAUNCH("gn_stats_kernel");
    rc = gn_launch_apply(xp + off, yp + off, scale_shift + static_cast<size_t>(n0) * C * 2, nr, HW, C, g, apply_silu, s);
    if (rc != ST_OK) return rc;
  }
  return ST_OK;
}

int st_groupnorm_from_partials_nhwc_bf16(const void* x, void* y, const void* gamma, const void* beta, void* workspace,
                                         int N, int HW, int C, int groups, float eps, int apply_silu,
                                         const void* part_a, int C_a, const void* part_b, int C_b,
                                         st_stream_t stream) {
  using namespace st;
  int rc = gn_check_common(x, y, workspace, N, HW, C, groups);
  if (rc != ST_OK) return rc;
  ST_CHECK_ARG(HW % 128 == 0, "groupnorm_from_partials: H*W (%d) must be a multiple of the 128-row producer tile", HW);
  ST_CHECK_ARG(part_a && C_a > 0 && C_b >= 0 && C_a + C_b == C && (C_b == 0 || part_b),
               "groupnorm_from_partials: channel split %d + %d does not match C = %d", C_a, C_b, C);
  ST_CHECK_ARG(aligned16(part_a) && (!part_b || aligned16(part_b)) && C_a % 2 == 0,
               "groupnorm_from_partials: partial buffers must be 16-byte aligned");
  const GnGeom g = gn_geometry(N, HW, C, groups);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* scale_shift = static_cast<float*>(workspace);  // N * C * 2 floats (st_groupnorm_workspace_bytes covers it)
  launch_kernel(gn_finalize_kernel, dim3(groups, N), dim3(kGnFinThreads), 0, s, static_cast<const float*>(part_a), C_a,
                static_cast<const float*>(part_b), C_b, static_cast<const __nv_bfloat16*>(gamma),
                static_cast<const __nv_bfloat16*>(beta), scale_shift, HW / 128, C, groups, g.cpg, eps);
  ST_CHECK_LAUNCH("gn_finalize_kernel");
  return gn_launch_apply(x, y, scale_shift, N, HW, C, g, apply_silu, s);
}

int st_layernorm_bf16(const void* x, int ldx, void* y, int ldy, const void* gamma, const void* beta, int M, int N,
                      float eps, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && y && gamma, "layernorm: null pointer");
  ST_CHECK_ARG(M > 0 && N > 0, "layernorm: sizes must be positive");
  ST_CHECK_ARG(N % 8 == 0 && N <= 4096, "layernorm: N (%d) must be a multiple of 8 and <= 4096", N);
  ST_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= N && ldy >= N, "layernorm: bad row pitch");
  ST_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && (!beta || aligned16(beta)),
               "layernorm: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_cta = 8;
  const int grid = (M + rows_per_cta - 1) / rows_per_cta;
  const int vpl = (N / 8 + 31) / 32;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(beta);
  // latency-bound launches (a few CTAs per SM: every LayerNorm of the SDXL step at CFG batch 2) fetch gamma / beta first
  const bool params_first = grid <= 8 * device_sm_count();
#define ST_LN_CASE(V)                                                                                               \
  case V:                                                                                                           \
    if (params_first)                                                                                               \
      launch_kernel(layernorm_kernel<V, true>, dim3(grid), dim3(256), 0, s, xp, ldx, yp, ldy, gp, bp, M, N, eps);   \
    else                                                                                                            \
      launch_kernel(layernorm_kernel<V, false>, dim3(grid), dim3(256), 0, s, xp, ldx, yp, ldy, gp, bp, M, N, eps);  \
    break;
  switch (vpl <= 3 ? 3 : vpl <= 5 ? 5 : vpl <= 8 ? 8 : 16) {
    ST_LN_CASE(3)
    ST_LN_CASE(5)
    ST_LN_CASE(8)
    ST_LN_CASE(16)
  }
#undef ST_LN_CASE
  ST_CHECK_LAUNCH("layernorm_kernel");
  return ST_OK;
}

int st_softmax_rows_f32_bf16(const float* S, long long lds, void* P, long long ldp, int M, int N, float scale,
                             st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(S && P, "softmax_rows: null pointer");
  ST_CHECK_ARG(M > 0 && N > 0 && N % 4 == 0, "softmax_rows: N (%d) must be a positive multiple of 4", N);
  ST_CHECK_ARG(scale > 0.f, "softmax_rows: scale must be positive");
  ST_CHECK_ARG(lds >= N && lds % 4 == 0 && ldp >= N && ldp % 4 == 0, "softmax_rows: bad row pitch");
  ST_CHECK_ARG(aligned16(S) && (reinterpret_cast<uintptr_t>(P) & 7) == 0, "softmax_rows: S must be 16-byte, P 8-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* pp = static_cast<__nv_bfloat16*>(P);
  const float sl2 = scale * 1.4426950408889634f;
  const int nvec = N / 4;
  if (nvec <= 4 * kSmThreads)
    launch_kernel(softmax_rows_kernel<4>, dim3(M), dim3(kSmThreads), 0, s, S, lds, pp, ldp, N, sl2);
  else if (nvec <= 16 * kSmThreads)
    launch_kernel(softmax_rows_kernel<16>, dim3(M), dim3(kSmThreads), 0, s, S, lds, pp, ldp, N, sl2);
  else
    launch_kernel(softmax_rows_kernel<0>, dim3(M), dim3(kSmThreads), 0, s, S, lds, pp, ldp, N, sl2);
  ST_CHECK_LAUNCH("softmax_rows_kernel");
  return ST_OK;
}

}  // extern "C"
