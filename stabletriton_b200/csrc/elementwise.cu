// CUDA-core kernels around the tensor-core path: standalone GEGLU, tiny-M Linear (embedding MLPs),
// small-channel direct 3x3 conv (conv_in / conv_out), im2col for strided conv, nearest 2x upsample,
// channel concat, sinusoidal timestep embedding, and the Euler + classifier-free-guidance update.
// All are HBM/L2-bandwidth or latency bound; 128-bit accesses throughout.
#include "common.cuh"
#include "ptx.cuh"

namespace st {

__device__ __forceinline__ void unpack8e(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8e(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ---- GEGLU (reference: kernels/geglu.py:17-26) ---------------------------------------------------
__global__ void geglu_kernel(const __nv_bfloat16* __restrict__ state, int lds, const __nv_bfloat16* __restrict__ gate,
                             int ldg, __nv_bfloat16* __restrict__ out, int ldo, int rows, int vec_per_row) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vec_per_row);
    const int c = static_cast<int>(i - static_cast<long long>(r) * vec_per_row) * 8;
    float s[8], g[8];
    unpack8e(*reinterpret_cast<const uint4*>(state + static_cast<size_t>(r) * lds + c), s);
    unpack8e(*reinterpret_cast<const uint4*>(gate + static_cast<size_t>(r) * ldg + c), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] *= gelu_erf_f(g[e]);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * ldo + c) = pack8e(s);
  }
}

// ---- tiny-M Linear: one warp per output column, all M rows at once ---------------------------------
// Weight-bandwidth bound (M <= 32 rows against up to 13 760 x 1280 weights = 35 MB for the batched resnet time-embedding
// projection).  The weight row of a warp is a parameter, so ALL of its 16-byte loads are issued before the programmatic
// dependency on the previous kernel resolves (kIters independent loads in flight per lane); the activations -- with
// SiLU applied once per CTA, not once per output column -- are staged in shared memory as fp32 when they fit.
constexpr int kSmallMMax = 32;
constexpr int kSmallMStageFloats = 12 * 1024;  // 48 KB of staged activations (M * K fp32)

// kRows bounds M at compile time: the per-row loops are fully unrolled and predicated, and predicated-off iterations
// still take issue slots (with a fixed bound of 32 the M = 2 time-embedding projection spent its time issuing them).
template <int kIters, int kRows>  // 16-byte weight vectors per lane: K <= kIters * 256; M <= kRows
__global__ void __launch_bounds__(256)
linear_small_m_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ W, int ldw,
                      const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y, int ldy, int M, int N,
                      int K, int silu_in, int silu_out, int stage_x, int w_static) {
  pdl_launch_dependents();
  if (!w_static) pdl_wait();  // W may be the output of the preceding kernel: no loads ahead of the dependency
  extern __shared__ float s_x[];  // [M][K] act(x), if stage_x
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  const bool active = n < N;
  uint4 wv[kIters];
  {
    const __nv_bfloat16* wr = W + static_cast<size_t>(active ? n : 0) * ldw;
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int k = lane * 8 + i * 256;
      wv[i] = (active && k < K) ? __ldg(reinterpret_cast<const uint4*>(wr + k)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (w_static) pdl_wait();
  if (stage_x) {
    for (int i = threadIdx.x * 8; i < M * K; i += blockDim.x * 8) {
      const int m = i / K, k = i - m * K;  // K % 8 == 0: a vector never straddles two rows
      float xv[8];
      unpack8e(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(m) * ldx + k), xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) s_x[i + e] = silu_in ? silu_f(xv[e]) : xv[e];
    }
    __syncthreads();
  }
  if (!active) return;
  float acc[kRows];
#pragma unroll
  for (int m = 0; m < kRows; ++m) acc[m] = 0.f;
#pragma unroll
  for (int i = 0; i < kIters; ++i) {
    const int k = lane * 8 + i * 256;
    if (k < K) {
      float wf[8];
      unpack8e(wv[i], wf);
#pragma unroll
      for (int m = 0; m < kRows; ++m) {
        if (m < M) {
          float xv[8];
          if (stage_x) {
            const float4 a = *reinterpret_cast<const float4*>(s_x + m * K + k);
            const float4 b = *reinterpret_cast<const float4*>(s_x + m * K + k + 4);
            xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
            xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
          } else {
            unpack8e(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(m) * ldx + k), xv);
            if (silu_in) {
#pragma unroll
              for (int e = 0; e < 8; ++e) xv[e] = silu_f(xv[e]);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[m] = fmaf(xv[e], wf[e], acc[m]);
        }
      }
    }
  }
  const float b = bias ? __bfloat162float(bias[n]) : 0.f;
#pragma unroll
  for (int m = 0; m < kRows; ++m) {
    if (m < M) {
      float v = acc[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      v += b;
      if (silu_out) v = silu_f(v);
      if (lane == 0) y[static_cast<size_t>(m) * ldy + n] = __float2bfloat16(v);
    }
  }
}

// ---- direct 3x3 conv, tiny C (conv_in): thread = (pixel, 8 output channels) ---------------------
// x is addressed through explicit element strides so an NCHW fp32-converted latent can be consumed
// as-is; weights [K][3][3][C] are staged in shared memory as fp32 [(tap*C+c)][K].
__global__ void __launch_bounds__(256)
conv3x3_small_c_kernel(const __nv_bfloat16* __restrict__ x, long long xs_n, long long xs_h, long long xs_w,
                       long long xs_c, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                       __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, int K) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_w[];  // [9*C][K]
  const int taps = 9 * C;
  for (int i = threadIdx.x; i < taps * K; i += blockDim.x) {
    const int k = i / taps, t = i - k * taps;
    s_w[t * K + k] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const int kv = K / 8;
  const long long total = static_cast<long long>(N) * H * W * kv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % kv);
    const long long pix = i / kv;
    const int q = static_cast<int>(pix % W);
    const int p = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[8];
    if (bias) {
      unpack8e(*reinterpret_cast<const uint4*>(bias + v * 8), acc);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    }
    for (int r = 0; r < 3; ++r) {
      const int ih = p + r - 1;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = q + s - 1;
        if (iw < 0 || iw >= W) continue;
        const __nv_bfloat16* xp = x + n * xs_n + ih * xs_h + iw * xs_w;
        for (int c = 0; c < C; ++c) {
          const float xv = __bfloat162float(xp[c * xs_c]);
          const float* wp = s_w + ((r * 3 + s) * C + c) * K + v * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv, wp[e], acc[e]);
        }
      }
    }
    *reinterpret_cast<uint4*>(y + pix * K + v * 8) = pack8e(acc);
  }
}

// ---- direct 3x3 conv, tiny K (conv_out): one warp per output pixel --------------------------------
// y is addressed through explicit element strides so the result can be written straight into an
// NCHW tensor.  Weights [K][3][3][C] staged in shared memory (bf16).
template <int KMAX>
__global__ void __launch_bounds__(256)
conv3x3_small_k_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                       const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y, long long ys_n,
                       long long ys_h, long long ys_w, long long ys_c, int N, int H, int W, int C, int K) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t s_raw[];
  __nv_bfloat16* s_w = reinterpret_cast<__nv_bfloat16*>(s_raw);  // [K][9*C]
  const int wlen = K * 9 * C;
  for (int i = threadIdx.x * 8; i < wlen; i += blockDim.x * 8)
    *reinterpret_cast<uint4*>(s_w + i) = *reinterpret_cast<const uint4*>(w + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int cv = C / 8;
  const long long total = static_cast<long long>(N) * H * W;
  for (long long pix = blockIdx.x * static_cast<long long>(warps) + warp; pix < total;
       pix += static_cast<long long>(gridDim.x) * warps) {
    const int q = static_cast<int>(pix % W);
    const int p = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int ih = p + r - 1;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = q + s - 1;
        if (iw < 0 || iw >= W) continue;
        const __nv_bfloat16* xp = x + ((static_cast<size_t>(n) * H + ih) * W + iw) * C;
        const int tap = r * 3 + s;
        for (int v = lane; v < cv; v += 32) {
          float xv[8];
          unpack8e(*reinterpret_cast<const uint4*>(xp + v * 8), xv);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              float wv[8];
              unpack8e(*reinterpret_cast<const uint4*>(s_w + (static_cast<size_t>(k) * 9 + tap) * C + v * 8), wv);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[k] = fmaf(xv[e], wv[e], acc[k]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          const float b = bias ? __bfloat162float(bias[k]) : 0.f;
          y[n * ys_n + p * ys_h + q * ys_w + k * ys_c] = __float2bfloat16(acc[k] + b);
        }
    }
  }
}

// ---- im2col 3x3 pad 1 stride s, NHWC -> [N*Ho*Wo, 9*C] (tap-major) --------------------------------
__global__ void im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, int N, int H,
                                 int W, int C, int Ho, int Wo, int stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * 9 * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    long long t = i / cv;
    const int tap = static_cast<int>(t % 9);
    t /= 9;
    const int q = static_cast<int>(t % Wo);
    t /= Wo;
    const int p = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const int ih = p * stride + tap / 3 - 1;
    const int iw = q * stride + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
      val = *reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + ih) * W + iw) * C + v * 8);
    *reinterpret_cast<uint4*>(col + i * 8) = val;
  }
}

// ---- im2col for tiny C (conv_in: C = 4): rows of 9*C values (tap-major) zero-padded to 64, so that the
// convolution becomes one K = 64 tensor-core GEMM.  x is addressed through element strides (NCHW or NHWC).
__global__ void im2col3x3_smallc_kernel(const __nv_bfloat16* __restrict__ x, long long xs_n, long long xs_h,
                                        long long xs_w, long long xs_c, __nv_bfloat16* __restrict__ col, int N, int H,
                                        int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(N) * H * W * 8;  // 8 vectors of 8 elements per output row
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i & 7);
    const long long pix = i >> 3;
    const int q = static_cast<int>(pix % W);
    const int p = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = v * 8 + e;
      unsigned short bits = 0;
      if (idx < 9 * C) {
        const int tap = idx / C, c = idx - tap * C;
        const int ih = p + tap / 3 - 1, iw = q + tap % 3 - 1;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W)
          bits = __bfloat16_as_ushort(x[n * xs_n + ih * xs_h + iw * xs_w + c * xs_c]);
      }
      w[e >> 1] |= static_cast<uint32_t>(bits) << ((e & 1) * 16);
    }
    *reinterpret_cast<uint4*>(col + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- [P, ld] rows (NHWC, first C channels) -> dense NCHW; used for the 4-channel conv_out result ----
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int ld, __nv_bfloat16* __restrict__ dst,
                                    int N, int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(N) * C * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int hw = static_cast<int>(i % HW);
    const int c = static_cast<int>((i / HW) % C);
    const int n = static_cast<int>(i / (static_cast<long long>(HW) * C));
    dst[i] = src[(static_cast<long long>(n) * HW + hw) * ld + c];
  }
}

// ---- nearest 2x upsample, NHWC ---------------------------------------------------------------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                  int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(N) * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    long long t = i / cv;
    const int q = static_cast<int>(t % Wo);
    t /= Wo;
    const int p = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    *reinterpret_cast<uint4*>(y + i * 8) =
        *reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + (p >> 1)) * W + (q >> 1)) * C + v * 8);
  }
}

// ---- channel concat ---------------------------------------------------------------------------------
__global__ void concat_channels_kernel(const __nv_bfloat16* __restrict__ a, int Ca, const __nv_bfloat16* __restrict__ b,
                                       int Cb, __nv_bfloat16* __restrict__ y, long long P) {
  pdl_launch_dependents();
  pdl_wait();
  const int va = Ca / 8, vb = Cb / 8, vt = va + vb;
  const long long total = P * vt;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vt);
    const long long pix = i / vt;
    const uint4 val = v < va ? *reinterpret_cast<const uint4*>(a + pix * Ca + v * 8)
                             : *reinterpret_cast<const uint4*>(b + pix * Cb + (v - va) * 8);
    *reinterpret_cast<uint4*>(y + i * 8) = val;
  }
}

// ---- sinusoidal timestep embedding (unet_pt.py:22-36) -----------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int ldo, int B,
                                          int half) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float exponent = (-9.210340371976184f * static_cast<float>(j)) / static_cast<float>(half);
  const float arg = t[b] * expf(exponent);
  out[static_cast<size_t>(b) * ldo + j] = __float2bfloat16(cosf(arg));
  out[static_cast<size_t>(b) * ldo + half + j] = __float2bfloat16(sinf(arg));
}

// ---- Euler-discrete + CFG -------------------------------------------------------------------------------
// scale_model_input: model_in[r, :] = bf16(x / sqrt(sigma_i^2 + 1)) for r in {0, 1} (the CFG pair)
__global__ void scale_model_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ model_in,
                                         long long n, int copies, const float* __restrict__ sigmas,
                                         const int* __restrict__ step) {
  pdl_launch_dependents();
  pdl_wait();
  const float sigma = sigmas[*step];
  const float inv = rsqrtf(sigma * sigma + 1.f);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const __nv_bfloat16 v = __float2bfloat16(x[i] * inv);
    for (int c = 0; c < copies; ++c) model_in[c * n + i] = v;
  }
}
// x += (sigma_{i+1} - sigma_i) * (eps_u + g (eps_c - eps_u)); eps rows: [uncond ; cond]
__global__ void euler_cfg_update_kernel(const __nv_bfloat16* __restrict__ eps_uncond,
                                        const __nv_bfloat16* __restrict__ eps_cond, float* __restrict__ x,
                                        long long n, float guidance, const float* __restrict__ sigmas,
                                        const int* __restrict__ step) {
  pdl_launch_dependents();
  pdl_wait();
  const int s = *step;
  const float dt = sigmas[s + 1] - sigmas[s];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float u = __bfloat162float(eps_uncond[i]);
    const float c = eps_cond ? __bfloat162float(eps_cond[i]) : u;
    const float e = u + guidance * (c - u);
    x[i] = fmaf(dt, e, x[i]);
  }
}
__global__ void advance_step_kernel(int* step, float* t_out, const float* timesteps) {
  pdl_launch_dependents();
  pdl_wait();
  const int s = *step + 1;
  *step = s;
  if (t_out && timesteps) *t_out = timesteps[s];
}

// ---- 1x1 convolution between tiny channel counts (<= 8 -> <= 8), NCHW in, NCHW out ----------------
// The VAE's post_quant_conv (4 -> 4 on the latent, Diffusers AutoencoderKL.decode): 16 multiply-adds per pixel, far too
// small for the tensor-core GEMM (K would be padded from 4 to 64).  y[n, o, p] = b[o] + sum_i w[o, i] * in_scale * x[n, i, p].
__global__ void pointwise_conv_small_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                            const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int N,
                                            long long HW, int Ci, int Co, float in_scale) {
  pdl_launch_dependents();
  pdl_wait();
  float wr[64], br[8];
  for (int i = 0; i < Co * Ci; ++i) wr[i] = __bfloat162float(w[i]) * in_scale;
  for (int o = 0; o < Co; ++o) br[o] = b ? __bfloat162float(b[o]) : 0.f;
  const long long total = static_cast<long long>(N) * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    float xi[8];
    for (int c = 0; c < Ci; ++c) xi[c] = __bfloat162float(x[(n * Ci + c) * HW + p]);
    for (int o = 0; o < Co; ++o) {
      float acc = br[o];
      for (int c = 0; c < Ci; ++c) acc = fmaf(wr[o * Ci + c], xi[c], acc);
      y[(n * Co + o) * HW + p] = __float2bfloat16(acc);
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int grid_for(long long total, int block, int max_blocks_per_sm = 8) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(device_sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace st

extern "C" {

int st_geglu_bf16(const void* state, int ld_state, const void* gate, int ld_gate, void* out, int ld_out, int rows,
                  int cols, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(state && gate && out, "geglu: null pointer");
  ST_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0, "geglu: cols (%d) must be a positive multiple of 8", cols);
  ST_CHECK_ARG(ld_state % 8 == 0 && ld_gate % 8 == 0 && ld_out % 8 == 0, "geglu: pitches must be multiples of 8");
  ST_CHECK_ARG(aligned16(state) && aligned16(gate) && aligned16(out), "geglu: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(rows) * (cols / 8);
  launch_kernel(geglu_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(state), ld_state, static_cast<const __nv_bfloat16*>(gate), ld_gate,
      static_cast<__nv_bfloat16*>(out), ld_out, rows, cols / 8);
  ST_CHECK_LAUNCH("geglu_kernel");
  return ST_OK;
}

int st_linear_small_m_bf16(const void* x, int ldx, const void* W, int ldw, const void* bias, void* y, int ldy, int M,
                           int N, int K, int silu_in, int silu_out, unsigned flags, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && W && y, "linear_small_m: null pointer");
  const int w_static = (flags & ST_W_STATIC) ? 1 : 0;
  ST_CHECK_ARG(M > 0 && M <= kSmallMMax, "linear_small_m: M (%d) must be in [1, %d]", M, kSmallMMax);
  ST_CHECK_ARG(N > 0 && K > 0 && K % 8 == 0, "linear_small_m: K (%d) must be a positive multiple of 8", K);
  ST_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "linear_small_m: pitches must be multiples of 8");
  ST_CHECK_ARG(aligned16(x) && aligned16(W), "linear_small_m: pointers must be 16-byte aligned");
  ST_CHECK_ARG(K <= 16 * 256, "linear_small_m: K (%d) must be <= 4096", K);
  const int warps = 8;
  const int stage_x = static_cast<long>(M) * K <= kSmallMStageFloats ? 1 : 0;
  const size_t smem = stage_x ? static_cast<size_t>(M) * K * sizeof(float) : 0;
  const dim3 grid((N + warps - 1) / warps), block(warps * 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(W);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(bias);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  const int iters = (K + 255) / 256;
#define ST_SMALL_M_LAUNCH(I, R) \
  launch_kernel(linear_small_m_kernel<I, R>, grid, block, smem, s, xp, ldx, wp, ldw, bp, yp, ldy, M, N, K, silu_in, silu_out, stage_x, w_static)
#define ST_SMALL_M_CASE(I)                  \
  do {                                      \
    if (M <= 2) ST_SMALL_M_LAUNCH(I, 2);    \
    else if (M <= 4) ST_SMALL_M_LAUNCH(I, 4); \
    else if (M <= 8) ST_SMALL_M_LAUNCH(I, 8); \
    else ST_SMALL_M_LAUNCH(I, kSmallMMax);  \
  } while (0)
  if (iters <= 2) ST_SMALL_M_CASE(2);
  else if (iters <= 5) ST_SMALL_M_CASE(5);
  else if (iters <= 11) ST_SMALL_M_CASE(11);
  else ST_SMALL_M_CASE(16);
#undef ST_SMALL_M_CASE
#undef ST_SMALL_M_LAUNCH
  ST_CHECK_LAUNCH("linear_small_m_kernel");
  return ST_OK;
}

int st_conv3x3_direct_bf16(const void* x, long long xs_n, long long xs_h, long long xs_w, long long xs_c,
                           const void* w, const void* bias, void* y, long long ys_n, long long ys_h, long long ys_w,
                           long long ys_c, int N, int H, int W, int C, int K, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && w && y, "conv3x3_direct: null pointer");
  ST_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && K > 0, "conv3x3_direct: sizes must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C <= 8) {
    ST_CHECK_ARG(K % 8 == 0, "conv3x3_direct: K (%d) must be a multiple of 8 when C <= 8", K);
    ST_CHECK_ARG(ys_c == 1 && ys_w == K && ys_h == (long long)W * K && ys_n == (long long)H * W * K,
                 "conv3x3_direct: output must be dense NHWC when C <= 8");
    ST_CHECK_ARG(aligned16(y) && (!bias || aligned16(bias)), "conv3x3_direct: y/bias must be 16-byte aligned");
    const size_t smem = static_cast<size_t>(9) * C * K * sizeof(float);
    ST_CHECK_ARG(smem <= 96 * 1024, "conv3x3_direct: weights do not fit in shared memory");
    static PerDeviceOnce configured;
    const int dev = current_device();
    ST_CHECK_ARG(dev >= 0, "conv3x3_direct: device ordinal outside [0, %d)", kMaxDevices);
    if (!configured.done(dev)) {
      cudaFuncSetAttribute(conv3x3_small_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      configured.mark(dev);
    }
    const long long total = static_cast<long long>(N) * H * W * (K / 8);
    launch_kernel(conv3x3_small_c_kernel, dim3(grid_for(total, 256, 4)), dim3(256), smem, s, static_cast<const __nv_bfloat16*>(x), xs_n, xs_h, xs_w, xs_c, static_cast<const __nv_bfloat16*>(w),
        static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), N, H, W, C, K);
    ST_CHECK_LAUNCH("conv3x3_small_c_kernel");
    return ST_OK;
  }
  ST_CHECK_ARG(K <= 8, "conv3x3_direct: needs C <= 8 or K <= 8 (got C=%d, K=%d)", C, K);
  ST_CHECK_ARG(C % 8 == 0, "conv3x3_direct: C (%d) must be a multiple of 8 when K <= 8", C);
  ST_CHECK_ARG(xs_c == 1 && xs_w == C && xs_h == (long long)W * C && xs_n == (long long)H * W * C,
               "conv3x3_direct: input must be dense NHWC when K <= 8");
  ST_CHECK_ARG(aligned16(x) && aligned16(w), "conv3x3_direct: x/w must be 16-byte aligned");
  const size_t smem = static_cast<size_t>(K) * 9 * C * 2;
  ST_CHECK_ARG(smem <= 96 * 1024 && (K * 9 * C) % 8 == 0, "conv3x3_direct: weights do not fit in shared memory");
  static PerDeviceOnce configured_k;
  const int dev_k = current_device();
  ST_CHECK_ARG(dev_k >= 0, "conv3x3_direct: device ordinal outside [0, %d)", kMaxDevices);
  if (!configured_k.done(dev_k)) {
    cudaFuncSetAttribute(conv3x3_small_k_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(conv3x3_small_k_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    configured_k.mark(dev_k);
  }
  const long long pixels = static_cast<long long>(N) * H * W;
  const int grid = grid_for(pixels * 32, 256, 4);
  if (K <= 4)
    launch_kernel(conv3x3_small_k_kernel<4>, dim3(grid), dim3(256), smem, s, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
        static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), ys_n, ys_h, ys_w, ys_c, N, H, W, C, K);
  else
    launch_kernel(conv3x3_small_k_kernel<8>, dim3(grid), dim3(256), smem, s, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
        static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), ys_n, ys_h, ys_w, ys_c, N, H, W, C, K);
  ST_CHECK_LAUNCH("conv3x3_small_k_kernel");
  return ST_OK;
}

int st_im2col3x3_nhwc_bf16(const void* x, void* col, int N, int H, int W, int C, int stride, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && col, "im2col: null pointer");
  ST_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "im2col: C (%d) must be a positive multiple of 8", C);
  ST_CHECK_ARG(stride == 1 || stride == 2, "im2col: stride must be 1 or 2");
  ST_CHECK_ARG(aligned16(x) && aligned16(col), "im2col: pointers must be 16-byte aligned");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = static_cast<long long>(N) * Ho * Wo * 9 * (C / 8);
  launch_kernel(im2col3x3_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(col), N, H, W, C, Ho, Wo, stride);
  ST_CHECK_LAUNCH("im2col3x3_kernel");
  return ST_OK;
}

int st_im2col3x3_smallc_bf16(const void* x, long long xs_n, long long xs_h, long long xs_w, long long xs_c, void* col,
                             int N, int H, int W, int C, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && col, "im2col_smallc: null pointer");
  ST_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && 9 * C <= 64, "im2col_smallc: needs 9*C <= 64 (got C=%d)", C);
  ST_CHECK_ARG(aligned16(col), "im2col_smallc: col must be 16-byte aligned");
  const long long total = static_cast<long long>(N) * H * W * 8;
  launch_kernel(im2col3x3_smallc_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                static_cast<const __nv_bfloat16*>(x), xs_n, xs_h, xs_w, xs_c, static_cast<__nv_bfloat16*>(col), N, H, W, C);
  ST_CHECK_LAUNCH("im2col3x3_smallc_kernel");
  return ST_OK;
}

int st_nhwc_to_nchw_bf16(const void* src, int ld, void* dst, int N, int HW, int C, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(src && dst, "nhwc_to_nchw: null pointer");
  ST_CHECK_ARG(N > 0 && HW > 0 && C > 0 && ld >= C, "nhwc_to_nchw: bad sizes");
  const long long total = static_cast<long long>(N) * C * HW;
  launch_kernel(nhwc_to_nchw_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                static_cast<const __nv_bfloat16*>(src), ld, static_cast<__nv_bfloat16*>(dst), N, HW, C);
  ST_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return ST_OK;
}

int st_pointwise_conv_small_bf16(const void* x, const void* w, const void* bias, void* y, int N, long long HW, int Ci,
                                 int Co, float in_scale, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && w && y, "pointwise_conv_small: null pointer");
  ST_CHECK_ARG(N > 0 && HW > 0 && Ci > 0 && Ci <= 8 && Co > 0 && Co <= 8, "pointwise_conv_small: channel counts must be in [1, 8]");
  launch_kernel(pointwise_conv_small_kernel, dim3(grid_for(static_cast<long long>(N) * HW, 256)), dim3(256), 0,
                static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
                static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), N, HW, Ci, Co, in_scale);
  ST_CHECK_LAUNCH("pointwise_conv_small_kernel");
  return ST_OK;
}

int st_upsample_nearest2x_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && y, "upsample: null pointer");
  ST_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample: C (%d) must be a positive multiple of 8", C);
  ST_CHECK_ARG(aligned16(x) && aligned16(y), "upsample: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(N) * 4 * H * W * (C / 8);
  launch_kernel(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, H, W, C);
  ST_CHECK_LAUNCH("upsample2x_kernel");
  return ST_OK;
}

int st_concat_channels_bf16(const void* a, int Ca, const void* b, int Cb, void* y, long long P, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(a && b && y, "concat: null pointer");
  ST_CHECK_ARG(P > 0 && Ca > 0 && Cb > 0 && Ca % 8 == 0 && Cb % 8 == 0, "concat: channel counts must be multiples of 8");
  ST_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(y), "concat: pointers must be 16-byte aligned");
  const long long total = P * ((Ca + Cb) / 8);
  launch_kernel(concat_channels_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(a), Ca, static_cast<const __nv_bfloat16*>(b), Cb,
      static_cast<__nv_bfloat16*>(y), P);
  ST_CHECK_LAUNCH("concat_channels_kernel");
  return ST_OK;
}

int st_timestep_embedding_bf16(const float* t, void* out, int ldo, int B, int half, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(t && out, "timestep_embedding: null pointer");
  ST_CHECK_ARG(B > 0 && half > 0 && ldo >= 2 * half, "timestep_embedding: bad sizes");
  const int total = B * half;
  launch_kernel(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), t, static_cast<__nv_bfloat16*>(out), ldo, B, half);
  ST_CHECK_LAUNCH("timestep_embedding_kernel");
  return ST_OK;
}

int st_scale_model_input(const float* x, void* model_in, long long n, int copies, const float* sigmas,
                         const int* step, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(x && model_in && sigmas && step, "scale_model_input: null pointer");
  ST_CHECK_ARG(n > 0 && copies > 0, "scale_model_input: bad sizes");
  launch_kernel(scale_model_input_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, static_cast<__nv_bfloat16*>(model_in), n, copies, sigmas, step);
  ST_CHECK_LAUNCH("scale_model_input_kernel");
  return ST_OK;
}

int st_euler_cfg_update(const void* eps_uncond, const void* eps_cond, float* x, long long n, float guidance,
                        const float* sigmas, const int* step, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(eps_uncond && x && sigmas && step, "euler_cfg_update: null pointer");
  ST_CHECK_ARG(n > 0, "euler_cfg_update: bad sizes");
  launch_kernel(euler_cfg_update_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(eps_uncond), static_cast<const __nv_bfloat16*>(eps_cond), x, n, guidance,
      sigmas, step);
  ST_CHECK_LAUNCH("euler_cfg_update_kernel");
  return ST_OK;
}

int st_advance_step(int* step, float* t_out, const float* timesteps, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(step, "advance_step: null pointer");
  launch_kernel(advance_step_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), step, t_out, timesteps);
  ST_CHECK_LAUNCH("advance_step_kernel");
  return ST_OK;
}

}  // extern "C"
