// Pipe-throughput microbenchmark for the softmax inner loop (B200): cycles per warp-instruction of MUFU.EX2,
// F2FP.BF16.PACK_AB and their mix, with 1 / 2 / 4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/pipebench tools/pipebench.cu
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <int MODE>
__global__ void bench(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  float facc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 ex2
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 1) {  // 8 packs
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(x[i], x[i + 1]);
    } else if (MODE == 2) {  // 16 ex2 + 8 packs (the softmax chunk)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(x[i], x[i + 1]);
    } else if (MODE == 3) {  // 16 ex2 + 16 fma + 16 fadd + 8 packs
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(fmaf(x[i], 0.99f, -0.01f));
#pragma unroll
      for (int i = 0; i < 16; ++i) facc += x[i];
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(x[i], x[i + 1]);
    } else if (MODE == 5) {  // 8 packed half2 ex2 = 16 exponentials
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = pack_f16(x[2 * i], x[2 * i + 1]) ^ acc;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = ex2_f16x2(h[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc ^= h[i];
    } else if (MODE == 6) {  // the attention chunk: 8 FFMA2 + 16 EX2 + 8 FADD2 + 8 F2FP (packed pairs)
      float f0 = 0.f, f1 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        asm volatile("{.reg .b64 a, b, c, d; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mov.b64 c, {%3, %3};\n"
                     "fma.rn.ftz.f32x2 d, a, b, c; mov.b64 {%0, %1}, d;}" : "+f"(x[i]), "+f"(x[i + 1]) : "f"(0.99f), "f"(-0.01f));
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; add.rn.ftz.f32x2 a, a, b; mov.b64 {%0, %1}, a;}"
                     : "+f"(f0), "+f"(f1) : "f"(x[i]), "f"(x[i + 1]));
      facc += f0 + f1;
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(x[i], x[i + 1]);
    } else if (MODE == 7) {  // 16 EX2 + 8 integer-op packs (round-to-nearest-even by hand + PRMT): no F2FP
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t a = __float_as_uint(x[i]), b = __float_as_uint(x[i + 1]);
        a += 0x7fffu + ((a >> 16) & 1u);
        b += 0x7fffu + ((b >> 16) & 1u);
        acc ^= __byte_perm(a, b, 0x7632);
      }
    } else if (MODE == 8) {  // 16 EX2 + 16 independent FFMA (does FMA-pipe work issue under the MUFU stream?)
      float y[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = fmaf(x[i], 0.99f, -0.01f);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) facc += y[i];
    } else if (MODE == 4) {  // 8 packs done with integer ops instead (round-to-nearest-even by hand: 4 ALU ops / pair)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t a = __float_as_uint(x[i]), b = __float_as_uint(x[i + 1]);
        a += 0x7fffu + ((a >> 16) & 1u);
        b += 0x7fffu + ((b >> 16) & 1u);
        acc ^= __byte_perm(a, b, 0x7632);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  float s = facc;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
}

template <int MODE>
void run(const char* name, int per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 2000;
    bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %2d warps/SM (%d per sub-partition): %7.1f cycles / iteration / warp, %5.2f cycles per instr per sub-partition\n",
           name, warps, warps / 4, double(h) / iters, double(h) / iters / (per_iter * (warps / 4)));
  }
}

int main() {
  run<0>("16 x MUFU.EX2", 16);
  run<1>("8 x F2FP.BF16.PACK_AB", 8);
  run<2>("16 x EX2 + 8 x F2FP", 24);
  run<3>("16 x (FFMA, EX2, FADD) + 8 x F2FP", 56);
  run<4>("8 x bf16 pack by integer ops", 8);
  run<5>("8 x F2FP.F16 + 32 x EX2.F16x2 (64 exps)", 40);
  run<6>("8 FFMA2 + 16 EX2 + 8 FADD2 + 8 F2FP (attention chunk)", 40);
  run<7>("16 EX2 + 8 integer-op bf16 packs", 24);
  run<8>("16 EX2 + 16 FFMA + 16 FADD", 48);
  return 0;
}
