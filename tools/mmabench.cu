// tcgen05.mma issue-rate microbenchmark (B200, sm_100a): one CTA per SM issues `reps` x 4 back-to-back
// 128 x N x 16 bf16 MMAs into one TMEM accumulator and reports cycles per instruction, for
//   SS (A and B from shared memory) / TS (A from TMEM), K-major / MN-major B, SWIZZLE_128B / SWIZZLE_32B tiles.
// Operand contents are irrelevant (uninitialised shared memory); only the descriptors must be legal.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mmabench tools/mmabench.cu
//   tools/mmabench [grid]
#include <cstdio>
#include <cstdlib>

#include "../stabletriton_b200/csrc/ptx.cuh"

using namespace st;

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

struct Case {
  int ts;      // 0: SS, 1: TS
  int M, N;
  int b_mn;    // B MN-major (N <= 64 here)
  int sw32;    // operands laid out as SWIZZLE_32B K-major tiles ([rows][16 bf16], 8-row atoms of 256 B)
  int reps;
  int variant;  // 0: descriptors rebuilt per MMA under `if (thread 0)`; 1: descriptors hoisted; 2: hoisted + whole
                // warp runs the loop, the MMA itself predicated by elect.sync inside the asm block
};

__device__ __forceinline__ void mma_ss_elect(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(Case c, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // give the operands finite contents (denormal / NaN handling must not be what we time)
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = umma_idesc_bf16(c.M, c.N, 0, c.b_mn);
  const uint32_t a_addr = smem_u32(smem);
  const uint32_t b_addr = a_addr + 32 * 1024;
  auto make = [&](int k, uint64_t& da, uint64_t& db) {
    if (c.sw32) {
      da = smem_desc(a_addr + k * (c.M * 32), 0, 256, 6);
      db = smem_desc(b_addr + k * (c.N * 32), 0, 256, 6);
    } else {
      da = smem_desc(a_addr + k * 32, 0, 1024, 2);
      db = c.b_mn ? smem_desc(b_addr + k * 16 * 128, 8192, 1024, 2) : smem_desc(b_addr + k * 32, 0, 1024, 2);
    }
  };
  long long t0 = 0, t1 = 0, t2 = 0;
  if (c.variant == 0 && threadIdx.x == 0) {
    t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da, db;
        make(k, da, db);
        if (c.ts)
          umma_bf16_ts(tmem, tmem + 256 + k * 8, db, idesc, 1);
        else
          umma_bf16_ss(tmem, da, db, idesc, 1);
      }
    }
    umma_commit(&bar);
    t1 = clock64();
    mbar_wait(&bar, 0);
    t2 = clock64();
  } else if (c.variant == 1 && threadIdx.x == 0) {
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) make(k, da[k], db[k]);
    t0 = clock64();
    if (c.ts) {
      for (int r = 0; r < c.reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem, tmem + 256 + k * 8, db[k], idesc, 1);
      }
    } else {
      for (int r = 0; r < c.reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da[k], db[k], idesc, 1);
      }
    }
    umma_commit(&bar);
    t1 = clock64();
    mbar_wait(&bar, 0);
    t2 = clock64();
  } else if (c.variant == 2 && threadIdx.x < 32) {
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) make(k, da[k], db[k]);
    t0 = clock64();
    if (c.ts) {
      for (int r = 0; r < c.reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ts_elect(tmem, tmem + 256 + k * 8, db[k], idesc);
      }
    } else {
      for (int r = 0; r < c.reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss_elect(tmem, da[k], db[k], idesc);
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    t1 = clock64();
    mbar_wait(&bar, 0);
    t2 = clock64();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  unsigned long long* out;
  cudaMalloc(&out, 16);
  const int smem = 97 * 1024 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Case cases[] = {
      {0, 128, 64, 0, 0, 0},  {0, 128, 128, 0, 0, 0}, {0, 128, 256, 0, 0, 0},   // SS, SW128 K-major
      {0, 64, 64, 0, 0, 0},   {0, 64, 128, 0, 0, 0},  {0, 64, 256, 0, 0, 0},    // SS, M = 64
      {0, 128, 64, 0, 1, 0},  {0, 128, 128, 0, 1, 0}, {0, 128, 256, 0, 1, 0},   // SS, SW32 K-major
      {1, 128, 64, 0, 0, 0},  {1, 128, 128, 0, 0, 0}, {1, 128, 256, 0, 0, 0},   // TS, B K-major
      {1, 128, 64, 1, 0, 0},                                                    // TS, B MN-major (P.V)
      {0, 128, 64, 1, 0, 0},                                                    // SS, B MN-major
  };
  printf("grid %d\n", grid);
  for (int variant = 0; variant < 3; ++variant)
  for (const Case& c00 : cases) {
    Case c0 = c00;
    c0.variant = variant;
    if (variant > 0 && (c0.M == 64 || c0.sw32)) continue;
    unsigned long long h[2][2];
    for (int i = 0; i < 2; ++i) {
      Case c = c0;
      c.reps = i ? 512 : 256;
      bench<<<grid, 128, smem>>>(c, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("case failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h[i], out, 16, cudaMemcpyDeviceToHost);
    }
    const double per = double(h[1][1] - h[0][1]) / (256 * 4);
    printf("v%d %s M=%3d N=%3d B %s %s : %7.1f cycles / MMA (issue-only %6.1f)  -> %5.1f%% of 8192 FLOP/clk\n",
           c0.variant, c0.ts ? "TS" : "SS", c0.M, c0.N, c0.b_mn ? "MN-major" : "K-major ", c0.sw32 ? "SW32 " : "SW128",
           per, double(h[1][0] - h[0][0]) / (256 * 4), 100.0 * (2.0 * c0.M * c0.N * 16 / per) / 8192.0);
  }
  return 0;
}
