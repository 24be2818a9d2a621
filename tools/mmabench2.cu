// cta_group::2 tcgen05.mma rate microbenchmark: a 2-CTA cluster, the leader issues reps x 4 256 x N x 16 MMAs
// (each CTA holds its 128 rows of A and N/2 rows of B at the same shared-memory offsets).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mmabench2 tools/mmabench2.cu
#include <cstdio>
#include <cstdlib>

#include "../stabletriton_b200/csrc/ptx.cuh"
using namespace st;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(int N, int reps, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = align_smem_1024(raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 32) tmem_alloc_pair<512>(&slot);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();
  const uint32_t tmem = slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) {
      da[k] = umma_smem_desc_sw128(smem_u32(smem) + k * 32, 0, 1024);
      db[k] = umma_smem_desc_sw128(smem_u32(smem) + 32 * 1024 + k * 32, 0, 1024);
    }
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss_pair_elect(tmem, da[k], db[k], idesc, 1);
    }
    umma_commit_pair_elect(&bar, 0x1);
    __syncwarp();
    t1 = clock64();
    mbar_wait(&bar, 0);
    t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 2;
  unsigned long long* out;
  cudaMalloc(&out, 16);
  const int smem = 66 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N = 64; N <= 256; N *= 2) {
    unsigned long long h[2][2];
    for (int i = 0; i < 2; ++i) {
      bench<<<grid, 128, smem>>>(N, i ? 512 : 256, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h[i], out, 16, cudaMemcpyDeviceToHost);
    }
    const double per = double(h[1][1] - h[0][1]) / (256 * 4);
    printf("pair SS M=256 N=%3d : %7.1f cycles / MMA -> %5.1f%% of 2 x 8192 FLOP/clk\n", N, per,
           100.0 * (2.0 * 256 * N * 16 / per) / 16384.0);
  }
  return 0;
}
