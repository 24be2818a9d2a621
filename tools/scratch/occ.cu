#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k_tmem(uint32_t* out) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  uint32_t t = slot;
  if (threadIdx.x == 0) out[blockIdx.x] = t;
  // stay resident a bit so that two CTAs can overlap, and record the SM id
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[1024 + blockIdx.x] = smid;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(t) : "memory");
}
__global__ void __launch_bounds__(256, 2) k_plain(uint32_t* out) { if (threadIdx.x == 0) out[blockIdx.x] = 1; }
int main() {
  int n1 = -1, n2 = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_tmem, 192, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n2, k_plain, 192, 0);
  printf("occupancy API: tmem kernel %d, plain kernel %d\n", n1, n2);
  uint32_t* d; cudaMalloc(&d, 8192); cudaMemset(d, 0xff, 8192);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_tmem<<<296, 192>>>(d);
  cudaEventRecord(b); cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("296 CTAs of ~1.05 ms spin each: %.3f ms total (%s)  -> %s\n", ms, cudaGetErrorString(e), ms < 1.8 ? "2 CTAs/SM co-resident" : "1 CTA/SM");
  uint32_t h[2048]; cudaMemcpy(h, d, 8192, cudaMemcpyDeviceToHost);
  printf("tmem addrs of first CTAs: %x %x %x %x\n", h[0], h[1], h[2], h[3]);
  return 0;
}
