// NVLink peer-memory exchange for the 2-GPU CFG split: one prompt on two GPUs, rank 0 evaluates the
// uncond row, rank 1 the cond row, and every denoise step both ranks need both eps rows to form
// eps_u + g (eps_c - eps_u).  Instead of an NCCL all-gather followed by the Euler update, ONE kernel per rank
//   (1) stores its own eps row straight into the peer GPU's exchange slab (P2P stores over NVLink / NVSwitch),
//   (2) publishes a sequence number on the peer with a system-scope release once the whole row is out,
//   (3) waits (system-scope acquire) for the peer's row to land in its own slab, and
//   (4) applies the guidance mix + Euler update to its replica of the latents.
// No host involvement, no extra launch, capturable in the step graph.  The slab is double-buffered by the parity of a
// device-resident epoch counter: a rank can run at most one exchange ahead of its peer (it needs the peer's row of
// exchange e to finish exchange e), so parity e & 1 is never overwritten before it has been read.
//
// The slab comes from cudaMalloc (not from the framework's caching allocator) so that it can be exported with the legacy
// CUDA IPC handles; the two processes swap the 64-byte handles through torch.distributed (plumbing) once, before capture.
// The reference has no multi-GPU path at all; the spec is SURVEY 8e.
#include "common.cuh"
#include "ptx.cuh"

#include <string.h>

namespace st {

constexpr int kPeerCtas = 16;         // all co-resident (the wait in phase 3 needs no forward progress from late CTAs)
constexpr int kPeerThreads = 256;
constexpr int kPeerHeaderBytes = 1024;  // flags + tickets + epoch, one 128-byte line each

// slab = [header 1 KB][parity 0: row 0, row 1][parity 1: row 0, row 1], each row `row_bytes` (multiple of 16)
struct PeerHeader {
  unsigned flag[2][32];    // flag[parity][0]: sequence number of the PEER's row that has fully landed here
  unsigned ticket[2][32];  // ticket[0][0]: CTAs of this launch whose stores are out; ticket[1][0]: CTAs that are done
  unsigned epoch[32];      // exchanges completed by this rank
  unsigned error[32];      // non-zero: a wait timed out (the peer never published)
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "header");

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// eps_local: this rank's eps row, n bf16 (n % 8 == 0).  row: 0 = uncond (this rank's row), 1 = cond.
__global__ void __launch_bounds__(kPeerThreads)
cfg_exchange_euler_kernel(const uint4* __restrict__ eps_local, int row, uint8_t* slab_local, uint8_t* slab_peer,
                          float* __restrict__ x, long long n, float guidance, const float* __restrict__ sigmas,
                          const int* __restrict__ step, unsigned long long timeout_ns) {
  pdl_launch_dependents();
  pdl_wait();
  PeerHeader* mine = reinterpret_cast<PeerHeader*>(slab_local);
  PeerHeader* theirs = reinterpret_cast<PeerHeader*>(slab_peer);
  const unsigned e = mine->epoch[0];  // only the last CTA of a launch advances it, after every CTA has read it
  const unsigned par = e & 1u;
  const size_t row_bytes = static_cast<size_t>(n) * 2;
  const long long nvec = n / 8;
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;

  // (1) my row -> the peer's slab, parity `par`, slot `row`
  uint4* dst = reinterpret_cast<uint4*>(slab_peer + kPeerHeaderBytes + (par * 2 + row) * row_bytes);
  for (long long i = tid; i < nvec; i += nthreads) dst[i] = eps_local[i];
  __threadfence_system();
  __syncthreads();
  // (2) the last CTA to get its stores out publishes the sequence number on the peer
  __shared__ unsigned s_last;
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&mine->ticket[0][0], 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) {
      mine->ticket[0][0] = 0u;
      __threadfence_system();  // every CTA's stores (fenced before its ticket) are ordered before the flag
      st_release_sys(&theirs->flag[par][0], e + 1u);
    }
  }
  // (3) wait for the peer's row of this exchange
  if (threadIdx.x == 0) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(&mine->flag[par][0]) != e + 1u) {
      if (globaltimer_ns() - t0 > timeout_ns) {  // never hang the GPU on a dead peer: flag the error and go on
        mine->error[0] = e + 1u;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  // (4) guidance mix + Euler step on my replica of the latents
  const uint4* other = reinterpret_cast<const uint4*>(slab_local + kPeerHeaderBytes + (par * 2 + (row ^ 1)) * row_bytes);
  const int s = *step;
  const float dt = sigmas[s + 1] - sigmas[s];
  for (long long i = tid; i < nvec; i += nthreads) {
    const uint4 own = eps_local[i];
    const uint4 got = __ldcg(&other[i]);  // written by the peer GPU: read it from L2, never from a stale L1 line
    const uint4 u4 = row == 0 ? own : got;
    const uint4 c4 = row == 0 ? got : own;
    const uint32_t uw[4] = {u4.x, u4.y, u4.z, u4.w};
    const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
    float4 a = *reinterpret_cast<const float4*>(x + 8 * i);
    float4 b = *reinterpret_cast<const float4*>(x + 8 * i + 4);
    float* xs[8] = {&a.x, &a.y, &a.z, &a.w, &b.x, &b.y, &b.z, &b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 u = unpack_bf16x2(uw[k]);
      const float2 c = unpack_bf16x2(cw[k]);
      *xs[2 * k] = fmaf(dt, u.x + guidance * (c.x - u.x), *xs[2 * k]);
      *xs[2 * k + 1] = fmaf(dt, u.y + guidance * (c.y - u.y), *xs[2 * k + 1]);
    }
    *reinterpret_cast<float4*>(x + 8 * i) = a;
    *reinterpret_cast<float4*>(x + 8 * i + 4) = b;
  }
  // the last CTA to finish advances the epoch (every CTA has read it long before)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&mine->ticket[1][0], 1u);
    if (t == gridDim.x - 1) {
      mine->ticket[1][0] = 0u;
      mine->epoch[0] = e + 1u;
    }
  }
}

}  // namespace st

extern "C" {

size_t st_peer_slab_bytes(long long n) { return st::kPeerHeaderBytes + 4 * static_cast<size_t>(n) * 2; }

int st_peer_alloc(size_t bytes, void** ptr) {
  using namespace st;
  ST_CHECK_ARG(ptr && bytes > 0, "peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("peer_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int st_peer_free(void* ptr) {
  using namespace st;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_error("peer_free: %s", cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int st_peer_export(void* ptr, unsigned char* handle64) {
  using namespace st;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  ST_CHECK_ARG(ptr && handle64, "peer_export: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    set_error("peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return ST_OK;
}

int st_peer_import(const unsigned char* handle64, void** ptr) {
  using namespace st;
  ST_CHECK_ARG(ptr && handle64, "peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("peer_import: cudaIpcOpenMemHandle: %s (the two GPUs need P2P access: NVLink / NVSwitch or PCIe P2P)",
              cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int st_peer_close(void* ptr) {
  using namespace st;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_error("peer_close: %s", cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int st_cfg_exchange_euler_update(const void* eps_local, int row, void* slab_local, void* slab_peer, float* x, long long n,
                                 float guidance, const float* sigmas, const int* step, st_stream_t stream) {
  using namespace st;
  ST_CHECK_ARG(eps_local && slab_local && slab_peer && x && sigmas && step, "cfg_exchange: null pointer");
  ST_CHECK_ARG(row == 0 || row == 1, "cfg_exchange: row must be 0 (uncond) or 1 (cond)");
  ST_CHECK_ARG(n > 0 && n % 8 == 0, "cfg_exchange: n (%lld) must be a positive multiple of 8", n);
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(eps_local) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(slab_local) & 15) == 0 && (reinterpret_cast<uintptr_t>(slab_peer) & 15) == 0,
               "cfg_exchange: pointers must be 16-byte aligned");
  const unsigned long long timeout_ns = 2000000000ull;  // 2 s
  launch_kernel(cfg_exchange_euler_kernel, dim3(kPeerCtas), dim3(kPeerThreads), 0, static_cast<cudaStream_t>(stream),
                static_cast<const uint4*>(eps_local), row, static_cast<uint8_t*>(slab_local),
                static_cast<uint8_t*>(slab_peer), x, n, guidance, sigmas, step, timeout_ns);
  ST_CHECK_LAUNCH("cfg_exchange_euler_kernel");
  return ST_OK;
}

// Non-zero: the sequence number of an exchange whose wait for the peer timed out (host read; NOT capturable).
int st_peer_error(const void* slab_local, unsigned* out) {
  using namespace st;
  ST_CHECK_ARG(slab_local && out, "peer_error: null pointer");
  cudaError_t e = cudaMemcpy(out, reinterpret_cast<const uint8_t*>(slab_local) + offsetof(PeerHeader, error), 4,
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_error("peer_error: %s", cudaGetErrorString(e));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

}  // extern "C"
