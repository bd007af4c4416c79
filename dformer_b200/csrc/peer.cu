// Small-message all-reduce over NVLink peer memory for the SyncBatchNorm statistics (torch SyncBatchNorm of the reference,
// utils/train.py:182-194; 9 layers x forward + backward per step, messages of <= 2 x 512 numbers).
//
// NCCL's latency for such messages (tens of microseconds per call, 18 calls on the compute stream per step) is what kept the
// 8-GPU step 1.2 ms above the 1-GPU step in round 1.  Here every rank owns one cudaMalloc'ed exchange buffer that all peers of
// the box map through CUDA IPC.  One single-CTA kernel per call:
//   push   write the local vector into slot [parity][my rank] of EVERY peer's buffer (plain stores over NVLink / NVSwitch),
//          fence at system scope, then raise flag [parity][my rank] of every peer to the call's sequence number;
//   wait   spin (bounded) until the flags of all peers in the local buffer reach the sequence number;
//   reduce sum the world's slots of the local buffer in rank order (bitwise identical on every rank).
// The sequence number lives in the local buffer and is advanced by the kernel itself, so a captured CUDA graph can be replayed.
// Two slot parities suffice: a rank cannot finish call k+1 before every peer has finished call k (it needs their k+1 flags).
#include <string.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int MAXW = 16;                 // ranks per box
constexpr int SLOT_BYTES = 16 * 1024;    // per (parity, source rank): 2048 doubles
constexpr int FLAG_OFF = 256;            // uint64 flags[2][MAXW]
constexpr int DATA_OFF = 4096;           // slots[2][MAXW][SLOT_BYTES]
constexpr size_t BUF_BYTES = DATA_OFF + 2 * (size_t)MAXW * SLOT_BYTES;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const T* __restrict__ in, T* __restrict__ out, int n, uint8_t* const* __restrict__ bases, int rank,
                                                            int world) {
  pdl_sync();
  uint8_t* mine = bases[rank];
  __shared__ unsigned long long s_seq;
  const int tid = threadIdx.x;
  if (tid == 0) {
    unsigned long long* sp = reinterpret_cast<unsigned long long*>(mine);
    s_seq = *sp + 1;
    *sp = s_seq;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int par = (int)(seq & 1);
  for (int p = 0; p < world; ++p) {
    T* dst = reinterpret_cast<T*>(bases[p] + DATA_OFF + ((size_t)(par * MAXW + rank)) * SLOT_BYTES);
    for (int i = tid; i < n; i += blockDim.x) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    st_release_sys(reinterpret_cast<unsigned long long*>(bases[tid] + FLAG_OFF) + par * MAXW + rank, seq);
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + FLAG_OFF) + par * MAXW + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > 40000000000LL) {       // ~20 s: a peer died or the ranks' call sequences diverged -> fail the launch, never hang the box
        printf("dfb200 peer_allreduce: rank %d timed out waiting for rank %d (call %llu)\n", rank, tid, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    T acc = 0;
    for (int p = 0; p < world; ++p)
      acc += *reinterpret_cast<const volatile T*>(mine + DATA_OFF + ((size_t)(par * MAXW + p)) * SLOT_BYTES + (size_t)i * sizeof(T));
    out[i] = acc;
  }
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_peer_alloc(void** ptr) {
  DFB_REQUIRE(ptr != nullptr, "peer_alloc: null argument");
  cudaError_t e = cudaMalloc(ptr, BUF_BYTES);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, BUF_BYTES);
  if (e != cudaSuccess) { dfb_set_error("peer_alloc: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  return DFB_OK;
}
extern "C" int dfb200_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { dfb_set_error("peer_free: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  return DFB_OK;
}
extern "C" int dfb200_peer_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { dfb_set_error("peer_export: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  memcpy(handle64, &h, 64);
  return DFB_OK;
}
extern "C" int dfb200_peer_open(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); dfb_set_error("peer_open: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  return DFB_OK;
}
extern "C" int dfb200_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { dfb_set_error("peer_close: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  return DFB_OK;
}

extern "C" int dfb200_peer_allreduce(const void* in, void* out, int dtype, int n, void* const* bases_dev, int rank, int world, void* stream) {
  DFB_REQUIRE(world >= 1 && world <= MAXW && rank >= 0 && rank < world, "peer_allreduce: rank %d / world %d (at most %d ranks)", rank, world, MAXW);
  DFB_REQUIRE(n >= 0 && (size_t)n * (dtype == 2 ? 8 : 4) <= SLOT_BYTES, "peer_allreduce: %d elements exceed the %d-byte slot", n, SLOT_BYTES);
  DFB_REQUIRE(dtype == 0 || dtype == 2, "peer_allreduce: dtype must be 0 (float32) or 2 (float64)");
  if (n == 0) return DFB_OK;
  uint8_t* const* bases = reinterpret_cast<uint8_t* const*>(bases_dev);
  if (dtype == 0) dfb_launch(peer_allreduce_kernel<float>, 1, 256, 0, ST, (const float*)in, (float*)out, n, bases, rank, world);
  else dfb_launch(peer_allreduce_kernel<double>, 1, 256, 0, ST, (const double*)in, (double*)out, n, bases, rank, world);
  return dfb_check_launch("peer_allreduce");
}
