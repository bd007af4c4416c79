// Shared pieces of the TMA-fed depthwise-convolution kernels (mlp_dw.cu, dw7.cu): mbarrier / bulk-tensor PTX wrappers,
// bf16 <-> packed-fp32 helpers and the cached 4-D tensor map over a channels-last activation.
#pragma once
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string.h>

#include "common.cuh"

namespace tile {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && ++spins > (1L << 26)) {
      if ((threadIdx.x & 31) == 0) printf("dfb200 tile: mbarrier wait timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void unpack4(const uint2& u, float2& lo, float2& hi) {
  lo = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
  hi = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ uint2 pack4(const float2& lo, const float2& hi) {
  uint2 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(lo.x, lo.y), b = __floats2bfloat162_rn(hi.x, hi.y);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

__device__ __forceinline__ uint8_t* align128(uint8_t* p) {
  return p + ((128u - (smem_u32(p) & 127u)) & 127u);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int B, H, W, C, bw, bh, bc;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = 1469598103934665603ull;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
};

// 4-D bf16 tensor map over a channels-last activation [B, H, W, C]; box = bc channels x bw x bh pixels of one image.
inline int make_map_nhwc(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bc = 64) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.B = B; key.H = H; key.W = W; key.C = C; key.bw = bw; key.bh = bh; key.bc = bc;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DFB_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { dfb_set_error("cuTensorMapEncodeTiled entry point not available"); return DFB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dfb_set_error("mlp_dw: cuTensorMapEncodeTiled failed (%d): ptr=%p B=%d H=%d W=%d C=%d box=%dx%d", (int)r, ptr, B, H, W, C, bw, bh);
    return DFB_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return DFB_OK;
}

}  // namespace tile
