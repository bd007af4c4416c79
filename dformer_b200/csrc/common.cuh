// Shared device/host helpers for the dformer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#define DFB_OK 0
#define DFB_ERR_ARG (-1)
#define DFB_ERR_CUDA (-2)
#define DFB_ERR_UNSUPPORTED (-3)

void dfb_set_error(const char* fmt, ...);
int dfb_check_launch(const char* what);

#define DFB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      dfb_set_error(__VA_ARGS__);         \
      return DFB_ERR_ARG;                 \
    }                                     \
  } while (0)

#define DFB_DISPATCH_DTYPE(dtype, T, ...)                         \
  do {                                                            \
    if ((dtype) == 0) { typedef float T; __VA_ARGS__ }            \
    else if ((dtype) == 1) { typedef bf16 T; __VA_ARGS__ }        \
    else { dfb_set_error("bad dtype %d", (int)(dtype)); return DFB_ERR_ARG; } \
  } while (0)

static inline int dfb_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL).  Every kernel of this library is launched with the programmatic-stream-serialization
// attribute (DFB200_PDL=0 turns it off) and begins with pdl_sync() = griddepcontrol.wait: the kernel may be scheduled while the
// previous kernel of the stream drains, and blocks there until that kernel has completed and its memory is visible.  All
// global-memory accesses of a kernel come after its pdl_sync(), so what overlaps is launch latency / CTA scheduling (and, in
// the tcgen05 GEMM, the barrier / tensor-memory prologue).  Inside a captured CUDA graph the same-stream kernel->kernel edges
// become programmatic edges.  Measured on the DFormer-L step: 28.72 -> 28.44 ms.  An early griddepcontrol.launch_dependents at
// kernel entry (-DDFB_PDL_EARLY_TRIGGER) was measured SLOWER (29.42 ms): the waiting CTAs of the next kernel take SM slots
// from the kernels of the three other streams of the step.  Round 2 repeated it restricted to grids of <= 296 CTAs (the launch-latency-bound
// kernels): training 21.66 -> 22.22 ms, batch-1 inference unchanged (2.506 ms), batch 8 7.74 -> 7.88 ms.  Not kept.
__device__ __forceinline__ void pdl_sync() {
#ifdef DFB_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool dfb_pdl_enabled();
template <typename... KArgs, typename... Args>
inline void dfb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dfb_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through dfb_check_launch (cudaPeekAtLastError)
}

// same with a thread-block cluster of `cluster_x` CTAs along x (gridDim.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void dfb_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dfb_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// Exact-erf GELU (nn.GELU() default) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, below fp32 GEMM
// round-off): one MUFU.EX2 + one MUFU.RCP (approx forms: no IEEE fix-up subroutine) replace the ~25-instruction erff,
// and the same exp(-x^2/2) serves the Gaussian pdf term of the derivative.
//   q(x) = 0.5 * erfc(|x|/sqrt2) = t (a1/2 + t (a2/2 + ...)) exp(-x^2/2),  t = 1 / (1 + p |x| / sqrt2)
//   Phi(x) = x >= 0 ? 1 - q : q          GELU(x) = max(x, 0) - |x| q          GELU'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_q(float x, float& q, float& e) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
  e = ex2_approx(x * x * -0.72134752044448170368f);                     // exp(-x^2 / 2)
  const float poly = fmaf(fmaf(fmaf(fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f), t, 0.5f * 1.421413741f), t, 0.5f * -0.284496736f), t,
                          0.5f * 0.254829592f) * t;
  q = poly * e;
}
__device__ __forceinline__ float gelu_f(float x) {
  float q, e;
  gelu_q(x, q, e);
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float q, e;
  gelu_q(x, q, e);
  const float cdf = x >= 0.f ? 1.0f - q : q;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// Packed-pair (FFMA2 / FMUL2) forms of the same functions: the polynomial runs on two values per issue slot.
__device__ __forceinline__ float2 fma2(const float2& a, const float2& b, const float2& c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 mul2(const float2& a, const float2& b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(const float2& a, const float2& b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ void gelu_q2(const float2& x, float2& q, float2& e, float2& ax) {
  ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = fma2(splat2(0.3275911f * 0.70710678118654752440f), ax, splat2(1.0f));
  const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  const float2 a = mul2(mul2(x, x), splat2(-0.72134752044448170368f));
  e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
  float2 p = fma2(splat2(0.5f * 1.061405429f), t, splat2(0.5f * -1.453152027f));
  p = fma2(p, t, splat2(0.5f * 1.421413741f));
  p = fma2(p, t, splat2(0.5f * -0.284496736f));
  p = fma2(p, t, splat2(0.5f * 0.254829592f));
  q = mul2(mul2(p, t), e);
}
__device__ __forceinline__ float2 gelu2(const float2& x) {
  float2 q, e, ax;
  gelu_q2(x, q, e, ax);
  const float2 r = make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f));
  return fma2(make_float2(-ax.x, -ax.y), q, r);
}
// GELU and its derivative from one evaluation of (q, e)
__device__ __forceinline__ void gelu_both2(const float2& x, float2& g, float2& dg) {
  float2 q, e, ax;
  gelu_q2(x, q, e, ax);
  g = fma2(make_float2(-ax.x, -ax.y), q, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
  const float2 cdf = make_float2(x.x >= 0.f ? 1.0f - q.x : q.x, x.y >= 0.f ? 1.0f - q.y : q.y);
  dg = fma2(mul2(x, splat2(0.39894228040143267794f)), e, cdf);
}
__device__ __forceinline__ float2 gelu_grad2(const float2& x) {
  float2 q, e, ax;
  gelu_q2(x, q, e, ax);
  const float2 cdf = make_float2(x.x >= 0.f ? 1.0f - q.x : q.x, x.y >= 0.f ? 1.0f - q.y : q.y);
  return fma2(mul2(x, splat2(0.39894228040143267794f)), e, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// Blackwell packed fp32 FMA: two independent FMAs per issue slot (SASS FFMA2).  acc += a * b on (x, y) pairs.
__device__ __forceinline__ void ffma2(float2& acc, const float2& a, const float2& b) {
  unsigned long long c = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(aa), "l"(bb));
  acc = *reinterpret_cast<float2*>(&c);
}
// 8 bf16 (one 16-byte vector) -> 4 fp32 pairs (channel 2q, 2q+1): one shift / one mask per value
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float2* v) {
  v[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
  v[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  v[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
  v[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
}

// 8-wide vector load/store of activations as floats (16B for bf16, 32B for fp32).
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  struct raw_t { float4 a, b; };                                  // the loaded-but-not-yet-used form (software-pipelined loops)
  static __device__ __forceinline__ raw_t load_raw(const float* p) { raw_t r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r; }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  typedef uint4 raw_t;
  static __device__ __forceinline__ raw_t load_raw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  // bf16 -> fp32 is a 16-bit shift: one SHL / one LOP per value instead of the cvt sequence __bfloat1622float2 emits
  static __device__ __forceinline__ void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  static __device__ __forceinline__ void load(const bf16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    unpack(u, v);
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
