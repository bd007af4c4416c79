// Global Awareness Attention pieces (DFormer.py:107-108,120-131):
//   pool7      AdaptiveAvgPool2d((7,7)) of cat[LN(x), LN(x_e)] on channels-last tensors
//   gaa        49 pooled queries attend over all H*W pixel keys/values (softmax over pixels)
//   resize     bilinear align_corners=False resampling into a column slice of a wider buffer
#include <string.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

__device__ __forceinline__ int win_start(int i, int n) { return (i * n) / 7; }
__device__ __forceinline__ int win_end(int i, int n) { return ((i + 1) * n + 6) / 7; }

// one CTA per (image, pooled cell): thread = (8-channel vector, pixel lane); the pixel lanes' partial sums meet in shared memory.
// (One thread per output vector walked up to H/7 x W/7 pixels serially with 14 k threads in flight: latency-bound.)
template <typename T>
__global__ void __launch_bounds__(256) pool7_fwd_kernel(const T* __restrict__ xn, int C1, const T* __restrict__ en, int C2, int B, int H, int W,
                                                        T* __restrict__ out) {
  pdl_sync();
  extern __shared__ float psm[];                        // [lanes][nvec * 8]
  const int Ct = C1 + C2, nvec = Ct >> 3;
  const int cell = blockIdx.x % 49, b = blockIdx.x / 49;
  const int py = cell / 7, px = cell % 7;
  const int y0 = win_start(py, H), y1 = win_end(py, H), x0 = win_start(px, W), x1 = win_end(px, W);
  const int ww = x1 - x0, npix = (y1 - y0) * ww;
  const int lanes = max(1, min((int)blockDim.x / nvec, npix));
  const int t = threadIdx.x, v = t % nvec, lane = t / nvec;
  if (lane < lanes && t < lanes * nvec) {
    const int c = v * 8;
    const T* src; int Cs, cc;
    if (c < C1) { src = xn; Cs = C1; cc = c; } else { src = en; Cs = C2; cc = c - C1; }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = lane; p < npix; p += lanes) {
      const int y = y0 + p / ww, x = x0 + p % ww;
      float vv[8];
      Vec8<T>::load(src + (((long)b * H + y) * W + x) * Cs + cc, vv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += vv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) psm[(lane * nvec + v) * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int i = t; i < nvec; i += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += psm[(l * nvec + i) * 8 + j];
    const float inv = 1.f / (float)npix;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    Vec8<T>::store(out + ((long)b * 49 + cell) * Ct + i * 8, acc);
  }
}

// gather form: for maps of >= 7 rows / columns the cells that can contain coordinate v are floor(7 v / n) - 1 .. + 1
template <typename T>
__global__ void pool7_bwd_kernel(const T* __restrict__ dout, int C1, int C2, int B, int H, int W, T* __restrict__ dxn, T* __restrict__ den) {
  pdl_sync();
  const int Ct = C1 + C2, nvec = Ct >> 3;
  const unsigned n = (unsigned)B * H * W * nvec;           // 32-bit index arithmetic (the launcher checks the range)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % nvec) * 8;
    const unsigned pix = i / nvec;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((unsigned)W * H));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int pyc = (y * 7) / H, pxc = (x * 7) / W;
    // maps with fewer than 7 rows / columns repeat pixels over more than three cells: scan all seven there
    const int py_lo = H >= 7 ? max(0, pyc - 1) : 0, py_hi = H >= 7 ? min(6, pyc + 1) : 6;
    const int px_lo = W >= 7 ? max(0, pxc - 1) : 0, px_hi = W >= 7 ? min(6, pxc + 1) : 6;
    for (int py = py_lo; py <= py_hi; ++py) {
      const int y0 = win_start(py, H), y1 = win_end(py, H);
      if (y < y0 || y >= y1) continue;
      for (int px = px_lo; px <= px_hi; ++px) {
        const int x0 = win_start(px, W), x1 = win_end(px, W);
        if (x < x0 || x >= x1) continue;
        float v[8];
        Vec8<T>::load(dout + ((long)b * 49 + py * 7 + px) * Ct + c, v);
        const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], inv, acc[j]);
      }
    }
    if (c < C1) Vec8<T>::store(dxn + (long)pix * C1 + c, acc);
    else Vec8<T>::store(den + (long)pix * C2 + (c - C1), acc);
  }
}

// ------------------------------------------------------------------ bilinear resize
struct Lerp { int i0, i1; float w1; };
__device__ __forceinline__ Lerp lerp_coord(int o, int n_in, int n_out) {
  Lerp l;
  if (n_in == n_out) { l.i0 = o; l.i1 = o; l.w1 = 0.f; return l; }
  const float scale = (float)n_in / (float)n_out;
  float src = ((float)o + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  l.i0 = min((int)src, n_in - 1);
  l.i1 = min(l.i0 + 1, n_in - 1);
  l.w1 = src - (float)l.i0;
  return l;
}

template <typename TI, typename TO, typename IT>          // IT: 32-bit index arithmetic whenever the element count allows it
__global__ void resize_fwd_kernel(const TI* __restrict__ in, int B, int Hi, int Wi, int C, TO* __restrict__ out, int Ho, int Wo, long ldo, int col0) {
  pdl_sync();
  const int nvec = C >> 3;
  const IT n = (IT)B * Ho * Wo * nvec;
  for (IT i = blockIdx.x * (IT)blockDim.x + threadIdx.x; i < n; i += (IT)gridDim.x * blockDim.x) {
    const int c = (int)(i % nvec) * 8;
    const IT pix = i / nvec;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((IT)Wo * Ho));
    const Lerp ly = lerp_coord(oy, Hi, Ho), lx = lerp_coord(ox, Wi, Wo);
    const TI* base = in + (long)b * Hi * Wi * C + c;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    Vec8<TI>::load(base + ((long)ly.i0 * Wi + lx.i0) * C, v00);
    Vec8<TI>::load(base + ((long)ly.i0 * Wi + lx.i1) * C, v01);
    Vec8<TI>::load(base + ((long)ly.i1 * Wi + lx.i0) * C, v10);
    Vec8<TI>::load(base + ((long)ly.i1 * Wi + lx.i1) * C, v11);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float top = (1.f - lx.w1) * v00[j] + lx.w1 * v01[j];
      const float bot = (1.f - lx.w1) * v10[j] + lx.w1 * v11[j];
      o[j] = (1.f - ly.w1) * top + ly.w1 * bot;
    }
    Vec8<TO>::store(out + (long)pix * ldo + col0 + c, o);
  }
}

// gather form of the adjoint: each input pixel scans the output pixels that can reference it
template <typename TO, typename TI>
__global__ void resize_bwd_kernel(const TO* __restrict__ dout, long ldo, int col0, int B, int Hi, int Wi, int C, int Ho, int Wo, TI* __restrict__ din,
                                  int accumulate) {
  pdl_sync();
  const int nvec = C >> 3;
  const long n = (long)B * Hi * Wi * nvec;
  const float ry = (float)Ho / (float)Hi, rx = (float)Wo / (float)Wi;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % nvec) * 8;
    const long pix = i / nvec;
    const int ix = (int)(pix % Wi), iy = (int)((pix / Wi) % Hi), b = (int)(pix / ((long)Wi * Hi));
    int oy_lo = max(0, (int)floorf((iy - 1) * ry) - 1), oy_hi = min(Ho - 1, (int)ceilf((iy + 2) * ry) + 1);
    int ox_lo = max(0, (int)floorf((ix - 1) * rx) - 1), ox_hi = min(Wo - 1, (int)ceilf((ix + 2) * rx) + 1);
    if (iy == 0) oy_lo = 0;
    if (iy == Hi - 1) oy_hi = Ho - 1;
    if (ix == 0) ox_lo = 0;
    if (ix == Wi - 1) ox_hi = Wo - 1;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const Lerp ly = lerp_coord(oy, Hi, Ho);
      const float wy = (ly.i0 == iy ? 1.f - ly.w1 : 0.f) + (ly.i1 == iy ? ly.w1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const Lerp lx = lerp_coord(ox, Wi, Wo);
        const float wx = (lx.i0 == ix ? 1.f - lx.w1 : 0.f) + (lx.i1 == ix ? lx.w1 : 0.f);
        if (wx == 0.f) continue;
        float g[8];
        Vec8<TO>::load(dout + (((long)b * Ho + oy) * Wo + ox) * ldo + col0 + c, g);
        const float w = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
      }
    }
    TI* dst = din + pix * C + c;
    if (accumulate) {
      float old[8];
      Vec8<TI>::load(dst, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += old[j];
    }
    Vec8<TI>::store(dst, acc);
  }
}

// Adjoint for strong up-sampling (7x7 -> HxW): one warp per (input pixel, 8-channel vector); the lanes split the candidate
// output window (hundreds of pixels) and combine with shuffles, instead of one thread walking the whole window.
template <typename TO, typename TI>
__global__ void __launch_bounds__(256) resize_bwd_warp_kernel(const TO* __restrict__ dout, long ldo, int col0, int B, int Hi, int Wi, int C, int Ho, int Wo,
                                                             TI* __restrict__ din, int accumulate) {
  pdl_sync();
  const int nvec = C >> 3;
  const long n = (long)B * Hi * Wi * nvec;
  const float ry = (float)Ho / (float)Hi, rx = (float)Wo / (float)Wi;
  const int lane = threadIdx.x & 31;
  for (long i = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += (long)gridDim.x * (blockDim.x >> 5)) {
    const int c = (int)(i % nvec) * 8;
    const long pix = i / nvec;
    const int ix = (int)(pix % Wi), iy = (int)((pix / Wi) % Hi), b = (int)(pix / ((long)Wi * Hi));
    int oy_lo = max(0, (int)floorf((iy - 1) * ry) - 1), oy_hi = min(Ho - 1, (int)ceilf((iy + 2) * ry) + 1);
    int ox_lo = max(0, (int)floorf((ix - 1) * rx) - 1), ox_hi = min(Wo - 1, (int)ceilf((ix + 2) * rx) + 1);
    if (iy == 0) oy_lo = 0;
    if (iy == Hi - 1) oy_hi = Ho - 1;
    if (ix == 0) ox_lo = 0;
    if (ix == Wi - 1) ox_hi = Wo - 1;
    const int nx = ox_hi - ox_lo + 1, total = (oy_hi - oy_lo + 1) * nx;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = lane; t < total; t += 32) {
      const int oy = oy_lo + t / nx, ox = ox_lo + t % nx;
      const Lerp ly = lerp_coord(oy, Hi, Ho), lx = lerp_coord(ox, Wi, Wo);
      const float wy = (ly.i0 == iy ? 1.f - ly.w1 : 0.f) + (ly.i1 == iy ? ly.w1 : 0.f);
      const float wx = (lx.i0 == ix ? 1.f - lx.w1 : 0.f) + (lx.i1 == ix ? lx.w1 : 0.f);
      const float w = wy * wx;
      if (w == 0.f) continue;
      float g[8];
      Vec8<TO>::load(dout + (((long)b * Ho + oy) * Wo + ox) * ldo + col0 + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      TI* dst = din + pix * C + c;
      if (accumulate) {
        float old[8];
        Vec8<TI>::load(dst, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += old[j];
      }
      Vec8<TI>::store(dst, acc);
    }
  }
}

inline int ew_grid(long n) {
  long b = (n + 255) / 256;
  if (b < 1) b = 1;
  const long cap = 148L * 16;
  return (int)(b > cap ? cap : b);
}

// dS = P * (dP - rowsum(dP * P)), in place on dP
__global__ void softmax_bwd_inplace_kernel(float* __restrict__ dP, const float* __restrict__ P, int rows, int cols) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* g = dP + (long)row * cols;
  const float* p = P + (long)row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(g[c], p[c], s);
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) g[c] = p[c] * (g[c] - s);
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_pool7_fwd(const void* xn, int C1, const void* en, int C2, int dtype, int B, int H, int W, void* out, void* stream) {
  DFB_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "pool7: channels must be multiples of 8");
  DFB_DISPATCH_DTYPE(dtype, T, {
    const int nvec = (C1 + C2) / 8;
    DFB_REQUIRE(nvec <= 256, "pool7_fwd: %d channels exceed one CTA (2048)", C1 + C2);
    const int lanes = 256 / nvec;
    dfb_launch(pool7_fwd_kernel<T>, B * 49, 256, (size_t)lanes * nvec * 8 * sizeof(float), ST, (const T*)xn, C1, (const T*)en, C2, B, H, W, (T*)out);
  });
  return dfb_check_launch("pool7_fwd");
}
extern "C" int dfb200_pool7_bwd(const void* dout, int C1, int C2, int dtype, int B, int H, int W, void* dxn, void* den, void* stream) {
  DFB_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "pool7: channels must be multiples of 8");
  DFB_DISPATCH_DTYPE(dtype, T, {
    DFB_REQUIRE((long)B * H * W * (C1 + C2) / 8 < (1L << 31), "pool7_bwd: problem too large for 32-bit indexing");
    dfb_launch(pool7_bwd_kernel<T>, ew_grid((long)B * H * W * (C1 + C2) / 8), 256, 0, ST, (const T*)dout, C1, C2, B, H, W, (T*)dxn, (T*)den);
  });
  return dfb_check_launch("pool7_bwd");
}

extern "C" int dfb200_resize_fwd(const void* in, int in_dtype, int B, int Hi, int Wi, int C, void* out, int out_dtype, int Ho, int Wo, long ldo, int col0,
                                 void* stream) {
  DFB_REQUIRE(C % 8 == 0 && ldo % 8 == 0 && col0 % 8 == 0, "resize: C, ldo, col0 must be multiples of 8");
  const long nfw = (long)B * Ho * Wo * C / 8;
  const int g = ew_grid(nfw);
#define L(TI, TO) do { if (nfw < (1L << 30)) dfb_launch(resize_fwd_kernel<TI, TO, unsigned>, g, 256, 0, ST, (const TI*)in, B, Hi, Wi, C, (TO*)out, Ho, Wo, ldo, col0); \
                       else dfb_launch(resize_fwd_kernel<TI, TO, long>, g, 256, 0, ST, (const TI*)in, B, Hi, Wi, C, (TO*)out, Ho, Wo, ldo, col0); } while (0)
  if (in_dtype == 0 && out_dtype == 0) L(float, float);
  else if (in_dtype == 0 && out_dtype == 1) L(float, bf16);
  else if (in_dtype == 1 && out_dtype == 1) L(bf16, bf16);
  else if (in_dtype == 1 && out_dtype == 0) L(bf16, float);
  else { dfb_set_error("resize: bad dtypes"); return DFB_ERR_ARG; }
#undef L
  return dfb_check_launch("resize_fwd");
}
// Block-cooperative adjoint for large up-sampling ratios (the 7x7 attention map -> H x W, the stage-3 feature map -> 60 x 80): one CTA
// per input pixel.  The separable bilinear weights of its candidate output rows / columns are computed once into shared memory
// (no per-candidate coordinate arithmetic), thread = (8-channel vector, pixel lane) walks the candidates with one 16-byte load and
// 8 FMAs each, and the pixel lanes meet in shared memory.
constexpr int RB_MAXR = 96;          // max candidate rows / columns per input pixel handled by this kernel
template <typename TO, typename TI>
__global__ void __launch_bounds__(256) resize_bwd_block_kernel(const TO* __restrict__ dout, long ldo, int col0, int B, int Hi, int Wi, int C, int Ho, int Wo,
                                                              TI* __restrict__ din, int accumulate) {
  pdl_sync();
  __shared__ float wys[RB_MAXR], wxs[RB_MAXR];
  __shared__ float red[256 * 8];
  const int nvec = C >> 3;
  const long pix = blockIdx.x;
  const int ix = (int)(pix % Wi), iy = (int)((pix / Wi) % Hi), b = (int)(pix / ((long)Wi * Hi));
  const float ry = (float)Ho / (float)Hi, rx = (float)Wo / (float)Wi;
  int oy_lo = max(0, (int)floorf((iy - 1) * ry) - 1), oy_hi = min(Ho - 1, (int)ceilf((iy + 2) * ry) + 1);
  int ox_lo = max(0, (int)floorf((ix - 1) * rx) - 1), ox_hi = min(Wo - 1, (int)ceilf((ix + 2) * rx) + 1);
  if (iy == 0) oy_lo = 0;
  if (iy == Hi - 1) oy_hi = Ho - 1;
  if (ix == 0) ox_lo = 0;
  if (ix == Wi - 1) ox_hi = Wo - 1;
  const int ny = oy_hi - oy_lo + 1, nx = ox_hi - ox_lo + 1;
  for (int i = threadIdx.x; i < ny; i += 256) {
    const Lerp l = lerp_coord(oy_lo + i, Hi, Ho);
    wys[i] = (l.i0 == iy ? 1.f - l.w1 : 0.f) + (l.i1 == iy ? l.w1 : 0.f);
  }
  for (int i = threadIdx.x; i < nx; i += 256) {
    const Lerp l = lerp_coord(ox_lo + i, Wi, Wo);
    wxs[i] = (l.i0 == ix ? 1.f - l.w1 : 0.f) + (l.i1 == ix ? l.w1 : 0.f);
  }
  __syncthreads();
  const int total = ny * nx;
  for (int v0 = 0; v0 < nvec; v0 += 256) {                     // channel-vector chunks (one for every DFormer width)
    const int nv = min(256, nvec - v0), lanes = 256 / nv;
    const int v = threadIdx.x % nv, lane = threadIdx.x / nv;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (lane < lanes) {
      const TO* base = dout + (long)b * Ho * Wo * ldo + col0 + (v0 + v) * 8;
      for (int t = lane; t < total; t += lanes) {
        const int yy = t / nx, xx = t - yy * nx;
        const float w = wys[yy] * wxs[xx];
        if (w == 0.f) continue;
        float g[8];
        Vec8<TO>::load(base + ((long)(oy_lo + yy) * Wo + ox_lo + xx) * ldo, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    if (lane == 0) {
      for (int q = 1; q < lanes; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += red[(q * nv + v) * 8 + j];
      TI* dst = din + pix * C + (v0 + v) * 8;
      if (accumulate) {
        float old[8];
        Vec8<TI>::load(dst, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += old[j];
      }
      Vec8<TI>::store(dst, acc);
    }
    __syncthreads();
  }
}

extern "C" int dfb200_resize_bwd(const void* dout, int out_dtype, long ldo, int col0, int B, int Hi, int Wi, int C, int Ho, int Wo, void* din, int in_dtype,
                                 int accumulate, void* stream) {
  DFB_REQUIRE(C % 8 == 0 && ldo % 8 == 0 && col0 % 8 == 0, "resize: C, ldo, col0 must be multiples of 8");
  const long work = (long)B * Hi * Wi * C / 8;
  const bool coop = ((long)Ho * Wo >= 16L * Hi * Wi);              // >= 4x up-sampling per axis: hundreds of candidates per input pixel
  const int g = coop ? (int)min(work / 8 + 1, 148L * 16) : ew_grid(work);
  // candidate rows / columns per input pixel: 3 * ratio + 4 at most (see the range computation in the kernels)
  const bool block_ok = (3L * Ho / Hi + 5 <= RB_MAXR) && (3L * Wo / Wi + 5 <= RB_MAXR) && (long)B * Hi * Wi < (1L << 31);
#define L(TO, TI)                                                                                                                        \
  do {                                                                                                                                   \
    if (coop && block_ok) dfb_launch(resize_bwd_block_kernel<TO, TI>, (unsigned)((long)B * Hi * Wi), 256, 0, ST, (const TO*)dout, ldo, col0, B, Hi, Wi, C, Ho, Wo, (TI*)din, accumulate); \
    else if (coop) dfb_launch(resize_bwd_warp_kernel<TO, TI>, g, 256, 0, ST, (const TO*)dout, ldo, col0, B, Hi, Wi, C, Ho, Wo, (TI*)din, accumulate); \
    else dfb_launch(resize_bwd_kernel<TO, TI>, g, 256, 0, ST, (const TO*)dout, ldo, col0, B, Hi, Wi, C, Ho, Wo, (TI*)din, accumulate);           \
  } while (0)
  if (in_dtype == 0 && out_dtype == 0) L(float, float);
  else if (in_dtype == 0 && out_dtype == 1) L(bf16, float);
  else if (in_dtype == 1 && out_dtype == 1) L(bf16, bf16);
  else if (in_dtype == 1 && out_dtype == 0) L(float, bf16);
  else { dfb_set_error("resize: bad dtypes"); return DFB_ERR_ARG; }
#undef L
  return dfb_check_launch("resize_bwd");
}

// ------------------------------------------------------------------ GAA (batched SIMT GEMMs + row softmax)
static dfb200_gemm_args gaa_args(int B, int heads) {
  dfb200_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.batch = B; g.batch_inner = heads;
  g.backend = DFB200_BACKEND_SIMT; g.splitk = 1;
  g.act_col_start = 0;
  return g;
}

extern "C" int dfb200_gaa_fwd(const void* m, const void* kv, int dtype, int B, int HW, int heads, int d, float* out, float* probs, void* stream) {
  const int Cp = heads * d;
  const size_t es = dtype == 1 ? 2 : 4;
  // S = scale * Q K^T   (per (b, head):  [49, d] x [HW, d]^T)
  dfb200_gemm_args g = gaa_args(B, heads);
  g.A = m; g.lda = Cp; g.strideA = 49L * Cp; g.strideA_in = d; g.a_dtype = dtype;
  g.B = kv; g.ldb = 2L * Cp; g.strideB = (long)HW * 2 * Cp; g.strideB_in = d; g.b_dtype = dtype; g.transB = 1;
  g.C = probs; g.ldc = HW; g.strideC = (long)heads * 49 * HW; g.strideC_in = 49L * HW; g.out_dtype = 0;
  g.M = 49; g.N = HW; g.K = d; g.alpha = 1.0f / sqrtf((float)d);
  int rc = dfb_gemm_simt(g, ST);
  if (rc) return rc;
  rc = dfb200_softmax_rows(probs, B * heads * 49, HW, probs, stream);
  if (rc) return rc;
  // O = P V  ([49, HW] x [HW, d]) -> out[b, :, head*d : (head+1)*d]
  g = gaa_args(B, heads);
  g.A = probs; g.lda = HW; g.strideA = (long)heads * 49 * HW; g.strideA_in = 49L * HW; g.a_dtype = 0;
  g.B = (const char*)kv + (size_t)Cp * es; g.ldb = 2L * Cp; g.strideB = (long)HW * 2 * Cp; g.strideB_in = d; g.b_dtype = dtype; g.transB = 0;
  g.C = out; g.ldc = Cp; g.strideC = 49L * Cp; g.strideC_in = d; g.out_dtype = 0;
  g.M = 49; g.N = d; g.K = HW;
  if (HW >= 1024) {   // long reduction, tiny output: split it and accumulate with fp32 atomics
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * 49 * Cp, ST);
    g.splitk = dfb_cdiv(HW, 256); g.accumulate = 1;
  }
  return dfb_gemm_simt(g, ST);
}

extern "C" int dfb200_gaa_bwd(const float* dout, const void* m, const void* kv, const float* probs, int dtype, int B, int HW, int heads, int d, float* dm,
                              void* dkv, float* scratch, void* stream) {
  const int Cp = heads * d;
  const size_t es = dtype == 1 ? 2 : 4;
  const float scale = 1.0f / sqrtf((float)d);
  float* dP = scratch;
  // dV = P^T dO  -> dkv[:, Cp + head*d ...]
  dfb200_gemm_args g = gaa_args(B, heads);
  g.A = probs; g.lda = HW; g.strideA = (long)heads * 49 * HW; g.strideA_in = 49L * HW; g.a_dtype = 0; g.transA = 1;
  g.B = dout; g.ldb = Cp; g.strideB = 49L * Cp; g.strideB_in = d; g.b_dtype = 0; g.transB = 0;
  g.C = (char*)dkv + (size_t)Cp * es; g.ldc = 2L * Cp; g.strideC = (long)HW * 2 * Cp; g.strideC_in = d; g.out_dtype = dtype;
  g.M = HW; g.N = d; g.K = 49;
  int rc = dfb_gemm_simt(g, ST);
  if (rc) return rc;
  // dP = dO V^T
  g = gaa_args(B, heads);
  g.A = dout; g.lda = Cp; g.strideA = 49L * Cp; g.strideA_in = d; g.a_dtype = 0;
  g.B = (const char*)kv + (size_t)Cp * es; g.ldb = 2L * Cp; g.strideB = (long)HW * 2 * Cp; g.strideB_in = d; g.b_dtype = dtype; g.transB = 1;
  g.C = dP; g.ldc = HW; g.strideC = (long)heads * 49 * HW; g.strideC_in = 49L * HW; g.out_dtype = 0;
  g.M = 49; g.N = HW; g.K = d;
  rc = dfb_gemm_simt(g, ST);
  if (rc) return rc;
  // dS = P * (dP - rowsum(dP*P))   (in place)
  dfb_launch(softmax_bwd_inplace_kernel, dfb_cdiv(B * heads * 49, 8), 256, 0, ST, dP, probs, B * heads * 49, HW);
  rc = dfb_check_launch("gaa softmax bwd");
  if (rc) return rc;
  // dQ = scale * dS K
  g = gaa_args(B, heads);
  g.A = dP; g.lda = HW; g.strideA = (long)heads * 49 * HW; g.strideA_in = 49L * HW; g.a_dtype = 0;
  g.B = kv; g.ldb = 2L * Cp; g.strideB = (long)HW * 2 * Cp; g.strideB_in = d; g.b_dtype = dtype; g.transB = 0;
  g.C = dm; g.ldc = Cp; g.strideC = 49L * Cp; g.strideC_in = d; g.out_dtype = 0;
  g.M = 49; g.N = d; g.K = HW; g.alpha = scale;
  if (HW >= 1024) {
    cudaMemsetAsync(dm, 0, sizeof(float) * (size_t)B * 49 * Cp, ST);
    g.splitk = dfb_cdiv(HW, 256); g.accumulate = 1;
  }
  rc = dfb_gemm_simt(g, ST);
  if (rc) return rc;
  // dK = scale * dS^T Q -> dkv[:, head*d ...]
  g = gaa_args(B, heads);
  g.A = dP; g.lda = HW; g.strideA = (long)heads * 49 * HW; g.strideA_in = 49L * HW; g.a_dtype = 0; g.transA = 1;
  g.B = m; g.ldb = Cp; g.strideB = 49L * Cp; g.strideB_in = d; g.b_dtype = dtype; g.transB = 0;
  g.C = dkv; g.ldc = 2L * Cp; g.strideC = (long)HW * 2 * Cp; g.strideC_in = d; g.out_dtype = dtype;
  g.M = HW; g.N = d; g.K = 49; g.alpha = scale;
  return dfb_gemm_simt(g, ST);
}
