// Depthwise k x k (k = 3 or 7) stride-1 'same' convolution on channels-last activations, with the
// element-wise neighbours of the reference fused in:
//   MLP:        y = GELU( dw3x3(h) + b + h )                       (DFormer.py:62-64)
//   Attention:  y = dw7x7(l) + b   /   y = dw7x7(e_fore) + b        (DFormer.py:115,133)
// One thread = 8 channels (one 16-byte bf16 vector) x TW consecutive output pixels of a row, so each
// input vector fetched from L1/L2 is reused for up to TW taps of the sliding window; weights for the
// CTA's 64-channel slab sit in shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int TW = 4;            // output pixels per thread along W
constexpr int CB = 64;           // channels per CTA (8 vectors)
constexpr int STRIPS = 32;       // strips per CTA -> 256 threads

// MODE 0: y = act(conv(x) + bias [+ x])
// MODE 1: y = dy * act'(conv(x) + bias [+ x])                      (dz for the backward pass)
// MODE 2: y = conv_flipped(x) [+ x]          (x = dz; data gradient of the depthwise conv, no bias)
template <typename T, int K, int MODE>
__global__ void __launch_bounds__(CB / 8 * STRIPS) dwconv_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ weight,
                                                                const float* __restrict__ bias, int B, int H, int W, int C, int add_input, int act,
                                                                T* __restrict__ y, T* __restrict__ zout) {
  pdl_sync();
  __shared__ float wsm[K * K][CB];
  __shared__ float bsm[CB];
  const int c_base = blockIdx.y * CB;
  const int cb = min(CB, C - c_base);
  for (int i = threadIdx.x; i < K * K * CB; i += blockDim.x) {
    const int tap = i / CB, c = i % CB;
    float w = 0.f;
    if (c < cb) {
      const int src_tap = (MODE == 2) ? (K * K - 1 - tap) : tap;
      w = weight[(long)(c_base + c) * K * K + src_tap];
    }
    wsm[tap][c] = w;
  }
  for (int i = threadIdx.x; i < CB; i += blockDim.x) bsm[i] = (MODE != 2 && i < cb && bias) ? bias[c_base + i] : 0.f;
  __syncthreads();

  const int cv = threadIdx.x & 7, strip_l = threadIdx.x >> 3;
  const int c0 = cv * 8;
  if (c0 >= cb) return;
  const int strips_per_row = (W + TW - 1) / TW;
  const long total_strips = (long)B * H * strips_per_row;
  constexpr int R = K / 2;
  for (long s = (long)blockIdx.x * STRIPS + strip_l; s < total_strips; s += (long)gridDim.x * STRIPS) {
    const int xs = (int)(s % strips_per_row) * TW;
    const int yy = (int)((s / strips_per_row) % H);
    const int b = (int)(s / ((long)strips_per_row * H));
    float acc[TW][8];
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = bsm[c0 + j];
    float center[TW][8];
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = yy + ky - R;
      if (iy < 0 || iy >= H) continue;
      const T* row = x + (((long)b * H + iy) * W) * C + c_base + c0;
#pragma unroll
      for (int ix_l = 0; ix_l < TW + K - 1; ++ix_l) {
        const int ix = xs + ix_l - R;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        Vec8<T>::load(row + (long)ix * C, v);
        if (ky == R) {
#pragma unroll
          for (int t = 0; t < TW; ++t)
            if (ix_l == t + R) {
#pragma unroll
              for (int j = 0; j < 8; ++j) center[t][j] = v[j];
            }
        }
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          const int kx = ix_l - t;
          if (kx >= 0 && kx < K) {
            const float* wp = &wsm[ky * K + kx][c0];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(v[j], wp[j], acc[t][j]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TW; ++t) {
      const int ox = xs + t;
      if (ox >= W) continue;
      const long off = (((long)b * H + yy) * W + ox) * C + c_base + c0;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = acc[t][j];
        if (add_input) z += center[t][j];
        o[j] = z;
      }
      if (MODE == 0) {
        if (zout) Vec8<T>::store(zout + off, o);
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = gelu_f(o[j]);
        } else if (act == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
      } else if (MODE == 1) {
        float g[8];
        Vec8<T>::load(dy + off, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * (act == 1 ? gelu_grad_f(o[j]) : (act == 2 ? (o[j] > 0.f ? 1.f : 0.f) : 1.f));
      }
      Vec8<T>::store(y + off, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// bf16 fast path: shared-memory tiled version of the kernel above.  A CTA stages a (TY+K-1) x (TX+K-1) pixel halo
// tile of one 64-channel slab (16-byte vectors, coalesced 128-byte rows) in shared memory once, so vertical taps
// are re-used on chip instead of being re-fetched from L2 by other CTAs; each thread then produces TH x TW output
// pixels of one 8-channel vector with a register sliding window along W.
// ---------------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers: all 16-byte tile loads of a CTA are in flight at once (no register staging, no
// per-iteration load->store dependency); out-of-image / out-of-slab vectors are zero-filled (src-size 0).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int TL_TX = 32, TL_TY = 8, TL_TW = 4, TL_TH = 2;    // tile 8 x 32 pixels, thread 2 x 4 pixels -> 256 threads

// halo tile [TY+K-1][TX+K-1][8 vectors] of the 64-channel slab starting at c_base
template <int K>
__device__ __forceinline__ void load_halo_tile_async(uint4* tile, const bf16* __restrict__ x, int b, int ty0, int tx0, int H, int W, int C,
                                                     int c_base, int nv) {
  constexpr int R = K / 2, SW = TL_TX + K - 1, SH = TL_TY + K - 1, N = SH * SW * 8;
#pragma unroll
  for (int it = 0; it < (N + 255) / 256; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < N) {
      const int cv = i & 7, px = (i >> 3) % SW, py = (i >> 3) / SW;
      const int gy = ty0 + py - R, gx = tx0 + px - R;
      const bool ok = cv < nv && gy >= 0 && gy < H && gx >= 0 && gx < W;
      const bf16* src = ok ? x + (((long)b * H + gy) * W + gx) * C + c_base + cv * 8 : x;
      cp_async16(tile + i, src, ok);
    }
  }
}

template <int K, int MODE>
__global__ void __launch_bounds__(256) dwconv_tiled_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ weight,
                                                          const float* __restrict__ bias, int B, int H, int W, int C, int add_input, int act,
                                                          bf16* __restrict__ y, bf16* __restrict__ zout, int tiles_x, int tiles_y) {
  pdl_sync();
  constexpr int R = K / 2, SW = TL_TX + K - 1, SH = TL_TY + K - 1;
  extern __shared__ __align__(16) uint8_t dsm[];
  uint4* tile = reinterpret_cast<uint4*>(dsm);                       // [SH][SW][8]
  float* wsm = reinterpret_cast<float*>(dsm + (size_t)SH * SW * 8 * 16);   // [K*K][64]
  float* bsm = wsm + K * K * 64;                                     // [64]
  const int c_base = blockIdx.y * 64;
  const int cb = min(64, C - c_base);
  const int tile_id = blockIdx.x;
  const int tx0 = (tile_id % tiles_x) * TL_TX, ty0 = ((tile_id / tiles_x) % tiles_y) * TL_TY, b = tile_id / (tiles_x * tiles_y);
  for (int i = threadIdx.x; i < K * K * 64; i += 256) {
    const int tap = i >> 6, c = i & 63;
    float w = 0.f;
    if (c < cb) w = weight[(long)(c_base + c) * K * K + ((MODE == 2) ? (K * K - 1 - tap) : tap)];
    wsm[i] = w;
  }
  if (threadIdx.x < 64) bsm[threadIdx.x] = (MODE != 2 && threadIdx.x < cb && bias) ? bias[c_base + threadIdx.x] : 0.f;
  const int nv = cb >> 3;
  load_halo_tile_async<K>(tile, x, b, ty0, tx0, H, W, C, c_base, nv);
  cp_async_wait_all();
  __syncthreads();
  const int cv = threadIdx.x & 7, sx = (threadIdx.x >> 3) & 7, sy = threadIdx.x >> 6;
  if (cv >= nv) return;
  const int lx = sx * TL_TW, ly = sy * TL_TH;
  float acc[TL_TH][TL_TW][8];
#pragma unroll
  for (int r = 0; r < TL_TH; ++r)
#pragma unroll
    for (int t = 0; t < TL_TW; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][t][j] = bsm[cv * 8 + j];
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    float wk[K][8];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(&wsm[(ky * K + kx) * 64 + cv * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&wsm[(ky * K + kx) * 64 + cv * 8 + 4]);
      wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w; wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
    }
#pragma unroll
    for (int r = 0; r < TL_TH; ++r) {
      const uint4* rowp = tile + ((ly + r + ky) * SW + lx) * 8 + cv;
#pragma unroll
      for (int i = 0; i < TL_TW + K - 1; ++i) {
        const uint4 u = rowp[i * 8];
        float v[8];
        Vec8<bf16>::unpack(u, v);
#pragma unroll
        for (int t = 0; t < TL_TW; ++t) {
          const int kx = i - t;
          if (kx >= 0 && kx < K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][t][j] = fmaf(v[j], wk[kx][j], acc[r][t][j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < TL_TH; ++r) {
    const int oy = ty0 + ly + r;
    if (oy >= H) continue;
#pragma unroll
    for (int t = 0; t < TL_TW; ++t) {
      const int ox = tx0 + lx + t;
      if (ox >= W) continue;
      const long off = (((long)b * H + oy) * W + ox) * C + c_base + cv * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = acc[r][t][j];
      if (add_input) {
        const uint4 u = tile[((ly + r + R) * SW + lx + t + R) * 8 + cv];
        float cval[8];
        Vec8<bf16>::unpack(u, cval);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += cval[j];
      }
      if (MODE == 0) {
        if (zout) Vec8<bf16>::store(zout + off, o);
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = gelu_f(o[j]);
        } else if (act == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
      } else if (MODE == 1) {
        float g[8];
        Vec8<bf16>::load(dy + off, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * (act == 1 ? gelu_grad_f(o[j]) : (act == 2 ? (o[j] > 0.f ? 1.f : 0.f) : 1.f));
      }
      Vec8<bf16>::store(y + off, o);
    }
  }
}

template <int K, int MODE>
int launch_tiled(const bf16* x, const bf16* dy, const float* w, const float* b, int B, int H, int W, int C, int add_input, int act, bf16* y, bf16* zout,
                 cudaStream_t st) {
  constexpr int SW = TL_TX + K - 1, SH = TL_TY + K - 1;
  constexpr int smem = SH * SW * 8 * 16 + (K * K * 64 + 64) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_tiled_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("dwconv smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int tiles_x = dfb_cdiv(W, TL_TX), tiles_y = dfb_cdiv(H, TL_TY);
  dim3 grid((unsigned)((long)B * tiles_x * tiles_y), dfb_cdiv(C, 64));
  dfb_launch(dwconv_tiled_kernel<K, MODE>, grid, 256, smem, st, x, dy, w, b, B, H, W, C, add_input, act, y, zout, tiles_x, tiles_y);
  return dfb_check_launch("dwconv_tiled");
}

// Weight / bias gradients.  dW[c, ky, kx] = sum_p dz[p, c] * x[p + (ky-R, kx-R), c];  db[c] = sum_p dz[p, c].
// Thread = (8-channel vector, kernel row ky, strip lane).  It walks strips of WG_TW output pixels along W with a
// register sliding window over the matching input row, so each 16-byte vector fetched serves up to K taps, and keeps
// its K x 8 partial sums in registers for its whole pixel range; partials are combined across strip lanes in shared
// memory and leave as one atomicAdd per (channel, tap) per CTA.
constexpr int WG_TW = 8;

template <typename T, int K>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const T* __restrict__ dz, const T* __restrict__ x, int B, int H, int W, int C,
                                                          float* __restrict__ dweight, float* __restrict__ dbias, int strips_per_block) {
  pdl_sync();
  constexpr int R = K / 2;
  constexpr int LANES = 256 / (8 * K);             // strip lanes per CTA (k=3: 10, k=7: 4)
  __shared__ float red[K][8][K * 8 + 8];           // [ky][cv][kx*8 + j] (+8: bias partial), strip lanes combined by smem atomics
  for (int i = threadIdx.x; i < K * 8 * (K * 8 + 8); i += blockDim.x) (&red[0][0][0])[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x & 7;
  const int ky = (threadIdx.x >> 3) % K;
  const int sl = threadIdx.x / (8 * K);
  const int c0 = blockIdx.y * 64 + cv * 8;
  const bool active = sl < LANES && c0 < C;
  float acc[K][8];
  float accb[8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) accb[j] = 0.f;
  if (active) {
    const int strips_per_row = (W + WG_TW - 1) / WG_TW;
    const long total_strips = (long)B * H * strips_per_row;
    const long s_begin = (long)blockIdx.x * strips_per_block, s_end = min(total_strips, s_begin + strips_per_block);
    for (long s = s_begin + sl; s < s_end; s += LANES) {
      const int xs = (int)(s % strips_per_row) * WG_TW;
      const int yy = (int)((s / strips_per_row) % H);
      const long b = s / ((long)strips_per_row * H);
      const int iy = yy + ky - R;
      const T* zrow = dz + ((b * H + yy) * W) * C + c0;
      if (ky == R) {
#pragma unroll
        for (int t = 0; t < WG_TW; ++t) {
          if (xs + t < W) {
            float g[8];
            Vec8<T>::load(zrow + (long)(xs + t) * C, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) accb[j] += g[j];
          }
        }
      }
      if (iy < 0 || iy >= H) continue;
      const T* xrow = x + ((b * H + iy) * W) * C + c0;
      float win[K][8];                             // x[iy, xs + t + kx - R] for the current t, kx = 0..K-1
#pragma unroll
      for (int k = 0; k < K - 1; ++k) {
        const int ix = xs + k - R;
        if (ix >= 0 && ix < W) Vec8<T>::load(xrow + (long)ix * C, win[k + 1]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) win[k + 1][j] = 0.f;
        }
      }
#pragma unroll
      for (int t = 0; t < WG_TW; ++t) {
#pragma unroll
        for (int k = 0; k < K - 1; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) win[k][j] = win[k + 1][j];
        const int ix = xs + t + K - 1 - R;
        if (ix >= 0 && ix < W) Vec8<T>::load(xrow + (long)ix * C, win[K - 1]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) win[K - 1][j] = 0.f;
        }
        if (xs + t < W) {
          float g[8];
          Vec8<T>::load(zrow + (long)(xs + t) * C, g);
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g[j], win[k][j], acc[k][j]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[ky][cv][k * 8 + j], acc[k][j]);
    if (ky == R) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[ky][cv][K * 8 + j], accb[j]);
    }
  }
  __syncthreads();
  // combine strip lanes: one thread per (ky, cv, kx*8+j | bias)
  for (int i = threadIdx.x; i < K * 8 * (K * 8 + 8); i += blockDim.x) {
    const int e = i % (K * 8 + 8), rest = i / (K * 8 + 8);
    const int v = rest % 8, kyy = rest / 8;
    const int c = blockIdx.y * 64 + v * 8 + (e & 7);
    if (c >= C) continue;
    const float sum = red[kyy][v][e];
    if (e < K * 8) atomicAdd(dweight + (long)c * K * K + kyy * K + (e >> 3), sum);
    else if (kyy == R) atomicAdd(dbias + c, sum);
  }
}

// bf16 fast path of the weight gradient: a CTA walks a list of 8x32-pixel tiles of one 64-channel slab; per tile the
// x halo tile and the dz tile are staged in shared memory with cp.async, and worker (cv, ky, part) slides a K-wide
// register window along 8-pixel row segments: acc[kx] += dz[p] * x[p + (ky-R, kx-R)].  Partials live in registers
// across all of the CTA's tiles and leave through shared-memory atomics + one global atomicAdd per (channel, tap).
template <int K, bool DB>
__device__ __forceinline__ void wgrad_issue_tile(uint4* xt, uint4* zt, const bf16* __restrict__ dz, const bf16* __restrict__ x, int tile_id, int tiles_x,
                                                 int tiles_y, int H, int W, int C, int c_base, int nv) {
  const int tx0 = (tile_id % tiles_x) * TL_TX, ty0 = ((tile_id / tiles_x) % tiles_y) * TL_TY, b = tile_id / (tiles_x * tiles_y);
  load_halo_tile_async<K>(xt, x, b, ty0, tx0, H, W, C, c_base, nv);
#pragma unroll
  for (int it = 0; it < TL_TY * TL_TX * 8 / 256; ++it) {
    const int i = threadIdx.x + it * 256;
    const int v = i & 7, px = (i >> 3) % TL_TX, py = (i >> 3) / TL_TX;
    const int gy = ty0 + py, gx = tx0 + px;
    const bool ok = v < nv && gy < H && gx < W;
    const bf16* src = ok ? dz + (((long)b * H + gy) * W + gx) * C + c_base + v * 8 : dz;
    cp_async16(zt + i, src, ok);
  }
  cp_async_commit();
}

// DB = true: the CTA double-buffers its tiles (prefetch of tile i+1 overlaps the FMAs of tile i); used where the
// shared-memory footprint allows only one CTA per SM (7x7), otherwise two resident CTAs provide the overlap.
template <int K, bool DB>
__global__ void __launch_bounds__(256) dwconv_wgrad_tiled_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ x, int B, int H, int W, int C,
                                                                float* __restrict__ dweight, float* __restrict__ dbias, int tiles_x, int tiles_y) {
  pdl_sync();
  constexpr int R = K / 2, SW = TL_TX + K - 1, SH = TL_TY + K - 1;
  constexpr int P = 32 / K;                          // parts per kernel row (k=3: 10, k=7: 4)
  constexpr int ITEMS = TL_TY * (TL_TX / 8);         // (row, 8-pixel segment) items per tile = 32
  constexpr int BUF_VECS = SH * SW * 8 + TL_TY * TL_TX * 8;
  extern __shared__ __align__(16) uint8_t dsm[];
  uint4* buf0 = reinterpret_cast<uint4*>(dsm);                                 // per buffer: x halo [SH][SW][8] + dz [TY][TX][8]
  float* red = reinterpret_cast<float*>(buf0 + (DB ? 2 : 1) * BUF_VECS);       // [K][8][K*8+8]
  const int c_base = blockIdx.y * 64;
  const int nv = min(64, C - c_base) >> 3;
  const int cv = threadIdx.x & 7, wk = threadIdx.x >> 3;
  const int ky = wk / P, part = wk % P;
  const bool worker = wk < K * P && cv < nv;
  for (int i = threadIdx.x; i < K * 8 * (K * 8 + 8); i += 256) red[i] = 0.f;
  float acc[K][8], accb[8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) accb[j] = 0.f;
  const int n_tiles = B * tiles_x * tiles_y;
  int cur = 0;
  if (DB) wgrad_issue_tile<K, DB>(buf0, buf0 + SH * SW * 8, dz, x, blockIdx.x, tiles_x, tiles_y, H, W, C, c_base, nv);
  for (int tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    uint4* xt = buf0 + cur * BUF_VECS;
    uint4* zt = xt + SH * SW * 8;
    if (DB) {
      const int next = tile_id + gridDim.x;
      if (next < n_tiles) {
        uint4* nx = buf0 + (cur ^ 1) * BUF_VECS;
        wgrad_issue_tile<K, DB>(nx, nx + SH * SW * 8, dz, x, next, tiles_x, tiles_y, H, W, C, c_base, nv);
        cp_async_wait_group<1>();                    // everything but the prefetch just issued has landed
      } else {
        cp_async_wait_group<0>();
      }
    } else {
      __syncthreads();                               // previous tile fully consumed
      wgrad_issue_tile<K, DB>(xt, zt, dz, x, tile_id, tiles_x, tiles_y, H, W, C, c_base, nv);
      cp_async_wait_group<0>();
    }
    __syncthreads();
    if (worker) {
      for (int item = part; item < ITEMS; item += P) {
        const int row = item / (TL_TX / 8), xs = (item % (TL_TX / 8)) * 8;
        const uint4* xrow = xt + ((row + ky) * SW + xs) * 8 + cv;          // x[row + ky - R][xs + i - R] at index i
        const uint4* zrow = zt + (row * TL_TX + xs) * 8 + cv;
        float win[K][8];
#pragma unroll
        for (int k = 1; k < K; ++k) Vec8<bf16>::unpack(xrow[(k - 1) * 8], win[k]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
#pragma unroll
          for (int k = 0; k < K - 1; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) win[k][j] = win[k + 1][j];
          Vec8<bf16>::unpack(xrow[(t + K - 1) * 8], win[K - 1]);
          float g[8];
          Vec8<bf16>::unpack(zrow[t * 8], g);
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g[j], win[k][j], acc[k][j]);
          if (ky == R) {
#pragma unroll
            for (int j = 0; j < 8; ++j) accb[j] += g[j];
          }
        }
      }
    }
    if (DB) {
      __syncthreads();                               // buffer `cur` is free again before the next prefetch overwrites it
      cur ^= 1;
    }
  }
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[(ky * 8 + cv) * (K * 8 + 8) + k * 8 + j], acc[k][j]);
    if (ky == R) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[(ky * 8 + cv) * (K * 8 + 8) + K * 8 + j], accb[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 8 * (K * 8 + 8); i += 256) {
    const int e = i % (K * 8 + 8), rest = i / (K * 8 + 8);
    const int v = rest % 8, kyy = rest / 8;
    const int c = c_base + v * 8 + (e & 7);
    if (c >= C) continue;
    const float sum = red[i];
    if (e < K * 8) atomicAdd(dweight + (long)c * K * K + kyy * K + (e >> 3), sum);
    else if (kyy == R) atomicAdd(dbias + c, sum);
  }
}

template <int K, bool DB>
int launch_wgrad_tiled(const bf16* dz, const bf16* x, int B, int H, int W, int C, float* dweight, float* dbias, cudaStream_t st) {
  constexpr int SW = TL_TX + K - 1, SH = TL_TY + K - 1;
  constexpr int smem = (DB ? 2 : 1) * (SH * SW * 8 + TL_TY * TL_TX * 8) * 16 + K * 8 * (K * 8 + 8) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_wgrad_tiled_kernel<K, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("dwconv wgrad smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int tiles_x = dfb_cdiv(W, TL_TX), tiles_y = dfb_cdiv(H, TL_TY);
  const int n_tiles = B * tiles_x * tiles_y, chunks = dfb_cdiv(C, 64);
  int gx = dfb_cdiv(148 * (DB ? 1 : 2), chunks);      // one (double-buffered) or two CTAs per SM over all channel slabs
  if (gx > n_tiles) gx = n_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, chunks);
  dfb_launch(dwconv_wgrad_tiled_kernel<K, DB>, grid, 256, smem, st, dz, x, B, H, W, C, dweight, dbias, tiles_x, tiles_y);
  return dfb_check_launch("dwconv_wgrad_tiled");
}

template <typename T, int K, int MODE>
int launch_conv(const T* x, const T* dy, const float* w, const float* b, int B, int H, int W, int C, int add_input, int act, T* y, T* zout,
                cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    return launch_tiled<K, MODE>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), w, b, B, H, W, C, add_input, act,
                                 reinterpret_cast<bf16*>(y), reinterpret_cast<bf16*>(zout), st);
  }
  const long strips = (long)B * H * ((W + TW - 1) / TW);
  long gx = (strips + STRIPS - 1) / STRIPS;
  const long cap = 148L * 32;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, dfb_cdiv(C, CB));
  dfb_launch(dwconv_kernel<T, K, MODE>, grid, CB / 8 * STRIPS, 0, st, x, dy, w, b, B, H, W, C, add_input, act, y, zout);
  return dfb_check_launch("dwconv");
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_dwconv_fwd(const void* x, int dtype, const float* weight, const float* bias, int B, int H, int W, int C, int k, int add_input,
                                 int act, void* y, void* z_out, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "dwconv: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(k == 3 || k == 7, "dwconv: k must be 3 or 7");
  if (k == 7 && dtype == 1 && !add_input && act == 0 && z_out == nullptr) return dfb_dw7_conv(x, weight, bias, B, H, W, C, 0, y, ST);
  DFB_DISPATCH_DTYPE(dtype, T, {
    if (k == 3) return launch_conv<T, 3, 0>((const T*)x, nullptr, weight, bias, B, H, W, C, add_input, act, (T*)y, (T*)z_out, ST);
    return launch_conv<T, 7, 0>((const T*)x, nullptr, weight, bias, B, H, W, C, add_input, act, (T*)y, (T*)z_out, ST);
  });
}

extern "C" int dfb200_dwconv_bwd(const void* dy, const void* x, const void* z, int dtype, const float* weight, const float* bias, int B, int H, int W,
                                 int C, int k, int add_input, int act, void* dz_buf, void* dx, float* dweight, float* dbias, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "dwconv: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(k == 3 || k == 7, "dwconv: k must be 3 or 7");
  DFB_REQUIRE(act == 0 || dz_buf != nullptr, "dwconv_bwd: dz_buf required when act != 0");
  if (k == 7 && dtype == 1 && !add_input && act == 0) {
    if (dx) {
      const int rc = dfb_dw7_conv(dy, weight, nullptr, B, H, W, C, 1, dx, ST);
      if (rc) return rc;
    }
    return dweight ? dfb_dw7_wgrad(dy, x, B, H, W, C, dweight, dbias, ST) : DFB_OK;
  }
  const long total_strips = (long)B * H * ((W + WG_TW - 1) / WG_TW);
  const int cchunks = dfb_cdiv(C, 64);
  int spb = dfb_cdiv(total_strips, dfb_cdiv(148 * 6, cchunks));     // ~6 CTAs per SM overall
  if (spb < 16) spb = 16;
  dim3 wgrid(dfb_cdiv(total_strips, spb), cchunks);
  const int ppb = spb;
  DFB_DISPATCH_DTYPE(dtype, T, {
    const T* dz = (const T*)dy;
    int rc = DFB_OK;
    if (act != 0) {
      if (z != nullptr) {      // saved pre-activation: dz = dy * act'(z) is a single element-wise pass
        rc = dfb200_act_bwd(dy, C, nullptr, 0, z, C, dz_buf, C, dtype, act, (int)((long)B * H * W), C, nullptr, stream);
      } else {
        rc = (k == 3) ? launch_conv<T, 3, 1>((const T*)x, (const T*)dy, weight, bias, B, H, W, C, add_input, act, (T*)dz_buf, nullptr, ST)
                      : launch_conv<T, 7, 1>((const T*)x, (const T*)dy, weight, bias, B, H, W, C, add_input, act, (T*)dz_buf, nullptr, ST);
      }
      if (rc) return rc;
      dz = (const T*)dz_buf;
    }
    if (dx) {
      rc = (k == 3) ? launch_conv<T, 3, 2>(dz, nullptr, weight, nullptr, B, H, W, C, add_input, 0, (T*)dx, nullptr, ST)
                    : launch_conv<T, 7, 2>(dz, nullptr, weight, nullptr, B, H, W, C, add_input, 0, (T*)dx, nullptr, ST);
      if (rc) return rc;
    }
    if (dweight) {
      if constexpr (sizeof(T) == 2) {
        if (k == 3) return launch_wgrad_tiled<3, false>((const bf16*)dz, (const bf16*)x, B, H, W, C, dweight, dbias, ST);
        return launch_wgrad_tiled<7, true>((const bf16*)dz, (const bf16*)x, B, H, W, C, dweight, dbias, ST);
      }
      if (k == 3) dfb_launch(dwconv_wgrad_kernel<T, 3>, wgrid, 256, 0, ST, dz, (const T*)x, B, H, W, C, dweight, dbias, ppb);
      else dfb_launch(dwconv_wgrad_kernel<T, 7>, wgrid, 256, 0, ST, dz, (const T*)x, B, H, W, C, dweight, dbias, ppb);
      return dfb_check_launch("dwconv_wgrad");
    }
    return DFB_OK;
  });
}
