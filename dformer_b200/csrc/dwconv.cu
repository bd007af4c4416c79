// Depthwise k x k (k = 3 or 7) stride-1 'same' convolution on channels-last activations, with the
// element-wise neighbours of the reference fused in:
//   MLP:        y = GELU( dw3x3(h) + b + h )                       (DFormer.py:62-64)
//   Attention:  y = dw7x7(l) + b   /   y = dw7x7(e_fore) + b        (DFormer.py:115,133)
// One thread = 8 channels (one 16-byte bf16 vector) x TW consecutive output pixels of a row, so each
// input vector fetched from L1/L2 is reused for up to TW taps of the sliding window; weights for the
// CTA's 64-channel slab sit in shared memory.
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
                                                                T* __restrict__ y) {
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
constexpr int TL_TX = 32, TL_TY = 8, TL_TW = 4, TL_TH = 2;    // tile 8 x 32 pixels, thread 2 x 4 pixels -> 256 threads

template <int K, int MODE>
__global__ void __launch_bounds__(256) dwconv_tiled_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ weight,
                                                          const float* __restrict__ bias, int B, int H, int W, int C, int add_input, int act,
                                                          bf16* __restrict__ y, int tiles_x, int tiles_y) {
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
  for (int i = threadIdx.x; i < SH * SW * 8; i += 256) {
    const int cv = i & 7, px = (i >> 3) % SW, py = (i >> 3) / SW;
    const int gy = ty0 + py - R, gx = tx0 + px - R;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (cv < nv && gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((long)b * H + gy) * W + gx) * C + c_base + cv * 8));
    tile[i] = v;
  }
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
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
        float v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hh[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
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
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hh[q]); o[2 * q] += f.x; o[2 * q + 1] += f.y; }
      }
      if (MODE == 0) {
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
int launch_tiled(const bf16* x, const bf16* dy, const float* w, const float* b, int B, int H, int W, int C, int add_input, int act, bf16* y, cudaStream_t st) {
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
  dwconv_tiled_kernel<K, MODE><<<grid, 256, smem, st>>>(x, dy, w, b, B, H, W, C, add_input, act, y, tiles_x, tiles_y);
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

template <typename T, int K, int MODE>
int launch_conv(const T* x, const T* dy, const float* w, const float* b, int B, int H, int W, int C, int add_input, int act, T* y, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    return launch_tiled<K, MODE>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), w, b, B, H, W, C, add_input, act,
                                 reinterpret_cast<bf16*>(y), st);
  }
  const long strips = (long)B * H * ((W + TW - 1) / TW);
  long gx = (strips + STRIPS - 1) / STRIPS;
  const long cap = 148L * 32;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, dfb_cdiv(C, CB));
  dwconv_kernel<T, K, MODE><<<grid, CB / 8 * STRIPS, 0, st>>>(x, dy, w, b, B, H, W, C, add_input, act, y);
  return dfb_check_launch("dwconv");
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_dwconv_fwd(const void* x, int dtype, const float* weight, const float* bias, int B, int H, int W, int C, int k, int add_input,
                                 int act, void* y, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "dwconv: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(k == 3 || k == 7, "dwconv: k must be 3 or 7");
  DFB_DISPATCH_DTYPE(dtype, T, {
    if (k == 3) return launch_conv<T, 3, 0>((const T*)x, nullptr, weight, bias, B, H, W, C, add_input, act, (T*)y, ST);
    return launch_conv<T, 7, 0>((const T*)x, nullptr, weight, bias, B, H, W, C, add_input, act, (T*)y, ST);
  });
}

extern "C" int dfb200_dwconv_bwd(const void* dy, const void* x, int dtype, const float* weight, const float* bias, int B, int H, int W, int C, int k,
                                 int add_input, int act, void* dz_buf, void* dx, float* dweight, float* dbias, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "dwconv: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(k == 3 || k == 7, "dwconv: k must be 3 or 7");
  DFB_REQUIRE(act == 0 || dz_buf != nullptr, "dwconv_bwd: dz_buf required when act != 0");
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
      rc = (k == 3) ? launch_conv<T, 3, 1>((const T*)x, (const T*)dy, weight, bias, B, H, W, C, add_input, act, (T*)dz_buf, ST)
                    : launch_conv<T, 7, 1>((const T*)x, (const T*)dy, weight, bias, B, H, W, C, add_input, act, (T*)dz_buf, ST);
      if (rc) return rc;
      dz = (const T*)dz_buf;
    }
    if (dx) {
      rc = (k == 3) ? launch_conv<T, 3, 2>(dz, nullptr, weight, nullptr, B, H, W, C, add_input, 0, (T*)dx, ST)
                    : launch_conv<T, 7, 2>(dz, nullptr, weight, nullptr, B, H, W, C, add_input, 0, (T*)dx, ST);
      if (rc) return rc;
    }
    if (dweight) {
      if (k == 3) dwconv_wgrad_kernel<T, 3><<<wgrid, 256, 0, ST>>>(dz, (const T*)x, B, H, W, C, dweight, dbias, ppb);
      else dwconv_wgrad_kernel<T, 7><<<wgrid, 256, 0, ST>>>(dz, (const T*)x, B, H, W, C, dweight, dbias, ppb);
      return dfb_check_launch("dwconv_wgrad");
    }
    return DFB_OK;
  });
}
