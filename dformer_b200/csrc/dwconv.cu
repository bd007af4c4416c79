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

// Weight / bias gradients: thread = (channel, tap group); consecutive threads -> consecutive channels.
// dW[c, tap] += sum_p dz[p, c] * x[p + off(tap), c];  db[c] += sum_p dz[p, c].
template <typename T, int K>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const T* __restrict__ dz, const T* __restrict__ x, int B, int H, int W, int C,
                                                          float* __restrict__ dweight, float* __restrict__ dbias, int pix_per_block) {
  constexpr int NT_TAPS = K * K;
  constexpr int GROUPS = 4;
  constexpr int TPG = (NT_TAPS + GROUPS - 1) / GROUPS;
  const int c_l = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + c_l;
  if (c >= C) return;
  float acc[TPG];
#pragma unroll
  for (int i = 0; i < TPG; ++i) acc[i] = 0.f;
  float accb = 0.f;
  const long total = (long)B * H * W;
  const long p0 = (long)blockIdx.x * pix_per_block, p1 = min(total, p0 + pix_per_block);
  constexpr int R = K / 2;
  for (long p = p0; p < p1; ++p) {
    const int px = (int)(p % W), py = (int)((p / W) % H);
    const long bimg = p / ((long)W * H);
    const float g = to_f(dz[p * C + c]);
    if (grp == 0) accb += g;
#pragma unroll
    for (int i = 0; i < TPG; ++i) {
      const int tap = grp * TPG + i;
      if (tap < NT_TAPS) {
        const int ky = tap / K, kx = tap % K;
        const int iy = py + ky - R, ix = px + kx - R;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) acc[i] = fmaf(g, to_f(x[((bimg * H + iy) * W + ix) * C + c]), acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TPG; ++i) {
    const int tap = grp * TPG + i;
    if (tap < NT_TAPS) atomicAdd(dweight + (long)c * NT_TAPS + tap, acc[i]);
  }
  if (grp == 0) atomicAdd(dbias + c, accb);
}

template <typename T, int K, int MODE>
int launch_conv(const T* x, const T* dy, const float* w, const float* b, int B, int H, int W, int C, int add_input, int act, T* y, cudaStream_t st) {
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
  const long total = (long)B * H * W;
  int ppb = dfb_cdiv(total, 148 * 8);
  if (ppb < 64) ppb = 64;
  dim3 wgrid(dfb_cdiv(total, ppb), dfb_cdiv(C, 64));
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
