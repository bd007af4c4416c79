// Depthwise 7x7 'same' convolution of the attention branches (DFormer.py:115,133: Attention.conv / e_conv) for bf16
// channels-last activations: forward, data gradient (same kernel, flipped taps) and weight/bias gradient.
//
// FP32-FMA bound (49 taps per output), so the design minimises instructions per FMA rather than bytes:
//   * a CTA owns one channel slab (48 or 64 channels: DFormer-L's widths 48/96/144/288/576 are multiples of 48, the other
//     variants' of 64) and walks TY x TX pixel tiles; the (TY+6) x (TX+6) halo tile arrives by one TMA bulk-tensor copy
//     (out-of-image pixels / channels zero-filled = the convolution's zero padding);
//   * a thread owns 4 channels (8-byte shared-memory vectors) x 8 consecutive output pixels of a row: per kernel row it
//     reads 14 input vectors and 7 weight vectors and issues 8 x 7 x 2 packed FFMA2;
//   * the weight gradient keeps a 7 x 4 register accumulator per (channel group, kernel row) worker and slides a 7-wide
//     register window along 8-pixel segments of the x halo tile against the dz tile.
#include <stdlib.h>

#include "tile_common.cuh"
#include "dfb200_internal.h"

namespace {

using namespace tile;

constexpr int NT = 256;

template <int SW> struct Slab;
template <> struct Slab<48> { static constexpr int NQ = 12, NPG = 21, TX = 40; };      // 21 x 12 = 252 active threads
template <> struct Slab<64> { static constexpr int NQ = 16, NPG = 16, TX = 32; };

// ------------------------------------------------------------------------------------------------ forward / data gradient
template <int SW, int TY, bool FLIP>
__global__ void __launch_bounds__(NT, 3) dw7_conv_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ weight,
                                                        const float* __restrict__ bias, bf16* __restrict__ y, int B, int H, int W, int C, int tiles_x,
                                                        int tiles_y) {
  pdl_sync();
  constexpr int NQ = Slab<SW>::NQ, NPG = Slab<SW>::NPG, TX = Slab<SW>::TX;
  constexpr int PW = TX + 6, PH = TY + 6, TILE_BYTES = PH * PW * SW * 2;
  constexpr int SEGS = TX / 8, NBLK = TY * SEGS;
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* dsm = align128(dsm_raw);
  const uint2* xT = reinterpret_cast<const uint2*>(dsm);
  float* wsm = reinterpret_cast<float*>(dsm + TILE_BYTES);          // [49][SW]
  float* bsm = wsm + 49 * SW;                                        // [SW]
  const int tid = threadIdx.x, cq = tid % NQ, pg = tid / NQ;
  const int c_base = blockIdx.y * SW, c0 = c_base + cq * 4;
  const int n_tiles = B * tiles_x * tiles_y;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 49 * SW; i += NT) {
    const int tap = i / SW, c = i % SW;
    wsm[i] = (c_base + c < C) ? weight[(long)(c_base + c) * 49 + (FLIP ? 48 - tap : tap)] : 0.f;
  }
  for (int i = tid; i < SW; i += NT) bsm[i] = (bias && c_base + i < C) ? bias[c_base + i] : 0.f;
  __syncthreads();
  const float2 b_lo = make_float2(bsm[cq * 4], bsm[cq * 4 + 1]), b_hi = make_float2(bsm[cq * 4 + 2], bsm[cq * 4 + 3]);
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    if (tid == 0) {
      mbar_expect_tx(&bar, TILE_BYTES);
      tma_load_4d(&tmX, &bar, dsm, c_base, tx0 - 3, ty0 - 3, b);
    }
    mbar_wait(&bar, it & 1);
    if (pg < NPG) {
      for (int blk = pg; blk < NBLK; blk += NPG) {
        const int row = blk / SEGS, x0 = (blk % SEGS) * 8;
        float2 acc[8][2];
#pragma unroll
        for (int t = 0; t < 8; ++t) { acc[t][0] = b_lo; acc[t][1] = b_hi; }
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
          float2 w[7][2];
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const float4 wv = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * SW + cq * 4);
            w[kx][0] = make_float2(wv.x, wv.y);
            w[kx][1] = make_float2(wv.z, wv.w);
          }
          const uint2* rp = xT + ((row + ky) * PW + x0) * NQ + cq;
#pragma unroll
          for (int i = 0; i < 14; ++i) {
            float2 lo, hi;
            unpack4(rp[i * NQ], lo, hi);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int kx = i - t;
              if (kx >= 0 && kx < 7) {
                ffma2(acc[t][0], lo, w[kx][0]);
                ffma2(acc[t][1], hi, w[kx][1]);
              }
            }
          }
        }
        const int oy = ty0 + row, ox0 = tx0 + x0;
        bf16* yp = y + (((long)b * H + oy) * W + ox0) * C + c0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint2 o = pack4(acc[t][0], acc[t][1]);
          if (c0 < C && oy < H && ox0 + t < W) *reinterpret_cast<uint2*>(yp + (long)t * C) = o;
        }
      }
    }
    __syncthreads();                                   // tile fully consumed before the next TMA refill
  }
}

// ------------------------------------------------------------------------------------------------ weight / bias gradient
template <int SW, int TY>
__global__ void __launch_bounds__(NT, 2) dw7_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDZ,
                                                         float* __restrict__ dweight, float* __restrict__ dbias, int B, int H, int W, int C, int tiles_x,
                                                         int tiles_y) {
  pdl_sync();
  constexpr int NQ = Slab<SW>::NQ, NPG = Slab<SW>::NPG, TX = Slab<SW>::TX;
  constexpr int PW = TX + 6, PH = TY + 6, X_BYTES = PH * PW * SW * 2, Z_BYTES = TY * TX * SW * 2;
  constexpr int SEGS = TX / 8, ITEMS = TY * SEGS, PARTS = NPG / 7;
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* dsm = align128(dsm_raw);
  const uint2* xT = reinterpret_cast<const uint2*>(dsm);
  const uint2* zT = reinterpret_cast<const uint2*>(dsm + X_BYTES);
  float* red_w = reinterpret_cast<float*>(dsm + X_BYTES + Z_BYTES);     // [SW][49]
  float* red_b = red_w + SW * 49;                                        // [SW]
  const int tid = threadIdx.x, cq = tid % NQ, pg = tid / NQ;
  const int c_base = blockIdx.y * SW;
  const int n_tiles = B * tiles_x * tiles_y;
  const bool worker = pg < 7 * PARTS;
  const int ky = pg / PARTS, part = pg % PARTS;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < SW * 50; i += NT) red_w[i] = 0.f;
  float2 gw[7][2], gb[2];
#pragma unroll
  for (int k = 0; k < 7; ++k) gw[k][0] = gw[k][1] = make_float2(0.f, 0.f);
  gb[0] = gb[1] = make_float2(0.f, 0.f);
  __syncthreads();
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    if (tid == 0) {
      mbar_expect_tx(&bar, X_BYTES + Z_BYTES);
      tma_load_4d(&tmX, &bar, dsm, c_base, tx0 - 3, ty0 - 3, b);
      tma_load_4d(&tmDZ, &bar, dsm + X_BYTES, c_base, tx0, ty0, b);
    }
    mbar_wait(&bar, it & 1);
    if (worker) {
      for (int item = part; item < ITEMS; item += PARTS) {
        const int r = item / SEGS, xs = (item % SEGS) * 8;
        const uint2* xrow = xT + ((r + ky) * PW + xs) * NQ + cq;         // x[r + ky - 3][xs + i - 3] at index i
        const uint2* zrow = zT + (r * TX + xs) * NQ + cq;
        float2 win[7][2];
#pragma unroll
        for (int k = 1; k < 7; ++k) unpack4(xrow[(k - 1) * NQ], win[k][0], win[k][1]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
#pragma unroll
          for (int k = 0; k < 6; ++k) { win[k][0] = win[k + 1][0]; win[k][1] = win[k + 1][1]; }
          unpack4(xrow[(t + 6) * NQ], win[6][0], win[6][1]);
          float2 glo, ghi;
          unpack4(zrow[t * NQ], glo, ghi);
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            ffma2(gw[k][0], glo, win[k][0]);
            ffma2(gw[k][1], ghi, win[k][1]);
          }
          if (ky == 3) { gb[0] = add2(gb[0], glo); gb[1] = add2(gb[1], ghi); }
        }
      }
    }
    __syncthreads();
  }
  if (worker) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float v[4] = {gw[k][0].x, gw[k][0].y, gw[k][1].x, gw[k][1].y};
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&red_w[(cq * 4 + j) * 49 + ky * 7 + k], v[j]);
    }
    if (ky == 3) {
      const float v[4] = {gb[0].x, gb[0].y, gb[1].x, gb[1].y};
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&red_b[cq * 4 + j], v[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < SW * 49; i += NT) {
    const int c = c_base + i / 49;
    if (c < C) atomicAdd(dweight + (long)c * 49 + i % 49, red_w[i]);
  }
  if (dbias && tid < SW && c_base + tid < C) atomicAdd(dbias + c_base + tid, red_b[tid]);
}

// ------------------------------------------------------------------------------------------------ host
int pick_slab(int C) {
  if (C % 64 == 0) return 64;
  if (C % 48 == 0) return 48;
  return dfb_cdiv(C, 48) * 48 < dfb_cdiv(C, 64) * 64 ? 48 : 64;
}

// rows per tile: the static round-robin schedule finishes after ceil(tiles / CTAs) tile-times; take the tile height with
// the shorter makespan (in pixel rows), e.g. stage 0 of DFormer-L (480 tiles of 8 rows on 222 CTAs -> 3 rounds = 24 rows,
// 960 tiles of 4 rows -> 5 rounds = 20 rows)
int pick_ty(int B, int H, int tiles_x, int slots) {
  const int r8 = dfb_cdiv((long)B * tiles_x * dfb_cdiv(H, 8), slots) * (8 + 2), r4 = dfb_cdiv((long)B * tiles_x * dfb_cdiv(H, 4), slots) * (4 + 2);
  return r4 < r8 ? 4 : 8;
}

template <int SW, int TY, bool FLIP>
int launch_conv(const bf16* x, const float* weight, const float* bias, int B, int H, int W, int C, bf16* y, int gx_max, cudaStream_t st) {
  constexpr int TX = Slab<SW>::TX;
  constexpr int smem = (TY + 6) * (TX + 6) * SW * 2 + 50 * SW * 4 + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dw7_conv_kernel<SW, TY, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("dw7_conv smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  CUtensorMap tm;
  int rc = make_map_nhwc(&tm, x, B, H, W, C, TX + 6, TY + 6, SW);
  if (rc) return rc;
  const int tiles_x = dfb_cdiv(W, TX), tiles_y = dfb_cdiv(H, TY), n_tiles = B * tiles_x * tiles_y;
  dim3 grid(gx_max < n_tiles ? gx_max : n_tiles, dfb_cdiv(C, SW));
  dfb_launch(dw7_conv_kernel<SW, TY, FLIP>, grid, NT, smem, st, tm, weight, bias, y, B, H, W, C, tiles_x, tiles_y);
  return dfb_check_launch("dw7_conv");
}

template <int SW, int TY>
int launch_wgrad(const bf16* dz, const bf16* x, int B, int H, int W, int C, float* dweight, float* dbias, int gx_max, cudaStream_t st) {
  constexpr int TX = Slab<SW>::TX;
  constexpr int smem = (TY + 6) * (TX + 6) * SW * 2 + TY * TX * SW * 2 + 50 * SW * 4 + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dw7_wgrad_kernel<SW, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("dw7_wgrad smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  CUtensorMap tmX, tmZ;
  int rc = make_map_nhwc(&tmX, x, B, H, W, C, TX + 6, TY + 6, SW);
  if (rc) return rc;
  rc = make_map_nhwc(&tmZ, dz, B, H, W, C, TX, TY, SW);
  if (rc) return rc;
  const int tiles_x = dfb_cdiv(W, TX), tiles_y = dfb_cdiv(H, TY), n_tiles = B * tiles_x * tiles_y;
  dim3 grid(gx_max < n_tiles ? gx_max : n_tiles, dfb_cdiv(C, SW));
  dfb_launch(dw7_wgrad_kernel<SW, TY>, grid, NT, smem, st, tmX, tmZ, dweight, dbias, B, H, W, C, tiles_x, tiles_y);
  return dfb_check_launch("dw7_wgrad");
}

}  // namespace

// y = dw7x7(x) + bias (flip = 0)   /   dx = dw7x7^T(dz) (flip = 1, no bias)
int dfb_dw7_conv(const void* x, const float* weight, const float* bias, int B, int H, int W, int C, int flip, void* y, cudaStream_t st) {
  const int sw = pick_slab(C), nslab = dfb_cdiv(C, sw);
  int gx = (148 * 3) / nslab;
  if (gx < 1) gx = 1;
  const int ty = pick_ty(B, H, dfb_cdiv(W, sw == 48 ? 40 : 32), gx);
  const bf16* xp = (const bf16*)x;
  bf16* yp = (bf16*)y;
#define GO(SW, TY) (flip ? launch_conv<SW, TY, true>(xp, weight, nullptr, B, H, W, C, yp, gx, st) : launch_conv<SW, TY, false>(xp, weight, bias, B, H, W, C, yp, gx, st))
  if (sw == 48) return ty == 4 ? GO(48, 4) : GO(48, 8);
  return ty == 4 ? GO(64, 4) : GO(64, 8);
#undef GO
}

int dfb_dw7_wgrad(const void* dz, const void* x, int B, int H, int W, int C, float* dweight, float* dbias, cudaStream_t st) {
  const int sw = pick_slab(C), nslab = dfb_cdiv(C, sw);
  int gx = (148 * 2) / nslab;
  if (gx < 1) gx = 1;
  const int ty = pick_ty(B, H, dfb_cdiv(W, sw == 48 ? 40 : 32), gx);
  const bf16* zp = (const bf16*)dz;
  const bf16* xp = (const bf16*)x;
  if (sw == 48) return ty == 4 ? launch_wgrad<48, 4>(zp, xp, B, H, W, C, dweight, dbias, gx, st) : launch_wgrad<48, 8>(zp, xp, B, H, W, C, dweight, dbias, gx, st);
  return ty == 4 ? launch_wgrad<64, 4>(zp, xp, B, H, W, C, dweight, dbias, gx, st) : launch_wgrad<64, 8>(zp, xp, B, H, W, C, dweight, dbias, gx, st);
}
