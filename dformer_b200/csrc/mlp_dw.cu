// The middle of the DFormer MLP (DFormer.py:62-64) as HBM-streaming kernels for bf16 channels-last activations:
//
//   forward   u  = GELU( dw3x3(h) + b + h )         reads h, writes u (+ gp = GELU'(z) in training)        2 (3) passes
//   backward  dz = du * gp                          (training default: GELU'(z) kept by the forward pass)
//             dh = dz + dw3x3^T(dz)                 reads du, gp, h, writes dh                              4 passes
//             dW[c,ky,kx] += sum_p dz[p,c] h[p+(ky-1,kx-1),c]   db[c] += sum_p dz[p,c]   dfc1_b[c] += sum_p dh[p,c]
//   (mlp_dw_bwd_kernel: same outputs with GELU'(dw3x3(h) + b + h) recomputed on a one-pixel halo ring instead of read: 3 passes,
//    but FMA-bound -- kept for callers that did not keep gp.)
//
// The unfused chain -- save z, act_bwd, data-gradient conv, weight-gradient conv, column sum -- moved 10 passes.
//
// A CTA owns one 64-channel slab and walks TY x TX pixel tiles of it.  Halo tiles are brought in by TMA
// (cp.async.bulk.tensor.4d over [B, H, W, C]; out-of-image pixels and channels are zero-filled by the copy engine, which
// is exactly the zero padding of the convolution) and completion is signalled on an mbarrier.  A thread owns one
// 4-channel group (8-byte shared-memory vectors; 16 threads cover one pixel's 128-byte slab row, conflict-free) and
// keeps its 9 x 4 filter taps in registers for the whole kernel, the centre tap with the "+ h" folded in.  All FMAs are
// packed fp32 pairs (FFMA2).  dz lives only in shared memory (bf16, written in place over the du tile).
#include "tile_common.cuh"
#include "dfb200_internal.h"

namespace {

using namespace tile;

constexpr int NT = 256;            // threads per CTA
constexpr int NQ = 16;             // 4-channel groups per 64-channel slab
constexpr int NPG = NT / NQ;       // pixel groups (= workers per channel group)
constexpr int PIX_BYTES = 128;     // one pixel of a slab in shared memory

// acc[r][t] += sum_{ky,kx} w[tap(ky,kx)] * tile[row0 + r + ky][col0 + t + kx]   for this thread's 4 channels;
// every input vector is fetched once and fanned out to the (up to 9) outputs it contributes to.
template <int ROWS, int PX, bool FLIP>
__device__ __forceinline__ void conv3x3_block(const uint2* __restrict__ tile, int pitch, int row0, int col0, int cq, const float2 (&w)[9][2],
                                              float2 (&acc)[ROWS][PX][2]) {
#pragma unroll
  for (int rr = 0; rr < ROWS + 2; ++rr) {
    const uint2* rp = tile + ((row0 + rr) * pitch + col0) * NQ + cq;
#pragma unroll
    for (int i = 0; i < PX + 2; ++i) {
      float2 lo, hi;
      unpack4(rp[i * NQ], lo, hi);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int ky = rr - r;
        if (ky < 0 || ky > 2) continue;
#pragma unroll
        for (int t = 0; t < PX; ++t) {
          const int kx = i - t;
          if (kx < 0 || kx > 2) continue;
          const int tap = FLIP ? 8 - (ky * 3 + kx) : ky * 3 + kx;
          ffma2(acc[r][t][0], lo, w[tap][0]);
          ffma2(acc[r][t][1], hi, w[tap][1]);
        }
      }
    }
  }
}

// filter taps of this thread's 4 channels -> registers; the residual "+ h" (and "+ dz" of the adjoint) is the centre tap
__device__ __forceinline__ void load_taps(const float* __restrict__ weight, const float* __restrict__ bias, int c0, int C, float2 (&w)[9][2], float2 (&b)[2]) {
  float wv[4][9], bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool ok = c0 + j < C;
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[j][t] = ok ? weight[(long)(c0 + j) * 9 + t] : 0.f;
    wv[j][4] += 1.0f;
    bv[j] = (ok && bias) ? bias[c0 + j] : 0.f;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    w[t][0] = make_float2(wv[0][t], wv[1][t]);
    w[t][1] = make_float2(wv[2][t], wv[3][t]);
  }
  b[0] = make_float2(bv[0], bv[1]);
  b[1] = make_float2(bv[2], bv[3]);
}


// ------------------------------------------------------------------------------------------------ forward
template <int TX, int TY>
__global__ void __launch_bounds__(NT, 2) mlp_dw_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const float* __restrict__ weight,
                                                          const float* __restrict__ bias, bf16* __restrict__ u, bf16* __restrict__ gp, int B, int H, int W,
                                                          int C, int tiles_x, int tiles_y) {
  pdl_sync();
  constexpr int PW = TX + 2, PH = TY + 2, STAGE_BYTES = PH * PW * PIX_BYTES;
  constexpr int BR = TY / 2, BC = TX / 4;
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  uint8_t* dsm = align128(dsm_raw);
  const int tid = threadIdx.x, cq = tid & (NQ - 1), pg = tid >> 4;
  const int c_base = blockIdx.y * 64;
  const int c0 = c_base + cq * 4;
  const int n_tiles = B * tiles_x * tiles_y;
  const long row_stride = (long)W * C;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2 w[9][2], bq[2];
  load_taps(weight, bias, c0, C, w, bq);
  __syncthreads();
  auto issue = [&](int tile, int s) {
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    mbar_expect_tx(&bar[s], STAGE_BYTES);
    tma_load_4d(&tmH, &bar[s], dsm + s * STAGE_BYTES, c_base, tx0 - 1, ty0 - 1, b);
  };
  if (tid == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, s ^ 1);     // stage s^1 was released by the barrier that ended iteration it-1
    mbar_wait(&bar[s], (it >> 1) & 1);
    const uint2* hT = reinterpret_cast<const uint2*>(dsm + s * STAGE_BYTES);
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    for (int blk = pg; blk < BR * BC; blk += NPG) {
      const int r0 = (blk / BC) * 2, x0 = (blk % BC) * 4;
      float2 acc[2][4][2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[r][t][0] = bq[0]; acc[r][t][1] = bq[1]; }
      conv3x3_block<2, 4, false>(hT, PW, r0, x0, cq, w, acc);
      const int oy0 = ty0 + r0, ox0 = tx0 + x0;
      const long off0 = (((long)b * H + oy0) * W + ox0) * C + c0;
      if (gp) {                                        // training: GELU'(z) is kept (bf16) so that the backward pass is a pure stream
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 g0, g1, d0, d1;
            gelu_both2(acc[r][t][0], g0, d0);
            gelu_both2(acc[r][t][1], g1, d1);
            if (c0 < C && oy0 + r < H && ox0 + t < W) {
              *reinterpret_cast<uint2*>(u + off0 + (long)r * row_stride + t * C) = pack4(g0, g1);
              *reinterpret_cast<uint2*>(gp + off0 + (long)r * row_stride + t * C) = pack4(d0, d1);
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint2 o = pack4(gelu2(acc[r][t][0]), gelu2(acc[r][t][1]));
            if (c0 < C && oy0 + r < H && ox0 + t < W) *reinterpret_cast<uint2*>(u + off0 + (long)r * row_stride + t * C) = o;
          }
        }
      }
    }
    __syncthreads();                                  // stage s fully consumed before the next iteration refills it
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int TX, int TY>
struct BwdGeom {
  static constexpr int HP = TX + 6, HR = TY + 4;      // h tile: origin (ty0-2, tx0-2); two spare columns feed the padded last dz segment
  static constexpr int ZP = TX + 2, ZR = TY + 2;      // du -> dz tile: origin (ty0-1, tx0-1)
  static constexpr int H_BYTES = HR * HP * PIX_BYTES, Z_BYTES = ZR * ZP * PIX_BYTES;
  static constexpr int RED_FLOATS = 64 * 9 + 64 + 64;
  static constexpr int SMEM = H_BYTES + Z_BYTES + RED_FLOATS * 4 + 128;
};

template <int TX, int TY>
__global__ void __launch_bounds__(NT, 2) mlp_dw_bwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmDU,
                                                          const float* __restrict__ weight, const float* __restrict__ bias, bf16* __restrict__ dh,
                                                          float* __restrict__ dweight, float* __restrict__ dbias, float* __restrict__ dh_colsum, int B,
                                                          int H, int W, int C, int tiles_x, int tiles_y) {
  pdl_sync();
  using G = BwdGeom<TX, TY>;
  constexpr int HP = G::HP, ZP = G::ZP, ZR = G::ZR;
  constexpr int SEGA = (ZP + 3) / 4;                  // 4-pixel segments per dz row (the last one is partly padding)
  constexpr int BR = TY / 2, BC = TX / 4;
  constexpr int PARTS = NPG / 3;                      // weight-gradient workers per kernel row (5)
  constexpr int ITEMS = TY * (TX / 8);                // (row, 8-pixel segment) items per tile
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* dsm = align128(dsm_raw);
  uint2* hT = reinterpret_cast<uint2*>(dsm);
  uint2* zT = reinterpret_cast<uint2*>(dsm + G::H_BYTES);
  float* red_w = reinterpret_cast<float*>(dsm + G::H_BYTES + G::Z_BYTES);     // [64][9]
  float* red_b = red_w + 64 * 9;                                                 // [64] sum dz
  float* red_c = red_b + 64;                                                     // [64] sum dh
  const int tid = threadIdx.x, cq = tid & (NQ - 1), pg = tid >> 4;
  const int c_base = blockIdx.y * 64;
  const int c0 = c_base + cq * 4;
  const int n_tiles = B * tiles_x * tiles_y;
  const long row_stride = (long)W * C;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < G::RED_FLOATS; i += NT) red_w[i] = 0.f;
  float2 w[9][2], bq[2];
  load_taps(weight, bias, c0, C, w, bq);
  const int ky = pg / PARTS, part = pg % PARTS;       // weight-gradient role of this thread
  const bool wg_worker = pg < 3 * PARTS;
  float2 gw[3][2], gb[2], gc[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) gw[k][0] = gw[k][1] = make_float2(0.f, 0.f);
  gb[0] = gb[1] = gc[0] = gc[1] = make_float2(0.f, 0.f);
  __syncthreads();
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    if (tid == 0) {
      mbar_expect_tx(&bar, G::H_BYTES + G::Z_BYTES);
      tma_load_4d(&tmH, &bar, hT, c_base, tx0 - 2, ty0 - 2, b);
      tma_load_4d(&tmDU, &bar, zT, c_base, tx0 - 1, ty0 - 1, b);
    }
    mbar_wait(&bar, it & 1);
    // ---- phase A: dz = du * GELU'(conv(h) + b + h) on the tile plus a one-pixel ring, in place over du
    for (int item = pg; item < ZR * SEGA; item += NPG) {
      const int row = item / SEGA, col0 = (item % SEGA) * 4;
      float2 acc[1][4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t) { acc[0][t][0] = bq[0]; acc[0][t][1] = bq[1]; }
      conv3x3_block<1, 4, false>(hT, HP, row, col0, cq, w, acc);
      uint2* zp = zT + (row * ZP + col0) * NQ + cq;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (col0 + t < ZP) {
          float2 lo, hi;
          unpack4(zp[t * NQ], lo, hi);
          zp[t * NQ] = pack4(mul2(lo, gelu_grad2(acc[0][t][0])), mul2(hi, gelu_grad2(acc[0][t][1])));
        }
      }
    }
    __syncthreads();
    // ---- phase B: dh = dz + conv^T(dz)  (+ running column sum of dh = gradient of the fc1 bias)
    for (int blk = pg; blk < BR * BC; blk += NPG) {
      const int r0 = (blk / BC) * 2, x0 = (blk % BC) * 4;
      float2 acc[2][4][2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[r][t][0] = acc[r][t][1] = make_float2(0.f, 0.f);
      conv3x3_block<2, 4, true>(zT, ZP, r0, x0, cq, w, acc);
      const int oy0 = ty0 + r0, ox0 = tx0 + x0;
      bf16* dp = dh + (((long)b * H + oy0) * W + ox0) * C + c0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint2 o = pack4(acc[r][t][0], acc[r][t][1]);
          if (c0 < C && oy0 + r < H && ox0 + t < W) {
            *reinterpret_cast<uint2*>(dp + (long)r * row_stride + t * C) = o;
            float2 lo, hi;
            unpack4(o, lo, hi);                        // the bias gradient sums the values fc1's weight gradient will read
            gc[0] = add2(gc[0], lo);
            gc[1] = add2(gc[1], hi);
          }
        }
      }
    }
    // ---- phase C: dW[ky][kx] += dz[p] * h[p + (ky-1, kx-1)], register sliding window along 8-pixel row segments
    if (wg_worker) {
      for (int item = part; item < ITEMS; item += PARTS) {
        const int r = item / (TX / 8), xs = (item % (TX / 8)) * 8;
        const uint2* xrow = hT + ((r + ky + 1) * HP + xs + 1) * NQ + cq;      // h[r + ky - 1][xs + i - 1] at index i
        const uint2* zrow = zT + ((r + 1) * ZP + xs + 1) * NQ + cq;
        float2 win[3][2];
        unpack4(xrow[0], win[1][0], win[1][1]);
        unpack4(xrow[NQ], win[2][0], win[2][1]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          win[0][0] = win[1][0]; win[0][1] = win[1][1];
          win[1][0] = win[2][0]; win[1][1] = win[2][1];
          unpack4(xrow[(t + 2) * NQ], win[2][0], win[2][1]);
          float2 glo, ghi;
          unpack4(zrow[t * NQ], glo, ghi);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            ffma2(gw[k][0], glo, win[k][0]);
            ffma2(gw[k][1], ghi, win[k][1]);
          }
          if (ky == 1) { gb[0] = add2(gb[0], glo); gb[1] = add2(gb[1], ghi); }
        }
      }
    }
    fence_proxy_async();                               // generic-proxy writes of dz are ordered before the next TMA refill
    __syncthreads();
  }
  // ---- combine the per-thread partial sums: shared-memory atomics, then one global atomic per (channel, tap) per CTA
  {
    const float gcv[4] = {gc[0].x, gc[0].y, gc[1].x, gc[1].y};
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&red_c[cq * 4 + j], gcv[j]);
    if (wg_worker) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float v[4] = {gw[k][0].x, gw[k][0].y, gw[k][1].x, gw[k][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&red_w[(cq * 4 + j) * 9 + ky * 3 + k], v[j]);
      }
      if (ky == 1) {
        const float v[4] = {gb[0].x, gb[0].y, gb[1].x, gb[1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&red_b[cq * 4 + j], v[j]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 64 * 9; i += NT) {
    const int c = c_base + i / 9;
    if (c < C) atomicAdd(dweight + (long)c * 9 + i % 9, red_w[i]);
  }
  if (tid < 64 && c_base + tid < C) {
    if (dbias) atomicAdd(dbias + c_base + tid, red_b[tid]);
    if (dh_colsum) atomicAdd(dh_colsum + c_base + tid, red_c[tid]);
  }
}

// ------------------------------------------------------------------------------------------------ backward, GELU'(z) kept by forward
// Same outputs as mlp_dw_bwd_kernel, but dz = du * gp is a two-load product (no convolution / GELU' recompute, no second halo
// ring of h).  Two (TY+2) x (TX+2) buffers: [du -> dz in place] and [gp, then h]: the h tile is fetched by a second TMA into
// the buffer gp has just vacated while phase B (the transposed convolution, which only reads dz) runs.
template <int TX, int TY>
__global__ void __launch_bounds__(NT, 2) mlp_dw_bwd_saved_kernel(const __grid_constant__ CUtensorMap tmDU, const __grid_constant__ CUtensorMap tmGP,
                                                                const __grid_constant__ CUtensorMap tmH, const float* __restrict__ weight,
                                                                bf16* __restrict__ dh, float* __restrict__ dweight, float* __restrict__ dbias,
                                                                float* __restrict__ dh_colsum, int B, int H, int W, int C, int tiles_x, int tiles_y) {
  pdl_sync();
  constexpr int ZP = TX + 2, ZR = TY + 2, BUF_BYTES = ZR * ZP * PIX_BYTES;
  constexpr int BR = TY / 2, BC = TX / 4;
  constexpr int PARTS = NPG / 3;
  constexpr int ITEMS = TY * (TX / 8);
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  uint8_t* dsm = align128(dsm_raw);
  uint2* zT = reinterpret_cast<uint2*>(dsm);                         // du, then dz
  uint2* hT = reinterpret_cast<uint2*>(dsm + BUF_BYTES);             // gp, then h      (both with origin (ty0-1, tx0-1))
  float* red_w = reinterpret_cast<float*>(dsm + 2 * BUF_BYTES);      // [64][9]
  float* red_b = red_w + 64 * 9;
  float* red_c = red_b + 64;
  const int tid = threadIdx.x, cq = tid & (NQ - 1), pg = tid >> 4;
  const int c_base = blockIdx.y * 64;
  const int c0 = c_base + cq * 4;
  const int n_tiles = B * tiles_x * tiles_y;
  const long row_stride = (long)W * C;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 64 * 9 + 128; i += NT) red_w[i] = 0.f;
  float2 w[9][2], bq[2];
  load_taps(weight, nullptr, c0, C, w, bq);
  const int ky = pg / PARTS, part = pg % PARTS;
  const bool wg_worker = pg < 3 * PARTS;
  float2 gw[3][2], gb[2], gc[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) gw[k][0] = gw[k][1] = make_float2(0.f, 0.f);
  gb[0] = gb[1] = gc[0] = gc[1] = make_float2(0.f, 0.f);
  __syncthreads();
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int tx0 = (tile % tiles_x) * TX, ty0 = ((tile / tiles_x) % tiles_y) * TY, b = tile / (tiles_x * tiles_y);
    if (tid == 0) {
      mbar_expect_tx(&bar[0], 2 * BUF_BYTES);
      tma_load_4d(&tmDU, &bar[0], zT, c_base, tx0 - 1, ty0 - 1, b);
      tma_load_4d(&tmGP, &bar[0], hT, c_base, tx0 - 1, ty0 - 1, b);
    }
    mbar_wait(&bar[0], it & 1);
    // ---- phase A: dz = du * GELU'(z) on the tile plus its one-pixel ring (zero outside the image: du is zero-filled there)
    for (int i = tid; i < ZR * ZP * NQ; i += NT) {
      float2 dlo, dhi, glo, ghi;
      unpack4(zT[i], dlo, dhi);
      unpack4(hT[i], glo, ghi);
      zT[i] = pack4(mul2(dlo, glo), mul2(dhi, ghi));
    }
    fence_proxy_async();                               // gp has been consumed (generic proxy) before the TMA refill below
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar[1], BUF_BYTES);
      tma_load_4d(&tmH, &bar[1], hT, c_base, tx0 - 1, ty0 - 1, b);
    }
    // ---- phase B: dh = dz + conv^T(dz), running column sum of dh (fc1 bias gradient); overlaps the h fetch
    for (int blk = pg; blk < BR * BC; blk += NPG) {
      const int r0 = (blk / BC) * 2, x0 = (blk % BC) * 4;
      float2 acc[2][4][2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[r][t][0] = acc[r][t][1] = make_float2(0.f, 0.f);
      conv3x3_block<2, 4, true>(zT, ZP, r0, x0, cq, w, acc);
      const int oy0 = ty0 + r0, ox0 = tx0 + x0;
      bf16* dp = dh + (((long)b * H + oy0) * W + ox0) * C + c0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint2 o = pack4(acc[r][t][0], acc[r][t][1]);
          if (c0 < C && oy0 + r < H && ox0 + t < W) {
            *reinterpret_cast<uint2*>(dp + (long)r * row_stride + t * C) = o;
            float2 lo, hi;
            unpack4(o, lo, hi);
            gc[0] = add2(gc[0], lo);
            gc[1] = add2(gc[1], hi);
          }
        }
      }
    }
    // ---- phase C: dW[ky][kx] += dz[p] * h[p + (ky-1, kx-1)]
    mbar_wait(&bar[1], it & 1);
    if (wg_worker) {
      for (int item = part; item < ITEMS; item += PARTS) {
        const int r = item / (TX / 8), xs = (item % (TX / 8)) * 8;
        const uint2* xrow = hT + ((r + ky) * ZP + xs) * NQ + cq;           // h[r + ky - 1][xs + i - 1] at index i  (origin -1)
        const uint2* zrow = zT + ((r + 1) * ZP + xs + 1) * NQ + cq;
        float2 win[3][2];
        unpack4(xrow[0], win[1][0], win[1][1]);
        unpack4(xrow[NQ], win[2][0], win[2][1]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          win[0][0] = win[1][0]; win[0][1] = win[1][1];
          win[1][0] = win[2][0]; win[1][1] = win[2][1];
          unpack4(xrow[(t + 2) * NQ], win[2][0], win[2][1]);
          float2 glo, ghi;
          unpack4(zrow[t * NQ], glo, ghi);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            ffma2(gw[k][0], glo, win[k][0]);
            ffma2(gw[k][1], ghi, win[k][1]);
          }
          if (ky == 1) { gb[0] = add2(gb[0], glo); gb[1] = add2(gb[1], ghi); }
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
  }
  {
    const float gcv[4] = {gc[0].x, gc[0].y, gc[1].x, gc[1].y};
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&red_c[cq * 4 + j], gcv[j]);
    if (wg_worker) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float v[4] = {gw[k][0].x, gw[k][0].y, gw[k][1].x, gw[k][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&red_w[(cq * 4 + j) * 9 + ky * 3 + k], v[j]);
      }
      if (ky == 1) {
        const float v[4] = {gb[0].x, gb[0].y, gb[1].x, gb[1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&red_b[cq * 4 + j], v[j]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 64 * 9; i += NT) {
    const int c = c_base + i / 9;
    if (c < C) atomicAdd(dweight + (long)c * 9 + i % 9, red_w[i]);
  }
  if (tid < 64 && c_base + tid < C) {
    if (dbias) atomicAdd(dbias + c_base + tid, red_b[tid]);
    if (dh_colsum) atomicAdd(dh_colsum + c_base + tid, red_c[tid]);
  }
}

// ------------------------------------------------------------------------------------------------ host
// tile geometry: 8 x 32 pixels unless 6 x 40 wastes less of the image on partial tiles
bool wide_tile(int H, int W) {
  const long a = (long)dfb_cdiv(H, 8) * 8 * dfb_cdiv(W, 32) * 32, b = (long)dfb_cdiv(H, 6) * 6 * dfb_cdiv(W, 40) * 40;
  return b < a;
}

int grid_x(int n_tiles, int nslab) {
  int gx = (2 * 148) / nslab;                          // two resident CTAs per SM over all channel slabs, no second wave
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  return gx;
}

template <int TX, int TY>
int launch_fwd(const void* h, const float* weight, const float* bias, int B, int H, int W, int C, void* u, void* gp, cudaStream_t st) {
  constexpr int smem = 2 * (TY + 2) * (TX + 2) * PIX_BYTES + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mlp_dw_fwd_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("mlp_dw_fwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  CUtensorMap tm;
  int rc = make_map_nhwc(&tm, h, B, H, W, C, TX + 2, TY + 2);
  if (rc) return rc;
  const int tiles_x = dfb_cdiv(W, TX), tiles_y = dfb_cdiv(H, TY), nslab = dfb_cdiv(C, 64);
  dim3 grid(grid_x(B * tiles_x * tiles_y, nslab), nslab);
  dfb_launch(mlp_dw_fwd_kernel<TX, TY>, grid, NT, smem, st, tm, weight, bias, (bf16*)u, (bf16*)gp, B, H, W, C, tiles_x, tiles_y);
  return dfb_check_launch("mlp_dw_fwd");
}

template <int TX, int TY>
int launch_bwd(const void* du, const void* h, const float* weight, const float* bias, int B, int H, int W, int C, void* dh, float* dweight, float* dbias,
               float* dh_colsum, cudaStream_t st) {
  using G = BwdGeom<TX, TY>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mlp_dw_bwd_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) { dfb_set_error("mlp_dw_bwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  CUtensorMap tmH, tmDU;
  int rc = make_map_nhwc(&tmH, h, B, H, W, C, G::HP, G::HR);
  if (rc) return rc;
  rc = make_map_nhwc(&tmDU, du, B, H, W, C, G::ZP, G::ZR);
  if (rc) return rc;
  const int tiles_x = dfb_cdiv(W, TX), tiles_y = dfb_cdiv(H, TY), nslab = dfb_cdiv(C, 64);
  dim3 grid(grid_x(B * tiles_x * tiles_y, nslab), nslab);
  dfb_launch(mlp_dw_bwd_kernel<TX, TY>, grid, NT, G::SMEM, st, tmH, tmDU, weight, bias, (bf16*)dh, dweight, dbias, dh_colsum, B, H, W, C, tiles_x, tiles_y);
  return dfb_check_launch("mlp_dw_bwd");
}

template <int TX, int TY>
int launch_bwd_saved(const void* du, const void* gp, const void* h, const float* weight, int B, int H, int W, int C, void* dh, float* dweight, float* dbias,
                     float* dh_colsum, cudaStream_t st) {
  constexpr int smem = 2 * (TY + 2) * (TX + 2) * PIX_BYTES + (64 * 9 + 128) * 4 + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mlp_dw_bwd_saved_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { dfb_set_error("mlp_dw_bwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  CUtensorMap tmDU, tmGP, tmH;
  int rc = make_map_nhwc(&tmDU, du, B, H, W, C, TX + 2, TY + 2);
  if (rc) return rc;
  rc = make_map_nhwc(&tmGP, gp, B, H, W, C, TX + 2, TY + 2);
  if (rc) return rc;
  rc = make_map_nhwc(&tmH, h, B, H, W, C, TX + 2, TY + 2);
  if (rc) return rc;
  const int tiles_x = dfb_cdiv(W, TX), tiles_y = dfb_cdiv(H, TY), nslab = dfb_cdiv(C, 64);
  dim3 grid(grid_x(B * tiles_x * tiles_y, nslab), nslab);
  dfb_launch(mlp_dw_bwd_saved_kernel<TX, TY>, grid, NT, smem, st, tmDU, tmGP, tmH, weight, (bf16*)dh, dweight, dbias, dh_colsum, B, H, W, C, tiles_x, tiles_y);
  return dfb_check_launch("mlp_dw_bwd_saved");
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_mlp_dw_fwd(const void* h, int dtype, const float* weight, const float* bias, int B, int H, int W, int C, void* u, void* gp,
                                 void* stream) {
  DFB_REQUIRE(dtype == 1, "mlp_dw_fwd: bf16 activations only (use dfb200_dwconv_fwd for fp32)");
  DFB_REQUIRE(C % 8 == 0, "mlp_dw_fwd: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(B > 0 && H > 0 && W > 0, "mlp_dw_fwd: empty input");
  return wide_tile(H, W) ? launch_fwd<40, 6>(h, weight, bias, B, H, W, C, u, gp, ST) : launch_fwd<32, 8>(h, weight, bias, B, H, W, C, u, gp, ST);
}

extern "C" int dfb200_mlp_dw_bwd(const void* du, const void* gp, const void* h, int dtype, const float* weight, const float* bias, int B, int H, int W, int C,
                                 void* dh, float* dweight, float* dbias, float* dh_colsum, void* stream) {
  DFB_REQUIRE(dtype == 1, "mlp_dw_bwd: bf16 activations only (use dfb200_dwconv_bwd for fp32)");
  DFB_REQUIRE(C % 8 == 0, "mlp_dw_bwd: C %% 8 != 0 (C=%d)", C);
  DFB_REQUIRE(B > 0 && H > 0 && W > 0, "mlp_dw_bwd: empty input");
  DFB_REQUIRE(dweight != nullptr && dh != nullptr, "mlp_dw_bwd: dh and dweight are required");
  if (gp)
    return wide_tile(H, W) ? launch_bwd_saved<40, 6>(du, gp, h, weight, B, H, W, C, dh, dweight, dbias, dh_colsum, ST)
                           : launch_bwd_saved<32, 8>(du, gp, h, weight, B, H, W, C, dh, dweight, dbias, dh_colsum, ST);
  return wide_tile(H, W) ? launch_bwd<40, 6>(du, h, weight, bias, B, H, W, C, dh, dweight, dbias, dh_colsum, ST)
                         : launch_bwd<32, 8>(du, h, weight, bias, B, H, W, C, dh, dweight, dbias, dh_colsum, ST);
}
