// Global Awareness Attention core (DFormer.py:122-130), bf16, on the tensor cores.
//
//   S = scale * Q K^T  [49, HW]      P = softmax_HW(S)      O = P V  [49, d]             per (image, head)
//
// 49 query rows (padded to 64) x d in {16, 32, 36, 48} (padded to a multiple of 16) is far too small for a tcgen05 tile
// pipeline: a whole (image, head) problem is 0.5 MFLOP per pixel chunk, the head slices of `kv` start at 72-byte offsets
// (not TMA-addressable) and a TMEM allocation + mbarrier round trip per CTA would cost more than the math.  The kernels
// below therefore use warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate; SASS HMMA) in the flash-attention register
// layout: a warp owns 16 query rows and a whole pixel chunk, so row maxima / sums never leave the warp's quads and the
// probabilities go from accumulator registers straight into the A operand of the next product.
//
//   forward   CTA = 128 pixels of one (image, head), 4 warps x 16 query rows: S -> local max / sum -> P~ (bf16) -> P~ V,
//             written as an unnormalised partial next to (max, sum); the last CTA of the (image, head) (atomic ticket)
//             merges the partials flash-decoding style and stores O and the row log-sum-exp.
//   backward  CTA = 64 pixels: phase 1 (warp = 16 query rows) recomputes P from the saved log-sum-exp, dP = dO V^T,
//             dS = P o (dP - rowsum(dO o O)); dQ partial = dS K goes out through fp32 atomics; P and dS are parked in
//             shared memory (bf16); phase 2 (warp = 16 pixels) dV = P^T dO, dK = dS^T Q, complete per pixel.
// Nothing of size 49 x HW reaches HBM in either direction.
#include <string.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int NQ = 49;             // pooled query tokens (7 x 7)
constexpr int QR = 64;             // padded to four m16 tiles
constexpr int NT = 128;            // 4 warps

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int D>
struct Geo {
  static constexpr int DP = (D + 15) / 16 * 16;     // reduction / output width in the MMAs (zero padded)
  static constexpr int PITCH = DP + 8;              // smem row pitch in elements: 8 consecutive rows hit 32 distinct banks for ldmatrix
  static constexpr int W8 = D / 4, WP8 = DP / 4;    // 8-byte words per real / padded row
};

// rows of bf16 [.., D] slices (row pitch `pitch` elements, 8-byte aligned) -> smem [rows][PITCH], columns D..DP-1 and rows >= nvalid zeroed
template <int D, int ROWS>
__device__ __forceinline__ void stage_bf16(const bf16* __restrict__ src, long pitch, int nvalid, bf16* __restrict__ dst) {
  constexpr int WP8 = Geo<D>::WP8, W8 = Geo<D>::W8, PITCH = Geo<D>::PITCH;
#pragma unroll
  for (int i = threadIdx.x; i < ROWS * WP8; i += NT) {
    const int r = i / WP8, w = i % WP8;
    uint2 v = make_uint2(0u, 0u);
    if (r < nvalid && w < W8) v = *reinterpret_cast<const uint2*>(src + (long)r * pitch + 4 * w);
    *reinterpret_cast<uint2*>(dst + r * PITCH + 4 * w) = v;
  }
}
// same from an fp32 source (the gradient of the attention output), rounded to bf16
template <int D, int ROWS>
__device__ __forceinline__ void stage_f32(const float* __restrict__ src, long pitch, int nvalid, bf16* __restrict__ dst) {
  constexpr int WP8 = Geo<D>::WP8, W8 = Geo<D>::W8, PITCH = Geo<D>::PITCH;
#pragma unroll
  for (int i = threadIdx.x; i < ROWS * WP8; i += NT) {
    const int r = i / WP8, w = i % WP8;
    uint2 v = make_uint2(0u, 0u);
    if (r < nvalid && w < W8) {
      const float4 f = *reinterpret_cast<const float4*>(src + (long)r * pitch + 4 * w);
      v = make_uint2(pack2(f.x, f.y), pack2(f.z, f.w));
    }
    *reinterpret_cast<uint2*>(dst + r * PITCH + 4 * w) = v;
  }
}

// acc[nt][0..3] (16 rows x NTILES*8 columns) = A[16 x DP] (rows row0.. of As) * B^T, B = Bs[NTILES*8 rows][DP] (both K-contiguous)
template <int D, int NTILES>
__device__ __forceinline__ void mma_rows_by_rows(const bf16* As, int row0, const bf16* Bs, float (*acc)[4], int lane) {
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH;
#pragma unroll
  for (int kk = 0; kk < DP / 16; ++kk) {
    uint32_t a[4];
    ldsm_x4(smem_addr(As + (row0 + (lane & 15)) * PITCH + kk * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int np = 0; np < NTILES / 2; ++np) {
      uint32_t b0, b1, b2, b3;       // (n-tile 2np: k 0-7, k 8-15), (n-tile 2np+1: k 0-7, k 8-15)
      const int i = lane >> 3;
      ldsm_x4(smem_addr(Bs + (np * 16 + (lane & 7) + (i >> 1) * 8) * PITCH + kk * 16 + (i & 1) * 8), b0, b1, b2, b3);
      mma16816(acc[2 * np], a, b0, b1);
      mma16816(acc[2 * np + 1], a, b2, b3);
    }
  }
}

// out[nt][0..3] (16 rows x DP columns) += A (accumulator-layout registers `pa`, 16 rows x KT*16 reduction) * Bs[KT*16 rows][DP] (N-contiguous)
template <int D, int KT>
__device__ __forceinline__ void mma_regs_by_cols(const uint32_t (*pa)[4], const bf16* Bs, float (*out)[4], int lane) {
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH;
#pragma unroll
  for (int j = 0; j < KT; ++j) {
#pragma unroll
    for (int np = 0; np < DP / 16; ++np) {
      uint32_t b0, b1, b2, b3;       // (n-tile 2np: k 0-7, k 8-15), (n-tile 2np+1: k 0-7, k 8-15), transposed on the way in
      const int i = lane >> 3;
      ldsm_x4_t(smem_addr(Bs + (j * 16 + (lane & 7) + (i & 1) * 8) * PITCH + np * 16 + (i >> 1) * 8), b0, b1, b2, b3);
      mma16816(out[2 * np], pa[j], b0, b1);
      mma16816(out[2 * np + 1], pa[j], b2, b3);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(NT) gaa_mma_fwd_kernel(const bf16* __restrict__ m, const bf16* __restrict__ kv, int HW, int heads, float scale, int nchunks,
                                                        float* __restrict__ out, float* __restrict__ lse, float* __restrict__ part, int* __restrict__ counters) {
  pdl_sync();
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH, PC = 128, PS = D + 4;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);          // [64][PITCH]
  bf16* Ks = Qs + QR * PITCH;                          // [PC][PITCH]
  bf16* Vs = Ks + PC * PITCH;                          // [PC][PITCH]
  float* ms = reinterpret_cast<float*>(Vs + PC * PITCH);   // [64] merge: 1 / total sum
  float* pm = ms + QR;                                 // [nchunks][49] merge staging (chunk maxima -> weights)
  float* pl = pm + nchunks * NQ;                       // [nchunks][49] chunk sums
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, head = bh % heads, c = blockIdx.x;
  const int Cp = heads * D;
  const int nvalid = min(PC, HW - c * PC);
  stage_bf16<D, QR>(m + (long)b * NQ * Cp + head * D, Cp, NQ, Qs);
  {
    const bf16* rows = kv + ((long)b * HW + (long)c * PC) * 2 * Cp + head * D;
    stage_bf16<D, PC>(rows, 2L * Cp, nvalid, Ks);
    stage_bf16<D, PC>(rows + Cp, 2L * Cp, nvalid, Vs);
  }
  __syncthreads();
  // ---- S = Q K^T for this warp's 16 query rows x 128 pixels
  float s[PC / 8][4];
#pragma unroll
  for (int i = 0; i < PC / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
  mma_rows_by_rows<D, PC / 8>(Qs, warp * 16, Ks, s, lane);
  // ---- local softmax statistics (rows g and g + 8 of the warp's slab; a row lives in the 4 lanes of a quad)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < PC / 8; ++i) {
    const int col = i * 8 + 2 * t;
    s[i][0] = col < nvalid ? s[i][0] * scale : -INFINITY;
    s[i][1] = col + 1 < nvalid ? s[i][1] * scale : -INFINITY;
    s[i][2] = col < nvalid ? s[i][2] * scale : -INFINITY;
    s[i][3] = col + 1 < nvalid ? s[i][3] * scale : -INFINITY;
    mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
    mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[PC / 16][4];           // P~ = exp(S - local max) as A fragments (pixel k-steps of 16)
#pragma unroll
  for (int j = 0; j < PC / 16; ++j) {
    float e[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      e[4 * h + 0] = __expf(s[2 * j + h][0] - mx0); e[4 * h + 1] = __expf(s[2 * j + h][1] - mx0);     // exp(-inf) = 0 for the padded tail
      e[4 * h + 2] = __expf(s[2 * j + h][2] - mx1); e[4 * h + 3] = __expf(s[2 * j + h][3] - mx1);
    }
    l0 += e[0] + e[1] + e[4] + e[5];
    l1 += e[2] + e[3] + e[6] + e[7];
    pa[j][0] = pack2(e[0], e[1]); pa[j][1] = pack2(e[2], e[3]); pa[j][2] = pack2(e[4], e[5]); pa[j][3] = pack2(e[6], e[7]);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- unnormalised partial context O~ = P~ V
  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  mma_regs_by_cols<D, PC / 16>(pa, Vs, o, lane);
  {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    float* d0 = part + (((long)bh * nchunks + c) * NQ + r0) * PS;
    float* d1 = part + (((long)bh * nchunks + c) * NQ + r1) * PS;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * t;
      if (col < D) {
        if (r0 < NQ) *reinterpret_cast<float2*>(d0 + col) = make_float2(o[i][0], o[i][1]);
        if (r1 < NQ) *reinterpret_cast<float2*>(d1 + col) = make_float2(o[i][2], o[i][3]);
      }
    }
    if (t == 0) {
      if (r0 < NQ) *reinterpret_cast<float2*>(d0 + D) = make_float2(mx0, l0);
      if (r1 < NQ) *reinterpret_cast<float2*>(d1 + D) = make_float2(mx1, l1);
    }
  }
  // ---- the last CTA of this (image, head) merges all partials
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[bh], 1) == nchunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* base = part + (long)bh * nchunks * NQ * PS;
  for (int i = tid; i < nchunks * NQ; i += NT) {
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(base + (long)i * PS + D));
    pm[i] = ml.x;
    pl[i] = ml.y;
  }
  __syncthreads();
  for (int r = warp; r < NQ; r += NT / 32) {
    float M = -INFINITY;
    for (int cc = lane; cc < nchunks; cc += 32) M = fmaxf(M, pm[cc * NQ + r]);
    M = warp_max(M);
    float L = 0.f;
    for (int cc = lane; cc < nchunks; cc += 32) {
      const float w = __expf(pm[cc * NQ + r] - M);
      L = fmaf(pl[cc * NQ + r], w, L);
      pm[cc * NQ + r] = w;
    }
    L = warp_sum(L);
    if (lane == 0) { ms[r] = 1.0f / L; lse[(long)bh * NQ + r] = M + __logf(L); }
  }
  __syncthreads();
  constexpr int D4 = D / 4;
  for (int idx = tid; idx < NQ * D4; idx += NT) {
    const int r = idx / D4, j = (idx % D4) * 4;
    const float* src = base + (long)r * PS + j;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int cc = 0; cc < nchunks; ++cc) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long)cc * NQ * PS));
      const float w = pm[cc * NQ + r];
      acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
    }
    const float inv = ms[r];
    *reinterpret_cast<float4*>(out + ((long)b * NQ + r) * Cp + head * D + j) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
  if (tid == 0) counters[bh] = 0;                        // self-resetting ticket: the buffer is reusable by the next launch
}

// ------------------------------------------------------------------------------------------------ backward
template <int D>
__global__ void __launch_bounds__(NT) gaa_mma_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ lse,
                                                        const bf16* __restrict__ m, const bf16* __restrict__ kv, int HW, int heads, float scale,
                                                        float* __restrict__ dm, bf16* __restrict__ dkv) {
  pdl_sync();
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH, PC = 64, PP = PC + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);          // [64][PITCH]
  bf16* dOs = Qs + QR * PITCH;                         // [64][PITCH]
  bf16* Ks = dOs + QR * PITCH;                         // [PC][PITCH]
  bf16* Vs = Ks + PC * PITCH;                          // [PC][PITCH]
  bf16* Ps = Vs + PC * PITCH;                          // [64][PP]  probabilities, query-major
  bf16* dSs = Ps + QR * PP;                            // [64][PP]
  float* lses = reinterpret_cast<float*>(dSs + QR * PP);   // [64]
  float* Dr = lses + QR;                               // [64] rowsum(dO o O) = rowsum(dP o P)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, head = bh % heads, c = blockIdx.x;
  const int Cp = heads * D;
  const int nvalid = min(PC, HW - c * PC);
  const long qoff = (long)b * NQ * Cp + head * D;
  stage_bf16<D, QR>(m + qoff, Cp, NQ, Qs);
  stage_f32<D, QR>(dout + qoff, Cp, NQ, dOs);
  {
    const bf16* rows = kv + ((long)b * HW + (long)c * PC) * 2 * Cp + head * D;
    stage_bf16<D, PC>(rows, 2L * Cp, nvalid, Ks);
    stage_bf16<D, PC>(rows + Cp, 2L * Cp, nvalid, Vs);
  }
  if (tid < QR) lses[tid] = tid < NQ ? lse[(long)bh * NQ + tid] : 0.f;
  for (int r = warp; r < QR; r += NT / 32) {             // Dr from the fp32 sources (one warp per row)
    float sdo = 0.f;
    if (r < NQ)
      for (int j = lane; j < D; j += 32) sdo = fmaf(dout[qoff + (long)r * Cp + j], out[qoff + (long)r * Cp + j], sdo);
    sdo = warp_sum(sdo);
    if (lane == 0) Dr[r] = sdo;
  }
  __syncthreads();
  // ---- phase 1: this warp's 16 query rows x 64 pixels
  {
    float s[PC / 8][4], dp[PC / 8][4];
#pragma unroll
    for (int i = 0; i < PC / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_rows_by_rows<D, PC / 8>(Qs, warp * 16, Ks, s, lane);
    mma_rows_by_rows<D, PC / 8>(dOs, warp * 16, Vs, dp, lane);
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const float ls0 = lses[r0], ls1 = lses[r1], dr0 = Dr[r0], dr1 = Dr[r1];
    const bool ok0 = r0 < NQ, ok1 = r1 < NQ;
    uint32_t da[PC / 16][4];         // dS as A fragments for dQ = dS K
#pragma unroll
    for (int i = 0; i < PC / 8; ++i) {
      const int col = i * 8 + 2 * t;
      const bool c0 = col < nvalid, c1 = col + 1 < nvalid;
      const float p00 = (ok0 && c0) ? __expf(fmaf(s[i][0], scale, -ls0)) : 0.f, p01 = (ok0 && c1) ? __expf(fmaf(s[i][1], scale, -ls0)) : 0.f;
      const float p10 = (ok1 && c0) ? __expf(fmaf(s[i][2], scale, -ls1)) : 0.f, p11 = (ok1 && c1) ? __expf(fmaf(s[i][3], scale, -ls1)) : 0.f;
      const float d00 = p00 * (dp[i][0] - dr0), d01 = p01 * (dp[i][1] - dr0), d10 = p10 * (dp[i][2] - dr1), d11 = p11 * (dp[i][3] - dr1);
      const uint32_t pk0 = pack2(p00, p01), pk1 = pack2(p10, p11), dk0 = pack2(d00, d01), dk1 = pack2(d10, d11);
      *reinterpret_cast<uint32_t*>(Ps + r0 * PP + col) = pk0;
      *reinterpret_cast<uint32_t*>(Ps + r1 * PP + col) = pk1;
      *reinterpret_cast<uint32_t*>(dSs + r0 * PP + col) = dk0;
      *reinterpret_cast<uint32_t*>(dSs + r1 * PP + col) = dk1;
      da[i >> 1][(i & 1) * 2 + 0] = dk0;
      da[i >> 1][(i & 1) * 2 + 1] = dk1;
    }
    float dq[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    mma_regs_by_cols<D, PC / 16>(da, Ks, dq, lane);
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * t;
      if (col < D) {
        if (ok0) { atomicAdd(dm + qoff + (long)r0 * Cp + col, dq[i][0] * scale); atomicAdd(dm + qoff + (long)r0 * Cp + col + 1, dq[i][1] * scale); }
        if (ok1) { atomicAdd(dm + qoff + (long)r1 * Cp + col, dq[i][2] * scale); atomicAdd(dm + qoff + (long)r1 * Cp + col + 1, dq[i][3] * scale); }
      }
    }
  }
  __syncthreads();
  // ---- phase 2: this warp's 16 pixels: dV = P^T dO, dK = scale * dS^T Q   (reduction over the 64 query rows)
  {
    float dv[DP / 8][4], dk[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) { dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; }
    const int m0 = warp * 16;
    const int i4 = lane >> 3;
#pragma unroll
    for (int kq = 0; kq < QR / 16; ++kq) {
      uint32_t ap[4], as[4];         // A[m = pixel][k = query] fragments, transposed out of the query-major tiles
      const int qrow = kq * 16 + (lane & 7) + (i4 >> 1) * 8, pcol = m0 + (i4 & 1) * 8;
      ldsm_x4_t(smem_addr(Ps + qrow * PP + pcol), ap[0], ap[1], ap[2], ap[3]);
      ldsm_x4_t(smem_addr(dSs + qrow * PP + pcol), as[0], as[1], as[2], as[3]);
#pragma unroll
      for (int np = 0; np < DP / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const int brow = kq * 16 + (lane & 7) + (i4 & 1) * 8, bcol = np * 16 + (i4 >> 1) * 8;
        ldsm_x4_t(smem_addr(dOs + brow * PITCH + bcol), b0, b1, b2, b3);
        mma16816(dv[2 * np], ap, b0, b1);
        mma16816(dv[2 * np + 1], ap, b2, b3);
        ldsm_x4_t(smem_addr(Qs + brow * PITCH + bcol), b0, b1, b2, b3);
        mma16816(dk[2 * np], as, b0, b1);
        mma16816(dk[2 * np + 1], as, b2, b3);
      }
    }
    const int p0 = m0 + g, p1 = p0 + 8;
    bf16* row0 = dkv + ((long)b * HW + (long)c * PC + p0) * 2 * Cp + head * D;
    bf16* row1 = dkv + ((long)b * HW + (long)c * PC + p1) * 2 * Cp + head * D;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * t;
      if (col < D) {
        if (p0 < nvalid) {
          *reinterpret_cast<uint32_t*>(row0 + col) = pack2(dk[i][0] * scale, dk[i][1] * scale);
          *reinterpret_cast<uint32_t*>(row0 + Cp + col) = pack2(dv[i][0], dv[i][1]);
        }
        if (p1 < nvalid) {
          *reinterpret_cast<uint32_t*>(row1 + col) = pack2(dk[i][2] * scale, dk[i][3] * scale);
          *reinterpret_cast<uint32_t*>(row1 + Cp + col) = pack2(dv[i][2], dv[i][3]);
        }
      }
    }
  }
}

template <int D> constexpr int fwd_smem_fixed() { return (QR + 2 * 128) * Geo<D>::PITCH * 2 + QR * 4; }
template <int D> constexpr int bwd_smem() { return (2 * QR + 2 * 64) * Geo<D>::PITCH * 2 + 2 * QR * (64 + 8) * 2 + 2 * QR * 4; }

template <int D>
int launch_fwd(const void* m, const void* kv, int B, int HW, int heads, float* out, float* lse, float* part, int* counters, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_mma_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { dfb_set_error("gaa_mma_fwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int nchunks = dfb_cdiv(HW, 128);
  const int smem = fwd_smem_fixed<D>() + 2 * nchunks * NQ * 4;            // + merge staging of the (max, sum) pairs
  if (smem > 200 * 1024) { dfb_set_error("gaa_mma_fwd: HW=%d too large for the one-launch merge", HW); return DFB_ERR_UNSUPPORTED; }
  dim3 grid(nchunks, B * heads);
  dfb_launch(gaa_mma_fwd_kernel<D>, grid, NT, smem, st, (const bf16*)m, (const bf16*)kv, HW, heads, 1.0f / sqrtf((float)D), nchunks, out, lse, part, counters);
  return dfb_check_launch("gaa_mma_fwd");
}

template <int D>
int launch_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, float* dm, void* dkv,
               cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_mma_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<D>());
    if (e != cudaSuccess) { dfb_set_error("gaa_mma_bwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  cudaMemsetAsync(dm, 0, sizeof(float) * (size_t)B * NQ * heads * D, st);
  dim3 grid(dfb_cdiv(HW, 64), B * heads);
  dfb_launch(gaa_mma_bwd_kernel<D>, grid, NT, bwd_smem<D>(), st, dout, out, lse, (const bf16*)m, (const bf16*)kv, HW, heads, 1.0f / sqrtf((float)D), dm,
             (bf16*)dkv);
  return dfb_check_launch("gaa_mma_bwd");
}

}  // namespace

#define GAA_MMA_DISPATCH_D(d, ...)                                    \
  switch (d) {                                                        \
    case 16: { constexpr int D = 16; __VA_ARGS__ } break;             \
    case 32: { constexpr int D = 32; __VA_ARGS__ } break;             \
    case 36: { constexpr int D = 36; __VA_ARGS__ } break;             \
    case 48: { constexpr int D = 48; __VA_ARGS__ } break;             \
    default: dfb_set_error("gaa_mma: head dim %d not instantiated (16, 32, 36, 48)", d); return DFB_ERR_UNSUPPORTED; \
  }

int dfb_gaa_mma_fwd(const void* m, const void* kv, int B, int HW, int heads, int d, float* out, float* lse, float* scratch, int* counters, cudaStream_t st) {
  GAA_MMA_DISPATCH_D(d, { return launch_fwd<D>(m, kv, B, HW, heads, out, lse, scratch, counters, st); });
  return DFB_OK;
}

int dfb_gaa_mma_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, int d, float* dm,
                    void* dkv, cudaStream_t st) {
  GAA_MMA_DISPATCH_D(d, { return launch_bwd<D>(dout, out, lse, m, kv, B, HW, heads, dm, dkv, st); });
  return DFB_OK;
}
