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
//   forward   the CTAs of a thread-block cluster share one (image, head); each walks 128-pixel chunks, 4 warps x 16 query rows:
//             S -> online max / sum -> P~ (bf16) -> O~ += P~ V; the partials are merged flash-decoding style through distributed
//             shared memory after one cluster barrier (no HBM scratch, no atomics).
//   backward  64-pixel chunks: phase 1 (warp = 16 query rows) recomputes P from the saved log-sum-exp, dP = dO V^T,
//             dS = P o (dP - rowsum(dO o O)), dQ += dS K in registers; P and dS are parked in shared memory (bf16); phase 2
//             (warp = 16 pixels) dV = P^T dO, dK = dS^T Q, complete per pixel.  dQ is reduced across the cluster through DSMEM.
// Nothing of size 49 x HW reaches HBM in either direction.
#include <string.h>
#include <cooperative_groups.h>

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

// 8-byte cp.async with zero fill (src_bytes = 0 -> the destination is zeroed, nothing is read)
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous copy of `ROWS` pixel rows of one head's K and V slices into a shared-memory buffer pair (real columns only: the
// zero padding of columns D..DP-1 is written once at kernel start and never overwritten); rows >= nvalid are zero-filled
template <int D, int ROWS>
__device__ __forceinline__ void prefetch_kv(const bf16* __restrict__ rows, long pitch, int Cp, int nvalid, bf16* __restrict__ Ks, bf16* __restrict__ Vs) {
  constexpr int W8 = Geo<D>::W8, PITCH = Geo<D>::PITCH;
#pragma unroll
  for (int i = threadIdx.x; i < ROWS * W8; i += NT) {
    const int r = i / W8, w = i % W8;
    const bool ok = r < nvalid;
    const bf16* src = rows + (ok ? (long)r * pitch : 0) + 4 * w;
    cp_async8(Ks + r * PITCH + 4 * w, src, ok ? 8 : 0);
    cp_async8(Vs + r * PITCH + 4 * w, src + Cp, ok ? 8 : 0);
  }
}
template <int D, int ROWS>
__device__ __forceinline__ void zero_pad_cols(bf16* __restrict__ buf) {
  constexpr int W8 = Geo<D>::W8, WP8 = Geo<D>::WP8, PITCH = Geo<D>::PITCH;
  if constexpr (WP8 > W8) {
    for (int i = threadIdx.x; i < ROWS * (WP8 - W8); i += NT) {
      const int r = i / (WP8 - W8), w = W8 + i % (WP8 - W8);
      *reinterpret_cast<uint2*>(buf + r * PITCH + 4 * w) = make_uint2(0u, 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// grid (nsplit, B * heads), thread-block cluster (nsplit, 1, 1): the CTAs of a cluster share one (image, head); CTA `split` walks the
// 128-pixel chunks split, split + nsplit, ... with an online softmax (K / V double-buffered through cp.async), parks its partial
// (O~, max, sum) in its own shared memory, and after one cluster barrier every CTA merges a slice of the 49 rows by reading its
// peers' partials through distributed shared memory.  No scratch buffer in HBM, no atomics, no tickets.
template <int D>
__global__ void __launch_bounds__(NT) gaa_mma_fwd_kernel(const bf16* __restrict__ m, const bf16* __restrict__ kv, int HW, int heads, float scale, int nchunks,
                                                        float* __restrict__ out, float* __restrict__ lse) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH, PC = 128;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);          // [64][PITCH]
  bf16* KV = Qs + QR * PITCH;                          // 2 x (K [PC][PITCH], V [PC][PITCH])
  float* Op = reinterpret_cast<float*>(KV + 4 * PC * PITCH);   // [64][DP] partial context
  float* mp = Op + QR * DP;                            // [64] running max
  float* lp = mp + QR;                                 // [64] running sum
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, head = bh % heads;
  const int nsplit = gridDim.x, split = blockIdx.x;
  const int Cp = heads * D;
  const int n_my = (nchunks - split + nsplit - 1) / nsplit;
  zero_pad_cols<D, 4 * PC>(KV);                        // touches no global memory: overlaps the previous kernel's tail under PDL
  pdl_sync();
  const bf16* kvb = kv + (long)b * HW * 2 * Cp + head * D;
  if (n_my > 0) prefetch_kv<D, PC>(kvb + (long)split * PC * 2 * Cp, 2L * Cp, Cp, min(PC, HW - split * PC), KV, KV + PC * PITCH);
  cp_async_commit();
  stage_bf16<D, QR>(m + (long)b * NQ * Cp + head * D, Cp, NQ, Qs);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  for (int it = 0; it < n_my; ++it) {
    const int c = split + it * nsplit;
    const int nvalid = min(PC, HW - c * PC);
    if (it + 1 < n_my) {
      const int cn = c + nsplit;
      bf16* nb = KV + ((it + 1) & 1) * 2 * PC * PITCH;
      prefetch_kv<D, PC>(kvb + (long)cn * PC * 2 * Cp, 2L * Cp, Cp, min(PC, HW - cn * PC), nb, nb + PC * PITCH);
    }
    cp_async_commit();
    cp_async_wait<1>();                               // everything but the newest group (the next chunk) has landed
    __syncthreads();
    const bf16* Ks = KV + (it & 1) * 2 * PC * PITCH;
    const bf16* Vs = Ks + PC * PITCH;
    // ---- S = Q K^T for this warp's 16 query rows x 128 pixels
    float s[PC / 8][4];
#pragma unroll
    for (int i = 0; i < PC / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_rows_by_rows<D, PC / 8>(Qs, warp * 16, Ks, s, lane);
    // ---- online softmax (rows g and g + 8 of the warp's slab; a row lives in the 4 lanes of a quad)
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
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);       // finite: pixel 0 of every chunk is real
    const float al0 = __expf(m0 - mn0), al1 = __expf(m1 - mn1);   // exp(-inf) = 0 on the first chunk
    m0 = mn0; m1 = mn1;
    float cs0 = 0.f, cs1 = 0.f;
    uint32_t pa[PC / 16][4];           // P~ = exp(S - running max) as A fragments (pixel k-steps of 16)
#pragma unroll
    for (int j = 0; j < PC / 16; ++j) {
      float e[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        e[4 * h + 0] = __expf(s[2 * j + h][0] - mn0); e[4 * h + 1] = __expf(s[2 * j + h][1] - mn0);
        e[4 * h + 2] = __expf(s[2 * j + h][2] - mn1); e[4 * h + 3] = __expf(s[2 * j + h][3] - mn1);
      }
      cs0 += e[0] + e[1] + e[4] + e[5];
      cs1 += e[2] + e[3] + e[6] + e[7];
      pa[j][0] = pack2(e[0], e[1]); pa[j][1] = pack2(e[2], e[3]); pa[j][2] = pack2(e[4], e[5]); pa[j][3] = pack2(e[6], e[7]);
    }
    l0 = fmaf(l0, al0, cs0);           // per-lane partial sums (the quad is reduced once, after the last chunk)
    l1 = fmaf(l1, al1, cs1);
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
    mma_regs_by_cols<D, PC / 16>(pa, Vs, o, lane);
    __syncthreads();                                   // this buffer is the prefetch target of the next iteration
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int col = i * 8 + 2 * t;
      *reinterpret_cast<float2*>(Op + r0 * DP + col) = make_float2(o[i][0], o[i][1]);
      *reinterpret_cast<float2*>(Op + r1 * DP + col) = make_float2(o[i][2], o[i][3]);
    }
    if (t == 0) { mp[r0] = m0; lp[r0] = l0; mp[r1] = m1; lp[r1] = l1; }
  }
  cluster.sync();
  // ---- merge: this CTA finishes rows split, split + nsplit, ... (thread = one output element)
  const int rows_mine = (NQ - split + nsplit - 1) / nsplit;
  for (int idx = tid; idx < rows_mine * D; idx += NT) {
    const int r = split + (idx / D) * nsplit, j = idx % D;
    float M = -INFINITY;
    for (int pr = 0; pr < nsplit; ++pr) M = fmaxf(M, cluster.map_shared_rank(mp, pr)[r]);
    float L = 0.f, acc = 0.f;
    for (int pr = 0; pr < nsplit; ++pr) {
      const float w = __expf(cluster.map_shared_rank(mp, pr)[r] - M);     // a CTA that saw no chunk carries max = -inf -> weight 0
      L = fmaf(cluster.map_shared_rank(lp, pr)[r], w, L);
      acc = fmaf(cluster.map_shared_rank(Op, pr)[r * DP + j], w, acc);
    }
    out[((long)b * NQ + r) * Cp + head * D + j] = acc / L;
    if (j == 0) lse[(long)bh * NQ + r] = M + __logf(L);
  }
  cluster.sync();                                      // peers may still be reading this CTA's partial
}

// ------------------------------------------------------------------------------------------------ backward
// Same cluster layout; a CTA walks 64-pixel chunks.  dQ partials stay in registers across the chunks and are reduced across the
// cluster through distributed shared memory (plain stores to dm: no memset, no atomics).
template <int D>
__global__ void __launch_bounds__(NT) gaa_mma_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ lse,
                                                        const bf16* __restrict__ m, const bf16* __restrict__ kv, int HW, int heads, float scale, int nchunks,
                                                        float* __restrict__ dm, bf16* __restrict__ dkv, float* __restrict__ dkv_colsum,
                                                        float* __restrict__ dm_colsum, bf16* __restrict__ dm_lo) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int DP = Geo<D>::DP, PITCH = Geo<D>::PITCH, PC = 64, PP = PC + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);          // [64][PITCH]
  bf16* dOs = Qs + QR * PITCH;                         // [64][PITCH]
  bf16* KV = dOs + QR * PITCH;                         // 2 x (K [PC][PITCH], V [PC][PITCH])
  bf16* Ps = KV + 4 * PC * PITCH;                      // [64][PP]  probabilities, query-major
  bf16* dSs = Ps + QR * PP;                            // [64][PP]
  float* lses = reinterpret_cast<float*>(dSs + QR * PP);   // [64]
  float* Dr = lses + QR;                               // [64] rowsum(dO o O) = rowsum(dP o P)
  float* dqp = Dr + QR;                                // [64][DP] this CTA's dQ partial (read by the cluster)
  float* csw = dqp + QR * DP;                          // [4 warps][2][DP] column sums of dK / dV, then [DP] column sums of dQ
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, head = bh % heads;
  const int nsplit = gridDim.x, split = blockIdx.x;
  const int Cp = heads * D;
  const int n_my = (nchunks - split + nsplit - 1) / nsplit;
  const long qoff = (long)b * NQ * Cp + head * D;
  zero_pad_cols<D, 4 * PC>(KV);
  pdl_sync();
  const bf16* kvb = kv + (long)b * HW * 2 * Cp + head * D;
  if (n_my > 0) prefetch_kv<D, PC>(kvb + (long)split * PC * 2 * Cp, 2L * Cp, Cp, min(PC, HW - split * PC), KV, KV + PC * PITCH);
  cp_async_commit();
  stage_bf16<D, QR>(m + qoff, Cp, NQ, Qs);
  stage_f32<D, QR>(dout + qoff, Cp, NQ, dOs);
  if (tid < QR) lses[tid] = tid < NQ ? lse[(long)bh * NQ + tid] : 0.f;
  {                                                      // Dr from the fp32 sources: a thread pair per row, all loads independent
    const int r = tid >> 1, hf = tid & 1;
    float sdo = 0.f;
    if (r < NQ) {
      const float4* a4 = reinterpret_cast<const float4*>(dout + qoff + (long)r * Cp + hf * (D / 2));
      const float4* b4 = reinterpret_cast<const float4*>(out + qoff + (long)r * Cp + hf * (D / 2));
      if constexpr ((D / 2) % 4 == 0) {
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          const float4 a = a4[j], c4 = b4[j];
          sdo = fmaf(a.x, c4.x, fmaf(a.y, c4.y, fmaf(a.z, c4.z, fmaf(a.w, c4.w, sdo))));
        }
      } else {                                           // D = 36: half rows of 18 floats (8-byte aligned)
        const float2* a2 = reinterpret_cast<const float2*>(a4);
        const float2* b2 = reinterpret_cast<const float2*>(b4);
#pragma unroll
        for (int j = 0; j < D / 4; ++j) {
          const float2 a = a2[j], c2 = b2[j];
          sdo = fmaf(a.x, c2.x, fmaf(a.y, c2.y, sdo));
        }
      }
    }
    sdo += __shfl_xor_sync(0xffffffffu, sdo, 1);
    if (hf == 0) Dr[r] = sdo;
  }
  float dq[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const bool ok0 = r0 < NQ, ok1 = r1 < NQ;
  float csk[DP / 8][2], csv[DP / 8][2];               // per-thread column sums of dK / dV over this CTA's pixels (bias gradient of the kv projection)
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) csk[i][0] = csk[i][1] = csv[i][0] = csv[i][1] = 0.f;
  for (int it = 0; it < n_my; ++it) {
    const int c = split + it * nsplit;
    const int nvalid = min(PC, HW - c * PC);
    if (it + 1 < n_my) {
      const int cn = c + nsplit;
      bf16* nb = KV + ((it + 1) & 1) * 2 * PC * PITCH;
      prefetch_kv<D, PC>(kvb + (long)cn * PC * 2 * Cp, 2L * Cp, Cp, min(PC, HW - cn * PC), nb, nb + PC * PITCH);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();                                   // K / V of this chunk (and, first time round, Q / dO / lse / Dr) are visible
    const bf16* Ks = KV + (it & 1) * 2 * PC * PITCH;
    const bf16* Vs = Ks + PC * PITCH;
    // ---- phase 1: this warp's 16 query rows x 64 pixels
    {
      float s[PC / 8][4], dp[PC / 8][4];
#pragma unroll
      for (int i = 0; i < PC / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
      mma_rows_by_rows<D, PC / 8>(Qs, warp * 16, Ks, s, lane);
      mma_rows_by_rows<D, PC / 8>(dOs, warp * 16, Vs, dp, lane);
      const float ls0 = lses[r0], ls1 = lses[r1], dr0 = Dr[r0], dr1 = Dr[r1];
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
      mma_regs_by_cols<D, PC / 16>(da, Ks, dq, lane);
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
#pragma unroll
      for (int i = 0; i < DP / 8; ++i) {               // padded pixels carry P = dS = 0, so no masking is needed
        csk[i][0] += dk[i][0] + dk[i][2]; csk[i][1] += dk[i][1] + dk[i][3];
        csv[i][0] += dv[i][0] + dv[i][2]; csv[i][1] += dv[i][1] + dv[i][3];
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
    __syncthreads();                                   // P / dS tiles and this K / V buffer are free again
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) {
    const int col = i * 8 + 2 * t;
    *reinterpret_cast<float2*>(dqp + r0 * DP + col) = make_float2(dq[i][0] * scale, dq[i][1] * scale);
    *reinterpret_cast<float2*>(dqp + r1 * DP + col) = make_float2(dq[i][2] * scale, dq[i][3] * scale);
  }
  if (dkv_colsum != nullptr) {                         // column sums of this CTA's dK / dV rows: reduce the 8 row groups of a warp, then the 4 warps
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = csk[i][e], v = csv[i][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
        if (g == 0) { csw[(warp * 2 + 0) * DP + i * 8 + 2 * t + e] = a * scale; csw[(warp * 2 + 1) * DP + i * 8 + 2 * t + e] = v; }
      }
    }
  }
  cluster.sync();
  if (dkv_colsum != nullptr) {
    for (int idx = tid; idx < 2 * D; idx += NT) {
      const int which = idx / D, j = idx % D;
      const float v = csw[(0 * 2 + which) * DP + j] + csw[(1 * 2 + which) * DP + j] + csw[(2 * 2 + which) * DP + j] + csw[(3 * 2 + which) * DP + j];
      atomicAdd(dkv_colsum + which * Cp + head * D + j, v);
    }
  }
  __syncthreads();
  if (tid < DP) csw[tid] = 0.f;                        // reused: column sums of this CTA's dQ rows
  __syncthreads();
  const int rows_mine = (NQ - split + nsplit - 1) / nsplit;
  for (int idx = tid; idx < rows_mine * D; idx += NT) {
    const int r = split + (idx / D) * nsplit, j = idx % D;
    float acc = 0.f;
    for (int pr = 0; pr < nsplit; ++pr) acc += cluster.map_shared_rank(dqp, pr)[r * DP + j];
    dm[qoff + (long)r * Cp + j] = acc;
    if (dm_lo != nullptr) dm_lo[qoff + (long)r * Cp + j] = __float2bfloat16_rn(acc);
    if (dm_colsum != nullptr) atomicAdd(&csw[j], acc);
  }
  if (dm_colsum != nullptr) {
    __syncthreads();
    if (tid < D) atomicAdd(dm_colsum + head * D + tid, csw[tid]);
  }
  cluster.sync();
}

template <int D> constexpr int fwd_smem() { return (QR + 4 * 128) * Geo<D>::PITCH * 2 + (QR * Geo<D>::DP + 2 * QR) * 4; }
template <int D> constexpr int bwd_smem() { return (2 * QR + 4 * 64) * Geo<D>::PITCH * 2 + 2 * QR * (64 + 8) * 2 + (2 * QR + QR * Geo<D>::DP + 8 * Geo<D>::DP) * 4; }

// CTAs per (image, head): enough to put about two CTAs on every SM, at most one per chunk, at most the portable cluster size
inline int pick_nsplit(int nchunks, int bh) {
  int ns = (2 * 148 + bh - 1) / bh;
  if (ns > 8) ns = 8;
  if (ns > nchunks) ns = nchunks;
  return ns < 1 ? 1 : ns;
}

template <int D>
int launch_fwd(const void* m, const void* kv, int B, int HW, int heads, float* out, float* lse, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_mma_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<D>());
    if (e != cudaSuccess) { dfb_set_error("gaa_mma_fwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int nchunks = dfb_cdiv(HW, 128);
  const int nsplit = pick_nsplit(nchunks, B * heads);
  dim3 grid(nsplit, B * heads);
  dfb_launch_cluster(gaa_mma_fwd_kernel<D>, grid, NT, fwd_smem<D>(), st, nsplit, (const bf16*)m, (const bf16*)kv, HW, heads, 1.0f / sqrtf((float)D), nchunks, out, lse);
  return dfb_check_launch("gaa_mma_fwd");
}

template <int D>
int launch_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, float* dm, void* dkv,
               float* dkv_colsum, float* dm_colsum, void* dm_lo, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_mma_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<D>());
    if (e != cudaSuccess) { dfb_set_error("gaa_mma_bwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int nchunks = dfb_cdiv(HW, 64);
  const int nsplit = pick_nsplit(nchunks, B * heads);
  dim3 grid(nsplit, B * heads);
  dfb_launch_cluster(gaa_mma_bwd_kernel<D>, grid, NT, bwd_smem<D>(), st, nsplit, dout, out, lse, (const bf16*)m, (const bf16*)kv, HW, heads,
                     1.0f / sqrtf((float)D), nchunks, dm, (bf16*)dkv, dkv_colsum, dm_colsum, (bf16*)dm_lo);
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
  GAA_MMA_DISPATCH_D(d, { return launch_fwd<D>(m, kv, B, HW, heads, out, lse, st); });
  return DFB_OK;
}

int dfb_gaa_mma_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, int d, float* dm,
                    void* dkv, float* dkv_colsum, float* dm_colsum, void* dm_lo, cudaStream_t st) {
  GAA_MMA_DISPATCH_D(d, { return launch_bwd<D>(dout, out, lse, m, kv, B, HW, heads, dm, dkv, dkv_colsum, dm_colsum, dm_lo, st); });
  return DFB_OK;
}
