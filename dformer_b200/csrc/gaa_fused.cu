// Global Awareness Attention core (DFormer.py:122-130) as ONE kernel per direction -- the exact-fp32 (CUDA-core) form.
// The bf16 configuration runs the same decomposition on the tensor cores (gaa_mma.cu); the C entry points at the end of this
// file dispatch on the dtype.
//
// The 49 pooled query tokens of an image attend over all H*W pixel keys/values, softmax over the pixels:
//   S = scale * Q K^T  [49, HW]      P = softmax_HW(S)      O = P V  [49, d]             per (image, head)
// The unfused form ran three CUDA-core GEMM launches plus a row softmax (five launches backward) and kept the
// [B, heads, 49, HW] probabilities for the backward pass.  Here a CTA owns 128 pixels of one (image, head):
//
//   forward   scores for its pixels -> local row max / sum -> unnormalised partial context, written next to (max, sum);
//             the last CTA of an (image, head) to finish (atomic ticket) merges the partials flash-decoding style and
//             stores O and the row log-sum-exp.  Nothing of size 49 x HW reaches HBM.
//   backward  P is recomputed from Q, K and the saved log-sum-exp; dV and dK of the CTA's pixels are complete in
//             registers, dQ partial sums go out as one fp32 atomic per element (dm is zeroed by the launcher).
//             rowsum(dP o P) = rowsum(dO o O), so no second pass over the pixels is needed.
//
// Thread pair (lanes 2i, 2i+1) = one pixel; each lane keeps half of the pixel's k, v (and dk, dv) rows in registers and
// the two halves of every dot product meet through one shuffle.  Q / dO rows are broadcast reads from shared memory.
#include <string.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int NQ = 49;             // pooled query tokens (7 x 7)
// Threads per CTA (NT) and threads per pixel (TPP) are template parameters.  Forward: NT = 256, thread pair per pixel (128-pixel
// chunks keep the number of partials to merge small).  Backward: NT = 128, thread pair per pixel (64-pixel chunks, 3-4 CTAs per SM).
// Measured on B200 (stage 1 / 2 / 3 shapes of DFormer-L, batch 8, us): NT 256 TPP 2: 136 / 56 / 37;  NT 128 TPP 2: 114 / 57 / 37;
// NT 256 TPP 4 (two thread pairs per pixel splitting the 49 query rows): 188 / 61 / 42.
constexpr int NT_FWD = 256, NT_BWD = 128, TPP_BWD = 2;

// Cooperative, coalesced staging of a [PC pixels][D] slice of kv (row pitch `pitch` elements) into shared memory as fp32:
// consecutive threads fetch consecutive 32-bit words of a pixel row (per-lane row loads would touch one sector per lane).
template <typename T, int D, int NT, int PC>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, long pitch, int nvalid, float* __restrict__ dst) {
  constexpr int WPR = D / 2;
#pragma unroll
  for (int i = threadIdx.x; i < PC * WPR; i += NT) {        // fully unrolled: all loads of a thread are in flight together
    const int p = i / WPR, w = i % WPR;
    float2 f = make_float2(0.f, 0.f);
    if (p < nvalid) {
      if constexpr (sizeof(T) == 2) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(src + (long)p * pitch + 2 * w);
        f = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
      } else {
        f = *reinterpret_cast<const float2*>(src + (long)p * pitch + 2 * w);
      }
    }
    *reinterpret_cast<float2*>(dst + p * D + 2 * w) = f;
  }
}
template <typename T>
__device__ __forceinline__ void store_pair(T* dst, float a, float b) {
  if constexpr (sizeof(T) == 2) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(&h);
  } else {
    *reinterpret_cast<float2*>(dst) = make_float2(a, b);
  }
}

template <int D, int NT>
struct Geo {
  static constexpr int DH = D / 2;
  static constexpr int G = NT / D;                        // row groups of the [49, D] output phase
  static constexpr int RPG = (NQ + G - 1) / G;            // rows per group
};

template <typename T, int D, int NT>
__global__ void __launch_bounds__(NT) gaa_fused_fwd_kernel(const T* __restrict__ m, const T* __restrict__ kv, int HW, int heads, float scale, int nchunks,
                                                          float* __restrict__ out, float* __restrict__ lse, float* __restrict__ part, int* __restrict__ counters) {
  pdl_sync();
  constexpr int PC = NT / 2, SP = PC + 4;
  constexpr int DH = Geo<D, NT>::DH, G = Geo<D, NT>::G, RPG = Geo<D, NT>::RPG, PS = D + 4;
  extern __shared__ __align__(16) float smf[];
  float* Qs = smf;                       // [49][D]
  float* Vs = Qs + NQ * D;               // [PC][D]
  float* Ss = Vs + PC * D;               // [49][SP]
  float* ms = Ss + NQ * SP;              // [49] local row max
  float* ls = ms + NQ;                   // [49] local row sum
  __shared__ int s_last;
  const int tid = threadIdx.x, bh = blockIdx.y, b = bh / heads, head = bh % heads, c = blockIdx.x;
  const int Cp = heads * D;
#ifdef GAA_TIMING
  long long tt[8]; int ti = 0;
#define TICK() do { __syncthreads(); tt[ti++] = clock64(); } while (0)
#else
#define TICK() do {} while (0)
#endif
  TICK();
#pragma unroll
  for (int i = tid; i < NQ * D; i += NT) Qs[i] = to_f(m[((long)b * NQ + i / D) * Cp + head * D + i % D]);
  TICK();
  const int p = tid >> 1, half = tid & 1;
  const int pix = c * PC + p;
  const bool valid = pix < HW;
  {
    const T* rows = kv + ((long)b * HW + (long)c * PC) * 2 * Cp + head * D;
    const int nvalid = min(PC, HW - c * PC);
    stage_rows<T, D, NT, PC>(rows, 2L * Cp, nvalid, Ss);            // K chunk, parked in the (not yet used) score tile
    stage_rows<T, D, NT, PC>(rows + Cp, 2L * Cp, nvalid, Vs);       // V chunk
  }
  __syncthreads();
  TICK();
  // ---- scores of this CTA's pixels
  float2 k2[DH / 2];
#pragma unroll
  for (int j = 0; j < DH / 2; ++j) k2[j] = *reinterpret_cast<const float2*>(Ss + p * D + half * DH + 2 * j);
  __syncthreads();                                          // every thread holds its k half-row before scores overwrite the tile
#pragma unroll 7
  for (int r = 0; r < NQ; ++r) {
    const float2* q = reinterpret_cast<const float2*>(Qs + r * D + half * DH);      // 8-byte aligned: D and DH are even
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) ffma2(s2, q[j], k2[j]);
    float s = s2.x + s2.y;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (half == (r & 1)) Ss[r * SP + p] = valid ? s * scale : -INFINITY;
  }
  __syncthreads();
  TICK();
  // ---- local softmax statistics, P~ = exp(S - local max) in place
  for (int r = tid >> 5; r < NQ; r += NT / 32) {
    const int lane = tid & 31;
    float v[PC / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < PC / 32; ++i) { v[i] = Ss[r * SP + lane + 32 * i]; mx = fmaxf(mx, v[i]); }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PC / 32; ++i) {
      const float e = __expf(v[i] - mx);                  // exp(-inf) = 0 for the padded tail (pixel 0 of a chunk is always real)
      Ss[r * SP + lane + 32 * i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) { ms[r] = mx; ls[r] = sum; }
  }
  __syncthreads();
  TICK();
  // ---- unnormalised partial context O~[r][j] = sum_p P~[r][p] V[p][j]
  {
    const int j = tid % D, g = tid / D;
    if (g < G) {
      float2 acc[RPG];
#pragma unroll
      for (int i = 0; i < RPG; ++i) acc[i] = make_float2(0.f, 0.f);
      for (int p4 = 0; p4 < PC; p4 += 4) {
        const float2 v01 = make_float2(Vs[(p4 + 0) * D + j], Vs[(p4 + 1) * D + j]), v23 = make_float2(Vs[(p4 + 2) * D + j], Vs[(p4 + 3) * D + j]);
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
          const int r = g + i * G;
          if (r < NQ) {                                   // even / odd pixels accumulate in the two halves of a packed FFMA2
            const float4 pr = *reinterpret_cast<const float4*>(Ss + r * SP + p4);
            ffma2(acc[i], make_float2(pr.x, pr.y), v01);
            ffma2(acc[i], make_float2(pr.z, pr.w), v23);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPG; ++i) {
        const int r = g + i * G;
        if (r < NQ) {
          float* dst = part + (((long)bh * nchunks + c) * NQ + r) * PS;
          dst[j] = acc[i].x + acc[i].y;
          if (j == 0) { dst[D] = ms[r]; dst[D + 1] = ls[r]; }
        }
      }
    }
  }
  // ---- the last CTA of this (image, head) merges all partials
  TICK();
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[bh], 1) == nchunks - 1);
  __syncthreads();
  TICK();
#ifdef GAA_TIMING
  if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 3) printf("gaa fwd phases (cycles): Q %lld  KV %lld  score %lld  stats %lld  PV %lld  fence+ticket %lld\n", tt[1]-tt[0], tt[2]-tt[1], tt[3]-tt[2], tt[4]-tt[3], tt[5]-tt[4], tt[6]-tt[5]);
#endif
  if (!s_last) return;
  __threadfence();
  // (1) all (max, sum) pairs of this (image, head) -> shared memory, independent loads
  float* pm = ls + NQ;                   // [nchunks][49] chunk maxima -> weights exp(m_c - M)   (tail of the dynamic allocation)
  float* pl = pm + nchunks * NQ;         // [nchunks][49] chunk sums
  const float* base = part + (long)bh * nchunks * NQ * PS;
  for (int i = tid; i < nchunks * NQ; i += NT) {
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(base + (long)i * PS + D));
    pm[i] = ml.x;
    pl[i] = ml.y;
  }
  __syncthreads();
  // (2) per row: global max, total sum, per-chunk weights
  for (int r = tid >> 5; r < NQ; r += NT / 32) {
    const int lane = tid & 31;
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
  // (3) O[r][j..j+3] = sum_c w[c][r] * O~_c[r][j..j+3] / L   (float4 loads, 4 chunks in flight per thread)
  constexpr int D4 = D / 4;
  for (int idx = tid; idx < NQ * D4; idx += NT) {
    const int r = idx / D4, j = (idx % D4) * 4;
    const float* src = base + (long)r * PS + j;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int cc = 0; cc < nchunks; ++cc) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long)cc * NQ * PS));
      const float w = pm[cc * NQ + r];
      o.x = fmaf(v.x, w, o.x); o.y = fmaf(v.y, w, o.y); o.z = fmaf(v.z, w, o.z); o.w = fmaf(v.w, w, o.w);
    }
    const float inv = ms[r];
    *reinterpret_cast<float4*>(out + ((long)b * NQ + r) * Cp + head * D + j) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
  }
  if (tid == 0) counters[bh] = 0;                        // self-resetting ticket: the buffer is reusable by the next launch
}

// TPP threads per pixel: 2 = (half-row, half-row); 4 = two such pairs that split the 49 query rows between them (even / odd), which
// halves the serial row loop -- the latency that bounds this kernel -- and meet again through one shuffle per accumulator at the end.
template <typename T, int D, int NT, int TPP>
__global__ void __launch_bounds__(NT) gaa_fused_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ lse,
                                                          const T* __restrict__ m, const T* __restrict__ kv, int HW, int heads, float scale,
                                                          float* __restrict__ dm, T* __restrict__ dkv) {
  pdl_sync();
  constexpr int PC = NT / TPP, SP = PC + 4, RGS = TPP / 2;
  constexpr int DH = Geo<D, NT>::DH, G = Geo<D, NT>::G, RPG = Geo<D, NT>::RPG;
  extern __shared__ __align__(16) float smf[];
  float* Qs = smf;                       // [49][D]
  float* dOs = Qs + NQ * D;              // [49][D]
  float* Ks = dOs + NQ * D;              // [PC][D]
  float* dSs = Ks + PC * D;              // [49][SP]
  float* lses = dSs + NQ * SP;           // [49]
  float* Dr = lses + NQ;                 // [49] rowsum(dO o O) = rowsum(dP o P)
  const int tid = threadIdx.x, bh = blockIdx.y, b = bh / heads, head = bh % heads, c = blockIdx.x;
  const int Cp = heads * D;
#ifdef GAA_TIMING
  long long tt[8]; int ti = 0;
#endif
  TICK();
  {
    // every thread's loads (Q, dO, O) are issued before any is consumed; Dr[r] = sum_j dO[r][j] O[r][j] is then reduced from
    // shared memory (dO) and a staged copy of O parked in the (not yet used) dS tile
    float* Os = dSs;
#pragma unroll
    for (int i = tid; i < NQ * D; i += NT) {
      const long g = ((long)b * NQ + i / D) * Cp + head * D + i % D;
      Qs[i] = to_f(m[g]);
      dOs[i] = dout[g];
      Os[i] = out[g];
    }
    if (tid < NQ) lses[tid] = lse[(long)bh * NQ + tid];
    __syncthreads();
    for (int r = tid >> 5; r < NQ; r += NT / 32) {
      float s = 0.f;
      for (int j = tid & 31; j < D; j += 32) s = fmaf(dOs[r * D + j], Os[r * D + j], s);
      s = warp_sum(s);
      if ((tid & 31) == 0) Dr[r] = s;
    }
    __syncthreads();                                        // Os (in the dS tile) is dead before V is staged there
  }
  TICK();
  const int p = tid / TPP, half = tid & 1, rg = (tid >> 1) & (RGS - 1);
  const int pix = c * PC + p;
  const bool valid = pix < HW;
  float2 k2[DH / 2], v2[DH / 2], dk2[DH / 2], dv2[DH / 2];
  const long rowoff = ((long)b * HW + (valid ? pix : 0)) * 2 * Cp + head * D + half * DH;
  {
    const T* rows = kv + ((long)b * HW + (long)c * PC) * 2 * Cp + head * D;
    const int nvalid = min(PC, HW - c * PC);
    stage_rows<T, D, NT, PC>(rows, 2L * Cp, nvalid, Ks);        // K chunk (kept: phase 2 reads it column-wise)
    stage_rows<T, D, NT, PC>(rows + Cp, 2L * Cp, nvalid, dSs);  // V chunk, parked in the (not yet used) dS tile
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < DH / 2; ++j) {
    k2[j] = *reinterpret_cast<const float2*>(Ks + p * D + half * DH + 2 * j);
    v2[j] = *reinterpret_cast<const float2*>(dSs + p * D + half * DH + 2 * j);
    dk2[j] = dv2[j] = make_float2(0.f, 0.f);
  }
  __syncthreads();                                          // v half-rows are in registers before dS overwrites the tile
  TICK();
  constexpr int ROW_ITERS = (NQ + RGS - 1) / RGS;             // warp-uniform trip count (the shuffles below use the full mask)
#pragma unroll 2
  for (int it = 0; it < ROW_ITERS; ++it) {
    const int r_raw = rg + it * RGS;
    const bool row_ok = r_raw < NQ;
    const int r = row_ok ? r_raw : NQ - 1;
    const float2* q = reinterpret_cast<const float2*>(Qs + r * D + half * DH);
    const float2* go = reinterpret_cast<const float2*>(dOs + r * D + half * DH);
    float2 s2 = make_float2(0.f, 0.f), dp2 = make_float2(0.f, 0.f);
    float2 qv[DH / 2], gv[DH / 2];
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) { qv[j] = q[j]; gv[j] = go[j]; ffma2(s2, qv[j], k2[j]); ffma2(dp2, gv[j], v2[j]); }
    float sd = s2.x + s2.y, dp = dp2.x + dp2.y;
    sd += __shfl_xor_sync(0xffffffffu, sd, 1);
    dp += __shfl_xor_sync(0xffffffffu, dp, 1);
    const float P = (valid && row_ok) ? __expf(fmaf(sd, scale, -lses[r])) : 0.f;
    const float ds = P * (dp - Dr[r]);
    const float2 P2 = make_float2(P, P), ds2 = make_float2(ds, ds);
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) { ffma2(dv2[j], P2, gv[j]); ffma2(dk2[j], ds2, qv[j]); }
    if (row_ok && half == (RGS == 1 ? (r & 1) : 0)) dSs[r * SP + p] = ds;
  }
  TICK();
  if (RGS == 2) {                                             // the two row groups of a pixel add up their partial dK / dV
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) {
      dk2[j].x += __shfl_xor_sync(0xffffffffu, dk2[j].x, 2); dk2[j].y += __shfl_xor_sync(0xffffffffu, dk2[j].y, 2);
      dv2[j].x += __shfl_xor_sync(0xffffffffu, dv2[j].x, 2); dv2[j].y += __shfl_xor_sync(0xffffffffu, dv2[j].y, 2);
    }
  }
  if (valid) {
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) {
      if (RGS == 1 || rg == 0) store_pair<T>(dkv + rowoff + 2 * j, dk2[j].x * scale, dk2[j].y * scale);
      if (RGS == 1 || rg == 1) store_pair<T>(dkv + rowoff + Cp + 2 * j, dv2[j].x, dv2[j].y);
    }
  }
  __syncthreads();
  TICK();
  // ---- dQ[r][j] += scale * sum_p dS[r][p] K[p][j]   (partial over this CTA's pixels)
  const int j = tid % D, g = tid / D;
  if (g < G) {
    float2 acc[RPG];
#pragma unroll
    for (int i = 0; i < RPG; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int p4 = 0; p4 < PC; p4 += 4) {
      const float2 k01 = make_float2(Ks[(p4 + 0) * D + j], Ks[(p4 + 1) * D + j]), k23 = make_float2(Ks[(p4 + 2) * D + j], Ks[(p4 + 3) * D + j]);
#pragma unroll
      for (int i = 0; i < RPG; ++i) {
        const int r = g + i * G;
        if (r < NQ) {
          const float4 d4 = *reinterpret_cast<const float4*>(dSs + r * SP + p4);
          ffma2(acc[i], make_float2(d4.x, d4.y), k01);
          ffma2(acc[i], make_float2(d4.z, d4.w), k23);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RPG; ++i) {
      const int r = g + i * G;
      if (r < NQ) atomicAdd(dm + ((long)b * NQ + r) * Cp + head * D + j, (acc[i].x + acc[i].y) * scale);
    }
  }
  TICK();
#ifdef GAA_TIMING
  if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 3) printf("gaa bwd phases (cycles): QdO+Dr %lld  KV %lld  rloop %lld  store %lld  dQ+atomics %lld\n", tt[1]-tt[0], tt[2]-tt[1], tt[3]-tt[2], tt[4]-tt[3], tt[5]-tt[4]);
#endif
}

template <int D, int NT> constexpr int fwd_smem() { return (NQ * D + (NT / 2) * D + NQ * (NT / 2 + 4) + 2 * NQ) * 4; }
template <int D, int PC> constexpr int bwd_smem() { return (2 * NQ * D + PC * D + NQ * (PC + 4) + 2 * NQ) * 4; }

template <typename T, int D>
int launch_fwd(const void* m, const void* kv, int B, int HW, int heads, float* out, float* lse, float* part, int* counters, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_fused_fwd_kernel<T, D, NT_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { dfb_set_error("gaa_fused_fwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  const int nchunks = dfb_cdiv(HW, NT_FWD / 2);
  const int smem = fwd_smem<D, NT_FWD>() + 2 * nchunks * NQ * 4;            // + merge staging of the (max, sum) pairs
  if (smem > 200 * 1024) { dfb_set_error("gaa_fused_fwd: HW=%d too large for the one-launch merge", HW); return DFB_ERR_UNSUPPORTED; }
  dim3 grid(nchunks, B * heads);
  dfb_launch(gaa_fused_fwd_kernel<T, D, NT_FWD>, grid, NT_FWD, smem, st, (const T*)m, (const T*)kv, HW, heads, 1.0f / sqrtf((float)D), nchunks, out, lse, part, counters);
  return dfb_check_launch("gaa_fused_fwd");
}

template <typename T, int D>
int launch_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, float* dm, void* dkv,
               cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gaa_fused_bwd_kernel<T, D, NT_BWD, TPP_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<D, NT_BWD / TPP_BWD>());
    if (e != cudaSuccess) { dfb_set_error("gaa_fused_bwd smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    attr = true;
  }
  cudaMemsetAsync(dm, 0, sizeof(float) * (size_t)B * NQ * heads * D, st);
  dim3 grid(dfb_cdiv(HW, NT_BWD / TPP_BWD), B * heads);
  dfb_launch(gaa_fused_bwd_kernel<T, D, NT_BWD, TPP_BWD>, grid, NT_BWD, bwd_smem<D, NT_BWD / TPP_BWD>(), st, dout, out, lse, (const T*)m, (const T*)kv, HW, heads, 1.0f / sqrtf((float)D), dm, (T*)dkv);
  return dfb_check_launch("gaa_fused_bwd");
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)
#define GAA_DISPATCH_D(d, ...)                                        \
  switch (d) {                                                        \
    case 16: { constexpr int D = 16; __VA_ARGS__ } break;             \
    case 32: { constexpr int D = 32; __VA_ARGS__ } break;             \
    case 36: { constexpr int D = 36; __VA_ARGS__ } break;             \
    case 48: { constexpr int D = 48; __VA_ARGS__ } break;             \
    default: dfb_set_error("gaa_fused: head dim %d not instantiated (16, 32, 36, 48)", d); return DFB_ERR_UNSUPPORTED; \
  }

extern "C" int dfb200_gaa_fused_fwd(const void* m, const void* kv, int dtype, int B, int HW, int heads, int d, float* out, float* lse, float* scratch,
                                    int* counters, void* stream) {
  DFB_REQUIRE(B > 0 && HW > 0 && heads > 0, "gaa_fused_fwd: empty problem");
  DFB_REQUIRE(dtype == 0 || dtype == 1, "gaa_fused_fwd: bad dtype %d", dtype);
  if (dtype == 1) return dfb_gaa_mma_fwd(m, kv, B, HW, heads, d, out, lse, scratch, counters, ST);      // bf16: tensor cores (gaa_mma.cu)
  GAA_DISPATCH_D(d, { return launch_fwd<float, D>(m, kv, B, HW, heads, out, lse, scratch, counters, ST); });
  return DFB_OK;
}

extern "C" int dfb200_gaa_fused_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int dtype, int B, int HW,
                                    int heads, int d, float* dm, void* dkv, void* stream) {
  DFB_REQUIRE(B > 0 && HW > 0 && heads > 0, "gaa_fused_bwd: empty problem");
  DFB_REQUIRE(dtype == 0 || dtype == 1, "gaa_fused_bwd: bad dtype %d", dtype);
  if (dtype == 1) return dfb_gaa_mma_bwd(dout, out, lse, m, kv, B, HW, heads, d, dm, dkv, nullptr, nullptr, nullptr, ST);      // bf16: tensor cores (gaa_mma.cu)
  GAA_DISPATCH_D(d, { return launch_bwd<float, D>(dout, out, lse, m, kv, B, HW, heads, dm, dkv, ST); });
  return DFB_OK;
}

// bf16 only: the backward pass that also emits what the two projections around the attention core need next --
// dkv_colsum[2*heads*d] += column sums of dkv (bias gradient of `kv`, DFormer.py:121), dm_colsum[heads*d] += column sums of dm
// (bias gradient of `short_cut_linear`, DFormer.py:108), dm_lo = dm rounded to bf16 (operand of that layer's gradient GEMMs).
extern "C" int dfb200_gaa_fused_bwd_ex(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads,
                                       int d, float* dm, void* dkv, float* dkv_colsum, float* dm_colsum, void* dm_lo, void* stream) {
  DFB_REQUIRE(B > 0 && HW > 0 && heads > 0, "gaa_fused_bwd_ex: empty problem");
  return dfb_gaa_mma_bwd(dout, out, lse, m, kv, B, HW, heads, d, dm, dkv, dkv_colsum, dm_colsum, dm_lo, ST);
}
