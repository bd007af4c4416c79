// LayerNorm (channels-last, eps 1e-6) and BatchNorm2d kernels.  HBM-bound, one pass per tensor.
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int LN_MAX_PER_LANE = 32;   // C <= 1024

// ------------------------------------------------------------------ LayerNorm: one warp per row, NPL values per lane
// (NPL = ceil(C / 32) rounded up to a supported size; the row lives in registers between the two passes)
// Optional fused layer-scale residual in front of the LayerNorm (DFormer.py:173-179 followed by :59/:104):
//   x_out[m, c] = x[m, c] + scale_b[sample(m)] * ls[c] * y[m, c]   is written once and normalised in the same pass.
struct LnResidual {
  const void* y; long ldy; const float* ls; const float* scale_b; int rows_per_sample; float* x_out;
};

template <typename T, int NPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                     int M, int C, T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, const LnResidual rp) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float gm[NPL], bt[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + i * 32;
    gm[i] = c < C ? gamma[c] : 0.f;
    bt[i] = c < C ? beta[c] : 0.f;
  }
  const float invC = 1.f / (float)C;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    const float* xr = x + (long)row * C;
    float v[NPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < C) ? __ldg(xr + c) : 0.f;
      if (rp.y && c < C) {
        const float sb = rp.scale_b ? rp.scale_b[row / rp.rows_per_sample] : 1.f;
        v[i] = fmaf(sb * rp.ls[c], to_f(reinterpret_cast<const T*>(rp.y)[(long)row * rp.ldy + c]), v[i]);
        rp.x_out[(long)row * C + c] = v[i];
      }
      s += v[i];
    }
    const float mu = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * 32;
      const float d = (c < C) ? v[i] - mu : 0.f;
      q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    T* yr = y + (long)row * C;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * 32;
      if (c < C) yr[c] = from_f<T>((v[i] - mu) * rs * gm[i] + bt[i]);
    }
  }
}

// Backward: warp per row for dx; dgamma/dbeta partials stay in registers over the warp's rows, are combined across
// the block's warps in shared memory and leave as one atomicAdd per channel per block.
template <typename T, int NPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const float* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int M,
                                                     int C, const float* dx_in, float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  extern __shared__ float sm[];   // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float pg[NPL], pb[NPL], gm[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    pg[i] = 0.f; pb[i] = 0.f;
    const int c = lane + i * 32;
    gm[i] = c < C ? gamma[c] : 0.f;
  }
  const float invC = 1.f / (float)C;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    const float* xr = x + (long)row * C;
    const T* gr = dy + (long)row * C;
    const T* gr2 = dy2 ? dy2 + (long)row * C : nullptr;
    const float mu = mean[row], rs = rstd[row];
    float xh[NPL], g[NPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * 32;
      if (c < C) {
        const float d = gr2 ? to_f(from_f<T>(to_f(gr[c]) + to_f(gr2[c]))) : to_f(gr[c]);      // the sum is rounded like the separate axpy pass was
        xh[i] = (__ldg(xr + c) - mu) * rs;
        g[i] = d * gm[i];
        pg[i] = fmaf(d, xh[i], pg[i]);
        pb[i] += d;
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
      } else { xh[i] = 0.f; g[i] = 0.f; }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    float* dr = dx + (long)row * C;
    const float* di = dx_in ? dx_in + (long)row * C : nullptr;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * 32;
      if (c < C) {
        const float v = rs * (g[i] - s1 - xh[i] * s2);
        dr[c] = di ? di[c] + v : v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + i * 32;
    if (c < C) { atomicAdd(&sm[c], pg[i]); atomicAdd(&sm[C + c], pb[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(dgamma + c, sm[c]); atomicAdd(dbeta + c, sm[C + c]); }
}

// ------------------------------------------------------------------ vectorised LayerNorm (C % 8 == 0)
// L lanes cooperate on one row (L = 8, 16 or 32 so that small C still fills the warp with several rows); each lane owns
// VPL 8-channel vectors -> every global access is a 16/32-byte vector and the row stays in registers.
template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int L, int VPL>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                         int M, int C, T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
                                                         const LnResidual rp) {
  pdl_sync();
  constexpr int RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rsel = lane / L;
  const int nvec = C >> 3;
  const int groups_per_block = (blockDim.x >> 5) * RPW;
  float gm[VPL][8], bt[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = sub + i * L;
    if (v < nvec) { Vec8<float>::load(gamma + v * 8, gm[i]); Vec8<float>::load(beta + v * 8, bt[i]); }
  }
  const float invC = 1.f / (float)C;
  // the trip count is warp-uniform (shuffles below use the full mask); rows past M are masked, not skipped
  for (long row0 = (long)blockIdx.x * groups_per_block + (threadIdx.x >> 5) * RPW; row0 < M; row0 += (long)gridDim.x * groups_per_block) {
    const long row = row0 + rsel;
    const bool valid = row < M;
    const float* xr = x + (valid ? row : 0) * C;
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec) {
        Vec8<float>::load(xr + vv * 8, v[i]);
        if (rp.y && valid) {
          const float sb = rp.scale_b ? rp.scale_b[row / rp.rows_per_sample] : 1.f;
          float yv[8], lv[8];
          Vec8<T>::load(reinterpret_cast<const T*>(rp.y) + row * rp.ldy + vv * 8, yv);
          Vec8<float>::load(rp.ls + vv * 8, lv);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = fmaf(sb * lv[j], yv[j], v[i][j]);
          Vec8<float>::store(rp.x_out + row * C + vv * 8, v[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mu = group_sum<L>(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (sub + i * L < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rs = rsqrtf(group_sum<L>(q) * invC + eps);
    if (!valid) continue;
    if (sub == 0) { mean[row] = mu; rstd[row] = rs; }
    T* yr = y + row * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * gm[i][j] + bt[i][j];
        Vec8<T>::store(yr + vv * 8, o);
      }
    }
  }
}

template <typename T, int L, int VPL>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const float* __restrict__ x,
                                                         const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         int M, int C, const float* dx_in, float* dx, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
  pdl_sync();
  extern __shared__ float sm[];   // [8 warps][2][C] per-warp partial column sums
  constexpr int RPW = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rsel = lane / L;
  const int nvec = C >> 3;
  const int groups_per_block = (blockDim.x >> 5) * RPW;
  float gm[VPL][8], pg[VPL][8], pb[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = sub + i * L;
    if (v < nvec) Vec8<float>::load(gamma + v * 8, gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; }
  }
  const float invC = 1.f / (float)C;
  for (long row0 = (long)blockIdx.x * groups_per_block + (threadIdx.x >> 5) * RPW; row0 < M; row0 += (long)gridDim.x * groups_per_block) {
    const long row = row0 + rsel;
    const bool valid = row < M;                     // warp-uniform trip count; masked rows contribute zeros
    const float mu = valid ? mean[row] : 0.f, rs = valid ? rstd[row] : 0.f;
    float xh[VPL][8], g[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec && valid) {
        float d[8], xv[8];
        Vec8<T>::load(dy + row * C + vv * 8, d);
        if (dy2) {                                  // fused gradient fan-in (was a separate axpy pass): sum rounded to T like that pass did
          float d2[8];
          Vec8<T>::load(dy2 + row * C + vv * 8, d2);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = to_f(from_f<T>(d[j] + d2[j]));
        }
        Vec8<float>::load(x + row * C + vv * 8, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xv[j] - mu) * rs;
          g[i][j] = d[j] * gm[i][j];
          pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
          pb[i][j] += d[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], xh[i][j], s2);
        }
      }
    }
    s1 = group_sum<L>(s1) * invC;
    s2 = group_sum<L>(s2) * invC;
    if (!valid) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec) {
        float o[8];
        if (dx_in) Vec8<float>::load(dx_in + row * C + vv * 8, o);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rs * (g[i][j] - s1 - xh[i][j] * s2);
        Vec8<float>::store(dx + row * C + vv * 8, o);
      }
    }
  }
  // column sums: lanes of a warp that own the same channels (different rows) meet through shuffles, every warp stores its
  // partial row to its own shared-memory slab (no atomics), then one thread per channel adds the 8 slabs and issues the
  // single global atomic of this CTA for that channel
  float* slab = sm + (threadIdx.x >> 5) * 2 * C;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vv = sub + i * L;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = pg[i][j], b = pb[i][j];
#pragma unroll
      for (int o = L; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (rsel == 0 && vv < nvec) { slab[vv * 8 + j] = a; slab[C + vv * 8 + j] = b; }
    }
  }
  __syncthreads();
  const int nwarp = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nwarp; ++w) v += sm[w * 2 * C + c];
    atomicAdd((c < C ? dgamma : dbeta - C) + c, v);
  }
}

// raw (unconverted) 8-element vectors: lets the loads of the NEXT row stay in flight while the current row is reduced
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* f) const { Vec8<bf16>::unpack(v, f); }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void get(float* f) const { f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w; }
};

template <typename T, int L>
__global__ void __launch_bounds__(256, 2) ln_bwd_vec_pf_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const float* __restrict__ x,
                                                         const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         int M, int C, const float* dx_in, float* dx, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
  pdl_sync();
  extern __shared__ float sm[];   // [8 warps][2][C] per-warp partial column sums
  constexpr int RPW = 32 / L, VPL = 1;
  const int lane = threadIdx.x & 31, sub = lane % L, rsel = lane / L;
  const int nvec = C >> 3;
  const int groups_per_block = (blockDim.x >> 5) * RPW;
  float gm[VPL][8], pg[VPL][8], pb[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = sub + i * L;
    if (v < nvec) Vec8<float>::load(gamma + v * 8, gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; }
  }
  const float invC = 1.f / (float)C;
  // Software pipeline (one-vector-per-lane shapes, C <= 256: the register budget allows it): the loads of the next row of this warp
  // are issued before the current row is reduced, doubling the bytes in flight of a kernel that runs at 16 warps per SM.
  constexpr bool PF = (VPL == 1);
  Raw8<T> nd[VPL], nd2[VPL];
  Raw8<float> nx[VPL], ni[VPL];
  float nmu = 0.f, nrs = 0.f;
  const long stride = (long)gridDim.x * groups_per_block;
  auto fetch = [&](long row0) {
    const long row = row0 + rsel;
    const bool valid = row < M;
    nmu = valid ? mean[row] : 0.f;
    nrs = valid ? rstd[row] : 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec && valid) {
        nd[i].load(dy + row * C + vv * 8);
        if (dy2) nd2[i].load(dy2 + row * C + vv * 8);
        nx[i].load(x + row * C + vv * 8);
        if (dx_in) ni[i].load(dx_in + row * C + vv * 8);
      }
    }
  };
  long row0 = (long)blockIdx.x * groups_per_block + (threadIdx.x >> 5) * RPW;
  if (PF && row0 < M) fetch(row0);
  for (; row0 < M; row0 += stride) {
    const long row = row0 + rsel;
    const bool valid = row < M;                     // warp-uniform trip count; masked rows contribute zeros
    if (!PF) fetch(row0);
    const float mu = nmu, rs = nrs;
    float xh[VPL][8], g[VPL][8], o[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec && valid) {
        float d[8], xv[8];
        nd[i].get(d);
        if (dy2) {                                  // fused gradient fan-in (was a separate axpy pass): sum rounded to T like that pass did
          float d2[8];
          nd2[i].get(d2);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = to_f(from_f<T>(d[j] + d2[j]));
        }
        nx[i].get(xv);
        if (dx_in) ni[i].get(o[i]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xv[j] - mu) * rs;
          g[i][j] = d[j] * gm[i][j];
          pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
          pb[i][j] += d[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], xh[i][j], s2);
        }
      }
    }
    if (PF && row0 + stride < M) fetch(row0 + stride);          // next row's loads fly during the reductions and the store below
    s1 = group_sum<L>(s1) * invC;
    s2 = group_sum<L>(s2) * invC;
    if (!valid) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vv = sub + i * L;
      if (vv < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] += rs * (g[i][j] - s1 - xh[i][j] * s2);
        Vec8<float>::store(dx + row * C + vv * 8, o[i]);
      }
    }
  }
  // column sums: lanes of a warp that own the same channels (different rows) meet through shuffles, every warp stores its
  // partial row to its own shared-memory slab (no atomics), then one thread per channel adds the 8 slabs and issues the
  // single global atomic of this CTA for that channel
  float* slab = sm + (threadIdx.x >> 5) * 2 * C;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vv = sub + i * L;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = pg[i][j], b = pb[i][j];
#pragma unroll
      for (int o = L; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (rsel == 0 && vv < nvec) { slab[vv * 8 + j] = a; slab[C + vv * 8 + j] = b; }
    }
  }
  __syncthreads();
  const int nwarp = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nwarp; ++w) v += sm[w * 2 * C + c];
    atomicAdd((c < C ? dgamma : dbeta - C) + c, v);
  }
}

#define LN_DISPATCH_VEC(C, ...)                                                            \
  do {                                                                                     \
    const int nvec_ = (C) >> 3;                                                            \
    if (nvec_ <= 8) { constexpr int L = 8, VPL = 1; __VA_ARGS__ }                          \
    else if (nvec_ <= 16) { constexpr int L = 16, VPL = 1; __VA_ARGS__ }                   \
    else if (nvec_ <= 32) { constexpr int L = 32, VPL = 1; __VA_ARGS__ }                   \
    else if (nvec_ <= 64) { constexpr int L = 32, VPL = 2; __VA_ARGS__ }                   \
    else if (nvec_ <= 96) { constexpr int L = 32, VPL = 3; __VA_ARGS__ }                   \
    else { constexpr int L = 32, VPL = 4; __VA_ARGS__ }                                    \
  } while (0)

#define LN_DISPATCH_NPL(C, ...)                                      \
  do {                                                               \
    const int npl_ = ((C) + 31) / 32;                                \
    if (npl_ <= 1) { constexpr int NPL = 1; __VA_ARGS__ }            \
    else if (npl_ <= 2) { constexpr int NPL = 2; __VA_ARGS__ }       \
    else if (npl_ <= 3) { constexpr int NPL = 3; __VA_ARGS__ }       \
    else if (npl_ <= 4) { constexpr int NPL = 4; __VA_ARGS__ }       \
    else if (npl_ <= 6) { constexpr int NPL = 6; __VA_ARGS__ }       \
    else if (npl_ <= 9) { constexpr int NPL = 9; __VA_ARGS__ }       \
    else if (npl_ <= 16) { constexpr int NPL = 16; __VA_ARGS__ }     \
    else if (npl_ <= 18) { constexpr int NPL = 18; __VA_ARGS__ }     \
    else { constexpr int NPL = 32; __VA_ARGS__ }                     \
  } while (0)

// ------------------------------------------------------------------ BatchNorm
constexpr int BN_THREADS = 256;

// Column sums in double (block-level partials in float over <= rows_per_block rows).
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T* __restrict__ x, int M, int C, double* sum, double* sumsq, int rows_per_block) {
  pdl_sync();
  __shared__ float s1[BN_THREADS * 8];
  __shared__ float s2[BN_THREADS * 8];
  const int nvec_all = C >> 3;
  const int v0 = blockIdx.y * BN_THREADS;
  const int nvec = min(BN_THREADS, nvec_all - v0);
  const int rl_count = BN_THREADS / nvec, active = rl_count * nvec;
  const int t = threadIdx.x;
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (t < active) {
    const int cv = t % nvec, rl = t / nvec;
    const T* xp = x + (v0 + cv) * 8;
    int r = r0 + rl;
    // four rows in flight per thread: the loop is a dependent chain of DRAM latencies otherwise (one 16-byte load per ~0.8 us)
    for (; r + 3 * rl_count < r1; r += 4 * rl_count) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) Vec8<T>::load(xp + (long)(r + u * rl_count) * C, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { a1[j] += v[u][j]; a2[j] = fmaf(v[u][j], v[u][j], a2[j]); }
    }
    for (; r < r1; r += rl_count) {
      float v[8];
      Vec8<T>::load(xp + (long)r * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a1[j] += v[j]; a2[j] = fmaf(v[j], v[j], a2[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[t * 8 + j] = a1[j]; s2[t * 8 + j] = a2[j]; }
  __syncthreads();
  for (int i = t; i < nvec * 8; i += BN_THREADS) {
    const int v = i >> 3, j = i & 7;
    double d1 = 0.0, d2 = 0.0;
    for (int q = 0; q < rl_count; ++q) { d1 += s1[(q * nvec + v) * 8 + j]; d2 += s2[(q * nvec + v) * 8 + j]; }
    atomicAdd(sum + (v0 + v) * 8 + j, d1);
    atomicAdd(sumsq + (v0 + v) * 8 + j, d2);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count, float eps, float momentum, int C,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* running_mean, float* running_var) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sum[c] / count;
  double var = sumsq[c] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}
__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C, float* mean, float* invstd) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = rsqrtf(rv[c] + eps);
}

__device__ __forceinline__ float act_fwd(float z, int act) { return act == 1 ? gelu_f(z) : (act == 2 ? fmaxf(z, 0.f) : z); }
__device__ __forceinline__ float act_bwd(float z, int act) { return act == 1 ? gelu_grad_f(z) : (act == 2 ? (z > 0.f ? 1.f : 0.f) : 1.f); }

// Thread = one fixed 8-channel vector x a lane of rows (same decomposition as bn_stats): the per-channel affine
// z = x * (gamma * invstd) + (beta - mean * gamma * invstd) sits in 16 registers for the thread's whole row range, so the
// streaming loop is one 16-byte load, 8 FMAs (+ activation) and one 16-byte store.
template <typename TX, typename TY>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const TX* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const TY* __restrict__ residual, int act, const float* __restrict__ chan_scale,
                                                              int rows_per_sample, int M, int C, TY* __restrict__ y, int rows_per_block) {
  pdl_sync();
  const int nvec_all = C >> 3;
  const int v0 = blockIdx.y * BN_THREADS;
  const int nvec = min(BN_THREADS, nvec_all - v0);
  const int rl_count = BN_THREADS / nvec, active = rl_count * nvec;
  const int t = threadIdx.x;
  if (t >= active) return;
  const int cv = t % nvec, rl = t / nvec, c = (v0 + cv) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = gamma[c + j] * invstd[c + j];
    sh[j] = fmaf(-mean[c + j], sc[j], beta[c + j]);
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  auto finish = [&](int r, float* v, const float* res) {
    const long off = (long)r * C + c;
    const float* cs = chan_scale ? chan_scale + (long)(r / rows_per_sample) * C + c : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z = fmaf(v[j], sc[j], sh[j]);
      if (residual) z += res[j];
      z = act_fwd(z, act);
      if (cs) z *= cs[j];
      v[j] = z;
    }
    Vec8<TY>::store(y + off, v);
  };
  int r = r0 + rl;
  for (; r + rl_count < r1; r += 2 * rl_count) {          // two rows in flight per thread
    const long offa = (long)r * C + c, offb = (long)(r + rl_count) * C + c;
    float va[8], vb[8], ra[8], rb[8];
    Vec8<TX>::load(x + offa, va);
    Vec8<TX>::load(x + offb, vb);
    if (residual) { Vec8<TY>::load(residual + offa, ra); Vec8<TY>::load(residual + offb, rb); }
    finish(r, va, ra);
    finish(r + rl_count, vb, rb);
  }
  if (r < r1) {
    float v[8], res[8];
    Vec8<TX>::load(x + (long)r * C + c, v);
    if (residual) Vec8<TY>::load(residual + (long)r * C + c, res);
    finish(r, v, res);
  }
}

// ACT: 0 none, 1 GELU, 2 ReLU (compile time).  EXTRA = false: no residual / second gradient / channel scale (the encoder's BatchNorms):
// those operands are not even tested.  The row loop is software-pipelined: the loads of the next pair of rows are issued before the
// current pair is processed (ncu of the plain loop: 46 % of the warp samples sat on the first use of the freshly loaded row, 16 warps
// per SM cannot hide a DRAM round trip plus ~400 issue slots of arithmetic behind each other).
template <typename TX, typename TY, int ACT, bool EXTRA>
__global__ void __launch_bounds__(BN_THREADS, 2) bn_bwd_reduce_kernel(const TY* __restrict__ dy, const TY* __restrict__ dy2, const TX* __restrict__ x, const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   const TY* __restrict__ residual, const float* __restrict__ chan_scale, int rows_per_sample,
                                                                   int M, int C, TY* __restrict__ gbuf, float* sum_g, float* sum_gx, float* dbeta, float* dgamma, int rows_per_block) {
  pdl_sync();
  __shared__ float s1[BN_THREADS * 8];
  __shared__ float s2[BN_THREADS * 8];
  const int nvec_all = C >> 3;
  const int v0 = blockIdx.y * BN_THREADS;
  const int nvec = min(BN_THREADS, nvec_all - v0);
  const int rl_count = BN_THREADS / nvec, active = rl_count * nvec;
  const int t = threadIdx.x;
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (t < active) {
    const int cv = t % nvec, rl = t / nvec, c = (v0 + cv) * 8;
    float mu[8], is[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[c + j]; is[j] = invstd[c + j]; ga[j] = ACT ? gamma[c + j] : 0.f; be[j] = ACT ? beta[c + j] : 0.f; }
    typedef typename Vec8<TX>::raw_t RX;
    typedef typename Vec8<TY>::raw_t RY;
    struct Row { RX x; RY g, g2, res; };
    const bool has2 = EXTRA && dy2 != nullptr, hasr = EXTRA && residual != nullptr, hascs = EXTRA && chan_scale != nullptr;
    auto fetch = [&](int r, Row& w) {
      if (r < r1) {
        const long off = (long)r * C + c;
        w.x = Vec8<TX>::load_raw(x + off);
        w.g = Vec8<TY>::load_raw(dy + off);
        if (has2) w.g2 = Vec8<TY>::load_raw(dy2 + off);
        if (hasr) w.res = Vec8<TY>::load_raw(residual + off);
      }
    };
    auto finish = [&](int r, const Row& w) {
      if (r >= r1) return;
      const long off = (long)r * C + c;
      float xv[8], g[8], g2[8], res[8];
      Vec8<TX>::unpack(w.x, xv);
      Vec8<TY>::unpack(w.g, g);
      if (has2) Vec8<TY>::unpack(w.g2, g2);
      if (hasr) Vec8<TY>::unpack(w.res, res);
      const float* cs = hascs ? chan_scale + (long)(r / rows_per_sample) * C + c : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[j]) * is[j];
        float gj = g[j];
        if (has2) gj += g2[j];                                  // gradient fan-in (residual branch): summed here instead of in an axpy pass
        if (hascs) gj *= cs[j];
        if (ACT) {
          float z = fmaf(xh, ga[j], be[j]);
          if (hasr) z += res[j];
          gj *= act_bwd(z, ACT);
        }
        g[j] = gj;
        a1[j] += gj;
        a2[j] = fmaf(gj, xh, a2[j]);
      }
      Vec8<TY>::store(gbuf + off, g);
    };
    constexpr int NR = EXTRA ? 1 : 2;                           // rows per pipeline stage (the four-operand form has no registers for pairs)
    const int step = NR * rl_count;
    int r = r0 + rl;
    Row cur[NR], nxt[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) fetch(r + u * rl_count, cur[u]);
    for (; r < r1; r += step) {
#pragma unroll
      for (int u = 0; u < NR; ++u) fetch(r + step + u * rl_count, nxt[u]);      // next stage in flight while this one is processed
#pragma unroll
      for (int u = 0; u < NR; ++u) finish(r + u * rl_count, cur[u]);
#pragma unroll
      for (int u = 0; u < NR; ++u) cur[u] = nxt[u];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[t * 8 + j] = a1[j]; s2[t * 8 + j] = a2[j]; }
  __syncthreads();
  for (int i = t; i < nvec * 8; i += BN_THREADS) {
    const int v = i >> 3, j = i & 7;
    float d1 = 0.f, d2 = 0.f;
    for (int q = 0; q < rl_count; ++q) { d1 += s1[(q * nvec + v) * 8 + j]; d2 += s2[(q * nvec + v) * 8 + j]; }
    atomicAdd(sum_g + (v0 + v) * 8 + j, d1);
    atomicAdd(sum_gx + (v0 + v) * 8 + j, d2);
    if (dbeta) atomicAdd(dbeta + (v0 + v) * 8 + j, d1);        // the local parameter gradients are these very sums (before any cross-rank reduction)
    if (dgamma) atomicAdd(dgamma + (v0 + v) * 8 + j, d2);
  }
}

// dx = gamma*invstd * (g - mean(g) - xhat * mean(g*xhat)) = p*g + q*x + r with per-channel p, q, r in registers
// (same thread decomposition as bn_apply): two FMAs per element.
template <typename TX, typename TY, typename TD>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const TY* __restrict__ gbuf, const TX* __restrict__ x, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                  const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float inv_count,
                                                                  int training, int M, int C, TD* __restrict__ dx, int rows_per_block) {
  pdl_sync();
  const int nvec_all = C >> 3;
  const int v0 = blockIdx.y * BN_THREADS;
  const int nvec = min(BN_THREADS, nvec_all - v0);
  const int rl_count = BN_THREADS / nvec, active = rl_count * nvec;
  const int t = threadIdx.x;
  if (t >= active) return;
  const int cv = t % nvec, rl = t / nvec, c = (v0 + cv) * 8;
  float p[8], q[8], rr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float is = invstd[c + j], k = gamma[c + j] * is;
    p[j] = k;
    if (training) {
      const float a = sum_g[c + j] * inv_count, b = sum_gx[c + j] * inv_count * is;     // xhat * mean(g xhat) = (x - mu) * b
      q[j] = -k * b;
      rr[j] = k * (b * mean[c + j] - a);
    } else {
      q[j] = 0.f;
      rr[j] = 0.f;
    }
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  int r = r0 + rl;
  for (; r + rl_count < r1; r += 2 * rl_count) {          // two rows in flight per thread
    const long offa = (long)r * C + c, offb = (long)(r + rl_count) * C + c;
    float ga_[8], gb_[8], xa[8], xb[8];
    Vec8<TY>::load(gbuf + offa, ga_);
    Vec8<TX>::load(x + offa, xa);
    Vec8<TY>::load(gbuf + offb, gb_);
    Vec8<TX>::load(x + offb, xb);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga_[j] = fmaf(p[j], ga_[j], fmaf(q[j], xa[j], rr[j])); gb_[j] = fmaf(p[j], gb_[j], fmaf(q[j], xb[j], rr[j])); }
    Vec8<TD>::store(dx + offa, ga_);
    Vec8<TD>::store(dx + offb, gb_);
  }
  if (r < r1) {
    const long off = (long)r * C + c;
    float g[8], xv[8];
    Vec8<TY>::load(gbuf + off, g);
    Vec8<TX>::load(x + off, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(p[j], g[j], fmaf(q[j], xv[j], rr[j]));
    Vec8<TD>::store(dx + off, g);
  }
}

// rows per CTA of the streaming BN kernels: ~8 resident CTAs per SM, each thread walking >= 4 rows
inline int bn_stream_rows(int M, int C) {
  const int nvec = C / 8 < BN_THREADS ? C / 8 : BN_THREADS;
  const int lanes = BN_THREADS / nvec;
  int rpb = dfb_cdiv(M, 148 * 8);
  if (rpb < 4 * lanes) rpb = 4 * lanes;
  return rpb;
}


}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

static int ln_fwd_launch(const float* x, const float* gamma, const float* beta, float eps, int M, int C, void* y, int y_dtype, float* mean, float* rstd,
                         const LnResidual& rp, cudaStream_t st) {
  const int grid = min(dfb_cdiv(M, 8), 148 * 8);
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(rp.y) |
                       reinterpret_cast<uintptr_t>(rp.x_out) | reinterpret_cast<uintptr_t>(rp.ls);
  if (C % 8 == 0 && (al & 15) == 0 && (rp.y == nullptr || rp.ldy % 8 == 0)) {
    DFB_DISPATCH_DTYPE(y_dtype, T, { LN_DISPATCH_VEC(C, { dfb_launch(ln_fwd_vec_kernel<T, L, VPL>, grid, 256, 0, st, x, gamma, beta, eps, M, C, (T*)y, mean, rstd, rp); }); });
    return dfb_check_launch("layernorm_fwd_vec");
  }
  DFB_DISPATCH_DTYPE(y_dtype, T, { LN_DISPATCH_NPL(C, { dfb_launch(ln_fwd_kernel<T, NPL>, grid, 256, 0, st, x, gamma, beta, eps, M, C, (T*)y, mean, rstd, rp); }); });
  return dfb_check_launch("layernorm_fwd");
}

extern "C" int dfb200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int M, int C, void* y, int y_dtype, float* mean,
                                    float* rstd, void* stream) {
  DFB_REQUIRE(C >= 1 && C <= 32 * LN_MAX_PER_LANE, "layernorm: C=%d out of range", C);
  if (M <= 0) return DFB_OK;
  const LnResidual none = {nullptr, 0, nullptr, nullptr, 1, nullptr};
  return ln_fwd_launch(x, gamma, beta, eps, M, C, y, y_dtype, mean, rstd, none, ST);
}

// x_out = res + scale_b * ls * branch (fp32), y = LayerNorm(x_out) -- one pass (branch and y share `dtype`)
extern "C" int dfb200_scale_residual_layernorm_fwd(const float* res, const void* branch, long ld_branch, int dtype, const float* ls, const float* scale_b,
                                                   int rows_per_sample, int M, int C, float* x_out, const float* gamma, const float* beta, float eps,
                                                   void* y, float* mean, float* rstd, void* stream) {
  DFB_REQUIRE(C >= 1 && C <= 32 * LN_MAX_PER_LANE, "layernorm: C=%d out of range", C);
  DFB_REQUIRE(branch && ls && x_out && rows_per_sample > 0, "scale_residual_layernorm_fwd: missing operand");
  if (M <= 0) return DFB_OK;
  const LnResidual rp = {branch, ld_branch, ls, scale_b, rows_per_sample, x_out};
  return ln_fwd_launch(res, gamma, beta, eps, M, C, y, dtype, mean, rstd, rp, ST);
}

#define LN_BWD_ARGS grid, 256, 16 * C * sizeof(float), ST, (const T*)dy, (const T*)dy2, x, gamma, mean, rstd, M, C, dx_in, dx, dgamma, dbeta
extern "C" int dfb200_layernorm_bwd(const void* dy, const void* dy2, int dy_dtype, const float* x, const float* gamma, const float* mean, const float* rstd, int M, int C,
                                    const float* dx_in, float* dx, float* dgamma, float* dbeta, void* stream) {
  DFB_REQUIRE(C >= 1 && C <= 32 * LN_MAX_PER_LANE, "layernorm: C=%d out of range", C);
  if (M <= 0) return DFB_OK;
  // every CTA ends with 2*C global atomics (dgamma, dbeta): >= 32 rows per CTA keeps that tail below the streaming work
  // at the small-M stages (M = 9600: 300 CTAs instead of 1184 contending for the same 2*C addresses)
  const int grid = min(dfb_cdiv(M, 32), 148 * 8);
  if (C % 8 == 0 && C <= 704 /* 8 per-warp slabs of 2*C floats in 48 KB */ && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dy2) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dx_in)) & 15) == 0) {
    DFB_DISPATCH_DTYPE(dy_dtype, T, {
      const int nvec = C >> 3;
      if (nvec <= 8) dfb_launch(ln_bwd_vec_pf_kernel<T, 8>, LN_BWD_ARGS);              // one vector per lane: software-pipelined rows
      else if (nvec <= 16) dfb_launch(ln_bwd_vec_pf_kernel<T, 16>, LN_BWD_ARGS);
      else if (nvec <= 32) dfb_launch(ln_bwd_vec_pf_kernel<T, 32>, LN_BWD_ARGS);
      else if (nvec <= 64) dfb_launch(ln_bwd_vec_kernel<T, 32, 2>, LN_BWD_ARGS);
      else if (nvec <= 96) dfb_launch(ln_bwd_vec_kernel<T, 32, 3>, LN_BWD_ARGS);
      else dfb_launch(ln_bwd_vec_kernel<T, 32, 4>, LN_BWD_ARGS);

    });
    return dfb_check_launch("layernorm_bwd_vec");
  }
  DFB_DISPATCH_DTYPE(dy_dtype, T, {
    LN_DISPATCH_NPL(C, { dfb_launch(ln_bwd_kernel<T, NPL>, grid, 256, 2 * C * sizeof(float), ST, (const T*)dy, (const T*)dy2, x, gamma, mean, rstd, M, C, dx_in, dx, dgamma, dbeta); });
  });
  return dfb_check_launch("layernorm_bwd");
}

// Inference-time BatchNorm folding for conv -> BN pairs: the freshly packed GEMM weight [rows, ld] (compute
// dtype) is scaled row-wise in place by s = gamma * rsqrt(running_var + eps) and bias_out = (conv_bias - running_mean) * s + beta,
// so that BN(conv(x)) = x * W' + bias' and the BN kernels drop out of the forward pass.  The fp32 parameters stay untouched.
template <typename T>
__global__ void bn_fold_kernel(T* __restrict__ w, int rows, int cols, long ld, const float* __restrict__ conv_bias, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ bias_out) {
  pdl_sync();
  const long n = (long)rows * cols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float sc = gamma[r] * rsqrtf(rv[r] + eps);
    w[(long)r * ld + c] = from_f<T>(to_f(w[(long)r * ld + c]) * sc);
    if (c == 0) bias_out[r] = ((conv_bias ? conv_bias[r] : 0.f) - rm[r]) * sc + beta[r];
  }
}

extern "C" int dfb200_bn_fold(void* w_packed, int dtype, int rows, int cols, long ld, const float* conv_bias, const float* running_mean,
                              const float* running_var, float eps, const float* gamma, const float* beta, float* bias_out, void* stream) {
  DFB_REQUIRE(rows > 0 && cols > 0 && ld >= cols && bias_out, "bn_fold: bad arguments");
  const long n = (long)rows * cols;
  const int grid = (int)min((n + 255) / 256, 148L * 8);
  DFB_DISPATCH_DTYPE(dtype, T, { dfb_launch(bn_fold_kernel<T>, grid, 256, 0, ST, (T*)w_packed, rows, cols, ld, conv_bias, running_mean, running_var, eps, gamma, beta, bias_out); });
  return dfb_check_launch("bn_fold");
}

extern "C" int dfb200_bn_stats(const void* x, int dtype, int M, int C, double* sum, double* sumsq, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "bn_stats: C %% 8 != 0 (C=%d)", C);
  cudaMemsetAsync(sum, 0, sizeof(double) * C, ST);
  cudaMemsetAsync(sumsq, 0, sizeof(double) * C, ST);
  if (M <= 0) return DFB_OK;
  int rpb = dfb_cdiv(M, 148 * 4);
  if (rpb < 32) rpb = 32;
  if (rpb > 4096) rpb = 4096;          // bound the length of float partial sums
  dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(C / 8, BN_THREADS));
  DFB_DISPATCH_DTYPE(dtype, T, { dfb_launch(bn_stats_kernel<T>, grid, BN_THREADS, 0, ST, (const T*)x, M, C, sum, sumsq, rpb); });
  return dfb_check_launch("bn_stats");
}

extern "C" int dfb200_bn_finalize(const double* sum, const double* sumsq, double count, float eps, float momentum, int C, float* mean, float* invstd,
                                  float* running_mean, float* running_var, void* stream) {
  dfb_launch(bn_finalize_kernel, dfb_cdiv(C, 128), 128, 0, ST, sum, sumsq, count, eps, momentum, C, mean, invstd, running_mean, running_var);
  return dfb_check_launch("bn_finalize");
}
extern "C" int dfb200_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean, float* invstd, void* stream) {
  dfb_launch(bn_eval_stats_kernel, dfb_cdiv(C, 128), 128, 0, ST, running_mean, running_var, eps, C, mean, invstd);
  return dfb_check_launch("bn_eval_stats");
}

extern "C" int dfb200_bn_apply(const void* x, int x_dtype, const float* mean, const float* invstd, const float* gamma, const float* beta,
                               const void* residual, int act, const float* chan_scale, int rows_per_sample, int M, int C, void* y, int y_dtype, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "bn_apply: C %% 8 != 0 (C=%d)", C);
  if (M <= 0) return DFB_OK;
  const int rpb = bn_stream_rows(M, C);
  dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(C / 8, BN_THREADS));
#define L(TX, TY) dfb_launch(bn_apply_kernel<TX, TY>, grid, BN_THREADS, 0, ST, (const TX*)x, mean, invstd, gamma, beta, (const TY*)residual, act, chan_scale, rows_per_sample, M, C, (TY*)y, rpb)
  if (x_dtype == 0 && y_dtype == 0) L(float, float);
  else if (x_dtype == 0 && y_dtype == 1) L(float, bf16);
  else if (x_dtype == 1 && y_dtype == 1) L(bf16, bf16);
  else if (x_dtype == 1 && y_dtype == 0) L(bf16, float);
  else { dfb_set_error("bn_apply: bad dtypes"); return DFB_ERR_ARG; }
#undef L
  return dfb_check_launch("bn_apply");
}

extern "C" int dfb200_bn_bwd_reduce(const void* dy, const void* dy2, int y_dtype, const void* x, int x_dtype, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const void* residual, int act, const float* chan_scale, int rows_per_sample,
                                    int M, int C, void* gbuf, float* sum_g, float* sum_gx, float* dbeta, float* dgamma, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "bn_bwd_reduce: C %% 8 != 0 (C=%d)", C);
  // the kernel keeps ~100 registers per thread (per-channel constants + two rows in flight): two CTAs per SM, so the grid is ONE wave
  // of 2 x 148 CTAs (a second, partial wave of this latency-sensitive kernel cost as much as the first)
  const int ygrid = dfb_cdiv(C / 8, BN_THREADS);
  int rpb = dfb_cdiv(M, max(1, (148 * 2) / ygrid));
  if (rpb < 32) rpb = 32;
  dim3 grid(dfb_cdiv(M, rpb), ygrid);
  DFB_REQUIRE(act >= 0 && act <= 2, "bn_bwd_reduce: bad activation %d", act);
  const bool extra = dy2 != nullptr || residual != nullptr || chan_scale != nullptr;
#define L3(TX, TY, A, E) dfb_launch(bn_bwd_reduce_kernel<TX, TY, A, E>, grid, BN_THREADS, 0, ST, (const TY*)dy, (const TY*)dy2, (const TX*)x, mean, invstd, gamma, beta, (const TY*)residual, chan_scale, rows_per_sample, M, C, (TY*)gbuf, sum_g, sum_gx, dbeta, dgamma, rpb)
#define L2(TX, TY, A) do { if (extra) L3(TX, TY, A, true); else L3(TX, TY, A, false); } while (0)
#define L(TX, TY) do { if (act == 0) L2(TX, TY, 0); else if (act == 1) L2(TX, TY, 1); else L2(TX, TY, 2); } while (0)
  if (x_dtype == 0 && y_dtype == 0) L(float, float);
  else if (x_dtype == 0 && y_dtype == 1) L(float, bf16);
  else if (x_dtype == 1 && y_dtype == 1) L(bf16, bf16);
  else if (x_dtype == 1 && y_dtype == 0) L(bf16, float);
  else { dfb_set_error("bn_bwd_reduce: bad dtypes"); return DFB_ERR_ARG; }
#undef L
#undef L2
#undef L3
  return dfb_check_launch("bn_bwd_reduce");
}

extern "C" int dfb200_bn_bwd_apply(const void* gbuf, int y_dtype, const void* x, int x_dtype, const float* mean, const float* invstd, const float* gamma,
                                   const float* sum_g, const float* sum_gx, float count, int training, int M, int C, void* dx, int dx_dtype, void* stream) {
  DFB_REQUIRE(C % 8 == 0, "bn_bwd_apply: C %% 8 != 0 (C=%d)", C);
  if (M <= 0) return DFB_OK;
  const float inv = 1.f / count;
  const int rpb = bn_stream_rows(M, C);
  dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(C / 8, BN_THREADS));
#define L(TX, TY, TD) dfb_launch(bn_bwd_apply_kernel<TX, TY, TD>, grid, BN_THREADS, 0, ST, (const TY*)gbuf, (const TX*)x, mean, invstd, gamma, sum_g, sum_gx, inv, training, M, C, (TD*)dx, rpb)
  const int key = x_dtype * 4 + y_dtype * 2 + dx_dtype;
  switch (key) {
    case 0: L(float, float, float); break;
    case 1: L(float, float, bf16); break;
    case 2: L(float, bf16, float); break;
    case 3: L(float, bf16, bf16); break;
    case 4: L(bf16, float, float); break;
    case 5: L(bf16, float, bf16); break;
    case 6: L(bf16, bf16, float); break;
    case 7: L(bf16, bf16, bf16); break;
    default: dfb_set_error("bn_bwd_apply: bad dtypes"); return DFB_ERR_ARG;
  }
#undef L
  return dfb_check_launch("bn_bwd_apply");
}
