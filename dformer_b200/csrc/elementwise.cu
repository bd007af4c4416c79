// HBM-bound element-wise / column-reduction kernels: parameter packing, gating, layer-scale residual,
// NMF multiplicative updates, softmax rows, casts, AdamW.  All are single-pass, 16-byte vectorised
// where the shape allows (C % 8 == 0), grid-stride, fp32 math.
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int EW_THREADS = 256;
inline int ew_grid(long n, int per_thread = 1) {
  long b = (n + (long)EW_THREADS * per_thread - 1) / ((long)EW_THREADS * per_thread);
  if (b < 1) b = 1;
  const long cap = 148L * 16;
  return (int)(b > cap ? cap : b);
}

// ------------------------------------------------------------------ column reductions
// Flat mapping: active threads A = (T / nvec) * nvec so that thread t always owns channel-vector t % nvec
// while consecutive threads touch consecutive 16/32-byte vectors (full coalescing across row boundaries).
template <typename T, int NACC, typename F>
__device__ __forceinline__ void colreduce_body(int M, int C, int rows_per_block, F f, float* const* outs, float* smem) {
  const int nvec_all = C >> 3;
  const int v0 = blockIdx.y * EW_THREADS;                         // channel-vector chunk
  const int nvec = min(EW_THREADS, nvec_all - v0);
  const int rl_count = EW_THREADS / nvec;
  const int active = rl_count * nvec;
  const int t = threadIdx.x;
  float acc[NACC][8];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
  const int cv = (t < active) ? (t % nvec) : 0;
  const int rl = t / nvec;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (t < active) {
    for (int r = r0 + rl; r < r1; r += rl_count) f(r, (v0 + cv) * 8, acc);
  }
  // reduce over rl through shared memory
  float* s = smem;   // [NACC][EW_THREADS][8]
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) s[(a * EW_THREADS + t) * 8 + j] = acc[a][j];
  __syncthreads();
  for (int i = t; i < NACC * nvec * 8; i += EW_THREADS) {
    const int a = i / (nvec * 8), rem = i % (nvec * 8), v = rem >> 3, j = rem & 7;
    float sum = 0.f;
    for (int q = 0; q < rl_count; ++q) sum += s[(a * EW_THREADS + q * nvec + v) * 8 + j];
    atomicAdd(outs[a] + (v0 + v) * 8 + j, sum);
  }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) colsum_kernel(const T* __restrict__ X, long ldx, int M, int N, float* out, int rows_per_block) {
  pdl_sync();
  extern __shared__ float smem[];
  float* outs[1] = {out};
  colreduce_body<T, 1>(M, N, rows_per_block,
                       [&](int r, int c, float(*acc)[8]) {
                         float v[8];
                         Vec8<T>::load(X + (long)r * ldx + c, v);
#pragma unroll
                         for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
                       },
                       outs, smem);
}

template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ X, long ldx, int M, int N, float* out, int rows_per_block) {
  pdl_sync();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += to_f(X[(long)r * ldx + n]);
  atomicAdd(out + n, s);
}

inline int pick_rows_per_block(int M) {
  int rpb = dfb_cdiv(M, 148 * 4);
  if (rpb < 32) rpb = 32;
  return rpb;
}

// ------------------------------------------------------------------ parameter packing
template <typename T>
__global__ void pack_params_kernel(const dfb200_pack_entry* __restrict__ table, int max_elems) {
  pdl_sync();
  const dfb200_pack_entry e = table[blockIdx.y];
  T* dst = reinterpret_cast<T*>(e.dst);
  if (e.kind == 0 && e.cols == e.dst_ld && (((long)e.rows * e.cols) & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(e.src) | reinterpret_cast<uintptr_t>(e.dst)) & 15) == 0) {
    // unpadded rows: a flat cast, eight elements per thread and iteration (two 16-byte loads, one or two 16-byte stores)
    const unsigned nv = (unsigned)(((long)e.rows * e.cols) >> 3);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
      float v[8];
      Vec8<float>::load(e.src + (size_t)i * 8, v);
      Vec8<T>::store(dst + (size_t)i * 8, v);
    }
  } else if (e.kind == 0) {
    const long n = (long)e.rows * e.cols;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
      const int r = (int)(i / e.cols), c = (int)(i % e.cols);
      dst[(long)r * e.dst_ld + c] = from_f<T>(e.src[i]);
    }
  } else {  // conv weight [Cout, Cin, 3, 3] -> [Cout, (ky*3+kx)*Cin + ci]
    const int cin = e.cols;
    const long n = (long)e.rows * cin * 9;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
      const int co = (int)(i / (cin * 9)), rem = (int)(i % (cin * 9));
      const int tap = rem / cin, ci = rem % cin;
      dst[(long)co * e.dst_ld + rem] = from_f<T>(e.src[((long)co * cin + ci) * 9 + tap]);
    }
  }
}

__global__ void unpack_conv_grad_kernel(const float* __restrict__ dWp, int ld, int Cout, int Cin, float* __restrict__ dW) {
  pdl_sync();
  const long n = (long)Cout * Cin * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i / (Cin * 9)), rem = (int)(i % (Cin * 9));
    const int ci = rem / 9, tap = rem % 9;
    dW[i] = dWp[(long)co * ld + tap * Cin + ci];
  }
}

// ------------------------------------------------------------------ gating multiply
template <typename T>
__global__ void mul_fwd_kernel(const T* __restrict__ a, long lda, const T* __restrict__ b, long ldb, T* __restrict__ o, long ldo, int M, int nvec) {
  pdl_sync();
  const long n = (long)M * nvec;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nvec; const int c = (int)(i % nvec) * 8;
    float x[8], y[8];
    Vec8<T>::load(a + r * lda + c, x);
    Vec8<T>::load(b + r * ldb + c, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= y[j];
    Vec8<T>::store(o + r * ldo + c, x);
  }
}
template <typename T>
__global__ void mul_bwd_kernel(const T* __restrict__ dout, long ldo, const T* __restrict__ a, long lda, const T* __restrict__ b, long ldb,
                               T* __restrict__ da, long ldda, T* __restrict__ db, long lddb, int M, int nvec) {
  pdl_sync();
  const long n = (long)M * nvec;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nvec; const int c = (int)(i % nvec) * 8;
    float g[8], x[8], y[8], ga[8], gb[8];
    Vec8<T>::load(dout + r * ldo + c, g);
    Vec8<T>::load(a + r * lda + c, x);
    Vec8<T>::load(b + r * ldb + c, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = g[j] * y[j]; gb[j] = g[j] * x[j]; }
    Vec8<T>::store(da + r * ldda + c, ga);
    Vec8<T>::store(db + r * lddb + c, gb);
  }
}

// ------------------------------------------------------------------ stand-alone activation
template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ in, long ldi, T* __restrict__ out, long ldo, int act, int M, int nvec) {
  pdl_sync();
  const long n = (long)M * nvec;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nvec; const int c = (int)(i % nvec) * 8;
    float v[8];
    Vec8<T>::load(in + r * ldi + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act == 1 ? gelu_f(v[j]) : fmaxf(v[j], 0.f);
    Vec8<T>::store(out + r * ldo + c, v);
  }
}
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dout, long lddo, const T* __restrict__ dout2, long lddo2, const T* __restrict__ z, long ldz,
                               T* __restrict__ din, long lddi, int act, int M, int nvec) {
  pdl_sync();
  const long n = (long)M * nvec;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nvec; const int c = (int)(i % nvec) * 8;
    float g[8], zz[8];
    Vec8<T>::load(dout + r * lddo + c, g);
    if (dout2) {                                     // fused gradient fan-in (sum rounded to T like the separate axpy pass did)
      float g2[8];
      Vec8<T>::load(dout2 + r * lddo2 + c, g2);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = to_f(from_f<T>(g[j] + g2[j]));
    }
    Vec8<T>::load(z + r * ldz + c, zz);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act == 1 ? gelu_grad_f(zz[j]) : (zz[j] > 0.f ? 1.f : 0.f);
    Vec8<T>::store(din + r * lddi + c, g);
  }
}

// Column-summing forms of the two kernels above: the element-wise gradient and the bias gradient (column sums of that gradient
// as the weight-gradient GEMM will read it, i.e. after rounding to T) of the Linear layers behind them leave in one pass.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) mul_bwd_colsum_kernel(const T* __restrict__ dout, long ldo, const T* __restrict__ a, long lda,
                                                                    const T* __restrict__ b, long ldb, T* __restrict__ da, long ldda, T* __restrict__ db,
                                                                    long lddb, int M, int N, float* da_colsum, float* db_colsum, int rows_per_block) {
  pdl_sync();
  extern __shared__ float smem[];
  float* outs[2] = {da_colsum, db_colsum};
  colreduce_body<T, 2>(M, N, rows_per_block,
                       [&](int r, int c, float(*acc)[8]) {
                         float g[8], x[8], y[8], ga[8], gb[8];
                         Vec8<T>::load(dout + (long)r * ldo + c, g);
                         Vec8<T>::load(a + (long)r * lda + c, x);
                         Vec8<T>::load(b + (long)r * ldb + c, y);
#pragma unroll
                         for (int j = 0; j < 8; ++j) {
                           ga[j] = g[j] * y[j];
                           gb[j] = g[j] * x[j];
                           acc[0][j] += to_f(from_f<T>(ga[j]));
                           acc[1][j] += to_f(from_f<T>(gb[j]));
                         }
                         Vec8<T>::store(da + (long)r * ldda + c, ga);
                         Vec8<T>::store(db + (long)r * lddb + c, gb);
                       },
                       outs, smem);
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) act_bwd_colsum_kernel(const T* __restrict__ dout, long lddo, const T* __restrict__ dout2, long lddo2,
                                                                    const T* __restrict__ z, long ldz, T* __restrict__ din, long lddi, int act, int M, int N,
                                                                    float* colsum, int rows_per_block) {
  pdl_sync();
  extern __shared__ float smem[];
  float* outs[1] = {colsum};
  colreduce_body<T, 1>(M, N, rows_per_block,
                       [&](int r, int c, float(*acc)[8]) {
                         float g[8], zz[8];
                         Vec8<T>::load(dout + (long)r * lddo + c, g);
                         if (dout2) {
                           float g2[8];
                           Vec8<T>::load(dout2 + (long)r * lddo2 + c, g2);
#pragma unroll
                           for (int j = 0; j < 8; ++j) g[j] = to_f(from_f<T>(g[j] + g2[j]));
                         }
                         Vec8<T>::load(z + (long)r * ldz + c, zz);
#pragma unroll
                         for (int j = 0; j < 8; ++j) {
                           g[j] *= act == 1 ? gelu_grad_f(zz[j]) : (zz[j] > 0.f ? 1.f : 0.f);
                           acc[0][j] += to_f(from_f<T>(g[j]));
                         }
                         Vec8<T>::store(din + (long)r * lddi + c, g);
                       },
                       outs, smem);
}

// ------------------------------------------------------------------ layer-scale residual
template <typename T>
__global__ void scale_residual_fwd_kernel(const float* __restrict__ res, const T* __restrict__ y, long ldy, const float* __restrict__ ls,
                                          const float* __restrict__ scale_b, int M, int nvec, int rows_per_sample, float* __restrict__ out) {
  pdl_sync();
  const long n = (long)M * nvec;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nvec; const int c = (int)(i % nvec) * 8;
    const float sb = scale_b ? scale_b[r / rows_per_sample] : 1.f;
    float rv[8], yv[8], lv[8];
    Vec8<float>::load(res + i * 8, rv);
    Vec8<T>::load(y + r * ldy + c, yv);
    Vec8<float>::load(ls + c, lv);
#pragma unroll
    for (int j = 0; j < 8; ++j) rv[j] = fmaf(sb * lv[j], yv[j], rv[j]);
    Vec8<float>::store(out + i * 8, rv);
  }
}
template <typename T, int NOUT>
__global__ void __launch_bounds__(EW_THREADS) scale_residual_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ y, long ldy, const float* __restrict__ ls,
                                                                        const float* __restrict__ scale_b, int M, int C, int rows_per_sample,
                                                                        T* __restrict__ dy, long lddy, float* dls, float* dy_colsum, int rows_per_block) {
  pdl_sync();
  extern __shared__ float smem[];
  float* outs[2] = {dls, dy_colsum};      // NOUT == 2: also the column sum of dy (= bias gradient of the Linear that produced y)
  colreduce_body<T, NOUT>(M, C, rows_per_block,
                          [&](int r, int c, float(*acc)[8]) {
                            const float sb = scale_b ? scale_b[r / rows_per_sample] : 1.f;
                            float g[8], yv[8], lv[8], o[8];
                            Vec8<float>::load(dout + (long)r * C + c, g);
                            Vec8<T>::load(y + (long)r * ldy + c, yv);
                            Vec8<float>::load(ls + c, lv);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                              o[j] = g[j] * lv[j] * sb;
                              acc[0][j] += g[j] * yv[j] * sb;
                              if (NOUT == 2) acc[1][j] += to_f(from_f<T>(o[j]));      // sums the values the weight-gradient GEMM will read
                            }
                            Vec8<T>::store(dy + (long)r * lddy + c, o);
                          },
                          outs, smem);
}

// ------------------------------------------------------------------ NMF helpers
__global__ void __launch_bounds__(256) normalize_cols_kernel(const float* __restrict__ in, int D, int R, float* __restrict__ out, float* __restrict__ norms) {
  pdl_sync();
  // one block per image b: out = in / max(||col||_2, 1e-12) per column of the [D, R] matrix.  Thread t owns column t % R and every
  // (256 / R)-th row (coalesced along R); the per-column partial sums meet in shared memory.
  extern __shared__ float nsm[];                       // [blockDim.x] partials, then [R] inverse norms
  const int b = blockIdx.x;
  const float* src = in + (long)b * D * R;
  float* dst = out + (long)b * D * R;
  const int t = threadIdx.x, nth = blockDim.x;
  const int rl = nth / R;                              // row lanes (>= 1: the launcher guarantees R <= blockDim.x)
  const int col = t % R, lane_r = t / R;
  float s = 0.f;
  if (lane_r < rl)
    for (int d = lane_r; d < D; d += rl) { const float v = src[(long)d * R + col]; s = fmaf(v, v, s); }
  nsm[t] = s;
  __syncthreads();
  if (t < R) {
    float tot = 0.f;
    for (int q = 0; q < rl; ++q) tot += nsm[q * R + t];
    const float nrm = fmaxf(sqrtf(tot), 1e-12f);
    if (norms) norms[b * R + t] = nrm;
    nsm[nth + t] = 1.f / nrm;
  }
  __syncthreads();
  if (lane_r < rl) {
    const float inv = nsm[nth + col];
    for (int d = lane_r; d < D; d += rl) dst[(long)d * R + col] = src[(long)d * R + col] * inv;
  }
}

__global__ void softmax_rows_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  pdl_sync();
  // one warp per row
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = in + (long)row * cols;
  float* y = out + (long)row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __expf(x[c] - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int c = lane; c < cols; c += 32) y[c] = __expf(x[c] - m) * inv;
}
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, int rows, int cols, float* __restrict__ din) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* g = dout + (long)row * cols;
  const float* p = out + (long)row * cols;
  float* d = din + (long)row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(g[c], p[c], s);
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) d[c] = p[c] * (g[c] - s);
}

// Multiplicative update of NMF2D (ham_head.py:126,133) with an optional second copy of the result in the compute dtype
// (bf16 operand of the next batched GEMM), written in the same pass.
template <typename TL>
__global__ void mu_update_kernel(const float* __restrict__ a, const float* __restrict__ num, const float* __restrict__ den, float eps, long n,
                                 float* __restrict__ out, TL* __restrict__ out_lo) {
  pdl_sync();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = a[i] * num[i] / (den[i] + eps);
    out[i] = v;
    if (out_lo) out_lo[i] = from_f<TL>(v);
  }
}
// Backward.  dnum / dden are only ever consumed as GEMM operands, so they are written straight in the compute dtype: dnum with a
// row length `cols` and leading dimension `ld_dnum` (it lands in a column slice of a K-concatenated operand), dden contiguous.
template <typename TL>
__global__ void mu_update_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ num, const float* __restrict__ den,
                                     float eps, long n, float* __restrict__ da, int acc_da, TL* __restrict__ dnum, long ld_dnum, int cols,
                                     TL* __restrict__ dden) {
  pdl_sync();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float inv = 1.f / (den[i] + eps), g = dout[i];
    const float ga = g * num[i] * inv;
    da[i] = acc_da ? da[i] + ga : ga;
    const long r = i / cols;
    dnum[r * ld_dnum + (i - r * cols)] = from_f<TL>(g * a[i] * inv);
    dden[i] = from_f<TL>(-g * a[i] * num[i] * inv * inv);
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n) {
  pdl_sync();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = from_f<TO>(to_f(in[i]));
}
// strided 2-D copy/convert: out[r, c] = in[r, c] with independent leading dimensions (writes into column slices of a
// concatenated operand buffer)
template <typename TI, typename TO>
__global__ void cast2d_kernel(const TI* __restrict__ in, long ld_in, TO* __restrict__ out, long ld_out, long rows, int cols) {
  pdl_sync();
  const long n = rows * cols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * ld_out + c] = from_f<TO>(to_f(in[r * ld_in + c]));
  }
}
// out[b, i, j] = in[b, i, j] + in[b, j, i] for a batch of small square matrices (NMF backward: the gradient of BtB = B^T B enters
// both factors, so bases @ (G + G^T) replaces two products)
template <typename TI, typename TO>
__global__ void sym_cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int batch, int R) {
  pdl_sync();
  const long n = (long)batch * R * R;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long b = i / (R * R);
    const int rc = (int)(i - b * R * R), r = rc / R, c = rc - r * R;
    out[i] = from_f<TO>(to_f(in[i]) + to_f(in[b * R * R + (long)c * R + r]));
  }
}
template <typename TI, typename TO>
__global__ void axpy_kernel(const TI* __restrict__ x, float alpha, TO* __restrict__ y, long n) {
  pdl_sync();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(fmaf(alpha, to_f(x[i]), to_f(y[i])));
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n, float lr_s, float b1,
                             float b2, float eps, float wd_s, float c1, float c2, float gs, const float* __restrict__ wd_arr,
                             const float* __restrict__ lr_arr, const float* __restrict__ dyn) {
  pdl_sync();
  if (dyn) {                                                // schedule state lives on the device (CUDA-graph replays)
    lr_s = dyn[0];
    c1 = 1.f - powf(b1, dyn[1]);
    c2 = 1.f - powf(b2, dyn[1]);
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float lrm = lr_arr ? lr_arr[i] : 1.f;
    if (lrm == 0.f) continue;                               // frozen element (parameter outside every optimizer group)
    const float wd = wd_arr ? wd_arr[i] : wd_s;
    const float lr = lr_s * lrm;
    const float gi = g[i] * gs;
    float pi = p[i];
    pi *= (1.f - lr * wd);                                  // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(c2) + eps;
    p[i] = pi - (lr / c1) * mi / denom;
  }
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_colsum(const void* X, int dtype, long ldx, int M, int N, float* out, int accumulate, void* stream) {
  if (M <= 0 || N <= 0) return DFB_OK;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, ST);
  const int rpb = pick_rows_per_block(M);
  const bool vec = (N % 8 == 0) && (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(X) & 31) == 0);
  DFB_DISPATCH_DTYPE(dtype, T, {
    if (vec) {
      dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(N / 8, EW_THREADS));
      dfb_launch(colsum_kernel<T>, grid, EW_THREADS, EW_THREADS * 8 * sizeof(float), ST, (const T*)X, ldx, M, N, out, rpb);
    } else {
      dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(N, 128));
      dfb_launch(colsum_scalar_kernel<T>, grid, 128, 0, ST, (const T*)X, ldx, M, N, out, rpb);
    }
  });
  return dfb_check_launch("colsum");
}

extern "C" int dfb200_pack_params(const dfb200_pack_entry* table_dev, int n_entries, int max_elems, int dst_dtype, void* stream) {
  if (n_entries <= 0) return DFB_OK;
  DFB_REQUIRE(n_entries <= 65535, "pack_params: too many entries (%d)", n_entries);
  int gx = dfb_cdiv(max_elems, EW_THREADS * 16);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, n_entries);
  DFB_DISPATCH_DTYPE(dst_dtype, T, { dfb_launch(pack_params_kernel<T>, grid, EW_THREADS, 0, ST, table_dev, max_elems); });
  return dfb_check_launch("pack_params");
}

extern "C" int dfb200_unpack_conv_grad(const float* dWp, int ld, int Cout, int Cin, float* dW, void* stream) {
  dfb_launch(unpack_conv_grad_kernel, ew_grid((long)Cout * Cin * 9), EW_THREADS, 0, ST, dWp, ld, Cout, Cin, dW);
  return dfb_check_launch("unpack_conv_grad");
}

extern "C" int dfb200_mul_fwd(const void* a, long lda, const void* b, long ldb, void* out, long ldo, int dtype, int M, int N, void* stream) {
  DFB_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "mul: N and leading dims must be multiples of 8");
  DFB_DISPATCH_DTYPE(dtype, T, { dfb_launch(mul_fwd_kernel<T>, ew_grid((long)M * N / 8), EW_THREADS, 0, ST, (const T*)a, lda, (const T*)b, ldb, (T*)out, ldo, M, N / 8); });
  return dfb_check_launch("mul_fwd");
}
extern "C" int dfb200_mul_bwd(const void* dout, long ldo, const void* a, long lda, const void* b, long ldb, void* da, long ldda, void* db, long lddb,
                              int dtype, int M, int N, float* da_colsum, float* db_colsum, void* stream) {
  DFB_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && ldda % 8 == 0 && lddb % 8 == 0, "mul_bwd: alignment");
  DFB_REQUIRE((da_colsum == nullptr) == (db_colsum == nullptr), "mul_bwd: give both column-sum outputs or neither");
  if (da_colsum) {
    const int rpb = pick_rows_per_block(M);
    dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(N / 8, EW_THREADS));
    DFB_DISPATCH_DTYPE(dtype, T, {
      dfb_launch(mul_bwd_colsum_kernel<T>, grid, EW_THREADS, 2 * EW_THREADS * 8 * sizeof(float), ST, (const T*)dout, ldo, (const T*)a, lda, (const T*)b, ldb,
                 (T*)da, ldda, (T*)db, lddb, M, N, da_colsum, db_colsum, rpb);
    });
    return dfb_check_launch("mul_bwd_colsum");
  }
  DFB_DISPATCH_DTYPE(dtype, T, {
    dfb_launch(mul_bwd_kernel<T>, ew_grid((long)M * N / 8), EW_THREADS, 0, ST, (const T*)dout, ldo, (const T*)a, lda, (const T*)b, ldb, (T*)da, ldda, (T*)db, lddb, M, N / 8);
  });
  return dfb_check_launch("mul_bwd");
}

extern "C" int dfb200_act_fwd(const void* in, long ldi, void* out, long ldo, int dtype, int act, int M, int N, void* stream) {
  DFB_REQUIRE(N % 8 == 0 && ldi % 8 == 0 && ldo % 8 == 0 && (act == 1 || act == 2), "act_fwd: bad arguments");
  DFB_DISPATCH_DTYPE(dtype, T, { dfb_launch(act_fwd_kernel<T>, ew_grid((long)M * N / 8), EW_THREADS, 0, ST, (const T*)in, ldi, (T*)out, ldo, act, M, N / 8); });
  return dfb_check_launch("act_fwd");
}
extern "C" int dfb200_act_bwd(const void* dout, long lddo, const void* dout2, long lddo2, const void* z, long ldz, void* din, long lddi, int dtype, int act,
                              int M, int N, float* colsum, void* stream) {
  DFB_REQUIRE(N % 8 == 0 && lddo % 8 == 0 && ldz % 8 == 0 && lddi % 8 == 0 && (!dout2 || lddo2 % 8 == 0) && (act == 1 || act == 2), "act_bwd: bad arguments");
  if (colsum) {
    const int rpb = pick_rows_per_block(M);
    dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(N / 8, EW_THREADS));
    DFB_DISPATCH_DTYPE(dtype, T, {
      dfb_launch(act_bwd_colsum_kernel<T>, grid, EW_THREADS, EW_THREADS * 8 * sizeof(float), ST, (const T*)dout, lddo, (const T*)dout2, lddo2, (const T*)z, ldz,
                 (T*)din, lddi, act, M, N, colsum, rpb);
    });
    return dfb_check_launch("act_bwd_colsum");
  }
  DFB_DISPATCH_DTYPE(dtype, T, {
    dfb_launch(act_bwd_kernel<T>, ew_grid((long)M * N / 8), EW_THREADS, 0, ST, (const T*)dout, lddo, (const T*)dout2, lddo2, (const T*)z, ldz, (T*)din, lddi, act, M, N / 8);
  });
  return dfb_check_launch("act_bwd");
}

extern "C" int dfb200_scale_residual_fwd(const float* res, const void* y, long ldy, int dtype, const float* ls, const float* scale_b, int M, int C,
                                         int rows_per_sample, float* out, void* stream) {
  DFB_REQUIRE(C % 8 == 0 && ldy % 8 == 0, "scale_residual: C, ldy %% 8 != 0");
  DFB_DISPATCH_DTYPE(dtype, T, {
    dfb_launch(scale_residual_fwd_kernel<T>, ew_grid((long)M * C / 8), EW_THREADS, 0, ST, res, (const T*)y, ldy, ls, scale_b, M, C / 8, rows_per_sample, out);
  });
  return dfb_check_launch("scale_residual_fwd");
}
extern "C" int dfb200_scale_residual_bwd(const float* dout, const void* y, long ldy, int dtype, const float* ls, const float* scale_b, int M, int C,
                                         int rows_per_sample, void* dy, long lddy, float* dls, float* dy_colsum, void* stream) {
  DFB_REQUIRE(C % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0, "scale_residual: C, ldy, lddy %% 8 != 0");
  const int rpb = pick_rows_per_block(M);
  dim3 grid(dfb_cdiv(M, rpb), dfb_cdiv(C / 8, EW_THREADS));
  DFB_DISPATCH_DTYPE(dtype, T, {
    if (dy_colsum)
      dfb_launch(scale_residual_bwd_kernel<T, 2>, grid, EW_THREADS, 2 * EW_THREADS * 8 * sizeof(float), ST, dout, (const T*)y, ldy, ls, scale_b, M, C,
                 rows_per_sample, (T*)dy, lddy, dls, dy_colsum, rpb);
    else
      dfb_launch(scale_residual_bwd_kernel<T, 1>, grid, EW_THREADS, EW_THREADS * 8 * sizeof(float), ST, dout, (const T*)y, ldy, ls, scale_b, M, C,
                 rows_per_sample, (T*)dy, lddy, dls, dy_colsum, rpb);
  });
  return dfb_check_launch("scale_residual_bwd");
}

extern "C" int dfb200_normalize_cols(const float* in, int B, int D, int R, float* out, float* norms, void* stream) {
  DFB_REQUIRE(R >= 1 && R <= 256, "normalize_cols: R=%d out of range (1..256)", R);
  dfb_launch(normalize_cols_kernel, B, 256, (256 + R) * sizeof(float), ST, in, D, R, out, norms);
  return dfb_check_launch("normalize_cols");
}
extern "C" int dfb200_softmax_rows(const float* in, int rows, int cols, float* out, void* stream) {
  if (rows <= 0) return DFB_OK;
  dfb_launch(softmax_rows_kernel, dfb_cdiv(rows, 8), 256, 0, ST, in, rows, cols, out);
  return dfb_check_launch("softmax_rows");
}
extern "C" int dfb200_softmax_rows_bwd(const float* dout, const float* out, int rows, int cols, float* din, void* stream) {
  if (rows <= 0) return DFB_OK;
  dfb_launch(softmax_rows_bwd_kernel, dfb_cdiv(rows, 8), 256, 0, ST, dout, out, rows, cols, din);
  return dfb_check_launch("softmax_rows_bwd");
}
extern "C" int dfb200_mu_update(const float* a, const float* num, const float* den, float eps, long n, float* out, void* out_lo, int lo_dtype, void* stream) {
  DFB_DISPATCH_DTYPE(lo_dtype, TL, { dfb_launch(mu_update_kernel<TL>, ew_grid(n), EW_THREADS, 0, ST, a, num, den, eps, n, out, (TL*)out_lo); });
  return dfb_check_launch("mu_update");
}
extern "C" int dfb200_mu_update_bwd(const float* dout, const float* a, const float* num, const float* den, float eps, long n, float* da,
                                    int accumulate_da, void* dnum, long ld_dnum, int cols, void* dden, int lo_dtype, void* stream) {
  DFB_REQUIRE(cols > 0 && ld_dnum >= cols && dnum && dden, "mu_update_bwd: bad arguments");
  DFB_DISPATCH_DTYPE(lo_dtype, TL, {
    dfb_launch(mu_update_bwd_kernel<TL>, ew_grid(n), EW_THREADS, 0, ST, dout, a, num, den, eps, n, da, accumulate_da, (TL*)dnum, ld_dnum, cols, (TL*)dden);
  });
  return dfb_check_launch("mu_update_bwd");
}

extern "C" int dfb200_cast(const void* in, int in_dtype, void* out, int out_dtype, long n, void* stream) {
  const int g = ew_grid(n, 4);
  if (in_dtype == 0 && out_dtype == 0) dfb_launch(cast_kernel<float, float>, g, EW_THREADS, 0, ST, (const float*)in, (float*)out, n);
  else if (in_dtype == 0 && out_dtype == 1) dfb_launch(cast_kernel<float, bf16>, g, EW_THREADS, 0, ST, (const float*)in, (bf16*)out, n);
  else if (in_dtype == 1 && out_dtype == 0) dfb_launch(cast_kernel<bf16, float>, g, EW_THREADS, 0, ST, (const bf16*)in, (float*)out, n);
  else if (in_dtype == 1 && out_dtype == 1) dfb_launch(cast_kernel<bf16, bf16>, g, EW_THREADS, 0, ST, (const bf16*)in, (bf16*)out, n);
  else { dfb_set_error("cast: bad dtypes"); return DFB_ERR_ARG; }
  return dfb_check_launch("cast");
}
extern "C" int dfb200_cast2d(const void* in, int in_dtype, long ld_in, void* out, int out_dtype, long ld_out, long rows, int cols, void* stream) {
  const long n = rows * cols;
  if (n <= 0) return DFB_OK;
  const int g = ew_grid(n, 4);
  if (in_dtype == 0 && out_dtype == 0) dfb_launch(cast2d_kernel<float, float>, g, EW_THREADS, 0, ST, (const float*)in, ld_in, (float*)out, ld_out, rows, cols);
  else if (in_dtype == 0 && out_dtype == 1) dfb_launch(cast2d_kernel<float, bf16>, g, EW_THREADS, 0, ST, (const float*)in, ld_in, (bf16*)out, ld_out, rows, cols);
  else if (in_dtype == 1 && out_dtype == 0) dfb_launch(cast2d_kernel<bf16, float>, g, EW_THREADS, 0, ST, (const bf16*)in, ld_in, (float*)out, ld_out, rows, cols);
  else if (in_dtype == 1 && out_dtype == 1) dfb_launch(cast2d_kernel<bf16, bf16>, g, EW_THREADS, 0, ST, (const bf16*)in, ld_in, (bf16*)out, ld_out, rows, cols);
  else { dfb_set_error("cast2d: bad dtypes"); return DFB_ERR_ARG; }
  return dfb_check_launch("cast2d");
}
extern "C" int dfb200_sym_cast(const void* in, int in_dtype, void* out, int out_dtype, int batch, int R, void* stream) {
  const long n = (long)batch * R * R;
  if (n <= 0) return DFB_OK;
  const int g = ew_grid(n, 1);
  if (in_dtype == 0 && out_dtype == 0) dfb_launch(sym_cast_kernel<float, float>, g, EW_THREADS, 0, ST, (const float*)in, (float*)out, batch, R);
  else if (in_dtype == 0 && out_dtype == 1) dfb_launch(sym_cast_kernel<float, bf16>, g, EW_THREADS, 0, ST, (const float*)in, (bf16*)out, batch, R);
  else if (in_dtype == 1 && out_dtype == 1) dfb_launch(sym_cast_kernel<bf16, bf16>, g, EW_THREADS, 0, ST, (const bf16*)in, (bf16*)out, batch, R);
  else { dfb_set_error("sym_cast: bad dtypes"); return DFB_ERR_ARG; }
  return dfb_check_launch("sym_cast");
}
extern "C" int dfb200_axpy(const void* x, int x_dtype, float alpha, void* y, int y_dtype, long n, void* stream) {
  const int g = ew_grid(n, 4);
  if (x_dtype == 0 && y_dtype == 0) dfb_launch(axpy_kernel<float, float>, g, EW_THREADS, 0, ST, (const float*)x, alpha, (float*)y, n);
  else if (x_dtype == 0 && y_dtype == 1) dfb_launch(axpy_kernel<float, bf16>, g, EW_THREADS, 0, ST, (const float*)x, alpha, (bf16*)y, n);
  else if (x_dtype == 1 && y_dtype == 0) dfb_launch(axpy_kernel<bf16, float>, g, EW_THREADS, 0, ST, (const bf16*)x, alpha, (float*)y, n);
  else if (x_dtype == 1 && y_dtype == 1) dfb_launch(axpy_kernel<bf16, bf16>, g, EW_THREADS, 0, ST, (const bf16*)x, alpha, (bf16*)y, n);
  else { dfb_set_error("axpy: bad dtypes"); return DFB_ERR_ARG; }
  return dfb_check_launch("axpy");
}

extern "C" int dfb200_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bias_c1, float bias_c2, float grad_scale, const float* wd_arr, const float* lr_arr,
                            const float* dyn, void* stream) {
  dfb_launch(adamw_kernel, ew_grid(n, 4), EW_THREADS, 0, ST, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2, grad_scale, wd_arr, lr_arr, dyn);
  return dfb_check_launch("adamw");
}
