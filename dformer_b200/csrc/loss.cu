// Fused x8 bilinear upsample (align_corners=False) + cross-entropy with ignore_index
// (models/builder.py:203,230).  Forward optionally materialises the NCHW fp32 logits the API returns;
// backward is a deterministic gather over the hi-res pixels that reference each low-res logit.
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

struct Lerp { int i0, i1; float w1; };
__device__ __forceinline__ Lerp lerp_coord(int o, int n_in, int n_out) {
  Lerp l;
  if (n_in == n_out) { l.i0 = o; l.i1 = o; l.w1 = 0.f; return l; }
  const float scale = (float)n_in / (float)n_out;
  float src = ((float)o + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  l.i0 = min((int)src, n_in - 1);
  l.i1 = min(l.i0 + 1, n_in - 1);
  l.w1 = src - (float)l.i0;
  return l;
}

template <typename T>
__device__ __forceinline__ float interp_logit(const T* __restrict__ base, int w, int ncls, const Lerp& ly, const Lerp& lx, int c) {
  const float v00 = to_f(base[((long)ly.i0 * w + lx.i0) * ncls + c]);
  const float v01 = to_f(base[((long)ly.i0 * w + lx.i1) * ncls + c]);
  const float v10 = to_f(base[((long)ly.i1 * w + lx.i0) * ncls + c]);
  const float v11 = to_f(base[((long)ly.i1 * w + lx.i1) * ncls + c]);
  const float top = (1.f - lx.w1) * v00 + lx.w1 * v01;
  const float bot = (1.f - lx.w1) * v10 + lx.w1 * v11;
  return (1.f - ly.w1) * top + ly.w1 * bot;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_ce_fwd_kernel(const T* __restrict__ small, int B, int h, int w, int ncls, int H, int W,
                                                              const int64_t* __restrict__ label, int ignore, float* __restrict__ out,
                                                              float* __restrict__ lse_out, float* loss_acc, T* __restrict__ up_lowp) {
  pdl_sync();
  __shared__ float red[32];
  float loss = 0.f, cnt = 0.f;
  const long n = (long)B * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long)W * H));
    const Lerp ly = lerp_coord(y, h, H), lx = lerp_coord(x, w, W);
    const T* base = small + (long)b * h * w * ncls;
    const long lab = label ? label[i] : (long)ignore;
    float m = -INFINITY, s = 0.f, picked = 0.f;
    for (int c = 0; c < ncls; ++c) {
      const float v = interp_logit(base, w, ncls, ly, lx, c);
      if (out) out[((long)b * ncls + c) * H * W + (long)y * W + x] = v;
      if (up_lowp) up_lowp[((long)b * ncls + c) * H * W + (long)y * W + x] = from_f<T>(v);
      if (v > m) { s = s * __expf(m - v) + 1.f; m = v; } else { s += __expf(v - m); }
      if ((long)c == lab) picked = v;
    }
    const float lse = m + __logf(s);
    if (lse_out) lse_out[i] = lse;
    if (label && lab != (long)ignore && lab >= 0 && lab < ncls) { loss += lse - picked; cnt += 1.f; }
  }
  if (loss_acc) {
    loss = block_sum(loss, red);
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) { atomicAdd(loss_acc, loss); atomicAdd(loss_acc + 1, cnt); }
  }
}

__global__ void ce_finalize_kernel(const float* acc, float* loss) {
  pdl_sync(); loss[0] = acc[0] / acc[1]; }

// thread = (low-res pixel, class); scans the hi-res pixels whose bilinear footprint includes the pixel
template <typename T, typename TD>
__global__ void upsample_ce_bwd_kernel(const T* __restrict__ small, int B, int h, int w, int ncls, int H, int W, const int64_t* __restrict__ label,
                                       int ignore, const float* __restrict__ lse, const float* __restrict__ loss_acc, const float* __restrict__ dloss,
                                       TD* __restrict__ dsmall) {
  pdl_sync();
  const long n = (long)B * h * w * ncls;
  const float gscale = dloss[0] / loss_acc[1];
  const float ry = (float)H / (float)h, rx = (float)W / (float)w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ncls);
    const long pix = i / ncls;
    const int ix = (int)(pix % w), iy = (int)((pix / w) % h), b = (int)(pix / ((long)w * h));
    int oy_lo = max(0, (int)floorf((iy - 1) * ry) - 1), oy_hi = min(H - 1, (int)ceilf((iy + 2) * ry) + 1);
    int ox_lo = max(0, (int)floorf((ix - 1) * rx) - 1), ox_hi = min(W - 1, (int)ceilf((ix + 2) * rx) + 1);
    if (iy == 0) oy_lo = 0;
    if (iy == h - 1) oy_hi = H - 1;
    if (ix == 0) ox_lo = 0;
    if (ix == w - 1) ox_hi = W - 1;
    const T* base = small + (long)b * h * w * ncls;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const Lerp ly = lerp_coord(oy, h, H);
      const float wy = (ly.i0 == iy ? 1.f - ly.w1 : 0.f) + (ly.i1 == iy ? ly.w1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const Lerp lx = lerp_coord(ox, w, W);
        const float wx = (lx.i0 == ix ? 1.f - lx.w1 : 0.f) + (lx.i1 == ix ? lx.w1 : 0.f);
        if (wx == 0.f) continue;
        const long hp = ((long)b * H + oy) * W + ox;
        const long lab = label[hp];
        if (lab == (long)ignore || lab < 0 || lab >= ncls) continue;
        const float v = interp_logit(base, w, ncls, ly, lx, c);
        float g = __expf(v - lse[hp]);
        if ((long)c == lab) g -= 1.f;
        acc = fmaf(wy * wx, g, acc);
      }
    }
    dsmall[i] = from_f<TD>(acc * gscale);
  }
}

// ---- separable adjoint.  Pass 1 (rows): t[b,c,ly,ox] = sum_oy wy(ly,oy) * (softmax(up)[b,c,oy,ox] - onehot) * valid
// One thread owns (b, ly, ox) and a chunk of CH classes, so the label / log-sum-exp of each hi-res pixel is fetched once
// per chunk instead of once per class; all loads are coalesced along ox.
template <typename TU, int CH>
__global__ void upsample_ce_bwd_rows_kernel(const TU* __restrict__ up, int B, int h, int ncls, int H, int W, const int64_t* __restrict__ label, int ignore,
                                            const float* __restrict__ lse, float* __restrict__ t) {
  pdl_sync();
  const int nchunk = (ncls + CH - 1) / CH;
  const long n = (long)B * nchunk * h * W;
  const float ry = (float)H / (float)h;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    const int ly = (int)((i / W) % h);
    const int ck = (int)((i / ((long)W * h)) % nchunk);
    const int b = (int)(i / ((long)W * h * nchunk));
    const int c0 = ck * CH;
    int oy_lo = max(0, (int)floorf((ly - 1) * ry) - 1), oy_hi = min(H - 1, (int)ceilf((ly + 2) * ry) + 1);
    if (ly == 0) oy_lo = 0;
    if (ly == h - 1) oy_hi = H - 1;
    float acc[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const Lerp l = lerp_coord(oy, h, H);
      const float wy = (l.i0 == ly ? 1.f - l.w1 : 0.f) + (l.i1 == ly ? l.w1 : 0.f);
      if (wy == 0.f) continue;
      const long hp = ((long)b * H + oy) * W + ox;
      const long lab = label[hp];
      if (lab == (long)ignore || lab < 0 || lab >= ncls) continue;
      const float ls = lse[hp];
      const TU* src = up + ((long)b * ncls + c0) * H * W + (long)oy * W + ox;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (c0 + j < ncls) {
          float g = __expf(to_f(src[(long)j * H * W]) - ls);
          if ((long)(c0 + j) == lab) g -= 1.f;
          acc[j] = fmaf(wy, g, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (c0 + j < ncls) t[(((long)b * ncls + c0 + j) * h + ly) * W + ox] = acc[j];
  }
}
// Pass 2 (columns): ds[b,ly,lx,c] = scale * sum_ox wx(lx,ox) * t[b,c,ly,ox]
template <typename TD>
__global__ void upsample_ce_bwd_cols_kernel(const float* __restrict__ t, int B, int h, int w, int ncls, int W, const float* __restrict__ loss_acc,
                                            const float* __restrict__ dloss, TD* __restrict__ dsmall) {
  pdl_sync();
  const long n = (long)B * h * w * ncls;
  const float gscale = dloss[0] / loss_acc[1];
  const float rx = (float)W / (float)w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ncls);
    const long pix = i / ncls;
    const int lx = (int)(pix % w), ly = (int)((pix / w) % h), b = (int)(pix / ((long)w * h));
    int ox_lo = max(0, (int)floorf((lx - 1) * rx) - 1), ox_hi = min(W - 1, (int)ceilf((lx + 2) * rx) + 1);
    if (lx == 0) ox_lo = 0;
    if (lx == w - 1) ox_hi = W - 1;
    const float* row = t + (((long)b * ncls + c) * h + ly) * W;
    float acc = 0.f;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const Lerp l = lerp_coord(ox, w, W);
      const float wx = (l.i0 == lx ? 1.f - l.w1 : 0.f) + (l.i1 == lx ? l.w1 : 0.f);
      acc = fmaf(wx, row[ox], acc);
    }
    dsmall[i] = from_f<TD>(acc * gscale);
  }
}

// ---- separable adjoint, recompute form (no hi-res copy of the logits is kept by the forward pass).
// Pass 1 (rows): thread = (b, source row ly0, hi-res column ox, chunk of CH classes).  All hi-res rows whose bilinear
// source row i0 equals ly0 share the two horizontally interpolated low-res rows top = lerp_x(small[ly0]) and
// bot = lerp_x(small[min(ly0+1, h-1)]), so a hi-res logit costs one lerp; g = softmax - onehot is evaluated exactly once
// per hi-res pixel and class and split between the two source rows:
//   tA[b, ly0, ox, c] = sum (1 - wy) g        (belongs to low-res row ly0)
//   tB[b, ly0, ox, c] = sum wy g              (belongs to low-res row min(ly0 + 1, h - 1))
__device__ __forceinline__ int first_row_of(int ly0, int h, int H) {      // first hi-res row whose source row i0 is >= ly0
  if (ly0 <= 0) return 0;
  if (ly0 >= h) return H;
  int o = (int)ceilf(((float)ly0 + 0.5f) * (float)H / (float)h - 0.5f);
  o = max(0, min(o, H));
  while (o > 0 && lerp_coord(o - 1, h, H).i0 >= ly0) --o;
  while (o < H && lerp_coord(o, h, H).i0 < ly0) ++o;
  return o;
}

template <typename T, int CH>
__global__ void __launch_bounds__(256) upsample_ce_bwd_rows_fused_kernel(const T* __restrict__ small, int B, int h, int w, int ncls, int H, int W,
                                                                         const int64_t* __restrict__ label, int ignore, const float* __restrict__ lse,
                                                                         float* __restrict__ tA, float* __restrict__ tB) {
  pdl_sync();
  const int nchunk = (ncls + CH - 1) / CH;
  const long n = (long)B * h * nchunk * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    const int ck = (int)((i / W) % nchunk);
    const int ly0 = (int)((i / ((long)W * nchunk)) % h);
    const int b = (int)(i / ((long)W * nchunk * h));
    const int c0 = ck * CH;
    const Lerp lx = lerp_coord(ox, w, W);
    const int i1 = min(ly0 + 1, h - 1);
    const T* r0 = small + (((long)b * h + ly0) * w) * ncls + c0;
    const T* r1 = small + (((long)b * h + i1) * w) * ncls + c0;
    float top[CH], bot[CH], a[CH], bb[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      a[j] = bb[j] = top[j] = bot[j] = 0.f;
      if (c0 + j < ncls) {
        const float v00 = to_f(r0[(long)lx.i0 * ncls + j]), v01 = to_f(r0[(long)lx.i1 * ncls + j]);
        const float v10 = to_f(r1[(long)lx.i0 * ncls + j]), v11 = to_f(r1[(long)lx.i1 * ncls + j]);
        top[j] = (1.f - lx.w1) * v00 + lx.w1 * v01;
        bot[j] = (1.f - lx.w1) * v10 + lx.w1 * v11;
      }
    }
    const int o_lo = first_row_of(ly0, h, H), o_hi = first_row_of(ly0 + 1, h, H);
    for (int oy = o_lo; oy < o_hi; ++oy) {
      const long hp = ((long)b * H + oy) * W + ox;
      const long lab = label[hp];
      if (lab == (long)ignore || lab < 0 || lab >= ncls) continue;
      const float ls = lse[hp];
      const float wy = lerp_coord(oy, h, H).w1;
      const int lj = (int)lab - c0;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float v = (1.f - wy) * top[j] + wy * bot[j];
        float g = __expf(v - ls);
        if (j == lj) g -= 1.f;
        a[j] += g;
        bb[j] = fmaf(wy, g, bb[j]);
      }
    }
    const long o = (((long)b * h + ly0) * W + ox) * ncls + c0;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (c0 + j < ncls) { tA[o + j] = a[j] - bb[j]; tB[o + j] = bb[j]; }
  }
}
// ---- training forward (loss + per-pixel log-sum-exp only, nothing hi-res materialised), same row grouping as the recompute
// backward above: thread = (b, source row ly0, hi-res column ox) walks the <= MAXR hi-res rows whose bilinear source row is ly0.
// The two horizontally interpolated low-res rows are built once per class and shared by those rows (one lerp + one exp per
// hi-res logit instead of four strided loads + three lerps), 16-byte loads over the classes when ncls % 8 == 0, and one common
// shift max_c max(top_c, bot_c) >= every interpolated logit of the group replaces the online-softmax rescaling.
template <typename T, int MAXR, bool VEC>
__global__ void __launch_bounds__(256) upsample_ce_fwd_rows_kernel(const T* __restrict__ small, int B, int h, int w, int ncls, int H, int W,
                                                                   const int64_t* __restrict__ label, int ignore, float* __restrict__ lse_out,
                                                                   float* loss_acc) {
  pdl_sync();
  __shared__ float red[32];
  float loss = 0.f, cnt = 0.f;
  const long n = (long)B * h * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W), ly0 = (int)((i / W) % h), b = (int)(i / ((long)W * h));
    const Lerp lx = lerp_coord(ox, w, W);
    const int i1 = min(ly0 + 1, h - 1);
    const T* r00 = small + (((long)b * h + ly0) * w + lx.i0) * ncls;
    const T* r01 = small + (((long)b * h + ly0) * w + lx.i1) * ncls;
    const T* r10 = small + (((long)b * h + i1) * w + lx.i0) * ncls;
    const T* r11 = small + (((long)b * h + i1) * w + lx.i1) * ncls;
    const int o_lo = first_row_of(ly0, h, H), nr = min(first_row_of(ly0 + 1, h, H) - o_lo, MAXR);
    const float wx = lx.w1, ux = 1.f - lx.w1;
    auto rows8 = [&](int c0, float* top, float* bot) {          // horizontally interpolated logits of classes c0 .. c0 + 7
      float a[8], bq[8], c[8], d[8];
      if (VEC) {
        Vec8<T>::load(r00 + c0, a); Vec8<T>::load(r01 + c0, bq); Vec8<T>::load(r10 + c0, c); Vec8<T>::load(r11 + c0, d);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = c0 + j < ncls;
          a[j] = ok ? to_f(r00[c0 + j]) : -INFINITY; bq[j] = ok ? to_f(r01[c0 + j]) : -INFINITY;
          c[j] = ok ? to_f(r10[c0 + j]) : -INFINITY; d[j] = ok ? to_f(r11[c0 + j]) : -INFINITY;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { top[j] = ux * a[j] + wx * bq[j]; bot[j] = ux * c[j] + wx * d[j]; }
    };
    float M = -INFINITY;
    for (int c0 = 0; c0 < ncls; c0 += 8) {
      float top[8], bot[8];
      rows8(c0, top, bot);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (VEC || c0 + j < ncls) M = fmaxf(M, fmaxf(top[j], bot[j]));
    }
    float wy[MAXR], sum[MAXR], pk[MAXR];
    int lab[MAXR];
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      sum[k] = 0.f; pk[k] = 0.f; wy[k] = 0.f; lab[k] = -1;
      if (k < nr) {
        wy[k] = lerp_coord(o_lo + k, h, H).w1;
        const long lb = label[((long)b * H + o_lo + k) * W + ox];
        lab[k] = (lb != (long)ignore && lb >= 0 && lb < ncls) ? (int)lb : -1;
      }
    }
    for (int c0 = 0; c0 < ncls; c0 += 8) {
      float top[8], bot[8];
      rows8(c0, top, bot);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (VEC || c0 + j < ncls) {
#pragma unroll
          for (int k = 0; k < MAXR; ++k) {
            const float v = (1.f - wy[k]) * top[j] + wy[k] * bot[j];
            sum[k] += __expf(v - M);
            if (c0 + j == lab[k]) pk[k] = v;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      if (k < nr) {
        const float lse = M + __logf(sum[k]);
        lse_out[((long)b * H + o_lo + k) * W + ox] = lse;
        if (lab[k] >= 0) { loss += lse - pk[k]; cnt += 1.f; }
      }
    }
  }
  if (loss_acc) {
    loss = block_sum(loss, red);
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) { atomicAdd(loss_acc, loss); atomicAdd(loss_acc + 1, cnt); }
  }
}


// ---- training step in ONE pass: loss + per-pixel log-sum-exp AND the (unscaled) gradient w.r.t. the low-res logits.
// CTA = (image, source row ly0, segment of CE_SEG hi-res columns); thread = one hi-res column, walking the <= MAXR hi-res rows whose
// bilinear source row is ly0 (same row grouping as upsample_ce_fwd_rows_kernel: the two horizontally interpolated low-res rows are
// shared by those rows).  Three sweeps over the classes: (1) common shift, (2) sum of exponentials -> lse, loss, (3) per chunk of 8
// classes g = softmax - onehot, split between source rows ly0 / ly0 + 1 (weights 1 - wy / wy) and -- through shared-memory
// reductions -- between the two source columns of the thread's hi-res column; the segment's partial sums are added to the fp32
// gradient with one atomic per (low-res pixel, class).  Nothing hi-res is written besides lse (optional); the gradient is scaled by
// dloss / #valid later (ce_grad_finalize_kernel), so this kernel needs neither.  Replaces forward + rows adjoint + columns adjoint
// (three launches, 200 MB of fp32 partials through HBM).
constexpr int CE_SEG = 128;
template <typename T, int MAXR, bool VEC>
__global__ void __launch_bounds__(CE_SEG) upsample_ce_train_kernel(const T* __restrict__ small, int B, int h, int w, int ncls, int H, int W,
                                                                   const int64_t* __restrict__ label, int ignore, float* __restrict__ lse_out,
                                                                   float* loss_acc, float* __restrict__ dgrad) {
  pdl_sync();
  __shared__ float red[32];
  __shared__ float tAs[CE_SEG][9], tBs[CE_SEG][9];  // per hi-res column of the segment: shares of source rows ly0 / ly0 + 1, 8 classes (+1: bank skew)
  __shared__ int s_i0[CE_SEG], s_i1[CE_SEG];        // horizontal source columns / weight of every hi-res column of the segment
  __shared__ float s_w1[CE_SEG];
  const int nseg = (W + CE_SEG - 1) / CE_SEG;
  const int seg = blockIdx.x % nseg, ly0 = (blockIdx.x / nseg) % h, b = blockIdx.x / (nseg * h);
  const int s0 = seg * CE_SEG, s1 = min(W, s0 + CE_SEG);
  const int ox = s0 + threadIdx.x;
  const bool col_ok = ox < s1;
  const Lerp lx = lerp_coord(col_ok ? ox : s1 - 1, w, W);
  const int lx_min = lerp_coord(s0, w, W).i0;
  const int nlx = lerp_coord(s1 - 1, w, W).i1 - lx_min + 1;      // <= CE_SEG + 2 because W >= w
  const int i1 = min(ly0 + 1, h - 1);
  const T* r00 = small + (((long)b * h + ly0) * w + lx.i0) * ncls;
  const T* r01 = small + (((long)b * h + ly0) * w + lx.i1) * ncls;
  const T* r10 = small + (((long)b * h + i1) * w + lx.i0) * ncls;
  const T* r11 = small + (((long)b * h + i1) * w + lx.i1) * ncls;
  const int o_lo = first_row_of(ly0, h, H);
  const int nr = col_ok ? min(first_row_of(ly0 + 1, h, H) - o_lo, MAXR) : 0;
  const float wx = lx.w1, ux = 1.f - lx.w1;
  s_i0[threadIdx.x] = col_ok ? lx.i0 : -1;
  s_i1[threadIdx.x] = col_ok ? lx.i1 : -1;
  s_w1[threadIdx.x] = wx;
  auto rows8 = [&](int c0, float* top, float* bot) {          // horizontally interpolated logits of classes c0 .. c0 + 7
    float a[8], bq[8], c[8], d[8];
    if (VEC) {
      Vec8<T>::load(r00 + c0, a); Vec8<T>::load(r01 + c0, bq); Vec8<T>::load(r10 + c0, c); Vec8<T>::load(r11 + c0, d);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = c0 + j < ncls;
        a[j] = ok ? to_f(r00[c0 + j]) : -INFINITY; bq[j] = ok ? to_f(r01[c0 + j]) : -INFINITY;
        c[j] = ok ? to_f(r10[c0 + j]) : -INFINITY; d[j] = ok ? to_f(r11[c0 + j]) : -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { top[j] = ux * a[j] + wx * bq[j]; bot[j] = ux * c[j] + wx * d[j]; }
  };
  // ---- sweep 1: one shift >= every interpolated logit of the group
  float M = -INFINITY;
  for (int c0 = 0; c0 < ncls; c0 += 8) {
    float top[8], bot[8];
    rows8(c0, top, bot);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (VEC || c0 + j < ncls) M = fmaxf(M, fmaxf(top[j], bot[j]));
  }
  float wy[MAXR], sum[MAXR], pk[MAXR];
  int lab[MAXR];
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    sum[k] = 0.f; pk[k] = 0.f; wy[k] = 0.f; lab[k] = -1;
    if (k < nr) {
      wy[k] = lerp_coord(o_lo + k, h, H).w1;
      const long lb = label[((long)b * H + o_lo + k) * W + ox];
      lab[k] = (lb != (long)ignore && lb >= 0 && lb < ncls) ? (int)lb : -1;
    }
  }
  // ---- sweep 2: sums of exponentials, picked logits
  for (int c0 = 0; c0 < ncls; c0 += 8) {
    float top[8], bot[8];
    rows8(c0, top, bot);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (VEC || c0 + j < ncls) {
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const float v = (1.f - wy[k]) * top[j] + wy[k] * bot[j];
          sum[k] += __expf(v - M);
          if (c0 + j == lab[k]) pk[k] = v;
        }
      }
    }
  }
  float loss = 0.f, cnt = 0.f;
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    if (k < nr) {
      const float lse = M + __logf(sum[k]);
      if (lse_out) lse_out[((long)b * H + o_lo + k) * W + ox] = lse;
      if (lab[k] >= 0) { loss += lse - pk[k]; cnt += 1.f; }
    }
    sum[k] = (k < nr && lab[k] >= 0) ? 1.f / sum[k] : 0.f;      // from here on: 1 / sum for the valid pixels, 0 (no gradient) otherwise
  }
  loss = block_sum(loss, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) { atomicAdd(loss_acc, loss); atomicAdd(loss_acc + 1, cnt); }
  // ---- sweep 3: gradient, one chunk of 8 classes at a time
  const float rx = (float)W / (float)w;
  for (int c0 = 0; c0 < ncls; c0 += 8) {
    float top[8], bot[8], ga[8], gb[8];
    rows8(c0, top, bot);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ga[j] = 0.f; gb[j] = 0.f;
      if (VEC || c0 + j < ncls) {
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const float v = (1.f - wy[k]) * top[j] + wy[k] * bot[j];
          float g = __expf(v - M) * sum[k];
          if (c0 + j == lab[k]) g -= 1.f;
          ga[j] += g;
          gb[j] = fmaf(wy[k], g, gb[j]);
        }
      }
    }
    __syncthreads();                                            // the previous chunk's shares have been consumed
#pragma unroll
    for (int j = 0; j < 8; ++j) { tAs[threadIdx.x][j] = ga[j] - gb[j]; tBs[threadIdx.x][j] = gb[j]; }
    __syncthreads();
    // gather form of the horizontal adjoint: one thread per (low-res column, class) sums the shares of the hi-res columns of this
    // segment whose source columns include it
    for (int i = threadIdx.x; i < nlx * 8; i += CE_SEG) {
      const int q = i >> 3, j = i & 7, lxq = lx_min + q;
      if (c0 + j >= ncls) continue;
      int t_lo = max(s0, (int)floorf((lxq - 1) * rx) - 1) - s0, t_hi = min(s1 - 1, (int)ceilf((lxq + 2) * rx) + 1) - s0;
      if (lxq == 0) t_lo = 0;
      if (lxq == w - 1) t_hi = s1 - 1 - s0;
      float va = 0.f, vb = 0.f;
      for (int t = t_lo; t <= t_hi; ++t) {
        const float w1 = s_w1[t];
        const float wgt = (s_i0[t] == lxq ? 1.f - w1 : 0.f) + (s_i1[t] == lxq ? w1 : 0.f);
        va = fmaf(wgt, tAs[t][j], va);
        vb = fmaf(wgt, tBs[t][j], vb);
      }
      if (i1 == ly0) {
        atomicAdd(dgrad + (((long)b * h + ly0) * w + lxq) * ncls + c0 + j, va + vb);
      } else {
        atomicAdd(dgrad + (((long)b * h + ly0) * w + lxq) * ncls + c0 + j, va);
        atomicAdd(dgrad + (((long)b * h + i1) * w + lxq) * ncls + c0 + j, vb);
      }
    }
  }
}

template <typename TD>
__global__ void ce_grad_finalize_kernel(const float* __restrict__ g, long n, const float* __restrict__ loss_acc, const float* __restrict__ dloss,
                                        TD* __restrict__ out) {
  pdl_sync();
  const float gscale = dloss[0] / loss_acc[1];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = from_f<TD>(g[i] * gscale);
}

// Pass 2 (columns): ds[b,ly,lx,c] = scale * sum_ox wx(lx,ox) * (tA[ly] + tB[ly-1] (+ tB[h-1] on the last row))[ox, c]
template <typename TD>
__global__ void __launch_bounds__(256) upsample_ce_bwd_cols_fused_kernel(const float* __restrict__ tA, const float* __restrict__ tB, int B, int h, int w,
                                                                         int ncls, int W, const float* __restrict__ loss_acc,
                                                                         const float* __restrict__ dloss, TD* __restrict__ dsmall) {
  pdl_sync();
  const long n = (long)B * h * w * ncls;
  const float gscale = dloss[0] / loss_acc[1];
  const float rx = (float)W / (float)w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ncls);
    const long pix = i / ncls;
    const int lx = (int)(pix % w), ly = (int)((pix / w) % h), b = (int)(pix / ((long)w * h));
    int ox_lo = max(0, (int)floorf((lx - 1) * rx) - 1), ox_hi = min(W - 1, (int)ceilf((lx + 2) * rx) + 1);
    if (lx == 0) ox_lo = 0;
    if (lx == w - 1) ox_hi = W - 1;
    const float* ra = tA + (((long)b * h + ly) * W) * ncls + c;
    const float* rb = ly > 0 ? tB + (((long)b * h + ly - 1) * W) * ncls + c : nullptr;
    const float* rl = ly == h - 1 ? tB + (((long)b * h + ly) * W) * ncls + c : nullptr;
    float acc = 0.f;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const Lerp l = lerp_coord(ox, w, W);
      const float wx = (l.i0 == lx ? 1.f - l.w1 : 0.f) + (l.i1 == lx ? l.w1 : 0.f);
      if (wx == 0.f) continue;
      float v = ra[(long)ox * ncls];
      if (rb) v += rb[(long)ox * ncls];
      if (rl) v += rl[(long)ox * ncls];
      acc = fmaf(wx, v, acc);
    }
    dsmall[i] = from_f<TD>(acc * gscale);
  }
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_upsample_ce_fwd(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label, int ignore,
                                      float* out_nchw, float* lse, float* loss_acc, void* up_lowp, void* stream) {
  const long n = (long)B * H * W;
  // training forward (only the loss and the per-pixel log-sum-exp are wanted): the row-grouped kernel; a source row feeds at most
  // ceil(1.5 * H / h) hi-res rows (the first group also takes the rows whose source coordinate clamps to 0)
  constexpr int MAXR = 12;
  if (out_nchw == nullptr && up_lowp == nullptr && lse != nullptr && label != nullptr && H >= h && (3L * H + 2 * h - 1) / (2L * h) <= MAXR) {
    const long n1 = (long)B * h * W;
    long g1 = (n1 + 255) / 256;
    if (g1 < 1) g1 = 1;
    const bool vec = (ncls % 8) == 0 && (reinterpret_cast<uintptr_t>(logits_small) & 15) == 0;
    DFB_DISPATCH_DTYPE(dtype, T, {
      if (vec) dfb_launch(upsample_ce_fwd_rows_kernel<T, MAXR, true>, (unsigned)g1, 256, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, loss_acc);
      else dfb_launch(upsample_ce_fwd_rows_kernel<T, MAXR, false>, (unsigned)g1, 256, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, loss_acc);
    });
    return dfb_check_launch("upsample_ce_fwd_rows");
  }
  long g = (n + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  DFB_DISPATCH_DTYPE(dtype, T, {
    dfb_launch(upsample_ce_fwd_kernel<T>, (int)g, 256, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, out_nchw, lse, loss_acc, (T*)up_lowp);
  });
  return dfb_check_launch("upsample_ce_fwd");
}

// One-pass training form: loss_acc += (sum of NLL, #valid), dgrad (fp32 [B*h*w, ncls], zero-initialised by the caller) += the gradient of
// the SUM of NLL w.r.t. the low-res logits; dfb200_ce_grad_finalize scales it by dloss / #valid.  Returns DFB_ERR_UNSUPPORTED for geometries
// the row-grouped kernel does not cover (down-sampling, more than 12 hi-res rows per source row): use fwd + bwd_fused there.
extern "C" int dfb200_upsample_ce_train(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label, int ignore,
                                        float* lse, float* loss_acc, float* dgrad, void* stream) {
  constexpr int MAXR = 12;
  if (!(label && loss_acc && dgrad && H >= h && W >= w && (3L * H + 2 * h - 1) / (2L * h) <= MAXR)) {
    dfb_set_error("upsample_ce_train: unsupported geometry (%dx%d -> %dx%d)", h, w, H, W);
    return DFB_ERR_UNSUPPORTED;
  }
  const long grid = (long)B * h * ((W + CE_SEG - 1) / CE_SEG);
  const bool vec = (ncls % 8) == 0 && (reinterpret_cast<uintptr_t>(logits_small) & 15) == 0;
  DFB_DISPATCH_DTYPE(dtype, T, {
    if (vec) dfb_launch(upsample_ce_train_kernel<T, MAXR, true>, (unsigned)grid, CE_SEG, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, loss_acc, dgrad);
    else dfb_launch(upsample_ce_train_kernel<T, MAXR, false>, (unsigned)grid, CE_SEG, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, loss_acc, dgrad);
  });
  return dfb_check_launch("upsample_ce_train");
}

extern "C" int dfb200_ce_grad_finalize(const float* dgrad, long n, const float* loss_acc, const float* dloss, void* out, int out_dtype, void* stream) {
  long g = (n + 255) / 256;
  if (g > 148L * 8) g = 148L * 8;
  if (g < 1) g = 1;
  DFB_DISPATCH_DTYPE(out_dtype, TD, { dfb_launch(ce_grad_finalize_kernel<TD>, (int)g, 256, 0, ST, dgrad, n, loss_acc, dloss, (TD*)out); });
  return dfb_check_launch("ce_grad_finalize");
}

extern "C" int dfb200_ce_finalize(const float* loss_acc, float* loss, void* stream) {
  dfb_launch(ce_finalize_kernel, 1, 1, 0, ST, loss_acc, loss);
  return dfb_check_launch("ce_finalize");
}

extern "C" int dfb200_upsample_ce_bwd_sep(const void* up, int up_dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label, int ignore,
                                          const float* lse, const float* loss_acc, const float* dloss, float* scratch, void* dlogits_small, int dl_dtype,
                                          void* stream) {
  constexpr int CH = 10;
  const long n1 = (long)B * ((ncls + CH - 1) / CH) * h * W, n2 = (long)B * h * w * ncls;
  long g1 = (n1 + 255) / 256, g2 = (n2 + 255) / 256;
  if (g1 > 148L * 32) g1 = 148L * 32;
  if (g1 < 1) g1 = 1;
  if (g2 < 1) g2 = 1;
  DFB_DISPATCH_DTYPE(up_dtype, TU, { dfb_launch(upsample_ce_bwd_rows_kernel<TU, CH>, (int)g1, 256, 0, ST, (const TU*)up, B, h, ncls, H, W, label, ignore, lse, scratch); });
  int rc = dfb_check_launch("upsample_ce_bwd_rows");
  if (rc) return rc;
  DFB_DISPATCH_DTYPE(dl_dtype, TD, { dfb_launch(upsample_ce_bwd_cols_kernel<TD>, (int)g2, 256, 0, ST, scratch, B, h, w, ncls, W, loss_acc, dloss, (TD*)dlogits_small); });
  return dfb_check_launch("upsample_ce_bwd_cols");
}

extern "C" int dfb200_upsample_ce_bwd(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label, int ignore,
                                      const float* lse, const float* loss_acc, const float* dloss, void* dlogits_small, int dl_dtype, void* stream) {
  const long n = (long)B * h * w * ncls;
  long g = (n + 255) / 256;
  if (g < 1) g = 1;
#define L(T, TD) dfb_launch(upsample_ce_bwd_kernel<T, TD>, (int)g, 256, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, loss_acc, dloss, (TD*)dlogits_small)
  const int key = dtype * 2 + dl_dtype;
  switch (key) { case 0: L(float, float); break; case 1: L(float, bf16); break; case 2: L(bf16, float); break; case 3: L(bf16, bf16); break;
    default: dfb_set_error("upsample_ce_bwd: bad dtypes"); return DFB_ERR_ARG; }
#undef L
  return dfb_check_launch("upsample_ce_bwd");
}

extern "C" int dfb200_upsample_ce_bwd_fused(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label,
                                            int ignore, const float* lse, const float* loss_acc, const float* dloss, float* scratch,
                                            void* dlogits_small, int dl_dtype, void* stream) {
  constexpr int CH = 10;
  const long per = (long)B * h * W * ncls;
  float* tA = scratch;
  float* tB = scratch + per;
  const long n1 = (long)B * h * ((ncls + CH - 1) / CH) * W, n2 = (long)B * h * w * ncls;
  long g1 = (n1 + 255) / 256, g2 = (n2 + 255) / 256;
  if (g1 < 1) g1 = 1;
  if (g2 < 1) g2 = 1;
  DFB_DISPATCH_DTYPE(dtype, T, {
    dfb_launch(upsample_ce_bwd_rows_fused_kernel<T, CH>, (unsigned)g1, 256, 0, ST, (const T*)logits_small, B, h, w, ncls, H, W, label, ignore, lse, tA, tB);
  });
  int rc = dfb_check_launch("upsample_ce_bwd_rows_fused");
  if (rc) return rc;
  DFB_DISPATCH_DTYPE(dl_dtype, TD, {
    dfb_launch(upsample_ce_bwd_cols_fused_kernel<TD>, (unsigned)g2, 256, 0, ST, tA, tB, B, h, w, ncls, W, loss_acc, dloss, (TD*)dlogits_small);
  });
  return dfb_check_launch("upsample_ce_bwd_cols_fused");
}
