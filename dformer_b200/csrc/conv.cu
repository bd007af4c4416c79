// Dense 3x3 stride-2 pad-1 convolutions of the stems / downsample layers (DFormer.py:194-228) as
// im2col gather + GEMM.  The gather writes [B*Ho*Wo, ld] rows with k = (ky*3+kx)*Cin + ci (zero padded
// to ld) in the compute dtype; the input is addressed with generic strides so NCHW network inputs and
// the channel-0 slice of modal_x are read in place.
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

// generic strides (NCHW network inputs, Cin = 3 / 1): one thread per OUTPUT ROW builds its ld columns in 8-wide vectors (one decode
// of the pixel coordinates per row instead of one 64-bit division chain per element, 16-byte stores)
template <typename TI, typename TO, typename IT>
__global__ void __launch_bounds__(256) im2col_scalar_kernel(const TI* __restrict__ in, long sb, long sy, long sx, long sc, int B, int H, int W, int Cin,
                                                            TO* __restrict__ out, int ld) {
  pdl_sync();
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const IT n = (IT)B * Ho * Wo;
  for (IT row = blockIdx.x * (IT)blockDim.x + threadIdx.x; row < n; row += (IT)gridDim.x * blockDim.x) {
    const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((IT)Wo * Ho));
    const TI* base = in + b * sb;
    TO* orow = out + (long)row * ld;
    for (int k0 = 0; k0 < ld; k0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        v[j] = 0.f;
        if (k < 9 * Cin) {
          const int tap = k / Cin, ci = k - tap * Cin;
          const int iy = 2 * oy - 1 + tap / 3, ix = 2 * ox - 1 + tap % 3;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[j] = to_f(base[iy * sy + ix * sx + ci * sc]);
        }
      }
      Vec8<TO>::store(orow + k0, v);
    }
  }
}

// channels-last, Cin % 8 == 0, ld == 9*Cin: one thread moves one 8-channel vector of one tap
template <typename TI, typename TO, typename IT>
__global__ void im2col_vec_kernel(const TI* __restrict__ in, int B, int H, int W, int Cin, TO* __restrict__ out) {
  pdl_sync();
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int nvec = Cin >> 3;
  const IT n = (IT)B * Ho * Wo * 9 * nvec;
  for (IT i = blockIdx.x * (IT)blockDim.x + threadIdx.x; i < n; i += (IT)gridDim.x * blockDim.x) {
    const int cv = (int)(i % nvec);
    const int tap = (int)((i / nvec) % 9);
    const IT row = i / ((IT)nvec * 9);
    const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((IT)Wo * Ho));
    const int iy = 2 * oy - 1 + tap / 3, ix = 2 * ox - 1 + tap % 3;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) Vec8<TI>::load(in + (((long)b * H + iy) * W + ix) * Cin + cv * 8, v);
    Vec8<TO>::store(out + (long)i * 8, v);
  }
}

// col2im (gather): input pixel (y, x) is referenced by <= 2 x 2 (output pixel, tap) pairs
template <typename TC, typename TI, typename IT>
__global__ void col2im_kernel(const TC* __restrict__ dcol, int ld, int B, int H, int W, int Cin, TI* __restrict__ din) {
  pdl_sync();
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int nvec = Cin >> 3;
  const IT n = (IT)B * H * W * nvec;
  for (IT i = blockIdx.x * (IT)blockDim.x + threadIdx.x; i < n; i += (IT)gridDim.x * blockDim.x) {
    const int cv = (int)(i % nvec);
    const IT pix = i / nvec;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((IT)W * H));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;            // = 2 * oy
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        float g[8];
        Vec8<TC>::load(dcol + (((long)b * Ho + oy) * Wo + ox) * ld + (ky * 3 + kx) * Cin + cv * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    }
    Vec8<TI>::store(din + (long)pix * Cin + cv * 8, acc);
  }
}

inline int ew_grid(long n) {
  long b = (n + 255) / 256;
  if (b < 1) b = 1;
  const long cap = 148L * 16;
  return (int)(b > cap ? cap : b);
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_im2col3x3s2_fwd(const void* in, int in_dtype, long sb, long sy, long sx, long sc, int B, int H, int W, int Cin, void* out,
                                      int out_dtype, int ld, void* stream) {
  DFB_REQUIRE(ld >= 9 * Cin, "im2col: ld (%d) < 9*Cin (%d)", ld, 9 * Cin);
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const bool cl = (sc == 1 && sx == Cin && sy == (long)W * Cin && sb == (long)H * W * Cin);
  const bool vec = cl && (Cin % 8 == 0) && ld == 9 * Cin;
  // 32-bit index arithmetic whenever the element count allows it (64-bit div / mod chains dominated these gathers)
  const long n_vec = (long)B * Ho * Wo * 9 * Cin / 8, n_rows = (long)B * Ho * Wo;
  const bool small = n_vec < (1L << 30) && (long)B * Ho * Wo * ld < (1L << 31);
  DFB_REQUIRE(vec || ld % 8 == 0, "im2col: ld (%d) must be a multiple of 8", ld);
  DFB_REQUIRE(vec || (reinterpret_cast<uintptr_t>(out) & 15) == 0, "im2col: output must be 16-byte aligned");
#define LV(TI, TO) do { if (small) dfb_launch(im2col_vec_kernel<TI, TO, unsigned>, ew_grid(n_vec), 256, 0, ST, (const TI*)in, B, H, W, Cin, (TO*)out); \
                        else dfb_launch(im2col_vec_kernel<TI, TO, long>, ew_grid(n_vec), 256, 0, ST, (const TI*)in, B, H, W, Cin, (TO*)out); } while (0)
#define LS(TI, TO) do { if (small) dfb_launch(im2col_scalar_kernel<TI, TO, unsigned>, ew_grid(n_rows), 256, 0, ST, (const TI*)in, sb, sy, sx, sc, B, H, W, Cin, (TO*)out, ld); \
                        else dfb_launch(im2col_scalar_kernel<TI, TO, long>, ew_grid(n_rows), 256, 0, ST, (const TI*)in, sb, sy, sx, sc, B, H, W, Cin, (TO*)out, ld); } while (0)
  const int key = in_dtype * 2 + out_dtype;
  if (vec) {
    switch (key) { case 0: LV(float, float); break; case 1: LV(float, bf16); break; case 2: LV(bf16, float); break; case 3: LV(bf16, bf16); break;
      default: dfb_set_error("im2col: bad dtypes"); return DFB_ERR_ARG; }
  } else {
    switch (key) { case 0: LS(float, float); break; case 1: LS(float, bf16); break; case 2: LS(bf16, float); break; case 3: LS(bf16, bf16); break;
      default: dfb_set_error("im2col: bad dtypes"); return DFB_ERR_ARG; }
  }
#undef LV
#undef LS
  return dfb_check_launch("im2col3x3s2_fwd");
}

extern "C" int dfb200_im2col3x3s2_bwd(const void* dcol, int col_dtype, int ld, int B, int H, int W, int Cin, void* din, int in_dtype, void* stream) {
  DFB_REQUIRE(Cin % 8 == 0 && ld % 8 == 0, "col2im: Cin and ld must be multiples of 8");
  const long n = (long)B * H * W * Cin / 8;
  const int g = ew_grid(n);
  const bool small = n < (1L << 30);
#define L(TC, TI) do { if (small) dfb_launch(col2im_kernel<TC, TI, unsigned>, g, 256, 0, ST, (const TC*)dcol, ld, B, H, W, Cin, (TI*)din); \
                       else dfb_launch(col2im_kernel<TC, TI, long>, g, 256, 0, ST, (const TC*)dcol, ld, B, H, W, Cin, (TI*)din); } while (0)
  const int key = col_dtype * 2 + in_dtype;
  switch (key) { case 0: L(float, float); break; case 1: L(float, bf16); break; case 2: L(bf16, float); break; case 3: L(bf16, bf16); break;
    default: dfb_set_error("col2im: bad dtypes"); return DFB_ERR_ARG; }
#undef L
  return dfb_check_launch("im2col3x3s2_bwd");
}
