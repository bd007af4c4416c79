// Training-time input pipeline on the GPU: the reference's TrainPre (utils/dataloader/dataloader.py:20-73,
// utils/transforms.py:27-75,182-187) -- random mirror, random scale (cv2.resize INTER_LINEAR for the images, INTER_NEAREST for the
// label), normalisation, random crop + constant padding, HWC -> CHW -- as ONE kernel per batch.  A thread produces one pixel of the
// crop for all three tensors: it maps the pixel back through pad -> crop -> scale -> mirror to the uint8 source, evaluates OpenCV's
// 11-bit fixed-point bilinear filter bit-exactly (coefficients as resize.cpp builds them: float source offset, cvRound(w * 2048),
// row = S[sx]*a0 + S[sx+1]*a1, dst = (((b0*(row0>>4))>>16) + ((b1*(row1>>4))>>16) + 2) >> 2; exact-2x down-scaling = 2x2 mean) and
// normalises through a 256-entry table per channel built on the host in float64 like numpy does.  Nothing but the uint8 sources is read.
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

struct Coef { int s0, s1, a0, a1; };

// x direction (resize.cpp: xofs / ialpha): out-of-range source columns collapse onto the border with weight (1, 0)
__device__ __forceinline__ Coef coef_x(int d, int n_src, int n_dst) {
  const double scale = (double)n_src / (double)n_dst;
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);       // no FMA contraction: numpy / OpenCV round twice
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
  Coef c;
  c.s0 = s;
  c.s1 = min(s + 1, n_src - 1);
  c.a0 = __float2int_rn(__fsub_rn(1.f, f) * 2048.f);
  c.a1 = __float2int_rn(f * 2048.f);
  return c;
}
// y direction (yofs / ibeta): the fractional weight is kept, the two row indices are clipped
__device__ __forceinline__ Coef coef_y(int d, int n_src, int n_dst) {
  const double scale = (double)n_src / (double)n_dst;
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  Coef c;
  c.s0 = min(max(s, 0), n_src - 1);
  c.s1 = min(max(s + 1, 0), n_src - 1);
  c.a0 = __float2int_rn(__fsub_rn(1.f, f) * 2048.f);
  c.a1 = __float2int_rn(f * 2048.f);
  return c;
}

// one resized uint8 sample of channel ch at (cy, cx) of the sh x sw image; src is [H, W, 3] uint8, optionally mirrored along W
__device__ __forceinline__ void resized_rgb(const uint8_t* __restrict__ src, int H, int W, int sh, int sw, int cy, int cx, int flip, int area2x,
                                            const Coef& ky, const Coef& kx, int out[3]) {
  auto px = [&](int y, int x, int ch) -> int { return src[((long)y * W + (flip ? W - 1 - x : x)) * 3 + ch]; };
  if (sh == H && sw == W) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[ch] = px(cy, cx, ch);
    return;
  }
  if (area2x) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[ch] = (px(2 * cy, 2 * cx, ch) + px(2 * cy, 2 * cx + 1, ch) + px(2 * cy + 1, 2 * cx, ch) + px(2 * cy + 1, 2 * cx + 1, ch) + 2) >> 2;
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int r0 = px(ky.s0, kx.s0, ch) * kx.a0 + px(ky.s0, kx.s1, ch) * kx.a1;
    const int r1 = px(ky.s1, kx.s0, ch) * kx.a0 + px(ky.s1, kx.s1, ch) * kx.a1;
    out[ch] = (((ky.a0 * (r0 >> 4)) >> 16) + ((ky.a1 * (r1 >> 4)) >> 16) + 2) >> 2;
  }
}

// params per sample: flip, sh, sw, pos_h, pos_w  (the reference's random draws); everything else follows from them
__global__ void __launch_bounds__(256) train_pre_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ modal, const uint8_t* __restrict__ label,
                                                        int B, int H, int W, const int* __restrict__ params, const float* __restrict__ lut_rgb,
                                                        const float* __restrict__ lut_modal, int crop_h, int crop_w, float* __restrict__ out_rgb,
                                                        float* __restrict__ out_modal, int64_t* __restrict__ out_label) {
  pdl_sync();
  const long n = (long)B * crop_h * crop_w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % crop_w), y = (int)((i / crop_w) % crop_h), b = (int)(i / ((long)crop_w * crop_h));
    const int* p = params + b * 5;
    const int flip = p[0], sh = p[1], sw = p[2], pos_h = p[3], pos_w = p[4];
    // crop (clipped at the end of the scaled image), then symmetric constant padding up to the crop size (transforms.py:36-40,61-73)
    const int got_h = min(crop_h, sh - pos_h), got_w = min(crop_w, sw - pos_w);
    const int pad_t = max(crop_h - got_h, 0) / 2, pad_l = max(crop_w - got_w, 0) / 2;
    const int yy = y - pad_t, xx = x - pad_l;
    const long plane = (long)crop_h * crop_w, o = (long)b * 3 * plane + (long)y * crop_w + x;
    if (yy < 0 || yy >= got_h || xx < 0 || xx >= got_w) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) { out_rgb[o + ch * plane] = 0.f; out_modal[o + ch * plane] = 0.f; }
      out_label[(long)b * plane + (long)y * crop_w + x] = 255;
      continue;
    }
    const int cy = pos_h + yy, cx = pos_w + xx;
    const int area2x = (H == 2 * sh && W == 2 * sw);
    const Coef ky = coef_y(cy, H, sh), kx = coef_x(cx, W, sw);
    int v[3];
    resized_rgb(rgb + (long)b * H * W * 3, H, W, sh, sw, cy, cx, flip, area2x, ky, kx, v);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out_rgb[o + ch * plane] = lut_rgb[ch * 256 + v[ch]];
    resized_rgb(modal + (long)b * H * W * 3, H, W, sh, sw, cy, cx, flip, area2x, ky, kx, v);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out_modal[o + ch * plane] = lut_modal[ch * 256 + v[ch]];
    // label: INTER_NEAREST, src = min(floor(d * src / dst), src - 1)
    const int ly = min((int)floor((double)cy * ((double)H / (double)sh)), H - 1);
    const int lx = min((int)floor((double)cx * ((double)W / (double)sw)), W - 1);
    out_label[(long)b * plane + (long)y * crop_w + x] = label[((long)b * H + ly) * W + (flip ? W - 1 - lx : lx)];
  }
}

}  // namespace

extern "C" int dfb200_train_pre(const void* rgb, const void* modal, const void* label, int B, int H, int W, const int* params, const float* lut_rgb,
                                const float* lut_modal, int crop_h, int crop_w, float* out_rgb, float* out_modal, int64_t* out_label, void* stream) {
  DFB_REQUIRE(B > 0 && H > 0 && W > 0 && crop_h > 0 && crop_w > 0, "train_pre: empty batch");
  const long n = (long)B * crop_h * crop_w;
  long g = (n + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  dfb_launch(train_pre_kernel, (unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream), (const uint8_t*)rgb, (const uint8_t*)modal, (const uint8_t*)label, B,
             H, W, params, lut_rgb, lut_modal, crop_h, crop_w, out_rgb, out_modal, out_label);
  return dfb_check_launch("train_pre");
}
