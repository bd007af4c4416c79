// Multi-scale + flip evaluation and the mIoU confusion matrix (utils/val_mm.py:257-470, utils/metrics_new.py:6-47), SURVEY
// section 8f row N3.  The reference makes four full passes over the [B, ncls, H, W] probability volume per scale and flip
// (interpolate, flip, softmax, +=) and three more for the metric (argmax, mask, bincount); here it is
//   resize_nchw_ac      : input rescale, bilinear align_corners=True (optionally mirrored along W)
//   ms_softmax_accum    : acc += softmax_c( resize_ac( [flip_W] logits ) )       one read of the low-res logits, one RMW of acc
//   argmax_confusion    : hist[target * n + argmax_c acc] += 1 over non-ignored pixels (shared-memory histogram per CTA)
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

// ATen's align_corners=True source index: src = dst * (in - 1) / (out - 1)
struct LerpAC { int i0, i1; float w1; };
__device__ __forceinline__ LerpAC lerp_ac(int o, int n_in, float scale) {
  LerpAC l;
  const float src = scale * (float)o;
  l.i0 = min((int)src, n_in - 1);
  l.i1 = l.i0 + (l.i0 < n_in - 1 ? 1 : 0);
  l.w1 = src - (float)l.i0;
  return l;
}
__host__ __device__ __forceinline__ float ac_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f; }

__global__ void resize_nchw_ac_kernel(const float* __restrict__ in, int B, int C, int Hi, int Wi, float* __restrict__ out, int Ho, int Wo, int flip) {
  pdl_sync();
  const long n = (long)B * C * Ho * Wo;
  const float sy = ac_scale(Hi, Ho), sx = ac_scale(Wi, Wo);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
    const long plane = i / ((long)Wo * Ho);
    const LerpAC ly = lerp_ac(y, Hi, sy), lx = lerp_ac(flip ? Wo - 1 - x : x, Wi, sx);      // flip(resize(img)) == resize sampled at mirrored x
    const float* p = in + plane * Hi * Wi;
    const float v00 = p[(long)ly.i0 * Wi + lx.i0], v01 = p[(long)ly.i0 * Wi + lx.i1];
    const float v10 = p[(long)ly.i1 * Wi + lx.i0], v11 = p[(long)ly.i1 * Wi + lx.i1];
    const float w0x = 1.f - lx.w1, w0y = 1.f - ly.w1;
    out[i] = w0y * (w0x * v00 + lx.w1 * v01) + ly.w1 * (w0x * v10 + lx.w1 * v11);
  }
}

// thread = one full-resolution pixel; classes are walked twice (max, then exp / sum) from the L2-resident low-res logits
__global__ void ms_softmax_accum_kernel(const float* __restrict__ logits, int B, int C, int h, int w, float* __restrict__ acc, int H, int W, int flip) {
  pdl_sync();
  const long n = (long)B * H * W;
  const float sy = ac_scale(h, H), sx = ac_scale(w, W);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long)W * H));
    const LerpAC ly = lerp_ac(y, h, sy), lx = lerp_ac(x, w, sx);
    // logits of a mirrored input are mirrored back before the resize: column j of the un-flipped map is column w-1-j here
    const int c0 = flip ? w - 1 - lx.i0 : lx.i0, c1 = flip ? w - 1 - lx.i1 : lx.i1;
    const float w0x = 1.f - lx.w1, w0y = 1.f - ly.w1;
    const float* base = logits + (long)b * C * h * w;
    const long o00 = (long)ly.i0 * w + c0, o01 = (long)ly.i0 * w + c1, o10 = (long)ly.i1 * w + c0, o11 = (long)ly.i1 * w + c1;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float* p = base + (long)c * h * w;
      const float v = w0y * (w0x * p[o00] + lx.w1 * p[o01]) + ly.w1 * (w0x * p[o10] + lx.w1 * p[o11]);
      m = fmaxf(m, v);
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* p = base + (long)c * h * w;
      const float v = w0y * (w0x * p[o00] + lx.w1 * p[o01]) + ly.w1 * (w0x * p[o10] + lx.w1 * p[o11]);
      s += __expf(v - m);
    }
    const float inv = 1.f / s;
    float* a = acc + (long)b * C * H * W + (long)y * W + x;
    for (int c = 0; c < C; ++c) {
      const float* p = base + (long)c * h * w;
      const float v = w0y * (w0x * p[o00] + lx.w1 * p[o01]) + ly.w1 * (w0x * p[o10] + lx.w1 * p[o11]);
      a[(long)c * H * W] += __expf(v - m) * inv;
    }
  }
}

__global__ void __launch_bounds__(256) argmax_confusion_kernel(const float* __restrict__ score, const int64_t* __restrict__ target, int B, int C, long HW,
                                                               int ignore, float* __restrict__ hist, int64_t* __restrict__ pred_out) {
  pdl_sync();
  extern __shared__ unsigned int sh[];                    // [C * C] per-CTA histogram
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long n = (long)B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long b = i / HW, p = i % HW;
    const float* s = score + b * C * HW + p;
    float best = s[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = s[(long)c * HW];
      if (v > best) { best = v; arg = c; }                 // first maximum wins, like torch.argmax
    }
    if (pred_out) pred_out[i] = arg;
    const long t = target ? target[i] : (long)ignore;
    if (target && t != (long)ignore && t >= 0 && t < C) atomicAdd(&sh[(int)t * C + arg], 1u);
  }
  __syncthreads();
  if (hist)
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (sh[i]) atomicAdd(hist + i, (float)sh[i]);
}

inline int grid_for(long n) {
  long b = (n + 255) / 256;
  if (b < 1) b = 1;
  const long cap = 148L * 16;
  return (int)(b > cap ? cap : b);
}

}  // namespace

#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int dfb200_resize_nchw_ac(const float* in, int B, int C, int Hi, int Wi, float* out, int Ho, int Wo, int flip, void* stream) {
  DFB_REQUIRE(B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "resize_nchw_ac: empty tensor");
  dfb_launch(resize_nchw_ac_kernel, grid_for((long)B * C * Ho * Wo), 256, 0, ST, in, B, C, Hi, Wi, out, Ho, Wo, flip);
  return dfb_check_launch("resize_nchw_ac");
}

extern "C" int dfb200_ms_softmax_accum(const float* logits, int B, int C, int h, int w, float* acc, int H, int W, int flip, void* stream) {
  DFB_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "ms_softmax_accum: empty tensor");
  dfb_launch(ms_softmax_accum_kernel, grid_for((long)B * H * W), 256, 0, ST, logits, B, C, h, w, acc, H, W, flip);
  return dfb_check_launch("ms_softmax_accum");
}

extern "C" int dfb200_argmax_confusion(const float* score, const int64_t* target, int B, int C, long HW, int ignore, float* hist, int64_t* pred,
                                       void* stream) {
  DFB_REQUIRE(B > 0 && C > 0 && HW > 0, "argmax_confusion: empty tensor");
  DFB_REQUIRE((size_t)C * C * 4 <= 48 * 1024, "argmax_confusion: %d classes exceed the shared-memory histogram", C);
  dfb_launch(argmax_confusion_kernel, grid_for((long)B * HW), 256, (size_t)C * C * 4, ST, score, target, B, C, HW, ignore, hist, pred);
  return dfb_check_launch("argmax_confusion");
}
