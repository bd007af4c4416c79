"""CPU restatement (numpy) of the reference's training-time preprocessing -- TEST INFRASTRUCTURE, never imported by the product.

Follows utils/dataloader/dataloader.py:20-73 (random_mirror, random_scale, TrainPre.__call__) and utils/transforms.py:27-75,182-187
(random_crop_pad_to_shape, generate_random_crop_pos, pad_image_to_shape, normalize).  The arithmetic that lives in a third-party
dependency absent from /root/reference is OpenCV's (opencv-python, unpinned in the reference; 4.13.0 in the build container):
  * cv2.resize(uint8, INTER_LINEAR): 11-bit fixed-point separable bilinear (resize.cpp: HResizeLinear / VResizeLinear<uchar>):
      fx = float((dx + 0.5) * (src / dst) - 0.5), sx = floor(fx), fx -= sx; x direction: sx < 0 -> (0, fx = 0), sx >= W-1 -> (W-1, fx = 0);
      y direction: indices clipped, fy kept;  alpha = cvRound(w * 2048) (round-half-even);
      row = S[sx] * a0 + S[sx+1] * a1;   dst = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2
      exactly-2x down-scaling in both directions is rerouted to INTER_AREA: (a + b + c + d + 2) >> 2;  equal sizes: copy.
  * cv2.resize(INTER_NEAREST): src index = min(floor(d * src / dst), src - 1).
  * cv2.flip(img, 1): mirror along W.  cv2.copyMakeBorder(BORDER_CONSTANT).
Pinned against the unmodified reference run with real cv2: oracle/make_golden_trainpre.py -> tests/golden/trainpre.pt
(tests/test_oracle_golden.py::test_trainpre_oracle_matches_reference_golden)."""
import random

import numpy as np


def _coef_x(n_src, n_dst):
    scale = n_src / n_dst
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo, hi = s < 0, s >= n_src - 1
    f[lo] = 0
    s[lo] = 0
    f[hi] = 0
    s[hi] = n_src - 1
    return s, np.minimum(s + 1, n_src - 1), f


def _coef_y(n_src, n_dst):
    scale = n_src / n_dst
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1), f


def resize_linear_u8(img, dw, dh):
    H, W = img.shape[:2]
    if (dw, dh) == (W, H):
        return img.copy()
    if W == 2 * dw and H == 2 * dh:
        i = img.astype(np.int32)
        return ((i[0::2, 0::2] + i[0::2, 1::2] + i[1::2, 0::2] + i[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, sx1, fx = _coef_x(W, dw)
    sy, sy1, fy = _coef_y(H, dh)
    one = np.float32(1)
    a0, a1 = np.rint((one - fx) * np.float32(2048)).astype(np.int64), np.rint(fx * np.float32(2048)).astype(np.int64)
    b0, b1 = np.rint((one - fy) * np.float32(2048)).astype(np.int64), np.rint(fy * np.float32(2048)).astype(np.int64)
    ii = img.astype(np.int64).reshape(H, W, -1)
    hb = ii[:, sx] * a0[None, :, None] + ii[:, sx1] * a1[None, :, None]
    out = (((b0[:, None, None] * (hb[sy] >> 4)) >> 16) + ((b1[:, None, None] * (hb[sy1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8).reshape((dh, dw) + img.shape[2:])


def resize_nearest(img, dw, dh):
    H, W = img.shape[:2]
    ys = np.minimum(np.floor(np.arange(dh) * (H / dh)).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(dw) * (W / dw)).astype(np.int64), W - 1)
    return img[ys][:, xs]


def draw_params(H, W, scales, crop_size):
    """The reference's `random` calls in its order (dataloader.py:21,29; transforms.py:53-57): mirror, scale, crop position."""
    flip = random.random() >= 0.5
    scale = random.choice(scales) if scales is not None else 1
    sh, sw = (int(H * scale), int(W * scale)) if scales is not None else (H, W)
    crop_h, crop_w = crop_size
    pos_h = random.randint(0, sh - crop_h + 1) if sh > crop_h else 0
    pos_w = random.randint(0, sw - crop_w + 1) if sw > crop_w else 0
    return flip, sh, sw, pos_h, pos_w


def _crop_pad(img, pos, crop_size, value):
    ch, cw = crop_size
    c = img[pos[0]:pos[0] + ch, pos[1]:pos[1] + cw, ...]
    ph, pw = max(ch - c.shape[0], 0), max(cw - c.shape[1], 0)
    pad = ((ph // 2, ph // 2 + ph % 2), (pw // 2, pw // 2 + pw % 2)) + ((0, 0),) * (img.ndim - 2)
    return np.pad(c, pad, mode="constant", constant_values=value)


def train_pre(rgb, gt, modal_x, norm_mean, norm_std, scales, crop_size, sign=False):
    """TrainPre.__call__ (dataloader.py:46-73): uint8 HWC rgb / modal_x, uint8 HW gt -> (3,ch,cw) float64, (ch,cw) uint8, (3,ch,cw) float64."""
    H, W = rgb.shape[:2]
    flip, sh, sw, pos_h, pos_w = draw_params(H, W, scales, crop_size)
    if flip:
        rgb, gt, modal_x = rgb[:, ::-1], gt[:, ::-1], modal_x[:, ::-1]
    if scales is not None:
        rgb, gt, modal_x = resize_linear_u8(rgb, sw, sh), resize_nearest(gt, sw, sh), resize_linear_u8(modal_x, sw, sh)
    norm = lambda im, m, s: (im.astype(np.float64) / 255.0 - m) / s
    rgb = norm(rgb, np.asarray(norm_mean), np.asarray(norm_std))
    mm, ms = ((np.array([0.48] * 3), np.array([0.28] * 3)) if sign else (np.asarray(norm_mean), np.asarray(norm_std)))
    modal_x = norm(modal_x, mm, ms)
    p_rgb = _crop_pad(rgb, (pos_h, pos_w), crop_size, 0)
    p_gt = _crop_pad(gt, (pos_h, pos_w), crop_size, 255)
    p_modal = _crop_pad(modal_x, (pos_h, pos_w), crop_size, 0)
    return p_rgb.transpose(2, 0, 1), p_gt, p_modal.transpose(2, 0, 1)
