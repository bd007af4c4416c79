"""Multi-scale + flip evaluation and segmentation metrics on the GPU.

Mirrors `utils/val_mm.py:evaluate_msf` (:325-470), `slide_inference` (:257-321) and `utils/metrics_new.py:Metrics` of the reference:
same scale rounding (multiples of 32), `align_corners=True` resampling, horizontal flip, softmax accumulation, argmax,
ignore-label masking and the confusion-matrix statistics -- with the element-wise passes fused into three kernels
(`csrc/eval.cu`).  The model call in the middle is the ordinary `EncoderDecoder.forward(rgb, modal_x)`."""
import math
from typing import Iterable, List, Sequence, Tuple

import torch

from . import kernels as K


class Metrics:
    """utils/metrics_new.py:6-47 with the histogram kept on the device and updated by one fused kernel."""

    def __init__(self, num_classes: int, ignore_label: int, device) -> None:
        self.ignore_label = ignore_label
        self.num_classes = num_classes
        self.hist = torch.zeros(num_classes, num_classes, device=device, dtype=torch.float32)
        self.index = 0

    def update_hist(self, hist: torch.Tensor) -> None:
        self.hist += hist.to(self.hist.device)

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        """pred: [B, ncls, H, W] scores (argmax is taken here, metrics_new.py:20); target: [B, H, W] int64."""
        self.index += 1
        K.argmax_confusion(pred.float().contiguous(), target.to(torch.int64), self.ignore_label, hist=self.hist)

    @staticmethod
    def _finish(v: torch.Tensor) -> Tuple[List[float], float]:
        v = v.clone()
        v[v.isnan()] = 0.0
        mean = v.mean().item()
        return (v * 100).cpu().numpy().round(2).tolist(), round(mean * 100, 2)

    def compute_iou(self):
        d = self.hist.diag()
        return self._finish(d / (self.hist.sum(0) + self.hist.sum(1) - d))

    def compute_f1(self):
        return self._finish(2 * self.hist.diag() / (self.hist.sum(0) + self.hist.sum(1)))

    def compute_pixel_acc(self):
        return self._finish(self.hist.diag() / self.hist.sum(1))


def scaled_size(H: int, W: int, scale: float) -> Tuple[int, int]:
    """val_mm.py:361-365: int(scale * size) rounded up to a multiple of 32."""
    nh, nw = int(scale * H), int(scale * W)
    return int(math.ceil(nh / 32)) * 32, int(math.ceil(nw / 32)) * 32


def slide_windows(h_img: int, w_img: int, crop: Tuple[int, int], stride_rate: float) -> List[Tuple[int, int, int, int]]:
    """Window corners (y1, y2, x1, x2) in the reference's visiting order (val_mm.py:287-305): stride = int(rate * crop), the last
    row / column of windows is shifted back inside the image, windows larger than the image are clipped to it."""
    h_crop, w_crop = crop
    h_stride, w_stride = int(stride_rate * h_crop), int(stride_rate * w_crop)
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    wins = []
    for h_idx in range(h_grids):
        for w_idx in range(w_grids):
            y2, x2 = min(h_idx * h_stride + h_crop, h_img), min(w_idx * w_stride + w_crop, w_img)
            wins.append((max(y2 - h_crop, 0), y2, max(x2 - w_crop, 0), x2))
    return wins


def slide_counts(h_img: int, w_img: int, wins: Sequence[Tuple[int, int, int, int]]) -> Tuple[List[int], List[int]]:
    """The reference's `count_mat` (val_mm.py:290,317) is separable because the windows form a grid: count[y, x] = rows[y] * cols[x]."""
    rows, cols = [0] * h_img, [0] * w_img
    for y1, y2 in sorted({(w[0], w[1]) for w in wins}):
        for y in range(y1, y2):
            rows[y] += 1
    for x1, x2 in sorted({(w[2], w[3]) for w in wins}):
        for x in range(x1, x2):
            cols[x] += 1
    return rows, cols


@torch.no_grad()
def slide_inference(model, imgs: torch.Tensor, modal_xs: torch.Tensor, config) -> torch.Tensor:
    """`slide_inference` (val_mm.py:257-321): overlapping `config.eval_crop_size` windows at `config.eval_stride_rate`, logits summed
    per pixel in the reference's window order and divided by the number of windows covering the pixel.  Inputs smaller than the
    crop are first resampled up to it (align_corners=True, :279-286).  The per-window model call is the hot path; the window
    bookkeeping around it is index arithmetic on views."""
    h_crop, w_crop = config.eval_crop_size
    if h_crop > imgs.shape[-2] or w_crop > imgs.shape[-1]:
        imgs, modal_xs = K.resize_nchw_ac(imgs, h_crop, w_crop), K.resize_nchw_ac(modal_xs, h_crop, w_crop)
    if imgs.shape[-2:] != modal_xs.shape[-2:]:
        raise ValueError(f"rgb {tuple(imgs.shape[-2:])} and modal_x {tuple(modal_xs.shape[-2:])} differ in size")
    B, _, h_img, w_img = imgs.shape
    wins = slide_windows(h_img, w_img, (h_crop, w_crop), config.eval_stride_rate)
    rows, cols = slide_counts(h_img, w_img, wins)
    if min(rows) == 0 or min(cols) == 0:
        raise ValueError("sliding windows do not cover the image (eval_stride_rate > 1?)")          # the reference asserts (:318)
    preds = None
    for y1, y2, x1, x2 in wins:
        logit = model(imgs[:, :, y1:y2, x1:x2].contiguous(), modal_xs[:, :, y1:y2, x1:x2].contiguous())
        if preds is None:
            preds = torch.zeros((B, logit.shape[1], h_img, w_img), device=imgs.device, dtype=torch.float32)
        preds[:, :, y1:y2, x1:x2] += logit
    count = torch.tensor(rows, device=imgs.device, dtype=torch.float32)[:, None] * torch.tensor(cols, device=imgs.device, dtype=torch.float32)[None, :]
    return preds / count


@torch.no_grad()
def multi_scale_predict(model, rgb: torch.Tensor, modal_x: torch.Tensor, scales: Sequence[float] = (1.0,), flip: bool = False,
                        sliding_config=None) -> torch.Tensor:
    """Sum over scales (and mirrored copies) of the class probabilities at the input resolution: [B, ncls, H, W] fp32
    (`scaled_logits` of val_mm.py:357-399).  `sliding_config` (with `eval_crop_size`, `eval_stride_rate`) selects the reference's
    `sliding=True` branch (:372-375, :387-390)."""
    B, _, H, W = rgb.shape
    acc = None
    for s in scales:
        nh, nw = scaled_size(H, W, s)
        for mirrored in ((False, True) if flip else (False,)):
            r = K.resize_nchw_ac(rgb, nh, nw, flip=mirrored)
            m = K.resize_nchw_ac(modal_x, nh, nw, flip=mirrored)
            logits = model(r, m) if sliding_config is None else slide_inference(model, r, m, sliding_config)
            if acc is None:
                acc = torch.zeros((B, logits.shape[1], H, W), device=rgb.device, dtype=torch.float32)
            K.ms_softmax_accum(logits, acc, flip=mirrored)
    return acc


@torch.no_grad()
def evaluate(model, batches: Iterable, num_classes: int, ignore_label: int = 255, scales: Sequence[float] = (1.0,), flip: bool = False,
             group=None, sliding_config=None) -> Metrics:
    """`evaluate_msf` over an iterable of dicts with keys rgb / modal_x / gt (the reference's loader format); with a process
    group the confusion matrices of all ranks are summed (val_mm.py:431-436)."""
    was_training = model.training
    model.eval()
    metrics = None
    for batch in batches:
        rgb, modal_x, gt = batch["rgb"].cuda(non_blocking=True), batch["modal_x"].cuda(non_blocking=True), batch["gt"].cuda(non_blocking=True)
        if metrics is None:
            metrics = Metrics(num_classes, ignore_label, rgb.device)
        probs = multi_scale_predict(model, rgb, modal_x, scales, flip, sliding_config)
        metrics.update(probs, gt)
    if metrics is not None and group is not None and torch.distributed.is_initialized():
        torch.distributed.all_reduce(metrics.hist, group=group)
    model.train(was_training)
    return metrics
