"""Multi-scale + flip evaluation and segmentation metrics on the GPU.

Mirrors `utils/val_mm.py:evaluate_msf` (:340-412, the non-sliding branch) and `utils/metrics_new.py:Metrics` of the reference:
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


@torch.no_grad()
def multi_scale_predict(model, rgb: torch.Tensor, modal_x: torch.Tensor, scales: Sequence[float] = (1.0,), flip: bool = False) -> torch.Tensor:
    """Sum over scales (and mirrored copies) of the class probabilities at the input resolution: [B, ncls, H, W] fp32
    (`scaled_logits` of val_mm.py:357-399)."""
    B, _, H, W = rgb.shape
    acc = None
    for s in scales:
        nh, nw = scaled_size(H, W, s)
        for mirrored in ((False, True) if flip else (False,)):
            r = K.resize_nchw_ac(rgb, nh, nw, flip=mirrored)
            m = K.resize_nchw_ac(modal_x, nh, nw, flip=mirrored)
            logits = model(r, m)
            if acc is None:
                acc = torch.zeros((B, logits.shape[1], H, W), device=rgb.device, dtype=torch.float32)
            K.ms_softmax_accum(logits, acc, flip=mirrored)
    return acc


@torch.no_grad()
def evaluate(model, batches: Iterable, num_classes: int, ignore_label: int = 255, scales: Sequence[float] = (1.0,), flip: bool = False,
             group=None) -> Metrics:
    """`evaluate_msf` over an iterable of dicts with keys rgb / modal_x / gt (the reference's loader format); with a process
    group the confusion matrices of all ranks are summed (val_mm.py:431-436)."""
    was_training = model.training
    model.eval()
    metrics = None
    for batch in batches:
        rgb, modal_x, gt = batch["rgb"].cuda(non_blocking=True), batch["modal_x"].cuda(non_blocking=True), batch["gt"].cuda(non_blocking=True)
        if metrics is None:
            metrics = Metrics(num_classes, ignore_label, rgb.device)
        probs = multi_scale_predict(model, rgb, modal_x, scales, flip)
        metrics.update(probs, gt)
    if metrics is not None and group is not None and torch.distributed.is_initialized():
        torch.distributed.all_reduce(metrics.hist, group=group)
    model.train(was_training)
    return metrics
