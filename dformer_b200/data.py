"""Training input pipeline on the GPU.

`TrainPre` mirrors the reference class of the same name (utils/dataloader/dataloader.py:38-73): same constructor arguments, same
sequence of `random` calls per sample (mirror, scale, crop position), same arithmetic -- cv2's uint8 bilinear / nearest resize,
float64 normalisation, crop + constant padding, HWC -> CHW -- but for a whole batch of device-resident uint8 images in ONE kernel
launch (`csrc/data.cu`), writing the fp32 NCHW tensors the model consumes.  The CPU never touches a pixel."""
import random
from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import lib
from .kernels import _chk, _s


def _lut(mean: Sequence[float], std: Sequence[float], device) -> torch.Tensor:
    """((v / 255 - mean) / std) for v = 0..255 per channel, evaluated in float64 like utils/transforms.py:182-187, then fp32."""
    v = np.arange(256, dtype=np.float64)[None, :] / 255.0
    t = (v - np.asarray(mean, dtype=np.float64)[:, None]) / np.asarray(std, dtype=np.float64)[:, None]
    return torch.from_numpy(t.astype(np.float32)).contiguous().to(device)


class TrainPre:
    def __init__(self, norm_mean, norm_std, sign: bool = False, config=None):
        self.config, self.norm_mean, self.norm_std, self.sign = config, norm_mean, norm_std, sign
        self._luts = {}

    def draw(self, H: int, W: int):
        """The reference's random draws for one sample, in its order (dataloader.py:21,29; transforms.py:53-57)."""
        cfg = self.config
        flip = random.random() >= 0.5
        sh, sw = H, W
        if cfg.train_scale_array is not None:
            scale = random.choice(cfg.train_scale_array)
            sh, sw = int(H * scale), int(W * scale)
        ch, cw = cfg.image_height, cfg.image_width
        pos_h = random.randint(0, sh - ch + 1) if sh > ch else 0
        pos_w = random.randint(0, sw - cw + 1) if sw > cw else 0
        return [int(flip), sh, sw, pos_h, pos_w]

    def __call__(self, rgb: torch.Tensor, gt: torch.Tensor, modal_x: torch.Tensor, params: Optional[Sequence[Sequence[int]]] = None):
        """rgb / modal_x [B,H,W,3] uint8, gt [B,H,W] uint8 (CUDA) -> rgb [B,3,h,w] fp32, gt [B,h,w] int64, modal_x [B,3,h,w] fp32."""
        _chk(rgb, "rgb")
        assert rgb.dtype == torch.uint8 and modal_x.dtype == torch.uint8 and gt.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[-1] == 3
        B, H, W, _ = rgb.shape
        dev = rgb.device
        if params is None:
            params = [self.draw(H, W) for _ in range(B)]
        ptab = torch.tensor(params, dtype=torch.int32).to(dev, non_blocking=True)
        if dev not in self._luts:
            mm, ms = ([0.48] * 3, [0.28] * 3) if self.sign else (self.norm_mean, self.norm_std)          # dataloader.py:55-60
            self._luts[dev] = (_lut(self.norm_mean, self.norm_std, dev), _lut(mm, ms, dev))
        lut_rgb, lut_modal = self._luts[dev]
        ch, cw = self.config.image_height, self.config.image_width
        out_rgb = torch.empty((B, 3, ch, cw), device=dev, dtype=torch.float32)
        out_modal = torch.empty((B, 3, ch, cw), device=dev, dtype=torch.float32)
        out_gt = torch.empty((B, ch, cw), device=dev, dtype=torch.int64)
        rgb, gt, modal_x = rgb.contiguous(), gt.contiguous(), modal_x.contiguous()
        lib().train_pre(rgb.data_ptr(), modal_x.data_ptr(), gt.data_ptr(), B, H, W, ptab.data_ptr(), lut_rgb.data_ptr(), lut_modal.data_ptr(), ch, cw,
                        out_rgb.data_ptr(), out_modal.data_ptr(), out_gt.data_ptr(), _s())
        return out_rgb, out_gt, out_modal


class ValPre:
    """utils/dataloader/dataloader.py:112-122 for a device-resident uint8 batch: normalisation (the depth image always with 0.48 / 0.28,
    as the reference does) and HWC -> CHW; same kernel as TrainPre with the identity geometry."""

    def __init__(self, norm_mean, norm_std, sign: bool = False, config=None):
        self.config, self.norm_mean, self.norm_std, self.sign = config, norm_mean, norm_std, sign
        self._luts = {}

    def __call__(self, rgb: torch.Tensor, gt: torch.Tensor, modal_x: torch.Tensor):
        _chk(rgb, "rgb")
        assert rgb.dtype == torch.uint8 and modal_x.dtype == torch.uint8 and gt.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[-1] == 3
        B, H, W, _ = rgb.shape
        dev = rgb.device
        if dev not in self._luts:
            self._luts[dev] = (_lut(self.norm_mean, self.norm_std, dev), _lut([0.48] * 3, [0.28] * 3, dev))
        lut_rgb, lut_modal = self._luts[dev]
        ptab = torch.tensor([[0, H, W, 0, 0]] * B, dtype=torch.int32).to(dev, non_blocking=True)
        out_rgb = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
        out_modal = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
        out_gt = torch.empty((B, H, W), device=dev, dtype=torch.int64)
        rgb, gt, modal_x = rgb.contiguous(), gt.contiguous(), modal_x.contiguous()
        lib().train_pre(rgb.data_ptr(), modal_x.data_ptr(), gt.data_ptr(), B, H, W, ptab.data_ptr(), lut_rgb.data_ptr(), lut_modal.data_ptr(), H, W,
                        out_rgb.data_ptr(), out_modal.data_ptr(), out_gt.data_ptr(), _s())
        return out_rgb, out_gt, out_modal
