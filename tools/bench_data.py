"""Throughput of the GPU input pipeline (dformer_b200.data.TrainPre) on NYUDepthv2-shaped batches: device-resident uint8 in,
fp32 NCHW out.  Algorithmic bytes per image: 2 x 3 x crop fp32 + crop int64 written, <= 7 bytes/pixel of uint8 read."""
import os
import random
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200.data import TrainPre  # noqa: E402

B, H, W = 8, 480, 640
cfg = SimpleNamespace(train_scale_array=[0.5, 0.75, 1, 1.25, 1.5, 1.75], image_height=480, image_width=640)
pre = TrainPre([0.485, 0.456, 0.406], [0.229, 0.224, 0.225], config=cfg)
rgb = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
modal = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
gt = torch.randint(0, 41, (B, H, W), dtype=torch.uint8, device="cuda")
random.seed(0)
for worst in (False, True):
    params = [[1, 840, 1120, 100, 200]] * B if worst else None          # worst case: mirrored 1.75x up-scaling everywhere
    for _ in range(3):
        pre(rgb, gt, modal, params=params)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pre(rgb, gt, modal, params=params)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    wr = B * (2 * 3 * 4 + 8) * 480 * 640
    print(f"{'all 1.75x + mirror' if worst else 'random params':20s}: {ms * 1e3:7.1f} us per batch of {B} ({B / ms * 1e3:9.0f} img/s, {wr / ms / 1e6:6.0f} GB/s written)")
