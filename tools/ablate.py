"""Graph-replayed step time of DFormer-L with stages removed (depth ablation) -> cost per stage under the real regime."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder  # noqa: E402
from dformer_b200.engine import GraphedTrainStep  # noqa: E402
from dformer_b200.models.encoders import DFormer as enc  # noqa: E402
from dformer_b200.optim import FusedAdamW  # noqa: E402

B = 8
rgb, hha = torch.randn(B, 3, 480, 640, device="cuda"), torch.randn(B, 3, 480, 640, device="cuda")
lab = torch.randint(0, 40, (B, 480, 640), device="cuda")
base = None
for depths in [(3, 3, 12, 2), (0, 3, 12, 2), (3, 0, 12, 2), (3, 3, 0, 2), (3, 3, 12, 0), (0, 0, 0, 0)]:
    orig = enc.DFormer_Large
    enc.DFormer_Large = lambda pretrained=False, drop_path_rate=0.1, **kw: enc.DFormer(
        dims=[96, 192, 288, 576], mlp_ratios=[8, 8, 4, 4], depths=list(depths), num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7],
        drop_path_rate=drop_path_rate, **kw)
    cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
    torch.manual_seed(0)
    m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
    enc.DFormer_Large = orig
    opt = FusedAdamW(m)
    run = GraphedTrainStep(m, opt, rgb, hha, lab, warmup=2)
    for _ in range(2):
        run.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    base = base or ms
    print(f"depths {depths}: {ms:.2f} ms  (delta vs full {base - ms:.2f} ms)", flush=True)
    del run, m, opt
    torch.cuda.empty_cache()
