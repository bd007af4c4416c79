"""One warm + N profiled training steps of the bench workload (for ncu launch lists)."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder  # noqa: E402
from dformer_b200.optim import FusedAdamW  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "DFormer-Large"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = SimpleNamespace(backbone=variant, decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0, device="cuda",
                      pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
torch.manual_seed(0)
m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
opt = FusedAdamW(m)
rgb, hha = torch.randn(B, 3, 480, 640, device="cuda"), torch.randn(B, 3, 480, 640, device="cuda")
lab = torch.randint(0, 40, (B, 480, 640), device="cuda")
for _ in range(steps):
    loss, _ = m(rgb, hha, lab)
    loss.backward()
    opt.step()
    opt.zero_grad()
torch.cuda.synchronize()
print("loss", loss.item())
