"""Kernel table of the CUDA-graph-replayed training step (torch.profiler / CUPTI)."""
import os
import sys
from collections import defaultdict
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder  # noqa: E402
from dformer_b200.engine import GraphedTrainStep  # noqa: E402
from dformer_b200.optim import FusedAdamW  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0,
                      device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
torch.manual_seed(0)
m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
opt = FusedAdamW(m)
rgb, hha = torch.randn(B, 3, 480, 640, device="cuda"), torch.randn(B, 3, 480, 640, device="cuda")
lab = torch.randint(0, 40, (B, 480, 640), device="cuda")
run = GraphedTrainStep(m, opt, rgb, hha, lab, warmup=2)
for _ in range(3):
    run.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run.step()
e1.record()
torch.cuda.synchronize()
print(f"graph step {e0.elapsed_time(e1) / 5:.2f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run.step()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
tmin, tmax = 1e30, 0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:80]
        agg[name][0] += ev.device_time_total
        agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"sum of kernel durations {tot / 1e3:.2f} ms over {sum(v[1] for v in agg.values())} kernels (overlapping streams)")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{n:<5d} {name}")
