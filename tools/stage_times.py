"""Per-module device time of one eager training step (CUDA events around every autograd.Function fwd/bwd)."""
import os
import sys
from collections import defaultdict
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder, functions as Fn  # noqa: E402
from dformer_b200.optim import FusedAdamW  # noqa: E402

REC = []


def wrap(cls, name_of):
    f0, b0 = cls.forward, cls.backward

    def fwd(ctx, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f0(ctx, *a)
        e1.record()
        REC.append((name_of(a) + " fwd", e0, e1))
        ctx._nm = name_of(a)
        return r

    def bwd(ctx, *g):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = b0(ctx, *g)
        e1.record()
        REC.append((ctx._nm + " bwd", e0, e1))
        return r
    cls.forward, cls.backward = staticmethod(fwd), staticmethod(bwd)


wrap(Fn.BlockFn, lambda a: "block stage" + a[2].prefix.split(".")[1])
wrap(Fn.StemFn, lambda a: "stem")
wrap(Fn.DownsampleFn, lambda a: "downsample")
wrap(Fn.HeadFn, lambda a: "head")
wrap(Fn.UpsampleCEFn, lambda a: "upsample+CE")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0,
                      device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
torch.manual_seed(0)
m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
opt = FusedAdamW(m)
rgb, hha = torch.randn(B, 3, 480, 640, device="cuda"), torch.randn(B, 3, 480, 640, device="cuda")
lab = torch.randint(0, 40, (B, 480, 640), device="cuda")
for it in range(3):
    REC.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    loss, _ = m(rgb, hha, lab)
    loss.backward()
    opt.step()
    opt.zero_grad()
    t1.record()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
for n, e0, e1 in REC:
    agg[n][0] += e0.elapsed_time(e1)
    agg[n][1] += 1
print(f"eager step {t0.elapsed_time(t1):.2f} ms; sum of module times {sum(v[0] for v in agg.values()):.2f} ms")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:8.3f} ms  x{c:<3d} ({t / c:6.3f} each)  {n}")
