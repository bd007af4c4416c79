"""Kernel timeline of the CUDA-graph-replayed DFormer-L inference forward (eval mode) at batch B: chronological kernel list
(start us, duration us, stream, name) + summary.  usage: python tools/profile_infer.py [B] [csv path]"""
import os
import sys
from collections import defaultdict
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.1, aux_rate=0.0,
                      device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16")
torch.manual_seed(0)
m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().eval()
rgb, hha = torch.rand(B, 3, 480, 640, device="cuda"), torch.rand(B, 3, 480, 640, device="cuda")
m.decode_head.injected_bases = torch.rand(B, 512, 64, device="cuda")
with torch.no_grad():
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            m(rgb, hha)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = m(rgb, hha)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay {e0.elapsed_time(e1) / 50:.3f} ms at batch {B}")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
agg = defaultdict(lambda: [0.0, 0])
rows = []
for ev in evs:
    nm = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
    rows.append((ev.time_range.start - t0, ev.time_range.end - ev.time_range.start, nm, getattr(ev, "device_resource_id", -1)))
    agg[nm][0] += ev.time_range.end - ev.time_range.start
    agg[nm][1] += 1
print(f"{len(rows)} kernels, span {rows[-1][0] + rows[-1][1]:.1f} us, sum of durations {sum(r[1] for r in rows):.1f} us")
for nm, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{t:9.1f} us  x{n:<4d} {nm}")
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        for s_, d_, nm, sid in rows:
            fh.write(f"{s_:.2f},{d_:.2f},{sid},{nm}\n")
