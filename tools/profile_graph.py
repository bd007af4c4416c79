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
DEPTHS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
if DEPTHS:
    from dformer_b200.models.encoders import DFormer as enc
    enc.DFormer_Large = lambda pretrained=False, drop_path_rate=0.1, **kw: enc.DFormer(
        dims=[96, 192, 288, 576], mlp_ratios=[8, 8, 4, 4], depths=DEPTHS, num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7],
        drop_path_rate=drop_path_rate, **kw)
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
# timeline analysis: busy union, idle gaps, concurrency
iv = sorted((ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA)
t0, t1 = iv[0][0], max(e for _, e, _ in iv)
if os.environ.get("DFB200_TIMELINE_CSV"):            # chronological kernel list (start us, duration us, stream, name) for offline analysis
    with open(os.environ["DFB200_TIMELINE_CSV"], "w") as fh:
        for ev in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start):
            nm = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
            fh.write(f"{ev.time_range.start - t0:.2f},{ev.time_range.end - ev.time_range.start:.2f},{getattr(ev, 'device_resource_id', -1)},{nm}\n")
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s_, e_, n_ in iv[1:]:
    if s_ > cur_e:
        busy += cur_e - cur_s
        gaps.append((s_ - cur_e, n_))
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print(f"timeline: span {(t1 - t0) / 1e3:.2f} ms, busy (union) {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms in {len(gaps)} gaps")
gaps.sort(reverse=True)
print("largest gaps (us, next kernel):", [(round(g, 1), n.replace('(anonymous namespace)::', '').replace('void ', '')[:40]) for g, n in gaps[:8]])
# time with exactly one kernel running
pts = sorted([(s_, 1) for s_, e_, _ in iv] + [(e_, -1) for s_, e_, _ in iv])
lvl, last, hist = 0, pts[0][0], defaultdict(float)
for t_, d_ in pts:
    hist[lvl] += t_ - last
    last, lvl = t_, lvl + d_
print("time by number of concurrently running kernels (ms):", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
# which kernels run ALONE (exclusive time per kernel name)
import heapq
ev2 = sorted([(s_, 0, i) for i, (s_, e_, _) in enumerate(iv)] + [(e_, 1, i) for i, (s_, e_, _) in enumerate(iv)])
active, last_t, alone = set(), ev2[0][0], defaultdict(float)
for t_, kind, i in ev2:
    if len(active) == 1:
        alone[iv[next(iter(active))][2]] += t_ - last_t
    last_t = t_
    if kind == 0:
        active.add(i)
    else:
        active.discard(i)
al = defaultdict(float)
for n_, v_ in alone.items():
    al[n_.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]] += v_
print("exclusive (alone-on-GPU) time by kernel, ms:")
for n_, v_ in sorted(al.items(), key=lambda kv: -kv[1])[:22]:
    print(f"   {v_ / 1e3:7.3f}  {n_}")
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
