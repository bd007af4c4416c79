"""Per-kernel device-time table of one training step (torch.profiler/CUPTI; debug aid, not a bench number).
usage: python tools/profile_step.py [variant] [batch] [precision] [H] [W]"""
import os
import sys
from collections import defaultdict
from types import SimpleNamespace

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dformer_b200 import EncoderDecoder  # noqa: E402
from dformer_b200.optim import FusedAdamW  # noqa: E402


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "DFormer-Large"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 480
    W = int(sys.argv[5]) if len(sys.argv) > 5 else 640
    cfg = SimpleNamespace(backbone=variant, decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=prec, return_logits=False)
    torch.manual_seed(0)
    m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
    opt = FusedAdamW(m)
    rgb, hha = torch.randn(B, 3, H, W, device="cuda"), torch.randn(B, 3, H, W, device="cuda")
    lab = torch.randint(0, 40, (B, H, W), device="cuda")

    def step():
        loss, _ = m(rgb, hha, lab)
        loss.backward()
        opt.step()
        opt.zero_grad()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"step: {e0.elapsed_time(e1) / 3:.2f} ms device, {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms wall, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = name.split("(")[0][:90]
            agg[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"total device kernel time {tot / 1e3:.2f} ms over {sum(v[1] for v in agg.values())} launches")
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{n:<5d} {name}")
    # per-shape GEMM table (CUDA events around every launch)
    from dformer_b200 import kernels as K
    K.GEMM_PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
    g = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for e0, e1, fl, by, tc, shape in prof:
        a = g[(tc,) + shape]
        a[0] += e0.elapsed_time(e1); a[1] += 1; a[2] += fl; a[3] += by
    tt = sum(a[0] for a in g.values())
    print(f"GEMM launches via gemm(): {len(prof)}, total {tt:.2f} ms (event-timed, includes launch gaps)")
    print("   ms     n   TFLOP/s   GB/s   tc (M, N, K, ta, tb)")
    for k, a in sorted(g.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{a[0]:7.3f} {a[1]:4d} {a[2] / a[0] / 1e9:8.1f} {a[3] / a[0] / 1e6:7.0f}   {int(k[0])} {k[1:]}")


if __name__ == "__main__":
    main()
