"""Collect the GEMM shapes of one DFormer-L training step and replay each shape back-to-back inside a CUDA graph:
gives the stand-alone per-shape cost (us) x count = GEMM time of a step without launch gaps.
usage: python tools/gemm_replay.py [collect|replay] (replay reads tools/gemm_shapes_L8.json)"""
import json
import os
import sys
from collections import Counter
from types import SimpleNamespace

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dformer_b200 import EncoderDecoder, kernels as K  # noqa: E402

PATH = os.path.join(ROOT, "tools", "gemm_shapes_L8.json")


def collect():
    cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.15, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
    torch.manual_seed(0)
    m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
    rgb, hha = torch.randn(8, 3, 480, 640, device="cuda"), torch.randn(8, 3, 480, 640, device="cuda")
    lab = torch.randint(0, 40, (8, 480, 640), device="cuda")
    loss, _ = m(rgb, hha, lab)
    loss.backward()
    K.GEMM_PROFILE = []
    loss, _ = m(rgb, hha, lab)
    loss.backward()
    torch.cuda.synchronize()
    c = Counter((tc,) + shape for _, _, _, _, tc, shape in K.GEMM_PROFILE)
    K.GEMM_PROFILE = None
    json.dump([[list(k), v] for k, v in c.items()], open(PATH, "w"))
    print("shapes:", len(c), "launches:", sum(c.values()))


def replay():
    shapes = json.load(open(PATH))
    dev = "cuda"
    rows = []
    s = torch.cuda.Stream()
    for (tc, M, N, Kd, ta, tb), cnt in shapes:
        if not tc:
            continue
        a = (torch.randn(Kd, M, device=dev) if ta else torch.randn(M, Kd, device=dev)).bfloat16()
        b = (torch.randn(N, Kd, device=dev) if tb else torch.randn(Kd, N, device=dev)).bfloat16()
        od = torch.float32 if ta else torch.bfloat16
        out = torch.zeros(M, N, device=dev, dtype=od)
        kw = dict(trans_a=bool(ta), trans_b=bool(tb), backend=K.TCGEN05, out=out, accumulate=bool(ta))
        for _ in range(3):
            K.gemm(a, b, **kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        reps = 50
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    K.gemm(a, b, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        by = 2 * (M * Kd + N * Kd) + out.element_size() * M * N
        rows.append((us * cnt, us, cnt, (M, N, Kd, ta, tb), 2.0 * M * N * Kd / us / 1e6, by / us / 1e3))
    rows.sort(reverse=True)
    print(f"total stand-alone GEMM time per step: {sum(r[0] for r in rows) / 1e3:.2f} ms over {sum(r[2] for r in rows)} launches")
    for tot, us, cnt, shp, tf, gb in rows[:40]:
        print(f"{tot:8.1f} us = {cnt:3d} x {us:7.2f} us  {tf:7.1f} TFLOP/s {gb:6.0f} GB/s  {shp}")


if __name__ == "__main__":
    (collect if (len(sys.argv) > 1 and sys.argv[1] == "collect") else replay)()
