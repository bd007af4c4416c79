"""Inference latency / throughput of DFormer-L (BASELINE config 5; protocol of the reference's utils/latency.py:29-63:
rand inputs, warm-up then CUDA-event timed repetitions) -- but in eval mode, and both with eager launches and as a
replayed CUDA graph.  usage: python tools/latency.py [precision] [batches...]"""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import EncoderDecoder  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
batches = [int(b) for b in sys.argv[2:]] or [1, 8, 32]
cfg = SimpleNamespace(backbone="DFormer-Large", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.1, aux_rate=0.0,
                      device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=prec)
torch.manual_seed(0)
m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().eval()
res = []
for B in batches:
    rgb, hha = torch.rand(B, 3, 480, 640, device="cuda"), torch.rand(B, 3, 480, 640, device="cuda")
    m.decode_head.injected_bases = torch.rand(B, 512, 64, device="cuda")
    warm, reps = (100, 300) if B == 1 else (10, 30)
    with torch.no_grad():
        for _ in range(warm):
            m(rgb, hha)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m(rgb, hha)
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / reps
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(rgb, hha)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = m(rgb, hha)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / reps
    r = {"batch": B, "precision": prec, "eager_ms": round(eager, 3), "graph_ms": round(graph, 3), "graph_img_per_s": round(B / graph * 1e3, 1)}
    res.append(r)
    print(json.dumps(r), flush=True)
    del g, out
