"""Tile-width sweep of the tcgen05 GEMM on the latency-bound shapes of a DFormer-L step (tuning aid for pick_bn in
csrc/gemm_tc.cu): forces BN through DFB200_TC_BN and times 20 back-to-back launches inside a CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as K  # noqa: E402

SHAPES = [(9600, 288, 288, 0, 1), (9600, 288, 288, 0, 0), (9600, 1152, 288, 0, 1), (9600, 288, 1152, 0, 1), (9600, 720, 288, 0, 1),
          (9600, 432, 576, 0, 1), (9600, 576, 432, 0, 0), (9600, 144, 144, 0, 1), (9600, 576, 144, 0, 1), (9600, 144, 576, 0, 1),
          (2400, 576, 576, 0, 1), (2400, 2304, 576, 0, 1), (2400, 576, 2304, 0, 1), (38400, 512, 512, 0, 1), (38400, 192, 192, 0, 1),
          (38400, 480, 192, 0, 1)]
s = torch.cuda.Stream()
for M, N, Kd, ta, tb in SHAPES:
    a = torch.randn(M, Kd, device="cuda").bfloat16()
    b = (torch.randn(N, Kd, device="cuda") if tb else torch.randn(Kd, N, device="cuda")).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for bn in [0] + [x for x in (32, 48, 64, 80, 96, 112, 128, 144, 160, 192, 208, 240, 256) if x <= ((N + 15) // 16) * 16]:
        if bn:
            os.environ["DFB200_TC_BN"] = str(bn)
        else:
            os.environ.pop("DFB200_TC_BN", None)
        kw = dict(trans_a=False, trans_b=bool(tb), backend=K.TCGEN05, out=out)
        for _ in range(3):
            K.gemm(a, b, **kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    K.gemm(a, b, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        res.append((bn, e0.elapsed_time(e1) / 20 * 1e3))
    best = min(res[1:], key=lambda r: r[1])
    print(f"M={M} N={N} K={Kd} tb={tb}: picker {res[0][1]:.2f} us | best BN={best[0]} {best[1]:.2f} us | " + " ".join(f"{bn}:{t:.1f}" for bn, t in res[1:]), flush=True)
