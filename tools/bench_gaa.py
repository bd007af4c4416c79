"""Stand-alone timing of the attention core: GEMM + softmax chain vs the fused kernels (DFormer-L batch-8 shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as K  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B = 8
SHAPES = [(4800, 2, 48), (1200, 4, 36), (300, 8, 36)]
if len(sys.argv) > 1:
    SHAPES = [SHAPES[int(sys.argv[1])]]
for HW, heads, d in SHAPES:
    Cp = heads * d
    m = torch.randn(B * 49, Cp, device="cuda").bfloat16()
    kv = torch.randn(B * HW, 2 * Cp, device="cuda").bfloat16()
    dout = torch.randn(B * 49, Cp, device="cuda")
    out, probs = K.gaa_fwd(m, kv, B, HW, heads, d)
    out2, lse = K.gaa_fused_fwd(m, kv, B, HW, heads, d)
    t = [timeit(lambda: K.gaa_fwd(m, kv, B, HW, heads, d)), timeit(lambda: K.gaa_fused_fwd(m, kv, B, HW, heads, d)),
         timeit(lambda: K.gaa_bwd(dout, m, kv, probs, B, HW, heads, d)), timeit(lambda: K.gaa_fused_bwd(dout, out2, lse, m, kv, B, HW, heads, d))]
    print(f"HW={HW} heads={heads} d={d}: fwd {t[0]:7.1f} -> {t[1]:7.1f} us   bwd {t[2]:7.1f} -> {t[3]:7.1f} us", flush=True)
