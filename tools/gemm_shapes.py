"""Runs a few representative tcgen05 GEMM shapes of DFormer-L (for ncu capture / timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as k  # noqa: E402

SHAPES = [  # (M, N, K, trans_a, trans_b, out_dtype) -- fc1 stage0, fc2 stage0, qcl stage1, wgrad fc1 stage0, dgrad fc2 stage0
    (153600, 768, 96, False, True, torch.bfloat16),
    (153600, 96, 768, False, True, torch.bfloat16),
    (38400, 480, 192, False, True, torch.bfloat16),
    (768, 96, 153600, True, False, torch.float32),
    (153600, 768, 96, False, False, torch.bfloat16),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = "cuda"
    for (M, N, K, ta, tb, od) in SHAPES:
        a = (torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)).bfloat16()
        b = (torch.randn(N, K, device=dev) if tb else torch.randn(K, N, device=dev)).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=od)
        for _ in range(2):
            k.gemm(a, b, trans_a=ta, trans_b=tb, backend=k.TCGEN05, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            k.gemm(a, b, trans_a=ta, trans_b=tb, backend=k.TCGEN05, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = a.numel() * 2 + b.numel() * 2 + out.numel() * out.element_size()
        print(f"M={M} N={N} K={K} ta={int(ta)} tb={int(tb)}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s", flush=True)


def tiny(reps=200):
    dev = "cuda"
    cases = [(128, 64, 64, 0, 1), (128, 64, 448, 0, 1), (128, 64, 1792, 0, 1), (128, 64, 448, 1, 0), (144, 432, 392, 1, 0), (144, 432, 392, 0, 1),
             (128, 256, 448, 0, 1), (9600, 144, 144, 0, 1), (9600, 288, 288, 0, 1), (2400, 576, 576, 0, 1)]
    for (M, N, K, ta, tb) in cases:
        a = (torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)).bfloat16()
        b = (torch.randn(N, K, device=dev) if tb else torch.randn(K, N, device=dev)).bfloat16()
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if ta else torch.bfloat16)
        _g = k.gemm
        def gemm_call(a=a, b=b, out=out, ta=ta, tb=tb):
            _g(a, b, trans_a=bool(ta), trans_b=bool(tb), backend=k.TCGEN05, out=out, accumulate=bool(ta))
        for _ in range(5):
            gemm_call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    gemm_call()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"tiny M={M} N={N} K={K} ta={ta} tb={tb}: {e0.elapsed_time(e1) / reps * 1e3:.2f} us per back-to-back launch (CUDA graph of {reps})", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "tiny":
        tiny()
        sys.exit(0)
    main()
