"""Stand-alone timing of the fused MLP-middle kernels (csrc/mlp_dw.cu) against the unfused chain they replace, on the
DFormer-L batch-8 shapes.  Algorithmic bytes: forward 2 passes over [M, C] bf16, backward 3 passes.
usage: python tools/bench_mlp_dw.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dformer_b200 import kernels as K  # noqa: E402

SHAPES = [(8, 120, 160, 768), (8, 120, 160, 384), (8, 60, 80, 1536), (8, 60, 80, 768), (8, 30, 40, 1152), (8, 30, 40, 576), (8, 15, 20, 2304),
          (8, 15, 20, 1152)]


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


def main():
    tot = {"fwd_old": 0.0, "fwd_new": 0.0, "bwd_old": 0.0, "bwd_new": 0.0, "fwd_gp": 0.0, "bwd_gp": 0.0}
    counts = {0: 3, 1: 3, 2: 3, 3: 3, 4: 12, 5: 12, 6: 2, 7: 1}
    only = int(sys.argv[1]) if len(sys.argv) > 1 else None       # profile mode: one shape, new kernels only
    for i, (B, H, W, C) in enumerate(SHAPES):
        if only is not None and i != only:
            continue
        M = B * H * W
        h = torch.randn(M, C, device="cuda").bfloat16()
        du = torch.randn(M, C, device="cuda").bfloat16()
        w, b = torch.randn(C, 1, 3, 3, device="cuda") * 0.2, torch.randn(C, device="cuda") * 0.1
        dw, db, dc = torch.zeros_like(w), torch.zeros_like(b), torch.zeros(C, device="cuda")
        zs = {}

        def fwd_old():
            zs["u"], zs["z"] = K.dwconv_fwd(h, w, b, B, H, W, 3, True, K.ACT_GELU, save_z=True)

        def fwd_new():
            K.mlp_dw_fwd(h, w, b, B, H, W)

        fwd_old()
        def bwd_old():
            K.dwconv_bwd(du, h, w, b, B, H, W, 3, True, K.ACT_GELU, dw, db, z=zs["z"])
            K.colsum(du, out=dc)

        def bwd_new():
            K.mlp_dw_bwd(du, h, w, b, B, H, W, dw, db, dc)

        _, gp = K.mlp_dw_fwd(h, w, b, B, H, W, save_gp=True)

        def fwd_gp():
            K.mlp_dw_fwd(h, w, b, B, H, W, save_gp=True)

        def bwd_gp():
            K.mlp_dw_bwd(du, h, w, b, B, H, W, dw, db, dc, gp=gp)

        if only is not None:
            x7 = torch.randn(M, 96, device="cuda").bfloat16()
            w7, b7 = torch.randn(96, 1, 7, 7, device="cuda") * 0.1, torch.randn(96, device="cuda") * 0.1
            dw7, db7 = torch.zeros_like(w7), torch.zeros_like(b7)
            for _ in range(2):
                fwd_gp()
                bwd_gp()
                K.dwconv_fwd(x7, w7, b7, B, H, W, 7)
                K.dwconv_bwd(x7, x7, w7, b7, B, H, W, 7, False, 0, dw7, db7)
            torch.cuda.synchronize()
            return

        t = {n: timeit(f) for n, f in (("fwd_old", fwd_old), ("fwd_new", fwd_new), ("bwd_old", bwd_old), ("bwd_new", bwd_new), ("fwd_gp", fwd_gp),
                                       ("bwd_gp", bwd_gp))}
        by = M * C * 2
        print(f"B{B} {H}x{W} C={C}: fwd {t['fwd_old']:7.1f} -> {t['fwd_new']:7.1f} us ({2 * by / t['fwd_new'] / 1e3:5.0f} GB/s)   "
              f"bwd {t['bwd_old']:7.1f} -> {t['bwd_new']:7.1f} us ({3 * by / t['bwd_new'] / 1e3:5.0f} GB/s)   "
              f"keep GELU': fwd {t['fwd_gp']:7.1f} us ({3 * by / t['fwd_gp'] / 1e3:5.0f} GB/s) bwd {t['bwd_gp']:7.1f} us ({4 * by / t['bwd_gp'] / 1e3:5.0f} GB/s)", flush=True)
        for n in tot:
            tot[n] += t[n] * counts[i]
    print("per DFormer-L step (us): " + "  ".join(f"{n} {v:8.0f}" for n, v in tot.items()))


if __name__ == "__main__":
    main()
