"""Device time of single kernels at the DFormer-L batch-8 stage shapes: each launcher is captured `reps` times back to back in a
CUDA graph and the replay is timed with CUDA events (no host launch overhead in the figure).
usage: python tools/ktime.py [family ...]   families: gaa ln dw7 mlp_dw elem pool bn gate loss"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as K  # noqa: E402

DEV = "cuda"
B = 8
STAGES = [(120, 160, 96, 1, 0), (60, 80, 192, 2, 1), (30, 40, 288, 4, 1), (15, 20, 576, 8, 1)]      # H, W, C, heads, window


def gtime(fn, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def rb(*s):
    return torch.randn(*s, device=DEV).bfloat16()


def fam_gaa():
    for H, W, C, heads, win in STAGES:
        if not win:
            continue
        HW, Ce = H * W, C // 2
        d = Ce // heads
        m, kv, dout = rb(B * 49, Ce), rb(B * HW, 2 * Ce), torch.randn(B * 49, Ce, device=DEV)
        out, lse = K.gaa_fused_fwd(m, kv, B, HW, heads, d)
        print(f"gaa HW={HW} heads={heads} d={d}: fwd {gtime(lambda: K.gaa_fused_fwd(m, kv, B, HW, heads, d)):.1f} us  "
              f"bwd {gtime(lambda: K.gaa_fused_bwd(dout, out, lse, m, kv, B, HW, heads, d)):.1f} us", flush=True)


def fam_ln():
    for H, W, C, _, _ in STAGES:
        for c in (C, C // 2):
            M = B * H * W
            x, g, b = torch.randn(M, c, device=DEV), torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
            y, mu, rs = K.layernorm_fwd(x, g, b, 1e-6, torch.bfloat16)
            dy, dx1 = rb(M, c), torch.randn(M, c, device=DEV)
            dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
            tf = gtime(lambda: K.layernorm_fwd(x, g, b, 1e-6, torch.bfloat16))
            tb = gtime(lambda: K.layernorm_bwd(dy, x, g, mu, rs, dx1, dg, db))
            tb2 = gtime(lambda: K.layernorm_bwd(dy, x, g, mu, rs, dx1, dg, db, dy2=dy))
            byf, byb = M * c * 6, M * c * 14
            print(f"ln M={M} C={c}: fwd {tf:.1f} us ({byf / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({byb / tb / 1e3:.0f} GB/s)  bwd+dy2 {tb2:.1f} us", flush=True)


def fam_dw7():
    for H, W, C, _, _ in STAGES:
        for c in (C, C // 2):
            x = rb(B * H * W, c)
            w, b = torch.randn(c, 1, 7, 7, device=DEV), torch.randn(c, device=DEV)
            dw, dbb = torch.zeros_like(w), torch.zeros_like(b)
            tf = gtime(lambda: K.dwconv_fwd(x, w, b, B, H, W, 7))
            tb = gtime(lambda: K.dwconv_bwd(x, x, w, b, B, H, W, 7, False, K.ACT_NONE, dw, dbb))
            print(f"dw7 [{B},{H},{W},{c}]: fwd {tf:.1f} us  bwd(dgrad+wgrad) {tb:.1f} us", flush=True)


def fam_mlp_dw():
    for (H, W, C, _, _), r in zip(STAGES, (8, 8, 4, 4)):
        for c in (C * r, C * r // 2):
            h = rb(B * H * W, c)
            w, b = torch.randn(c, 1, 3, 3, device=DEV), torch.randn(c, device=DEV)
            dw, dbb, dcs = torch.zeros_like(w), torch.zeros_like(b), torch.zeros_like(b)
            u, gp = K.mlp_dw_fwd(h, w, b, B, H, W, save_gp=True)
            tf = gtime(lambda: K.mlp_dw_fwd(h, w, b, B, H, W, save_gp=True))
            tb = gtime(lambda: K.mlp_dw_bwd(u, h, w, b, B, H, W, dw, dbb, dcs, gp=gp))
            n = B * H * W * c * 2
            print(f"mlp_dw [{B},{H},{W},{c}]: fwd {tf:.1f} us ({3 * n / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({4 * n / tb / 1e3:.0f} GB/s)", flush=True)


def fam_elem():
    for H, W, C, _, _ in STAGES:
        M, HW = B * H * W, H * W
        x, y = torch.randn(M, C, device=DEV), rb(M, C)
        ls, sb = torch.randn(C, device=DEV), torch.ones(B, device=DEV)
        dls, dcs = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        a, b2, o = rb(M, C), rb(M, C), rb(M, C)
        t1 = gtime(lambda: K.scale_residual_fwd(x, y, ls, sb, HW))
        t2 = gtime(lambda: K.scale_residual_bwd(x, y, ls, sb, HW, dls, dy_colsum=dcs))
        t3 = gtime(lambda: K.mul_fwd(a, b2, o))
        t4 = gtime(lambda: K.act_fwd(a, K.ACT_GELU))
        print(f"elem M={M} C={C}: scale_res fwd {t1:.1f} bwd {t2:.1f}  mul_fwd {t3:.1f}  act_fwd {t4:.1f} us", flush=True)


def fam_pool():
    for H, W, C, _, win in STAGES:
        if not win:
            continue
        M, Ce = B * H * W, C // 2
        xn, en = rb(M, C), rb(M, Ce)
        y = rb(M, 2 * C)
        o7 = torch.randn(B * 49, Ce, device=DEV)
        do7 = torch.empty(B * 49, Ce, device=DEV)
        dp = rb(B * 49, C + Ce)
        t1 = gtime(lambda: K.pool7_fwd(xn, en, B, H, W))
        t2 = gtime(lambda: K.pool7_bwd(dp, C, Ce, B, H, W))
        t3 = gtime(lambda: K.resize_fwd(o7, B, 7, 7, y, H, W, col0=C))
        t4 = gtime(lambda: K.resize_bwd(y, C, B, 7, 7, Ce, H, W, do7))
        print(f"pool/resize [{B},{H},{W}] C={C}: pool7 fwd {t1:.1f} bwd {t2:.1f}  resize fwd {t3:.1f} bwd {t4:.1f} us", flush=True)


def fam_bn():
    """BatchNorm streaming kernels at the head (M = 38400, C = 512), stem (M = 614400 / 153600) and downsample shapes"""
    for (M, C, xdt) in [(38400, 512, torch.bfloat16), (614400, 48, torch.bfloat16), (153600, 96, torch.bfloat16), (153600, 96, torch.float32),
                        (38400, 192, torch.float32), (9600, 288, torch.float32)]:
        x = torch.randn(M, C, device=DEV).to(xdt)
        dy, res = rb(M, C), rb(M, C)
        g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        st = K.bn_stats(x)
        ms = K.bn_finalize(st, M, 1e-5, 0.1, rm, rv)
        gbuf, sums = K.bn_bwd_reduce(dy, x, ms, g, b, res, K.ACT_RELU, None, M // B)
        t1 = gtime(lambda: K.bn_stats(x))
        t2 = gtime(lambda: K.bn_apply(x, ms, g, b, torch.bfloat16, residual=res, act=K.ACT_RELU))
        t3 = gtime(lambda: K.bn_bwd_reduce(dy, x, ms, g, b, res, K.ACT_RELU, None, M // B))
        t3b = gtime(lambda: K.bn_bwd_reduce(dy, x, ms, g, b, None, K.ACT_NONE, None, M // B))
        t4 = gtime(lambda: K.bn_bwd_apply(gbuf, x, ms, g, sums, M, True, torch.bfloat16))
        es = x.element_size()
        n = M * C
        print(f"bn M={M} C={C} x={'bf16' if es == 2 else 'fp32'}: stats {t1:.1f} us ({n * es / t1 / 1e3:.0f} GB/s)  apply+res+relu {t2:.1f} ({n * (es + 4) / t2 / 1e3:.0f})  "
              f"bwd_reduce+res+relu {t3:.1f} ({n * (es + 6) / t3 / 1e3:.0f})  bwd_reduce plain {t3b:.1f} ({n * (es + 4) / t3b / 1e3:.0f})  "
              f"bwd_apply {t4:.1f} ({n * (es + 4) / t4 / 1e3:.0f}) [incl. the small memset / alloc of each call]", flush=True)


def fam_loss():
    """x8 upsample + CE at the bench shape: three-launch path (forward, rows adjoint, columns adjoint) vs the one-pass training kernel"""
    Bq, h, w, ncls, H, W = 8, 60, 80, 40, 480, 640
    small = rb(Bq * h * w, ncls)
    label = torch.randint(0, ncls, (Bq, H, W), device=DEV)
    label[torch.rand(Bq, H, W, device=DEV) < 0.05] = 255
    one = torch.ones((), device=DEV)
    _, lse, acc, _, _ = K.upsample_ce_fwd(small, Bq, h, w, ncls, H, W, label, 255, want_out=False)
    t1 = gtime(lambda: K.upsample_ce_fwd(small, Bq, h, w, ncls, H, W, label, 255, want_out=False))
    t2 = gtime(lambda: K.upsample_ce_bwd_fused(small, Bq, h, w, ncls, H, W, label, 255, lse, acc, one))
    t3 = gtime(lambda: K.upsample_ce_train(small, Bq, h, w, ncls, H, W, label, 255))
    print(f"loss [{Bq},{ncls},{h}x{w} -> {H}x{W}]: forward {t1:.1f} us + adjoint (rows + columns) {t2:.1f} us = {t1 + t2:.1f} us  |  one-pass training kernel {t3:.1f} us "
          f"(each incl. its small fills)", flush=True)


def fam_gate():
    """the gating GEMMs (a, e_back) with the fused gate epilogue vs GEMM + mul_fwd"""
    for H, W, C, _, _ in STAGES:
        M = B * H * W
        for n in (C, C // 2):
            x, w, bias = rb(M, n), rb(n, n), torch.randn(n, device=DEV)
            q, y, a = rb(M, 5 * C // 2), rb(M, 2 * C), rb(M, n)
            gate = q[:, C:C + n]
            t1 = gtime(lambda: K.gemm(x, w, trans_b=True, bias=bias, out=y[:, :n], gate=gate, out2=a))
            t1b = gtime(lambda: K.gemm(x, w, trans_b=True, bias=bias, out=y[:, :n], gate=gate))
            t2 = gtime(lambda: K.gemm(x, w, trans_b=True, bias=bias, out=a))
            t3 = gtime(lambda: K.mul_fwd(gate, a, y[:, :n]))
            print(f"gate M={M} N=K={n}: fused (keeps a) {t1:.1f} us  fused (inference) {t1b:.1f}  |  gemm {t2:.1f} + mul {t3:.1f} = {t2 + t3:.1f} us", flush=True)


if __name__ == "__main__":
    fams = sys.argv[1:] or ["gaa", "ln", "dw7", "mlp_dw", "elem", "pool"]
    for f in fams:
        globals()["fam_" + f]()
