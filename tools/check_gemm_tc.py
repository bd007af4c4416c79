"""Stand-alone parity check of the tcgen05 GEMM (run in its own process: a protocol bug traps the
kernel and poisons the CUDA context).  Prints one line per case and exits non-zero on mismatch."""
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as k  # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    cases = []
    # (M, N, K, trans_a, trans_b, out_dtype, act, splitk)
    for M, N, K in [(128, 64, 64), (256, 96, 96), (300, 48, 48), (4800, 192, 192), (1000, 288, 1152), (19200, 768, 96),
                    (777, 40, 512), (512, 2304, 576), (200, 144, 72), (4800, 512, 1056)]:
        cases.append((M, N, K, False, True, torch.bfloat16, 1, 1))       # forward: A K-major, W K-major
    for M, N, K in [(256, 64, 128), (300, 96, 240), (4800, 192, 480), (1000, 576, 288), (130, 48, 40)]:
        cases.append((M, N, K, False, False, torch.bfloat16, 0, 1))      # dgrad: dY K-major, W MN-major
    for M, N, K in [(64, 64, 512), (96, 96, 3000), (240, 96, 4800), (768, 96, 19200), (40, 512, 1234), (288, 1152, 2400)]:
        cases.append((M, N, K, True, False, torch.float32, 0, 0))        # wgrad: both MN-major, split-K auto
    cases.append((1000, 200, 320, False, True, torch.float32, 0, 1))
    for (M, N, K, ta, tb, od, act, sk) in cases:
        a = (torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)).bfloat16()
        b = (torch.randn(N, K, device=dev) if tb else torch.randn(K, N, device=dev)).bfloat16()
        bias = torch.randn(N, device=dev)
        out = k.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, backend=k.TCGEN05, out_dtype=od, act=act, act_col_start=N // 2 // 8 * 8, splitk=sk)
        torch.cuda.synchronize()
        A = a.float().t() if ta else a.float()
        Bm = b.float().t() if tb else b.float()
        ref = A @ Bm + bias
        if act:
            c0 = N // 2 // 8 * 8
            ref[:, c0:] = torch.nn.functional.gelu(ref[:, c0:])
        err = (out.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        lim = (2e-2 if od == torch.bfloat16 else 2e-3) * max(scale, 1.0)
        good = err <= lim and bool(torch.isfinite(out.float()).all())
        ok &= good
        print(f"{'OK ' if good else 'BAD'} M={M} N={N} K={K} ta={int(ta)} tb={int(tb)} out={od} act={act} err={err:.4g} lim={lim:.4g}", flush=True)
    # strided C / A slices (column slice of a wider buffer), accumulate
    buf = torch.randn(1000, 3 * 96, device=dev).bfloat16()
    w = torch.randn(64, 96, device=dev).bfloat16()
    outbuf = torch.zeros(1000, 256, device=dev, dtype=torch.bfloat16)
    k.gemm(buf[:, 96:192], w, trans_b=True, backend=k.TCGEN05, out=outbuf[:, 64:128])
    ref = buf[:, 96:192].float() @ w.float().t()
    err = (outbuf[:, 64:128].float() - ref).abs().max().item()
    good = err < 0.3 and outbuf[:, :64].abs().max().item() == 0 and outbuf[:, 128:].abs().max().item() == 0
    ok &= good
    print(f"{'OK ' if good else 'BAD'} strided slices err={err:.4g}")
    acc = torch.ones(96, 64, device=dev)
    x = torch.randn(5000, 64, device=dev).bfloat16()
    dy = torch.randn(5000, 96, device=dev).bfloat16()
    k.gemm(dy, x, trans_a=True, trans_b=False, backend=k.TCGEN05, out=acc, accumulate=True)
    ref = dy.float().t() @ x.float() + 1
    err = (acc - ref).abs().max().item()
    good = err < 0.05
    ok &= good
    print(f"{'OK ' if good else 'BAD'} wgrad accumulate err={err:.4g}")
    # TMA-store epilogue: ragged N / M tails, relu + bias, column slices at 64-column and odd (16-byte aligned) offsets, several n-tiles
    for (M, N, K, off, width) in [(1000, 720, 288, 0, 720), (333, 432, 576, 64, 512), (4800, 144, 144, 8, 160), (130, 48, 40, 48, 96),
                                  (9600, 288, 1152, 0, 288), (257, 1152, 288, 128, 1408), (19200, 40, 512, 0, 40)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.2).bfloat16()
        bias = torch.randn(N, device=dev)
        outbuf = torch.full((M, width), 7.0, device=dev, dtype=torch.bfloat16)
        k.gemm(a, w, trans_b=True, bias=bias, backend=k.TCGEN05, out=outbuf[:, off:off + N], act=2)
        ref = torch.relu(a.float() @ w.float().t() + bias)
        err = (outbuf[:, off:off + N].float() - ref).abs().max().item()
        untouched = (outbuf[:, :off] == 7).all().item() and (outbuf[:, off + N:] == 7).all().item()
        good = err <= 2e-2 * max(1.0, ref.abs().max().item()) and untouched
        ok &= good
        print(f"{'OK ' if good else 'BAD'} tma-store slice M={M} N={N} K={K} off={off} width={width} err={err:.4g} untouched={untouched}", flush=True)
    # fused "gate" epilogue (q * a, cut * e written into column slices of the concat buffer; optional un-gated second output)
    for (M, N, K, ycols, off, keep) in [(1000, 96, 96, 144, 0, True), (1000, 48, 48, 144, 96, True), (4800, 192, 192, 384, 0, False),
                                        (1200, 144, 144, 576, 432, True), (300, 288, 288, 1152, 864, False), (130, 40, 72, 200, 160, True)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.2).bfloat16()
        bias = torch.randn(N, device=dev)
        qbuf = torch.randn(M, 3 * N + 8, device=dev).bfloat16()
        gate = qbuf[:, N:2 * N] if N % 8 == 0 else qbuf[:, :N]
        y = torch.full((M, ycols), 7.0, device=dev, dtype=torch.bfloat16)
        plain = torch.full((M, N + 8), 7.0, device=dev, dtype=torch.bfloat16) if keep else None
        k.gemm(a, w, trans_b=True, bias=bias, backend=k.TCGEN05, out=y[:, off:off + N], gate=gate, out2=plain[:, :N] if keep else None)
        ref = a.float() @ w.float().t() + bias
        refy = ref * gate.float()
        e1 = (y[:, off:off + N].float() - refy).abs().max().item()
        e2 = (plain[:, :N].float() - ref).abs().max().item() if keep else 0.0
        untouched = (y[:, :off] == 7).all().item() and (y[:, off + N:] == 7).all().item() and (not keep or (plain[:, N:] == 7).all().item())
        good = e1 <= 2e-2 * max(1.0, refy.abs().max().item()) and e2 <= 2e-2 * max(1.0, ref.abs().max().item()) and untouched
        ok &= good
        print(f"{'OK ' if good else 'BAD'} gate M={M} N={N} K={K} off={off} keep={keep} err={e1:.4g} err_plain={e2:.4g} untouched={untouched}", flush=True)
    # strided-batched bf16 output (NMF products)
    for (Bz, M, N, K, ta, tb) in [(3, 4800, 64, 512, False, False), (2, 512, 64, 4800, True, False), (3, 300, 512, 64, False, True)]:
        a = (torch.randn(Bz, K, M, device=dev) if ta else torch.randn(Bz, M, K, device=dev)).bfloat16()
        b = (torch.randn(Bz, N, K, device=dev) if tb else torch.randn(Bz, K, N, device=dev)).bfloat16()
        for od in (torch.bfloat16, torch.float32):
            out = torch.empty(Bz, M, N, device=dev, dtype=od)
            k.bgemm(a, b, out, trans_a=ta, trans_b=tb, M=M, N=N, K=K)
            A = a.float().transpose(1, 2) if ta else a.float()
            Bm = b.float().transpose(1, 2) if tb else b.float()
            ref = A @ Bm
            err = (out.float() - ref).abs().max().item()
            good = err <= (2e-2 if od == torch.bfloat16 else 2e-3) * max(1.0, ref.abs().max().item())
            ok &= good
            print(f"{'OK ' if good else 'BAD'} batched Bz={Bz} M={M} N={N} K={K} ta={int(ta)} tb={int(tb)} out={od} err={err:.4g}", flush=True)
    # timing of a few representative shapes
    for (M, N, K) in [(153600, 768, 96), (153600, 96, 768), (38400, 1536, 192), (38400, 192, 1536), (9600, 1152, 288), (38400, 512, 1056), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            k.gemm(a, b, trans_b=True, backend=k.TCGEN05, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            k.gemm(a, b, trans_b=True, backend=k.TCGEN05, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            torch.matmul(a, b.t(), out=out)
        t1.record()
        torch.cuda.synchronize()
        ms_t = t0.elapsed_time(t1) / 10
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K + M * N)
        print(f"time M={M} N={N} K={K}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s | cuBLAS {ms_t:.4f} ms {fl / ms_t / 1e9:.1f} TFLOP/s")
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
