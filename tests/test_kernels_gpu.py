"""Kernel-level parity: every C-ABI launcher against the plain-PyTorch fp32 statement of the same op.

fp32 kernels: tight tolerances.  bf16 kernels: inputs are rounded to bf16 first, the reference is
computed in fp32 from the rounded inputs, and the tolerance covers one bf16 rounding of the output."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)


def K():
    from dformer_b200 import kernels
    return kernels


def tol(dtype):
    return dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


DTYPES = [torch.float32, torch.bfloat16]


# ----------------------------------------------------------------------------- GEMM (SIMT)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K_", [(300, 96, 48), (49, 40, 130), (1000, 144, 288)])
def test_gemm_simt(dtype, ta, tb, M, N, K_):
    k = K()
    a = rnd(K_, M, dtype=dtype) if ta else rnd(M, K_, dtype=dtype)
    b = rnd(N, K_, dtype=dtype) if tb else rnd(K_, N, dtype=dtype)
    bias = rnd(N)
    out = k.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, backend=k.SIMT, act=k.ACT_GELU, act_col_start=N // 2)
    A = (a.float().t() if ta else a.float())
    Bm = (b.float().t() if tb else b.float())
    ref = A @ Bm + bias
    ref[:, N // 2:] = F.gelu(ref[:, N // 2:])
    torch.testing.assert_close(out.float(), ref, **(dict(rtol=2e-2, atol=5e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-3)))


def test_gemm_simt_splitk_and_accumulate():
    k = K()
    a, b = rnd(5000, 64), rnd(5000, 96)
    out = k.gemm(a, b, trans_a=True, trans_b=False, backend=k.SIMT, splitk=8)
    ref = a.t() @ b
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=5e-3)
    out2 = k.gemm(a, b, trans_a=True, trans_b=False, backend=k.SIMT, splitk=1, out=ref.clone(), accumulate=True)
    torch.testing.assert_close(out2, 2 * ref, rtol=1e-4, atol=1e-2)


def test_bgemm_mixed_dtypes():
    k = K()
    a = rnd(3, 70, 33, dtype=torch.bfloat16)
    b = rnd(3, 33, 20)
    out = torch.empty(3, 70, 20, device=DEV)
    k.bgemm(a, b, out, M=70, N=20, K=33, alpha=0.5)
    torch.testing.assert_close(out, 0.5 * torch.bmm(a.float(), b), rtol=1e-4, atol=1e-3)


def test_colsum():
    k = K()
    for dtype in DTYPES:
        for N in (96, 40, 37, 2304):
            x = rnd(1234, N, dtype=dtype)
            torch.testing.assert_close(k.colsum(x), x.float().sum(0), rtol=1e-3, atol=1e-2)


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C", [16, 48, 96, 288, 576])
def test_layernorm(dtype, C):
    k = K()
    M = 777
    x = rnd(M, C) * 2 + 0.5
    g, b = 1 + 0.1 * rnd(C), 0.1 * rnd(C)
    y, mean, rstd = k.layernorm_fwd(x, g, b, 1e-6, dtype)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-6)
    torch.testing.assert_close(y.float(), ref, **tol(dtype))
    dy = rnd(M, C, dtype=dtype)
    ref.backward(dy.float())
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx = k.layernorm_bwd(dy, x, g, mean, rstd, torch.ones(M, C, device=DEV), dg, db)
    # fused fan-in: LN_bwd(dy_a + dy_b) in one kernel
    dya = rnd(M, C, dtype=dtype)
    dyb = (dy.float() - dya.float()).to(dtype)
    dg2, db2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx2 = k.layernorm_bwd(dya, x, g, mean, rstd, torch.ones(M, C, device=DEV), dg2, db2, dy2=dyb)
    t2 = dict(rtol=2e-2, atol=5e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dx2, dx, **t2)
    torch.testing.assert_close(dg2, dg, rtol=2e-2, atol=0.5 if dtype == torch.bfloat16 else 1e-3)
    torch.testing.assert_close(dx - 1, xr.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dg, gr.grad, rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(db, br.grad, rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C", [96, 288, 52])
def test_scale_residual_layernorm_fused(dtype, C):
    """x1 = res + dp * ls * branch and LN(x1) in one pass == the two separate kernels."""
    k = K()
    B, rows = 3, 50
    M = B * rows
    res = rnd(M, C) * 2
    wide = rnd(M, C + 24, dtype=dtype)
    branch = wide[:, 8:8 + C] if C % 8 == 0 else rnd(M, C, dtype=dtype)
    ls, dp = rnd(C) * 0.3, torch.tensor([1.2, 0.0, 1.2], device=DEV)
    g, b = 1 + 0.1 * rnd(C), 0.1 * rnd(C)
    x1, y, mean, rstd = k.scale_residual_layernorm_fwd(res, branch, ls, dp, rows, g, b, 1e-6)
    x1_ref = res + dp.repeat_interleave(rows)[:, None] * ls * branch.float()
    torch.testing.assert_close(x1, x1_ref, rtol=1e-5, atol=1e-5)
    ref = F.layer_norm(x1_ref, (C,), g, b, 1e-6)
    torch.testing.assert_close(y.float(), ref, **tol(dtype))
    torch.testing.assert_close(mean, x1_ref.mean(1), rtol=1e-4, atol=1e-4)
    if C % 8 == 0:
        x1b = k.scale_residual_fwd(res, branch, ls, dp, rows)
        yb, mb, rb = k.layernorm_fwd(x1b, g, b, 1e-6, dtype)
        torch.testing.assert_close(x1, x1b)
        torch.testing.assert_close(y, yb)


# ----------------------------------------------------------------------------- depthwise conv
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k_,C,add,act", [(3, 256, True, 1), (7, 96, False, 0), (7, 48, False, 0), (3, 72, True, 1)])
def test_dwconv(dtype, k_, C, add, act):
    k = K()
    B, H, W = 2, 15, 21
    x = rnd(B, H, W, C, dtype=dtype)
    w, b = rnd(C, 1, k_, k_, scale=0.2), rnd(C, scale=0.1)
    y = k.dwconv_fwd(x.view(-1, C), w, b, B, H, W, k_, add, act)
    xr = x.float().clone().requires_grad_(True)
    wr, brr = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.conv2d(xr.permute(0, 3, 1, 2), wr, brr, padding=k_ // 2, groups=C).permute(0, 2, 3, 1)
    if add:
        z = z + xr
    ref = F.gelu(z) if act == 1 else z
    torch.testing.assert_close(y.view(B, H, W, C).float(), ref, **tol(dtype))
    dy = rnd(B, H, W, C, dtype=dtype)
    ref.backward(dy.float())
    dw, dbias = torch.zeros_like(w), torch.zeros_like(b)
    dx = k.dwconv_bwd(dy.view(-1, C), x.view(-1, C), w, b, B, H, W, k_, add, act, dw, dbias)
    if act:      # variant with the pre-activation saved by the forward kernel
        y2, zsv = k.dwconv_fwd(x.view(-1, C), w, b, B, H, W, k_, add, act, save_z=True)
        torch.testing.assert_close(y2, y)
        torch.testing.assert_close(zsv.view(B, H, W, C).float(), z, **tol(dtype))
        dw2, db2 = torch.zeros_like(w), torch.zeros_like(b)
        dx2 = k.dwconv_bwd(dy.view(-1, C), x.view(-1, C), w, b, B, H, W, k_, add, act, dw2, db2, z=zsv)
        torch.testing.assert_close(dx2.float(), dx.float(), **(dict(rtol=3e-2, atol=5e-2) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)))
    t = dict(rtol=3e-2, atol=5e-2) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dx.view(B, H, W, C).float(), xr.grad, **t)
    torch.testing.assert_close(dw, wr.grad, rtol=3e-2 if dtype == torch.bfloat16 else 1e-3, atol=0.3 if dtype == torch.bfloat16 else 1e-2)
    torch.testing.assert_close(dbias, brr.grad, rtol=3e-2 if dtype == torch.bfloat16 else 1e-3, atol=0.3 if dtype == torch.bfloat16 else 1e-2)


@pytest.mark.parametrize("B,H,W,C", [(2, 30, 40, 144), (1, 17, 70, 64), (2, 60, 80, 96), (1, 9, 45, 48), (3, 15, 20, 288), (1, 5, 3, 40), (1, 33, 50, 192)])
def test_dw7_tma(B, H, W, C):
    """TMA-fed depthwise 7x7 (bf16): forward, data gradient and weight/bias gradient for both slab widths (48 / 64 channels),
    both tile heights, partial slabs and partial tiles."""
    k = K()
    dtype = torch.bfloat16
    x = rnd(B, H, W, C, dtype=dtype)
    w, b = rnd(C, 1, 7, 7, scale=0.1), rnd(C, scale=0.1)
    y = k.dwconv_fwd(x.view(-1, C), w, b, B, H, W, 7)
    xr = x.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=3, groups=C).permute(0, 2, 3, 1)
    torch.testing.assert_close(y.view(B, H, W, C).float(), ref, **tol(dtype))
    dy = rnd(B, H, W, C, dtype=dtype)
    ref.backward(dy.float())
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    dx = k.dwconv_bwd(dy.view(-1, C), x.view(-1, C), w, b, B, H, W, 7, False, 0, dw, db)
    torch.testing.assert_close(dx.view(B, H, W, C).float(), xr.grad, rtol=3e-2, atol=5e-2)
    n = math.sqrt(B * H * W)
    torch.testing.assert_close(dw, wr.grad, rtol=3e-2, atol=0.03 * n)
    torch.testing.assert_close(db, br.grad, rtol=3e-2, atol=0.03 * n)


@pytest.mark.parametrize("B,H,W,C", [(2, 15, 21, 256), (1, 30, 40, 64), (2, 17, 70, 72), (1, 8, 32, 128), (3, 60, 80, 192), (1, 1, 1, 8)])
def test_mlp_dw_fused(B, H, W, C):
    """TMA-fed fused MLP middle (bf16): u = GELU(dw3x3(h)+b+h) and its whole backward (dh, dW, db, colsum(dh)) against
    PyTorch autograd in fp32 on the bf16-rounded inputs; also against the unfused kernels it replaces."""
    k = K()
    dtype = torch.bfloat16
    h = rnd(B, H, W, C, dtype=dtype)
    w, b = rnd(C, 1, 3, 3, scale=0.2), rnd(C, scale=0.1)
    u = k.mlp_dw_fwd(h.view(-1, C), w, b, B, H, W)
    hr = h.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.conv2d(hr.permute(0, 3, 1, 2), wr, br, padding=1, groups=C).permute(0, 2, 3, 1) + hr
    ref = F.gelu(z)
    torch.testing.assert_close(u.view(B, H, W, C).float(), ref, **tol(dtype))
    torch.testing.assert_close(u, k.dwconv_fwd(h.view(-1, C), w, b, B, H, W, 3, True, 1), rtol=1e-2, atol=1e-2)
    du = rnd(B, H, W, C, dtype=dtype)
    ref.backward(du.float())
    dw, db, dc = torch.zeros_like(w), torch.zeros_like(b), torch.zeros(C, device=DEV)
    dh = k.mlp_dw_bwd(du.view(-1, C), h.view(-1, C), w, b, B, H, W, dw, db, dc)
    t = dict(rtol=3e-2, atol=5e-2)
    torch.testing.assert_close(dh.view(B, H, W, C).float(), hr.grad, **t)
    n = math.sqrt(B * H * W)
    torch.testing.assert_close(dw, wr.grad, rtol=3e-2, atol=0.03 * n)
    torch.testing.assert_close(db, br.grad, rtol=3e-2, atol=0.03 * n)
    torch.testing.assert_close(dc, dh.float().sum(0), rtol=1e-3, atol=1e-2 * n)
    torch.testing.assert_close(dc, hr.grad.reshape(-1, C).sum(0), rtol=3e-2, atol=0.03 * n)
    # accumulation semantics: a second call adds onto the same buffers
    k.mlp_dw_bwd(du.view(-1, C), h.view(-1, C), w, b, B, H, W, dw, db, None)
    torch.testing.assert_close(dw, 2 * wr.grad, rtol=3e-2, atol=0.06 * n)
    # training form: forward keeps GELU'(z), backward is a pure stream (no recompute)
    u2, gp = k.mlp_dw_fwd(h.view(-1, C), w, b, B, H, W, save_gp=True)
    torch.testing.assert_close(u2, u)
    zd = z.detach()
    gp_ref = 0.5 * (1 + torch.erf(zd / math.sqrt(2))) + zd * torch.exp(-0.5 * zd * zd) / math.sqrt(2 * math.pi)
    torch.testing.assert_close(gp.view(B, H, W, C).float(), gp_ref, **tol(dtype))
    dw3, db3, dc3 = torch.zeros_like(w), torch.zeros_like(b), torch.zeros(C, device=DEV)
    dh3 = k.mlp_dw_bwd(du.view(-1, C), h.view(-1, C), w, b, B, H, W, dw3, db3, dc3, gp=gp)
    torch.testing.assert_close(dh3.view(B, H, W, C).float(), hr.grad, **t)
    torch.testing.assert_close(dw3, wr.grad, rtol=3e-2, atol=0.03 * n)
    torch.testing.assert_close(db3, br.grad, rtol=3e-2, atol=0.03 * n)
    torch.testing.assert_close(dc3, dh3.float().sum(0), rtol=1e-3, atol=1e-2 * n)


# ----------------------------------------------------------------------------- gating / residual
@pytest.mark.parametrize("dtype", DTYPES)
def test_mul_and_scale_residual(dtype):
    k = K()
    M, C = 600, 96
    buf = rnd(M, 3 * C, dtype=dtype)
    a, b = buf[:, :C], buf[:, C:2 * C]
    cat = torch.zeros(M, 2 * C, device=DEV, dtype=dtype)
    k.mul_fwd(a, b, cat[:, C:])
    torch.testing.assert_close(cat[:, C:].float(), a.float() * b.float(), **tol(dtype))
    assert cat[:, :C].abs().max() == 0
    dout = rnd(M, 2 * C, dtype=dtype)
    da, db = torch.empty(M, C, device=DEV, dtype=dtype), torch.empty(M, C, device=DEV, dtype=dtype)
    k.mul_bwd(dout[:, C:], a, b, da, db)
    torch.testing.assert_close(da.float(), dout[:, C:].float() * b.float(), **tol(dtype))
    torch.testing.assert_close(db.float(), dout[:, C:].float() * a.float(), **tol(dtype))
    # column-summing form: same element-wise results + bias gradients (column sums of what was written), accumulated
    da2, db2 = torch.empty_like(da), torch.empty_like(db)
    ca, cb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    k.mul_bwd(dout[:, C:], a, b, da2, db2, ca, cb)
    torch.testing.assert_close(da2, da)
    torch.testing.assert_close(db2, db)
    torch.testing.assert_close(ca - 1, da.float().sum(0), rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(cb, db.float().sum(0), rtol=1e-3, atol=2e-2)
    # layer-scale residual with per-sample DropPath scale
    B, hw = 3, 200
    res, y, ls = rnd(M, C), rnd(M, 2 * C, dtype=dtype)[:, C:], rnd(C)
    sb = torch.tensor([0.0, 1.25, 1.25], device=DEV)
    out = k.scale_residual_fwd(res, y, ls, sb, hw)
    ref = res + sb.repeat_interleave(hw)[:, None] * ls * y.float()
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    g = rnd(M, C)
    dls = torch.zeros(C, device=DEV)
    dybuf = torch.zeros(M, 3 * C, device=DEV, dtype=dtype)
    dy = k.scale_residual_bwd(g, y, ls, sb, hw, dls, dy=dybuf[:, C:2 * C])
    torch.testing.assert_close(dy.float(), g * ls * sb.repeat_interleave(hw)[:, None], **tol(dtype))
    torch.testing.assert_close(dls, (g * y.float() * sb.repeat_interleave(hw)[:, None]).sum(0), rtol=1e-3, atol=1e-2)
    # same pass also emits the bias gradient of the layer that produced y (column sums of dy), accumulated
    dls2, dbias = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    dy2 = k.scale_residual_bwd(g, y, ls, sb, hw, dls2, dy_colsum=dbias)
    torch.testing.assert_close(dy2, dy)
    torch.testing.assert_close(dls2, dls)
    torch.testing.assert_close(dbias - 1, dy2.float().sum(0), rtol=1e-3, atol=1e-2)
    # stand-alone activation on column slices
    z = rnd(M, 3 * C, dtype=dtype)
    o = k.act_fwd(z[:, C:2 * C], k.ACT_GELU)
    torch.testing.assert_close(o.float(), F.gelu(z[:, C:2 * C].float()), **tol(dtype))
    zr = z[:, C:2 * C].float().clone().requires_grad_(True)
    F.gelu(zr).backward(dout[:, :C].float())
    dz = k.act_bwd(dout[:, :C], z[:, C:2 * C], k.ACT_GELU)
    cz = torch.zeros(C, device=DEV)
    dz_b = k.act_bwd(dout[:, :C], z[:, C:2 * C], k.ACT_GELU, colsum=cz)
    torch.testing.assert_close(dz_b, dz)
    torch.testing.assert_close(cz, dz.float().sum(0), rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(dz.float(), zr.grad, **tol(dtype))


# ----------------------------------------------------------------------------- pooling / attention / resize
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("H,W", [(60, 80), (15, 20), (2, 3), (9, 7)])
def test_pool7(dtype, H, W):
    k = K()
    B, C1, C2 = 2, 64, 32
    xn, en = rnd(B, H, W, C1, dtype=dtype), rnd(B, H, W, C2, dtype=dtype)
    out = k.pool7_fwd(xn.view(-1, C1), en.view(-1, C2), B, H, W)
    cat = torch.cat([xn, en], 3).float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.adaptive_avg_pool2d(cat, (7, 7))
    torch.testing.assert_close(out.view(B, 7, 7, C1 + C2).float(), ref.permute(0, 2, 3, 1), **tol(dtype))
    dout = rnd(B, 7, 7, C1 + C2, dtype=dtype)
    ref.backward(dout.float().permute(0, 3, 1, 2))
    dxn, den = k.pool7_bwd(dout.view(-1, C1 + C2), C1, C2, B, H, W)
    gref = cat.grad.permute(0, 2, 3, 1)
    torch.testing.assert_close(dxn.view(B, H, W, C1).float(), gref[..., :C1], **tol(dtype))
    torch.testing.assert_close(den.view(B, H, W, C2).float(), gref[..., C1:], **tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("HW,heads,d", [(1200, 4, 36), (300, 8, 32), (4800, 2, 48), (6, 8, 16), (129, 1, 48), (256, 3, 16)])
def test_gaa(dtype, HW, heads, d):
    k = K()
    B, Cp = 2, heads * d
    m = rnd(B * 49, Cp, dtype=dtype)
    kv = rnd(B * HW, 2 * Cp, dtype=dtype)
    out, probs = k.gaa_fwd(m, kv, B, HW, heads, d)
    mr = m.float().clone().requires_grad_(True)
    kvr = kv.float().clone().requires_grad_(True)
    q = mr.view(B, 49, heads, d).permute(0, 2, 1, 3)
    kk, vv = kvr.view(B, HW, 2, heads, d).permute(2, 0, 3, 1, 4)
    att = ((q * d ** -0.5) @ kk.transpose(-2, -1)).softmax(-1)
    ref = (att @ vv).permute(0, 2, 1, 3).reshape(B * 49, Cp)
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-3)
    dout = rnd(B * 49, Cp)
    ref.backward(dout)
    dm, dkv = k.gaa_bwd(dout, m, kv, probs, B, HW, heads, d)
    t = dict(rtol=3e-2, atol=3e-2) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dm, mr.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dkv.float(), kvr.grad, **t)
    # one-launch fused form (keeps only the row log-sum-exp); run twice: the ticket counters must reset themselves
    # bf16 runs on the tensor cores (gaa_mma.cu): the probabilities / dS / dO enter the second products rounded to bf16, as
    # the reference's autocast bmm does (DFormer.py:125-130 under torch.autocast) -> bf16-level tolerances on out and dm
    tf = dict(rtol=2e-2, atol=1e-2) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)
    for _ in range(2):
        out2, lse = k.gaa_fused_fwd(m, kv, B, HW, heads, d)
        torch.testing.assert_close(out2, ref, **tf)
        ref_lse = torch.logsumexp((q * d ** -0.5) @ kk.transpose(-2, -1), dim=-1).reshape(-1)
        torch.testing.assert_close(lse, ref_lse.detach(), rtol=1e-3, atol=1e-3)
        dm2, dkv2 = k.gaa_fused_bwd(dout, out2, lse, m, kv, B, HW, heads, d)
        torch.testing.assert_close(dm2, mr.grad, **(t if dtype == torch.bfloat16 else tf))
        torch.testing.assert_close(dkv2.float(), kvr.grad, **t)
        if dtype == torch.bfloat16:           # the variant that also emits the two bias gradients (accumulating) and dm in bf16
            cs_kv, cs_m = torch.ones(2 * Cp, device=DEV), torch.ones(Cp, device=DEV)
            dm3, dkv3 = k.gaa_fused_bwd_ex(dout, out2, lse, m, kv, B, HW, heads, d, cs_kv, cs_m)
            torch.testing.assert_close(dkv3.float(), dkv2.float(), rtol=0, atol=0)
            torch.testing.assert_close(dm3.float(), dm2, rtol=1e-2, atol=1e-2)
            torch.testing.assert_close(cs_kv - 1, kvr.grad.sum(0), rtol=3e-2, atol=6e-2 * max(1.0, (B * HW) ** 0.5 / 8))
            torch.testing.assert_close(cs_m - 1, mr.grad.sum(0), rtol=3e-2, atol=5e-2)


@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("hi,wi,ho,wo", [(7, 7, 60, 80), (15, 20, 60, 80), (30, 40, 60, 80), (60, 80, 60, 80), (7, 7, 2, 3)])
def test_resize(in_dtype, out_dtype, hi, wi, ho, wo):
    k = K()
    B, C = 2, 32
    x = rnd(B, hi, wi, C, dtype=in_dtype)
    out = torch.zeros(B * ho * wo, 3 * C, device=DEV, dtype=out_dtype)
    k.resize_fwd(x.view(-1, C), B, hi, wi, out, ho, wo, col0=C)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=False)
    torch.testing.assert_close(out.view(B, ho, wo, 3 * C)[..., C:2 * C].float(), ref.permute(0, 2, 3, 1), **tol(out_dtype))
    assert out.view(B, ho, wo, 3 * C)[..., :C].abs().max() == 0
    dout = rnd(B * ho * wo, 3 * C, dtype=out_dtype)
    ref.backward(dout.view(B, ho, wo, 3 * C)[..., C:2 * C].float().permute(0, 3, 1, 2))
    din = torch.ones(B * hi * wi, C, device=DEV, dtype=in_dtype)
    k.resize_bwd(dout, C, B, hi, wi, C, ho, wo, din, accumulate=(in_dtype == torch.float32))
    want = xr.grad.permute(0, 2, 3, 1).reshape(-1, C) + (1 if in_dtype == torch.float32 else 0)
    t = dict(rtol=3e-2, atol=0.2) if torch.bfloat16 in (in_dtype, out_dtype) else dict(rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(din.float(), want, **t)


@pytest.mark.parametrize("C,hi,wi,ho,wo", [(48, 7, 7, 120, 160), (144, 7, 7, 30, 40), (288, 7, 7, 15, 20), (96, 7, 7, 60, 80), (2056, 2, 2, 9, 11)])
def test_resize_bwd_block_cooperative(C, hi, wi, ho, wo):
    """The one-CTA-per-input-pixel adjoint at the attention branch's real widths (vector counts that do not divide the CTA,
    more vectors than threads) against autograd of F.interpolate (`DFormer.py:131`)."""
    k = K()
    B = 2
    xr = torch.zeros(B, C, hi, wi, device=DEV, requires_grad=True)
    ref = F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=False)
    dout = rnd(B * ho * wo, C + 8, dtype=torch.bfloat16)
    ref.backward(dout.view(B, ho, wo, C + 8)[..., 8:].float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1).reshape(-1, C)
    for accumulate in (0, 1):
        din = torch.ones(B * hi * wi, C, device=DEV, dtype=torch.float32)
        k.resize_bwd(dout, 8, B, hi, wi, C, ho, wo, din, accumulate=bool(accumulate))
        torch.testing.assert_close(din, want + accumulate, rtol=1e-4, atol=2e-3)


# ----------------------------------------------------------------------------- im2col conv
@pytest.mark.parametrize("dtype", DTYPES)
def test_im2col_conv3x3s2(dtype):
    k = K()
    B, H, W = 2, 30, 42
    # NCHW fp32 network input with Cin = 3 (scalar path, K padded to 32)
    x = rnd(B, 3, H, W)
    col = k.im2col_fwd(x, (3 * H * W, W, 1, H * W), B, H, W, 3, dtype, 32)
    w = rnd(16, 3, 3, 3, scale=0.3)
    wp = torch.zeros(16, 32, device=DEV)
    wp[:, :27] = w.permute(0, 2, 3, 1).reshape(16, 27)
    out = col.float() @ wp.t()
    ref = F.conv2d(x.to(dtype).float(), w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 16)
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-3)
    # channel-0 slice of a 3-channel NCHW tensor with Cin = 1
    col1 = k.im2col_fwd(x, (3 * H * W, W, 1, H * W), B, H, W, 1, dtype, 16)
    w1 = rnd(8, 1, 3, 3)
    wp1 = torch.zeros(8, 16, device=DEV)
    wp1[:, :9] = w1.reshape(8, 9)
    ref1 = F.conv2d(x[:, 0:1].to(dtype).float(), w1, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 8)
    torch.testing.assert_close(col1.float() @ wp1.t(), ref1, rtol=1e-3, atol=1e-3)
    # channels-last vector path (+ odd spatial size) and its adjoint
    Cin, H2, W2 = 24, 15, 21
    xc = rnd(B, H2, W2, Cin, dtype=dtype)
    col2 = k.im2col_fwd(xc.view(-1, Cin), (H2 * W2 * Cin, W2 * Cin, Cin, 1), B, H2, W2, Cin, dtype, 9 * Cin)
    w2 = rnd(40, Cin, 3, 3, scale=0.1)
    xr = xc.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref2 = F.conv2d(xr, w2, stride=2, padding=1)
    out2 = col2.float() @ w2.permute(0, 2, 3, 1).reshape(40, -1).t()
    torch.testing.assert_close(out2, ref2.permute(0, 2, 3, 1).reshape(-1, 40), rtol=1e-3, atol=1e-3)
    dy = rnd(*ref2.shape)
    ref2.backward(dy)
    dcol = (dy.permute(0, 2, 3, 1).reshape(-1, 40) @ w2.permute(0, 2, 3, 1).reshape(40, -1)).to(dtype)
    din = k.im2col_bwd(dcol, B, H2, W2, Cin, dtype)
    torch.testing.assert_close(din.float(), xr.grad.permute(0, 2, 3, 1).reshape(-1, Cin), **(dict(rtol=3e-2, atol=5e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)))
    # pack / unpack of conv weights
    from dformer_b200.kernels import build_pack_table, pack_params, unpack_conv_grad
    dst = torch.zeros(40, 9 * Cin, device=DEV, dtype=dtype)
    lin = rnd(10, 24)
    dst2 = torch.zeros(12, 24, device=DEV, dtype=dtype)
    table, n, mx = build_pack_table([(w2, dst, 40, Cin, 9 * Cin, 1), (lin, dst2[2:], 10, 24, 24, 0)], DEV)
    pack_params(table, n, mx, k.dt(dst))
    torch.testing.assert_close(dst.float(), w2.permute(0, 2, 3, 1).reshape(40, -1).to(dtype).float())
    torch.testing.assert_close(dst2[2:].float(), lin.to(dtype).float())
    back = unpack_conv_grad(w2.permute(0, 2, 3, 1).reshape(40, -1).contiguous(), 40, Cin, torch.empty_like(w2))
    torch.testing.assert_close(back, w2)


# ----------------------------------------------------------------------------- BatchNorm
@pytest.mark.parametrize("x_dtype,y_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("act,use_res", [(0, False), (1, False), (2, False), (2, True)])
@pytest.mark.parametrize("M,C", [(2400, 48), (4099, 512), (37, 2056)])
def test_batchnorm_train(x_dtype, y_dtype, act, use_res, M, C):
    k = K()
    B = 1 if M % 2 else 2
    x = (rnd(M, C) * 1.5 + 0.7).to(x_dtype)
    gamma, beta = 1 + 0.1 * rnd(C), 0.1 * rnd(C)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    res = rnd(M, C, dtype=y_dtype) if use_res else None
    st = k.bn_stats(x)
    ms = k.bn_finalize(st, M, 1e-5, 0.1, rm, rv)
    y = k.bn_apply(x, ms, gamma, beta, y_dtype, residual=res, act=act)
    xr = x.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    z = F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    rr = res.float().clone().requires_grad_(True) if use_res else None
    if use_res:
        z = z + rr
    ref = F.gelu(z) if act == 1 else (F.relu(z) if act == 2 else z)
    torch.testing.assert_close(y.float(), ref, **tol(y_dtype))
    torch.testing.assert_close(rm, rm2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv, rv2, rtol=1e-4, atol=1e-5)
    dy = rnd(M, C, dtype=y_dtype)
    dy_b = rnd(M, C, dtype=y_dtype) if use_res else None             # second incoming gradient, summed inside the kernel
    ref.backward(dy.float() + (dy_b.float() if use_res else 0))
    gbuf, sums = k.bn_bwd_reduce(dy, x, ms, gamma, beta, res, act, None, M // B, dy2=dy_b)
    dx = k.bn_bwd_apply(gbuf, x, ms, gamma, sums, M, True, x_dtype)
    t = dict(rtol=3e-2, atol=3e-2) if torch.bfloat16 in (x_dtype, y_dtype) else dict(rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(dx.float(), xr.grad, **t)
    torch.testing.assert_close(sums[1], gr.grad, rtol=2e-2, atol=0.5 if y_dtype == torch.bfloat16 else 1e-2)   # dgamma
    torch.testing.assert_close(sums[0], br.grad, rtol=2e-2, atol=0.5 if y_dtype == torch.bfloat16 else 1e-2)   # dbeta
    if use_res:
        torch.testing.assert_close(gbuf.float(), rr.grad, **t)


def test_batchnorm_eval_and_channel_dropout():
    k = K()
    M, C, B = 960, 64, 2
    x = rnd(M, C)
    gamma, beta, rm, rv = 1 + 0.1 * rnd(C), 0.1 * rnd(C), 0.1 * rnd(C), 0.5 + torch.rand(C, device=DEV)
    ms = k.bn_eval_stats(rm, rv, 1e-3)
    mask = (torch.rand(B, C, device=DEV) > 0.3).float() / 0.7
    y = k.bn_apply(x, ms, gamma, beta, torch.float32, act=2, chan_scale=mask, rows_per_sample=M // B)
    ref = F.relu(F.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-3)) * mask.repeat_interleave(M // B, 0)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------- NMF helpers
def test_nmf_elementwise():
    k = K()
    b = torch.rand(2, 512, 64, device=DEV)
    nb, norms = k.normalize_cols(b)
    torch.testing.assert_close(nb, F.normalize(b, dim=1), rtol=1e-5, atol=1e-6)
    x = rnd(300, 64)
    sm = k.softmax_rows(x)
    torch.testing.assert_close(sm, x.softmax(-1), rtol=1e-5, atol=1e-6)
    g = rnd(300, 64)
    xr = x.clone().requires_grad_(True)
    xr.softmax(-1).backward(g)
    torch.testing.assert_close(k.softmax_rows_bwd(g, sm), xr.grad, rtol=1e-4, atol=1e-6)
    a, num, den = torch.rand(1000, device=DEV), torch.rand(1000, device=DEV), torch.rand(1000, device=DEV) + 0.1
    ar, nr, dr = (t.clone().requires_grad_(True) for t in (a, num, den))
    ref = ar * nr / (dr + 1e-6)
    torch.testing.assert_close(k.mu_update(a, num, den), ref, rtol=1e-5, atol=1e-6)
    ref.backward(g.flatten()[:1000])
    da = torch.ones(1000, device=DEV)
    out32, out_lo = k.mu_update(a, num, den, lo_dtype=torch.bfloat16)          # fp32 state + bf16 operand from one launch
    torch.testing.assert_close(out32, ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out_lo, ref.detach().bfloat16())
    # dnum goes into a column slice of a wider operand buffer (rows of 50 values, leading dimension 120), dden contiguous
    a2, n2, d2, g2 = (t.view(20, 50) for t in (a, num, den, g.flatten()[:1000].contiguous()))
    for dt_ in (torch.float32, torch.bfloat16):
        da = torch.ones(20, 50, device=DEV)
        wide = torch.zeros(20, 120, device=DEV, dtype=dt_)
        dden = k.mu_update_bwd(g2, a2, n2, d2, da, True, wide[:, 40:90], dt_)
        t_ = dict(rtol=1e-4, atol=1e-4) if dt_ == torch.float32 else dict(rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(da.flatten() - 1, ar.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(wide[:, 40:90].float().flatten(), nr.grad, **t_)
        torch.testing.assert_close(dden.float().flatten(), dr.grad, **t_)
        assert wide[:, :40].abs().max() == 0 and wide[:, 90:].abs().max() == 0
    y = torch.ones(100, device=DEV, dtype=torch.bfloat16)
    k.axpy(rnd(100), 2.0, y)
    assert k.cast(y, torch.float32).dtype == torch.float32


# ----------------------------------------------------------------------------- upsample + CE
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ncls", [40, 37])
def test_upsample_ce(dtype, ncls):
    k = K()
    B, h, w, H, W = 2, 12, 16, 96, 128
    small = rnd(B, h, w, ncls, dtype=dtype)
    label = torch.randint(0, ncls, (B, H, W), device=DEV)
    label[torch.rand(B, H, W, device=DEV) < 0.1] = 255
    out, lse, acc, loss, _ = k.upsample_ce_fwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255)
    sr = small.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    up = F.interpolate(sr, (H, W), mode="bilinear", align_corners=False)
    ref_loss = F.cross_entropy(up, label, reduction="none", ignore_index=255)[label != 255].mean()
    torch.testing.assert_close(out, up, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss, ref_loss, rtol=1e-4, atol=1e-5)
    (ref_loss * 3.0).backward()
    dl = torch.full((), 3.0, device=DEV)
    ds = k.upsample_ce_bwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, lse, acc, dl)
    t = dict(rtol=3e-2, atol=1e-5) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(ds.view(B, h, w, ncls).float(), sr.grad.permute(0, 2, 3, 1), **t)
    # separable form fed by the kept up-sampled logits
    out2, lse2, acc2, loss2, up = k.upsample_ce_fwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, want_out=False, keep_up=True)
    ds2 = k.upsample_ce_bwd_sep(up, dtype, B, h, w, ncls, H, W, label, 255, lse2, acc2, dl)
    t2 = dict(rtol=5e-2, atol=2e-5) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(ds2.view(B, h, w, ncls).float(), sr.grad.permute(0, 2, 3, 1), **t2)
    # training forward (loss + per-pixel log-sum-exp only): the row-grouped kernel (16-byte class vectors when ncls % 8 == 0)
    _, lse4, acc4, loss4, _ = k.upsample_ce_fwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, want_out=False)
    torch.testing.assert_close(loss4, ref_loss, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lse4, lse, rtol=1e-4, atol=1e-4)          # lse: the per-pixel kernel's, checked through the loss above
    torch.testing.assert_close(acc4[1], acc[1])
    # recompute form (what training uses): nothing but lse kept by the forward pass
    ds3 = k.upsample_ce_bwd_fused(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, lse, acc, dl)
    torch.testing.assert_close(ds3.view(B, h, w, ncls).float(), sr.grad.permute(0, 2, 3, 1), **t)


@pytest.mark.parametrize("h,w,H,W", [(12, 16, 96, 128), (7, 9, 30, 40), (5, 6, 5, 6), (1, 3, 8, 9), (15, 20, 33, 47), (6, 8, 6, 64)])
def test_upsample_ce_bwd_fused_general_scales(h, w, H, W):
    """Integer, fractional, identity and degenerate scale factors of the separable recompute-form adjoint."""
    k = K()
    B, ncls = 2, 13
    small = rnd(B, h, w, ncls)
    label = torch.randint(0, ncls, (B, H, W), device=DEV)
    label[torch.rand(B, H, W, device=DEV) < 0.2] = 255
    out, lse, acc, loss, _ = k.upsample_ce_fwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, want_out=False)
    sr = small.permute(0, 3, 1, 2).clone().requires_grad_(True)
    up = F.interpolate(sr, (H, W), mode="bilinear", align_corners=False)
    ref_loss = F.cross_entropy(up, label, reduction="none", ignore_index=255)[label != 255].mean()
    torch.testing.assert_close(loss, ref_loss, rtol=1e-4, atol=1e-5)
    ref_loss.backward()
    ds = k.upsample_ce_bwd_fused(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, lse, acc, torch.ones((), device=DEV))
    torch.testing.assert_close(ds.view(B, h, w, ncls), sr.grad.permute(0, 2, 3, 1), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ncls,h,w,H,W", [(40, 12, 16, 96, 128), (37, 12, 16, 96, 128), (40, 60, 80, 480, 640), (13, 7, 9, 30, 40), (13, 5, 6, 5, 6),
                                          (13, 1, 3, 8, 9), (13, 15, 20, 33, 47), (8, 6, 8, 6, 300), (40, 3, 40, 24, 320)])
def test_upsample_ce_one_pass_training_kernel(dtype, ncls, h, w, H, W):
    """loss + gradient of the masked-mean CE through the x8 (or any up-sampling) bilinear resize from ONE launch (what training uses):
    several column segments per row, ragged last segment, 16-byte and scalar class loads, ignore labels, identity / fractional scales."""
    k = K()
    B = 2
    assert k.upsample_ce_train_supported(h, w, H, W)
    small = rnd(B, h, w, ncls, dtype=dtype)
    label = torch.randint(0, ncls, (B, H, W), device=DEV)
    label[torch.rand(B, H, W, device=DEV) < 0.15] = 255
    loss, acc, dgrad = k.upsample_ce_train(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255)
    sr = small.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    up = F.interpolate(sr, (H, W), mode="bilinear", align_corners=False)
    ref_loss = F.cross_entropy(up, label, reduction="none", ignore_index=255)[label != 255].mean()
    torch.testing.assert_close(loss, ref_loss, rtol=1e-4, atol=1e-5)
    assert acc[1].item() == (label != 255).sum().item()
    (ref_loss * 3.0).backward()
    ds = k.ce_grad_finalize(dgrad, acc, torch.full((), 3.0, device=DEV), dtype)
    t = dict(rtol=2e-2, atol=2e-5) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(ds.view(B, h, w, ncls).float(), sr.grad.permute(0, 2, 3, 1), **t)
    # agrees with the three-launch path it replaces
    _, lse, acc2, loss2, _ = k.upsample_ce_fwd(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, want_out=False)
    ds2 = k.upsample_ce_bwd_fused(small.view(-1, ncls), B, h, w, ncls, H, W, label, 255, lse, acc2, torch.full((), 3.0, device=DEV))
    torch.testing.assert_close(loss, loss2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ds.float(), ds2.float(), **t)


def test_upsample_ce_one_pass_rejects_downsampling():
    k = K()
    assert not k.upsample_ce_train_supported(16, 16, 8, 8)
    small = rnd(1, 16, 16, 8)
    with pytest.raises(RuntimeError, match="unsupported geometry"):
        k.upsample_ce_train(small.view(-1, 8), 1, 16, 16, 8, 8, 8, torch.zeros(1, 8, 8, dtype=torch.long, device=DEV), 255)


def test_adamw_matches_torch():
    k = K()
    p = rnd(1000)
    g = rnd(1000)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        pr.grad = g.clone()
        opt.step()
        k.adamw(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    torch.testing.assert_close(p, pr.detach(), rtol=1e-5, atol=1e-6)
    # device-side schedule state (CUDA-graph friendly) gives the same update
    p2, m2, v2 = rnd(1000), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    p3, m3, v3 = p2.clone(), m2.clone(), v2.clone()
    dyn = torch.tensor([2e-3, 0.0], device=DEV)
    for step in (1, 2):
        dyn[1] += 1
        k.adamw(p2, g, m2, v2, 123.0, 0.9, 0.999, 1e-8, 0.01, 77, dyn=dyn)
        k.adamw(p3, g, m3, v3, 2e-3, 0.9, 0.999, 1e-8, 0.01, step)
    torch.testing.assert_close(p2, p3, rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- multi-scale evaluation + metrics (row N3)
@pytest.mark.parametrize("Hi,Wi,Ho,Wo", [(48, 64, 96, 128), (48, 64, 64, 96), (60, 80, 32, 32), (5, 7, 5, 7), (9, 1, 4, 3)])
@pytest.mark.parametrize("flip", [False, True])
def test_resize_nchw_align_corners(Hi, Wi, Ho, Wo, flip):
    k = K()
    x = rnd(2, 3, Hi, Wi)
    ref = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True)
    if flip:
        ref = torch.flip(ref, dims=(3,))
    torch.testing.assert_close(k.resize_nchw_ac(x, Ho, Wo, flip=flip), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("ncls,h,w,H,W", [(40, 24, 32, 48, 64), (37, 32, 32, 30, 45), (5, 7, 9, 7, 9)])
def test_ms_softmax_accum_and_confusion(ncls, h, w, H, W):
    k = K()
    B = 2
    acc = torch.zeros(B, ncls, H, W, device=DEV)
    ref = torch.zeros_like(acc)
    for flip in (False, True, False):
        logits = rnd(B, ncls, h, w) * 3
        k.ms_softmax_accum(logits, acc, flip=flip)
        lr = torch.flip(logits, dims=(3,)) if flip else logits
        ref += F.interpolate(lr, size=(H, W), mode="bilinear", align_corners=True).softmax(dim=1)
    torch.testing.assert_close(acc, ref, rtol=1e-4, atol=1e-5)
    target = torch.randint(0, ncls, (B, H, W), device=DEV)
    target[torch.rand(B, H, W, device=DEV) < 0.15] = 255
    hist = torch.zeros(ncls, ncls, device=DEV)
    pred = k.argmax_confusion(acc, target, 255, hist=hist, want_pred=True)
    pr = ref.argmax(dim=1)
    assert (pred == pr).float().mean() > 0.999
    keep = target != 255
    href = torch.bincount(target[keep] * ncls + pred[keep], minlength=ncls ** 2).view(ncls, ncls).float()
    torch.testing.assert_close(hist, href)
    k.argmax_confusion(acc, target, 255, hist=hist)              # accumulates
    torch.testing.assert_close(hist, 2 * href)


def test_metrics_match_reference_formulas():
    from dformer_b200.evaluation import Metrics, scaled_size
    ncls = 7
    pred = rnd(3, ncls, 20, 30)
    target = torch.randint(0, ncls, (3, 20, 30), device=DEV)
    target[:, :3] = 255
    m = Metrics(ncls, 255, DEV)
    m.update(pred, target)
    m.update(pred, target)
    keep = target != 255
    hist = 2 * torch.bincount(target[keep] * ncls + pred.argmax(1)[keep], minlength=ncls ** 2).view(ncls, ncls).float()
    ious = hist.diag() / (hist.sum(0) + hist.sum(1) - hist.diag())
    ious[ious.isnan()] = 0.0
    got, miou = m.compute_iou()
    assert got == (ious * 100).cpu().numpy().round(2).tolist() and miou == round(ious.mean().item() * 100, 2)
    f1 = 2 * hist.diag() / (hist.sum(0) + hist.sum(1))
    assert m.compute_f1()[1] == round(f1.mean().item() * 100, 2)
    acc = hist.diag() / hist.sum(1)
    assert m.compute_pixel_acc()[1] == round(acc.mean().item() * 100, 2)
    assert scaled_size(480, 640, 0.75) == (384, 480) and scaled_size(480, 640, 1.25) == (608, 800)


# ----------------------------------------------------------------------------- inference-time BatchNorm folding (row N4)
@pytest.mark.parametrize("dtype", DTYPES)
def test_bn_fold(dtype):
    k = K()
    rows, cols, ld = 48, 27, 32
    w = torch.zeros(rows, ld, device=DEV, dtype=dtype)
    w[:, :cols] = rnd(rows, cols, dtype=dtype)
    w0 = w.clone()
    cb, rm, rv, g, b = rnd(rows), rnd(rows), torch.rand(rows, device=DEV) + 0.2, 1 + 0.2 * rnd(rows), rnd(rows)
    bias = k.bn_fold(w, ld, cb, rm, rv, 1e-5, g, b)
    sc = g * torch.rsqrt(rv + 1e-5)
    torch.testing.assert_close(w.float(), (w0.float() * sc[:, None]), **tol(dtype))
    torch.testing.assert_close(bias, (cb - rm) * sc + b, rtol=1e-5, atol=1e-6)
    # conv -> BN(eval) == one GEMM with the folded weight and bias
    x = rnd(200, ld, dtype=dtype)
    ref = ((x.float() @ w0.float().t() + cb) - rm) * sc + b
    out = k.gemm(x, w, trans_b=True, bias=bias, backend=k.SIMT)
    torch.testing.assert_close(out.float(), ref, **(dict(rtol=3e-2, atol=8e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)))
    bias2 = k.bn_fold(w0.clone(), ld, None, rm, rv, 1e-5, g, b)               # conv without bias (the head's ConvModules)
    torch.testing.assert_close(bias2, -rm * sc + b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
def test_sym_cast(in_dtype, out_dtype):
    """x + x^T per matrix (the symmetrised Gram-matrix gradient of the NMF backward, ham_head.py:122-141)"""
    k = K()
    x = rnd(5, 64, 64, dtype=in_dtype)
    out = k.sym_cast(x, out_dtype)
    ref = (x.float() + x.float().transpose(1, 2)).to(out_dtype)
    torch.testing.assert_close(out, ref, rtol=0, atol=0)
