"""Thin torch-tensor front-ends for the C-ABI launchers (pointer / stride extraction only).

PyTorch is used here for device memory and the current CUDA stream -- no torch compute.
Every function launches asynchronously on ``torch.cuda.current_stream()``."""
import ctypes

import torch

from ._lib import GemmArgs, PackEntry, lib

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
SIMT, TCGEN05, AUTO = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16}
_TD = {F32: torch.float32, BF16: torch.bfloat16}


def dt(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def _s():
    return torch.cuda.current_stream().cuda_stream


def fork(side):
    """side stream waits for everything enqueued so far on the current stream"""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    side.wait_event(ev)


def join(side):
    """current stream waits for everything enqueued so far on the side stream"""
    ev = torch.cuda.Event()
    ev.record(side)
    torch.cuda.current_stream().wait_event(ev)


def signal(stream=None):
    ev = torch.cuda.Event()
    ev.record(stream if stream is not None else torch.cuda.current_stream())
    return ev


def share(stream, *tensors):
    """Tell the caching allocator that tensors created on another stream are also used on `stream`."""
    for t in tensors:
        if t is not None:
            t.record_stream(stream)


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"dformer_b200: {name} must be a CUDA tensor (there is no CPU path)")
    return t


GEMM_PROFILE = None      # bench.py sets this to a list to time every GEMM launch with CUDA events
GEMM_SHAPES_ONLY = False  # ... or just to collect the (shape, flops, bytes) census of a step


def gemm(a, b, *, trans_a=False, trans_b=True, bias=None, out=None, out_dtype=None, act=ACT_NONE, act_col_start=0,
         accumulate=False, backend=AUTO, splitk=0, alpha=0.0, M=None, N=None, K=None, gate=None, out2=None):
    """out[M,N] = act(alpha * op(a) @ op(b) + bias).  a/b are 2-D with unit inner stride (row slices allowed).
    Fused epilogue (tcgen05 backend, bf16): `gate` [M,N] -> out = (a @ b + bias) * gate, `out2` (optional) = a @ b + bias."""
    _chk(a), _chk(b)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if M is None:
        M = a.shape[1] if trans_a else a.shape[0]
    if K is None:
        K = a.shape[0] if trans_a else a.shape[1]
    if N is None:
        N = b.shape[0] if trans_b else b.shape[1]
    kb = b.shape[1] if trans_b else b.shape[0]
    assert kb >= K, (a.shape, b.shape, trans_a, trans_b)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype or a.dtype)
    assert out.stride(1) == 1
    g = GemmArgs()
    g.A, g.B, g.C, g.bias = a.data_ptr(), b.data_ptr(), out.data_ptr(), _p(bias)
    g.lda, g.ldb, g.ldc = a.stride(0), b.stride(0), out.stride(0)
    g.M, g.N, g.K, g.batch, g.batch_inner = M, N, K, 1, 1
    g.transA, g.transB = int(trans_a), int(trans_b)
    g.a_dtype, g.b_dtype, g.out_dtype = dt(a), dt(b), dt(out)
    g.act, g.act_col_start, g.accumulate, g.backend, g.splitk, g.alpha = act, act_col_start, int(accumulate), backend, splitk, alpha
    if gate is not None:
        _chk(gate)
        assert gate.shape == (M, N) and gate.stride(1) == 1 and gate.dtype == out.dtype
        g.epi_mode, g.aux, g.ld_aux = 1, gate.data_ptr(), gate.stride(0)
        if out2 is not None:
            _chk(out2)
            assert out2.shape == (M, N) and out2.stride(1) == 1 and out2.dtype == out.dtype
            g.out2, g.ld_out2 = out2.data_ptr(), out2.stride(0)
    if GEMM_PROFILE is not None and GEMM_SHAPES_ONLY:
        tc = backend != SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
        GEMM_PROFILE.append((None, None, 2.0 * M * N * K, a.element_size() * M * K + b.element_size() * N * K + out.element_size() * M * N, tc,
                             (M, N, K, int(trans_a), int(trans_b), out.dtype == torch.float32, bias is not None, int(accumulate))))
    elif GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib().gemm(ctypes.byref(g), _s())
        e1.record()
        tc = backend != SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
        GEMM_PROFILE.append((e0, e1, 2.0 * M * N * K, a.element_size() * M * K + b.element_size() * N * K + out.element_size() * M * N, tc,
                             (M, N, K, int(trans_a), int(trans_b))))
        return out
    lib().gemm(ctypes.byref(g), _s())
    return out


def gemm_sm_budget(main_sms=0, split_sms=0):
    """SMs the persistent GEMM grids may cover (0 = the in-step defaults; 148 = whole chip, for stand-alone timing)."""
    lib().gemm_sm_budget(main_sms, split_sms)


def bgemm(a, b, out, *, trans_a=False, trans_b=False, M, N, K, accumulate=False, alpha=0.0, splitk=1, backend=AUTO):
    """Strided-batched GEMM over the leading dimension of 3-D tensors: tcgen05 (3-D TMA maps) when both operands are
    bf16, otherwise the CUDA-core kernel."""
    assert a.dim() == 3 and b.dim() == 3 and out.dim() == 3 and a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
    if not (a.shape[0] == b.shape[0] == out.shape[0]):
        raise ValueError(f"bgemm: batch counts differ (a {a.shape[0]}, b {b.shape[0]}, out {out.shape[0]})")
    g = GemmArgs()
    g.A, g.B, g.C, g.bias = a.data_ptr(), b.data_ptr(), out.data_ptr(), None
    g.lda, g.ldb, g.ldc = a.stride(1), b.stride(1), out.stride(1)
    g.strideA, g.strideB, g.strideC = a.stride(0), b.stride(0), out.stride(0)
    g.M, g.N, g.K, g.batch, g.batch_inner = M, N, K, a.shape[0], 1
    g.transA, g.transB = int(trans_a), int(trans_b)
    g.a_dtype, g.b_dtype, g.out_dtype = dt(a), dt(b), dt(out)
    tc = (backend != SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and alpha in (0.0, 1.0)
          and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0 and a.stride(1) % 8 == 0 and b.stride(1) % 8 == 0)
    g.accumulate, g.backend, g.alpha = int(accumulate), (TCGEN05 if tc else SIMT), alpha
    g.splitk = (0 if out.dtype == torch.float32 else 1) if tc else splitk
    lib().gemm(ctypes.byref(g), _s())
    return out


def colsum(x, out=None, accumulate=False):
    M, N = x.shape
    if out is None:
        out = torch.empty(N, device=x.device, dtype=torch.float32)
    lib().colsum(x.data_ptr(), dt(x), x.stride(0), M, N, out.data_ptr(), int(accumulate), _s())
    return out


def build_pack_table(entries, device):
    """entries: list of (src_tensor, dst_tensor_view, rows, cols, dst_ld, kind) -> (device uint8 table, n, max_elems)"""
    arr = (PackEntry * len(entries))()
    mx = 1
    for i, (src, dst, rows, cols, ld, kind) in enumerate(entries):
        arr[i].src, arr[i].dst, arr[i].rows, arr[i].cols, arr[i].dst_ld, arr[i].kind = src.data_ptr(), dst.data_ptr(), rows, cols, ld, kind
        mx = max(mx, rows * cols * (9 if kind == 1 else 1))
    raw = bytes(arr)
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    return host.to(device), len(entries), mx


def pack_params(table, n, max_elems, dst_dtype):
    lib().pack_params(table.data_ptr(), n, max_elems, dst_dtype, _s())


def unpack_conv_grad(dwp, cout, cin, out):
    lib().unpack_conv_grad(dwp.data_ptr(), dwp.stride(0), cout, cin, out.data_ptr(), _s())
    return out


def layernorm_fwd(x, gamma, beta, eps, out_dtype):
    M, C = x.shape
    y = torch.empty((M, C), device=x.device, dtype=out_dtype)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    lib().layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, M, C, y.data_ptr(), dt(y), mean.data_ptr(), rstd.data_ptr(), _s())
    return y, mean, rstd


def scale_residual_layernorm_fwd(res, branch, ls, scale_b, rows_per_sample, gamma, beta, eps, out=None):
    """x1 = res + scale_b * ls * branch (fp32) and LN(x1) in branch.dtype in one pass: returns (x1, y, mean, rstd).
    `out` (optional, fp32 [M, C] contiguous) receives x1 instead of a fresh tensor."""
    M, C = res.shape
    if out is None:
        x1 = torch.empty_like(res)
    else:
        assert out.shape == res.shape and out.dtype == res.dtype and out.is_contiguous() and out.device == res.device
        x1 = out
    y = torch.empty((M, C), device=res.device, dtype=branch.dtype)
    mean = torch.empty(M, device=res.device, dtype=torch.float32)
    rstd = torch.empty(M, device=res.device, dtype=torch.float32)
    lib().scale_residual_layernorm_fwd(res.data_ptr(), branch.data_ptr(), branch.stride(0), dt(branch), ls.data_ptr(), _p(scale_b), rows_per_sample,
                                       M, C, x1.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _s())
    return x1, y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_in, dgamma, dbeta, dy2=None):
    """returns dx = (dx_in or 0) + LN-gradient of (dy [+ dy2]) (fresh fp32 tensor); dgamma/dbeta accumulate."""
    M, C = x.shape
    dx = torch.empty((M, C), device=x.device, dtype=torch.float32)
    assert dy2 is None or (dy2.dtype == dy.dtype and dy2.shape == dy.shape and dy2.is_contiguous())
    lib().layernorm_bwd(dy.data_ptr(), _p(dy2), dt(dy), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, C,
                        _p(dx_in), dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _s())
    return dx


def dwconv_fwd(x, weight, bias, B, H, W, k, add_input=False, act=ACT_NONE, save_z=False):
    C = x.shape[-1]
    y = torch.empty_like(x)
    z = torch.empty_like(x) if save_z else None
    lib().dwconv_fwd(x.data_ptr(), dt(x), weight.data_ptr(), _p(bias), B, H, W, C, k, int(add_input), act, y.data_ptr(), _p(z), _s())
    return (y, z) if save_z else y


def dwconv_bwd(dy, x, weight, bias, B, H, W, k, add_input, act, dweight, dbias, need_dx=True, z=None, wgrad_stream=None):
    """dx (and dz) on the current stream; the weight/bias gradient optionally on `wgrad_stream` (it is a leaf of the
    backward graph: nothing downstream waits for it except the end-of-module join)."""
    C = x.shape[-1]
    dz = torch.empty_like(dy) if act != ACT_NONE else None
    dx = torch.empty_like(dy) if need_dx else None
    if wgrad_stream is None:
        lib().dwconv_bwd(dy.data_ptr(), x.data_ptr(), _p(z), dt(x), weight.data_ptr(), _p(bias), B, H, W, C, k, int(add_input), act,
                         _p(dz), _p(dx), dweight.data_ptr(), dbias.data_ptr(), _s())
        return dx
    lib().dwconv_bwd(dy.data_ptr(), x.data_ptr(), _p(z), dt(x), weight.data_ptr(), _p(bias), B, H, W, C, k, int(add_input), act,
                     _p(dz), _p(dx), None, None, _s())
    g = dz if dz is not None else dy
    wgrad_stream.wait_event(signal())
    g.record_stream(wgrad_stream)
    x.record_stream(wgrad_stream)
    with torch.cuda.stream(wgrad_stream):
        lib().dwconv_bwd(g.data_ptr(), x.data_ptr(), None, dt(x), weight.data_ptr(), _p(bias), B, H, W, C, k, int(add_input), ACT_NONE,
                         None, None, dweight.data_ptr(), dbias.data_ptr(), _s())
    return dx


def mlp_dw_fwd(h, weight, bias, B, H, W, save_gp=False):
    """u = GELU(dw3x3(h) + bias + h) for bf16 channels-last h [B*H*W, C] (DFormer.py:62-64).  save_gp: also return
    GELU'(pre-activation) (bf16), which turns the backward kernel into a pure stream."""
    u = torch.empty_like(h)
    gp = torch.empty_like(h) if save_gp else None
    lib().mlp_dw_fwd(h.data_ptr(), dt(h), weight.data_ptr(), _p(bias), B, H, W, h.shape[-1], u.data_ptr(), _p(gp), _s())
    return (u, gp) if save_gp else u


def mlp_dw_bwd(du, h, weight, bias, B, H, W, dweight, dbias, dh_colsum=None, gp=None):
    """Fused backward of mlp_dw_fwd: returns dh; accumulates dweight, dbias and (optionally) colsum(dh) in place.
    gp = GELU'(pre-activation) kept by the forward pass, or None to recompute it from h."""
    dh = torch.empty_like(du)
    lib().mlp_dw_bwd(du.data_ptr(), _p(gp), h.data_ptr(), dt(h), weight.data_ptr(), _p(bias), B, H, W, h.shape[-1], dh.data_ptr(),
                     dweight.data_ptr(), _p(dbias), _p(dh_colsum), _s())
    return dh


def mul_fwd(a, b, out):
    M, N = a.shape
    lib().mul_fwd(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0), dt(a), M, N, _s())
    return out


def mul_bwd(dout, a, b, da, db, da_colsum=None, db_colsum=None):
    """da = dout * b, db = dout * a; with the colsum pair also accumulates the column sums of both (bias gradients)."""
    M, N = a.shape
    lib().mul_bwd(dout.data_ptr(), dout.stride(0), a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), da.data_ptr(), da.stride(0),
                  db.data_ptr(), db.stride(0), dt(a), M, N, _p(da_colsum), _p(db_colsum), _s())


def scale_residual_fwd(res, y, ls, scale_b, rows_per_sample):
    M, C = res.shape
    out = torch.empty_like(res)
    lib().scale_residual_fwd(res.data_ptr(), y.data_ptr(), y.stride(0), dt(y), ls.data_ptr(), _p(scale_b), M, C, rows_per_sample, out.data_ptr(), _s())
    return out


def scale_residual_bwd(dout, y, ls, scale_b, rows_per_sample, dls, dy=None, dy_colsum=None):
    """dy = dout * ls * scale_b; dls += colsum(dout * y * scale_b); optionally dy_colsum += colsum(dy) (the bias gradient of the
    Linear that produced y), from the same pass."""
    M, C = dout.shape
    if dy is None:
        dy = torch.empty((M, C), device=y.device, dtype=y.dtype)
    lib().scale_residual_bwd(dout.data_ptr(), y.data_ptr(), y.stride(0), dt(y), ls.data_ptr(), _p(scale_b), M, C, rows_per_sample,
                             dy.data_ptr(), dy.stride(0), dls.data_ptr(), _p(dy_colsum), _s())
    return dy


def act_fwd(x, act, out=None):
    M, N = x.shape
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=x.dtype)
    lib().act_fwd(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), dt(x), act, M, N, _s())
    return out


def act_bwd(dout, z, act, out=None, dout2=None, colsum=None):
    """out = (dout [+ dout2]) * act'(z); colsum (optional) += column sums of out"""
    M, N = z.shape
    if out is None:
        out = torch.empty((M, N), device=z.device, dtype=z.dtype)
    lib().act_bwd(dout.data_ptr(), dout.stride(0), _p(dout2), dout2.stride(0) if dout2 is not None else 0, z.data_ptr(), z.stride(0),
                  out.data_ptr(), out.stride(0), dt(z), act, M, N, _p(colsum), _s())
    return out


def pool7_fwd(xn, en, B, H, W):
    C1, C2 = xn.shape[-1], en.shape[-1]
    out = torch.empty((B * 49, C1 + C2), device=xn.device, dtype=xn.dtype)
    lib().pool7_fwd(xn.data_ptr(), C1, en.data_ptr(), C2, dt(xn), B, H, W, out.data_ptr(), _s())
    return out


def pool7_bwd(dout, C1, C2, B, H, W):
    dxn = torch.empty((B * H * W, C1), device=dout.device, dtype=dout.dtype)
    den = torch.empty((B * H * W, C2), device=dout.device, dtype=dout.dtype)
    lib().pool7_bwd(dout.data_ptr(), C1, C2, dt(dout), B, H, W, dxn.data_ptr(), den.data_ptr(), _s())
    return dxn, den


def gaa_fwd(m, kv, B, HW, heads, d):
    out = torch.empty((B * 49, heads * d), device=m.device, dtype=torch.float32)
    probs = torch.empty((B, heads, 49, HW), device=m.device, dtype=torch.float32)
    lib().gaa_fwd(m.data_ptr(), kv.data_ptr(), dt(m), B, HW, heads, d, out.data_ptr(), probs.data_ptr(), _s())
    return out, probs


def gaa_bwd(dout, m, kv, probs, B, HW, heads, d):
    dm = torch.empty((B * 49, heads * d), device=m.device, dtype=torch.float32)
    dkv = torch.empty_like(kv)
    scratch = torch.empty((B, heads, 49, HW), device=m.device, dtype=torch.float32)
    lib().gaa_bwd(dout.data_ptr(), m.data_ptr(), kv.data_ptr(), probs.data_ptr(), dt(m), B, HW, heads, d, dm.data_ptr(), dkv.data_ptr(),
                  scratch.data_ptr(), _s())
    return dm, dkv


GAA_FUSED_DIMS = (16, 32, 36, 48)
_GAA_COUNTERS = {}


def gaa_fused_fwd(m, kv, B, HW, heads, d):
    """One-launch attention core: returns (out [B*49, heads*d] fp32, lse [B*heads*49] fp32)."""
    dev = m.device
    out = torch.empty((B * 49, heads * d), device=dev, dtype=torch.float32)
    lse = torch.empty((B * heads * 49,), device=dev, dtype=torch.float32)
    if m.dtype == torch.bfloat16:                     # tensor-core kernel: partials are merged inside a thread-block cluster (no scratch, no tickets)
        lib().gaa_fused_fwd(m.data_ptr(), kv.data_ptr(), dt(m), B, HW, heads, d, out.data_ptr(), lse.data_ptr(), None, None, _s())
        return out, lse
    scratch = torch.empty((B * heads * ((HW + 127) // 128) * 49 * (d + 4),), device=dev, dtype=torch.float32)
    key = (dev, torch.cuda.current_stream().cuda_stream)
    cnt = _GAA_COUNTERS.get(key)                      # self-resetting tickets: one persistent zeroed buffer per (device, stream)
    if cnt is None or cnt.numel() < B * heads:
        cnt = _GAA_COUNTERS[key] = torch.zeros(max(B * heads, 1024), device=dev, dtype=torch.int32)
    lib().gaa_fused_fwd(m.data_ptr(), kv.data_ptr(), dt(m), B, HW, heads, d, out.data_ptr(), lse.data_ptr(), scratch.data_ptr(), cnt.data_ptr(), _s())
    return out, lse


def gaa_fused_bwd(dout, out, lse, m, kv, B, HW, heads, d):
    dm = torch.empty((B * 49, heads * d), device=m.device, dtype=torch.float32)
    dkv = torch.empty_like(kv)
    lib().gaa_fused_bwd(dout.data_ptr(), out.data_ptr(), lse.data_ptr(), m.data_ptr(), kv.data_ptr(), dt(m), B, HW, heads, d, dm.data_ptr(),
                        dkv.data_ptr(), _s())
    return dm, dkv


def gaa_fused_bwd_ex(dout, out, lse, m, kv, B, HW, heads, d, dkv_colsum, dm_colsum):
    """bf16 tensor-core backward that also accumulates the bias gradients of the kv / short_cut_linear projections and returns
    dm already in bf16: (dm_lo, dkv)."""
    assert m.dtype == torch.bfloat16
    dm = torch.empty((B * 49, heads * d), device=m.device, dtype=torch.float32)
    dm_lo = torch.empty((B * 49, heads * d), device=m.device, dtype=torch.bfloat16)
    dkv = torch.empty_like(kv)
    lib().gaa_fused_bwd_ex(dout.data_ptr(), out.data_ptr(), lse.data_ptr(), m.data_ptr(), kv.data_ptr(), B, HW, heads, d, dm.data_ptr(),
                           dkv.data_ptr(), _p(dkv_colsum), _p(dm_colsum), dm_lo.data_ptr(), _s())
    return dm_lo, dkv


def resize_fwd(inp, B, Hi, Wi, out, Ho, Wo, col0=0):
    C = inp.shape[-1]
    lib().resize_fwd(inp.data_ptr(), dt(inp), B, Hi, Wi, C, out.data_ptr(), dt(out), Ho, Wo, out.stride(0), col0, _s())
    return out


def resize_bwd(dout, col0, B, Hi, Wi, C, Ho, Wo, din, accumulate=False):
    lib().resize_bwd(dout.data_ptr(), dt(dout), dout.stride(0), col0, B, Hi, Wi, C, Ho, Wo, din.data_ptr(), dt(din), int(accumulate), _s())
    return din


def im2col_fwd(inp, strides, B, H, W, Cin, out_dtype, ld):
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    out = torch.empty((B * Ho * Wo, ld), device=inp.device, dtype=out_dtype)
    sb, sy, sx, sc = strides
    lib().im2col3x3s2_fwd(inp.data_ptr(), dt(inp), sb, sy, sx, sc, B, H, W, Cin, out.data_ptr(), dt(out), ld, _s())
    return out


def im2col_bwd(dcol, B, H, W, Cin, in_dtype):
    din = torch.empty((B * H * W, Cin), device=dcol.device, dtype=in_dtype)
    lib().im2col3x3s2_bwd(dcol.data_ptr(), dt(dcol), dcol.stride(0), B, H, W, Cin, din.data_ptr(), dt(din), _s())
    return din


def bn_stats(x):
    M, C = x.shape
    st = torch.empty((2, C), device=x.device, dtype=torch.float64)
    lib().bn_stats(x.data_ptr(), dt(x), M, C, st[0].data_ptr(), st[1].data_ptr(), _s())
    return st


def bn_finalize(st, count, eps, momentum, running_mean, running_var):
    C = st.shape[1]
    ms = torch.empty((2, C), device=st.device, dtype=torch.float32)
    lib().bn_finalize(st[0].data_ptr(), st[1].data_ptr(), float(count), eps, momentum, C, ms[0].data_ptr(), ms[1].data_ptr(),
                      _p(running_mean), _p(running_var), _s())
    return ms


def bn_eval_stats(running_mean, running_var, eps):
    C = running_mean.numel()
    ms = torch.empty((2, C), device=running_mean.device, dtype=torch.float32)
    lib().bn_eval_stats(running_mean.data_ptr(), running_var.data_ptr(), eps, C, ms[0].data_ptr(), ms[1].data_ptr(), _s())
    return ms


def bn_apply(x, ms, gamma, beta, out_dtype, residual=None, act=ACT_NONE, chan_scale=None, rows_per_sample=1):
    M, C = x.shape
    y = torch.empty((M, C), device=x.device, dtype=out_dtype)
    lib().bn_apply(x.data_ptr(), dt(x), ms[0].data_ptr(), ms[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(residual), act,
                   _p(chan_scale), rows_per_sample, M, C, y.data_ptr(), dt(y), _s())
    return y


def bn_bwd_reduce(dy, x, ms, gamma, beta, residual, act, chan_scale, rows_per_sample, dbeta=None, dgamma=None, dy2=None):
    """dbeta / dgamma (optional): accumulate the local parameter gradients (= the two sums) in the same pass;
    dy2 (optional): a second incoming gradient, summed with dy in the same pass"""
    assert dy2 is None or (dy2.shape == dy.shape and dy2.dtype == dy.dtype and dy2.is_contiguous())
    M, C = x.shape
    gbuf = torch.empty_like(dy)
    sums = torch.zeros((2, C), device=x.device, dtype=torch.float32)
    lib().bn_bwd_reduce(dy.data_ptr(), _p(dy2), dt(dy), x.data_ptr(), dt(x), ms[0].data_ptr(), ms[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                        _p(residual), act, _p(chan_scale), rows_per_sample, M, C, gbuf.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), _p(dbeta), _p(dgamma), _s())
    return gbuf, sums


def bn_bwd_apply(gbuf, x, ms, gamma, sums, count, training, dx_dtype):
    M, C = x.shape
    dx = torch.empty((M, C), device=x.device, dtype=dx_dtype)
    lib().bn_bwd_apply(gbuf.data_ptr(), dt(gbuf), x.data_ptr(), dt(x), ms[0].data_ptr(), ms[1].data_ptr(), gamma.data_ptr(),
                       sums[0].data_ptr(), sums[1].data_ptr(), float(count), int(training), M, C, dx.data_ptr(), dt(dx), _s())
    return dx


def bn_fold(w_packed, cols, conv_bias, running_mean, running_var, eps, gamma, beta):
    """Fold an eval-mode BatchNorm into the freshly packed weight of the conv in front of it (in place); returns the folded bias."""
    rows = w_packed.shape[0]
    bias = torch.empty(rows, device=w_packed.device, dtype=torch.float32)
    lib().bn_fold(w_packed.data_ptr(), dt(w_packed), rows, cols, w_packed.stride(0), _p(conv_bias), running_mean.data_ptr(), running_var.data_ptr(),
                  eps, gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(), _s())
    return bias


def normalize_cols(x):
    B, D, R = x.shape
    out = torch.empty_like(x)
    norms = torch.empty((B, R), device=x.device, dtype=torch.float32)
    lib().normalize_cols(x.data_ptr(), B, D, R, out.data_ptr(), norms.data_ptr(), _s())
    return out, norms


def softmax_rows(x):
    out = torch.empty_like(x)
    lib().softmax_rows(x.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], out.data_ptr(), _s())
    return out


def softmax_rows_bwd(dout, out):
    din = torch.empty_like(out)
    lib().softmax_rows_bwd(dout.data_ptr(), out.data_ptr(), out.numel() // out.shape[-1], out.shape[-1], din.data_ptr(), _s())
    return din


def mu_update(a, num, den, eps=1e-6, lo_dtype=None):
    """a * num / (den + eps); with lo_dtype also returns a copy in that dtype written by the same launch."""
    out = torch.empty_like(a)
    lo = torch.empty(a.shape, device=a.device, dtype=lo_dtype) if lo_dtype is not None and lo_dtype != a.dtype else None
    lib().mu_update(a.data_ptr(), num.data_ptr(), den.data_ptr(), eps, a.numel(), out.data_ptr(), _p(lo), _DT[lo_dtype] if lo is not None else 0, _s())
    if lo_dtype is None:
        return out
    return out, (lo if lo is not None else out)


def mu_update_bwd(dout, a, num, den, da, accumulate_da, dnum_out, dden_dtype, eps=1e-6):
    """da (fp32) in place; dnum written into `dnum_out` (a [..., cols] view with uniform row stride, e.g. a column slice of a
    K-concatenated GEMM operand) and dden returned contiguous, both in the GEMM compute dtype."""
    cols = a.shape[-1]
    assert dnum_out.shape == a.shape and dnum_out.stride(-1) == 1 and dnum_out.dtype == dden_dtype
    dden = torch.empty(a.shape, device=a.device, dtype=dden_dtype)
    lib().mu_update_bwd(dout.data_ptr(), a.data_ptr(), num.data_ptr(), den.data_ptr(), eps, a.numel(), da.data_ptr(), int(accumulate_da),
                        dnum_out.data_ptr(), dnum_out.stride(-2), cols, dden.data_ptr(), _DT[dden_dtype], _s())
    return dden


def cast(x, dtype):
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    lib().cast(x.data_ptr(), dt(x), out.data_ptr(), dt(out), x.numel(), _s())
    return out


def sym_cast(x, dtype):
    """x [B, R, R] -> x + x^T (per matrix) in `dtype`"""
    B, R, R2 = x.shape
    assert R == R2 and x.is_contiguous()
    out = torch.empty((B, R, R), device=x.device, dtype=dtype)
    lib().sym_cast(x.data_ptr(), dt(x), out.data_ptr(), dt(out), B, R, _s())
    return out


def cast_into(x, dst):
    """dst[..., :] = x converted to dst.dtype; x, dst are [..., cols] with unit inner stride and uniform row strides
    (dst is typically a column slice of a concatenated operand buffer)."""
    cols = x.shape[-1]
    rows = x.numel() // cols
    assert x.stride(-1) == 1 and dst.stride(-1) == 1 and dst.shape == x.shape
    lib().cast2d(x.data_ptr(), dt(x), x.stride(-2), dst.data_ptr(), dt(dst), dst.stride(-2), rows, cols, _s())
    return dst


def axpy(x, alpha, y):
    lib().axpy(x.data_ptr(), dt(x), alpha, y.data_ptr(), dt(y), x.numel(), _s())
    return y


def upsample_ce_fwd(small, B, h, w, ncls, H, W, label, ignore, want_out=True, want_loss=True, keep_up=False):
    """keep_up: also keep the up-sampled logits in small's dtype for the separable backward (the fp32 `out` doubles as
    that copy when small is fp32 and out is requested)."""
    dev = small.device
    out = torch.empty((B, ncls, H, W), device=dev, dtype=torch.float32) if want_out else None
    lse = torch.empty((B, H, W), device=dev, dtype=torch.float32) if want_loss else None
    acc = torch.zeros(2, device=dev, dtype=torch.float32) if want_loss else None
    up = None
    if keep_up and want_loss:
        up = out if (out is not None and small.dtype == torch.float32) else torch.empty((B, ncls, H, W), device=dev, dtype=small.dtype)
    lib().upsample_ce_fwd(small.data_ptr(), dt(small), B, h, w, ncls, H, W, _p(label), ignore, _p(out), _p(lse), _p(acc),
                          _p(up) if up is not out else None, _s())
    loss = None
    if want_loss:
        loss = torch.empty((), device=dev, dtype=torch.float32)
        lib().ce_finalize(acc.data_ptr(), loss.data_ptr(), _s())
    return out, lse, acc, loss, up


def upsample_ce_train_supported(h, w, H, W):
    """geometries the one-pass training kernel covers (dfb200_upsample_ce_train)"""
    return H >= h and W >= w and (3 * H + 2 * h - 1) // (2 * h) <= 12


def upsample_ce_train(small, B, h, w, ncls, H, W, label, ignore):
    """loss and the UNSCALED gradient w.r.t. the low-res logits in one launch: returns (loss, acc, dgrad fp32 [B*h*w, ncls])"""
    dev = small.device
    acc = torch.zeros(2, device=dev, dtype=torch.float32)
    dgrad = torch.zeros((B * h * w, ncls), device=dev, dtype=torch.float32)
    lib().upsample_ce_train(small.data_ptr(), dt(small), B, h, w, ncls, H, W, label.data_ptr(), ignore, None, acc.data_ptr(), dgrad.data_ptr(), _s())
    loss = torch.empty((), device=dev, dtype=torch.float32)
    lib().ce_finalize(acc.data_ptr(), loss.data_ptr(), _s())
    return loss, acc, dgrad


def ce_grad_finalize(dgrad, acc, dloss, out_dtype):
    """gradient of the masked-mean loss: dgrad * dloss / #valid, in out_dtype"""
    out = torch.empty(dgrad.shape, device=dgrad.device, dtype=out_dtype)
    lib().ce_grad_finalize(dgrad.data_ptr(), dgrad.numel(), acc.data_ptr(), dloss.data_ptr(), out.data_ptr(), dt(out), _s())
    return out


def upsample_ce_bwd_sep(up, small_dtype, B, h, w, ncls, H, W, label, ignore, lse, acc, dloss):
    dsmall = torch.empty((B * h * w, ncls), device=up.device, dtype=small_dtype)
    scratch = torch.empty((B, ncls, h, W), device=up.device, dtype=torch.float32)
    lib().upsample_ce_bwd_sep(up.data_ptr(), dt(up), B, h, w, ncls, H, W, label.data_ptr(), ignore, lse.data_ptr(), acc.data_ptr(),
                              dloss.data_ptr(), scratch.data_ptr(), dsmall.data_ptr(), dt(dsmall), _s())
    return dsmall


def upsample_ce_bwd_fused(small, B, h, w, ncls, H, W, label, ignore, lse, acc, dloss):
    """Separable adjoint of upsample+CE recomputed from the low-res logits (no hi-res copy kept by the forward pass)."""
    dsmall = torch.empty_like(small)
    scratch = torch.empty((2, B, h, W, ncls), device=small.device, dtype=torch.float32)
    lib().upsample_ce_bwd_fused(small.data_ptr(), dt(small), B, h, w, ncls, H, W, label.data_ptr(), ignore, lse.data_ptr(), acc.data_ptr(),
                                dloss.data_ptr(), scratch.data_ptr(), dsmall.data_ptr(), dt(dsmall), _s())
    return dsmall


def upsample_ce_bwd(small, B, h, w, ncls, H, W, label, ignore, lse, acc, dloss):
    dsmall = torch.empty_like(small)
    lib().upsample_ce_bwd(small.data_ptr(), dt(small), B, h, w, ncls, H, W, label.data_ptr(), ignore, lse.data_ptr(), acc.data_ptr(),
                          dloss.data_ptr(), dsmall.data_ptr(), dt(dsmall), _s())
    return dsmall


def resize_nchw_ac(x, Ho, Wo, flip=False):
    """F.interpolate(x, (Ho, Wo), mode='bilinear', align_corners=True) [then torch.flip(dims=(3,))] for NCHW fp32."""
    _chk(x, "x")
    B, C, Hi, Wi = x.shape
    x = x.contiguous().float()
    out = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32)
    lib().resize_nchw_ac(x.data_ptr(), B, C, Hi, Wi, out.data_ptr(), Ho, Wo, int(flip), _s())
    return out


def ms_softmax_accum(logits, acc, flip=False):
    """acc += softmax(F.interpolate(flip_W(logits) if flip else logits, acc.shape[2:], bilinear, align_corners=True), dim=1)"""
    B, C, h, w = logits.shape
    assert acc.shape[:2] == (B, C) and acc.dtype == torch.float32 and acc.is_contiguous()
    logits = logits.contiguous().float()
    lib().ms_softmax_accum(logits.data_ptr(), B, C, h, w, acc.data_ptr(), acc.shape[2], acc.shape[3], int(flip), _s())
    return acc


def argmax_confusion(score, target, ignore, hist=None, want_pred=False):
    """pred = score.argmax(1); hist[target * C + pred] += 1 over target != ignore.  Returns pred (or None)."""
    B, C = score.shape[:2]
    HW = score[0, 0].numel()
    assert score.dtype == torch.float32 and score.is_contiguous()
    pred = torch.empty((B,) + tuple(score.shape[2:]), device=score.device, dtype=torch.int64) if want_pred else None
    if target is not None:
        target = target.contiguous()
        assert target.dtype == torch.int64 and target.numel() == B * HW
    lib().argmax_confusion(score.data_ptr(), _p(target), B, C, HW, ignore, _p(hist), _p(pred), _s())
    return pred


def adamw(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, wd_arr=None, lr_arr=None, dyn=None):
    c1 = 1.0 - beta1 ** step
    c2 = 1.0 - beta2 ** step
    lib().adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps, wd, c1, c2, grad_scale,
                _p(wd_arr), _p(lr_arr), _p(dyn), _s())
