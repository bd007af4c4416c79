"""autograd.Functions with hand-scheduled forward AND backward for the DFormer hot path.

Granularity follows the reference's modules -- one Function per stem / downsample layer / Block /
LightHamHead / upsample+CE -- so the autograd graph has ~50 nodes instead of ~3500 ATen ops, fan-in
gradients are accumulated inside our kernels, and parameter gradients are written straight into the
flat GradArena.  Everything below is launch orchestration: pointers, shapes and kernel order; all
arithmetic happens in libdformer_b200.so."""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import torch
import torch.distributed as dist

from . import kernels as K
from .parallel import small_all_reduce_
from .runtime import GradArena, backend_for

F32 = torch.float32
import os as _os


# ============================================================================================ helpers
def _lin(x, wb, T, act=K.ACT_NONE, act_col_start=0, out=None, out_dtype=None, **epi):
    """epi: the fused gate epilogue of the tcgen05 GEMM (gate= / out2=, kernels.gemm)"""
    w, b = wb
    return K.gemm(x, w, trans_b=True, bias=b, out=out, out_dtype=out_dtype or T, act=act, act_col_start=act_col_start,
                  backend=backend_for(T), K=x.shape[1], **epi)


_WGRAD_STREAM = None      # set by BlockFn.backward: weight/bias gradients are leaves of the backward graph


def _lin_bwd(dy, x, w, dW, db, T, need_dx=True, dx_out=None):
    """dW[N,K] = dy^T x (fp32, into the arena), db = colsum(dy), returns dx = dy @ W.
    The parameter gradients go to the wgrad side stream when one is active (nothing downstream consumes them)."""
    be = backend_for(T)
    ws = _WGRAD_STREAM
    if ws is not None:
        ws.wait_event(K.signal())
        dy.record_stream(ws)
        x.record_stream(ws)                         # saved activation: may be released before the trailing wgrad stream has read it
        with torch.cuda.stream(ws):
            K.gemm(dy, x, trans_a=True, trans_b=False, out=dW, backend=be, accumulate=True)
            if db is not None:
                K.colsum(dy, out=db)
    else:
        # dW lives in the zero-initialised gradient arena (or a zeroed scratch): accumulate so that split-K needs no memset
        K.gemm(dy, x, trans_a=True, trans_b=False, out=dW, backend=be, accumulate=True)
        if db is not None:
            K.colsum(dy, out=db)
    if not need_dx:
        return None
    return K.gemm(dy, w, trans_a=False, trans_b=False, out=dx_out, out_dtype=T, backend=be, N=w.shape[1])


def _views(arena: GradArena, prefix: str, names):
    return {n: arena.view(prefix + n) for n in names}


class BNState:
    """Plain container describing one BatchNorm layer for the kernels (parameters + buffers + mode)."""

    def __init__(self, mod, prefix, training, sync_group=None):
        self.weight, self.bias = mod.weight, mod.bias
        self.running_mean, self.running_var, self.nbt = mod.running_mean, mod.running_var, mod.num_batches_tracked
        self.eps, self.momentum = mod.eps, (mod.momentum if mod.momentum is not None else 0.1)
        training = bool(training and mod.training)             # honour a per-layer .eval() (frozen statistics) like nn.Module does
        self.training = training or mod.running_mean is None
        self.prefix = prefix
        self.sync_group = sync_group if (training and dist.is_available() and dist.is_initialized() and sync_group is not False) else None
        self.sync = self.sync_group is not None


def _bn_fwd(x2d, bn: BNState, out_dtype, act=K.ACT_NONE, residual=None, chan_scale=None, rows_per_sample=1):
    M, C = x2d.shape
    if bn.training:
        st = K.bn_stats(x2d)                                    # [2, C] fp64: sum, sum of squares
        count = float(M)
        if bn.sync:                                             # SyncBatchNorm: one all-reduce of (sum, sumsq); equal shards per rank
            grp = bn.sync_group if bn.sync_group is not True else None
            small_all_reduce_(st, grp)                          # NVLink peer-memory kernel on one box (parallel.PeerExchange), else NCCL / gloo
            count = float(M) * dist.get_world_size(grp)
        ms = K.bn_finalize(st, count, bn.eps, bn.momentum, bn.running_mean, bn.running_var)
        if bn.nbt is not None:
            bn.nbt.add_(1)
    else:
        ms = K.bn_eval_stats(bn.running_mean, bn.running_var, bn.eps)
        count = float(M)
    y = K.bn_apply(x2d, ms, bn.weight, bn.bias, out_dtype, residual=residual, act=act, chan_scale=chan_scale, rows_per_sample=rows_per_sample)
    return y, ms, count


_FOLD_BN = _os.environ.get("DFB200_FOLD_BN", "1") == "1"       # inference: conv -> BN(eval) pairs run as one GEMM


def _wants_grad(ctx, st) -> bool:
    """Will this call ever be differentiated?  `ctx.needs_input_grad` alone is not the answer: autograd fills it from the inputs'
    `requires_grad` flags even under `torch.no_grad()` (parameters always require grad), i.e. it is True in plain inference; and
    inside Function.forward the grad mode is always off.  The module that applies the Function records the caller's grad mode in
    `st.grad` (absent = assume a backward pass may follow)."""
    return bool(getattr(st, "grad", True)) and any(ctx.needs_input_grad)


def _can_fold(bn: BNState, ctx, st) -> bool:
    """eval-mode statistics and no gradient wanted from this call: the BN is a per-channel affine that folds into the conv's GEMM"""
    return _FOLD_BN and not bn.training and not _wants_grad(ctx, st)


def _lin_bn_folded(x, wb, bn: BNState, T, act=K.ACT_NONE, out_dtype=None, cache=None):
    """act(BN_eval(x W^T + b)) as ONE GEMM: the packed weight is scaled in place.  `cache` (frozen inference, ParamPacker.fold_cache):
    the fold of this layer is applied once per repack and its bias kept, keyed by the BatchNorm tensors' versions."""
    w, b = wb
    key = (w.data_ptr(), bn.running_mean._version, bn.running_var._version, bn.weight._version, bn.bias._version)
    if cache is not None and bn.prefix in cache and cache[bn.prefix][0] == key:
        bias = cache[bn.prefix][1]
    else:
        if cache is not None and bn.prefix in cache:
            raise RuntimeError("dformer_b200: BatchNorm buffers of %s changed under a frozen, already folded weight copy; call "
                               "dformer_b200.runtime.bump_weights_epoch() after editing buffers in place" % bn.prefix)
        bias = K.bn_fold(w, w.shape[1], b, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
        if cache is not None:
            cache[bn.prefix] = (key, bias)
    return K.gemm(x, w, trans_b=True, bias=bias, out_dtype=out_dtype or T, act=act, backend=backend_for(T), K=x.shape[1])


def _bn_bwd(dy, x2d, ms, bn: BNState, count, dx_dtype, dgamma, dbeta, act=K.ACT_NONE, residual=None, chan_scale=None, rows_per_sample=1, dy2=None):
    """returns (dx, g) with g = gradient w.r.t. the pre-activation (= gradient of the residual branch); dy2: second incoming gradient."""
    # local parameter gradients (DP averages them later) come out of the same pass as the two sums
    gbuf, sums = K.bn_bwd_reduce(dy, x2d, ms, bn.weight, bn.bias, residual, act, chan_scale, rows_per_sample, dbeta=dbeta, dgamma=dgamma, dy2=dy2)
    if bn.sync:
        small_all_reduce_(sums, bn.sync_group if bn.sync_group is not True else None)
    dx = K.bn_bwd_apply(gbuf, x2d, ms, bn.weight, sums, count, bn.training, dx_dtype)
    return dx, gbuf


# ============================================================================================ stem
class StemFn(torch.autograd.Function):
    """DFormer.py:194-211 -- conv3x3 s2 -> BN -> GELU -> conv3x3 s2 -> BN on an NCHW network input.
    Output: channels-last fp32 residual stream [B*H/4*W/4, C0]."""

    @staticmethod
    def forward(ctx, inp, st, w1, b1, g1, be1, w2, b2, g2, be2):
        T = st.dtype
        B, H, W = inp.shape[0], inp.shape[2], inp.shape[3]
        cin = st.cin
        H1, W1, H2, W2 = (H + 1) // 2, (W + 1) // 2, ((H + 1) // 2 + 1) // 2, ((W + 1) // 2 + 1) // 2
        pk1, pk2 = st.packed[st.g1], st.packed[st.g2]
        col1 = K.im2col_fwd(inp, (inp.stride(0), inp.stride(2), inp.stride(3), inp.stride(1)), B, H, W, cin, T, pk1[0].shape[1])
        if getattr(st, "ev_pack", None) is not None:
            torch.cuda.current_stream().wait_event(st.ev_pack)          # the packed weights (side stream) are first needed here
        if _can_fold(st.bn1, ctx, st) and _can_fold(st.bn2, ctx, st):          # inference: both BatchNorms ride in their conv's GEMM
            fc = getattr(st, "fold_cache", None)
            a1 = _lin_bn_folded(col1, pk1, st.bn1, T, act=K.ACT_GELU, cache=fc)
            cm = a1.shape[1]
            col2 = K.im2col_fwd(a1, (H1 * W1 * cm, W1 * cm, cm, 1), B, H1, W1, cm, T, pk2[0].shape[1])
            return _lin_bn_folded(col2, pk2, st.bn2, T, out_dtype=F32, cache=fc)
        c1 = _lin(col1, pk1, T)
        a1, ms1, n1 = _bn_fwd(c1, st.bn1, T, act=K.ACT_GELU)
        cm = a1.shape[1]
        col2 = K.im2col_fwd(a1, (H1 * W1 * cm, W1 * cm, cm, 1), B, H1, W1, cm, T, pk2[0].shape[1])
        c2 = _lin(col2, pk2, T)
        x0, ms2, n2 = _bn_fwd(c2, st.bn2, F32)
        ctx.st, ctx.dims = st, (B, H, W, H1, W1, H2, W2, cm)
        ctx.n = (n1, n2)
        ctx.save_for_backward(col1, c1, ms1, col2, c2, ms2)
        return x0

    @staticmethod
    def backward(ctx, dx0):
        st = ctx.st
        T = st.dtype
        col1, c1, ms1, col2, c2, ms2 = ctx.saved_tensors
        B, H, W, H1, W1, H2, W2, cm = ctx.dims
        n1, n2 = ctx.n
        ar: GradArena = st.arena
        p = st.prefix
        dx0 = dx0.contiguous()
        dc2, _ = _bn_bwd(dx0, c2, ms2, st.bn2, n2, T, ar.view(p + "4.weight"), ar.view(p + "4.bias"))
        pk1, pk2 = st.packed[st.g1], st.packed[st.g2]
        dW2p = torch.zeros(pk2[0].shape, device=dx0.device, dtype=F32)
        dcol2 = _lin_bwd(dc2, col2, pk2[0], dW2p, ar.view(p + "3.bias"), T)
        K.unpack_conv_grad(dW2p, pk2[0].shape[0], cm, ar.view(p + "3.weight"))
        da1 = K.im2col_bwd(dcol2, B, H1, W1, cm, T)
        # a1 = GELU(BN1(c1)) is recomputed inside the BN backward (pre-activation from c1)
        dc1, _ = _bn_bwd(da1, c1, ms1, st.bn1, n1, T, ar.view(p + "1.weight"), ar.view(p + "1.bias"), act=K.ACT_GELU)
        dW1p = torch.zeros(pk1[0].shape, device=dx0.device, dtype=F32)
        _lin_bwd(dc1, col1, pk1[0], dW1p, ar.view(p + "0.bias"), T, need_dx=False)
        K.unpack_conv_grad(dW1p, pk1[0].shape[0], st.cin, ar.view(p + "0.weight"))
        if getattr(st, "wstream", None) is not None:
            K.join(st.wstream)                        # every Block's parameter gradients (wgrad side stream) are final from here on
        ar.done(st.tag)
        return (None, None, ar.view(p + "0.weight"), ar.view(p + "0.bias"), ar.view(p + "1.weight"), ar.view(p + "1.bias"),
                ar.view(p + "3.weight"), ar.view(p + "3.bias"), ar.view(p + "4.weight"), ar.view(p + "4.bias"))


# ============================================================================================ downsample
class DownsampleFn(torch.autograd.Function):
    """DFormer.py:216-228 -- BN -> conv3x3 s2 on the channels-last fp32 stage output [B*H*W, Cin]."""

    @staticmethod
    def forward(ctx, x, st, g, be, w, b):
        T = st.dtype
        B, H, W = st.B, st.H, st.W
        cin = x.shape[1]
        xb, ms, n = _bn_fwd(x, st.bn, T)
        pk = st.packed[st.g]
        col = K.im2col_fwd(xb, (H * W * cin, W * cin, cin, 1), B, H, W, cin, T, pk[0].shape[1])
        y = _lin(col, pk, T, out_dtype=F32)
        ctx.st, ctx.n = st, n
        ctx.save_for_backward(x, ms, col)
        return y

    @staticmethod
    def backward(ctx, dy):
        st = ctx.st
        T = st.dtype
        x, ms, col = ctx.saved_tensors
        ar: GradArena = st.arena
        p = st.prefix
        cin = x.shape[1]
        dyT = dy.contiguous() if T == F32 else K.cast(dy.contiguous(), T)
        pk = st.packed[st.g]
        dWp = torch.zeros(pk[0].shape, device=dy.device, dtype=F32)
        dcol = _lin_bwd(dyT, col, pk[0], dWp, ar.view(p + "1.bias"), T)
        K.unpack_conv_grad(dWp, pk[0].shape[0], cin, ar.view(p + "1.weight"))
        dxb = K.im2col_bwd(dcol, st.B, st.H, st.W, cin, T)
        dx, _ = _bn_bwd(dxb, x, ms, st.bn, ctx.n, F32, ar.view(p + "0.weight"), ar.view(p + "0.bias"))
        ar.done(st.tag)
        return dx, None, ar.view(p + "0.weight"), ar.view(p + "0.bias"), ar.view(p + "1.weight"), ar.view(p + "1.bias")


# ============================================================================================ Block
BLOCK_PARAM_ORDER_DOC = "see models/encoders/DFormer.py: Block.param_names()"


def _mlp_fwd(x, pfx, st, P, sv, scale_b, pre=None):
    """DFormer.py:58-67 + layer-scale residual :176/:179.  x fp32 [M,C] -> fp32 [M,C].
    pre = (branch, layer_scale, drop_path_scale): the attention branch's residual x + dp * ls * branch (:173-175) is applied
    in the same pass as this MLP's LayerNorm."""
    T = st.dtype
    B, H, W = st.B, st.H, st.W
    if pre is not None:
        x, hn, mu, rs = K.scale_residual_layernorm_fwd(x, pre[0], pre[1], pre[2], H * W, P[pfx + "norm.weight"], P[pfx + "norm.bias"], 1e-6)
    else:
        hn, mu, rs = K.layernorm_fwd(x, P[pfx + "norm.weight"], P[pfx + "norm.bias"], 1e-6, T)
    h = _lin(hn, st.packed[st.key + pfx + "fc1"], T)
    if T == torch.bfloat16:                    # TMA-fed fused kernel (csrc/mlp_dw.cu); training keeps GELU'(z) (bf16) for a streaming backward
        if sv.get("_bwd", True):
            u, z = K.mlp_dw_fwd(h, P[pfx + "pos.weight"], P[pfx + "pos.bias"], B, H, W, save_gp=True)      # z slot holds GELU'(z)
        else:
            u, z = K.mlp_dw_fwd(h, P[pfx + "pos.weight"], P[pfx + "pos.bias"], B, H, W), None
    else:
        u, z = K.dwconv_fwd(h, P[pfx + "pos.weight"], P[pfx + "pos.bias"], B, H, W, 3, add_input=True, act=K.ACT_GELU, save_z=True)
    ls = P["layer_scale_2" if pfx == "mlp." else "layer_scale_2_e"]
    f = _lin(u, st.packed[st.key + pfx + "fc2"], T)
    carry = getattr(st, "carry", None)
    if carry is not None:
        # not the last Block of its stage: the residual x + dp * ls * f (:176/:179) is applied by the NEXT Block's first LayerNorm
        # kernel (`_carried_layernorm`), which writes it into the buffer returned here -- one pass over x less per stream and Block
        out = torch.empty_like(x)
        carry[pfx] = (x, f, ls, scale_b, out)
    else:
        out = K.scale_residual_fwd(x, f, ls, scale_b, H * W)
    sv.update({pfx + "x": x, pfx + "mu": mu, pfx + "rs": rs, pfx + "hn": hn, pfx + "h": h, pfx + "u": u, pfx + "f": f, pfx + "z": z})
    return out


def _mlp_bwd(dout, pfx, st, P, sv, scale_b, G):
    """returns d(input of the MLP residual branch) = dout + LN-path gradient (fp32)."""
    T = st.dtype
    B, H, W = st.B, st.H, st.W
    lsn = "layer_scale_2" if pfx == "mlp." else "layer_scale_2_e"
    fused = T == torch.bfloat16
    df = K.scale_residual_bwd(dout, sv[pfx + "f"], P[lsn], scale_b, H * W, G[lsn], dy_colsum=G[pfx + "fc2.bias"] if fused else None)
    w2 = st.packed[st.key + pfx + "fc2"][0]
    if fused:                                 # GELU' . dw3x3^T . weight/bias gradients . fc1 bias gradient: one kernel, dz stays on chip
        du = _lin_bwd(df, sv[pfx + "u"], w2, G[pfx + "fc2.weight"], None, T)                # fc2 bias gradient came with df above
        dh = K.mlp_dw_bwd(du, sv[pfx + "h"], P[pfx + "pos.weight"], P[pfx + "pos.bias"], B, H, W, G[pfx + "pos.weight"], G[pfx + "pos.bias"],
                          G[pfx + "fc1.bias"], gp=sv[pfx + "z"])
        dhn = _lin_bwd(dh, sv[pfx + "hn"], st.packed[st.key + pfx + "fc1"][0], G[pfx + "fc1.weight"], None, T)
        return K.layernorm_bwd(dhn, sv[pfx + "x"], P[pfx + "norm.weight"], sv[pfx + "mu"], sv[pfx + "rs"], dout,
                               G[pfx + "norm.weight"], G[pfx + "norm.bias"])
    du = _lin_bwd(df, sv[pfx + "u"], w2, G[pfx + "fc2.weight"], G[pfx + "fc2.bias"], T)
    dh = K.dwconv_bwd(du, sv[pfx + "h"], P[pfx + "pos.weight"], P[pfx + "pos.bias"], B, H, W, 3, True, K.ACT_GELU,
                      G[pfx + "pos.weight"], G[pfx + "pos.bias"], z=sv[pfx + "z"], wgrad_stream=_WGRAD_STREAM)
    w1 = st.packed[st.key + pfx + "fc1"][0]
    dhn = _lin_bwd(dh, sv[pfx + "hn"], w1, G[pfx + "fc1.weight"], G[pfx + "fc1.bias"], T)
    return K.layernorm_bwd(dhn, sv[pfx + "x"], P[pfx + "norm.weight"], sv[pfx + "mu"], sv[pfx + "rs"], dout,
                           G[pfx + "norm.weight"], G[pfx + "norm.bias"])


def _carried_layernorm(x, pfx, st, gamma, beta):
    """First LayerNorm of a Block (DFormer.py:104-105).  When the previous Block of the stage left its MLP residual pending
    (`_mlp_fwd`), x is that Block's still unwritten output buffer: the residual is formed, stored into x and normalised in one pass."""
    pend = st.pending.pop(pfx, None) if getattr(st, "pending", None) is not None else None
    if pend is None:
        return K.layernorm_fwd(x, gamma, beta, 1e-6, st.dtype)
    res, f, ls, scale_b, out = pend
    assert out.data_ptr() == x.data_ptr() and out.shape == x.shape, "pending residual does not belong to this Block's input"
    _, xn, mu, rs = K.scale_residual_layernorm_fwd(res, f, ls, scale_b, st.H * st.W, gamma, beta, 1e-6, out=out)
    return xn, mu, rs


class BlockFn(torch.autograd.Function):
    """DFormer.py:147-181 (Block) with Attention :102-145 and both MLPs :58-67; x, x_e are the fp32
    channels-last residual streams [M, C], [M, C/2].

    The depth-stream work that does not depend on the RGB stream (LN_e -> e_fore -> dw7x7 -> e_back, and the whole
    depth MLP) is launched on a side CUDA stream: at stages 2-3 the kernels are too small to fill 148 SMs one at a
    time, and under graph capture the two branches become parallel graph branches."""

    @staticmethod
    def forward(ctx, x, x_e, st, *params):
        T = st.dtype
        P = dict(zip(st.names, params))
        B, H, W, C = st.B, st.H, st.W, st.C
        Ce, M, HW = C // 2, x.shape[0], st.H * st.W
        win, dd = st.window != 0, st.drop_depth
        pk = lambda n: st.packed[st.key + n]
        sv = {"_bwd": _wants_grad(ctx, st)}               # inference: nothing is kept for a backward pass
        side = st.side
        main = torch.cuda.current_stream()
        # bf16: the gating products q * a / cut * e (:134-135) are epilogues of the GEMMs that produce a / e (csrc/gemm_tc.cu "gate"):
        # they land directly in their column slices of the concat buffer y (:137-140)
        fuse = T == torch.bfloat16
        keep = sv["_bwd"]
        ycols = 2 * C if win else C + Ce
        y = torch.empty((M, ycols), device=x.device, dtype=T)
        # ---- depth gate path on the side stream
        K.fork(side)
        with torch.cuda.stream(side):
            en, mu2, rs2 = _carried_layernorm(x_e, "mlp_e2.", st, P["attn.norm_e.weight"], P["attn.norm_e.bias"])
            ev_en = K.signal(side)
            ef = _lin(en, pk("attn.e_fore"), T)
            ec = K.dwconv_fwd(ef, P["attn.e_conv.weight"], P["attn.e_conv.bias"], B, H, W, 7)
            if not fuse:
                e = _lin(ec, pk("attn.e_back"), T)
                ev_e = K.signal(side)
        # ---- RGB path
        xn, mu1, rs1 = _carried_layernorm(x, "mlp.", st, P["attn.norm.weight"], P["attn.norm.bias"])
        if win:                                                       # global-awareness branch on a second side stream:
            side2 = st.side2                                          # the pooled queries only need the two LayerNorm outputs, so they
            K.fork(side2)                                             # are computed while the main stream runs the q|cut|l GEMM
            with torch.cuda.stream(side2):
                side2.wait_event(ev_en)
                pooled = K.pool7_fwd(xn, en, B, H, W)
                m = _lin(pooled, pk("attn.short_cut_linear"), T)
        qcl = _lin(xn, pk("attn.qcl"), T)                                         # [M, 2.5C] = z_l | q | cut
        l = K.act_fwd(qcl[:, :C], K.ACT_GELU)
        if fuse:
            ev_qcl = K.signal()
            with torch.cuda.stream(side):                             # e_back needs `cut` as its gate
                side.wait_event(ev_qcl)
                e = torch.empty((M, Ce), device=x.device, dtype=T) if keep else None
                _lin(ec, pk("attn.e_back"), T, out=y[:, ycols - Ce:], gate=qcl[:, 2 * C:], out2=e)
                ev_e = K.signal(side)
            K.share(side, y, qcl)
        K.share(main, en, mu2, rs2, ef, ec, e)
        if win:
            K.fork(side2)
            with torch.cuda.stream(side2):
                kv = _lin(l, pk("attn.kv"), T)
                dh_ = Ce // st.heads
                if dh_ in K.GAA_FUSED_DIMS:                         # one launch (tensor cores in bf16); keeps only the row log-sum-exp
                    o7, lse7 = K.gaa_fused_fwd(m, kv, B, HW, st.heads, dh_)
                    probs = None
                else:
                    o7, probs = K.gaa_fwd(m, kv, B, HW, st.heads, dh_)
                    lse7 = None
                K.resize_fwd(o7, B, 7, 7, y, H, W, col0=C)
            K.share(main, kv, pooled, m, probs, o7, lse7)
            K.share(side2, y, l)
            sv.update(kv=kv, pooled=pooled, m=m, probs=probs, o7=o7, lse7=lse7)
        cv = K.dwconv_fwd(l, P["attn.conv.weight"], P["attn.conv.bias"], B, H, W, 7)
        if fuse:
            a = torch.empty((M, C), device=x.device, dtype=T) if keep else None
            _lin(cv, pk("attn.a"), T, out=y[:, :C], gate=qcl[:, C:2 * C], out2=a)
        else:
            a = _lin(cv, pk("attn.a"), T)
            K.mul_fwd(qcl[:, C:2 * C], a, y[:, :C])
        if win:
            K.join(side2)
        main.wait_event(ev_e)
        if not fuse:
            K.mul_fwd(qcl[:, 2 * C:], e, y[:, ycols - Ce:])
        pp = _lin(y, pk("attn.pp"), T)                                            # [M, C (+Ce)] = proj | proj_e
        sv.update(x=x, x_e=x_e, mu1=mu1, rs1=rs1, mu2=mu2, rs2=rs2, xn=xn, en=en, qcl=qcl, l=l, cv=cv, a=a, y=y, ef=ef, ec=ec, e=e, pp=pp)
        # ---- MLPs: depth stream on the side stream, RGB stream on the main stream
        if not dd:
            K.fork(side)
            with torch.cuda.stream(side):
                n_before = set(sv)
                xe2 = _mlp_fwd(x_e, "mlp_e2.", st, P, sv, st.dp[3], pre=(pp[:, C:], P["layer_scale_1_e"], st.dp[2]))
            K.share(main, xe2, *[sv[k] for k in sv if k not in n_before])
        else:
            xe2 = x_e
        x2 = _mlp_fwd(x, "mlp.", st, P, sv, st.dp[1], pre=(pp[:, :C], P["layer_scale_1"], st.dp[0]))
        K.join(side)
        ctx.st, ctx.sv, ctx.P = st, sv, P
        return x2, xe2

    @staticmethod
    def backward(ctx, dx2, dxe2):
        st, sv = ctx.st, ctx.sv
        T = st.dtype
        P = ctx.P
        ar: GradArena = st.arena
        G = {n: ar.view(st.prefix + n) for n in st.names}           # allocates / zero-fills the arena on the main stream
        B, H, W, C = st.B, st.H, st.W, st.C
        Ce, HW = C // 2, st.H * st.W
        M = sv["x"].shape[0]
        win, dd = st.window != 0, st.drop_depth
        pk = lambda n: st.packed[st.key + n]
        dev = sv["x"].device
        side, side2, wstream = st.side, st.side2, st.wstream
        main = torch.cuda.current_stream()
        global _WGRAD_STREAM
        _WGRAD_STREAM = wstream
        K.fork(wstream)                                             # wgrad stream starts after the arena memset / previous module
        dx2 = dx2.contiguous()
        ppw = pk("attn.pp")[0]
        dpp = torch.empty((M, ppw.shape[0]), device=dev, dtype=T)
        # ---- MLPs (reverse): depth stream on the side stream
        dxe1 = None
        if not dd:
            dxe2 = dxe2.contiguous() if dxe2 is not None else torch.zeros_like(sv["x_e"])
            K.fork(side)
            with torch.cuda.stream(side):
                dxe1 = _mlp_bwd(dxe2, "mlp_e2.", st, P, sv, st.dp[3], G)
                K.scale_residual_bwd(dxe1, sv["pp"][:, C:], P["layer_scale_1_e"], st.dp[2], HW, G["layer_scale_1_e"], dy=dpp[:, C:],
                                     dy_colsum=G["attn.proj_e.bias"])                       # proj_e bias gradient in the same pass
                ev_e = K.signal(side)
            K.share(main, dxe1)
        dx1 = _mlp_bwd(dx2, "mlp.", st, P, sv, st.dp[1], G)
        K.scale_residual_bwd(dx1, sv["pp"][:, :C], P["layer_scale_1"], st.dp[0], HW, G["layer_scale_1"], dy=dpp[:, :C],
                             dy_colsum=G["attn.proj.bias"])                                 # proj bias gradient in the same pass
        if not dd:
            main.wait_event(ev_e)
        # ---- proj | proj_e
        if dd:
            dWpp = G["attn.proj.weight"]
        else:
            dWpp = ar.span(st.prefix + "attn.proj.weight", st.prefix + "attn.proj_e.weight", ppw.shape)
        dy = _lin_bwd(dpp, sv["y"], ppw, dWpp, None, T)
        ycols = dy.shape[1]
        qcl = sv["qcl"]
        dqcl = torch.empty_like(qcl)
        da = torch.empty((M, C), device=dev, dtype=T)
        de = torch.empty((M, Ce), device=dev, dtype=T)
        # the element-wise gradient kernels below also emit the bias gradients of q | q_cut | l, a and e_back (column sums of what they write)
        K.mul_bwd(dy[:, ycols - Ce:], qcl[:, 2 * C:], sv["e"], dqcl[:, 2 * C:], de, G["attn.q_cut.bias"], G["attn.e_back.bias"])
        # ---- depth gate path e = e_back(dw7(e_fore(en))) on the side stream
        K.fork(side)
        with torch.cuda.stream(side):
            dec = _lin_bwd(de, sv["ec"], pk("attn.e_back")[0], G["attn.e_back.weight"], None, T)
            def_ = K.dwconv_bwd(dec, sv["ef"], P["attn.e_conv.weight"], P["attn.e_conv.bias"], B, H, W, 7, False, K.ACT_NONE,
                                G["attn.e_conv.weight"], G["attn.e_conv.bias"], wgrad_stream=_WGRAD_STREAM)
            den = _lin_bwd(def_, sv["en"], pk("attn.e_fore")[0], G["attn.e_fore.weight"], G["attn.e_fore.bias"], T)
        K.share(main, den)
        # ---- global-awareness branch (resize^T -> attention -> pooled queries / kv) on a second side stream
        dxn_pool = den_pool = dl_kv = None
        if win:
            K.fork(side2)
            with torch.cuda.stream(side2):
                do7 = torch.empty((B * 49, Ce), device=dev, dtype=F32)
                K.resize_bwd(dy, C, B, 7, 7, Ce, H, W, do7)
                bias_done = False
                if sv["probs"] is None and T == torch.bfloat16:      # tensor-core kernel: also emits both bias gradients and dm in bf16
                    dmT, dkv = K.gaa_fused_bwd_ex(do7, sv["o7"], sv["lse7"], sv["m"], sv["kv"], B, HW, st.heads, Ce // st.heads,
                                                  G["attn.kv.bias"], G["attn.short_cut_linear.bias"])
                    bias_done = True
                else:
                    if sv["probs"] is None:
                        dm, dkv = K.gaa_fused_bwd(do7, sv["o7"], sv["lse7"], sv["m"], sv["kv"], B, HW, st.heads, Ce // st.heads)
                    else:
                        dm, dkv = K.gaa_bwd(do7, sv["m"], sv["kv"], sv["probs"], B, HW, st.heads, Ce // st.heads)
                    dmT = dm if T == F32 else K.cast(dm, T)
                dl_kv = _lin_bwd(dkv, sv["l"], pk("attn.kv")[0], G["attn.kv.weight"], None if bias_done else G["attn.kv.bias"], T)
                ev_dlkv = K.signal(side2)                 # the main stream needs dl_kv early (GELU' of l); the pooled-query
                # gradients below are only consumed by the final LayerNorm backward
                dpooled = _lin_bwd(dmT, sv["pooled"], pk("attn.short_cut_linear")[0], G["attn.short_cut_linear.weight"],
                                   None if bias_done else G["attn.short_cut_linear.bias"], T)
                dxn_pool, den_pool = K.pool7_bwd(dpooled, C, Ce, B, H, W)
            K.share(main, dxn_pool, den_pool, dl_kv)
        # ---- RGB path: a = a(dw7(l))
        K.mul_bwd(dy[:, :C], qcl[:, C:2 * C], sv["a"], dqcl[:, C:2 * C], da, G["attn.q.bias"], G["attn.a.bias"])
        dcv = _lin_bwd(da, sv["cv"], pk("attn.a")[0], G["attn.a.weight"], None, T)
        dl = K.dwconv_bwd(dcv, sv["l"], P["attn.conv.weight"], P["attn.conv.bias"], B, H, W, 7, False, K.ACT_NONE,
                          G["attn.conv.weight"], G["attn.conv.bias"], wgrad_stream=_WGRAD_STREAM)
        if win:
            main.wait_event(ev_dlkv)
        K.act_bwd(dl, qcl[:, :C], K.ACT_GELU, out=dqcl[:, :C], dout2=dl_kv, colsum=G["attn.l.bias"])      # kv branch's gradient of l joins here
        qclw = pk("attn.qcl")[0]
        dWq = ar.span(st.prefix + "attn.l.weight", st.prefix + "attn.q_cut.weight", qclw.shape)
        dxn = _lin_bwd(dqcl, sv["xn"], qclw, dWq, None, T)
        # the pooled-query branch's gradient joins inside the LayerNorm backward kernels (dy2)
        if win:
            K.join(side2)
        dx = K.layernorm_bwd(dxn, sv["x"], P["attn.norm.weight"], sv["mu1"], sv["rs1"], dx1, G["attn.norm.weight"], G["attn.norm.bias"],
                             dy2=dxn_pool)
        K.join(side)
        dxe = K.layernorm_bwd(den, sv["x_e"], P["attn.norm_e.weight"], sv["mu2"], sv["rs2"], dxe1, G["attn.norm_e.weight"], G["attn.norm_e.bias"],
                              dy2=den_pool)
        if dd and dxe2 is not None:
            K.axpy(dxe2.contiguous(), 1.0, dxe)                                  # x_e passes through the last block unchanged
        _WGRAD_STREAM = None
        ctx.sv = ctx.P = None
        # the wgrad stream is NOT joined here (it may trail the dgrad chain across Blocks); the DP engine gets an event,
        # and the stem's backward -- the last encoder node on the main stream -- performs the final join
        ar.done(st.tag, K.signal(wstream))
        return (dx, dxe, None) + tuple(G[n] for n in st.names)


# ============================================================================================ NMF2D
def nmf_prepare(bases_raw, T):
    """`F.normalize(bases, dim=1)` of the freshly drawn bases (ham_head.py:111-115) + their compute-dtype copy.  Depends on nothing the
    network computes, so the head issues it on its side stream under the resize / squeeze kernels (LightHamHead.forward)."""
    bases, _ = K.normalize_cols(bases_raw)
    return bases, (bases if T == F32 else K.cast(bases, torch.bfloat16))


def _nmf_fwd(x, bases_raw, steps, T, side=None, prepared=None):
    """ham_head.py:60-100,109-145 on channels-last x [B, N, D] (compute dtype); bases_raw [B, D, R] fp32.

    The factor state (coef, bases) and every multiplicative update stay fp32; in bf16 mode the batched matrix
    products run on the tcgen05 GEMM with bf16 copies of the factors (what the reference's autocast bmm does).
    `side`: a second stream for the small Gram products (B^T B, C^T C), which only depend on one factor: the chain is
    latency-bound (5-8 us launches on a few SMs), so every launch taken off the critical path is time saved."""
    B, N, D = x.shape
    R = bases_raw.shape[2]
    dev = x.device
    f = lambda *s: torch.empty(s, device=dev, dtype=F32)
    lo = (lambda t: t) if T == F32 else (lambda t: K.cast(t, torch.bfloat16))
    sk = max(1, N // 512)
    main = torch.cuda.current_stream()

    def on_side(fn):
        """run fn() on the side stream after everything enqueued so far; returns (result, event to wait for)"""
        if side is None:
            return fn(), None
        K.fork(side)
        with torch.cuda.stream(side):
            r = fn()
            ev = K.signal(side)
        K.share(main, *(r if isinstance(r, tuple) else (r,)))
        return r, ev

    def wait(ev):
        if ev is not None:
            main.wait_event(ev)

    # split-K products accumulate with fp32 reductions into zeroed outputs: ONE zero-filled pool per call (one memset node) instead of
    # one memset in front of every such GEMM of the latency-bound chain
    zpool = torch.zeros(max(1, steps) * (B * D * R + B * R * R), device=dev, dtype=F32) if sk > 1 else None
    zoff = [0]

    def fz(*shape):
        if zpool is None:
            return f(*shape)
        n = 1
        for d in shape:
            n *= d
        v = zpool[zoff[0]:zoff[0] + n].view(*shape)
        zoff[0] += n
        return v

    acc_z = zpool is not None                                   # accumulate into the pre-zeroed view (no memset launched by the GEMM)
    bases, bases_l = prepared if prepared is not None else nmf_prepare(bases_raw, T)
    S = K.bgemm(x, bases_l, f(B, N, R), M=N, N=R, K=D)
    coef = K.softmax_rows(S)
    coef_l = lo(coef)
    tape = []

    lo_out = lambda *s: torch.empty(s, device=dev, dtype=T)     # short-K products land directly in the compute dtype (no split-K, no cast)

    def coef_update(coef, coef_l, bases, bases_l):
        btb_l, ev = on_side(lambda: K.bgemm(bases_l, bases_l, lo_out(B, R, R), trans_a=True, M=R, N=R, K=D))
        num = K.bgemm(x, bases_l, f(B, N, R), M=N, N=R, K=D)
        wait(ev)
        den = K.bgemm(coef_l, btb_l, f(B, N, R), M=N, N=R, K=R)
        coef_n, coef_nl = K.mu_update(coef, num, den, lo_dtype=T)          # fp32 state + compute-dtype operand from one launch
        return coef_n, coef_nl, (coef, coef_l, num, den, bases_l, btb_l)

    for _ in range(steps):
        coef_n, coef_nl, rec_c = coef_update(coef, coef_l, bases, bases_l)
        ctc_buf, num2_buf = fz(B, R, R), fz(B, D, R)
        ctc_l, ev = on_side(lambda: lo(K.bgemm(coef_nl, coef_nl, ctc_buf, trans_a=True, M=R, N=R, K=N, splitk=sk, accumulate=acc_z)))
        num2 = K.bgemm(x, coef_nl, num2_buf, trans_a=True, M=D, N=R, K=N, splitk=sk, accumulate=acc_z)
        wait(ev)
        den2 = K.bgemm(bases_l, ctc_l, f(B, D, R), M=D, N=R, K=R)
        bases_n, bases_nl = K.mu_update(bases, num2, den2, lo_dtype=T)
        tape.append((rec_c, (bases, bases_l, num2, den2, coef_nl, ctc_l)))
        coef, coef_l, bases, bases_l = coef_n, coef_nl, bases_n, bases_nl
    coef_f, coef_fl, rec_f = coef_update(coef, coef_l, bases, bases_l)
    out = torch.empty((B, N, D), device=dev, dtype=T)
    K.bgemm(coef_fl, bases_l, out, trans_b=True, M=N, N=D, K=R)
    return out, (tape, rec_f, coef_fl, bases_l)


def _nmf_bwd(dout, x, saved, T, side=None):
    """Back-propagation through every multiplicative update (the reference does not detach them, ham_head.py:45,119).

    The gradient w.r.t. x is a sum of 2*steps + 2 rank-R products (one per use of x in a numerator, plus the initial
    softmax).  Their factors are gathered side by side along the reduction dimension -- A_cat [B, N, slots*R],
    B_cat [B, D, slots*R] -- and contracted by ONE GEMM with K = slots*R that writes dx once, instead of `slots`
    read-modify-write passes over an fp32 [B, N, D] accumulator.
    The gradient of a Gram matrix G = F^T F enters F twice (F dG + F dG^T): the two products are one product with the
    symmetrised dG; the Gram-gradient products run on `side` next to the long x^T (.) / x (.) products."""
    tape, rec_f, coef_fl, bases_Tl = saved
    B, N, D = x.shape
    R = coef_fl.shape[2]
    dev = x.device
    f = lambda *s: torch.empty(s, device=dev, dtype=F32)
    slots = 2 * len(tape) + 2
    a_cat = torch.empty((B, N, slots * R), device=dev, dtype=T)
    b_cat = torch.empty((B, D, slots * R), device=dev, dtype=T)
    slot = [0]
    main = torch.cuda.current_stream()

    def on_side(fn):
        if side is None:
            return fn(), None
        K.fork(side)
        with torch.cuda.stream(side):
            r = fn()
            ev = K.signal(side)
        K.share(main, r)
        return r, ev

    def wait(ev):
        if ev is not None:
            main.wait_event(ev)

    def push(a_src, b_src):
        """register one product  dx += a[B,N,R] @ b[B,D,R]^T : returns the two column slices of the concatenated operands; a source
        given as None is written into its slice by the caller (mu_update_bwd emits dnum straight into it)"""
        k0 = slot[0] * R
        slot[0] += 1
        av, bv = a_cat[:, :, k0:k0 + R], b_cat[:, :, k0:k0 + R]
        if a_src is not None:
            K.cast_into(a_src, av)
        if b_src is not None:
            K.cast_into(b_src, bv)
        return av, bv

    sk = max(1, N // 512)
    # one zero-filled pool for the split-K outputs of the chain (see _nmf_fwd)
    zpool = torch.zeros(B * D * R + (len(tape) + 1) * B * R * R, device=dev, dtype=F32) if sk > 1 else None
    zoff = [0]

    def fz(*shape):
        if zpool is None:
            return f(*shape)
        n = 1
        for d in shape:
            n *= d
        v = zpool[zoff[0]:zoff[0] + n].view(*shape)
        zoff[0] += n
        return v

    acc_z = zpool is not None
    # out = coef_f @ bases^T
    dcoef = K.bgemm(dout, bases_Tl, f(B, N, R), M=N, N=R, K=D)
    dbases = K.bgemm(dout, coef_fl, fz(B, D, R), trans_a=True, M=D, N=R, K=N, splitk=sk, accumulate=acc_z)

    def coef_update_bwd(dcoef_new, rec, dbases):
        coef, coef_l, num, den, bases_l, btb_l = rec
        dco = f(B, N, R)
        dnum_l, _ = push(None, bases_l)                                                     # num = x @ bases
        dden_l = K.mu_update_bwd(dcoef_new, coef, num, den, dco, False, dnum_l, T)
        # BtB = bases^T bases: d(BtB) = coef^T dden, symmetrised, on the side stream
        dbtb_buf = fz(B, R, R)
        dbtb_s, ev = on_side(lambda: K.sym_cast(K.bgemm(coef_l, dden_l, dbtb_buf, trans_a=True, M=R, N=R, K=N, splitk=sk, accumulate=acc_z), T))
        K.bgemm(x, dnum_l, dbases, trans_a=True, M=D, N=R, K=N, accumulate=True, splitk=sk)
        K.bgemm(dden_l, btb_l, dco, M=N, N=R, K=R, accumulate=True)                         # den = coef @ BtB (BtB symmetric)
        wait(ev)
        K.bgemm(bases_l, dbtb_s, dbases, M=D, N=R, K=R, accumulate=True)
        return dco

    def bases_update_bwd(dbases_new, rec, dcoef):
        bases, bases_l, num2, den2, coef_l, ctc_l = rec
        dba = f(B, D, R)
        _, dnum2_l = push(coef_l, None)                                                     # num2 = x^T @ coef
        dden2_l = K.mu_update_bwd(dbases_new, bases, num2, den2, dba, False, dnum2_l, T)
        # CtC = coef^T coef: d(CtC) = bases^T dden2, symmetrised, on the side stream
        dctc_s, ev = on_side(lambda: K.sym_cast(K.bgemm(bases_l, dden2_l, f(B, R, R), trans_a=True, M=R, N=R, K=D), T))
        K.bgemm(x, dnum2_l, dcoef, M=N, N=R, K=D, accumulate=True)
        K.bgemm(dden2_l, ctc_l, dba, M=D, N=R, K=R, accumulate=True)                        # den2 = bases @ CtC
        wait(ev)
        K.bgemm(coef_l, dctc_s, dcoef, M=N, N=R, K=R, accumulate=True)
        return dba

    dcoef = coef_update_bwd(dcoef, rec_f, dbases)
    for rec_c, rec_b in reversed(tape):
        dbases = bases_update_bwd(dbases, rec_b, dcoef)
        dcoef = coef_update_bwd(dcoef, rec_c, dbases)
    # coef0 = softmax(x @ bases0)
    first = tape[0][0] if tape else rec_f
    coef0, bases0_l = first[0], first[4]
    dS = K.softmax_rows_bwd(dcoef, coef0)
    push(dS, bases0_l)
    assert slot[0] == slots
    dx = torch.empty((B, N, D), device=dev, dtype=T)
    K.bgemm(a_cat, b_cat, dx, trans_b=True, M=N, N=D, K=slots * R)
    return dx


# ============================================================================================ LightHamHead
class HeadFn(torch.autograd.Function):
    """ham_head.py:222-240 (LightHamHead.forward), :173-180 (Hamburger), decode_head.py:226-231 (cls_seg).
    Inputs: the three channels-last fp32 stage outputs; output: channels-last logits [B*h*w, ncls]."""

    @staticmethod
    def forward(ctx, o1, o2, o3, bases_raw, st, *params):
        T = st.dtype
        P = dict(zip(st.names, params))
        B = st.B
        (h1, w1), (h2, w2), (h3, w3) = st.sizes
        C1, C2, C3 = o1.shape[1], o2.shape[1], o3.shape[1]
        M = o1.shape[0]
        pk = lambda n: st.packed[n]
        cat = torch.empty((M, C1 + C2 + C3), device=o1.device, dtype=T)
        K.resize_fwd(o1, B, h1, w1, cat, h1, w1, col0=0)
        K.resize_fwd(o2, B, h2, w2, cat, h1, w1, col0=C1)
        K.resize_fwd(o3, B, h3, w3, cat, h1, w1, col0=C1 + C2)
        if getattr(st, "ev_prologue", None) is not None:         # weight packing / Dropout2d mask / normalised bases from the side stream
            torch.cuda.current_stream().wait_event(st.ev_prologue)
        fold = _can_fold(st.bn_sq, ctx, st) and _can_fold(st.bn_al, ctx, st) and st.drop_mask is None
        if fold:
            s_pre, ms_s, n_s = None, None, 0
            s = _lin_bn_folded(cat, pk("squeeze"), st.bn_sq, T, act=K.ACT_RELU, cache=getattr(st, "fold_cache", None))
        else:
            s_pre = _lin(cat, pk("squeeze"), T)
            s, ms_s, n_s = _bn_fwd(s_pre, st.bn_sq, T, act=K.ACT_RELU)
        hin = _lin(s, pk("ham_in"), T, act=K.ACT_RELU)
        D = hin.shape[1]
        nmf, nmf_saved = _nmf_fwd(hin.view(B, h1 * w1, D), bases_raw, st.steps, T, side=getattr(st, "side", None),
                                  prepared=getattr(st, "bases_prepared", None))
        nmf2d = nmf.view(M, D)
        ho_pre = _lin(nmf2d, pk("ham_out"), T)
        hs, ms_o, n_o = _bn_fwd(ho_pre, st.bn_out, T, act=K.ACT_RELU, residual=s)
        if fold:
            al_pre, ms_a, n_a = None, None, 0
            al = _lin_bn_folded(hs, pk("align"), st.bn_al, T, act=K.ACT_RELU, cache=getattr(st, "fold_cache", None))
        else:
            al_pre = _lin(hs, pk("align"), T)
            al, ms_a, n_a = _bn_fwd(al_pre, st.bn_al, T, act=K.ACT_RELU, chan_scale=st.drop_mask, rows_per_sample=h1 * w1)
        logits = _lin(al, pk("conv_seg"), T)
        ctx.st = st
        ctx.sv = dict(cat=cat, s_pre=s_pre, ms_s=ms_s, n_s=n_s, s=s, hin=hin, nmf_saved=nmf_saved, nmf=nmf2d, ho_pre=ho_pre, ms_o=ms_o,
                      n_o=n_o, hs=hs, al_pre=al_pre, ms_a=ms_a, n_a=n_a, al=al, shapes=(C1, C2, C3, M, D))
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        st, sv = ctx.st, ctx.sv
        T = st.dtype
        ar: GradArena = st.arena
        G = {n: ar.view(st.prefix + n) for n in st.names}
        # the five weight gradients of the head are leaves: they run on a side stream under the long, latency-bound NMF chain
        global _WGRAD_STREAM
        ws = getattr(st, "wstream", None)
        if ws is not None:
            K.fork(ws)
            _WGRAD_STREAM = ws
        B = st.B
        (h1, w1), (h2, w2), (h3, w3) = st.sizes
        C1, C2, C3, M, D = sv["shapes"]
        pk = lambda n: st.packed[n]
        dev = dlogits.device
        dlogits = dlogits.contiguous()
        if dlogits.dtype != T:
            dlogits = K.cast(dlogits, T)
        dal = _lin_bwd(dlogits, sv["al"], pk("conv_seg")[0], G["conv_seg.weight"].view(pk("conv_seg")[0].shape), G["conv_seg.bias"], T)
        dal_pre, _ = _bn_bwd(dal, sv["al_pre"], sv["ms_a"], st.bn_al, sv["n_a"], T, G["align.bn.weight"], G["align.bn.bias"], act=K.ACT_RELU,
                             chan_scale=st.drop_mask, rows_per_sample=h1 * w1)
        dhs = _lin_bwd(dal_pre, sv["hs"], pk("align")[0], G["align.conv.weight"].view(pk("align")[0].shape), None, T)
        # hs = ReLU(s + BN(ho_pre)): g = gradient of the pre-activation = gradient of the residual s
        dho_pre, ds_res = _bn_bwd(dhs, sv["ho_pre"], sv["ms_o"], st.bn_out, sv["n_o"], T, G["hamburger.ham_out.bn.weight"],
                                  G["hamburger.ham_out.bn.bias"], act=K.ACT_RELU, residual=sv["s"])
        dnmf = _lin_bwd(dho_pre, sv["nmf"], pk("ham_out")[0], G["hamburger.ham_out.conv.weight"].view(pk("ham_out")[0].shape), None, T)
        hin = sv["hin"]
        dhin = _nmf_bwd(dnmf.view(B, h1 * w1, D), hin.view(B, h1 * w1, D), sv["nmf_saved"], T, side=getattr(st, "side", None)).view(M, D)
        dhin_pre = K.act_bwd(dhin, hin, K.ACT_RELU)
        ds = _lin_bwd(dhin_pre, sv["s"], pk("ham_in")[0], G["hamburger.ham_in.conv.weight"].view(pk("ham_in")[0].shape),
                      G["hamburger.ham_in.conv.bias"], T)
        # s feeds ham_in and the residual of ham_out: the two gradients are summed inside the BatchNorm backward
        ds_pre, _ = _bn_bwd(ds, sv["s_pre"], sv["ms_s"], st.bn_sq, sv["n_s"], T, G["squeeze.bn.weight"], G["squeeze.bn.bias"], act=K.ACT_RELU,
                            dy2=ds_res)
        dcat = _lin_bwd(ds_pre, sv["cat"], pk("squeeze")[0], G["squeeze.conv.weight"].view(pk("squeeze")[0].shape), None, T)
        do1 = torch.empty((M, C1), device=dev, dtype=F32)
        do2 = torch.empty((B * h2 * w2, C2), device=dev, dtype=F32)
        do3 = torch.empty((B * h3 * w3, C3), device=dev, dtype=F32)
        K.resize_bwd(dcat, 0, B, h1, w1, C1, h1, w1, do1)
        K.resize_bwd(dcat, C1, B, h2, w2, C2, h1, w1, do2)
        K.resize_bwd(dcat, C1 + C2, B, h3, w3, C3, h1, w1, do3)
        ctx.sv = None
        _WGRAD_STREAM = None
        if ws is not None:
            K.join(ws)
        ar.done(st.tag)
        return (do1, do2, do3, None, None) + tuple(G[n] for n in st.names)


# ============================================================================================ upsample + CE
class UpsampleCEFn(torch.autograd.Function):
    """builder.py:203 (F.interpolate x8, bilinear, align_corners=False) + :230 (masked-mean CE).
    Returns (loss, out_nchw); only the loss is differentiable."""

    @staticmethod
    def forward(ctx, small, label, meta):
        B, h, w, ncls, H, W, ignore, want_out = meta[:8]
        grad = meta[8] if len(meta) > 8 else True            # the caller's grad mode (inside forward it is always off)
        ctx.meta = meta[:8]
        ctx.one_pass = (label is not None and not want_out and grad and ctx.needs_input_grad[0] and K.upsample_ce_train_supported(h, w, H, W))
        if ctx.one_pass:          # training without the hi-res logits: loss and (unscaled) gradient from ONE launch, nothing hi-res kept
            loss, acc, dgrad = K.upsample_ce_train(small, B, h, w, ncls, H, W, label, ignore)
            ctx.save_for_backward(dgrad, acc)
            ctx.small_dtype = small.dtype
            out = torch.empty(0, device=small.device)
            ctx.mark_non_differentiable(out)
            return loss, out
        out, lse, acc, loss, _ = K.upsample_ce_fwd(small, B, h, w, ncls, H, W, label, ignore, want_out=want_out, want_loss=label is not None)
        if label is not None:
            ctx.save_for_backward(small, label, lse, acc)
        if out is None:
            out = torch.empty(0, device=small.device)
        if loss is None:
            loss = torch.zeros((), device=small.device)
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, dloss, _dout):
        B, h, w, ncls, H, W, ignore, _ = ctx.meta
        dl = dloss.contiguous().float()
        if ctx.one_pass:
            dgrad, acc = ctx.saved_tensors
            return K.ce_grad_finalize(dgrad, acc, dl, ctx.small_dtype), None, None
        small, label, lse, acc = ctx.saved_tensors
        ds = K.upsample_ce_bwd_fused(small, B, h, w, ncls, H, W, label, ignore, lse, acc, dl)
        return ds, None, None


class UpsampleFn(torch.autograd.Function):
    """Differentiable x8 upsample alone (eval / `decode` API): out NCHW fp32 from channels-last logits."""

    @staticmethod
    def forward(ctx, small, meta):
        B, h, w, ncls, H, W = meta
        out = K.upsample_ce_fwd(small, B, h, w, ncls, H, W, None, 255, want_out=True, want_loss=False)[0]
        ctx.meta, ctx.dtype = meta, small.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        B, h, w, ncls, H, W = ctx.meta
        # adjoint of the bilinear resize on an NCHW gradient: go through channels-last (layout plumbing only)
        g = dout.permute(0, 2, 3, 1).contiguous().view(B * H * W, ncls)
        pad = (8 - ncls % 8) % 8
        if pad:
            g = torch.nn.functional.pad(g, (0, pad))
        din = torch.empty((B * h * w, ncls + pad), device=dout.device, dtype=F32)
        K.resize_bwd(g, 0, B, h, w, ncls + pad, H, W, din)
        din = din[:, :ncls].contiguous()
        return (din if ctx.dtype == F32 else K.cast(din, ctx.dtype)), None
