"""CPU oracle for the DFormer RGB-D hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch fp32, functional restatement of the reference algorithm
(Originofamonia/DFormer): dual-stream encoder, LightHamHead (NMF2D "Hamburger")
decoder, x8 bilinear upsample and masked-mean cross-entropy.  It is driven by a
flat ``state_dict`` (the reference's own key layout) and nothing else, so the same
weights can be pushed through the reference, this oracle and the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``dformer_b200`` never does (it raises when its CUDA extension is missing).

Pinning: ``oracle/make_golden.py`` imports the UNMODIFIED reference modules from
/root/reference (through the mmcv/mmengine stand-ins in ``oracle/ref_shim``) in the
build container and writes golden input/output vectors to ``tests/golden``;
``tests/test_oracle_golden.py`` checks this restatement against them.  The reference
itself ships no tests or golden vectors for this path, so the
oracle is pinned by outputs of the reference run in the build container.

Reference lines followed (paths relative to the reference root):
  models/encoders/DFormer.py:21-45    LayerNorm (channels_last branch)
  models/encoders/DFormer.py:48-67    MLP
  models/encoders/DFormer.py:70-145   Attention
  models/encoders/DFormer.py:147-181  Block
  models/encoders/DFormer.py:184-305  DFormer (stems, downsample, stages, forward)
  models/encoders/DFormer.py:460-497  Tiny/Small/Base/Large hyper-parameters
  models/decoders/ham_head.py:46-55,60-100,109-145   NMF2D
  models/decoders/ham_head.py:148-180 Hamburger
  models/decoders/ham_head.py:184-240 LightHamHead
  models/decoders/decode_head.py:226-231  cls_seg
  models/builder.py:193-208,224-235   encode_decode / forward (loss)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# ATen-faithful mode.  The primitive restatements below (LayerNorm as mean / var arithmetic, GELU through erf, BatchNorm through
# mean / var, the 49-window pooling loop, explicit bilinear index math) are what pins the ALGORITHM; but under torch.autocast
# they round to bf16 after every primitive, which the reference never does: it dispatches the fused ATen ops `F.layer_norm`,
# `F.gelu`, `F.batch_norm`, `F.adaptive_avg_pool2d`, `F.interpolate`, `F.linear`, `F.cross_entropy`.  Inside
# `with aten_faithful():` every primitive below calls exactly the op the reference modules call, so that
# `torch.autocast(bf16)` around `forward` reproduces the reference's AMP rounding points (the bf16 yardstick of the tests and
# the B200-eager comparator of bench.py).  Both modes are pinned by the same golden vectors (tests/test_oracle_golden.py).
_ATEN = False


class aten_faithful:
    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        global _ATEN
        self.prev, _ATEN = _ATEN, self.on
        return self

    def __exit__(self, *exc):
        global _ATEN
        _ATEN = self.prev
        return False

# models/encoders/DFormer.py:460-497
VARIANTS = {
    "DFormer-Tiny": dict(dims=(32, 64, 128, 256), depths=(3, 3, 5, 2)),
    "DFormer-Small": dict(dims=(64, 128, 256, 512), depths=(2, 2, 4, 2)),
    "DFormer-Base": dict(dims=(64, 128, 256, 512), depths=(3, 3, 12, 2)),
    "DFormer-Large": dict(dims=(96, 192, 288, 576), depths=(3, 3, 12, 2)),
}
for _v in VARIANTS.values():
    _v.update(mlp_ratios=(8, 8, 4, 4), num_heads=(1, 2, 4, 8), windows=(0, 7, 7, 7))


# --------------------------------------------------------------------------- norms
def layer_norm_cl(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """DFormer.py:37-39 -- LayerNorm over the last (channel) axis of an NHWC tensor."""
    if _ATEN:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(dim=-1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def batch_norm_nchw(x: Tensor, P: Params, prefix: str, training: bool, eps: float,
                    momentum: float = 0.1, new_stats: Optional[Params] = None) -> Tensor:
    """nn.BatchNorm2d: batch statistics (biased var) in training, running stats in eval.

    When ``new_stats`` is a dict the updated running statistics (unbiased var,
    momentum 0.1) are written into it under the reference buffer names."""
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    if _ATEN:
        rm, rv = P[prefix + ".running_mean"].detach().clone(), P[prefix + ".running_var"].detach().clone()
        y = F.batch_norm(x, rm, rv, w, b, training, momentum, eps)
        if training and new_stats is not None:
            new_stats[prefix + ".running_mean"], new_stats[prefix + ".running_var"] = rm, rv
            new_stats[prefix + ".num_batches_tracked"] = P[prefix + ".num_batches_tracked"] + 1
        return y
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.numel() // x.shape[1]
            unbiased = var * (n / max(n - 1, 1))
            new_stats[prefix + ".running_mean"] = (1 - momentum) * P[prefix + ".running_mean"] + momentum * mean
            new_stats[prefix + ".running_var"] = (1 - momentum) * P[prefix + ".running_var"] + momentum * unbiased
            new_stats[prefix + ".num_batches_tracked"] = P[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form."""
    if _ATEN:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x: Tensor, P: Params, prefix: str) -> Tensor:
    if _ATEN:
        return F.linear(x, P[prefix + ".weight"], P[prefix + ".bias"])
    return x @ P[prefix + ".weight"].t() + P[prefix + ".bias"]


def dwconv_cl(x: Tensor, P: Params, prefix: str, k: int) -> Tensor:
    """Depthwise kxk 'same' conv with bias applied to an NHWC tensor (DFormer.py:54,80-81)."""
    c = x.shape[-1]
    y = F.conv2d(x.permute(0, 3, 1, 2), P[prefix + ".weight"], P[prefix + ".bias"], padding=k // 2, groups=c)
    return y.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------- encoder blocks
def mlp(P: Params, prefix: str, x: Tensor) -> Tensor:
    """DFormer.py:58-67 -- LN -> fc1 -> dw3x3 + identity -> GELU -> fc2 (NHWC)."""
    h = layer_norm_cl(x, P[prefix + ".norm.weight"], P[prefix + ".norm.bias"])
    h = linear(h, P, prefix + ".fc1")
    h = dwconv_cl(h, P, prefix + ".pos", 3) + h
    return linear(gelu(h), P, prefix + ".fc2")


def adaptive_avg_pool_7(x_nchw: Tensor) -> Tensor:
    """nn.AdaptiveAvgPool2d((7,7)): window i = [floor(i*In/7), ceil((i+1)*In/7))."""
    if _ATEN:
        return F.adaptive_avg_pool2d(x_nchw, (7, 7))
    B, C, H, W = x_nchw.shape
    out = x_nchw.new_zeros(B, C, 7, 7)
    for i in range(7):
        h0, h1 = (i * H) // 7, -((-(i + 1) * H) // 7)
        for j in range(7):
            w0, w1 = (j * W) // 7, -((-(j + 1) * W) // 7)
            out[:, :, i, j] = x_nchw[:, :, h0:h1, w0:w1].mean(dim=(2, 3))
    return out


def _bilinear_axis(n_in: int, n_out: int) -> Tuple[Tensor, Tensor, Tensor]:
    """align_corners=False source index / weights along one axis (ATen upsample_bilinear2d)."""
    scale = n_in / n_out
    dst = torch.arange(n_out, dtype=torch.float32)
    src = ((dst + 0.5) * scale - 0.5).clamp(min=0.0)
    i0 = src.floor().to(torch.long).clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    w1 = src - i0.to(torch.float32)
    return i0, i1, w1


def bilinear_resize_nchw(x: Tensor, size: Sequence[int]) -> Tensor:
    """F.interpolate(mode='bilinear', align_corners=False) restated with explicit index math."""
    if _ATEN:
        return F.interpolate(x, size=(int(size[0]), int(size[1])), mode="bilinear", align_corners=False)
    H, W = x.shape[-2:]
    Ho, Wo = int(size[0]), int(size[1])
    if (H, W) == (Ho, Wo):
        return x
    y0, y1, wy = (t.to(x.device) for t in _bilinear_axis(H, Ho))
    x0, x1, wx = (t.to(x.device) for t in _bilinear_axis(W, Wo))
    wy = wy.to(x.dtype)[:, None]
    wx = wx.to(x.dtype)[None, :]
    top = x[..., y0, :]
    bot = x[..., y1, :]
    tl, tr = top[..., x0], top[..., x1]
    bl, br = bot[..., x0], bot[..., x1]
    return (1 - wy) * ((1 - wx) * tl + wx * tr) + wy * ((1 - wx) * bl + wx * br)


def attention(P: Params, prefix: str, x: Tensor, x_e: Tensor, num_head: int, window: int,
              drop_depth: bool) -> Tuple[Tensor, Optional[Tensor]]:
    """DFormer.py:102-145.  x: (B,H,W,C), x_e: (B,H,W,C/2); queries are the 49 pooled tokens."""
    B, H, W, C = x.shape
    xn = layer_norm_cl(x, P[prefix + ".norm.weight"], P[prefix + ".norm.bias"])
    en = layer_norm_cl(x_e, P[prefix + ".norm_e.weight"], P[prefix + ".norm_e.bias"])
    q = linear(xn, P, prefix + ".q")
    cut = linear(xn, P, prefix + ".q_cut")
    l = gelu(linear(xn, P, prefix + ".l"))
    a = linear(dwconv_cl(l, P, prefix + ".conv", 7), P, prefix + ".a")
    if window != 0:
        d = C // num_head // 2
        kv = linear(l, P, prefix + ".kv").reshape(B, H * W, 2, num_head, d).permute(2, 0, 3, 1, 4)
        k, v = kv[0], kv[1]                                       # (B, heads, HW, d)
        sc = torch.cat([xn, en], dim=3).permute(0, 3, 1, 2)       # (B, 3C/2, H, W)
        m = adaptive_avg_pool_7(sc).permute(0, 2, 3, 1)           # (B, 7, 7, 3C/2)
        m = linear(m, P, prefix + ".short_cut_linear")            # (B, 7, 7, C/2)
        m = m.reshape(B, 49, num_head, d).permute(0, 2, 1, 3)     # (B, heads, 49, d)
        s = (m * d ** -0.5) @ k.transpose(-2, -1)                 # (B, heads, 49, HW)
        s = s.softmax(dim=-1)
        o = (s @ v).reshape(B, num_head, 7, 7, d).permute(0, 1, 4, 2, 3).reshape(B, C // 2, 7, 7)
        o = bilinear_resize_nchw(o, (H, W)).permute(0, 2, 3, 1)   # (B, H, W, C/2)
    e = linear(dwconv_cl(linear(en, P, prefix + ".e_fore"), P, prefix + ".e_conv", 7), P, prefix + ".e_back")
    cut = cut * e
    g = q * a
    y = torch.cat([g, o, cut], dim=3) if window != 0 else torch.cat([g, cut], dim=3)
    y_e = None if drop_depth else linear(y, P, prefix + ".proj_e")
    return linear(y, P, prefix + ".proj"), y_e


def block(P: Params, prefix: str, x: Tensor, x_e: Tensor, num_head: int, window: int,
          drop_depth: bool) -> Tuple[Tensor, Tensor]:
    """DFormer.py:168-181 with DropPath = identity (eval, or drop_path_rate = 0)."""
    ax, ae = attention(P, prefix + ".attn", x, x_e, num_head, window, drop_depth)
    x = x + P[prefix + ".layer_scale_1"] * ax
    x = x + P[prefix + ".layer_scale_2"] * mlp(P, prefix + ".mlp", x)
    if not drop_depth:
        x_e = x_e + P[prefix + ".layer_scale_1_e"] * ae
        x_e = x_e + P[prefix + ".layer_scale_2_e"] * mlp(P, prefix + ".mlp_e2", x_e)
    return x, x_e


def conv3x3_s2(x: Tensor, P: Params, prefix: str) -> Tensor:
    return F.conv2d(x, P[prefix + ".weight"], P[prefix + ".bias"], stride=2, padding=1)


def encoder(P: Params, rgb: Tensor, modal_x: Tensor, dims: Sequence[int], depths: Sequence[int],
            num_heads: Sequence[int] = (1, 2, 4, 8), windows: Sequence[int] = (0, 7, 7, 7),
            training: bool = False, prefix: str = "encoder_backbone", bn_eps: float = 1e-5,
            new_stats: Optional[Params] = None) -> List[Tensor]:
    """DFormer.py:278-305.  Returns the four NCHW stage outputs."""
    x = rgb
    x_e = modal_x[:, 0:1]                                         # :286 -- only channel 0 is used
    outs = []
    for i in range(4):
        for name, t in (("downsample_layers", "x"), ("downsample_layers_e", "e")):
            p = f"{prefix}.{name}.{i}"
            cur = x if t == "x" else x_e
            if i == 0:                                            # :194-211 conv-BN-GELU-conv-BN
                cur = conv3x3_s2(cur, P, p + ".0")
                cur = gelu(batch_norm_nchw(cur, P, p + ".1", training, bn_eps, new_stats=new_stats))
                cur = conv3x3_s2(cur, P, p + ".3")
                cur = batch_norm_nchw(cur, P, p + ".4", training, bn_eps, new_stats=new_stats)
            else:                                                 # :216-228 BN-conv
                cur = batch_norm_nchw(cur, P, p + ".0", training, bn_eps, new_stats=new_stats)
                cur = conv3x3_s2(cur, P, p + ".1")
            if t == "x":
                x = cur
            else:
                x_e = cur
        x = x.permute(0, 2, 3, 1)
        x_e = x_e.permute(0, 2, 3, 1)
        for j in range(depths[i]):
            drop_depth = (i == 3) and (j == depths[i] - 1)        # :243
            x, x_e = block(P, f"{prefix}.stages.{i}.{j}", x, x_e, num_heads[i], windows[i], drop_depth)
        x = x.permute(0, 3, 1, 2)
        x_e = x_e.permute(0, 3, 1, 2)
        outs.append(x)
    return outs


# --------------------------------------------------------------------------- decoder
def draw_bases(batch: int, D: int = 512, R: int = 64) -> Tensor:
    """ham_head.py:111 -- the reference draws NMF bases with a CPU torch.rand every forward."""
    return torch.rand((batch, D, R))


def nmf2d(x: Tensor, bases_raw: Tensor, steps: int) -> Tensor:
    """ham_head.py:60-100,109-145.  x: (B, D, N) >= 0, bases_raw: (B, D, R) uniform[0,1)."""
    bases = F.normalize(bases_raw.to(x.device, x.dtype), dim=1)                # :115
    coef = torch.bmm(x.transpose(1, 2), bases).softmax(dim=-1)                 # :48-49, inv_t = 1 (:107)

    def coef_update(coef, bases):                                              # :122-126 / :139-143
        num = torch.bmm(x.transpose(1, 2), bases)
        den = coef.bmm(bases.transpose(1, 2).bmm(bases))
        return coef * num / (den + 1e-6)

    for _ in range(steps):
        coef = coef_update(coef, bases)
        num = torch.bmm(x, coef)                                               # :129-133
        den = bases.bmm(coef.transpose(1, 2).bmm(coef))
        bases = bases * num / (den + 1e-6)
    coef = coef_update(coef, bases)                                            # compute_coef :86
    return torch.bmm(bases, coef.transpose(1, 2))                              # :89


def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.conv2d(x, w, b)


def ham_head(P: Params, outs: Sequence[Tensor], bases_raw: Tensor, training: bool = False,
             prefix: str = "decode_head", bn_eps: float = 1e-3, train_steps: int = 6,
             eval_steps: int = 7, new_stats: Optional[Params] = None,
             dropout_mask: Optional[Tensor] = None) -> Tensor:
    """ham_head.py:222-240 + 173-180 + decode_head.py:226-231 (Dropout2d = identity unless a
    (B, C) keep-mask already divided by keep-prob is injected)."""
    levels = [outs[1], outs[2], outs[3]]                          # in_index [1,2,3]
    size = levels[0].shape[2:]
    x = torch.cat([bilinear_resize_nchw(l, size) for l in levels], dim=1)
    x = conv1x1(x, P[prefix + ".squeeze.conv.weight"])
    x = F.relu(batch_norm_nchw(x, P, prefix + ".squeeze.bn", training, bn_eps, new_stats=new_stats))
    e = F.relu(conv1x1(x, P[prefix + ".hamburger.ham_in.conv.weight"], P[prefix + ".hamburger.ham_in.conv.bias"]))
    B, C, H, W = e.shape
    e = nmf2d(e.reshape(B, C, H * W), bases_raw, train_steps if training else eval_steps).reshape(B, C, H, W)
    e = conv1x1(e, P[prefix + ".hamburger.ham_out.conv.weight"])
    e = batch_norm_nchw(e, P, prefix + ".hamburger.ham_out.bn", training, bn_eps, new_stats=new_stats)
    x = F.relu(x + e)
    x = conv1x1(x, P[prefix + ".align.conv.weight"])
    x = F.relu(batch_norm_nchw(x, P, prefix + ".align.bn", training, bn_eps, new_stats=new_stats))
    if dropout_mask is not None:
        x = x * dropout_mask[:, :, None, None]
    return conv1x1(x, P[prefix + ".conv_seg.weight"], P[prefix + ".conv_seg.bias"])


def masked_ce(logits: Tensor, label: Tensor, background: int = 255) -> Tensor:
    """builder.py:230 -- CE(reduction='none', ignore_index=255) then mean over label != background."""
    label = label.long()
    if _ATEN:
        return F.cross_entropy(logits, label, reduction="none", ignore_index=255)[label != background].mean()
    lse = torch.logsumexp(logits, dim=1)
    valid = label != background
    safe = label.clamp(max=logits.shape[1] - 1)
    picked = logits.gather(1, safe[:, None]).squeeze(1)
    return (lse - picked)[valid].mean()


def forward(P: Params, rgb: Tensor, modal_x: Tensor, bases_raw: Tensor, dims: Sequence[int],
            depths: Sequence[int], label: Optional[Tensor] = None, training: bool = False,
            num_heads: Sequence[int] = (1, 2, 4, 8), windows: Sequence[int] = (0, 7, 7, 7),
            head_bn_eps: float = 1e-3, new_stats: Optional[Params] = None, return_all: bool = False):
    """builder.py:193-208,224-235 with the fork's `(outs, None)` tuple unwrapped."""
    outs = encoder(P, rgb, modal_x, dims, depths, num_heads, windows, training, new_stats=new_stats)
    small = ham_head(P, outs, bases_raw, training, bn_eps=head_bn_eps, new_stats=new_stats)
    out = bilinear_resize_nchw(small, rgb.shape[-2:])
    loss = masked_ce(out, label) if label is not None else None
    if return_all:
        return dict(outs=outs, small=small, out=out, loss=loss)
    return (loss, out) if label is not None else out
