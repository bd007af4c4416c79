"""Mirror of the reference `models/encoders/DFormer.py` (LayerNorm :21-45, MLP :48-67, Attention :70-145,
Block :147-181, DFormer :184-305, constructors :460-497).

The classes below own nn.Parameters with the reference's names and shapes (so state_dicts are
interchangeable, strict=True) and are constructed in the reference's order (so the same torch seed
yields the same random init), but they contain no arithmetic: `DFormer.forward` hands the parameters
to the hand-scheduled autograd.Functions of `dformer_b200.functions`, which launch the sm_100a kernels."""
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn as nn

from ... import functions as Fn
from ...runtime import GradArena, ParamLayout, ParamPacker, bump_weights_epoch, flatten_parameters, is_flat, require_cuda, resolve_dtype


def _norm_layer(norm_cfg, channels):
    cfg = dict(norm_cfg or dict(type="BN"))
    kind = cfg.get("type", "BN")
    cls = {"BN": nn.BatchNorm2d, "SyncBN": nn.SyncBatchNorm}[kind]
    layer = cls(channels, eps=cfg.get("eps", 1e-5))
    for p in layer.parameters():
        p.requires_grad = cfg.get("requires_grad", True)
    return layer


class LayerNorm(nn.Module):
    """channels_last LayerNorm parameter holder (DFormer.py:21-45); eps 1e-6."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format != "channels_last":
            raise NotImplementedError("only the channels_last branch is on the DFormer hot path")
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)


class MLP(nn.Module):
    def __init__(self, dim, mlp_ratio=4, norm_cfg=None):
        super().__init__()
        self.norm = LayerNorm(dim, eps=1e-6)
        self.fc1 = nn.Linear(dim, dim * mlp_ratio)
        self.pos = nn.Conv2d(dim * mlp_ratio, dim * mlp_ratio, 3, padding=1, groups=dim * mlp_ratio)
        self.fc2 = nn.Linear(dim * mlp_ratio, dim)
        self.act = nn.GELU()


class Attention(nn.Module):
    def __init__(self, dim, num_head=8, window=7, norm_cfg=None, drop_depth=False):
        super().__init__()
        self.num_head, self.window, self.drop_depth = num_head, window, drop_depth
        half = dim // 2
        self.q = nn.Linear(dim, dim)
        self.q_cut = nn.Linear(dim, half)
        self.a = nn.Linear(dim, dim)
        self.l = nn.Linear(dim, dim)
        self.conv = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.e_conv = nn.Conv2d(half, half, 7, padding=3, groups=half)
        self.e_fore = nn.Linear(half, half)
        self.e_back = nn.Linear(half, half)
        # the reference first builds the window-less projections and then replaces them (DFormer.py:86-95);
        # the discarded layers still consume RNG, so they are drawn here too to keep seeds interchangeable
        self.proj = nn.Linear(half * 3, dim)
        if not drop_depth:
            self.proj_e = nn.Linear(half * 3, half)
        if window != 0:
            self.short_cut_linear = nn.Linear(half * 3, half)
            self.kv = nn.Linear(dim, dim)
            self.pool = nn.AdaptiveAvgPool2d(output_size=(7, 7))
            self.proj = nn.Linear(dim * 2, dim)
            if not drop_depth:
                self.proj_e = nn.Linear(dim * 2, half)
        self.act = nn.GELU()
        self.norm = LayerNorm(dim, eps=1e-6)
        self.norm_e = LayerNorm(half, eps=1e-6)


class DropPathCfg(nn.Module):
    """Holder of the stochastic-depth probability (mmcv DropPath stand-in; no parameters)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)


class Block(nn.Module):
    def __init__(self, index, dim, num_head, norm_cfg=None, mlp_ratio=4., block_index=0, last_block_index=50, window=7,
                 dropout_layer=None, drop_depth=False):
        super().__init__()
        self.index, self.dim, self.num_head = index, dim, num_head
        if block_index > last_block_index:
            window = 0
        self.window = window
        self.attn = Attention(dim, num_head, window=window, norm_cfg=norm_cfg, drop_depth=drop_depth)
        self.mlp = MLP(dim, mlp_ratio, norm_cfg=norm_cfg)
        self.dropout_layer = DropPathCfg(dropout_layer.get("drop_prob", 0.0)) if dropout_layer else nn.Identity()
        init = 1e-6
        self.layer_scale_1 = nn.Parameter(init * torch.ones(dim), requires_grad=True)
        self.layer_scale_2 = nn.Parameter(init * torch.ones(dim), requires_grad=True)
        if not drop_depth:
            self.layer_scale_1_e = nn.Parameter(init * torch.ones(dim // 2), requires_grad=True)
            self.layer_scale_2_e = nn.Parameter(init * torch.ones(dim // 2), requires_grad=True)
            self.mlp_e2 = MLP(dim // 2, mlp_ratio)
        self.drop_depth = drop_depth

    # parameter order inside the gradient arena: members of one fused GEMM are adjacent
    def param_names(self):
        a = ["attn.l.weight", "attn.q.weight", "attn.q_cut.weight", "attn.l.bias", "attn.q.bias", "attn.q_cut.bias",
             "attn.a.weight", "attn.a.bias"]
        if self.window != 0:
            a += ["attn.kv.weight", "attn.kv.bias", "attn.short_cut_linear.weight", "attn.short_cut_linear.bias"]
        a += ["attn.e_fore.weight", "attn.e_fore.bias", "attn.e_back.weight", "attn.e_back.bias"]
        if self.drop_depth:
            a += ["attn.proj.weight", "attn.proj.bias"]
        else:
            a += ["attn.proj.weight", "attn.proj_e.weight", "attn.proj.bias", "attn.proj_e.bias"]
        a += ["attn.conv.weight", "attn.conv.bias", "attn.e_conv.weight", "attn.e_conv.bias", "attn.norm.weight", "attn.norm.bias",
              "attn.norm_e.weight", "attn.norm_e.bias", "layer_scale_1", "layer_scale_2"]
        mlp = ["norm.weight", "norm.bias", "fc1.weight", "fc1.bias", "pos.weight", "pos.bias", "fc2.weight", "fc2.bias"]
        a += ["mlp." + n for n in mlp]
        if not self.drop_depth:
            a += ["layer_scale_1_e", "layer_scale_2_e"] + ["mlp_e2." + n for n in mlp]
        return a

    def gemm_groups(self):
        g = [("attn.qcl", ["attn.l", "attn.q", "attn.q_cut"]), ("attn.a", ["attn.a"])]
        if self.window != 0:
            g += [("attn.kv", ["attn.kv"]), ("attn.short_cut_linear", ["attn.short_cut_linear"])]
        g += [("attn.e_fore", ["attn.e_fore"]), ("attn.e_back", ["attn.e_back"]),
              ("attn.pp", ["attn.proj"] if self.drop_depth else ["attn.proj", "attn.proj_e"]),
              ("mlp.fc1", ["mlp.fc1"]), ("mlp.fc2", ["mlp.fc2"])]
        if not self.drop_depth:
            g += [("mlp_e2.fc1", ["mlp_e2.fc1"]), ("mlp_e2.fc2", ["mlp_e2.fc2"])]
        return g


class DFormer(nn.Module):
    """Dual-stream RGB-D encoder.  forward(x, x_e) -> (outs, None) exactly like the fork (DFormer.py:305):
    `outs` are four (B, C_i, H_i, W_i) fp32 tensors (channels-last memory, NCHW shape)."""

    def __init__(self, in_channels=4, depths=(2, 2, 8, 2), dims=(32, 64, 128, 256), out_indices=(0, 1, 2, 3), windows=(7, 7, 7, 7),
                 norm_cfg=dict(type="SyncBN", requires_grad=True), mlp_ratios=(8, 8, 4, 4), num_heads=(2, 4, 10, 16),
                 last_block=(50, 50, 50, 50), drop_path_rate=0.1, init_cfg=None, precision=None):
        super().__init__()
        self.depths, self.dims, self.num_heads, self.windows = tuple(depths), tuple(dims), tuple(num_heads), tuple(windows)
        self.init_cfg, self.out_indices = init_cfg, out_indices
        self.precision = precision
        self.sync_bn = dict(norm_cfg or {}).get("type") == "SyncBN"
        d0 = dims[0]
        self.downsample_layers = nn.ModuleList()
        stem = nn.Sequential(nn.Conv2d(3, d0 // 2, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(d0 // 2), nn.GELU(),
                             nn.Conv2d(d0 // 2, d0, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(d0))
        self.stem_e_fc1 = nn.Linear(360, 640)      # present in the fork's state_dict, never used on the hot path (DFormer.py:202-203)
        self.stem_e_fc2 = nn.Linear(1, 480)
        self.downsample_layers_e = nn.ModuleList()
        stem_e = nn.Sequential(nn.Conv2d(1, d0 // 4, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(d0 // 4), nn.GELU(),
                               nn.Conv2d(d0 // 4, d0 // 2, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(d0 // 2))
        self.downsample_layers.append(stem)
        self.downsample_layers_e.append(stem_e)
        for i in range(len(dims) - 1):
            self.downsample_layers.append(nn.Sequential(_norm_layer(norm_cfg, dims[i]),
                                                        nn.Conv2d(dims[i], dims[i + 1], kernel_size=3, stride=2, padding=1)))
            self.downsample_layers_e.append(nn.Sequential(_norm_layer(norm_cfg, dims[i] // 2),
                                                          nn.Conv2d(dims[i] // 2, dims[i + 1] // 2, kernel_size=3, stride=2, padding=1)))
        self.stages = nn.ModuleList()
        dp_rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for i in range(len(dims)):
            self.stages.append(nn.Sequential(*[
                Block(index=cur + j, dim=dims[i], window=windows[i], dropout_layer=dict(type="DropPath", drop_prob=dp_rates[cur + j]),
                      num_head=num_heads[i], norm_cfg=norm_cfg, block_index=depths[i] - j, last_block_index=last_block[i],
                      mlp_ratio=mlp_ratios[i], drop_depth=((i == 3) and (j == depths[i] - 1))) for j in range(depths[i])]))
            cur += depths[i]
        self._plan = None
        self._side_stream = None
        self._last_arena = None
        self.grad_hook = None           # set by the data-parallel engine: callable(arena, lo, hi)

    # ------------------------------------------------------------------ plan: arena layout + GEMM packing
    def _build_plan(self):
        named = dict(self.named_parameters())
        layout, packer = ParamLayout(), ParamPacker(named)
        for sfx in ("", "_e"):
            p = f"downsample_layers{sfx}.0."
            layout.begin_mark(p)
            for n in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "4.weight", "4.bias"):
                layout.add(p + n, named[p + n])
            layout.end_mark(p)
            packer.add(p + "c1", [p + "0.weight"], [p + "0.bias"], kind=1)
            packer.add(p + "c2", [p + "3.weight"], [p + "3.bias"], kind=1)
        for i in range(4):
            if i > 0:
                for sfx in ("", "_e"):
                    p = f"downsample_layers{sfx}.{i}."
                    layout.begin_mark(p)
                    for n in ("0.weight", "0.bias", "1.weight", "1.bias"):
                        layout.add(p + n, named[p + n])
                    layout.end_mark(p)
                    packer.add(p + "c", [p + "1.weight"], [p + "1.bias"], kind=1)
            for j, blk in enumerate(self.stages[i]):
                p = f"stages.{i}.{j}."
                layout.begin_mark(p)
                for n in blk.param_names():
                    layout.add(p + n, named[p + n])
                layout.end_mark(p)
                for gname, members in blk.gemm_groups():
                    packer.add(p + gname, [p + m + ".weight" for m in members], [p + m + ".bias" for m in members])
        self._plan = SimpleNamespace(layout=layout, packer=packer, named=named, flat=None)
        return self._plan

    def flat_parameters(self):
        """Flat fp32 buffer holding every hot-path parameter (arena layout); created on first use on the GPU."""
        plan = self._plan or self._build_plan()
        if not is_flat(plan.layout, plan.flat):
            plan.flat = flatten_parameters(self, plan.layout)
        return plan.flat

    def _apply(self, fn, *a, **k):
        self._plan = None               # parameters may be re-allocated (.cuda(), .to(), ...)
        return super()._apply(fn, *a, **k)

    def init_weights(self, pretrained):
        """DFormer.py:254-276: load a checkpoint (`state_dict_ema` / `state_dict`), strip `backbone.` / `module.`
        prefixes, load non-strictly and FREEZE every parameter that was found in the checkpoint."""
        ck = torch.load(pretrained, map_location="cpu", weights_only=False)
        ck = ck["state_dict_ema"] if "state_dict_ema" in ck else ck["state_dict"]
        sd = OrderedDict((k[9:] if k.startswith("backbone.") else k, v) for k, v in ck.items())
        if sd and next(iter(sd)).startswith("module."):
            sd = OrderedDict((k[7:], v) for k, v in sd.items())
        self.load_state_dict(sd, strict=False)
        for name, p in self.named_parameters():
            if any(name == k or name.startswith(k + ".") for k in sd):
                p.requires_grad = False

    # ------------------------------------------------------------------ forward
    def forward(self, x, x_e):
        if x_e is None:
            x_e = x
        if x.dim() == 3:
            x = x.unsqueeze(0)
        if x_e.dim() == 3:
            x_e = x_e.unsqueeze(2)
        x_e = x_e[:, 0:1]                                  # DFormer.py:286 -- only channel 0 of the modality tensor
        require_cuda(x, x_e)
        T = resolve_dtype(self.precision)
        x, x_e = x.float(), x_e.float()
        plan = self._plan or self._build_plan()
        dev = x.device
        if self._side_stream is None or self._side_stream[0].device != dev:
            self._side_stream = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
        side, side2, wstream = self._side_stream
        # compute-dtype copies of the GEMM weights: on a side stream, under the stems' im2col gathers (which need no weights)
        grad_mode = torch.is_grad_enabled()               # recorded for the Functions: inside Function.forward it is always off
        # inference (eval mode, no gradients, folding on): the packed copies and the BatchNorm folds applied to them are reused
        frozen = (not self.training) and (not grad_mode) and Fn._FOLD_BN
        bn_modes = tuple(m.training for m in self.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm))
        if not frozen:
            bump_weights_epoch()
        Fn.K.fork(side2)
        with torch.cuda.stream(side2):
            packed = plan.packer.pack(dev, T, frozen=frozen, token=bn_modes)
            ev_pack = Fn.K.signal(side2)
        fold_cache = plan.packer.fold_cache(dev, T) if frozen else None
        arena = GradArena(plan.layout, dev, self.grad_hook)
        self._last_arena = arena
        named = plan.named
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        training = self.training
        sync = self.sync_bn and training
        n_blocks = sum(self.depths)
        rates = [blk.dropout_layer.drop_prob if isinstance(blk.dropout_layer, DropPathCfg) else 0.0 for st in self.stages for blk in st]
        dp = None
        if training and any(r > 0 for r in rates):
            keep = getattr(plan, "dp_keep", None)
            if keep is None or keep.device != dev:     # cached on the device: no H2D copy inside a captured step
                keep = plan.dp_keep = (1.0 - torch.tensor(rates, dtype=torch.float32).view(-1, 1, 1)).to(dev)
            dp = torch.floor(keep + torch.rand(n_blocks, 4, B, device=dev)) / keep      # DropPath: mask / keep_prob per sample

        def stem(inp, sfx, cin):
            p = f"downsample_layers{sfx}.0."
            seq = self.downsample_layers[0] if sfx == "" else self.downsample_layers_e[0]
            st = SimpleNamespace(grad=grad_mode, dtype=T, cin=cin, packed=packed, ev_pack=ev_pack, fold_cache=fold_cache, g1=p + "c1", g2=p + "c2", arena=arena, prefix=p, tag=p,
                                 wstream=wstream,
                                 bn1=Fn.BNState(seq[1], p + "1", training, False), bn2=Fn.BNState(seq[4], p + "4", training, False))
            names = ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "4.weight", "4.bias")
            return Fn.StemFn.apply(inp, st, *[named[p + n] for n in names])

        def down(xx, sfx, i, h, w):
            p = f"downsample_layers{sfx}.{i}."
            seq = self.downsample_layers[i] if sfx == "" else self.downsample_layers_e[i]
            st = SimpleNamespace(grad=grad_mode, dtype=T, packed=packed, g=p + "c", arena=arena, prefix=p, tag=p, B=B, H=h, W=w,
                                 bn=Fn.BNState(seq[0], p + "0", training, True if sync else False))
            return Fn.DownsampleFn.apply(xx, st, *[named[p + n] for n in ("0.weight", "0.bias", "1.weight", "1.bias")])

        outs = []
        h, w = H, W
        bi = 0
        for i in range(4):
            # the two modality streams are independent here: the depth stem / downsample runs on the side stream
            main = torch.cuda.current_stream()
            Fn.K.fork(side)
            with torch.cuda.stream(side):
                x_e = stem(x_e, "_e", 1) if i == 0 else down(x_e, "_e", i, h, w)
            x_e.record_stream(main)
            x = stem(x, "", 3) if i == 0 else down(x, "", i, h, w)
            if i == 0:
                h, w = ((h + 1) // 2 + 1) // 2, ((w + 1) // 2 + 1) // 2
            else:
                h, w = (h + 1) // 2, (w + 1) // 2
            Fn.K.join(side)
            # MLP residuals left pending by a Block for the next Block's first LayerNorm kernel (functions._mlp_fwd / _carried_layernorm);
            # the last Block of a stage materialises its outputs (they are the stage output and the downsample input)
            handover = {}
            n_blk = len(self.stages[i])
            for j, blk in enumerate(self.stages[i]):
                p = f"stages.{i}.{j}."
                names = blk.param_names()
                st = SimpleNamespace(grad=grad_mode, dtype=T, packed=packed, key=p, arena=arena, prefix=p, tag=p, names=names, B=B, H=h, W=w,
                                     C=self.dims[i], heads=blk.num_head, window=blk.window, drop_depth=blk.drop_depth, side=side, side2=side2, wstream=wstream,
                                     dp=(tuple(dp[bi, k] if rates[bi] > 0 else None for k in range(4)) if dp is not None else (None,) * 4),
                                     pending=handover, carry=handover if j + 1 < n_blk else None)
                x, x_e = Fn.BlockFn.apply(x, x_e, st, *[named[p + n] for n in names])
                bi += 1
            assert not handover, "a Block left a residual pending at the end of its stage"
            outs.append(x.view(B, h, w, self.dims[i]).permute(0, 3, 1, 2))
        return outs, None


def DFormer_Tiny(pretrained=False, **kwargs):
    assert not pretrained, "use DFormer.init_weights(path); the fork's `load_model_weights` is undefined (DFormer.py:464)"
    return DFormer(dims=[32, 64, 128, 256], mlp_ratios=[8, 8, 4, 4], depths=[3, 3, 5, 2], num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7], **kwargs)


def DFormer_Small(pretrained=False, **kwargs):
    assert not pretrained
    return DFormer(dims=[64, 128, 256, 512], mlp_ratios=[8, 8, 4, 4], depths=[2, 2, 4, 2], num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7], **kwargs)


def DFormer_Base(pretrained=False, drop_path_rate=0.1, **kwargs):
    assert not pretrained
    return DFormer(dims=[64, 128, 256, 512], mlp_ratios=[8, 8, 4, 4], depths=[3, 3, 12, 2], num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7],
                   drop_path_rate=drop_path_rate, **kwargs)


def DFormer_Large(pretrained=False, drop_path_rate=0.1, **kwargs):
    assert not pretrained
    return DFormer(dims=[96, 192, 288, 576], mlp_ratios=[8, 8, 4, 4], depths=[3, 3, 12, 2], num_heads=[1, 2, 4, 8], windows=[0, 7, 7, 7],
                   drop_path_rate=drop_path_rate, **kwargs)
