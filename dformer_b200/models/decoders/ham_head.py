"""Mirror of `models/decoders/ham_head.py`: NMF2D (:103-145, base :11-100), Hamburger (:148-180),
LightHamHead (:184-240).  Parameter names / shapes / construction order follow the reference; the
arithmetic is `dformer_b200.functions.HeadFn` (resize+concat, 1x1 convs as GEMMs, BatchNorm, the NMF
multiplicative-update loop with full back-propagation, Dropout2d mask, classifier)."""
from types import SimpleNamespace

import torch
import torch.nn as nn

from ... import functions as Fn
from ...runtime import GradArena, ParamLayout, ParamPacker, bump_weights_epoch, flatten_parameters, is_flat, require_cuda, resolve_dtype
from .decode_head import BaseDecodeHead


def _norm_layer(norm_cfg, channels):
    cfg = dict(norm_cfg)
    cls = {"BN": nn.BatchNorm2d, "SyncBN": nn.SyncBatchNorm}[cfg.get("type", "BN")]
    layer = cls(channels, eps=cfg.get("eps", 1e-5))
    for p in layer.parameters():
        p.requires_grad = cfg.get("requires_grad", True)
    return layer


class ConvModule(nn.Module):
    """1x1 conv (+ BN named `bn`) (+ ReLU) parameter holder with mmcv's naming and default init
    (kaiming-normal fan_out for the conv, bias only when there is no norm)."""

    def __init__(self, in_channels, out_channels, kernel_size, conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU")):
        super().__init__()
        assert kernel_size == 1 and conv_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, 1, bias=norm_cfg is None)
        self.with_norm, self.with_activation = norm_cfg is not None, act_cfg is not None
        if self.with_norm:
            self.bn = _norm_layer(norm_cfg, out_channels)
        if self.with_activation:
            self.activate = nn.ReLU(inplace=True)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)


class NMF2D(nn.Module):
    """Hyper-parameters of the matrix decomposition (ham_head.py:15-27,107): S=1, D=512, R=64,
    6 train / 7 eval steps, inv_t forced to 1, random bases drawn on the CPU every forward (:111)."""

    def __init__(self, args=None, **kwargs):
        super().__init__()
        args = args if args is not None else {}
        self.args = args
        self.spatial = args.setdefault("SPATIAL", True)
        self.S = args.setdefault("MD_S", 1)
        self.D = args.setdefault("MD_D", 512)
        self.R = args.setdefault("MD_R", 64)
        self.train_steps = args.setdefault("TRAIN_STEPS", 6)
        self.eval_steps = args.setdefault("EVAL_STEPS", 7)
        args.setdefault("INV_T", 100)
        self.eta = args.setdefault("ETA", 0.9)
        self.rand_init = args.setdefault("RAND_INIT", True)
        self.inv_t = 1
        assert self.spatial and self.S == 1 and self.rand_init, "only the configuration used by LightHamHead is implemented"

    def draw_bases(self, B, D, device):
        """Identical RNG call to the reference (`torch.rand((B*S, D, R))` on the CPU, then .to(device))."""
        return torch.rand((B * self.S, D, self.R)).to(device, non_blocking=True)


class Hamburger(nn.Module):
    def __init__(self, ham_channels=512, ham_kwargs=None, norm_cfg=None, **kwargs):
        super().__init__()
        ham_kwargs = ham_kwargs if ham_kwargs is not None else {}
        self.ham_in = ConvModule(ham_channels, ham_channels, 1, norm_cfg=None, act_cfg=None)
        ham_kwargs["device"] = kwargs.get("device")
        self.ham = NMF2D(ham_kwargs)
        self.ham_out = ConvModule(ham_channels, ham_channels, 1, norm_cfg=norm_cfg, act_cfg=None)


class LightHamHead(BaseDecodeHead):
    def __init__(self, ham_channels=512, ham_kwargs=None, precision=None, **kwargs):
        super().__init__(input_transform="multiple_select", **kwargs)
        self.ham_channels = ham_channels
        self.precision = precision
        self.squeeze = ConvModule(sum(self.in_channels), ham_channels, 1, conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg)
        self.hamburger = Hamburger(ham_channels, ham_kwargs if ham_kwargs is not None else {}, norm_cfg=self.norm_cfg,
                                   device=kwargs.get("device"))
        self.align = ConvModule(ham_channels, self.channels, 1, conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg)
        assert self.hamburger.ham.D == ham_channels, "NMF2D runs with S=1, D=ham_channels"
        self.sync_bn = dict(self.norm_cfg or {}).get("type") == "SyncBN"
        self._plan = None
        self._last_arena = None
        self.grad_hook = None
        self.injected_bases = None       # tests / reproducibility: raw uniform bases (B, D, R) used instead of a fresh draw

    PARAMS = ["squeeze.conv.weight", "squeeze.bn.weight", "squeeze.bn.bias", "hamburger.ham_in.conv.weight", "hamburger.ham_in.conv.bias",
              "hamburger.ham_out.conv.weight", "hamburger.ham_out.bn.weight", "hamburger.ham_out.bn.bias", "align.conv.weight",
              "align.bn.weight", "align.bn.bias", "conv_seg.weight", "conv_seg.bias"]

    def _build_plan(self):
        named = dict(self.named_parameters())
        layout, packer = ParamLayout(), ParamPacker(named)
        layout.begin_mark("head")
        for n in self.PARAMS:
            layout.add(n, named[n])
        layout.end_mark("head")
        packer.add("squeeze", ["squeeze.conv.weight"], [None])
        packer.add("ham_in", ["hamburger.ham_in.conv.weight"], ["hamburger.ham_in.conv.bias"])
        packer.add("ham_out", ["hamburger.ham_out.conv.weight"], [None])
        packer.add("align", ["align.conv.weight"], [None])
        packer.add("conv_seg", ["conv_seg.weight"], ["conv_seg.bias"])
        self._plan = SimpleNamespace(layout=layout, packer=packer, named=named, flat=None)
        return self._plan

    def flat_parameters(self):
        """Flat fp32 buffer holding every hot-path parameter (arena layout); created on first use on the GPU."""
        plan = self._plan or self._build_plan()
        if not is_flat(plan.layout, plan.flat):
            plan.flat = flatten_parameters(self, plan.layout)
        return plan.flat

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def forward(self, inputs):
        """inputs: the encoder's four NCHW-shaped stage outputs -> logits (B, num_classes, H/8, W/8)."""
        lv = self._transform_inputs(inputs)
        require_cuda(*lv)
        T = resolve_dtype(self.precision)
        B = lv[0].shape[0]
        sizes = [tuple(l.shape[2:]) for l in lv]
        flat = [l.permute(0, 2, 3, 1).reshape(-1, l.shape[1]).float().contiguous() for l in lv]   # channels-last views (no copy for our encoder)
        plan = self._plan or self._build_plan()
        dev = flat[0].device
        grad_mode = torch.is_grad_enabled()               # recorded for the Functions: inside Function.forward it is always off
        # inference (eval mode, no gradients, folding on): the packed copies and the BatchNorm folds applied to them are reused
        frozen = (not self.training) and (not grad_mode) and Fn._FOLD_BN
        bn_modes = tuple(m.training for m in self.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm))
        if not frozen:
            bump_weights_epoch()
        arena = GradArena(plan.layout, dev, self.grad_hook)
        self._last_arena = arena
        training = self.training
        ham = self.hamburger.ham
        sync = True if (self.sync_bn and training) else False
        ws = getattr(self, "_wstream", None)
        if ws is None or ws.device != dev:
            ws = self._wstream = torch.cuda.Stream(device=dev)
        sd = getattr(self, "_side_stream", None)
        if sd is None or sd.device != dev:
            sd = self._side_stream = torch.cuda.Stream(device=dev)
        # Prologue on the side stream, under the head's resize kernels: nothing here depends on the encoder's output -- the compute-dtype
        # copies of the GEMM weights, the Dropout2d channel mask and the normalised NMF bases.  HeadFn waits for `ev_prologue`
        # in front of its first GEMM.
        main = torch.cuda.current_stream()
        Fn.K.fork(sd)
        with torch.cuda.stream(sd):
            packed = plan.packer.pack(dev, T, frozen=frozen, token=bn_modes)
            bases = self.injected_bases if self.injected_bases is not None else ham.draw_bases(B, ham.D, dev)
            bases = bases.to(dev, torch.float32).contiguous()
            if tuple(bases.shape) != (B * ham.S, ham.D, ham.R):
                raise ValueError(f"NMF bases have shape {tuple(bases.shape)}, this batch needs {(B * ham.S, ham.D, ham.R)} "
                                 "(injected_bases must match the batch of the call)")
            bases.record_stream(sd)                         # injected bases usually live on the caller's stream
            bases_prepared = Fn.nmf_prepare(bases, T)
            drop_mask = None
            if training and self.dropout is not None and self.dropout.p > 0:
                keep = 1.0 - self.dropout.p
                drop_mask = (torch.rand(B, self.channels, device=dev) < keep).float() / keep         # Dropout2d: whole channels per sample
            ev_prologue = Fn.K.signal(sd)
        Fn.K.share(main, bases, drop_mask, *bases_prepared)
        fold_cache = plan.packer.fold_cache(dev, T) if frozen else None
        st = SimpleNamespace(ev_prologue=ev_prologue, bases_prepared=bases_prepared, grad=grad_mode, dtype=T, packed=packed, fold_cache=fold_cache, arena=arena, prefix="", tag="head", names=self.PARAMS, B=B, sizes=sizes, wstream=ws, side=sd,
                             steps=ham.train_steps if training else ham.eval_steps, drop_mask=drop_mask,
                             bn_sq=Fn.BNState(self.squeeze.bn, "squeeze.bn", training, sync),
                             bn_out=Fn.BNState(self.hamburger.ham_out.bn, "hamburger.ham_out.bn", training, sync),
                             bn_al=Fn.BNState(self.align.bn, "align.bn", training, sync))
        named = plan.named
        logits = Fn.HeadFn.apply(flat[0], flat[1], flat[2], bases, st, *[named[n] for n in self.PARAMS])
        h, w = sizes[0]
        return logits.view(B, h, w, self.num_classes).permute(0, 3, 1, 2)
