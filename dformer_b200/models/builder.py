"""Mirror of `models/builder.py: EncoderDecoder` (:91-235) for the DFormer + LightHamHead configuration.

Same constructor signature, attributes (`encoder_backbone`, `decode_head`, `aux_head`, `criterion`,
`channels`, `cfg`), methods (`forward`, `encode_decode`, `encode`, `decode`, `init_weights`) and
state_dict layout as the reference.  Deviations, all documented in DESIGN.md:
  * the fork's `DFormer.forward` returns `(outs, None)` (DFormer.py:305) which its own `encode_decode`
    cannot consume; here `encode_decode` unwraps the tuple (the only way the fork's model can run);
  * bilinear x8 upsample + CrossEntropy(ignore_index) + masked mean run as one fused kernel pair;
  * decoders other than 'ham' and the few-shot wrappers are out of scope."""
import torch
import torch.nn as nn

from .. import functions as Fn
from ..runtime import require_cuda


def _init_decode_head(module, norm_layer, bn_eps, bn_momentum):
    """utils/init_func.py:7-23 as called by builder.py:185-187: kaiming-normal (fan_in, relu) on every conv
    of the head, BN eps/momentum from the config and affine reset."""
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
            nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")
        elif isinstance(m, norm_layer):
            m.eps, m.momentum = bn_eps, bn_momentum
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class EncoderDecoder(nn.Module):
    def __init__(self, cfg=None, criterion=nn.CrossEntropyLoss(reduction="none", ignore_index=255), norm_layer=nn.BatchNorm2d, syncbn=False):
        super().__init__()
        from .encoders import DFormer as enc
        self.norm_layer, self.cfg = norm_layer, cfg
        table = {"DFormer-Large": (enc.DFormer_Large, [96, 192, 288, 576]), "DFormer-Base": (enc.DFormer_Base, [64, 128, 256, 512]),
                 "DFormer-Small": (enc.DFormer_Small, [64, 128, 256, 512]), "DFormer-Tiny": (enc.DFormer_Tiny, [32, 64, 128, 256])}
        if cfg.backbone not in table:
            raise NotImplementedError(f"backbone {cfg.backbone!r}: only the DFormer-T/S/B/L hot path is implemented")
        backbone, self.channels = table[cfg.backbone]
        norm_cfg = dict(type="SyncBN" if syncbn else "BN", requires_grad=True)
        precision = getattr(cfg, "precision", None)
        dpr = cfg.drop_path_rate if getattr(cfg, "drop_path_rate", None) is not None else 0.1
        self.encoder_backbone = backbone(drop_path_rate=dpr, norm_cfg=norm_cfg, precision=precision)
        self.aux_head = None
        if cfg.decoder != "ham":
            raise NotImplementedError(f"decoder {cfg.decoder!r}: only LightHamHead ('ham') is on the hot path")
        if getattr(cfg, "aux_rate", 0) != 0:
            raise NotImplementedError("aux FCN head (aux_rate != 0) is not used by any shipped config and is out of scope")
        from .decoders.ham_head import LightHamHead
        self.decode_head = LightHamHead(in_channels=self.channels[1:], num_classes=cfg.num_classes, in_index=[1, 2, 3], norm_cfg=norm_cfg,
                                        channels=cfg.decoder_embed_dim, device=getattr(cfg, "device", None), precision=precision)
        self.criterion = criterion
        if self.criterion:
            self.init_weights(cfg, pretrained=getattr(cfg, "pretrained_model", None))

    def init_weights(self, cfg, pretrained=None):
        if pretrained:
            self.encoder_backbone.init_weights(pretrained=pretrained)
        _init_decode_head(self.decode_head, self.norm_layer, cfg.bn_eps, cfg.bn_momentum)

    # ---------------------------------------------------------------------------------------------
    def _ignore_index(self):
        ii = getattr(self.criterion, "ignore_index", 255) if self.criterion is not None else 255
        return int(getattr(self.cfg, "background", ii))

    def _small_logits(self, rgb, modal_x):
        feats = self.encoder_backbone(rgb, modal_x)
        if isinstance(feats, tuple):              # the fork returns (outs, None)
            feats = feats[0]
        return feats, self.decode_head.forward(feats)

    def _upsample(self, small, size, label=None):
        B, ncls, h, w = small.shape
        s2d = small.permute(0, 2, 3, 1).reshape(B * h * w, ncls)
        if not s2d.is_contiguous():
            s2d = s2d.contiguous()
        H, W = int(size[0]), int(size[1])
        if label is None:
            return None, Fn.UpsampleFn.apply(s2d, (B, h, w, ncls, H, W))
        lab = label.long().contiguous()
        want_out = getattr(self.cfg, "return_logits", True)
        loss, out = Fn.UpsampleCEFn.apply(s2d, lab, (B, h, w, ncls, H, W, self._ignore_index(), want_out, torch.is_grad_enabled()))
        return loss, (out if want_out else None)

    def encode_decode(self, rgb, modal_x):
        """backbone -> head -> bilinear resize to the input size (builder.py:193-208)."""
        _, small = self._small_logits(rgb, modal_x)
        return self._upsample(small, rgb.shape[-2:])[1]

    def encode(self, rgb, modal_x):
        return self.encoder_backbone(rgb, modal_x)

    def decode(self, x, rgb):
        if isinstance(x, tuple):
            x = x[0]
        return self._upsample(self.decode_head.forward(x), rgb.shape[-2:])[1]

    def forward(self, rgb, modal_x=None, label=None):
        """eval: logits (B, ncls, H, W); train (label given): (loss, logits) with
        loss = CE(reduction='none', ignore_index)[label != background].mean()   (builder.py:224-235)."""
        require_cuda(rgb, modal_x, label)
        _, small = self._small_logits(rgb, modal_x)
        if label is None:
            return self._upsample(small, rgb.shape[-2:])[1]
        loss, out = self._upsample(small, rgb.shape[-2:], label)
        return loss, out
