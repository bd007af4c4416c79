"""mmcv.cnn stand-in: build_norm_layer + ConvModule (see ../__init__.py)."""
import torch.nn as nn


def build_norm_layer(cfg, num_features, postfix=""):
    # mmcv semantics: returns (name, layer); BN/SyncBN are abbreviated "bn";
    # eps defaults to 1e-5; `requires_grad` is applied to the affine params.
    cfg = dict(cfg)
    kind = cfg.pop("type")
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if kind == "BN":
        layer = nn.BatchNorm2d(num_features, **cfg)
    elif kind == "SyncBN":
        layer = nn.SyncBatchNorm(num_features, **cfg)
    else:
        raise KeyError(kind)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return "bn" + str(postfix), layer


class ConvModule(nn.Module):
    """conv -> norm -> act bundle; `bias` defaults to "no norm => bias"."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias="auto", conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type="ReLU"), inplace=True, **kwargs):
        super().__init__()
        assert conv_cfg is None
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            name, norm = build_norm_layer(norm_cfg, out_channels)
            self.norm_name = name
            self.add_module(name, norm)
        if self.with_activation:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)
        # mmcv default init: kaiming-normal (fan_out, relu) conv, constant norm
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)
        if self.with_norm:
            nn.init.constant_(getattr(self, self.norm_name).weight, 1)
            nn.init.constant_(getattr(self, self.norm_name).bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x
