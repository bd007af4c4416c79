"""mmcv.cnn.bricks.transformer stand-in: DropPath via build_dropout; FFN unused."""
import torch
import torch.nn as nn


class FFN(nn.Module):  # imported by the reference, never instantiated on the hot path
    pass


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        return x.div(keep) * mask.floor()


def build_dropout(cfg, default_args=None):
    cfg = dict(cfg)
    kind = cfg.pop("type")
    assert kind == "DropPath", kind
    return DropPath(**cfg)
