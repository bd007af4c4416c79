"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

Weights are a pure function of (parameter name, shape, seed) so that golden fixtures only
need to store shapes + outputs, never multi-MB weight blobs.  A "stress" init is used on
purpose: layer scales O(1) (the reference default 1e-6 would hide Block bugs),
non-trivial norm affine parameters and running statistics.
"""
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def make_tensor(name: str, shape, seed: int = 0) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "running_var":
        return 0.5 + torch.rand(shape, generator=g)
    if leaf.startswith("layer_scale"):
        return 0.5 + torch.rand(shape, generator=g)
    if leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "weight" and len(shape) == 1:           # LayerNorm / BatchNorm gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) / max(fan_in, 1) ** 0.5
    raise KeyError(name)


def make_state(shapes: dict, seed: int = 0) -> dict:
    return {k: make_tensor(k, s, seed) for k, s in shapes.items()}


def make_inputs(batch: int, height: int, width: int, num_classes: int, seed: int = 0,
                ham_channels: int = 512, rank: int = 64):
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    rgb = torch.randn(batch, 3, height, width, generator=g)
    hha = torch.randn(batch, 3, height, width, generator=g)
    label = torch.randint(0, num_classes, (batch, height, width), generator=g)
    ignore = torch.rand(batch, height, width, generator=g) < 0.05
    label = label.masked_fill(ignore, 255)
    bases = torch.rand(batch, ham_channels, rank, generator=g)
    return rgb, hha, label, bases
