"""Host-side runtime shared by the model mirrors: precision selection, parameter packing into
compute-dtype GEMM operands (one launch per module forward) and the flat gradient arena that the
backward kernels write into (and that the data-parallel engine all-reduces bucket by bucket)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import kernels as K

_PRECISION_DEFAULT = "fp32"


def resolve_dtype(module_precision: Optional[str]) -> torch.dtype:
    """bf16 when the module asks for it or when the caller runs under torch.autocast(bf16)
    (the reference's AMP switch, utils/train.py:322); fp32 otherwise."""
    if torch.is_autocast_enabled():
        ad = torch.get_autocast_gpu_dtype()
        if ad == torch.bfloat16:
            return torch.bfloat16
        raise RuntimeError("dformer_b200 supports fp32 and bf16 compute (autocast dtype %s requested)" % ad)
    p = module_precision or _PRECISION_DEFAULT
    if p in ("bf16", torch.bfloat16):
        return torch.bfloat16
    if p in ("fp32", torch.float32):
        return torch.float32
    raise ValueError(p)


def backend_for(dtype: torch.dtype) -> int:
    # bf16 -> tcgen05 whenever the shape qualifies; fp32 -> exact CUDA-core GEMM
    return K.AUTO if dtype == torch.bfloat16 else K.SIMT


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dformer_b200 runs on sm_100a CUDA devices only: got a %s tensor. "
                               "There is no CPU fallback (use the oracle in oracle/ for CPU checks)." % t.device)


# --------------------------------------------------------------------------------------------- layout
@dataclass
class Slot:
    name: str
    param: nn.Parameter
    offset: int
    numel: int


class ParamLayout:
    """Order + offsets of a module's parameters inside the flat fp32 gradient arena.  Members of a fused
    GEMM group are adjacent (their concatenated weight gradient is ONE contiguous [sum N, K] matrix)."""

    def __init__(self):
        self.slots: Dict[str, Slot] = {}
        self.order: List[str] = []
        self.total = 0
        self.marks: Dict[str, Tuple[int, int]] = {}     # named contiguous ranges (e.g. one Block) for DP buckets

    def add(self, name: str, p: nn.Parameter):
        off = (self.total + 7) // 8 * 8
        self.slots[name] = Slot(name, p, off, p.numel())
        self.order.append(name)
        self.total = off + p.numel()

    def begin_mark(self, tag: str):
        self.marks[tag] = ((self.total + 7) // 8 * 8, -1)

    def end_mark(self, tag: str):
        self.marks[tag] = (self.marks[tag][0], self.total)


class GradArena:
    """One zero-initialised flat fp32 buffer per backward pass; kernels write parameter gradients straight
    into views of it.  `on_range_done(lo, hi)` lets the DP engine launch a bucket all-reduce as soon as the
    gradients of a contiguous range are final (reverse execution order)."""

    def __init__(self, layout: ParamLayout, device, on_range_done: Optional[Callable] = None):
        self.layout = layout
        self.device = device
        self.buf: Optional[torch.Tensor] = None
        self.on_range_done = on_range_done

    def flat(self) -> torch.Tensor:
        if self.buf is None:
            self.buf = torch.zeros(self.layout.total, device=self.device, dtype=torch.float32)
        return self.buf

    def view(self, name: str) -> torch.Tensor:
        s = self.layout.slots[name]
        return self.flat()[s.offset:s.offset + s.numel].view(s.param.shape)

    def span(self, first: str, last: str, shape) -> torch.Tensor:
        a, b = self.layout.slots[first], self.layout.slots[last]
        return self.flat()[a.offset:b.offset + b.numel].view(shape)

    def done(self, tag: str, event=None):
        """`event`: CUDA event after which the range is final when its kernels ran on a side (wgrad) stream."""
        if self.on_range_done is not None:
            lo, hi = self.layout.marks[tag]
            self.on_range_done(self, lo, hi, event)


# --------------------------------------------------------------------------------------------- packing
_WEIGHTS_EPOCH = [0]


def bump_weights_epoch():
    """called by everything in this package that changes parameter values behind autograd's back (fused optimizer, graph replays)"""
    _WEIGHTS_EPOCH[0] += 1


def weights_epoch() -> int:
    return _WEIGHTS_EPOCH[0]


@dataclass
class GemmGroup:
    """Row-concatenation of nn.Linear / conv weights that feed ONE GEMM (e.g. l|q|q_cut)."""
    name: str
    weights: List[str]                  # parameter names in concat order
    biases: List[Optional[str]]
    kind: int = 0                       # 0 linear / 1x1 conv, 1 dense 3x3 conv (im2col order)
    n_rows: int = 0
    k: int = 0                          # logical K
    ld: int = 0                         # padded K (multiple of 8)
    w_off: int = 0
    b_off: int = -1


class ParamPacker:
    """Compute-dtype copies of every GEMM weight of a module, refreshed by a single pack_params launch
    (table cached per (device, dtype)); fp32 concatenated biases by a second launch."""

    def __init__(self, named_params: Dict[str, nn.Parameter]):
        self.params = named_params
        self.groups: Dict[str, GemmGroup] = {}
        self.w_total = 0
        self.b_total = 0
        self._cache = {}

    def add(self, name: str, weights: Sequence[str], biases: Sequence[Optional[str]], kind: int = 0) -> GemmGroup:
        g = GemmGroup(name, list(weights), list(biases), kind)
        rows = 0
        for w in weights:
            p = self.params[w]
            rows += p.shape[0]
            k = p.shape[1] * (9 if kind == 1 else 1)
            assert g.k in (0, k)
            g.k = k
        g.n_rows = rows
        g.ld = (g.k + 7) // 8 * 8
        g.w_off = (self.w_total + 63) // 64 * 64
        self.w_total = g.w_off + rows * g.ld
        if any(b is not None for b in biases):
            assert all(b is not None for b in biases)
            g.b_off = (self.b_total + 7) // 8 * 8
            self.b_total = g.b_off + rows
        self.groups[name] = g
        return g

    def _build(self, device, dtype):
        wbuf = torch.zeros(max(self.w_total, 1), device=device, dtype=dtype)       # zero padding columns stay zero
        bbuf = torch.zeros(max(self.b_total, 1), device=device, dtype=torch.float32)
        w_entries, b_entries = [], []
        views = {}
        for g in self.groups.values():
            wv = wbuf[g.w_off:g.w_off + g.n_rows * g.ld].view(g.n_rows, g.ld)
            bv = bbuf[g.b_off:g.b_off + g.n_rows] if g.b_off >= 0 else None
            views[g.name] = (wv, bv)
            r = 0
            for wn, bn in zip(g.weights, g.biases):
                p = self.params[wn]
                n = p.shape[0]
                if g.kind == 0:
                    w_entries.append((p, wv[r:], n, g.k, g.ld, 0))
                else:
                    w_entries.append((p, wv[r:], n, p.shape[1], g.ld, 1))
                if bn is not None:
                    b_entries.append((self.params[bn], bv[r:], 1, n, n, 0))
                r += n
        wt = K.build_pack_table(w_entries, device)
        bt = K.build_pack_table(b_entries, device) if b_entries else None
        ptrs = tuple(p.data_ptr() for p in self.params.values())
        return dict(wbuf=wbuf, bbuf=bbuf, views=views, wt=wt, bt=bt, ptrs=ptrs)

    def pack(self, device, dtype, frozen=False, token=None) -> Dict[str, Tuple[torch.Tensor, Optional[torch.Tensor]]]:
        """Refresh the compute-dtype copies.  frozen=True (inference: eval mode, no gradients): the copies -- and the BatchNorm folds
        applied to them, `fold_cache` -- are reused as long as neither a parameter's `_version` nor the weights epoch (bumped by every
        optimizer step / training-graph replay / training-mode forward of this package, whose kernels update parameters without
        touching autograd's version counters) nor `token` (the caller's fold-relevant state: BatchNorm modes) has moved: the 234 MB
        repack of DFormer-L per forward is 4 % of a batch-1 inference."""
        key = (str(device), dtype)
        c = self._cache.get(key)
        if c is not None and c["ptrs"] != tuple(p.data_ptr() for p in self.params.values()):
            c = None                                   # parameters were re-allocated (.to(), load with assign, ...)
        if c is None:
            c = self._build(device, dtype)
            c["stamp"], c["folded"] = None, {}
            self._cache = {key: c}
        stamp = (weights_epoch(), tuple(p._version for p in self.params.values()), token) if frozen else None
        if frozen and c["stamp"] is not None and c["stamp"] == stamp:
            return c["views"]
        c["stamp"], c["folded"] = stamp, {}
        K.pack_params(*c["wt"], K._DT[dtype])
        if c["bt"] is not None:
            K.pack_params(*c["bt"], K.F32)
        return c["views"]

    def fold_cache(self, device, dtype):
        """dict of the BatchNorm folds already applied to the current packed copies (frozen mode), else None"""
        c = self._cache.get((str(device), dtype))
        return c["folded"] if (c is not None and c.get("stamp") is not None) else None


# --------------------------------------------------------------------------------------------- flat parameters
def flatten_parameters(module, layout: ParamLayout):
    """Re-home every parameter of `layout` into ONE flat fp32 device buffer laid out exactly like the gradient
    arena (parameter i and its gradient share offsets), so the optimizer is a single fused launch over
    (params, grads, m, v) and DP buckets are plain slices.  Values are preserved; state_dict keys/shapes are
    untouched because each nn.Parameter keeps its identity and simply views into the buffer."""
    dev = next(iter(layout.slots.values())).param.device
    flat = torch.zeros(layout.total, device=dev, dtype=torch.float32)
    with torch.no_grad():
        for s in layout.slots.values():
            v = flat[s.offset:s.offset + s.numel].view(s.param.shape)
            v.copy_(s.param.data)
            s.param.data = v
    return flat


def is_flat(layout: ParamLayout, flat: Optional[torch.Tensor]) -> bool:
    if flat is None:
        return False
    first = layout.slots[layout.order[0]]
    last = layout.slots[layout.order[-1]]
    base = flat.data_ptr()
    return (first.param.data_ptr() == base + 4 * first.offset) and (last.param.data_ptr() == base + 4 * last.offset)
