"""Fused AdamW over the flat parameter / gradient buffers (next-row N1 of SURVEY.md section 8f).

One `dfb200_adamw` launch per flat module buffer replaces the ~1000 per-tensor launches of an eager
optimizer.  Parameter groups follow the reference exactly (utils/init_func.py:26-70 as used by
utils/train.py:207-216): Linear/Conv weights decay; Linear/Conv biases and BatchNorm affine parameters do
not; and -- a quirk of the reference worth keeping for parity -- `layer_scale_*` and the custom channels-last
`LayerNorm` parameters land in NEITHER group and are therefore never updated (lr multiplier 0)."""
import torch
import torch.nn as nn

from . import kernels as K


def _flat_modules(model):
    return [m for m in model.modules() if hasattr(m, "flat_parameters") and hasattr(m, "_build_plan")]


class FusedAdamW:
    def __init__(self, model, lr=6e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, reference_groups=True):
        self.model, self.lr, self.betas, self.eps, self.weight_decay = model, lr, betas, eps, weight_decay
        self.step_count = 0
        self.state = []
        self._dyn = None          # device {lr, step}: the schedule state a captured CUDA graph reads at replay time
        for mod in _flat_modules(model):
            flat = mod.flat_parameters()
            layout = mod._plan.layout
            decay_ids, nodecay_ids = set(), set()
            for sub in mod.modules():
                if isinstance(sub, (nn.Linear, nn.Conv1d, nn.Conv2d, nn.Conv3d)):
                    decay_ids.add(id(sub.weight))
                    if sub.bias is not None:
                        nodecay_ids.add(id(sub.bias))
                elif isinstance(sub, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.GroupNorm, nn.LayerNorm)):
                    nodecay_ids.update(id(p) for p in (sub.weight, sub.bias) if p is not None)
            wd = torch.zeros_like(flat)
            lrm = torch.zeros_like(flat)
            for s in layout.slots.values():
                sl = slice(s.offset, s.offset + s.numel)
                if not s.param.requires_grad:
                    continue
                if id(s.param) in decay_ids:
                    wd[sl], lrm[sl] = weight_decay, 1.0
                elif id(s.param) in nodecay_ids or not reference_groups:
                    lrm[sl] = 1.0
            self.state.append(dict(mod=mod, m=torch.zeros_like(flat), v=torch.zeros_like(flat), wd=wd, lrm=lrm))

    def set_lr(self, lr):
        """Update the learning rate read by (possibly graph-captured) optimizer launches."""
        self.lr = lr
        if self._dyn is not None:
            self._dyn[0] = lr

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    @torch.no_grad()
    def step(self, lr=None):
        """Uses the gradient arenas written by the last backward pass (p.grad views alias them)."""
        self.step_count += 1
        lr = self.lr if lr is None else lr
        dev = self.state[0]["m"].device if self.state else None
        if self._dyn is None and dev is not None:
            self._dyn = torch.tensor([lr, 0.0], device=dev, dtype=torch.float32)
        if self._dyn is not None:
            if not torch.cuda.is_current_stream_capturing() and lr != self.lr:
                self._dyn[0] = lr
            self._dyn[1:].add_(1.0)                   # device-side step counter (captured as a kernel)
        for st in self.state:
            mod = st["mod"]
            flat = mod.flat_parameters()
            arena = getattr(mod, "_last_arena", None)
            if arena is None or arena.buf is None:        # no backward pass since the last step
                continue
            K.adamw(flat, arena.buf, st["m"], st["v"], lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                    wd_arr=st["wd"], lr_arr=st["lrm"], dyn=self._dyn)


class WarmUpPolyLR:
    """utils/lr_policy.py:22-34."""

    def __init__(self, start_lr, lr_power, total_iters, warmup_steps):
        self.start_lr, self.lr_power, self.total_iters, self.warmup_steps = start_lr, lr_power, total_iters + 0.0, warmup_steps

    def get_lr(self, cur_iter):
        if cur_iter < self.warmup_steps:
            return self.start_lr * (cur_iter / self.warmup_steps)
        return self.start_lr * ((1 - float(cur_iter) / self.total_iters) ** self.lr_power)
