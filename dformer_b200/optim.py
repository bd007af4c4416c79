"""Fused AdamW over the flat parameter / gradient buffers (next-row N1 of SURVEY.md section 8f).

One `dfb200_adamw` launch per flat module buffer replaces the ~1000 per-tensor launches of an eager
optimizer.  Parameter groups follow the reference exactly (utils/init_func.py:26-70 as used by
utils/train.py:207-216): Linear/Conv weights decay; Linear/Conv biases and BatchNorm affine parameters do
not; and -- a quirk of the reference worth keeping for parity -- `layer_scale_*` and the custom channels-last
`LayerNorm` parameters land in NEITHER group and are therefore never updated (lr multiplier 0)."""
import torch
import torch.nn as nn

from . import kernels as K
from .runtime import bump_weights_epoch


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
            decays = torch.zeros_like(flat)
            for s in layout.slots.values():
                sl = slice(s.offset, s.offset + s.numel)
                if not s.param.requires_grad:
                    continue
                if id(s.param) in decay_ids:
                    wd[sl], lrm[sl], decays[sl] = weight_decay, 1.0, 1.0
                elif id(s.param) in nodecay_ids or not reference_groups:
                    lrm[sl] = 1.0
            # the layout is kept here: `mod._plan` is reset by any later `model._apply` (e.g. a redundant `.cuda()`)
            self.state.append(dict(mod=mod, layout=layout, m=torch.zeros_like(flat), v=torch.zeros_like(flat), wd=wd, lrm=lrm, decays=decays))

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
        bump_weights_epoch()                          # the kernel below rewrites parameters without touching autograd's version counters
        self.step_count += 1
        lr = self.lr if lr is None else lr
        dev = self.state[0]["m"].device if self.state else None
        if self._dyn is None and dev is not None:
            self._dyn = torch.tensor([lr, float(self.step_count - 1)], device=dev, dtype=torch.float32)
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


    # ------------------------------------------------------------------ checkpoint interchange with the reference
    def _reference_groups(self):
        """The two parameter lists `group_weight` hands to `torch.optim.AdamW` (utils/init_func.py:26-70, utils/train.py:207-216),
        in its module-traversal order: [weights that decay], [biases + BatchNorm affine parameters]."""
        decay, no_decay = [], []
        for sub in self.model.modules():
            if isinstance(sub, (nn.Linear, nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
                decay.append(sub.weight)
                if sub.bias is not None:
                    no_decay.append(sub.bias)
            elif isinstance(sub, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.GroupNorm, nn.LayerNorm)):
                no_decay.extend(p for p in (sub.weight, sub.bias) if p is not None)
        return decay, no_decay

    def _moment_views(self):
        """id(parameter) -> (exp_avg view, exp_avg_sq view, updated?) inside the flat moment buffers."""
        views = {}
        for st in self.state:
            updated = (st["lrm"] != 0).cpu()
            for s in st["layout"].slots.values():
                sl = slice(s.offset, s.offset + s.numel)
                views[id(s.param)] = (st["m"][sl].view(s.param.shape), st["v"][sl].view(s.param.shape), bool(updated[s.offset]))
        return views

    def _sync_step_count(self):
        """The device-side counter is authoritative: a replayed CUDA graph advances it without running `step()` on the host
        (`engine.GraphedTrainStep` mirrors it, but a graph replayed by other code would not).  One sync at checkpoint time."""
        if self._dyn is not None:
            self.step_count = int(round(float(self._dyn[1].item())))
        return self.step_count

    def state_dict(self):
        """A `torch.optim.AdamW.state_dict()` of the optimizer the reference builds (same parameter numbering, same two groups), so
        that `Engine.save_checkpoint` / `restore_checkpoint` files (utils/engine/engine.py:101-186) move between the two
        implementations.  Like torch, only parameters that have been stepped carry state; the never-used `stem_e_fc*` layers
        (no gradient, `DFormer.py:202-203`) and frozen parameters carry none."""
        self._sync_step_count()
        decay, no_decay = self._reference_groups()
        skeleton = torch.optim.AdamW([dict(params=decay, lr=self.lr), dict(params=no_decay, weight_decay=0.0, lr=self.lr)], lr=self.lr,
                                     betas=self.betas, eps=self.eps, weight_decay=self.weight_decay).state_dict()
        views = self._moment_views()
        state = {}
        if self.step_count > 0:
            for idx, p in enumerate(decay + no_decay):
                mv = views.get(id(p))
                if mv is not None and mv[2] and p.requires_grad:
                    state[idx] = {"step": torch.tensor(float(self.step_count)), "exp_avg": mv[0].clone(), "exp_avg_sq": mv[1].clone()}
        skeleton["state"] = state
        return skeleton

    def load_state_dict(self, sd):
        """Inverse of `state_dict()`; also accepts the state of a stock `torch.optim.AdamW` built with the reference's groups."""
        decay, no_decay = self._reference_groups()
        params = decay + no_decay
        saved = [i for g in sd["param_groups"] for i in g["params"]]
        if len(saved) != len(params):
            raise ValueError(f"optimizer state has {len(saved)} parameters, the model's reference groups have {len(params)}")
        views = self._moment_views()
        for st in self.state:
            st["m"].zero_()
            st["v"].zero_()
        steps = set()
        for pos, idx in enumerate(saved):
            e = sd["state"].get(idx)
            if e is None:
                continue
            mv = views.get(id(params[pos]))
            if mv is None:
                raise ValueError(f"optimizer state for parameter {idx}, which is not on the hot path")
            if tuple(e["exp_avg"].shape) != tuple(mv[0].shape):
                raise ValueError(f"optimizer state {idx}: shape {tuple(e['exp_avg'].shape)} != parameter shape {tuple(mv[0].shape)}")
            mv[0].copy_(e["exp_avg"])
            mv[1].copy_(e["exp_avg_sq"])
            steps.add(int(float(e["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}); the fused optimizer keeps one")
        self.step_count = steps.pop() if steps else 0
        g0 = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g0["lr"], tuple(g0["betas"]), g0["eps"], g0["weight_decay"]
        for st in self.state:                       # the per-element decay array follows the restored coefficient
            st["wd"].copy_(st["decays"] * self.weight_decay)
        if self._dyn is not None:
            self._dyn[0] = self.lr
            self._dyn[1] = float(self.step_count)


class WarmUpPolyLR:
    """utils/lr_policy.py:22-34."""

    def __init__(self, start_lr, lr_power, total_iters, warmup_steps):
        self.start_lr, self.lr_power, self.total_iters, self.warmup_steps = start_lr, lr_power, total_iters + 0.0, warmup_steps

    def get_lr(self, cur_iter):
        if cur_iter < self.warmup_steps:
            return self.start_lr * (cur_iter / self.warmup_steps)
        return self.start_lr * ((1 - float(cur_iter) / self.total_iters) ** self.lr_power)
