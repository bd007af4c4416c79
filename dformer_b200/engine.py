"""CUDA-graph training step: forward + loss + backward (+ gradient all-reduce) + fused AdamW captured ONCE and
replayed per iteration, so the ~2700 kernel launches of a DFormer-L step cost one graph launch on the host.

Streams and graphs instead of a tracing compiler: the step is ordinary eager code (our autograd.Functions issuing
C-ABI kernel launches on the current stream); `torch.cuda.graph` records it.  Inputs live in static device buffers
that `step()` refreshes with asynchronous H2D copies; the NMF bases keep the reference's CPU `torch.rand` draw
(ham_head.py:111), staged through pinned memory outside the graph."""
import os

import torch

from .runtime import bump_weights_epoch

_SKIP_BASES = os.environ.get("DFB200_PROFILE_SKIP_BASES", "0") == "1"


class GraphedTrainStep:
    """`preserve_state=True` (default) snapshots parameters, optimizer moments / step counter and BatchNorm buffers before the
    warm-up + capture passes and restores them afterwards, so building the runner (e.g. after `restore_checkpoint`) does not
    advance the training trajectory by `warmup + 1` steps on the example batch."""

    def __init__(self, model, optimizer, rgb, modal_x, label, grad_sync=None, warmup=3, use_graph=True, preserve_state=True):
        self.model, self.opt, self.sync = model, optimizer, grad_sync
        self.rgb, self.modal_x, self.label = rgb.clone(), modal_x.clone(), label.clone()
        head = model.decode_head
        ham = head.hamburger.ham
        B = rgb.shape[0]
        # two pinned host buffers for the per-step CPU draw of the NMF bases (ham_head.py:111), each guarded by the event of the
        # H2D copy that last read it: the host runs many steps ahead of the GPU, so a single buffer would be overwritten while the
        # previous step's asynchronous copy is still pending
        self._bases_host = [torch.empty((B * ham.S, ham.D, ham.R), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._bases_evt = [None, None]
        self._bases_slot = 0
        self._bases_dev = torch.empty_like(self._bases_host[0], device=rgb.device)
        self._head = head
        self.loss = None
        self.graph = None
        self._copy_stream = self._staging = self._staged = self._consumed = None
        snap = self._snapshot() if preserve_state else None
        self._draw_bases()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):          # allocator / tensor-map cache / attribute warm-up on the side stream
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._eager()
        if snap is not None:
            self._restore(snap)

    # ---- state that warm-up / capture passes must not advance
    def _snapshot(self):
        model, opt = self.model, self.opt
        snap = {"params": [p.detach().clone() for p in model.parameters()], "buffers": [b.detach().clone() for b in model.buffers()]}
        if hasattr(opt, "state") and isinstance(opt.state, list):           # optim.FusedAdamW
            snap["moments"] = [(st["m"].clone(), st["v"].clone()) for st in opt.state]
            snap["step_count"] = opt.step_count
        return snap

    def _restore(self, snap):
        model, opt = self.model, self.opt
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, q in zip(model.parameters(), snap["params"]):
                p.copy_(q)
            for b, q in zip(model.buffers(), snap["buffers"]):
                b.copy_(q)
            if "moments" in snap:
                for st, (m, v) in zip(opt.state, snap["moments"]):
                    st["m"].copy_(m)
                    st["v"].copy_(v)
                opt.step_count = snap["step_count"]
                if getattr(opt, "_dyn", None) is not None:
                    opt._dyn[1] = float(snap["step_count"])
        torch.cuda.synchronize()

    def _draw_bases(self):
        if _SKIP_BASES and self.graph is not None:       # diagnostic only (cost of the per-step host draw + H2D copy in front of the graph)
            return
        k = self._bases_slot
        self._bases_slot ^= 1
        if self._bases_evt[k] is not None:
            self._bases_evt[k].synchronize()         # the copy that last read this host buffer (two steps ago) has executed
        host = self._bases_host[k]
        torch.rand(host.shape, out=host)             # same CPU RNG stream as the reference
        self._bases_dev.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._bases_evt[k] = ev

    def _eager(self):
        # the injected bases belong to the training step only: evaluation / inference on the same model draws fresh ones
        head = self._head
        head.injected_bases = self._bases_dev
        try:
            loss, _ = self.model(self.rgb, self.modal_x, self.label)
        finally:
            head.injected_bases = None
        loss.backward()
        if self.sync is not None:
            self.sync.finish()
        self.opt.step()
        self.opt.zero_grad()
        return loss.detach()

    def stage(self, rgb, modal_x, label):
        """Input prefetch: start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream into staging
        buffers; it overlaps the training step that is currently running.  The following `step()` (called without tensors)
        consumes it.  This is the double-buffered loader every training loop uses, made explicit because the step itself is a
        CUDA graph reading fixed device buffers."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.rgb.device)
            self._staging = tuple(torch.empty_like(t) for t in (self.rgb, self.modal_x, self.label, self._bases_dev))
            self._bases_host2 = torch.empty_like(self._bases_host[0]).pin_memory()
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)              # the previous staged batch has been moved into the graph's input buffers
            self._consumed.synchronize()               # ... and its host-side bases buffer is free again (long past: one step ago)
        # the next step's NMF bases: same CPU draw as the reference (ham_head.py:111), one per step, made while the GPU is busy
        torch.rand(self._bases_host2.shape, out=self._bases_host2)
        with torch.cuda.stream(cs):
            for dst, src in zip(self._staging, (rgb, modal_x, label, self._bases_host2)):
                dst.copy_(src, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(cs)

    def step(self, rgb=None, modal_x=None, label=None):
        """Run one step on (a) the batch given here (copied into the static buffers first; async from pinned host memory), or
        (b) the batch previously handed to `stage()`, or (c) the buffers as they are."""
        if rgb is not None:
            self.rgb.copy_(rgb, non_blocking=True)
            self.modal_x.copy_(modal_x, non_blocking=True)
            self.label.copy_(label, non_blocking=True)
        elif self._staged is not None:
            torch.cuda.current_stream().wait_event(self._staged)
            for dst, src in zip((self.rgb, self.modal_x, self.label, self._bases_dev), self._staging):
                dst.copy_(src, non_blocking=True)       # device-to-device, ~0.05 ms for a DFormer-L batch of 8
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
            self._staged = None
            return self._run()
        self._draw_bases()
        return self._run()

    def _run(self):
        if self.graph is not None:
            self.graph.replay()
            bump_weights_epoch()                       # replayed AdamW: cached inference copies of the weights (runtime.ParamPacker) are stale
            if hasattr(self.opt, "step_count"):
                self.opt.step_count += 1               # host mirror of the device-side step counter the replayed AdamW kernel advances
            return self.loss
        self.loss = self._eager()
        return self.loss


# ---------------------------------------------------------------------------------------------- checkpoints (row N1)
def save_checkpoint(path, model, optimizer, epoch, iteration):
    """`Engine.save_checkpoint` (utils/engine/engine.py:101-129): one `torch.save` of
    `{"model": state_dict without a DDP "module." prefix, "optimizer": optimizer.state_dict(), "epoch", "iteration"}`.
    With `optim.FusedAdamW` the optimizer entry is in `torch.optim.AdamW`'s own format and numbering, so the file restores
    into the reference's trainer and vice versa."""
    from collections import OrderedDict
    sd = OrderedDict((k[7:] if k.split(".")[0] == "module" else k, v.detach().cpu()) for k, v in model.state_dict().items())
    torch.save({"model": sd, "optimizer": optimizer.state_dict(), "epoch": epoch, "iteration": iteration}, path)


def restore_checkpoint(path, model, optimizer=None):
    """`Engine.restore_checkpoint` (utils/engine/engine.py:159-186): loads on the CPU first, restores the model strictly and the
    optimizer, and returns `(epoch + 1, iteration)` -- the epoch to CONTINUE with, as the reference sets it.  Files whose model
    keys carry the DDP "module." prefix (the reference's `load_model(..., is_restore=True)` view) are accepted as well."""
    tmp = torch.load(path, map_location=torch.device("cpu"), weights_only=False)
    sd = tmp["model"]
    if sd and all(k.startswith("module.") for k in sd):
        sd = {k[7:]: v for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)
    if optimizer is not None:
        optimizer.load_state_dict(tmp["optimizer"])
    return tmp["epoch"] + 1, tmp["iteration"]


# ---------------------------------------------------------------------------------------------- training driver (row N1)
def is_eval(epoch, config):
    """utils/train.py:60-61."""
    return epoch > int(config.checkpoint_start_epoch) or epoch == 1 or epoch % 10 == 0


class CheckpointKeeper:
    """`Engine.save_and_link_checkpoint` (utils/engine/engine.py:136-157) without the log-directory symlinks: checkpoints are named
    `epoch-{epoch}_miou_{miou}.pt`, the five best by metric are kept.  (The reference tries to delete the sixth under a `.pth` name
    that was never written, so its files pile up; here the file that falls out of the top five is removed.)"""

    def __init__(self, checkpoint_dir, keep=5):
        self.dir, self.keep, self.state = checkpoint_dir, keep, []

    def path(self, epoch, metric):
        import os
        return os.path.join(self.dir, f"epoch-{epoch}_miou_{metric}.pt")

    def save(self, model, optimizer, epoch, iteration, metric):
        import os
        os.makedirs(self.dir, exist_ok=True)
        self.state.append({"epoch": epoch, "metric": metric})
        self.state.sort(key=lambda x: x["metric"], reverse=True)
        path = self.path(epoch, metric)
        save_checkpoint(path, model, optimizer, epoch, iteration)
        if len(self.state) > self.keep:
            worst = self.state.pop()
            try:
                os.remove(self.path(worst["epoch"], worst["metric"]))
            except OSError:
                pass
        return path


def train(runner, optimizer, config, batches, lr_policy, evaluate_fn=None, keeper=None, model=None, start_epoch=1, log=None, is_main=None):
    """The epoch / iteration loop of utils/train.py:290-470 around a step runner (normally `GraphedTrainStep`).

    * epochs run from `start_epoch` (what `restore_checkpoint` returned) to `config.nepochs` inclusive, `config.niters_per_epoch`
      iterations each; `batches(epoch)` yields the epoch's `(rgb, modal_x, label)` triples (pinned host or device tensors);
    * the learning rate follows the reference's order of operations (:352-356): the step runs first, then the rate for iteration
      `(epoch - 1) * niters + idx` is installed -- so it takes effect one step later and the very first step uses the optimizer's
      construction-time rate, exactly like the reference;
    * after the epochs selected by `is_eval` (:60-61) `evaluate_fn(epoch)` returns the mIoU; a new best is checkpointed through
      `keeper` (:405-416) with the iteration index of the epoch's last step (`engine.update_iteration`, :310).
    * only the main rank writes checkpoints (`is_main`, default: rank 0 of the default process group; the reference guards with
      `local_rank == 0`, :405); the other ranks wait at a barrier so nobody races ahead of a half-written file.
    Returns `(best_miou, history)` with one `{"epoch", "loss", "lr"[, "miou"]}` entry per epoch; the loss is read back to the host
    once per epoch (mean of the per-step device losses), not per step."""
    import torch as _torch
    import torch.distributed as _dist
    distributed = _dist.is_available() and _dist.is_initialized()
    if is_main is None:
        is_main = (not distributed) or _dist.get_rank() == 0
    n = int(config.niters_per_epoch)
    best, history = 0.0, []
    lr = optimizer.lr if hasattr(optimizer, "lr") else None
    for epoch in range(start_epoch, int(config.nepochs) + 1):
        if model is not None:
            model.train()
        it = iter(batches(epoch))
        total = None
        idx = -1
        for idx in range(n):
            rgb, modal_x, label = next(it)
            loss = runner.step(rgb, modal_x, label)
            total = loss.detach().clone() if total is None else total + loss.detach()
            lr = lr_policy.get_lr((epoch - 1) * n + idx)
            optimizer.set_lr(lr)
        entry = {"epoch": epoch, "loss": float(total / n) if total is not None else float("nan"), "lr": lr}
        if evaluate_fn is not None and is_eval(epoch, config):
            with _torch.no_grad():
                miou = float(evaluate_fn(epoch))
            entry["miou"] = miou
            if miou > best:
                best = miou
                if keeper is not None:
                    if is_main:
                        keeper.save(model if model is not None else runner.model, optimizer, epoch, idx, miou)
                    if distributed:
                        _dist.barrier()
        history.append(entry)
        if log is not None:
            log(entry)
    return best, history
