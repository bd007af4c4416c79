"""Data-parallel gradient exchange: one process per GPU, batch sharded across ranks,
parameters replicated, gradients averaged with NCCL all-reduce over NVLink/NVSwitch, overlapped with backward.

The backward Functions write parameter gradients into flat per-module arenas in reverse execution order and
call `arena.done(tag)` when a contiguous range (a Block, a downsample layer, the head) is final.  This engine
coalesces those ranges into buckets and launches `all_reduce(AVG)` on a side stream as soon as a bucket is
complete, exactly where DDP's hooks would fire -- without per-parameter hooks or bucket copies (the arena IS
the bucket).  `finish()` joins the side stream before the optimizer step."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, model, bucket_mb=25.0, group=None):
        self.model, self.group = model, group
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.stream = torch.cuda.Stream() if (self.enabled and torch.cuda.is_available()) else None
        self.pending = {}        # id(arena) -> [arena, lo, hi, events] contiguous finished-but-unsent range
        self.handles = []
        self.launched = 0
        for m in model.modules():
            if hasattr(m, "grad_hook") and hasattr(m, "_build_plan"):
                m.grad_hook = self._on_range_done

    def _fire(self, arena, lo, hi, events=()):
        buf = arena.flat()[lo:hi]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            for ev in events:                       # gradients written on the wgrad side stream
                self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            buf.record_stream(self.stream)
        else:                       # gloo (CPU tests): no AVG op, no streams
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(dist.get_world_size(self.group))
        self.launched += 1

    def _on_range_done(self, arena, lo, hi, event=None):
        if not self.enabled:
            return
        key = id(arena)
        cur = self.pending.get(key)
        if cur is not None and cur[1] == hi:          # ranges finish in reverse order: extend downwards
            cur[1] = lo
        elif cur is not None and cur[2] == lo:
            cur[2] = hi
        else:
            if cur is not None:
                self._fire(*cur)
            cur = [arena, lo, hi, []]
            self.pending[key] = cur
        if event is not None:
            cur[3].append(event)
        if cur[2] - cur[1] >= self.bucket_elems:
            self._fire(*cur)
            del self.pending[key]

    def finish(self):
        """Flush partial buckets and make the compute stream wait for all gradient all-reduces."""
        if not self.enabled:
            return
        for cur in list(self.pending.values()):
            self._fire(*cur)
        self.pending.clear()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
