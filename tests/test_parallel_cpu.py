"""Data-parallel gradient exchange on CPU (gloo, world_size 2): bucket coalescing of arena ranges finished in
reverse execution order, averaging, and the DP-equivalence of the exchanged gradients."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Layout:
    def __init__(self, marks, total):
        self.marks, self.total = marks, total


class _FakeArena:
    """GradArena stand-in with a CPU buffer (the real one is filled by CUDA kernels)."""

    def __init__(self, layout, on_range_done):
        self.layout, self.buf, self.on_range_done = layout, torch.zeros(layout.total), on_range_done

    def flat(self):
        return self.buf

    def done(self, tag, event=None):
        lo, hi = self.layout.marks[tag]
        self.on_range_done(self, lo, hi, event)


class _FakeModule(torch.nn.Module):
    grad_hook = None

    def _build_plan(self):
        return None


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dformer_b200.parallel import GradSync
    model = torch.nn.Sequential(_FakeModule())
    sync = GradSync(model, bucket_mb=4 * 300 / (1024 * 1024))       # bucket = 300 floats
    marks = {"stem": (0, 96), "b0": (96, 296), "b1": (296, 504), "head": (504, 640)}
    arena = _FakeArena(_Layout(marks, 640), model[0].grad_hook)
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(640, generator=g)
    arena.buf.copy_(local)
    for tag in ("head", "b1", "b0", "stem"):                        # reverse execution order, as backward finishes them
        arena.done(tag)
    fired_before_finish = sync.launched
    sync.finish()
    all_local = [torch.randn(640, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    want = torch.stack(all_local).mean(0)
    q.put((rank, torch.allclose(arena.buf, want, atol=1e-6), fired_before_finish, sync.launched))
    dist.destroy_process_group()


def test_gradsync_buckets_and_average_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, early, total in res:
        assert ok, f"rank {rank}: averaged gradients differ from the mean of the per-rank gradients"
        assert early >= 1, "a full bucket must be all-reduced before finish() (overlap with backward)"
        assert total == early + 1, (early, total)                   # one flush of the partial tail bucket


def test_gradsync_is_a_noop_without_process_group():
    sys.path.insert(0, ROOT)
    from dformer_b200.parallel import GradSync
    model = torch.nn.Sequential(_FakeModule())
    sync = GradSync(model)
    assert not sync.enabled
    model[0].grad_hook(None, 0, 10, None)
    sync.finish()
