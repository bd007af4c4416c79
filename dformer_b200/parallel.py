"""Data-parallel gradient exchange: one process per GPU, batch sharded across ranks,
parameters replicated, gradients averaged with NCCL all-reduce over NVLink/NVSwitch, overlapped with backward.

The backward Functions write parameter gradients into flat per-module arenas in reverse execution order and
call `arena.done(tag)` when a contiguous range (a Block, a downsample layer, the head) is final.  This engine
coalesces those ranges into buckets and launches `all_reduce(AVG)` on a side stream as soon as a bucket is
complete, exactly where DDP's hooks would fire -- without per-parameter hooks or bucket copies (the arena IS
the bucket).  `finish()` joins the side stream before the optimizer step."""
import torch
import torch.distributed as dist


import weakref

import os as _os
_SKIP_AR = _os.environ.get("DFB200_PROFILE_SKIP_AR", "0") == "1"      # profiling only: drop the gradient all-reduces (wrong results) to see their cost
_GRAD_SYNCS = weakref.WeakSet()
_PEER_CHAIN = {"event": None, "capturing": False}     # completion event of the most recent peer-memory exchange of this step


def _last_peer_event():
    ev = _PEER_CHAIN["event"]
    if ev is not None and _PEER_CHAIN["capturing"] != (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
        ev = _PEER_CHAIN["event"] = None              # recorded inside / outside a graph capture that is over: nothing to order against
    return ev


class GradSync:
    def __init__(self, model, bucket_mb=25.0, group=None):
        self.model, self.group = model, group
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.stream = torch.cuda.Stream() if (self.enabled and torch.cuda.is_available()) else None
        self.pending = {}        # id(arena) -> [arena, lo, hi, events] contiguous finished-but-unsent range
        self.handles = []
        self.launched = 0
        self.inflight = False    # an all-reduce has been launched on self.stream since the last join (see small_all_reduce_)
        _GRAD_SYNCS.add(self)
        for m in model.modules():
            if hasattr(m, "grad_hook") and hasattr(m, "_build_plan"):
                m.grad_hook = self._on_range_done

    def _fire(self, arena, lo, hi, events=()):
        buf = arena.flat()[lo:hi]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            for ev in events:                       # gradients written on the wgrad side stream
                self.stream.wait_event(ev)
            pe = _last_peer_event()
            if pe is not None:                      # collectives of the step form ONE total order on every rank (see small_all_reduce_)
                self.stream.wait_event(pe)
            self.inflight = True
            with torch.cuda.stream(self.stream):
                if not _SKIP_AR:
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
        self.inflight = False
        _PEER_CHAIN["event"] = None


class PeerExchange:
    """Small-message all-reduce over NVLink peer memory (csrc/peer.cu) for the SyncBatchNorm statistics.

    One exchange buffer per rank, mapped by every peer of the box through CUDA IPC; `all_reduce_(t)` sums a small fp32 / fp64
    tensor in place with one single-CTA kernel on the current stream (graph-capturable).  Built lazily per process group by
    `peer_exchange(group)`; unavailable (-> NCCL all-reduce) when the ranks are not all CUDA ranks of one host."""

    def __init__(self, group=None):
        import ctypes
        import socket

        from ._lib import lib
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ok = False
        self._ptrs = []
        L = lib()
        mine = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        err = ""
        try:
            L.peer_alloc(ctypes.byref(mine))
            L.peer_export(mine, handle)
        except RuntimeError as e:                                  # e.g. IPC not permitted in this container
            err = str(e)
        info = (socket.gethostname(), torch.cuda.current_device(), bytes(handle), err)
        infos = [None] * self.world
        dist.all_gather_object(infos, info, group=group)
        same_host = all(i[0] == infos[0][0] for i in infos) and len({i[1] for i in infos}) == self.world
        good = same_host and not any(i[3] for i in infos) and self.world <= 16
        bases = []
        if good:
            try:
                for r, (_, _, h, _) in enumerate(infos):
                    if r == self.rank:
                        bases.append(mine.value)
                    else:
                        p = ctypes.c_void_p()
                        L.peer_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p))
                        self._ptrs.append(p)
                        bases.append(p.value)
            except RuntimeError:
                good = False
        flag = torch.tensor([1 if good else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)          # everybody or nobody (also orders set-up before first use)
        self._mine = mine
        if int(flag.item()) == 1:
            self.table = torch.tensor(bases, device="cuda", dtype=torch.int64)
            self.ok = True

    def all_reduce_(self, t):
        from ._lib import lib
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)
        lib().peer_allreduce(t.data_ptr(), t.data_ptr(), 0 if t.dtype == torch.float32 else 2, t.numel(), self.table.data_ptr(), self.rank, self.world,
                             torch.cuda.current_stream().cuda_stream)
        return t


_PEER = {}


def peer_exchange(group=None):
    """The PeerExchange of a process group (None when it cannot be used: CPU / gloo ranks, several hosts, IPC refused)."""
    key = id(group) if group is not None else None
    if key not in _PEER:
        import os
        usable = (torch.cuda.is_available() and dist.get_backend(group) == "nccl" and dist.get_world_size(group) > 1
                  and os.environ.get("DFB200_PEER_EXCHANGE", "1") != "0")        # 0: measure against NCCL's all-reduce (bench A/B only)
        ex = PeerExchange(group) if usable else None
        _PEER[key] = ex if (ex is not None and ex.ok) else None
    return _PEER[key]


def small_all_reduce_(t, group=None):
    """In-place sum of a small statistics tensor across the group: peer-memory kernel on one NVLink box, else the backend's all_reduce."""
    ex = peer_exchange(group)
    if ex is not None and t.numel() * t.element_size() <= 16 * 1024:
        # The exchange kernel spins on its peers.  Two kernels that wait on other ranks must never be co-scheduled in different
        # orders on different ranks (rank 0 inside collective A waiting for rank 1, rank 1 inside collective B waiting for rank 0,
        # each blocking the other's launch queue), so every collective of a step -- these exchanges on the RGB and depth streams
        # and the NCCL gradient all-reduces on the communication stream -- is chained into one total order, identical on all
        # ranks because they run the same program: an exchange starts after the previous exchange and after any gradient
        # all-reduce launched before it; an all-reduce starts after the last exchange (GradSync._fire).
        cur = torch.cuda.current_stream()
        for gs in list(_GRAD_SYNCS):
            if gs.inflight and gs.stream is not None:
                cur.wait_stream(gs.stream)
                gs.inflight = False
        pe = _last_peer_event()
        if pe is not None:
            cur.wait_event(pe)
        ex.all_reduce_(t)
        ev = torch.cuda.Event()
        ev.record(cur)
        _PEER_CHAIN["event"], _PEER_CHAIN["capturing"] = ev, torch.cuda.is_current_stream_capturing()
        return t
    dist.all_reduce(t, group=group)
    return t


class NativeNccl:
    """Host mirror of the `dfb200_nccl_*` C-ABI wrappers (include/dfb200.h): an NCCL communicator created WITHOUT torch.distributed,
    for hosts that drive libdformer_b200.so directly.  The Python product path above keeps torch.distributed's NCCL backend; this
    class exists so that the wrappers are exercised (tests/test_nccl_abi*.py) and documents the call sequence a C/C++ host follows:

        id = NativeNccl.unique_id()            # rank 0; ship the 128 bytes to the other ranks out of band
        comm = NativeNccl(id, nranks, rank)    # collective; the current CUDA device is this rank's GPU
        comm.all_reduce_(arena_slice, average=True, stream=comm_stream)     # per finished gradient bucket
    """

    _DT = {torch.float32: 0, torch.bfloat16: 1, torch.float64: 2}

    @staticmethod
    def version():
        import ctypes

        from ._lib import lib
        v = ctypes.c_int(0)
        lib().nccl_version(ctypes.byref(v))
        return v.value

    @staticmethod
    def unique_id() -> bytes:
        import ctypes

        from ._lib import lib
        buf = (ctypes.c_ubyte * 128)()
        lib().nccl_unique_id(buf)
        return bytes(buf)

    def __init__(self, unique_id: bytes, nranks: int, rank: int):
        import ctypes

        from ._lib import lib
        assert len(unique_id) == 128
        self._comm = ctypes.c_void_p()
        self.nranks, self.rank = nranks, rank
        lib().nccl_comm_init((ctypes.c_ubyte * 128).from_buffer_copy(unique_id), nranks, rank, ctypes.byref(self._comm))

    def all_reduce_(self, t, average=True, stream=None):
        from ._lib import lib
        assert t.is_cuda and t.is_contiguous() and t.dtype in self._DT
        st = stream if stream is not None else torch.cuda.current_stream(t.device)
        lib().nccl_all_reduce(self._comm, t.data_ptr(), t.numel(), self._DT[t.dtype], 1 if average else 0, st.cuda_stream)
        return t

    def destroy(self):
        from ._lib import lib
        if self._comm is not None and self._comm.value:
            lib().nccl_comm_destroy(self._comm)
        self._comm = None
