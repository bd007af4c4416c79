"""dfb200_nccl_* thin wrappers (include/dfb200.h): libnccl is resolved at run time; argument errors are
reported through the C ABI's error channel.  CPU part: no device calls.  GPU part: a one-rank communicator (the by-value
ncclUniqueId hand-over, in-place average on a side stream) and, with >= 2 GPUs, two processes averaging a gradient-like buffer."""
import os
import subprocess
import sys

import pytest
import torch

from dformer_b200.parallel import NativeNccl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libnccl_is_found_at_run_time_and_reports_its_version():
    v = NativeNccl.version()
    assert v >= 21000, v                       # ncclAvg / ncclBfloat16 need NCCL >= 2.10


def test_argument_errors_come_back_as_messages_not_crashes():
    from dformer_b200._lib import lib
    with pytest.raises(RuntimeError, match="nccl_all_reduce"):
        lib().nccl_all_reduce(None, None, 4, 0, 1, None)
    with pytest.raises(RuntimeError, match="nccl_comm_init"):
        lib().nccl_comm_init(None, 2, 0, None)
    with pytest.raises(RuntimeError, match="null"):
        lib().nccl_unique_id(None)


@pytest.mark.gpu
def test_one_rank_communicator_averages_in_place_on_a_side_stream():
    torch.cuda.set_device(0)
    comm = NativeNccl(NativeNccl.unique_id(), 1, 0)
    try:
        side = torch.cuda.Stream()
        for dtype in (torch.float32, torch.bfloat16, torch.float64):
            x = torch.randn(100003, device="cuda").to(dtype)
            ref = x.clone()
            side.wait_stream(torch.cuda.current_stream())
            comm.all_reduce_(x, average=True, stream=side)
            torch.cuda.current_stream().wait_stream(side)
            assert torch.equal(x, ref), dtype
            comm.all_reduce_(x, average=False)
            torch.cuda.synchronize()
            assert torch.equal(x, ref), dtype
        from dformer_b200._lib import lib
        with pytest.raises(RuntimeError, match="dtype"):
            lib().nccl_all_reduce(comm._comm, x.data_ptr(), 4, 7, 1, None)
    finally:
        comm.destroy()


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from dformer_b200.parallel import NativeNccl
rank, world, idfile = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
torch.cuda.set_device(rank)
uid = open(idfile, "rb").read()
comm = NativeNccl(uid, world, rank)
g = torch.Generator().manual_seed(1234)
full = torch.randn(world, 1 << 20, generator=g)
x = full[rank].cuda()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
comm.all_reduce_(x, average=True, stream=s)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
ref = full.double().mean(0).float().cuda()
err = (x - ref).abs().max().item()
comm.destroy()
assert err < 1e-6, err
print("rank", rank, "ok", err)
"""


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_processes_average_a_buffer_without_torch_distributed(tmp_path):
    idfile = tmp_path / "nccl_id.bin"
    idfile.write_bytes(NativeNccl.unique_id())
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(idfile)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
