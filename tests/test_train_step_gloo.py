"""Host-side logic of the data-parallel train step on CPU (gloo, world_size 2): the NCCL unique id produced by the C ABI on
rank 0 reaches every rank unchanged and each rank hands it to its own context; the flat-buffer layout keeps every
parameter 16-byte aligned.  (The exchange itself needs GPUs: tests/test_train_step_gpu.py, tools/dp_exchange_check.py.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lip2speech_b200 import train_step


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _RecordingBackend:
    """Stands in for _lib.Backend (which needs a GPU): records what init_data_parallel hands to l2s_comm_init."""
    def comm_init(self, unique_id, rank, world):
        self.args = (unique_id, rank, world)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = _RecordingBackend()
    uid = train_step.init_data_parallel(be, rank, world)
    assert be.args == (uid, rank, world) and len(uid) == 128
    out.put((rank, uid))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_reaches_every_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert got[0] == got[1] and any(b != 0 for b in got[0])      # rank 0's id, not the zero placeholder


def test_flat_layout_alignment():
    offsets, total = train_step.flat_layout([7, 165, 263168, 3, 2560, 1])
    assert offsets == [0, 8, 176, 263344, 263348, 265908] and total == 265912
    assert all(o % 4 == 0 for o in offsets) and total % 4 == 0
    assert train_step.flat_layout([]) == ([], 0)
