"""N>1 path on CPU (gloo, world_size 2): the batch is sharded by clips with no data-path collective; only the
timing reduction (max over ranks) and the barrier use torch.distributed — the same calls bench.py makes with NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lip2speech_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.clip_range(total, rank, world)
    # every rank "decodes" its own clips: here the payload is just the clip ids
    mine = torch.arange(lo, hi, dtype=torch.int64)
    t = sharding.max_over_ranks(float(10 + rank))            # pretend rank r took 10+r ms
    gathered = sharding.gather_clip_ids(mine, total)
    if rank == 0:
        out.put((t, gathered.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_clip_ranges_partition_the_batch():
    for total in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            ranges = [sharding.clip_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 33, q)) for r in range(2)]
    for p in procs: p.start()
    t, ids = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert t == 11.0                      # max over ranks
    assert ids == list(range(33))         # contiguous shards cover every clip exactly once
