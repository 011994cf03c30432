"""Data-parallel gradient exchange check (run under torchrun, one rank per GPU): NCCL all-reduce through the C ABI
(l2s_comm_init / l2s_allreduce_grads) + clip + AdamW(amsgrad) vs the CPU reference fed with the mean gradient.
Also times the exchange of the reference's full gradient (38.44 M fp32 = 153.7 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from lip2speech_b200 import _lib
from lip2speech_b200.train_step import ClipAdamW, init_data_parallel
from oracle import train_oracle as TO

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = _lib.backend(local)
init_data_parallel(be, rank, world)

g = torch.Generator().manual_seed(3)
shapes = [(7,), (33, 5), (1024, 257), (3,)]
params = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
steps = 3
world_grads = []
for r in range(world):
    gr = torch.Generator().manual_seed(100 + r)
    world_grads.append([[torch.randn(p.shape, generator=gr) * 3.0 for p in params] for _ in range(steps)])
ref_p, ref_norms = TO.clip_adamw_steps(params, None, lr=1e-3, weight_decay=1e-2, max_norm=1.0, world_grads=world_grads)

cu = [torch.nn.Parameter(p.clone().cuda()) for p in params]
opt = ClipAdamW(cu, lr=1e-3, weight_decay=1e-2, max_norm=1.0, backend=be, world=world)
for s in range(steps):
    opt.zero_grad()
    for p, gr in zip(cu, world_grads[rank][s]):
        p.grad.copy_(gr)
    norm = opt.step()
    assert abs(float(norm) - float(ref_norms[s])) <= 1e-5 * float(ref_norms[s]), (s, float(norm), float(ref_norms[s]))
for p, r in zip(cu, ref_p):
    err = float((p.detach().cpu() - r).abs().max() / r.abs().max())
    assert err < 2e-6, err
# every rank holds bit-identical parameters
flat = opt.p.clone()
other = flat.clone()
dist.broadcast(other, 0)
assert torch.equal(flat, other)

# timing: the reference's gradient size
n = 38_436_836
gbuf = torch.randn((n + 3) // 4 * 4, device="cuda")
sq = torch.zeros(1, device="cuda")
for _ in range(3):
    be.allreduce_grads(gbuf, 1.0 / world, sq)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
iters = 10
e0.record()
for _ in range(iters):
    be.allreduce_grads(gbuf, 1.0 / world, sq)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    bus = 2 * (world - 1) / world * gbuf.numel() * 4 / (float(ms) * 1e-3) / 1e9
    print(f"dp exchange ok: world={world} allreduce+scale+norm of {gbuf.numel() * 4 / 1e6:.1f} MB: {float(ms):.3f} ms (bus {bus:.0f} GB/s)", flush=True)
be.comm_destroy()
dist.destroy_process_group()
