"""Host mirror of the reference's train-step tail (train.py:102-104, 167-193; train_utils/losses.py:13-79): the
reconstruction `Loss`, and `ClipAdamW` = gradient exchange over the data-parallel ranks + clip_grad_norm_(1.0) +
AdamW(amsgrad=True).step() on ONE flat fp32 buffer per state (p, g, m, v, vmax) that the module's parameters alias.
Every arithmetic step is a C-ABI call (csrc/train_step.cuh); there is no PyTorch fallback.

The forward-train / backward kernels of the model itself are not built yet: gradients enter here as
`p.grad` views of the flat gradient buffer, whoever produced them.
"""
from __future__ import annotations

import torch

from . import _lib


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, backend, mel_out, mel_post, gate_logits, content_dis, mel_target, gate_target):
        losses, grads = backend.loss_fwd_bwd(mel_out, mel_post, gate_logits, content_dis, mel_target, gate_target)
        ctx.save_for_backward(*grads)
        ctx.shapes = (mel_out.shape, mel_post.shape, gate_logits.shape, content_dis.shape)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        # d(sum_k w_k L_k): the kernel returned d(sum L)/d(outputs) per output; each output feeds exactly one loss term
        g_mel, g_post, g_gate, g_dis = ctx.saved_tensors
        w_kld, w_mel, w_post, w_gate = grad_losses.unbind(0)
        return (None, (g_mel * w_mel).view(ctx.shapes[0]), (g_post * w_post).view(ctx.shapes[1]),
                (g_gate * w_gate).view(ctx.shapes[2]), (g_dis * w_kld).view(ctx.shapes[3]), None, None)


class Loss(torch.nn.Module):
    """train_utils/losses.py:Loss — same call signature and dict keys (KLD, mel_loss, postnet_mel_loss, gate_loss)."""

    def forward(self, model_output, targets, losses=None):
        if losses is None:
            losses = dict()
        mel_target, gate_target = targets[0], targets[1]
        mel_out, mel_post, gate_out, qy = model_output[0], model_output[1], model_output[2], model_output[5]
        be = _lib.backend(mel_out.device.index or 0)
        v = _LossFn.apply(be, mel_out, mel_post, gate_out, qy, mel_target, gate_target)
        losses['KLD'], losses['mel_loss'], losses['postnet_mel_loss'], losses['gate_loss'] = v[0], v[1], v[2], v[3]
        return losses


def flat_layout(sizes):
    """Offsets of the parameters inside the flat buffers (every parameter starts 16-byte aligned) and the total length."""
    offsets, off = [], 0
    for n in sizes:
        offsets.append(off)
        off += (n + 3) // 4 * 4
    return offsets, off


class ClipAdamW:
    """optim.zero_grad() / [backward] / clip_grad_norm_ / optim.step() of train.py:180-193 on flat buffers.

    params: iterable of CUDA fp32 parameters; their storage is MOVED into one flat buffer (param.data become views) and
    param.grad are views of a flat gradient buffer.  `step()` returns the pre-clip gradient norm as a device scalar
    (what clip_grad_norm_ returns), after: sum over ranks (NCCL, if a communicator was set up) x 1/world, global L2
    norm, clip to max_norm, AdamW(amsgrad) update."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, max_norm=1.0, backend=None, world=None):
        """`params` may be an iterable of parameters or of {'params': [...]} groups, as train.py:102-104 passes them
        (every group shares the hyper-parameters).  `world`: data-parallel size the gradients are averaged over; by default
        the size of the communicator set up with init_data_parallel (1 without one)."""
        params = list(params)
        if params and isinstance(params[0], dict):
            groups = [list(g["params"]) for g in params]          # group["params"] may be a generator (train.py:102-103)
            self._group_sizes = [len(g) for g in groups]
            params = [p for g in groups for p in g]
        else:
            self._group_sizes = [len(params)]
        self.params = params
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        dev = self.params[0].device
        self.be = backend or _lib.backend(dev.index or 0)
        comm_world = getattr(self.be, "world", 1)
        if world is not None and comm_world > 1 and world != comm_world:
            raise ValueError(f"ClipAdamW(world={world}) but the backend's communicator spans {comm_world} ranks")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.world = world if world is not None else comm_world
        self.offsets, self.n = flat_layout([p.numel() for p in self.params])
        self.p = torch.zeros(self.n, device=dev)
        self.g = torch.zeros(self.n, device=dev)
        self.m = torch.zeros(self.n, device=dev)
        self.v = torch.zeros(self.n, device=dev)
        self.vmax = torch.zeros(self.n, device=dev)
        self.sqnorm = torch.zeros(1, device=dev)
        for p, o in zip(self.params, self.offsets):
            flat = self.p[o:o + p.numel()].view_as(p)
            flat.copy_(p.data)
            p.data = flat
            p.grad = self.g[o:o + p.numel()].view_as(p)
        self.t = 0

    def zero_grad(self):
        self.g.zero_()

    def step(self):
        self.t += 1
        self.be.allreduce_grads(self.g, 1.0 / self.world, self.sqnorm)
        self.be.clip_adamw_step(self.p, self.g, self.m, self.v, self.vmax, self.sqnorm, self.max_norm, self.lr, self.betas[0],
                                self.betas[1], self.eps, self.weight_decay, self.t)
        # the kernel wrote the parameters through raw pointers: neither data_ptr nor the autograd version counter moved, so
        # tell the backend that its packed copies (BN folded, tiles repacked) are stale
        self.be.invalidate_weights()
        return self.sqnorm.sqrt()

    # ---- torch.optim.AdamW-compatible checkpoint state (train.py:125 load_state_dict, :211 'optimize_state') ---------------
    @property
    def param_groups(self):
        groups, i = [], 0
        for n in self._group_sizes:
            groups.append({"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                           "amsgrad": True, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                           "fused": None, "params": list(range(i, i + n))})
            i += n
        return groups

    def state_dict(self):
        """Same layout as torch.optim.AdamW(amsgrad=True).state_dict(): per-parameter step / exp_avg / exp_avg_sq /
        max_exp_avg_sq (copies sliced from the flat buffers; empty before the first step, like torch) + param_groups."""
        state = {}
        if self.t > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                sl = slice(o, o + p.numel())
                state[i] = {"step": torch.tensor(float(self.t)), "exp_avg": self.m[sl].view_as(p).clone(),
                            "exp_avg_sq": self.v[sl].view_as(p).clone(), "max_exp_avg_sq": self.vmax[sl].view_as(p).clone()}
        return {"state": state, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if [len(g["params"]) for g in groups] != self._group_sizes:
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        g0 = groups[0]
        if not g0.get("amsgrad", False):
            raise ValueError("ClipAdamW implements AdamW(amsgrad=True) only (train.py:104)")
        self.lr, self.betas, self.eps, self.weight_decay = g0["lr"], tuple(g0["betas"]), g0["eps"], g0["weight_decay"]
        state = sd["state"]
        steps = {int(st["step"]) for st in state.values()}
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; the flat update keeps one step counter")
        self.t = steps.pop() if steps else 0
        self.m.zero_(); self.v.zero_(); self.vmax.zero_()
        for i, st in state.items():
            p, o = self.params[int(i)], self.offsets[int(i)]
            sl = slice(o, o + p.numel())
            self.m[sl].copy_(st["exp_avg"].reshape(-1)); self.v[sl].copy_(st["exp_avg_sq"].reshape(-1))
            self.vmax[sl].copy_(st["max_exp_avg_sq"].reshape(-1))


def init_data_parallel(backend: "_lib.Backend", rank: int, world: int, group=None):
    """Create the library's NCCL communicator for this rank; the unique id travels over torch.distributed."""
    import torch.distributed as dist
    lib = _lib.load()
    buf = torch.zeros(_lib.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        import ctypes as C
        raw = C.create_string_buffer(_lib.NCCL_UNIQUE_ID_BYTES)
        if lib.l2s_nccl_unique_id(raw, _lib.NCCL_UNIQUE_ID_BYTES) != 0:
            raise RuntimeError("l2s_nccl_unique_id failed: " + lib.l2s_last_error(None).decode())
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        dbuf = buf.cuda()
        dist.broadcast(dbuf, 0, group=group)
        buf = dbuf.cpu()
    else:
        dist.broadcast(buf, 0, group=group)
    backend.comm_init(bytes(buf.tolist()), rank, world)
    backend.world = world
    return bytes(buf.tolist())
