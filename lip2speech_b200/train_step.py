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

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, max_norm=1.0, backend=None, world=1):
        self.params = [p for p in params]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        dev = self.params[0].device
        self.be = backend or _lib.backend(dev.index or 0)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm, self.world = lr, betas, eps, weight_decay, max_norm, world
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
        return self.sqnorm.sqrt()


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
    return bytes(buf.tolist())
