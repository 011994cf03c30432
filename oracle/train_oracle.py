"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's train-step tail, used as the checker for
csrc/train_step.cuh.  Only tests/ may import this.

* `loss_forward` restates train_utils/losses.py:35-79 (Loss.forward) in plain torch; pinned bit for bit (values and
  gradients) against the unmodified reference module by tests/test_train_oracle_vs_reference.py where /root/reference exists.
* `clip_adamw_steps` runs the reference's own calls — torch.nn.utils.clip_grad_norm_ (train.py:191) and
  torch.optim.AdamW(lr, weight_decay, amsgrad=True) (train.py:102-104,193) — on CPU copies; torch is the third-party
  arithmetic the reference pins (requirements.txt:10), so parity is against torch 2.11 CPU fp32 semantics
  ("parity unpinned" in SURVEY.md §8c's sense).
"""
import torch
import torch.nn.functional as F


def loss_forward(model_output, targets):
    """losses.py:35-79.  model_output = [mel, mel_post, gate_logits [B,M,1], face, attn, content_dis [rows,501], ...];
    targets = (mel_target [B,80,M], gate_target [B,M]).  Returns the dict of the four loss terms."""
    mel_target, gate_target = targets[0], targets[1]
    gate_target = gate_target.view(-1, 1)                                   # 43
    mel_out, mel_post, gate_out = model_output[0], model_output[1], model_output[2].view(-1, 1)   # 45-49
    qy = model_output[5]                                                    # 69
    log_ratio = torch.log(qy * qy.shape[-1] + 1e-20)                        # 71
    return {
        "KLD": torch.sum(qy * log_ratio, dim=-1).mean(),                    # 72-73
        "mel_loss": F.mse_loss(mel_out, mel_target),                        # 75
        "postnet_mel_loss": 10 * F.mse_loss(mel_post, mel_target),          # 76
        "gate_loss": F.binary_cross_entropy_with_logits(gate_out, gate_target),   # 77
    }


def clip_adamw_steps(params, grads_per_step, lr=1e-4, weight_decay=1e-6, max_norm=1.0, world_grads=None):
    """params: list of CPU tensors (copied).  grads_per_step: list (steps) of lists (params) of gradient tensors — when
    `world_grads` ranks are given instead ([rank][step][param]) their mean is the gradient (data-parallel semantics).
    Returns (final params, list of pre-clip gradient norms)."""
    ps = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt = torch.optim.AdamW(ps, lr=lr, weight_decay=weight_decay, amsgrad=True)
    norms = []
    steps = len(grads_per_step) if world_grads is None else len(world_grads[0])
    for s in range(steps):
        for i, p in enumerate(ps):
            if world_grads is None:
                p.grad = grads_per_step[s][i].detach().clone()
            else:
                p.grad = torch.stack([wg[s][i] for wg in world_grads]).sum(0) / len(world_grads)
        norms.append(torch.nn.utils.clip_grad_norm_(ps, max_norm))
        opt.step()
    return [p.detach() for p in ps], norms
