"""GPU parity of the train-step tail (csrc/train_step.cuh through the C ABI) against the reference's own calls on CPU:
Loss.forward + its gradients (train_utils/losses.py:35-79), clip_grad_norm_ + AdamW(amsgrad) (train.py:102-104,191-193),
and the NCCL gradient exchange at world size 2 (runs only on a box with >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def be():
    from lip2speech_b200 import _lib, build
    build.build()
    return _lib.backend(0)


def _loss_inputs(B, M, rows, seed):
    g = torch.Generator().manual_seed(seed)
    mel_t = (torch.randn(B, 80, M, generator=g) * 2 - 5).clamp_min(-11.5129)      # SURVEY §8d mel targets
    mel = mel_t + 0.3 * torch.randn(B, 80, M, generator=g)
    post = mel_t + 0.2 * torch.randn(B, 80, M, generator=g)
    gate = torch.randn(B, M, 1, generator=g) * 3
    gate_t = torch.zeros(B, M); gate_t[:, M - 3:] = 1.0
    dis = torch.softmax(torch.randn(rows, 501, generator=g) * 2, -1)
    return mel, post, gate, dis, mel_t, gate_t


@pytest.mark.parametrize("B,M,rows", [(2, 24, 8), (8, 77, 32), (3, 5, 12)])
def test_loss_and_gradients(be, B, M, rows):
    from oracle import train_oracle as TO
    mel, post, gate, dis, mel_t, gate_t = _loss_inputs(B, M, rows, 5 + B)
    ref_in = [t.clone().requires_grad_(True) for t in (mel, post, gate, dis)]
    ref = TO.loss_forward([ref_in[0], ref_in[1], ref_in[2], None, None, ref_in[3]], (mel_t, gate_t))
    sum(ref.values()).backward()
    losses, grads = be.loss_fwd_bwd(mel.cuda(), post.cuda(), gate.cuda(), dis.cuda(), mel_t.cuda(), gate_t.cuda())
    got = losses.cpu()
    for j, k in enumerate(("KLD", "mel_loss", "postnet_mel_loss", "gate_loss")):
        assert abs(float(got[j]) - float(ref[k].detach())) <= 1e-5 * max(1.0, abs(float(ref[k].detach()))), k
    for g, r in zip(grads, ref_in):
        assert rel_err(g.cpu().view_as(r.grad), r.grad) < 1e-5


def test_loss_module_autograd(be):
    """The nn.Module mirror of train_utils/losses.py:Loss: same keys, gradients flow through the C-ABI call."""
    from lip2speech_b200.train_step import Loss
    from oracle import train_oracle as TO
    mel, post, gate, dis, mel_t, gate_t = _loss_inputs(4, 30, 16, 9)
    xs = [t.cuda().requires_grad_(True) for t in (mel, post, gate, dis)]
    out = Loss()([xs[0], xs[1], xs[2], None, None, xs[3]], (mel_t.cuda(), gate_t.cuda()))
    assert set(out) == {"KLD", "mel_loss", "postnet_mel_loss", "gate_loss"}
    (out["mel_loss"] + 2.0 * out["gate_loss"] + out["postnet_mel_loss"] + out["KLD"]).backward()
    rs = [t.clone().requires_grad_(True) for t in (mel, post, gate, dis)]
    ref = TO.loss_forward([rs[0], rs[1], rs[2], None, None, rs[3]], (mel_t, gate_t))
    (ref["mel_loss"] + 2.0 * ref["gate_loss"] + ref["postnet_mel_loss"] + ref["KLD"]).backward()
    for x, r in zip(xs, rs):
        assert rel_err(x.grad.cpu(), r.grad) < 1e-5


def _param_set(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(7,), (33, 5), (1024, 257), (3,), (512, 1, 5), (1,)]
    return [torch.randn(*s, generator=g) * 0.1 for s in shapes]


@pytest.mark.parametrize("max_norm,grad_scale", [(1.0, 5.0), (1.0, 1e-4), (0.0, 1.0)])
def test_clip_adamw_matches_torch(be, max_norm, grad_scale):
    """Five steps of clip_grad_norm_ + AdamW(amsgrad) on a ragged parameter set (sizes not multiples of 4): clipping
    active, clipping inactive (tiny gradients), clipping disabled."""
    from lip2speech_b200.train_step import ClipAdamW
    from oracle import train_oracle as TO
    params = _param_set(3)
    g = torch.Generator().manual_seed(11)
    steps = 5
    grads = [[torch.randn(p.shape, generator=g) * grad_scale for p in params] for _ in range(steps)]
    if max_norm > 0:
        ref_p, ref_norms = TO.clip_adamw_steps(params, grads, lr=1e-3, weight_decay=1e-2, max_norm=max_norm)
    else:                                                     # no clipping: plain AdamW
        ps = [torch.nn.Parameter(p.clone()) for p in params]
        opt = torch.optim.AdamW(ps, lr=1e-3, weight_decay=1e-2, amsgrad=True)
        for s in range(steps):
            for p, gr in zip(ps, grads[s]):
                p.grad = gr.clone()
            opt.step()
        ref_p, ref_norms = [p.detach() for p in ps], None
    cu = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = ClipAdamW(cu, lr=1e-3, weight_decay=1e-2, max_norm=max_norm, backend=be)
    for s in range(steps):
        opt.zero_grad()
        for p, gr in zip(cu, grads[s]):
            p.grad.copy_(gr)                                 # p.grad is a view of the flat gradient buffer
        norm = opt.step()
        if ref_norms is not None:
            assert abs(float(norm) - float(ref_norms[s])) <= 1e-5 * float(ref_norms[s])
    for p, r in zip(cu, ref_p):
        assert rel_err(p.detach().cpu(), r) < 2e-6
    # the padding between parameters never moves
    assert float(opt.p[opt.offsets[0] + 7:opt.offsets[1]].abs().max()) == 0.0


def test_clip_adamw_state_dict_round_trip(be):
    """the reference checkpoints optim.state_dict() (train.py:211) and restores it on resume (train.py:125).
    ClipAdamW's state_dict has torch.optim.AdamW's layout and values; loading it (from ours or from torch's) continues the
    run exactly."""
    from lip2speech_b200.train_step import ClipAdamW
    params = _param_set(4)
    g = torch.Generator().manual_seed(12)
    grads = [[torch.randn(p.shape, generator=g) for p in params] for _ in range(6)]
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    ref = torch.optim.AdamW([{"params": ps[:2]}, {"params": ps[2:]}], lr=1e-3, weight_decay=1e-2, amsgrad=True)   # two groups, as train.py:102
    cu = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = ClipAdamW([{"params": cu[:2]}, {"params": cu[2:]}], lr=1e-3, weight_decay=1e-2, max_norm=0.0, backend=be)
    assert opt.state_dict()["state"] == {}
    for s in range(3):
        for p, gr in zip(ps, grads[s]):
            p.grad = gr.clone()
        ref.step()
        opt.zero_grad()
        for p, gr in zip(cu, grads[s]):
            p.grad.copy_(gr)
        opt.step()
    sd, rsd = opt.state_dict(), ref.state_dict()
    assert [g_["params"] for g_ in sd["param_groups"]] == [g_["params"] for g_ in rsd["param_groups"]]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert sd["param_groups"][0][k] == rsd["param_groups"][0][k]
    assert set(sd["state"]) == set(rsd["state"])
    for i in rsd["state"]:
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"])
        for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            assert rel_err(sd["state"][i][k].cpu(), rsd["state"][i][k]) < 2e-6, (i, k)
    # resume in a fresh optimizer from TORCH's state dict and continue: same parameters as the uninterrupted torch run
    cu2 = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ps]
    opt2 = ClipAdamW([{"params": cu2[:2]}, {"params": cu2[2:]}], lr=5e-2, max_norm=0.0, backend=be)
    opt2.load_state_dict(rsd)
    assert opt2.t == 3 and opt2.lr == 1e-3
    for s in range(3, 6):
        for p, gr in zip(ps, grads[s]):
            p.grad = gr.clone()
        ref.step()
        opt2.zero_grad()
        for p, gr in zip(cu2, grads[s]):
            p.grad.copy_(gr)
        opt2.step()
    for p, r in zip(cu2, ps):
        assert rel_err(p.detach().cpu(), r.detach()) < 2e-6


def test_forward_after_optimizer_step_sees_new_weights(be):
    """ClipAdamW.step() writes the parameters through raw pointers (no data_ptr / version change); the next
    forward must run on the UPDATED weights, not on the copies packed before the step."""
    from lip2speech_b200 import modules, spec, synth
    from lip2speech_b200.train_step import ClipAdamW
    from oracle import l2s_oracle as O
    dec = modules.Decoder(seed=1234).cuda().eval()
    visual, face = synth.visual_features(2, 29, seed=3)
    g = synth.gumbel(2, 29, seed=3)
    with torch.no_grad():
        mel0, _ = dec.inference(visual.cuda(), face.cuda(), gumbel_noise=g.cuda())
    opt = ClipAdamW(dec.parameters(), lr=1e-2, max_norm=0.0)
    opt.zero_grad()
    gen = torch.Generator().manual_seed(5)
    for p in dec.parameters():
        p.grad.copy_(torch.randn(p.shape, generator=gen))
    opt.step()
    with torch.no_grad():
        mel1, len1 = dec.inference(visual.cuda(), face.cuda(), gumbel_noise=g.cuda())
    w = {"decoder." + k: v.detach().cpu() for k, v in dec.state_dict().items()}
    ref_mel, ref_len = O.decoder_inference(w, visual, face, g)
    assert rel_err(mel1.cpu(), ref_mel) < 1e-3 and torch.equal(len1.cpu(), ref_len)
    assert rel_err(mel0.cpu(), ref_mel) > 1e-2, "the step should have moved the output"


def test_gradient_exchange_single_rank_nccl(be):
    """The NCCL path of l2s_allreduce_grads on a 1-GPU box: a world-size-1 communicator (ncclCommInitRank through the
    library's dlopen'ed NCCL) leaves the gradient unchanged; scale and squared norm are applied on the device."""
    import ctypes as C
    from lip2speech_b200 import _lib
    lib = _lib.load()
    raw = C.create_string_buffer(_lib.NCCL_UNIQUE_ID_BYTES)
    assert lib.l2s_nccl_unique_id(raw, _lib.NCCL_UNIQUE_ID_BYTES) == 0
    b2 = _lib.Backend(0)
    b2.comm_init(raw.raw, 0, 1)
    g = torch.randn(1_000_003 // 4 * 4, generator=torch.Generator().manual_seed(2)).cuda()
    ref = g.clone()
    sq = torch.zeros(1, device="cuda")
    b2.allreduce_grads(g, 0.5, sq)
    assert torch.equal(g, ref * 0.5)
    assert abs(float(sq) - float((ref.double() * 0.5).pow(2).sum())) < 1e-4 * float(sq)
    b2.comm_destroy()
    b2.close()


def test_train_step_argument_errors(be):
    x = torch.zeros(16, device="cuda")
    with pytest.raises(RuntimeError):
        be.clip_adamw_step(x, x, x, x, x, None, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)      # clipping needs the norm
    with pytest.raises(RuntimeError):
        be.clip_adamw_step(x, x, x, x, x, x[:1], 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0)     # step starts at 1
    with pytest.raises(RuntimeError):
        be.allreduce_grads(x[1:], 1.0, x[:1])                                           # misaligned flat buffer


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_gradient_exchange_two_ranks():
    """world size 2 over NCCL: flat_grads <- mean over ranks, identical updated parameters on both ranks, equal to the
    single-process reference fed with the mean gradient."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "dp_exchange_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dp exchange ok" in r.stdout
