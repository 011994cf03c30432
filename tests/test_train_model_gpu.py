"""GPU parity of the train-mode forward + backward (csrc/train_model.cuh through the C ABI: l2s_train_bind,
l2s_decoder_train_fwd / _bwd) against torch.autograd on the train-mode oracle (oracle/train_oracle.py, itself pinned against
the seeded unmodified reference Decoder by tests/test_train_oracle_vs_reference.py): outputs, the gradient of EVERY decoder
parameter, the input gradients and the BatchNorm running statistics, with identical explicit noise on both sides."""
import re

import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import spec, synth

pytestmark = pytest.mark.gpu
ZERO_GRAD_BIAS = re.compile(r"(postnet\.convolutions\.\d\.0\.conv|[KV]\.0\.conv\.\d\.0|content\.agg\.\d\.0)\.bias$")


@pytest.fixture(scope="module")
def be():
    from lip2speech_b200 import _lib, build
    build.build()
    return _lib.backend(0)


def _decoder_weights(seed, soft):
    w = spec.seeded_state_dict(spec.decoder_spec("decoder."), seed)
    if soft:                                   # soft-attention regime: well-conditioned gradients through both softmaxes
        w["decoder.temperature"] = torch.full_like(w["decoder.temperature"], 1.0)
        w["decoder.content.temperature"] = torch.full_like(w["decoder.content.temperature"], 1.0)
    return w


def _bind(be, w):
    """Device copies of the parameters + zeroed gradient buffers, bound under the reference's keys."""
    dev, grads = {}, {}
    for k, v in w.items():
        if not v.is_floating_point():
            continue                            # num_batches_tracked
        dev[k] = v.clone().cuda().contiguous()
        trainable = not spec.is_buffer(k)
        grads[k] = torch.zeros_like(dev[k]) if trainable else None
        be.train_bind(k, dev[k], grads[k])
    return dev, grads


@pytest.mark.parametrize("B,T,M,tf_ratio,soft,seed", [(3, 29, 10, 0.5, True, 7), (2, 29, 8, 0.3, False, 11), (4, 31, 6, 0.9, True, 13)])
def test_decoder_train_forward_backward_vs_oracle(be, B, T, M, tf_ratio, soft, seed):
    from oracle import train_oracle as TO
    w = _decoder_weights(1234, soft)
    visual, face = synth.visual_features(B, T, seed=seed)
    mels = synth.mel_like(B, M, seed=seed) * 2 - 5
    gate_t = torch.zeros(B, M); gate_t[:, -2:] = 1
    noise = TO.reference_noise(B, T, M, tf_ratio, with_video=False, generator=torch.Generator().manual_seed(seed))

    # ---- oracle: autograd on the functional restatement -------------------------------------------------------------------
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not spec.is_buffer(k) else v.clone()) for k, v in w.items()}
    vin = visual.clone().requires_grad_(True)
    fin = face.clone().requires_grad_(True)
    bn_new = {}
    ref = TO.decoder_forward_train(sd, vin, fin, mels, noise, bn_updates=bn_new)
    ref_losses = TO.loss_forward(ref, (mels, gate_t))
    sum(ref_losses.values()).backward()

    # ---- CUDA path -----------------------------------------------------------------------------------------------------------
    dev, grads = _bind(be, w)
    out = be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise.to("cuda"))
    for got, want, name in zip(out, (ref[0], ref[1], ref[2], ref[4], ref[5]), ("outputs", "post", "stop", "attn_logits", "content_dis")):
        assert rel_err(got.cpu(), want.detach()) < 1e-3, name
    losses, g = be.loss_fwd_bwd(out[0], out[1], out[2], out[4], mels.cuda(), gate_t.cuda())
    for j, k in enumerate(("KLD", "mel_loss", "postnet_mel_loss", "gate_loss")):
        assert abs(float(losses[j]) - float(ref_losses[k])) <= 1e-3 * max(1.0, abs(float(ref_losses[k]))), k
    g_visual, g_spk = be.decoder_train_bwd(g[0], g[1], g[2], g[3], B, T)
    torch.cuda.synchronize()

    assert rel_err(g_visual.cpu(), vin.grad) < 2e-3
    assert rel_err(g_spk.cpu(), fin.grad[:, 0]) < 2e-3        # only face_features[:, 0] is read (decoder.py:323)
    errs = {}
    for k, p in sd.items():
        if not (torch.is_tensor(p) and p.requires_grad):
            continue
        assert p.grad is not None, k
        got = grads[k].cpu()
        if ZERO_GRAD_BIAS.search(k):             # analytically zero (bias before a train-mode BatchNorm): rounding noise on both sides
            wk = k[:-4] + "weight"
            assert float(got.abs().max()) < 1e-3 * float(sd[wk].grad.abs().max()) + 1e-7, k
            continue
        errs[k] = rel_err(got, p.grad)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("worst gradient deviations:", worst)
    for k, e in errs.items():
        assert e < (5e-2 if k.endswith("temperature") and not soft else 2e-3), (k, e)
    # BatchNorm running statistics were updated in place (momentum 0.1, unbiased variance)
    assert len(bn_new) == 2 * 17
    for k, v in bn_new.items():
        assert rel_err(dev[k].cpu(), v) < 1e-4, k


def test_train_backward_is_deterministic_and_accumulates(be):
    """Two identical forward/backward passes give bit-identical gradients (no floating-point atomics), and a second
    backward into the same gradient memory ACCUMULATES (as autograd does into p.grad)."""
    from oracle import train_oracle as TO
    B, T, M = 2, 29, 5
    w = _decoder_weights(1234, True)
    visual, face = synth.visual_features(B, T, seed=3)
    mels = synth.mel_like(B, M, seed=3) * 2 - 5
    noise = TO.reference_noise(B, T, M, 0.5, with_video=False, generator=torch.Generator().manual_seed(1)).to("cuda")
    dev, grads = _bind(be, w)
    g_out = [torch.randn(B, 80, M, device="cuda"), torch.randn(B, 80, M, device="cuda"), torch.randn(B, M, 1, device="cuda"), None]
    be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise)
    gv1, _ = be.decoder_train_bwd(*g_out, B, T)
    first = {k: v.clone() for k, v in grads.items() if v is not None}
    be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise)
    gv2, _ = be.decoder_train_bwd(*g_out, B, T)
    assert torch.equal(gv1, gv2)
    for k, v in first.items():
        assert torch.allclose(grads[k], 2 * v, rtol=1e-6, atol=0), k
    with pytest.raises(RuntimeError):
        be.decoder_train_bwd(*g_out, B, T)           # the tape was consumed
