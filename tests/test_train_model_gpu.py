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
# BatchNorm biases followed by a bias-free 1x1 conv and another train-mode BatchNorm: the shift is removed again, true gradient 0
ZERO_GRAD_BN_BIAS = re.compile(r"(banch2\.4|banch1\.1)\.bias$")


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


# (18, ...): more than 16 rows per step — the per-step layers leave the few-row kernels for the general GEMM path; (12, 75, ...): AVSpeech-length clips
@pytest.mark.parametrize("B,T,M,tf_ratio,soft,seed", [(3, 29, 10, 0.5, True, 7), (2, 29, 8, 0.3, False, 11), (4, 31, 6, 0.9, True, 13),
                                                     (18, 29, 4, 0.5, True, 17), (12, 75, 5, 0.5, True, 19)])
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
    # soft-attention cases agree to ~1e-4.  With the seeded sqrt(512) temperature the attention is one-hot-sharp and the
    # gradients that pass through it (the K path, Q, the temperatures) are ill-conditioned in fp32: the oracle itself deviates
    # from the unmodified reference by up to 4e-2 there (tests/test_train_oracle_vs_reference.py), so the bound is looser
    for k, e in errs.items():
        assert e < (2e-3 if soft else (5e-2 if k.endswith("temperature") else 1e-2)), (k, e)
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
    for k, v in first.items():          # second pass accumulated on top of the first (sums over steps re-associate: not bit-exact 2x)
        assert torch.allclose(grads[k], 2 * v, rtol=1e-4, atol=1e-6 * float(v.abs().max())), k
    with pytest.raises(RuntimeError):
        be.decoder_train_bwd(*g_out, B, T)           # the tape was consumed


def test_decoder_train_bench_shape_vs_oracle(be):
    """The shape bench.py --config c3 times (8 clips per GPU, T=29, M=77 teacher frames; BASELINE configs[3]) against autograd
    on the oracle: outputs, input gradients and every parameter gradient — on the eager first call and again on the
    CUDA-graph replay the timed steps actually run (third call with the key)."""
    from oracle import train_oracle as TO
    B, T, M = 8, 29, 77
    w = _decoder_weights(1234, True)
    visual, face = synth.visual_features(B, T, seed=41)
    mels = synth.mel_like(B, M, seed=41) * 2 - 5
    gate_t = torch.zeros(B, M); gate_t[:, -2:] = 1
    noise = TO.reference_noise(B, T, M, 0.5, with_video=False, generator=torch.Generator().manual_seed(41))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not spec.is_buffer(k) else v.clone()) for k, v in w.items()}
    vin = visual.clone().requires_grad_(True)
    fin = face.clone().requires_grad_(True)
    ref = TO.decoder_forward_train(sd, vin, fin, mels, noise)
    sum(TO.loss_forward(ref, (mels, gate_t)).values()).backward()

    be.train_set_graphs(True)
    dev, grads = _bind(be, w)
    for attempt in ("eager", "captured", "replayed"):
        for k, v in w.items():                       # undo the in-place running-statistics update of the previous pass
            if spec.is_buffer(k) and v.is_floating_point():
                dev[k].copy_(v)
        for v in grads.values():
            if v is not None:
                v.zero_()
        out = be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise.to("cuda"))
        for got, want, name in zip(out, (ref[0], ref[1], ref[2], ref[4], ref[5]), ("outputs", "post", "stop", "attn_logits", "content_dis")):
            assert rel_err(got.cpu(), want.detach()) < 1e-3, (attempt, name)
        _, g = be.loss_fwd_bwd(out[0], out[1], out[2], out[4], mels.cuda(), gate_t.cuda())
        g_visual, g_spk = be.decoder_train_bwd(g[0], g[1], g[2], g[3], B, T)
        torch.cuda.synchronize()
        assert rel_err(g_visual.cpu(), vin.grad) < 2e-3, attempt
        assert rel_err(g_spk.cpu(), fin.grad[:, 0]) < 2e-3, attempt
        for k, p in sd.items():
            if not (torch.is_tensor(p) and p.requires_grad):
                continue
            got = grads[k].cpu()
            if ZERO_GRAD_BIAS.search(k):
                assert float(got.abs().max()) < 1e-3 * float(sd[k[:-4] + "weight"].grad.abs().max()) + 1e-7, (attempt, k)
            else:
                assert rel_err(got, p.grad) < 2e-3, (attempt, k, rel_err(got, p.grad))


def test_graph_replay_is_bit_identical_to_eager(be):
    """The train-mode forward / backward are captured as CUDA graphs on the second call with a key and replayed afterwards
    (l2s_train_set_graphs).  Eager, captured and replayed passes must agree bit for bit — also when the teacher-forcing coins
    and every other noise tensor change between replays (the coins are read on the device, not baked into the graph), and when
    the caller's input tensors live at new addresses."""
    from oracle import train_oracle as TO
    B, T, M = 2, 29, 6
    w = _decoder_weights(1234, True)
    dev, grads = _bind(be, w)
    g_out = [torch.randn(B, 80, M, device="cuda"), torch.randn(B, 80, M, device="cuda"), torch.randn(B, M, 1, device="cuda"), None]

    def one_pass(seed):
        visual, face = synth.visual_features(B, T, seed=seed)
        mels = synth.mel_like(B, M, seed=seed) * 2 - 5
        noise = TO.reference_noise(B, T, M, 0.5, with_video=False, generator=torch.Generator().manual_seed(seed)).to("cuda")
        for v in grads.values():
            if v is not None:
                v.zero_()
        n0 = be.launch_count()
        out = be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise)
        gv, gs = be.decoder_train_bwd(*g_out, B, T)
        torch.cuda.synchronize()
        return ([o.clone() for o in out], gv.clone(), gs.clone(), {k: v.clone() for k, v in grads.items() if v is not None}, be.launch_count() - n0,
                noise.tf_mask.clone())

    def same(a, b):
        assert all(torch.equal(x, y) for x, y in zip(a[0], b[0])), "outputs"
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), "input gradients"
        for k in a[3]:
            assert torch.equal(a[3][k], b[3][k]), k
        assert a[4] == b[4], "launch accounting"

    be.train_set_graphs(False)
    eager = {seed: one_pass(seed) for seed in (5, 6, 7)}
    assert not all(torch.equal(eager[5][5], eager[s][5]) for s in (6, 7)), "the seeds must draw different teacher-forcing coins"
    be.train_set_graphs(True)
    try:
        same(one_pass(5), eager[5])          # captured + launched (the eager passes above already showed the key once)
        same(one_pass(5), eager[5])          # replayed
        same(one_pass(6), eager[6])          # replay, new inputs / noise / coins
        same(one_pass(7), eager[7])
        same(one_pass(5), eager[5])
    finally:
        be.train_set_graphs(True)


def test_decoder_train_without_input_gradients(be):
    """want_input_grads = 0 (the C ABI's switch for a decoder trained on frozen features): the parameter gradients are the
    same bits, no input-gradient buffers are touched — eager and replayed."""
    from oracle import train_oracle as TO
    B, T, M = 3, 29, 5
    w = _decoder_weights(1234, True)
    dev, grads = _bind(be, w)
    visual, face = synth.visual_features(B, T, seed=23)
    mels = synth.mel_like(B, M, seed=23) * 2 - 5
    noise = TO.reference_noise(B, T, M, 0.5, with_video=False, generator=torch.Generator().manual_seed(23)).to("cuda")
    g_out = [torch.randn(B, 80, M, device="cuda"), torch.randn(B, 80, M, device="cuda"), torch.randn(B, M, 1, device="cuda"), None]

    def one_pass(want):
        for v in grads.values():
            if v is not None:
                v.zero_()
        be.decoder_train_fwd(visual.cuda(), face[:, 0].cuda(), mels.cuda(), noise, want_input_grads=want)
        gv, gs = be.decoder_train_bwd(*g_out, B, T, want_input_grads=want)
        torch.cuda.synchronize()
        return gv, gs, {k: v.clone() for k, v in grads.items() if v is not None}

    gv, gs, ref = one_pass(True)
    assert gv is not None and gs is not None
    for attempt in range(3):                     # eager, captured, replayed
        gv0, gs0, got = one_pass(False)
        assert gv0 is None and gs0 is None
        for k in ref:
            assert torch.equal(got[k], ref[k]), (attempt, k)


def test_video_graph_replay_is_bit_identical_to_eager(be):
    B, T = 2, 5
    w = {k: v for k, v in spec.seeded_state_dict(spec.encoder_spec("encoder."), 1234).items()}
    dev, grads = _bind(be, w)

    def one_pass(seed):
        video = synth.video(B, T, seed=seed).cuda()
        drop = torch.empty(B, T, 768, device="cuda").bernoulli_(0.9, generator=torch.Generator("cuda").manual_seed(seed))
        g_feat = torch.randn(B, T, 768, device="cuda", generator=torch.Generator("cuda").manual_seed(seed + 100))
        for k, v in w.items():               # the forward updates the running statistics in place: restore them
            if spec.is_buffer(k) and v.is_floating_point():
                dev[k].copy_(v)
        for v in grads.values():
            if v is not None:
                v.zero_()
        feat = be.video_train_fwd(video, drop)
        be.video_train_bwd(g_feat)
        torch.cuda.synchronize()
        return feat.clone(), {k: v.clone() for k, v in grads.items() if v is not None}, {k: dev[k].clone() for k in dev if spec.is_buffer(k)}

    def same(a, b):
        assert torch.equal(a[0], b[0])
        for k in a[1]:
            assert torch.equal(a[1][k], b[1][k]), k
        for k in a[2]:
            assert torch.equal(a[2][k], b[2][k]), k

    be.train_set_graphs(False)
    eager = {seed: one_pass(seed) for seed in (1, 2)}
    be.train_set_graphs(True)
    same(one_pass(1), eager[1])
    same(one_pass(1), eager[1])
    same(one_pass(2), eager[2])
    same(one_pass(1), eager[1])


def _cast_sd(w, dt):
    return {k: (v.clone().to(dt).requires_grad_(True) if v.is_floating_point() and not spec.is_buffer(k)
                else (v.clone().to(dt) if v.is_floating_point() else v.clone())) for k, v in w.items()}


def _l2_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _check_against_fp64(got, yardsticks, g64, skip, floor=1e-2, slack=4.0):
    """Deep train-mode BatchNorm + ReLU stacks on a small batch are ill-conditioned: a 1e-6 relative difference in the forward
    pass (fp32 rounding in another summation order) flips individual ReLU decisions, and ONE flipped element moves single
    entries of some gradients by percents (measured: the fp64 oracle fed a clip perturbed by 1e-6 deviates from itself by
    4.3e-2 max-abs on trunk.1.0.weight — the very number the CUDA path shows; which element flips depends on the perturbation).
    So every gradient is compared with the fp64 oracle in relative L2 norm (a flip is sparse, a wrong kernel is not) and must
    be within `floor`, or as close as the yardstick runs are (torch fp32; fp64 with 2e-6 input noise) x slack."""
    worst = []
    for k, truth in g64.items():
        if skip(k):
            continue
        e_gpu = _l2_err(got[k], truth)
        e_ref = max(_l2_err(y[k], truth) for y in yardsticks)
        worst.append((e_gpu / max(floor, slack * e_ref), k, e_gpu, e_ref))
    worst.sort(reverse=True)
    print("worst (ratio to allowance, key, cuda-vs-fp64 rel-L2, yardstick-vs-fp64 rel-L2):", worst[:5])
    assert worst[0][0] <= 1.0, worst[:5]


def test_video_train_forward_backward_vs_oracle(be):
    """VideoExtractor.forward in train mode (Conv3d stem + BatchNorm batch statistics + PReLU + MaxPool3d + 16 ShuffleNetV2
    blocks + conv_last + AvgPool + L2 norm) + the dropout of model.py:26: features and the gradient of every encoder
    parameter against autograd on the oracle (fp64 truth, fp32 as the noise yardstick), plus the BatchNorm running statistics."""
    from oracle import l2s_oracle as O, train_oracle as TO
    B, T, H = 2, 5, 88
    w = spec.seeded_state_dict(spec.encoder_spec("encoder."), 1234)
    video = synth.video(B, T, H, H, seed=9)
    keep = torch.empty(B, T, 768).bernoulli_(0.9, generator=torch.Generator().manual_seed(4))
    g_out = torch.randn(B, T, 768, generator=torch.Generator().manual_seed(5))

    jitter = torch.randn(video.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)

    def oracle(dt, noise=0.0):
        sd = _cast_sd(w, dt)
        bn_new = {}
        v = video.to(dt) * (1 + noise * jitter.to(dt)) if noise else video.to(dt)
        with TO.bn_train(bn_new):
            ref = TO._drop(O.video_features(sd, v, "encoder."), keep.to(dt), 0.1)
        (ref * g_out.to(dt)).sum().backward()
        return ref.detach(), {k: p.grad for k, p in sd.items() if torch.is_tensor(p) and p.requires_grad}, bn_new

    ref32, g32, bn_new = oracle(torch.float32)
    ref64, g64, _ = oracle(torch.float64)
    _, gnz, _ = oracle(torch.float64, 2e-6)
    dev, grads = _bind(be, w)
    feat = be.video_train_fwd(video.cuda(), keep.cuda())
    assert rel_err(feat.cpu(), ref32) < 1e-3
    be.video_train_bwd(g_out.cuda())
    torch.cuda.synchronize()
    gmax = max(float(g.abs().max()) for g in g64.values())
    for k in g64:
        if ZERO_GRAD_BN_BIAS.search(k):
            assert float(grads[k].abs().max()) < 1e-4 * gmax, k
    _check_against_fp64({k: v.cpu() for k, v in grads.items() if v is not None}, (g32, gnz), g64, lambda k: bool(ZERO_GRAD_BN_BIAS.search(k)))
    assert len(bn_new) == 2 * 56
    for k, v in bn_new.items():
        assert rel_err(dev[k].cpu(), v) < 1e-4, k


def test_full_train_step_through_mirror_modules(be):
    """train.py:167-193 with the mirror modules: net.train(); net(...) -> Loss -> loss.backward() -> ClipAdamW.step().
    Gradients reach every encoder.* and decoder.* parameter through the two autograd nodes and match autograd on the oracle;
    after the optimizer step the (eval) forward runs on the UPDATED weights."""
    from lip2speech_b200 import modules
    from lip2speech_b200.train_step import ClipAdamW, Loss
    from oracle import train_oracle as TO
    B, T, M, H = 2, 7, 6, 88
    w = spec.seeded_state_dict(spec.full_spec(), 1234)
    w["decoder.temperature"] = torch.full_like(w["decoder.temperature"], 1.0)
    w["decoder.content.temperature"] = torch.full_like(w["decoder.content.temperature"], 1.0)
    net_sd = {k: v for k, v in w.items() if not k.startswith("speaker_encoder.")}
    net = modules.get_network("train")
    net.load_state_dict(net_sd, strict=True)
    net = net.cuda()
    video = synth.video(B, T, H, H, seed=2)
    spk = synth.speaker_embedding(B, seed=2)
    mels = synth.mel_like(B, M, seed=2) * 2 - 5
    gate_t = torch.zeros(B, M); gate_t[:, -2:] = 1
    noise = TO.reference_noise(B, T, M, 0.5, with_video=True, generator=torch.Generator().manual_seed(8))

    jitter = torch.randn(video.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)

    def oracle(dt, nz=0.0):
        sd = _cast_sd(net_sd, dt)
        n = TO.TrainNoise(noise.video_drop.to(dt), noise.gumbel.to(dt), noise.tf_mask, noise.prenet.to(dt), noise.attn.to(dt), noise.lstm.to(dt),
                          [t.to(dt) for t in noise.post])
        v = video.to(dt) * (1 + nz * jitter.to(dt)) if nz else video.to(dt)
        ref = TO.lip2speech_forward_train(sd, v, spk.to(dt), mels.to(dt), n)
        losses = TO.loss_forward(ref, (mels.to(dt), gate_t.to(dt)))
        sum(losses.values()).backward()
        return float(sum(losses.values())), {k: p.grad for k, p in sd.items() if torch.is_tensor(p) and p.requires_grad}

    loss32, g32 = oracle(torch.float32)
    loss64, g64 = oracle(torch.float64)
    _, gnz = oracle(torch.float64, 2e-6)
    # CUDA path through the reference-shaped API
    opt = ClipAdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-3, max_norm=1.0)   # train.py:102-104
    lens = torch.full((B,), T, dtype=torch.long)
    tn = modules.TrainNoise(noise.tf_mask, noise.gumbel.cuda(), noise.prenet.cuda(), noise.attn.cuda(), noise.lstm.cuda(),
                            [t.cuda() for t in noise.post], video_drop=noise.video_drop.cuda())
    opt.zero_grad()
    out = net(video.cuda(), None, None, mels.cuda(), lens, None, lens, 0.5, speaker_embedding=spk.cuda(), train_noise=tn)
    losses = Loss()(out, (mels.cuda(), gate_t.cuda()))
    loss = sum(losses.values())
    assert abs(float(loss) - loss64) < 1e-3 * abs(loss64)
    loss.backward()
    got = {}
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        got[k] = p.grad.cpu()
    _check_against_fp64(got, (g32, gnz), g64, lambda k: bool(ZERO_GRAD_BIAS.search(k) or ZERO_GRAD_BN_BIAS.search(k)))
    before = {k: p.detach().clone() for k, p in net.named_parameters()}
    opt.step()
    still = [k for k, p in net.named_parameters() if torch.equal(before[k], p.detach())]
    # T = 7 gives ONE content slot: the softmax over it is constant, so the content keys / query / temperature have an exactly
    # zero gradient (and a 1e-9 weight decay does not move an fp32 value); every other parameter moved
    assert all(re.search(r"decoder\.content\.(K\.|Q\.|temperature)", k) for k in still), still
    # eval forward after the step: packed weights were invalidated, BatchNorm running statistics moved
    net.eval()
    with torch.no_grad():
        mel, lengths = net.inference(video.cuda(), None, spk.cuda(), gumbel_noise=noise.gumbel.cuda())
    assert torch.isfinite(mel).all()
    assert int(net.decoder.state_dict()["postnet.convolutions.0.1.num_batches_tracked"]) == 1
