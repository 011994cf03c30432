"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors
produced by the unmodified reference.  Tolerance: max|a-b|/max|b| <= 1e-3 (north_star, fp32)."""
import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import spec, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def be(weights):
    from lip2speech_b200 import _lib, build
    build.build()
    b = _lib.backend(0)
    b.bind_state_dict(weights, "", _lib.PART_VIDEO | _lib.PART_SPEAKER | _lib.PART_DECODER)
    return b


@pytest.fixture(scope="module")
def O():
    from oracle import l2s_oracle
    return l2s_oracle


def test_speaker_encoder(be, O, golden, spk_weights):
    wav = synth.wav(2)
    raw = be.speaker_fwd(wav.cuda(), normalize=False).cpu()
    assert rel_err(raw, golden["A_spk_raw"]) < TOL
    assert rel_err(raw, O.speaker_forward(spk_weights, wav)) < TOL
    emb = be.speaker_fwd(wav.cuda(), normalize=True).cpu()
    assert rel_err(emb, golden["A_spk_emb"]) < TOL
    raw = be.speaker_fwd(synth.wav(1, 48000, seed=9).cuda(), normalize=False).cpu()
    assert rel_err(raw, golden["F_spk_raw"]) < TOL


def test_video_frontend(be, golden):
    f = be.video_fwd(synth.video(2, 29).cuda()).cpu()
    assert f.shape == (2, 29, 768)
    assert rel_err(f, golden["A_video_feat"]) < TOL
    f = be.video_fwd(synth.video(1, 5, 88, 88, seed=5).cuda()).cpu()
    assert rel_err(f, golden["D_video_feat"]) < TOL


def test_bf16_stem(be, golden):
    """precision = bf16 (BASELINE configs[2]: 'bf16 3D frontend + fp32 decoder step'): the Conv3d stem multiplies bf16
    operands on tcgen05; features and the full-span mel stay inside the parity bound."""
    from lip2speech_b200 import _lib
    f = be.video_fwd(synth.video(2, 29).cuda(), precision=_lib.PRECISION_BF16).cpu()
    assert rel_err(f, golden["A_video_feat"]) < TOL
    f = be.video_fwd(synth.video(1, 5, 88, 88, seed=5).cuda(), precision=_lib.PRECISION_BF16).cpu()
    assert rel_err(f, golden["D_video_feat"]) < TOL
    mel, lengths = be.infer(synth.video(2, 29).cuda(), synth.wav(2).cuda(), synth.gumbel(2, 29).cuda(), precision=_lib.PRECISION_BF16)
    assert torch.equal(lengths.cpu(), golden["A_lengths"])
    assert rel_err(mel.cpu(), golden["A_mel"]) < TOL
    with pytest.raises(RuntimeError):
        be.video_fwd(synth.video(1, 5).cuda(), precision=7)


def test_postnet(be, golden):
    y = be.postnet_fwd(synth.mel_like(2, 77).cuda()).cpu()
    assert rel_err(y, golden["E_postnet"]) < TOL


def test_decoder_preloop_tensors(be, O, weights):
    visual, face = synth.visual_features(3, 29)
    g = synth.gumbel(3, 29)
    be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=2)
    pre = O.decoder_preloop(weights, visual, face[:, 0], g)
    assert rel_err(be.debug_read("dec.rnn_out", (3, 29, 1024)), O._lstm(visual, pre_h0(O, weights, face), pre_h0(O, weights, face).clone(), weights, "decoder.encoder_rnn", 1, True)[0]) < TOL
    assert rel_err(be.debug_read("dec.enc", (3, 29, 512)), pre["enc"]) < TOL
    assert rel_err(be.debug_read("dec.enc_cell", (3, 512)), pre["enc_cell"]) < TOL
    assert rel_err(be.debug_read("dec.K", (3, 29, 512)), pre["k"].permute(0, 2, 1)) < TOL
    assert rel_err(be.debug_read("dec.V", (3, 29, 512)), pre["v"]) < TOL
    assert rel_err(be.debug_read("dec.ckey", (3, 4, 256)), pre["ckey"].permute(0, 2, 1)) < TOL
    assert rel_err(be.debug_read("dec.cval", (3, 4, 256)), pre["cval"]) < TOL


def pre_h0(O, weights, face):
    p = "decoder."
    s = O.psine(torch.nn.functional.linear(face[:, 0], weights[p + "encoder_site.0.linear_layer.weight"],
                                           weights[p + "encoder_site.0.linear_layer.bias"]), weights[p + "encoder_site.1.w"])
    return s.unsqueeze(0).repeat(2, 1, 1)


def test_decoder_few_steps(be, O, weights):
    visual, face = synth.visual_features(3, 29)
    g = synth.gumbel(3, 29)
    for steps in (1, 2, 5):
        mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
        pre = O.decoder_preloop(weights, visual, face[:, 0], g)
        outs, lens, _ = O.decoder_steps(weights, pre, steps)
        assert rel_err(be.debug_read("dec.outputs", (3, steps, 80)), outs) < TOL, steps
        assert torch.equal(lengths.cpu(), lens)


def test_decoder_inference_golden(be, golden):
    visual, face = synth.visual_features(3, 29)
    mel, lengths, attn = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), synth.gumbel(3, 29).cuda(), return_attention=True)
    assert mel.shape == (3, 80, 300) and attn.shape == (3, 300, 29)
    assert torch.equal(lengths.cpu(), golden["B_lengths"])
    assert rel_err(mel.cpu(), golden["B_mel"]) < TOL
    _attn_agree(attn.cpu(), golden["B_attn"])
    visual, face = synth.visual_features(1, 75, seed=77)
    mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), synth.gumbel(1, 75, seed=77).cuda())
    assert torch.equal(lengths.cpu(), golden["C_lengths"])
    assert rel_err(mel.cpu(), golden["C_mel"]) < TOL


def test_full_span_golden(be, golden):
    mel, lengths = be.infer(synth.video(2, 29).cuda(), synth.wav(2).cuda(), synth.gumbel(2, 29).cuda())
    assert torch.equal(lengths.cpu(), golden["A_lengths"])
    assert rel_err(mel.cpu(), golden["A_mel"]) < TOL


def test_host_entry_point_matches_device_path(be, golden):
    """l2s_infer_host (pinned host buffers, clip copy overlapped with the speaker encoder on a second stream) returns
    exactly what the device-pointer entry point returns."""
    video, wav, g = synth.video(2, 29), synth.wav(2), synth.gumbel(2, 29)
    mel_h = torch.empty(2, 80, 300).pin_memory()
    len_h = torch.empty(2, dtype=torch.int64).pin_memory()
    for _ in range(2):                                   # second call reuses the streams / workspaces
        mel_h.zero_()
        be.infer_host(video.pin_memory(), wav.pin_memory(), g.pin_memory(), mel_h, len_h)
        mel, lengths = be.infer(video.cuda(), wav.cuda(), g.cuda())
        assert torch.equal(mel_h, mel.cpu()) and torch.equal(len_h, lengths.cpu())
    assert rel_err(mel_h, golden["A_mel"]) < TOL
    # pipelined form: two different batches in flight in the two staging slots
    v2, w2, g2 = synth.video(3, 29, seed=8), synth.wav(3, seed=8), synth.gumbel(3, 29, seed=8)
    pins = [t.pin_memory() for t in (video, wav, g, v2, w2, g2)]
    out = [(torch.empty(2, 80, 300).pin_memory(), torch.empty(2, dtype=torch.int64).pin_memory()),
           (torch.empty(3, 80, 300).pin_memory(), torch.empty(3, dtype=torch.int64).pin_memory())]
    be.infer_host_submit(0, pins[0], pins[1], pins[2], *out[0])
    be.infer_host_submit(1, pins[3], pins[4], pins[5], *out[1])
    be.infer_host_submit(0, pins[0], pins[1], pins[2], *out[0])          # slot 0 again while slot 1 is still in flight
    be.infer_host_wait(1); be.infer_host_wait(0)
    assert torch.equal(out[0][0], mel.cpu())
    m2, l2 = be.infer(v2.cuda(), w2.cuda(), g2.cuda())
    assert torch.equal(out[1][0], m2.cpu()) and torch.equal(out[1][1], l2.cpu())
    with pytest.raises(RuntimeError):
        be.infer_host_wait(2)


def _attn_agree(attn, ref_attn):
    """Attention maps: the learned temperature sqrt(512) makes the softmax one-hot sharp, so values are compared at 2e-2
    absolute AND the attended position (argmax per step) must be identical wherever the oracle's own top-2 margin is not a
    numerical tie (> 1e-3)."""
    assert (attn - ref_attn).abs().max() < 2e-2
    top2 = ref_attn.topk(2, dim=-1).values
    decided = (top2[..., 0] - top2[..., 1]) > 1e-3
    assert torch.equal(attn.argmax(-1)[decided], ref_attn.argmax(-1)[decided])
    assert decided.float().mean() > 0.9


def test_full_span_b32_bf16_vs_oracle(be, O, weights, spk_weights):
    """The bench configuration itself (BASELINE configs[2]: B=32, T=29, 96x96, S=19456, 300 steps, bf16 stem) against the
    oracle: mel within 1e-3, identical lengths, identical attended position at every step."""
    from lip2speech_b200 import _lib
    B = 32
    video, wav, g = synth.video(B, 29), synth.wav(B), synth.gumbel(B, 29)
    mel, lengths = be.infer(video.cuda(), wav.cuda(), g.cuda(), 300, _lib.PRECISION_BF16)
    assert be.debug_flag("dec3") == 1
    emb = O.speaker_inference(spk_weights, wav)
    ref_mel, ref_len, ref_attn = O.lip2speech_inference(weights, video, emb, g, 300, return_attention=True)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    # the same span in pieces, to reach the attention maps (l2s_infer does not return them)
    feat = be.video_fwd(video.cuda(), precision=_lib.PRECISION_BF16)
    spk = be.speaker_fwd(wav.cuda(), normalize=True)
    visual = torch.cat([feat, spk.unsqueeze(1).expand(-1, 29, -1)], dim=2)
    mel2, len2, attn = be.decoder_infer(visual, spk, g.cuda(), return_attention=True)
    assert torch.equal(mel2, mel) and torch.equal(len2, lengths)
    _attn_agree(attn.cpu(), ref_attn)


def test_t75_b32_all_300_steps_vs_oracle(be, O, weights):
    """AVSpeech-shape encoder length (BASELINE configs[4]: T=75, minT=10; keys/values streamed from L2) at a full
    32-clip launch for all 300 steps."""
    visual, face = synth.visual_features(32, 75, seed=75)
    g = synth.gumbel(32, 75, seed=75)
    mel, lengths, attn = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), return_attention=True)
    ref_mel, ref_len, ref_attn = O.decoder_inference(weights, visual, face, g, return_attention=True)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    _attn_agree(attn.cpu(), ref_attn)


def test_soft_attention_regime(O, weights):
    """Second weight set: another seed and SMALL temperatures (1.0 instead of sqrt(512) / sqrt(256)), so both softmaxes are
    genuinely soft (many positions carry weight) instead of one-hot — the regime a trained checkpoint may sit in."""
    from lip2speech_b200 import _lib
    w = spec.seeded_state_dict(spec.full_spec(), 4321)
    w["decoder.temperature"] = torch.full_like(w["decoder.temperature"], 1.0)
    w["decoder.content.temperature"] = torch.full_like(w["decoder.content.temperature"], 1.0)
    visual, face = synth.visual_features(5, 29, seed=61)
    g = synth.gumbel(5, 29, seed=61)
    b2 = _lib.Backend(0)
    b2.bind_state_dict(w, "", _lib.PART_DECODER)
    mel, lengths, attn = b2.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=80, return_attention=True)
    ref_mel, ref_len, ref_attn = O.decoder_inference(w, visual, face, g, steps=80, return_attention=True)
    b2.close()
    assert float((ref_attn.max(-1).values < 0.9).float().mean()) > 0.2, "temperature 1.0 should give soft attention on many steps"
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    assert (attn.cpu() - ref_attn).abs().max() < 1e-3


def test_uint8_frames_match_normalised_float_path(be):
    """Raw uint8 frames [B,T,H,W,3] (loadframes, datasets/lrw/dataset.py:20-24) with /255 + Normalize fused on the device give
    exactly what the dataset's host-side normalisation + the fp32 NCDHW entry point give (same fp32 arithmetic)."""
    from lip2speech_b200 import _lib
    frames = synth.frames_u8(3, 29)
    video = synth.normalise_frames(frames)
    for prec in (_lib.PRECISION_FP32, _lib.PRECISION_BF16):
        f_u8 = be.video_fwd_u8(frames.cuda(), precision=prec)
        f_f32 = be.video_fwd(video.cuda(), precision=prec)
        assert torch.equal(f_u8, f_f32)
    wav, g = synth.wav(3), synth.gumbel(3, 29)
    mel, lengths = be.infer_u8(frames.cuda(), wav.cuda(), g.cuda(), 40, _lib.PRECISION_BF16)
    mel_f, len_f = be.infer(video.cuda(), wav.cuda(), g.cuda(), 40, _lib.PRECISION_BF16)
    assert torch.equal(mel, mel_f) and torch.equal(lengths, len_f)
    mel_h = torch.empty(3, 80, 40).pin_memory()
    len_h = torch.empty(3, dtype=torch.int64).pin_memory()
    be.infer_host_submit_u8(0, frames.pin_memory(), wav.pin_memory(), g.pin_memory(), mel_h, len_h, 40, _lib.PRECISION_BF16)
    be.infer_host_wait(0)
    assert torch.equal(mel_h, mel.cpu()) and torch.equal(len_h, lengths.cpu())


def test_batch_70_chunked_decode_vs_oracle(be, O, weights):
    """B > 32: the decode loop runs in consecutive 32-clip launches of the pipelined kernel (32 + 32 + 6 here, partial last
    chunk); the frontend runs in 32-clip chunks as well."""
    from lip2speech_b200 import _lib
    visual, face = synth.visual_features(70, 29, seed=23)
    g = synth.gumbel(70, 29, seed=23)
    mel, lengths, attn = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=30, return_attention=True)
    assert be.debug_flag("dec3") == 1
    ref_mel, ref_len, ref_attn = O.decoder_inference(weights, visual, face, g, steps=30, return_attention=True)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    _attn_agree(attn.cpu(), ref_attn)
    # a clip's result does not depend on the chunk it lands in (the pre-loop GEMMs pick their tiling from the row count, so
    # the comparison across batch sizes is to rounding, not bit-exact)
    m2, l2 = be.decoder_infer(visual[32:64].cuda(), face[32:64, 0].cuda(), g[4 * 32:4 * 64].cuda(), steps=30)
    assert rel_err(m2.cpu(), mel[32:64].cpu()) < 1e-4 and torch.equal(l2, lengths[32:64])
    video = synth.video(35, 5, 88, 88, seed=5)
    f = be.video_fwd(video.cuda(), precision=_lib.PRECISION_BF16)
    f1 = be.video_fwd(video[32:].cuda(), precision=_lib.PRECISION_BF16)
    assert torch.equal(f[32:], f1)


def test_batch_33_two_clip_groups_vs_oracle(be, O, weights):
    """B=33 exercises the second 32-clip chunk and batch padding."""
    visual, face = synth.visual_features(33, 29, seed=21)
    g = synth.gumbel(33, 29, seed=21)
    mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=40)
    ref_mel, ref_len = O.decoder_inference(weights, visual, face, g, steps=40)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL


def test_batch_invariance_full_size(be):
    """Size-independent property at the bench size (B=32, 300 steps): a clip's mel does not depend on which other
    clips share the batch.  The stage-pipelined kernel serves every B <= 32 and every clip is an independent MMA column, so
    sub-batches of any size — aligned to its 8-clip groups or not, down to a single clip — must agree (bit-exact for the
    group-sized ones; the parity tolerance is asserted for the 4-clip and single-clip cases)."""
    visual, face = synth.visual_features(32, 29, seed=5)
    g = synth.gumbel(32, 29, seed=5)
    mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
    for lo, n in ((0, 12), (7, 9), (15, 17)):            # sub-batches served by the pipelined kernel, unaligned to clip groups
        m, l = be.decoder_infer(visual[lo:lo + n].cuda(), face[lo:lo + n, 0].cuda(), g[4 * lo:4 * (lo + n)].cuda())
        assert torch.equal(m, mel[lo:lo + n]) and torch.equal(l, lengths[lo:lo + n])
    for lo in (0, 28):                                   # 4-clip sub-batches (one partial clip group)
        m4, l4 = be.decoder_infer(visual[lo:lo + 4].cuda(), face[lo:lo + 4, 0].cuda(), g[4 * lo:4 * lo + 16].cuda())
        assert rel_err(m4.cpu(), mel[lo:lo + 4].cpu()) < TOL and torch.equal(l4, lengths[lo:lo + 4])
    for i in (0, 17, 31):                                # single clips
        m1, l1 = be.decoder_infer(visual[i:i + 1].cuda(), face[i:i + 1, 0].cuda(), g[4 * i:4 * i + 4].cuda())
        assert rel_err(m1[0].cpu(), mel[i].cpu()) < TOL and int(l1[0]) == int(lengths[i])


def test_pipelined_kernel_vs_oracle(be, O, weights):
    """B <= 32 runs the stage-pipelined kernel (decode3.cuh): partial clip groups (B=11), attention maps, and the
    Decoder.forward flavour (teacher-forced steps, raw stop logits, pre-softmax attention logits) against the oracle."""
    visual, face = synth.visual_features(11, 29, seed=31)
    g = synth.gumbel(11, 29, seed=31)
    mel, lengths, attn = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=50, return_attention=True)
    assert be.debug_flag("dec3") == 1
    ref_mel, ref_len, ref_attn = O.decoder_inference(weights, visual, face, g, steps=50, return_attention=True)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    _attn_agree(attn.cpu(), ref_attn)
    mels = synth.mel_like(11, 20, seed=31)
    tf_mask = torch.tensor([i % 3 == 0 for i in range(20)])
    o = be.decoder_forward(visual.cuda(), face[:, 0].cuda(), g.cuda(), mels.cuda(), tf_mask)
    ref = O.decoder_forward(weights, visual, face, mels, tf_mask, g)
    for got, want in zip(o[:4], (ref[0], ref[1], ref[2], ref[4])):
        assert rel_err(got.cpu(), want) < TOL
    # T = 75: keys/values do not fit in shared memory twice, the attention CTAs stream them from L2
    visual, face = synth.visual_features(9, 75, seed=32)
    g = synth.gumbel(9, 75, seed=32)
    mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=12)
    ref_mel, ref_len = O.decoder_inference(weights, visual, face, g, steps=12)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL


def _stop_logit_trajectory(O, w, pre, steps):
    """Stop logits [B, steps] of the oracle run (captured from its stop_token_layer call)."""
    logs, orig = [], O._lin

    def lin(sd, name, x):
        y = orig(sd, name, x)
        if name.endswith("stop_token_layer.linear_layer"):
            logs.append(y[:, 0].clone())
        return y

    O._lin = lin
    try:
        O.decoder_steps(w, pre, steps)
    finally:
        O._lin = orig
    return torch.stack(logs, 1)


def test_stop_token_midway(O, weights):
    """Shift the stop bias into the range of the per-step stop logits (at the middle of the widest gap between
    observed logits, so fp32 noise cannot flip a comparison) so that gates fire mid-sequence, and check
    output_lengths against the oracle (first-crossing rule, decoder.py:429-435)."""
    from lip2speech_b200 import _lib
    w = dict(weights)
    steps = 60
    visual, face = synth.visual_features(4, 29, seed=8)
    g = synth.gumbel(4, 29, seed=8)
    # with the seeded weights the stop logit peaks at step 0; negating the stop weights makes it rise later instead
    wkey = "decoder.stop_token_layer.linear_layer.weight"
    w[wkey] = -weights[wkey]
    pre = O.decoder_preloop(w, visual, face[:, 0], g)
    logits = _stop_logit_trajectory(O, w, pre, steps)
    vals = torch.sort(logits.flatten()).values
    gaps = vals[1:] - vals[:-1]
    key = "decoder.stop_token_layer.linear_layer.bias"
    chosen, best = None, (-1, 0.0)
    for idx in range(len(gaps)):
        if float(gaps[idx]) <= 2e-5:
            continue
        thr = float((vals[idx] + vals[idx + 1]) / 2)
        first = [(logits[b] > thr).nonzero() for b in range(4)]
        lens = [int(f[0]) + 1 if len(f) else steps for f in first]
        score = (sum(1 for l in lens if 2 < l < steps), float(gaps[idx]))
        if score[0] > 0 and score > best:
            best, chosen = score, (thr, lens)
    assert chosen is not None, "no robust mid-sequence stop threshold in the logit trajectory"
    thr, lens = chosen
    w[key] = weights[key] - thr
    _, ref_lens, _ = O.decoder_steps(w, pre, steps)
    assert ref_lens.tolist() == lens
    b2 = _lib.Backend(0)
    b2.bind_state_dict(w, "", _lib.PART_DECODER)
    _, lengths = b2.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
    assert lengths.cpu().tolist() == lens
    b2.close()


@pytest.mark.parametrize("B,T,steps", [(9, 31, 5), (9, 32, 5), (10, 7, 4), (32, 30, 3), (1, 29, 6)])
def test_pipelined_kernel_shape_edges(be, O, weights, B, T, steps):
    """Edges of the pipelined decode kernel: T = 31 is the largest encoder length whose K/V images fit twice in shared
    memory (T = 32 streams them from L2), T = 7 the smallest the Content convolutions accept (minT = 1), a full batch at
    an even T, and a single clip (one clip group, three SM groups idle every turn)."""
    visual, face = synth.visual_features(B, T, seed=40 + T)
    g = synth.gumbel(B, T, seed=40 + T)
    mel, lengths, attn = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps, return_attention=True)
    assert be.debug_flag("dec3") == 1
    ref_mel, ref_len, ref_attn = O.decoder_inference(weights, visual, face, g, steps=steps, return_attention=True)
    assert torch.equal(lengths.cpu(), ref_len)
    assert rel_err(mel.cpu(), ref_mel) < TOL
    assert (attn.cpu() - ref_attn).abs().max() < 2e-2


def test_module_mirror_matches_backend(be, weights, golden):
    """The nn.Module mirror (reference-shaped API) drives the same kernels."""
    from lip2speech_b200 import modules
    sd = dict(weights)
    spk_sd = {k[len("speaker_encoder."):]: sd.pop(k) for k in list(sd) if k.startswith("speaker_encoder.")}
    net = modules.get_network("test")
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    spk = modules.SpeakerEncoder(state_dict=spk_sd).eval().cuda()
    emb = spk.inference(synth.wav(2).cuda())
    mel, lengths, attn = net.inference(synth.video(2, 29).cuda(), None, emb, return_attention_map=True,
                                       gumbel_noise=synth.gumbel(2, 29).cuda())
    assert rel_err(mel.cpu(), golden["A_mel"]) < TOL
    assert attn.shape == (2, 300, 29)
    # seeded internal gumbel draw is reproducible
    torch.manual_seed(7); m1, _ = net.inference(synth.video(2, 29).cuda(), None, emb)
    torch.manual_seed(7); m2, _ = net.inference(synth.video(2, 29).cuda(), None, emb)
    assert torch.equal(m1, m2)


def test_decoder_forward_eval(be, O, golden, weights):
    """Decoder.forward in eval mode (evaluate.py path) against the reference goldens and the oracle, with and without
    teacher-forced steps; the mirror module reproduces the reference's seeded coin flips."""
    from lip2speech_b200 import modules
    visual, face = synth.visual_features(2, 29, seed=11)
    g = synth.gumbel(2, 29, seed=11)
    mels = synth.mel_like(2, 24, seed=11)
    dec = modules.Decoder()
    dec.load_state_dict({k[len("decoder."):]: v for k, v in weights.items() if k.startswith("decoder.")}, strict=True)
    dec = dec.cuda().eval()
    lens = torch.full((2,), 29, dtype=torch.long)
    for name, tf in (("G5", 0.5), ("G1", 1.0)):
        torch.manual_seed(4321)
        o = dec(visual.cuda(), face.cuda(), mels.cuda(), lens, lens, tf, gumbel_noise=g.cuda())
        assert rel_err(o[0].cpu(), golden[name + "_outputs"]) < TOL
        assert rel_err(o[1].cpu(), golden[name + "_post"]) < TOL
        assert rel_err(o[2].cpu(), golden[name + "_stop"]) < TOL
        assert rel_err(o[4].cpu(), golden[name + "_attn_logits"]) < TOL
        assert rel_err(o[5].cpu(), golden[name + "_cdis"]) < TOL
        assert torch.equal(o[3].cpu(), face[:, 0])
    dec.train()                                   # train mode: the CUDA train path (tests/test_train_model_gpu.py)
    o = dec(visual.cuda(), face.cuda(), mels.cuda(), lens, lens, 0.5)
    assert o[1].requires_grad and o[1].shape == (2, 80, 24)


def test_unsupported_shapes_are_errors(be):
    visual, face = synth.visual_features(1, 29)
    with pytest.raises(RuntimeError):
        be.decoder_infer(visual.cuda(), face[:, 0].cuda(), synth.gumbel(1, 29).cuda(), steps=301)
    with pytest.raises(RuntimeError):
        be.video_fwd(torch.zeros(1, 3, 2, 64, 64).cuda())
