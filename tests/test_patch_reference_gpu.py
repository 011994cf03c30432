"""The reference's OWN classes (model/model.py, model/modules/*.py, unmodified) running on the B200 backend after
`lip2speech_b200.patch.patch()` — the "demo.py / evaluate.py / train.py run unchanged" claim of north_star, executed on
hardware.  Needs the reference tree: /root/reference in the build container, or the copy staged by
oracle/stage_reference.sh under oracle/_ref/reference (git-ignored test data that travels with the gpurun snapshot);
skipped where neither exists (the driver's box)."""
import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import spec, synth
from oracle import ref_import

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.available(), reason="needs the reference tree (oracle/stage_reference.sh)")]


@pytest.fixture(scope="module")
def ref_net(weights):
    import lip2speech_b200.patch as b200
    from lip2speech_b200 import build
    build.build()
    Decoder, VideoExtractor, SpeakerEncoder, Lip2Speech = ref_import.import_reference()
    import model.model as ref_model                       # the reference's module (model/model.py)
    import model.modules.decoder as ref_decoder
    ref_model.device = ref_decoder.device = torch.device("cuda")        # the reference's module-level `device` globals (model.py:10, decoder.py:16)
    b200.patch()
    net = ref_model.get_network("test")                   # model.py:62-72, unmodified
    sd = {k: v for k, v in weights.items() if not k.startswith("speaker_encoder.")}
    missing, unexpected = net.load_state_dict(sd, strict=False)         # vgg_face.* (third-party, out of scope) is not in the seeded set
    assert not unexpected and all(k.startswith("vgg_face.") for k in missing)
    net = net.cuda()
    spk_sd = {k[len("speaker_encoder."):]: v for k, v in weights.items() if k.startswith("speaker_encoder.")}
    spk = SpeakerEncoder(state_dict=spk_sd).eval().cuda() if _speaker_ctor_takes_state(SpeakerEncoder) else None
    yield net, spk
    b200.unpatch()


def _speaker_ctor_takes_state(cls):
    import inspect
    return "state_dict" in inspect.signature(cls.__init__).parameters


def test_reference_inference_on_b200_backend(ref_net, golden):
    """demo.py:84-86 with the reference's classes: speaker_encoder.inference(audio) -> net.inference(video, faces, emb)."""
    net, spk = ref_net
    video, wav, g = synth.video(2, 29), synth.wav(2), synth.gumbel(2, 29)
    if spk is not None:
        emb = spk.inference(wav.cuda())
        assert rel_err(emb.cpu(), golden["A_spk_emb"]) < 1e-3
    else:
        emb = golden["A_spk_emb"].cuda()
    mel, lengths, attn = net.inference(video.cuda(), None, emb, return_attention_map=True, gumbel_noise=g.cuda())
    assert mel.shape == (2, 80, 300) and attn.shape == (2, 300, 29)
    assert torch.equal(lengths.cpu(), golden["A_lengths"])
    assert rel_err(mel.cpu(), golden["A_mel"]) < 1e-3


def test_reference_train_step_on_b200_backend(ref_net):
    """train.py:167-193 with the reference's classes in train(): net(...) -> Loss -> loss.backward() -> clip -> AdamW.step(),
    the stock torch optimizer of train.py:102-104 over the reference modules' own parameters."""
    from lip2speech_b200.train_step import Loss
    net, _ = ref_net
    B, T, M = 2, 7, 5
    video, spk = synth.video(B, T, 88, 88, seed=2), synth.speaker_embedding(B, seed=2)
    mels = synth.mel_like(B, M, seed=2) * 2 - 5
    gate_t = torch.zeros(B, M); gate_t[:, -2:] = 1
    net.train()
    net.vgg_face.inference = lambda faces: spk.cuda()     # FaceRecognizer = third-party InceptionResnetV1 (out of scope, frozen)
    optim = torch.optim.AdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-4, weight_decay=1e-6, amsgrad=True)
    lens = torch.full((B,), T, dtype=torch.long)
    torch.manual_seed(3)
    out = net(video.cuda(), torch.zeros(B, 2, 3, 8, 8).cuda(), None, mels.cuda(), lens, None, lens, 0.5)     # model.py:23
    losses = Loss()(out, (mels.cuda(), gate_t.cuda()))
    loss = sum(losses.values())
    optim.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters() if not k.startswith("vgg_face.")}
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    assert sum(int(float(g.abs().max()) > 0) for g in grads.values()) > 0.9 * len(grads)
    before = net.decoder.fc_out.linear_layer.weight.detach().clone()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    optim.step()
    assert not torch.equal(before, net.decoder.fc_out.linear_layer.weight.detach())
    net.eval()
