"""GPU parity of the steps right after the path: the vocoder (MelSpec2Audio: exp -> InverseMelScale ->
GriffinLim, datasets/spectograms.py:76-95) against torchaudio on CPU with the SAME initial phase, and batched ESTOI against the
fp64 numpy restatement of pystoi (oracle/audio_metrics.py), through the C ABI (l2s_vocoder, l2s_estoi)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from lip2speech_b200 import _lib, audio, build
    build.build()
    b = _lib.backend(0)
    b.bind_vocoder(audio.inverse_mel_operator())
    return b


def _log_mel(B, L, seed):
    """A plausible log-mel: the reference's own analysis (spectograms.py:42-61: n_fft 1024, hop 256, 80 mels, log(clamp(.,1e-5)))
    of seeded coloured noise."""
    import torchaudio.transforms as T
    g = torch.Generator().manual_seed(seed)
    wav = torch.randn(B, (L - 1) * 256, generator=g).cumsum(-1)
    wav = 0.3 * wav / wav.abs().amax(-1, keepdim=True)
    m = T.MelSpectrogram(16000, n_fft=1024, win_length=1024, hop_length=256, f_min=0.0, f_max=8000.0, n_mels=80, power=2.0)(wav)
    return torch.log(m.clamp_min(1e-5)), wav


def test_inverse_mel_and_one_istft(be):
    """n_iter = 0: waveform = istft(sqrt(relu(InverseMelScale(exp(mel)))) * initial phase) — the GEMM-based inverse DFT,
    window folding and overlap-add against torch.istft."""
    import torchaudio.transforms as T
    mel, _ = _log_mel(3, 40, 1)
    init = torch.rand(3, 513, 40, generator=torch.Generator().manual_seed(2), dtype=torch.complex64)
    spec = T.InverseMelScale(n_stft=513, n_mels=80, sample_rate=16000, f_min=0.0, f_max=8000.0)(torch.exp(mel))
    mag = spec.clamp_min(0).sqrt()
    ref = torch.istft(mag * init, 1024, 256, 1024, torch.hann_window(1024), center=True, length=None)
    got = be.vocoder(mel.cuda(), init, n_iter=0).cpu()
    assert got.shape == ref.shape == (3, 39 * 256)
    # the least-squares spectrum has many entries near zero and sqrt() has an unbounded derivative there: a 3e-7 difference in
    # the solve (torchaudio's fp32 gelsd per call vs one fp32 operator applied as a GEMM) becomes ~5e-4 in the magnitude
    # (measured with an fp64 emulation of this formulation against torchaudio: 2.2e-4 on the waveform)
    assert rel_err(got, ref) < 3e-3


@pytest.mark.parametrize("n_iter", [1, 8])
def test_griffin_lim_matches_torchaudio(be, n_iter):
    """The Griffin-Lim iteration (istft -> stft -> momentum phase update) against torchaudio.functional.griffinlim with the
    same random initial phase (it is the first draw after the seed)."""
    import torchaudio.functional as AF
    import torchaudio.transforms as T
    mel, _ = _log_mel(2, 48, 3)
    spec = T.InverseMelScale(n_stft=513, n_mels=80, sample_rate=16000, f_min=0.0, f_max=8000.0)(torch.exp(mel)).clamp_min(0)
    torch.manual_seed(11)
    ref = AF.griffinlim(spec, torch.hann_window(1024), 1024, 256, 1024, 2.0, n_iter, 0.99, None, True)
    torch.manual_seed(11)
    init = torch.rand(spec.size(), dtype=torch.complex64)
    got = be.vocoder(mel.cuda(), init, n_iter=n_iter, momentum=0.99).cpu()
    # phase retrieval amplifies rounding differences iteration by iteration: tight after one, looser after eight
    assert rel_err(got, ref) < (5e-3 if n_iter == 1 else 5e-2), rel_err(got, ref)


def test_estoi_matches_numpy_restatement(be):
    from oracle import audio_metrics as AM
    g = np.random.default_rng(5)
    B, S = 6, 19456
    t = np.arange(S) / 16000.0
    clean = np.stack([np.sin(2 * np.pi * (200 + 90 * b) * t) * (0.2 + 0.8 * (np.sin(2 * np.pi * (2 + b) * t) > 0)) + 0.01 * g.standard_normal(S)
                      for b in range(B)]).astype(np.float32)
    clean[1, :6000] *= 1e-4                                    # a long silent stretch: frames are dropped
    proc = (clean + np.stack([(0.02 + 0.1 * b) * g.standard_normal(S) for b in range(B)])).astype(np.float32)
    proc[3] = g.standard_normal(S).astype(np.float32)         # unrelated noise
    got = be.estoi(torch.from_numpy(clean).cuda(), torch.from_numpy(proc).cuda()).cpu().numpy()
    ref = np.array([AM.estoi(clean[b], proc[b]) for b in range(B)])
    assert np.abs(got - ref).max() < 1e-6, (got, ref)
    assert got[0] > got[5] and got[3] < 0.3                   # more noise -> lower score


def test_sample_lrw_estoi_on_gpu(be, weights, spk_weights):
    """evaluate.py:38-45 on the device for the 10 SAMPLE_LRW clips: mel -> vocoder -> ESTOI, all on the GPU, against the host
    chain (torchaudio vocoder + numpy ESTOI) fed with the SAME mel and the same initial phase: |dESTOI| <= 1e-3."""
    import os
    from oracle import audio_metrics as AM
    from tests_helpers import load_sample_lrw
    from lip2speech_b200 import _lib
    video, audio = load_sample_lrw()
    be.bind_state_dict(weights, "", 7)
    g = synth.gumbel(10, 29, seed=2024)
    mel, _ = be.infer(video.cuda(), audio.cuda(), g.cuda())
    L = 77
    m = mel[:, :, :L].contiguous()
    torch.manual_seed(7)
    init = torch.rand(10, 513, L, dtype=torch.complex64)
    wav_gpu = be.vocoder(m, init, n_iter=32)
    e_gpu = be.estoi(audio[:, : wav_gpu.shape[1]].cuda(), wav_gpu).cpu().numpy()
    wav_ref = AM.mel_to_audio(m.cpu(), seed=7, n_iter=32).numpy()
    e_ref = np.array([AM.estoi(audio[i].numpy(), wav_ref[i]) for i in range(10)])
    assert np.abs(e_gpu - e_ref).max() <= 1e-3, (e_gpu, e_ref)
