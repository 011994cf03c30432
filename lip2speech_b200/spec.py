"""State-dict specification of the Lip2Speech hot path (names/shapes only) and a
platform-independent seeded initialiser.

The reference exposes no plugin API; its boundary is `nn.Module` signatures plus
`state_dict` key names (reference model/model.py:13-59).  To be a drop-in, the
backend must accept exactly those keys.  This module *generates* the key set
programmatically from the architecture constants:

  encoder.*  : 337 keys  (reference model/modules/video.py:62-72, shufflenetv2.py:42-152)
  decoder.*  : 191 keys  (reference model/modules/decoder.py:274-318)
  speaker    : 16 keys   (reference model/modules/audio.py:110-129)

`tests/test_spec.py` checks the generated set against `tests/golden/state_manifest.json`
(dumped from the real reference modules by `tests/golden/make_golden.py`).

Seeded weights are drawn with numpy's PCG64 (bit-reproducible across hosts) so the
GPU box regenerates exactly the tensors the golden vectors were made with; no 154 MB
checkpoint needs to be committed.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

# ---- live hyper-parameters (reference hparams.py:32-38,51,57,71-73) -------------------
N_MELS = 80
MAX_DECODER_STEPS = 300
ENCODER_EMBEDDING_DIM = 1024
HID = 512                      # FFN_HID_DIM, decoder.py:287
SPK_DIM = 256
CONTENT_DIM = 256              # latent_dim, decoder.py:200
VOCAB = 501                    # Content.vocab_size, decoder.py:200
GUMBEL_TAU = 0.1               # decoder.py:257
POSTNET_DIM = 512
POSTNET_K = 5
POSTNET_N = 5
MULTIHOP_KS = (1, 3, 7, 11)    # decoder.py:164-185
CONTENT_KS = (1, 3, 5, 7)      # decoder.py:209-230 (kernel == stride)
VIDEO_FEAT = 768               # stage_out_channels[-1] = 1024-256, shufflenetv2.py:118
STEM_CH = 24
TRUNK_STAGES = ((116, 4), (232, 8), (464, 4))   # width_mult 1.0, shufflenetv2.py:113,118
SPK_MEL = 40
SPK_NFFT = 400
SPK_HOP = 160
SPK_HID = 256
SPK_LAYERS = 3

# kind tags drive the seeded initialiser
W, B, BN_W, BN_B, BN_M, BN_V, BN_N, PSINE, PRELU, CONST, RAND01, NORMAL = (
    "w", "b", "bn_w", "bn_b", "bn_m", "bn_v", "bn_n", "psine", "prelu", "const", "rand01", "normal")


def _bn(spec, name, c):
    spec[name + ".weight"] = ((c,), BN_W)
    spec[name + ".bias"] = ((c,), BN_B)
    spec[name + ".running_mean"] = ((c,), BN_M)
    spec[name + ".running_var"] = ((c,), BN_V)
    spec[name + ".num_batches_tracked"] = ((), BN_N)


def encoder_spec(prefix: str = "") -> "OrderedDict[str, tuple]":
    """VideoExtractor parameters: trunk (ShuffleNetV2 x1.0 features + conv_last) then the
    Conv3d stem, in the reference's registration order (video.py:62-72)."""
    s: OrderedDict = OrderedDict()
    cin = STEM_CH
    blk = 0
    for cout, repeats in TRUNK_STAGES:
        half = cout // 2
        for r in range(repeats):
            p = f"{prefix}trunk.0.{blk}."
            if r == 0:  # stride-2 block with two branches (shufflenetv2.py:66-89)
                s[p + "banch1.0.weight"] = ((cin, 1, 3, 3), W)
                _bn(s, p + "banch1.1", cin)
                s[p + "banch1.2.weight"] = ((half, cin, 1, 1), W)
                _bn(s, p + "banch1.3", half)
                s[p + "banch2.0.weight"] = ((half, cin, 1, 1), W)
            else:       # stride-1 split block (shufflenetv2.py:51-65)
                s[p + "banch2.0.weight"] = ((half, half, 1, 1), W)
            _bn(s, p + "banch2.1", half)
            s[p + "banch2.3.weight"] = ((half, 1, 3, 3), W)
            _bn(s, p + "banch2.4", half)
            s[p + "banch2.5.weight"] = ((half, half, 1, 1), W)
            _bn(s, p + "banch2.6", half)
            cin = cout
            blk += 1
    s[f"{prefix}trunk.1.0.weight"] = ((VIDEO_FEAT, cin, 1, 1), W)
    _bn(s, f"{prefix}trunk.1.1", VIDEO_FEAT)
    s[f"{prefix}frontend3D.0.weight"] = ((STEM_CH, 3, 5, 7, 7), W)
    _bn(s, f"{prefix}frontend3D.1", STEM_CH)
    s[f"{prefix}frontend3D.2.weight"] = ((STEM_CH,), PRELU)
    return s


def _lin(s, name, cout, cin):
    s[name + ".weight"] = ((cout, cin), W)
    s[name + ".bias"] = ((cout,), B)


def _conv1d(s, name, cout, cin, k):
    s[name + ".weight"] = ((cout, cin, k), W)
    s[name + ".bias"] = ((cout,), B)


def _lstm(s, name, inp, hid, layers, bidir=False):
    for l in range(layers):
        for sfx in ([""] + (["_reverse"] if bidir else [])):
            i = inp if l == 0 else hid * (2 if bidir else 1)
            s[f"{name}.weight_ih_l{l}{sfx}"] = ((4 * hid, i), W)
            s[f"{name}.weight_hh_l{l}{sfx}"] = ((4 * hid, hid), W)
            s[f"{name}.bias_ih_l{l}{sfx}"] = ((4 * hid,), B)
            s[f"{name}.bias_hh_l{l}{sfx}"] = ((4 * hid,), B)


def decoder_spec(prefix: str = "") -> "OrderedDict[str, tuple]":
    """Decoder parameters in registration order (decoder.py:274-318)."""
    s: OrderedDict = OrderedDict()
    p = prefix
    s[p + "BOS"] = ((1, 1, N_MELS), NORMAL)
    s[p + "temperature"] = ((1,), CONST)
    chans = [N_MELS] + [POSTNET_DIM] * (POSTNET_N - 1) + [N_MELS]
    for i in range(POSTNET_N):
        _conv1d(s, f"{p}postnet.convolutions.{i}.0.conv", chans[i + 1], chans[i], POSTNET_K)
        _bn(s, f"{p}postnet.convolutions.{i}.1", chans[i + 1])
    for i in range(POSTNET_N - 1):
        s[f"{p}postnet.sin_activation.{i}.w"] = ((POSTNET_DIM,), PSINE)
    _lin(s, p + "encoder_proj.linear_layer", HID, 2 * HID)
    for site in ("encoder_site", "attention_site"):
        _lin(s, f"{p}{site}.0.linear_layer", HID, SPK_DIM)
        s[f"{p}{site}.1.w"] = ((HID,), PSINE)
    _conv1d(s, p + "residual_bottleneck", HID, ENCODER_EMBEDDING_DIM, 1)
    _lstm(s, p + "encoder_rnn", ENCODER_EMBEDDING_DIM, HID, 1, bidir=True)
    for kv in ("K", "V"):
        for j, k in enumerate(MULTIHOP_KS):
            _conv1d(s, f"{p}{kv}.0.conv.{j}.0", HID, HID, k)
            _bn(s, f"{p}{kv}.0.conv.{j}.1", HID)
        _conv1d(s, f"{p}{kv}.0.bottleneck", HID, HID * (len(MULTIHOP_KS) + 1), 1)
        s[f"{p}{kv}.1.w"] = ((HID,), PSINE)
    _lin(s, p + "Q.0.linear_layer", HID, 2 * HID)
    s[p + "Q.1.w"] = ((HID,), PSINE)
    s[p + "content.word_embeddings"] = ((VOCAB, CONTENT_DIM), RAND01)
    s[p + "content.temperature"] = ((1,), CONST)
    for j, k in enumerate(CONTENT_KS):
        _conv1d(s, f"{p}content.agg.{j}.0", HID, HID, k)
        _bn(s, f"{p}content.agg.{j}.1", HID)
    _conv1d(s, p + "content.bottleneck", CONTENT_DIM, HID * (len(CONTENT_KS) + 1), 1)
    _lin(s, p + "content.location_fc.0", CONTENT_DIM, CONTENT_DIM)
    _lin(s, p + "content.location_fc.2", CONTENT_DIM, CONTENT_DIM)
    _lin(s, p + "content.location_fc.4", VOCAB, CONTENT_DIM)
    _lin(s, p + "content.K.0", CONTENT_DIM, CONTENT_DIM)
    _lin(s, p + "content.K.2", CONTENT_DIM, CONTENT_DIM)
    _lin(s, p + "content.Q.0", CONTENT_DIM, 2 * HID)
    _lin(s, p + "attention_proj.linear_layer", HID // 2, HID)
    _lin(s, p + "prenet.0.linear_layer", HID // 2, N_MELS)
    s[p + "prenet.1.w"] = ((HID // 2,), PSINE)
    _lin(s, p + "prenet.3.linear_layer", HID // 2, HID // 2)
    s[p + "prenet.4.w"] = ((HID // 2,), PSINE)
    _lstm(s, p + "decoder_rnn", HID, HID, 2)
    _lin(s, p + "fc_out.linear_layer", N_MELS, HID)
    _lin(s, p + "E_C.linear_layer", HID, 2 * HID)
    _lin(s, p + "stop_token_layer.linear_layer", 1, 2 * HID)
    s[p + "positional_encodings.pos_table"] = ((1, MAX_DECODER_STEPS, HID), CONST)
    return s


def speaker_spec(prefix: str = "") -> "OrderedDict[str, tuple]":
    """SpeakerEncoder parameters + torchaudio buffers (audio.py:111-129)."""
    s: OrderedDict = OrderedDict()
    _lstm(s, prefix + "lstm", SPK_MEL, SPK_HID, SPK_LAYERS)
    _lin(s, prefix + "linear", SPK_HID, SPK_HID)
    s[prefix + "mel_spec.spectrogram.window"] = ((SPK_NFFT,), CONST)
    s[prefix + "mel_spec.mel_scale.fb"] = ((SPK_NFFT // 2 + 1, SPK_MEL), CONST)
    return s


BUFFER_SUFFIXES = (".running_mean", ".running_var", ".num_batches_tracked", ".pos_table",
                   ".spectrogram.window", ".mel_scale.fb")


def is_buffer(key: str) -> bool:
    return key.endswith(BUFFER_SUFFIXES)


# ---- fixed (non-random) tensors ------------------------------------------------------

def sinusoid_table(n_position: int = MAX_DECODER_STEPS, d_hid: int = HID) -> torch.Tensor:
    """tab[p, j] = p / 10000^(2*(j//2)/d); sin on even j, cos on odd j, computed in float64
    then cast (decoder.py:20-31)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    tab = np.where(j % 2 == 0, np.sin(ang), np.cos(ang))
    return torch.from_numpy(tab.astype(np.float32)).unsqueeze(0)


def hann_periodic(n: int = SPK_NFFT) -> torch.Tensor:
    k = np.arange(n, dtype=np.float64)
    return torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * k / n)).astype(np.float32))


def htk_mel_fbanks(n_freqs: int = SPK_NFFT // 2 + 1, n_mels: int = SPK_MEL,
                   f_min: float = 0.0, f_max: float = 8000.0, sample_rate: int = 16000) -> torch.Tensor:
    """Triangular HTK mel filterbank, norm=None — what torchaudio's MelScale builds for
    MelSpectrogram(16000, n_fft=400, n_mels=40) (audio.py:124)."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    hz2mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    m_pts = np.linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2)
    f_pts = mel2hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return torch.from_numpy(fb.astype(np.float32))


_CONSTS = {
    "temperature": lambda shape: torch.full(shape, float(HID) ** 0.5),
    "content.temperature": lambda shape: torch.full(shape, float(CONTENT_DIM) ** 0.5),
    "positional_encodings.pos_table": lambda shape: sinusoid_table(),
    "mel_spec.spectrogram.window": lambda shape: hann_periodic(),
    "mel_spec.mel_scale.fb": lambda shape: htk_mel_fbanks(),
}


def _const_for(key: str, shape):
    for sfx in sorted(_CONSTS, key=len, reverse=True):
        if key == sfx or key.endswith("." + sfx):
            return _CONSTS[sfx](shape)
    raise KeyError(key)


def _uniform(rng: np.random.Generator, shape, lo, hi) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    a = rng.random(n, dtype=np.float64) * (hi - lo) + lo
    return torch.from_numpy(a.astype(np.float32).reshape(shape))


def seeded_tensor(key: str, shape, kind: str, seed: int) -> torch.Tensor:
    """One tensor, a pure function of (key, seed): PCG64 stream keyed by crc32(key)."""
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
    if kind == W:
        fan_in = int(np.prod(shape[1:]))
        a = 1.0 / math.sqrt(fan_in)
        if ".weight_hh_" in key or ".weight_ih_" in key:      # nn.LSTM default: U(-1/sqrt(H), 1/sqrt(H))
            a = 1.0 / math.sqrt(shape[0] // 4)
        elif len(shape) >= 4 or (len(shape) == 3 and "postnet" not in key):
            a = math.sqrt(3.0 / fan_in)                         # conv: unit-gain variance
        elif len(shape) == 2 or len(shape) == 3:
            a = math.sqrt(6.0 / (fan_in + shape[0] * int(np.prod(shape[2:]))))  # xavier-uniform
        return _uniform(rng, shape, -a, a)
    if kind == B:
        return _uniform(rng, shape, -0.05, 0.05)
    if kind == BN_W:
        return _uniform(rng, shape, 0.5, 1.5)
    if kind == BN_B:
        return _uniform(rng, shape, -0.2, 0.2)
    if kind == BN_M:
        return _uniform(rng, shape, -0.2, 0.2)
    if kind == BN_V:
        return _uniform(rng, shape, 0.5, 1.5)
    if kind == BN_N:
        return torch.zeros((), dtype=torch.int64)
    if kind == PSINE:
        return _uniform(rng, shape, 0.5, 1.5)
    if kind == PRELU:
        return _uniform(rng, shape, 0.1, 0.4)
    if kind == RAND01:
        return _uniform(rng, shape, 0.0, 1.0)
    if kind == NORMAL:
        return _uniform(rng, shape, -1.7, 1.7)
    if kind == CONST:
        t = _const_for(key, shape)
        assert tuple(t.shape) == tuple(shape), (key, t.shape, shape)
        return t
    raise ValueError(kind)


_COMPONENT_PREFIXES = ("encoder.", "decoder.", "speaker_encoder.")


def canonical_key(key: str, shape) -> str:
    """Prefix-independent name used to seed a tensor, so `decoder_spec()` and
    `decoder_spec("decoder.")` draw identical weights."""
    for p in _COMPONENT_PREFIXES:
        if key.startswith(p):
            return key[len(p):]
    return key


def seeded_state_dict(spec: "OrderedDict[str, tuple]", seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, seeded_tensor(canonical_key(k, shape), shape, kind, seed)) for k, (shape, kind) in spec.items())


def full_spec() -> "OrderedDict[str, tuple]":
    """encoder.* + decoder.* + speaker_encoder.* — the layout of a demo.py checkpoint
    (demo.py:30-38) minus the out-of-scope vgg_face.* keys."""
    s: OrderedDict = OrderedDict()
    s.update(encoder_spec("encoder."))
    s.update(decoder_spec("decoder."))
    s.update(speaker_spec("speaker_encoder."))
    return s


def content_min_t(T: int) -> int:
    """min over the strided Content.agg outputs (decoder.py:240-245): floor((T-k)/k)+1."""
    return min([T] + [(T - k) // k + 1 for k in CONTENT_KS])


def manifest(spec) -> dict:
    return {k: list(shape) for k, (shape, _) in spec.items()}
