"""Deterministic synthetic inputs of LRW / AVSpeech shape.

Drawn with numpy PCG64 so the build container (golden generation), the GPU box (parity
tests, bench) and the CPU baseline all see bit-identical tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import spec


def _rng(seed, tag):
    return np.random.Generator(np.random.PCG64([int(seed), int(tag)]))


def video(B, T=29, H=96, W=96, seed=1234) -> torch.Tensor:
    """[B,3,T,H,W] fp32 ~ N(0,1): the ImageNet-normalised range of datasets/lrw/dataset.py:84-85."""
    a = _rng(seed, 1).standard_normal((B, 3, T, H, W), dtype=np.float32)
    return torch.from_numpy(a)


def frames_u8(B, T=29, H=96, W=96, seed=1234) -> torch.Tensor:
    """[B,T,H,W,3] uint8 RGB: decoded mouth crops as loadframes() returns them (datasets/lrw/dataset.py:20-24)."""
    a = _rng(seed, 7).integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    return torch.from_numpy(a)


def normalise_frames(frames: torch.Tensor) -> torch.Tensor:
    """The dataset's host-side arithmetic (datasets/lrw/dataset.py:82-86,132) + the collate permute
    (datasets/__init__.py:7-46): uint8 [B,T,H,W,3] -> fp32 [B,3,T,H,W] = (x/255 - mean) / std."""
    x = frames.permute(0, 1, 4, 2, 3).float() / 255.0                 # [B,T,3,H,W]
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 1, 3, 1, 1)
    return ((x - mean) / std).permute(0, 2, 1, 3, 4).contiguous()


def wav(B, S=19456, seed=1234) -> torch.Tensor:
    """[B,S] fp32, 0.1*N(0,1): 1.216 s of 16 kHz audio per LRW clip."""
    a = _rng(seed, 2).standard_normal((B, S), dtype=np.float32) * np.float32(0.1)
    return torch.from_numpy(a)


def speaker_embedding(B, seed=1234) -> torch.Tensor:
    """normalize(relu(N(0,1))) [B,256] — the range of SpeakerEncoder.inference (audio.py:144-150)."""
    a = np.maximum(_rng(seed, 3).standard_normal((B, spec.SPK_DIM), dtype=np.float32), 0)
    a = a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-12)
    return torch.from_numpy(a.astype(np.float32))


def gumbel(B, T=29, seed=1234) -> torch.Tensor:
    """g = -log(E), E~Exp(1), shape [B*minT, 501], row b*minT+m (decoder.py:253-257)."""
    n = B * spec.content_min_t(T)
    e = _rng(seed, 4).exponential(1.0, (n, spec.VOCAB))
    return torch.from_numpy((-np.log(e)).astype(np.float32))


def visual_features(B, T=29, seed=1234):
    """Decoder-only test input: (visual [B,T,1024], face_tiled [B,T,256]) as model.py:52-55 builds
    them — L2-normalised 768-d frame features concatenated with the tiled speaker embedding."""
    f = _rng(seed, 5).standard_normal((B, T, spec.VIDEO_FEAT), dtype=np.float32)
    f = f / np.linalg.norm(f, axis=2, keepdims=True)
    face = speaker_embedding(B, seed).unsqueeze(1).repeat(1, T, 1)
    return torch.cat([torch.from_numpy(f.astype(np.float32)), face], dim=2), face


def mel_like(B, L=300, seed=1234) -> torch.Tensor:
    """[B,80,L] postnet test input in the range of decoder outputs."""
    a = _rng(seed, 6).standard_normal((B, spec.N_MELS, L), dtype=np.float32)
    return torch.from_numpy(a)
