"""Host side of the steps right after the hot path: the reference's `MelSpec2Audio`
(datasets/spectograms.py:76-95) and the ESTOI metric of evaluate.py:44-45, as thin wrappers over the C ABI
(`l2s_vocoder`, `l2s_estoi`).  No arithmetic happens here except the one-time construction of the inverse mel operator."""
from __future__ import annotations

import torch

from . import _lib


def inverse_mel_operator(n_stft: int = 513, n_mels: int = 80, sample_rate: int = 16000, f_min: float = 0.0, f_max: float = 8000.0) -> torch.Tensor:
    """The linear operator torchaudio.transforms.InverseMelScale applies before its clamp: the least-squares solution of
    fb^T x = mel is linear in mel, so P = lstsq(fb^T, I).solution [n_stft, n_mels] (same `gelsd` driver torchaudio uses on CPU).
    A constant of the configuration (hparams.py:32-38), built once on the host and bound as the weight "vocoder.inv_mel"."""
    import torchaudio.functional as AF
    fb = AF.melscale_fbanks(n_stft, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk")       # [n_stft, n_mels]
    return torch.linalg.lstsq(fb.transpose(0, 1), torch.eye(n_mels), driver="gelsd").solution              # [n_stft, n_mels]


class MelSpec2Audio(torch.nn.Module):
    """datasets/spectograms.py:76-95 on the B200 backend: `forward(melspec [B,80,L]) -> waveform [B,(L-1)*256]`."""

    def __init__(self, max_iters: int = 256, momentum: float = 0.99, device: int = 0):
        super().__init__()
        self.max_iters, self.momentum = max_iters, momentum
        self.be = _lib.backend(device)
        self.be.bind_vocoder(inverse_mel_operator())

    def forward(self, melspec, init_angles=None):
        if init_angles is None:                 # GriffinLim(rand_init=True): torch.rand(size, dtype=complex) from the global generator
            B, _, L = melspec.shape
            init_angles = torch.rand(B, 513, L, dtype=torch.complex64)
        return self.be.vocoder(melspec, init_angles, self.max_iters, self.momentum)


def estoi(clean: torch.Tensor, processed: torch.Tensor, device: int = 0) -> torch.Tensor:
    """pystoi.stoi(gt, pred, 16000, extended=True) (evaluate.py:45) for every row of [B,S]; returns [B] float64 on the device."""
    return _lib.backend(device).estoi(clean, processed)
