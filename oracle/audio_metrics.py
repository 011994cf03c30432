"""Vocoder + ESTOI used ONLY to check the north-star criterion "ESTOI on SAMPLE_LRW within +-0.001 of the reference"
(TEST INFRASTRUCTURE; never imported by the product path).

* mel_to_audio restates datasets/spectograms.py:76-95 (MelSpec2Audio: exp -> InverseMelScale -> GriffinLim, n_fft 1024,
  hop 256, 80 mels, 0-8 kHz) with the torchaudio of this image.  torchaudio 2.11 dropped InverseMelScale(max_iter=...)
  (SURVEY.md §8c), so the inverse uses its least-squares solver; Griffin-Lim's random phase is seeded.
* estoi restates pystoi 0.3.3 `stoi(x, y, fs, extended=True)` (evaluate.py:45; Jensen & Taal 2016): resample to 10 kHz,
  drop frames more than 40 dB below the loudest (256/128 Hann frames), 512-point STFT, 15 one-third-octave bands from
  150 Hz, 30-frame segments, row+column normalisation, mean correlation.
The +-0.001 criterion compares the SAME implementation applied to the reference-path mel and to the CUDA-path mel.
"""
from __future__ import annotations

import numpy as np
import torch

FS = 10000
N_FRAME = 256
NFFT = 512
NUMBAND = 15
MINFREQ = 150
N_SEG = 30
DYN_RANGE = 40
EPS = np.finfo("float").eps


def mel_to_audio(mel: torch.Tensor, seed: int = 0, n_iter: int = 256) -> torch.Tensor:
    """mel [B,80,L] (log-mel as produced by the decoder) -> waveform [B, S]."""
    import torchaudio.transforms as T
    inv = T.InverseMelScale(sample_rate=16000, n_mels=80, f_min=0.0, f_max=8000.0, n_stft=513)
    gl = T.GriffinLim(n_fft=1024, win_length=1024, hop_length=256, n_iter=n_iter)
    torch.manual_seed(seed)                                # GriffinLim(rand_init=True) draws the initial phase
    return gl(inv(torch.exp(mel.double()).float()))


def _thirdoct(fs, nfft, num_bands, min_freq):
    f = np.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = np.arange(num_bands).astype(float)
    cf = np.power(2.0 ** (1.0 / 3), k) * min_freq
    freq_low = min_freq * np.power(2.0, (2 * k - 1) / 6)
    freq_high = min_freq * np.power(2.0, (2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    for i in range(len(cf)):
        fl = np.argmin(np.square(f - freq_low[i]))
        fh = np.argmin(np.square(f - freq_high[i]))
        obm[i, fl:fh] = 1
    return obm


def _frames(x, framelen, hop):
    n = 1 + (len(x) - framelen) // hop if len(x) >= framelen else 0
    idx = np.arange(framelen)[None, :] + hop * np.arange(n)[:, None]
    return x[idx] if n else np.zeros((0, framelen))


def _remove_silent_frames(x, y, dyn_range, framelen, hop):
    w = np.hanning(framelen + 2)[1:-1]
    xf, yf = _frames(x, framelen, hop) * w, _frames(y, framelen, hop) * w
    energies = 20 * np.log10(np.linalg.norm(xf, axis=1) + EPS)
    mask = (np.max(energies) - dyn_range - energies) < 0
    xf, yf = xf[mask], yf[mask]
    n = (len(xf) - 1) * hop + framelen if len(xf) else 0
    xs, ys = np.zeros(n), np.zeros(n)
    for i in range(len(xf)):
        xs[i * hop: i * hop + framelen] += xf[i]
        ys[i * hop: i * hop + framelen] += yf[i]
    return xs, ys


def _stft(x, win_size, fft_size, overlap=2):
    hop = win_size // overlap
    w = np.hanning(win_size + 2)[1:-1]
    fr = _frames(x, win_size, hop) * w
    return np.fft.rfft(fr, n=fft_size).T if len(fr) else np.zeros((fft_size // 2 + 1, 0))


def _row_col_normalize(x):
    x = x - np.mean(x, axis=-1, keepdims=True)
    x = x / (np.linalg.norm(x, axis=-1, keepdims=True) + EPS)
    x = x - np.mean(x, axis=1, keepdims=True)
    x = x / (np.linalg.norm(x, axis=1, keepdims=True) + EPS)
    return x


def estoi(clean: np.ndarray, processed: np.ndarray, fs_sig: int = 16000) -> float:
    from scipy.signal import resample_poly
    n = min(len(clean), len(processed))
    x, y = np.asarray(clean[:n], dtype=np.float64), np.asarray(processed[:n], dtype=np.float64)
    if fs_sig != FS:
        g = np.gcd(FS, fs_sig)
        x, y = resample_poly(x, FS // g, fs_sig // g), resample_poly(y, FS // g, fs_sig // g)
    x, y = _remove_silent_frames(x, y, DYN_RANGE, N_FRAME, N_FRAME // 2)
    xs, ys = _stft(x, N_FRAME, NFFT), _stft(y, N_FRAME, NFFT)
    if xs.shape[-1] < N_SEG:
        return 1e-5                                      # pystoi's "not enough frames" value
    obm = _thirdoct(FS, NFFT, NUMBAND, MINFREQ)
    x_tob = np.sqrt(obm @ np.square(np.abs(xs)))
    y_tob = np.sqrt(obm @ np.square(np.abs(ys)))
    segs = np.arange(N_SEG, x_tob.shape[1] + 1)
    x_seg = np.array([x_tob[:, m - N_SEG:m] for m in segs])
    y_seg = np.array([y_tob[:, m - N_SEG:m] for m in segs])
    xn, yn = _row_col_normalize(x_seg), _row_col_normalize(y_seg)
    return float(np.sum(xn * yn / N_SEG) / xn.shape[0])
