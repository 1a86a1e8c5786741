"""Oracle: Cnn14 front-end (test infrastructure only -- see oracle/__init__.py).

The reference builds its front-end from ``torchlibrosa.stft.Spectrogram`` and
``LogmelFilterBank`` (st_ito/models/panns.py:139-168, used at :230-231) and then
min-max normalises (panns.py:238-241).  torchlibrosa / librosa are not vendored
and absent here, so this is a restatement of their published algorithm
[recollection] -- PARITY UNPINNED; cross-checked in tests against
``torch.stft`` and ``torchaudio.functional.melscale_fbanks(slaney, slaney)``:

  * STFT as two Conv1d(1 -> n_fft/2+1, kernel n_fft, stride hop) whose weights
    are the (periodic-Hann-windowed) real / imaginary DFT matrix rows, computed
    in fp64 and stored as fp32; centre=True with reflect padding n_fft/2;
  * power spectrogram real^2 + imag^2;
  * mel = S @ melW, melW = librosa.filters.mel(htk=False, norm="slaney").T, fp32;
  * 10*log10(clamp(mel, amin)) - 10*log10(max(amin, ref)), top_db=None.
The modules register parameters under torchlibrosa's names
(``stft.conv_real.weight``, ``stft.conv_imag.weight``, ``melW``) so that a real
AFx-Rep checkpoint's keys line up.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window("hann", n, fftbins=True) in fp64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def windowed_dft_kernels(n_fft: int):
    """(real[F, n_fft], imag[F, n_fft]) fp32: rows of omega^(k*t) * window, omega = exp(-2*pi*i/n)."""
    t, k = np.meshgrid(np.arange(n_fft), np.arange(n_fft))
    omega = np.exp(-2 * np.pi * 1j / n_fft)
    W = np.power(omega, t * k)  # [k, t]; same formula torchlibrosa uses (complex power in fp64)
    n_freq = n_fft // 2 + 1
    Ww = W[:n_freq, :] * hann_periodic(n_fft)[None, :]
    return np.real(Ww).astype(np.float32), np.imag(Ww).astype(np.float32)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm="slaney") -> [n_mels, F] fp32."""
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, fftfreqs.size), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w.astype(np.float64) * enorm[:, None]).astype(np.float32)


class _STFT(nn.Module):
    def __init__(self, n_fft, hop_length):
        super().__init__()
        n_freq = n_fft // 2 + 1
        self.n_fft, self.hop = n_fft, hop_length
        self.conv_real = nn.Conv1d(1, n_freq, n_fft, stride=hop_length, padding=0, dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(1, n_freq, n_fft, stride=hop_length, padding=0, dilation=1, groups=1, bias=False)
        re, im = windowed_dft_kernels(n_fft)
        self.conv_real.weight.data = torch.from_numpy(re)[:, None, :].contiguous()
        self.conv_imag.weight.data = torch.from_numpy(im)[:, None, :].contiguous()
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        x = x[:, None, :]
        x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode="reflect")
        real = self.conv_real(x)[:, None, :, :].transpose(2, 3)
        imag = self.conv_imag(x)[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    """torchlibrosa.stft.Spectrogram(power=2.0): (B, L) -> (B, 1, T, n_fft/2+1)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", power=2.0, freeze_parameters=True):
        super().__init__()
        assert window == "hann" and center and pad_mode == "reflect" and power == 2.0
        assert win_length in (None, n_fft)
        self.stft = _STFT(n_fft, hop_length if hop_length is not None else n_fft // 4)

    def forward(self, x):
        real, imag = self.stft(x)
        return real ** 2 + imag ** 2


class LogmelFilterBank(nn.Module):
    """torchlibrosa.stft.LogmelFilterBank: (B,1,T,F) power -> (B,1,T,n_mels) dB."""

    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True, ref=1.0,
                 amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        assert top_db is None, "AFx-Rep uses top_db=None (panns.py:145)"
        self.is_log, self.ref, self.amin = is_log, ref, amin
        melW = slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax if fmax is not None else sr // 2).T
        self.melW = nn.Parameter(torch.from_numpy(np.ascontiguousarray(melW)), requires_grad=False)

    def forward(self, x):
        mel = torch.matmul(x, self.melW)
        if not self.is_log:
            return mel
        out = 10.0 * torch.log10(torch.clamp(mel, min=self.amin, max=np.inf))
        out -= 10.0 * np.log10(np.maximum(self.amin, self.ref))
        return out


def minmax_norm(logmel: torch.Tensor) -> torch.Tensor:
    """panns.py:238-241."""
    x = logmel.clamp(-80, 40.0)
    x = (x + 80) / 120
    return (x * 2) - 1
