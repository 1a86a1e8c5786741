"""AFx-Rep encoder (PANNs Cnn14) -- parameter container + B200 forward.

Mirrors the reference's ``st_ito.models.panns.Cnn14`` (panns.py:121-281): same constructor
arguments, same ``state_dict`` keys (including torchlibrosa's
``spectrogram_extractor.stft.conv_{real,imag}.weight`` and ``logmel_extractor.melW``), same
``forward(x[bs, chs, L]) -> (mid, side)`` returning RAW embeddings.  The arithmetic does not run in
torch: ``forward`` hands the audio to libstito (stito_embed), which computes mid/side, the
FFT-based log-mel front-end, the 12 convolutions (tcgen05 tensor cores) and the heads on the B200.
The STFT convolution kernels of a checkpoint are accepted but not used (the FFT is analytic);
``melW`` IS used, so a checkpoint's own filterbank is honoured.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

CHANNELS = (1, 64, 128, 256, 512, 1024, 2048)

# cfg/model/pretext/param-panns-concat-l2.yaml:16-25 of the reference
AFX_REP_ARGS = dict(embed_dim=512, sample_rate=48000, window_size=2048, hop_size=1024, mel_bins=128,
                    fmin=20, fmax=20000, use_batchnorm=True, input_norm="minmax")


def _slaney_hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3.0)
    log_region = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(f >= 1000.0, log_region, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * (200.0 / 3.0)
    log_region = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return np.where(m >= 15.0, log_region, lin)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax) -> np.ndarray:
    """Slaney-scale, Slaney-normalised triangular filterbank [n_fft//2+1, n_mels] (float32): what
    torchlibrosa's LogmelFilterBank stores as ``melW`` (= librosa.filters.mel(...).T)."""
    freqs = np.arange(n_fft // 2 + 1, dtype=np.float64) * (float(sr) / n_fft)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(fmin), _slaney_hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    out = np.zeros((n_mels, freqs.size), dtype=np.float32)
    for m in range(n_mels):
        rising = (freqs - edges[m]) / width[m]
        falling = (edges[m + 2] - freqs) / width[m + 1]
        out[m] = np.maximum(0.0, np.minimum(rising, falling))
    area_norm = 2.0 / (edges[2:] - edges[:-2])
    return np.ascontiguousarray((out.astype(np.float64) * area_norm[:, None]).astype(np.float32).T)


class _StftKernels(nn.Module):
    """Holds torchlibrosa's frozen DFT convolution weights so checkpoints load with strict keys."""

    def __init__(self, n_fft):
        super().__init__()
        n_freq = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, n_freq, n_fft, bias=False)
        self.conv_imag = nn.Conv1d(1, n_freq, n_fft, bias=False)
        n = np.arange(n_fft, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)
        ang = -2.0 * np.pi * np.outer(np.arange(n_freq, dtype=np.float64), n) / n_fft
        with torch.no_grad():
            self.conv_real.weight.copy_(torch.from_numpy((np.cos(ang) * window).astype(np.float32))[:, None, :])
            self.conv_imag.weight.copy_(torch.from_numpy((np.sin(ang) * window).astype(np.float32))[:, None, :])
        for p in self.parameters():
            p.requires_grad = False


class _Spectrogram(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.stft = _StftKernels(n_fft)


class _LogmelFilterBank(nn.Module):
    def __init__(self, sr, n_fft, n_mels, fmin, fmax):
        super().__init__()
        self.melW = nn.Parameter(torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, fmin, fmax)),
                                 requires_grad=False)


class ConvBlock(nn.Module):
    """reference panns.py:25-80 (parameters only; eval-mode BN is folded inside libstito)"""

    def __init__(self, in_channels, out_channels, use_batchnorm: bool = True):
        super().__init__()
        if not use_batchnorm:
            raise NotImplementedError("AFx-Rep uses use_batchnorm=True; the B200 path folds BatchNorm")
        self.conv1 = nn.Conv2d(in_channels, out_channels, (3, 3), (1, 1), (1, 1), bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, (3, 3), (1, 1), (1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        nn.init.xavier_uniform_(self.conv1.weight)  # init_layer, panns.py:10-16
        nn.init.xavier_uniform_(self.conv2.weight)


class Cnn14(nn.Module):
    def __init__(self, embed_dim: int, sample_rate: float, window_size: int, hop_size: int, mel_bins: int,
                 fmin: int, fmax: int, use_batchnorm: bool = True, input_norm: str = "minmax"):
        super().__init__()
        if input_norm != "minmax":
            raise NotImplementedError("the B200 path implements input_norm='minmax' (AFx-Rep)")
        if window_size != 2048 or mel_bins != 128:
            raise NotImplementedError("the B200 front-end is built for window_size=2048, mel_bins=128 (AFx-Rep)")
        self.embed_dim = embed_dim
        self.sample_rate = sample_rate
        self.window_size, self.hop_size, self.mel_bins = window_size, hop_size, mel_bins
        self.input_norm = input_norm
        self.spectrogram_extractor = _Spectrogram(window_size)
        self.logmel_extractor = _LogmelFilterBank(sample_rate, window_size, mel_bins, fmin, fmax)
        self.bn0 = nn.BatchNorm2d(mel_bins)  # in checkpoints; unused with minmax (panns.py:178, 233)
        for i in range(6):
            setattr(self, f"conv_block{i + 1}", ConvBlock(CHANNELS[i], CHANNELS[i + 1], use_batchnorm))
        self.fc_mid = nn.Linear(2048, embed_dim)
        self.fc_side = nn.Linear(2048, embed_dim)
        for fc in (self.fc_mid, self.fc_side):  # init_layer
            nn.init.xavier_uniform_(fc.weight)
            fc.bias.data.fill_(0.0)
        self._engine = None
        self._engine_key = None

    # -- B200 side ---------------------------------------------------------------------------
    def stito_engine(self, device_index=None):
        """The libstito handle holding this model's weights on a B200 (created on first use)."""
        from ..engine import Engine, default_device

        dev = default_device() if device_index is None else device_index
        key = (dev, self._weights_version())
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(model=self, device=dev)
            self._engine_key = key
        return self._engine

    def _weights_version(self):
        return sum(int(p._version) for p in self.parameters()) + sum(int(b._version) for b in self.buffers())

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        return out

    def forward(self, x: torch.Tensor):
        """x [bs, chs, seq_len] -> (mid_embed, side_embed) raw [bs, embed_dim] (panns.py:209-281)."""
        if self.training:
            raise NotImplementedError("the B200 encoder is inference-only; call model.eval()")
        bs, chs, seq_len = x.size()
        if chs not in (1, 2):
            raise ValueError(f"Invalid number of channels: {chs}")
        return self.stito_engine().embed(x, peak_normalize=False)
