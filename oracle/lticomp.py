"""Oracle: feed-forward compressor with LTI gain smoothing (test infrastructure only -- see oracle/__init__.py).

PARITY UNPINNED.  Restates what the reference's ``apply_compressor`` (st_ito/effects.py:623-648) and
``apply_random_compressor`` (st_ito/dsp.py:49-78) call: ``dasp_pytorch.compressor`` -- dasp-pytorch is an absent,
un-pinned dependency (setup.py:51), so the arithmetic below is its published algorithm as recalled
(dasp_pytorch/functional.py ``compressor`` + dasp_pytorch/signal.py ``lfilter_via_fsm``):

  1. side-chain = SUM over the channels; x_db = 20 log10(clamp(|side|, 1e-8))                       (float32)
  2. static curve with a soft knee W around the threshold T, ratio R:
        x_db >  T + W/2 :            x_sc = T + (x_db - T) / R
        T - W/2 <= x_db <= T + W/2 : x_sc = x_db + (1/R - 1) (x_db - T + W/2)^2 / (2 W)
        else :                       x_sc = x_db ;          g_c = x_sc - x_db  (dB, <= 0)          (float32)
  3. ONE-pole smoothing of g_c with the ATTACK constant only (the release time is accepted and ignored upstream):
        alpha = exp(-log(9) / (fs * attack_ms / 1000)),  H(z) = (1 - alpha) / (1 - alpha z^-1),
     applied by FREQUENCY SAMPLING (``lfilter_via_fsm``): n_fft = 2^ceil(log2(2 L - 1)), Y = rfft(g_c, n_fft) * H(e^jw),
     irfft, crop to L.  Sampling H on n_fft bins is a CIRCULAR convolution with the n_fft-periodic (time-aliased)
     impulse response, i.e. exactly the recursion y[n] = alpha y[n-1] + (1 - alpha) g_c[n] started from the periodic
     steady state y[-1] = y0[L-1] alpha^(n_fft - L) / (1 - alpha^n_fft) (y0 = the zero-state response).  That wrap-around
     term vanishes for the ES configurations (alpha^L < 1e-38 at 10 s) but not for short clips with a long attack;
     ``smooth_gain_recursive`` states it explicitly and ``tests/test_lticomp_oracle.py`` checks both forms agree.
  4. look-ahead: the INPUT is delayed by ``lookahead_samples`` (torch.roll + zeroing the head) against the gain;
     y = x_delayed * 10^((y_smooth + makeup_db) / 20).

Precision: the reference runs all of this in float32 torch (including the FFTs).  Steps 1, 2 and 4 are float32 here;
alpha is the correctly rounded float32 of the formula (torch's own expf may differ from it by an ulp, which for the
longest attacks moves the time constant by ~3e-4 relative: part of "unpinned"); the smoothing filter runs in float64.

The ES path of the reference only has Basic* (pedalboard) plugins; ``OracleLTICompressor`` wraps the function in that
plugin protocol with the six parameters and ranges of effects.py:629-646 and lookahead 512 (effects.py:646).
"""
from __future__ import annotations

import numpy as np

LOG9_F32 = np.float32(np.log(np.float32(9.0)))


def attack_alpha(attack_ms, sample_rate) -> np.float32:
    """alpha_A of dasp_pytorch.compressor in float32 steps: exp(-log(9) / (fs * (attack_ms / 1e3)))."""
    nat = np.float32(sample_rate) * (np.float32(attack_ms) / np.float32(1e3))
    arg = -LOG9_F32 / np.float32(nat)
    return np.float32(np.exp(np.float64(arg)))


def gain_computer_db(side: np.ndarray, threshold_db, ratio, knee_db) -> np.ndarray:
    """Steps 1-2 in float32: g_c[n] in dB."""
    f = np.float32
    T, R, W = f(threshold_db), f(ratio), f(knee_db)
    x_db = f(20.0) * np.log10(np.maximum(np.abs(side.astype(np.float32)), f(1e-8)), dtype=np.float32)
    x_sc = x_db.copy()
    half = W / f(2.0)
    knee = np.logical_and(x_db >= (T - half), x_db <= (T + half))
    below = x_db + ((f(1.0) / R) - f(1.0)) * ((x_db - T + half) ** 2) / (f(2.0) * W)
    x_sc[knee] = below[knee]
    above = x_db > (T + half)
    lin = T + ((x_db - T) / R)
    x_sc[above] = lin[above]
    return (x_sc - x_db).astype(np.float32)


def fsm_fft_size(L: int) -> int:
    return 1 << int(np.ceil(np.log2(max(2 * L - 1, 1))))


def smooth_gain_fsm(g_c: np.ndarray, alpha: np.float32) -> np.ndarray:
    """Step 3 as the reference does it (frequency sampling), in float64."""
    L = g_c.shape[-1]
    n_fft = fsm_fft_size(L)
    b0 = float(np.float32(1.0) - alpha)
    a1 = -float(alpha)
    w = np.exp(-2j * np.pi * np.arange(n_fft // 2 + 1) / n_fft)
    H = b0 / (1.0 + a1 * w)
    return np.fft.irfft(np.fft.rfft(g_c.astype(np.float64), n_fft) * H, n_fft)[..., :L]


def smooth_gain_recursive(g_c: np.ndarray, alpha: np.float32) -> np.ndarray:
    """Step 3 as a recursion plus the explicit wrap-around term (what libstito computes)."""
    from scipy.signal import lfilter

    L = g_c.shape[-1]
    n_fft = fsm_fft_size(L)
    a = float(alpha)
    b0 = float(np.float32(1.0) - alpha)
    y0 = lfilter([b0], [1.0, -a], g_c.astype(np.float64))
    if a <= 0.0:
        return y0
    ln_a = np.log(a)
    y_init = y0[..., -1] * np.exp(ln_a * (n_fft - L)) / (-np.expm1(ln_a * n_fft))
    return y0 + np.exp(ln_a * (np.arange(L) + 1.0)) * y_init


def lti_compressor(x: np.ndarray, sample_rate: float, threshold_db, ratio, attack_ms, release_ms, knee_db,
                   makeup_gain_db, lookahead_samples: int = 0) -> np.ndarray:
    """x [chs][L] float32 -> [chs][L] float32 (``release_ms`` is unused, as upstream)."""
    del release_ms
    x = np.asarray(x, dtype=np.float32)
    side = x.sum(axis=0, dtype=np.float32)
    g_c = gain_computer_db(side, threshold_db, ratio, knee_db)
    g_s = smooth_gain_fsm(g_c, attack_alpha(attack_ms, sample_rate))
    if lookahead_samples > 0:
        k = min(int(lookahead_samples), x.shape[1])
        xd = np.zeros_like(x)
        xd[:, k:] = x[:, : x.shape[1] - k]
    else:
        xd = x
    g_db = g_s.astype(np.float32) + np.float32(makeup_gain_db)
    g_lin = np.power(np.float32(10.0), g_db / np.float32(20.0), dtype=np.float32)
    return (xd * g_lin[None, :]).astype(np.float32)


class OracleLTICompressor:
    """Plugin-protocol wrapper (ranges: effects.py:629-646; lookahead 512: effects.py:646)."""

    def __init__(self, lookahead_samples: int = 512):
        from oracle.dsp import Parameter

        self.lookahead_samples = int(lookahead_samples)
        self.parameters = {
            "threshold_db": Parameter(-24.0, -60.0, 0.0),
            "ratio": Parameter(4.0, 1.0, 20.0),
            "attack_ms": Parameter(10.0, 0.1, 250.0),
            "release_ms": Parameter(100.0, 10.0, 2000.0),
            "knee_db": Parameter(6.0, 1.0, 24.0),
            "makeup_gain_db": Parameter(0.0, 0.0, 24.0),
        }

    def process(self, x, sample_rate):
        x = np.atleast_2d(np.asarray(x, dtype=np.float32))
        v = [np.float32(q.get_value()) for q in self.parameters.values()]
        return lti_compressor(x, sample_rate, *v, lookahead_samples=self.lookahead_samples)
