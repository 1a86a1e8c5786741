"""Oracle: noise-shaped convolution reverb (test infrastructure only -- see oracle/__init__.py).

PARITY UNPINNED.  Restates what the reference's ``apply_reverb`` (st_ito/effects.py:558-620) and
``apply_random_reverb`` (st_ito/dsp.py:26-46) call: ``dasp_pytorch.noise_shaped_reverberation`` --
dasp-pytorch is an absent, un-pinned dependency (setup.py:51), so the arithmetic below is its published
algorithm as recalled (dasp_pytorch/functional.py, ``noise_shaped_reverberation`` + ``octave_band_filterbank``):

  1. 12 octave-band FIR filters, 1023 taps, ``scipy.signal.firwin`` (Hamming): low-pass 12 Hz; band-passes
     fc/sqrt(2) .. fc*sqrt(2) for fc in 31.5 .. 16000 Hz (upper edge clipped to 0.999*Nyquist); high-pass 18 kHz;
     cast to float32.
  2. white noise [2 channels][12 bands][num_samples + 1022] filtered band by band ("valid" correlation).
  3. per band: envelope exp(-(10*decay + 1) * t), t = linspace(0, 1, num_samples); times gain; mean over bands
     = a stereo impulse response of num_samples taps.
  4. y = causal convolution of every channel with its impulse response (mono is up-mixed to stereo first),
     out = (1 - mix) * x + mix * y.

Deterministic choice (the reference draws a FRESH ``torch.randn`` on every call, i.e. it is not reproducible and its
objective is noisy): the white noise is a pure function of (seed, element index) -- splitmix64 -> Box-Muller, see
``white_noise`` -- drawn ONCE per plugin instance, so the filtered-noise bands are candidate-independent and
only the 12 gains, 12 decays and the mix vary per candidate.  libstito restates the same generator in C++.

The ES path of the reference only has Basic* (pedalboard) plugins; ``OracleNoiseShapedReverb`` wraps the
function above in that plugin protocol (25 parameters on [0, 1], used raw exactly as effects.py:564-588 does).
"""
from __future__ import annotations

import numpy as np

NUM_BANDS = 12
NUM_TAPS = 1023
BAND_CENTRES = (31.5, 63.0, 125.0, 250.0, 500.0, 1000.0, 2000.0, 4000.0, 8000.0, 16000.0)


def octave_band_filterbank(num_taps: int, sample_rate: float) -> np.ndarray:
    """[12][num_taps] float32 (dasp_pytorch.functional.octave_band_filterbank; the flip it applies is a no-op on
    these symmetric linear-phase filters)."""
    from scipy.signal import firwin

    filts = [firwin(num_taps, 12, fs=sample_rate)]
    for fc in BAND_CENTRES:
        f_min = fc / np.sqrt(2)
        f_max = float(np.clip(fc * np.sqrt(2), 0, (sample_rate / 2) * 0.999))
        filts.append(firwin(num_taps, [f_min, f_max], fs=sample_rate, pass_zero=False))
    filts.append(firwin(num_taps, 18000, fs=sample_rate, pass_zero=False))
    return np.stack(filts).astype(np.float32)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def white_noise(seed: int, count: int) -> np.ndarray:
    """Gaussian white noise, float32[count]: element i = Box-Muller of two uniforms hashed from (seed, i).

    u1 = (splitmix64(key + 2i) >> 11 + 1) * 2^-53 in (0, 1],  u2 = (splitmix64(key + 2i + 1) >> 11) * 2^-53 in [0, 1),
    key = splitmix64(seed);  z = sqrt(-2 ln u1) * cos(2 pi u2) evaluated in float64, rounded to float32.
    """
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([seed], dtype=np.uint64))[0]
        i = np.arange(count, dtype=np.uint64)
        a = _splitmix64(key + np.uint64(2) * i)
        b = _splitmix64(key + np.uint64(2) * i + np.uint64(1))
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * 2.0 ** -53
    u2 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)


def filtered_noise_bands(sample_rate: float, num_samples: int, seed: int) -> np.ndarray:
    """[2][12][num_samples] float32: steps 1-2 (candidate-independent)."""
    from scipy.signal import fftconvolve

    filt = octave_band_filterbank(NUM_TAPS, sample_rate).astype(np.float64)
    span = num_samples + NUM_TAPS - 1
    wn = white_noise(seed, 2 * NUM_BANDS * span).reshape(2, NUM_BANDS, span).astype(np.float64)
    out = np.empty((2, NUM_BANDS, num_samples), dtype=np.float32)
    for c in range(2):
        for b in range(NUM_BANDS):
            # conv1d = correlation: out[n] = sum_t wn[n + t] * filt[t]  ("valid")
            out[c, b] = fftconvolve(wn[c, b], filt[b][::-1], mode="valid").astype(np.float32)
    return out


def impulse_response(bands: np.ndarray, gains, decays) -> np.ndarray:
    """Step 3 in float32, operation by operation: [2][num_samples]."""
    num_samples = bands.shape[-1]
    t = (np.arange(num_samples, dtype=np.float64) / (num_samples - 1)).astype(np.float32)
    g = np.asarray(gains, dtype=np.float32)
    d = np.asarray(decays, dtype=np.float32) * np.float32(10.0) + np.float32(1.0)
    acc = np.zeros((2, num_samples), dtype=np.float32)
    for b in range(NUM_BANDS):
        env = np.exp(-d[b] * t).astype(np.float32)
        acc += bands[:, b, :] * (env * g[b])[None, :]
    return acc / np.float32(NUM_BANDS)


def noise_shaped_reverberation(x: np.ndarray, sample_rate: float, gains, decays, mix, bands: np.ndarray) -> np.ndarray:
    """x [chs, L] float32 -> [2, L] float32 (mono is repeated to stereo first)."""
    from scipy.signal import fftconvolve

    x = np.asarray(x, dtype=np.float32)
    if x.shape[0] == 1:
        x = np.concatenate((x, x), axis=0)
    ir = impulse_response(bands, gains, decays)
    L = x.shape[1]
    y = np.stack([fftconvolve(x[c].astype(np.float64), ir[c].astype(np.float64))[:L] for c in range(2)])
    mix = np.float32(mix)
    return ((np.float32(1.0) - mix) * x + mix * y.astype(np.float32)).astype(np.float32)


class OracleNoiseShapedReverb:
    """Plugin-protocol wrapper (builder-defined: the reference's ES path has no plugin for this effect).

    25 parameters on [0, 1] in the order of effects.py:564-588: band0_gain..band11_gain, band0_decay..band11_decay, mix.
    """

    def __init__(self, num_samples: int = 65536, seed: int = 0):
        from oracle.dsp import Parameter

        self.num_samples, self.seed = int(num_samples), int(seed)
        self.parameters = {}
        for b in range(NUM_BANDS):
            self.parameters[f"band{b}_gain"] = Parameter(1.0, 0.0, 1.0)
        for b, d in enumerate((0.6, 0.4, 0.4, 0.5, 0.2, 0.3, 0.3, 0.2, 0.1, 0.1, 0.2, 0.1)):  # dsp.py:28-30
            self.parameters[f"band{b}_decay"] = Parameter(d, 0.0, 1.0)
        self.parameters["mix"] = Parameter(0.5, 0.0, 1.0)
        self._bands = {}

    def process(self, x, sample_rate):
        key = float(sample_rate)
        if key not in self._bands:
            self._bands[key] = filtered_noise_bands(sample_rate, self.num_samples, self.seed)
        v = [np.float32(q.get_value()) for q in self.parameters.values()]
        return noise_shaped_reverberation(x, sample_rate, v[:12], v[12:24], v[24], self._bands[key])


class OracleNoiseShapedReverb2s(OracleNoiseShapedReverb):
    """BASELINE config 4: a 2 s impulse response (96 000 taps at 48 kHz)."""

    def __init__(self):
        super().__init__(num_samples=96000, seed=0)
