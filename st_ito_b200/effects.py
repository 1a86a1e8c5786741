"""Built-in effect plugins of the ES path, mirroring the reference's duck-typed plugin protocol
(st_ito/effects.py:784-959): ``plugin.parameters[name].raw_value`` on [0, 1] and
``plugin.process(x[chs, L] float32, sample_rate) -> ndarray[chs, L]``.

The classes carry only parameters; ``process`` renders on the B200 through libstito
(stito_process, include/stito.h).  When these plugins are used inside ``process_audio`` /
``run_es`` the whole chain is compiled into one descriptor and rendered for the entire population
at once (st_ito_b200/engine.py) -- ``process`` is the single-plugin entry the protocol requires.

Arithmetic notes (what the kernels implement; the CPU restatement used by the tests lives in the oracle package):
  * BasicParametricEQ   RBJ low-shelf + 4 peaking + high-shelf, DF-II-transposed, fp64
                        (reference effects.py:395-512 + scipy.signal.lfilter)
  * BasicCompressor     pedalboard.Compressor = juce::dsp::Compressor<float>      [recollection]
  * BasicDistortion     pedalboard.Distortion (tanh) + pedalboard.Gain            [recollection]
  * BasicDelay          pedalboard.Delay (whole-sample feedback delay, dry/wet)   [recollection]
  * BasicReverb         pedalboard.Reverb = juce::Reverb (Freeverb)               [recollection]
  * BasicNoiseShapedReverb  the convolution reverb of apply_reverb (effects.py:558-620) =
                        dasp_pytorch.noise_shaped_reverberation                   [recollection]
"""
from __future__ import annotations

import numpy as np

from . import _lib


class Parameter:
    """reference effects.py:784-797"""

    def __init__(self, init_value: float, min_value: float, max_value: float):
        self.min_value = min_value
        self.max_value = max_value
        self.set_value(init_value)

    def set_value(self, value: float):
        """Normalize the value to the range [0, 1] and store it as raw_value."""
        assert self.min_value <= value <= self.max_value
        self.raw_value = (value - self.min_value) / (self.max_value - self.min_value)

    def get_value(self):
        """Denormalize the value to the range [min_value, max_value]."""
        return self.raw_value * (self.max_value - self.min_value) + self.min_value


class _BasicPlugin:
    """Common part of the Basic* plugins: a ``parameters`` dict and a CUDA ``process``."""

    stito_kind: int = -1
    _spec: tuple = ()  # ((name, default, lo, hi), ...) in the order of the reference's dict

    def _init_parameters(self, values):
        self.parameters = {}
        for (name, _default, lo, hi), v in zip(self._spec, values):
            self.parameters[name] = Parameter(v, lo, hi)

    def process(self, x: np.ndarray, sample_rate: float):
        from .engine import render_single_plugin

        return render_single_plugin(self, x, sample_rate)


def _eq_spec():
    spec = []
    for name, fc, flo, fhi in (("low_shelf", 80.0, 20.0, 4000.0), ("band0", 300.0, 20.0, 10000.0),
                               ("band1", 1000.0, 20.0, 10000.0), ("band2", 3000.0, 20.0, 10000.0),
                               ("band3", 10000.0, 20.0, 10000.0), ("high_shelf", 1000.0, 200.0, 18000.0)):
        spec += [(f"{name}_gain_db", 0.0, -24.0, 24.0), (f"{name}_cutoff_freq", fc, flo, fhi),
                 (f"{name}_q_factor", 0.707, 0.1, 4.0)]
    return tuple(spec)


class BasicParametricEQ(_BasicPlugin):
    """reference effects.py:800-873"""

    stito_kind = _lib.FX_EQ
    _spec = _eq_spec()

    def __init__(self, low_shelf_gain_db=0.0, low_shelf_cutoff_freq=80.0, low_shelf_q_factor=0.707,
                 band0_gain_db=0.0, band0_cutoff_freq=300.0, band0_q_factor=0.707,
                 band1_gain_db=0.0, band1_cutoff_freq=1000.0, band1_q_factor=0.707,
                 band2_gain_db=0.0, band2_cutoff_freq=3000.0, band2_q_factor=0.707,
                 band3_gain_db=0.0, band3_cutoff_freq=10000.0, band3_q_factor=0.707,
                 high_shelf_gain_db=0.0, high_shelf_cutoff_freq=1000.0, high_shelf_q_factor=0.707):
        self._init_parameters([
            low_shelf_gain_db, low_shelf_cutoff_freq, low_shelf_q_factor,
            band0_gain_db, band0_cutoff_freq, band0_q_factor,
            band1_gain_db, band1_cutoff_freq, band1_q_factor,
            band2_gain_db, band2_cutoff_freq, band2_q_factor,
            band3_gain_db, band3_cutoff_freq, band3_q_factor,
            high_shelf_gain_db, high_shelf_cutoff_freq, high_shelf_q_factor])


class BasicCompressor(_BasicPlugin):
    """reference effects.py:876-897"""

    stito_kind = _lib.FX_COMPRESSOR
    _spec = (("threshold_db", 0.0, -80.0, 0.0), ("ratio", 4.0, 1.0, 20.0), ("attack_ms", 1.0, 0.1, 100.0),
             ("release_ms", 100.0, 10.0, 1000.0))

    def __init__(self, threshold_db=0.0, ratio=4.0, attack_ms=1.0, release_ms=100.0):
        self._init_parameters([threshold_db, ratio, attack_ms, release_ms])


class BasicDistortion(_BasicPlugin):
    """reference effects.py:900-914 (like the reference, the constructor ignores its arguments)"""

    stito_kind = _lib.FX_DISTORTION
    _spec = (("drive_db", 0.0, -48.0, 48.0), ("output_gain_db", 0.0, -24.0, 24.0))

    def __init__(self, drive_db=0.0, output_gain_db=0.0):
        self._init_parameters([0.0, 0.0])


class BasicDelay(_BasicPlugin):
    """reference effects.py:917-934"""

    stito_kind = _lib.FX_DELAY
    _spec = (("delay_seconds", 0.5, 0.01, 1.0), ("feedback", 0.5, 0.05, 1.0), ("mix", 0.5, 0.0, 1.0))

    def __init__(self, delay_seconds=0.5, feedback=0.5, mix=0.5):
        self._init_parameters([delay_seconds, feedback, mix])


class BasicReverb(_BasicPlugin):
    """reference effects.py:937-959"""

    stito_kind = _lib.FX_REVERB
    _spec = (("room_size", 0.5, 0.0, 1.0), ("damping", 0.5, 0.0, 1.0), ("wet_dry", 0.5, 0.0, 1.0),
             ("width", 0.5, 0.0, 1.0))

    def __init__(self, room_size=0.5, damping=0.5, wet_dry=0.5, width=0.5):
        self._init_parameters([room_size, damping, wet_dry, width])


def _nsr_spec():
    spec = [(f"band{b}_gain", 1.0, 0.0, 1.0) for b in range(12)]
    spec += [(f"band{b}_decay", d, 0.0, 1.0)
             for b, d in enumerate((0.6, 0.4, 0.4, 0.5, 0.2, 0.3, 0.3, 0.2, 0.1, 0.1, 0.2, 0.1))]  # dsp.py:28-30
    return tuple(spec + [("mix", 0.5, 0.0, 1.0)])


class BasicNoiseShapedReverb(_BasicPlugin):
    """Convolution reverb with a noise-shaped impulse response, in the plugin protocol of the ES path.

    The reference reaches this effect only through ``apply_reverb`` (effects.py:558-620, run_autodiff) and
    ``apply_random_reverb`` (dsp.py:26-46); BASELINE config 4 puts it on the ES path ("2 s-IR conv reverb").  The 25
    parameters are the ones of effects.py:564-588 (12 band gains, 12 band decays, mix), all used raw on [0, 1].
    ``num_samples`` is the impulse-response length (65 536 = dasp-pytorch's default), ``seed`` fixes the white noise
    that the reference re-draws on every call (see libstito / oracle/convreverb.py).  A 2-channel plugin.
    """

    stito_kind = _lib.FX_CONV_REVERB
    _spec = _nsr_spec()

    def __init__(self, num_samples: int = 65536, seed: int = 0):
        self.num_samples, self.seed = int(num_samples), int(seed)
        self._init_parameters([s[1] for s in self._spec])

    @property
    def stito_iopt(self):
        return (self.num_samples, self.seed, 0, 0)


class BasicNoiseShapedReverb2s(BasicNoiseShapedReverb):
    """BASELINE config 4: a 2 s impulse response (96 000 taps at 48 kHz)."""

    def __init__(self):
        super().__init__(num_samples=96000, seed=0)


class BasicLTICompressor(_BasicPlugin):
    """Compressor with LTI gain smoothing, in the plugin protocol of the ES path.

    The reference reaches this effect only through ``apply_compressor`` (effects.py:623-648, run_autodiff / the style
    chain) and ``apply_random_compressor`` (dsp.py:49-78): dasp-pytorch's ``compressor``.  The six parameters and their
    ranges are effects.py:629-634 (``release_ms`` is accepted and unused, as upstream); ``lookahead_samples`` = 512 is
    effects.py:646.  ``num_channels: 2`` in the chain entry links the channels (summed side-chain, what the reference
    does for a stereo tensor); ``1`` compresses every channel on its own.
    """

    stito_kind = _lib.FX_LTI_COMPRESSOR
    _spec = (("threshold_db", -24.0, -60.0, 0.0), ("ratio", 4.0, 1.0, 20.0), ("attack_ms", 10.0, 0.1, 250.0),
             ("release_ms", 100.0, 10.0, 2000.0), ("knee_db", 6.0, 1.0, 24.0), ("makeup_gain_db", 0.0, 0.0, 24.0))

    def __init__(self, lookahead_samples: int = 512):
        self.lookahead_samples = int(lookahead_samples)
        self._init_parameters([s[1] for s in self._spec])

    @property
    def stito_iopt(self):
        return (self.lookahead_samples, 0, 0, 0)


def is_native_plugin(obj) -> bool:
    """True for plugins libstito can render (their ranges are the ones compiled into the library)."""
    if not isinstance(obj, _BasicPlugin) or obj.stito_kind < 0:
        return False
    names = [s[0] for s in obj._spec]
    if list(obj.parameters.keys()) != names:
        return False
    return all(obj.parameters[n].min_value == lo and obj.parameters[n].max_value == hi
               for n, _d, lo, hi in obj._spec)


# Chain presets of the reference: run_optim.py:376-407 ("basic") and eval_pst.py:559-579
# ("mastering-pb"); "eq" is BASELINE config 1.
def make_chain(kind: str = "basic") -> dict:
    table = {
        "ParametricEQ": (BasicParametricEQ, 1), "Compressor": (BasicCompressor, 1),
        "Distortion": (BasicDistortion, 1), "Delay": (BasicDelay, 2), "Reverb": (BasicReverb, 2),
        "NoiseShapedReverb": (BasicNoiseShapedReverb2s, 2), "LTICompressor": (BasicLTICompressor, 2),
    }
    presets = {
        "basic": ["ParametricEQ", "Compressor", "Distortion", "Delay", "Reverb"],
        "mastering-pb": ["ParametricEQ", "Compressor", "Reverb"],
        "eq": ["ParametricEQ"],
        "mastering-conv": ["ParametricEQ", "Compressor", "NoiseShapedReverb"],  # BASELINE config 4
        # SURVEY row R2 in full: the dasp-style pair of effects.py:558-648 behind the EQ
        "mastering-dasp": ["ParametricEQ", "LTICompressor", "NoiseShapedReverb"],
    }
    if kind not in presets:
        raise ValueError(f"Unknown chain: {kind}")
    return {name: {"class_path": table[name][0], "num_params": None, "num_channels": table[name][1],
                   "fixed_parameters": {}} for name in presets[kind]}
