"""Oracle: effect chain (test infrastructure only -- see oracle/__init__.py).

Restates the reference's plugin protocol and chain walk on the CPU:
  * ``Parameter`` / ``Basic*`` plugin objects   reference st_ito/effects.py:784-985
  * ``load_plugins`` / ``process_audio`` / ``parameters_to_dict``
                                                reference st_ito/style_transfer.py:17-115, 324-359
The per-sample arithmetic lives in oracle/dsp_oracle.c (built by oracle/Makefile).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_dsp.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/dsp_oracle.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "dsp_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.oracle_biquad_coefs.argtypes = [ctypes.c_double] * 4 + [ctypes.c_int, f64p, f64p]
        L.oracle_eq.argtypes = [f32p, ctypes.c_int64, f64p, ctypes.c_double, f32p]
        L.oracle_compressor.argtypes = [f32p, ctypes.c_int64] + [ctypes.c_float] * 4 + [ctypes.c_double, f32p]
        L.oracle_reverb.argtypes = [f32p, f32p, ctypes.c_int64, ctypes.c_int] + [ctypes.c_float] * 5 + [ctypes.c_double]
        L.oracle_distortion.argtypes = [f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_float, f32p]
        L.oracle_delay.argtypes = [f32p, ctypes.c_int64] + [ctypes.c_float] * 3 + [ctypes.c_double, f32p]
        L.oracle_peak_normalize.argtypes = [f32p, ctypes.c_int64]
        L.oracle_peak_normalize.restype = ctypes.c_float
        _lib = L
    return _lib


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _rows(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError("audio must be [chs, L]")
    return x


def biquad_coefs(gain_db, cutoff, q, fs, kind):
    """(b[3], a[3]) for kind in {"low_shelf","peaking","high_shelf"} (effects.py:395-450)."""
    code = {"low_shelf": 0, "peaking": 1, "high_shelf": 2}[kind]
    b = np.zeros(3)
    a = np.zeros(3)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib().oracle_biquad_coefs(gain_db, cutoff, q, fs, code, b.ctypes.data_as(f64p), a.ctypes.data_as(f64p))
    return b, a


# --------------------------------------------------------------------------
# plugin objects (the reference's duck-typed protocol: .parameters + .process)
# --------------------------------------------------------------------------
class Parameter:
    """raw_value in [0,1] <-> value in [lo,hi] (effects.py:784-797)."""

    def __init__(self, init_value, lo, hi):
        self.min_value, self.max_value = lo, hi
        self.set_value(init_value)

    def set_value(self, value):
        assert self.min_value <= value <= self.max_value
        self.raw_value = (value - self.min_value) / (self.max_value - self.min_value)

    def get_value(self):
        return self.raw_value * (self.max_value - self.min_value) + self.min_value


_EQ_SPEC = [("low_shelf", 80.0, (20.0, 4000.0))] + [
    (f"band{i}", fc, (20.0, 10000.0)) for i, fc in enumerate((300.0, 1000.0, 3000.0, 10000.0))
] + [("high_shelf", 1000.0, (200.0, 18000.0))]


class OracleParametricEQ:
    """effects.py:800-873 (ranges at :823-840)."""

    def __init__(self):
        self.parameters = {}
        for name, fc, (flo, fhi) in _EQ_SPEC:
            self.parameters[f"{name}_gain_db"] = Parameter(0.0, -24.0, 24.0)
            self.parameters[f"{name}_cutoff_freq"] = Parameter(fc, flo, fhi)
            self.parameters[f"{name}_q_factor"] = Parameter(0.707, 0.1, 4.0)

    def process(self, x, sample_rate):
        x = _rows(x)
        p = np.array([q.get_value() for q in self.parameters.values()], dtype=np.float64)
        y = np.empty_like(x)
        for c in range(x.shape[0]):
            lib().oracle_eq(_f32p(x[c]), x.shape[1], p.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            float(sample_rate), _f32p(y[c]))
        return y


class OracleCompressor:
    """effects.py:876-897."""

    def __init__(self):
        self.parameters = {
            "threshold_db": Parameter(0.0, -80.0, 0.0),
            "ratio": Parameter(4.0, 1.0, 20.0),
            "attack_ms": Parameter(1.0, 0.1, 100.0),
            "release_ms": Parameter(100.0, 10.0, 1000.0),
        }

    def process(self, x, sample_rate):
        x = _rows(x)
        v = [np.float32(q.get_value()) for q in self.parameters.values()]
        y = np.empty_like(x)
        for c in range(x.shape[0]):  # JUCE keeps one envelope per channel
            lib().oracle_compressor(_f32p(x[c]), x.shape[1], v[0], v[1], v[2], v[3], float(sample_rate), _f32p(y[c]))
        return y


class OracleDistortion:
    """effects.py:900-914 (constructor ignores its arguments, like the reference)."""

    def __init__(self):
        self.parameters = {
            "drive_db": Parameter(0.0, -48.0, 48.0),
            "output_gain_db": Parameter(0.0, -24.0, 24.0),
        }

    def process(self, x, sample_rate):
        x = _rows(x)
        v = [np.float32(q.get_value()) for q in self.parameters.values()]
        y = np.empty_like(x)
        for c in range(x.shape[0]):
            lib().oracle_distortion(_f32p(x[c]), x.shape[1], v[0], v[1], _f32p(y[c]))
        return y


class OracleDelay:
    """effects.py:917-934."""

    def __init__(self):
        self.parameters = {
            "delay_seconds": Parameter(0.5, 0.01, 1.0),
            "feedback": Parameter(0.5, 0.05, 1.0),
            "mix": Parameter(0.5, 0.0, 1.0),
        }

    def process(self, x, sample_rate):
        x = _rows(x)
        v = [np.float32(q.get_value()) for q in self.parameters.values()]
        y = np.empty_like(x)
        for c in range(x.shape[0]):
            lib().oracle_delay(_f32p(x[c]), x.shape[1], v[0], v[1], v[2], float(sample_rate), _f32p(y[c]))
        return y


class OracleReverb:
    """effects.py:937-959: wet_level = wet_dry, dry_level = 1 - wet_dry (Python float, then f32)."""

    def __init__(self):
        self.parameters = {
            "room_size": Parameter(0.5, 0.0, 1.0),
            "damping": Parameter(0.5, 0.0, 1.0),
            "wet_dry": Parameter(0.5, 0.0, 1.0),
            "width": Parameter(0.5, 0.0, 1.0),
        }

    def process(self, x, sample_rate):
        y = _rows(x).copy()
        p = self.parameters
        wet = p["wet_dry"].get_value()
        args = [np.float32(p["room_size"].get_value()), np.float32(p["damping"].get_value()),
                np.float32(wet), np.float32(1 - wet), np.float32(p["width"].get_value())]
        if y.shape[0] == 2:
            lib().oracle_reverb(_f32p(y[0]), _f32p(y[1]), y.shape[1], 2, *args, float(sample_rate))
        else:
            lib().oracle_reverb(_f32p(y[0]), _f32p(y[0]), y.shape[1], 1, *args, float(sample_rate))
        return y


# --------------------------------------------------------------------------
# chain walk
# --------------------------------------------------------------------------
def make_plugins(kinds, channels=None, fixed=None):
    """Build a reference-style plugins dict from effect kinds, e.g. ("eq","comp","reverb")."""
    from oracle.convreverb import OracleNoiseShapedReverb, OracleNoiseShapedReverb2s
    from oracle.lticomp import OracleLTICompressor

    table = {
        "eq": ("ParametricEQ", OracleParametricEQ, 1),
        "comp": ("Compressor", OracleCompressor, 1),
        "dist": ("Distortion", OracleDistortion, 1),
        "delay": ("Delay", OracleDelay, 2),
        "reverb": ("Reverb", OracleReverb, 2),
        "convreverb": ("NoiseShapedReverb", OracleNoiseShapedReverb, 2),      # 65 536-tap IR (dasp default)
        "convreverb2s": ("NoiseShapedReverb", OracleNoiseShapedReverb2s, 2),  # 96 000-tap IR (BASELINE config 4)
        "lticomp": ("LTICompressor", OracleLTICompressor, 2),                 # stereo-linked, look-ahead 512
        "lticomp1": ("LTICompressor", OracleLTICompressor, 1),                # every channel on its own
    }
    plugins = {}
    for i, k in enumerate(kinds):
        name, cls, ch = table[k]
        plugins[name if name not in plugins else f"{name}{i}"] = {
            "class_path": cls,
            "num_params": None,
            "num_channels": ch if channels is None else channels[i],
            "fixed_parameters": {} if fixed is None else dict(fixed[i]),
        }
    return plugins


def load_plugins(plugins: dict):
    """style_transfer.py:17-42: instantiate, prepend the dead ``our_bypass`` slot, count."""
    total, init = 0, []
    for entry in plugins.values():
        if "class_path" not in entry:
            raise ValueError("Plugin must contain 'vst_filepath' or 'class_path'.")
        inst = entry["class_path"]()
        names = ["our_bypass"] + list(inst.parameters.keys())
        init += [0.0] + [q.raw_value for q in inst.parameters.values()]
        entry["parameter_names"] = names
        entry["num_params"] = len(names)
        entry["instance"] = inst
        total += len(names)
    return plugins, total, init


def _assign(entry, w, widx):
    """Parameter loop of style_transfer.py:76-92.  ``our_bypass`` only consumes a slot."""
    inst = entry["instance"]
    for name in entry["parameter_names"]:
        if name != "our_bypass":
            if name in entry["fixed_parameters"]:
                inst.parameters[name].set_value(entry["fixed_parameters"][name])
            else:
                inst.parameters[name].raw_value = w[widx]
        widx += 1
    return widx


def process_audio(x: np.ndarray, w, sr, plugins: dict, normalize_stages: bool = False) -> np.ndarray:
    """style_transfer.py:45-115."""
    x = np.array(x, dtype=np.float32, copy=True)
    widx = 0
    for entry in plugins.values():
        if "instance" not in entry:
            entry["instance"] = entry["class_path"]()
        widx = _assign(entry, w, widx)
        if entry["num_channels"] == 2 and x.shape[0] == 1:
            x = np.concatenate((x, x), axis=0)
        if entry["num_channels"] == 1 and x.shape[0] == 2:
            x = np.concatenate([entry["instance"].process(x[c:c + 1], sr) for c in (0, 1)], axis=0)
        else:
            x = entry["instance"].process(x, sr)
        if normalize_stages:
            x = x / np.clip(np.max(np.abs(x)), 1e-8, None)
    x = np.ascontiguousarray(x, dtype=np.float32)
    lib().oracle_peak_normalize(_f32p(x), x.size)
    return x


def parameters_to_dict(w, plugins: dict) -> dict:
    """style_transfer.py:324-359."""
    out, widx = {}, 0
    for pname, entry in plugins.items():
        d = out.setdefault(pname, {})
        start = widx
        widx = _assign(entry, w, widx)
        for k, name in enumerate(entry["parameter_names"]):
            if name == "our_bypass":
                d[name] = w[start + k]
            else:
                d[name] = entry["instance"].parameters[name].get_value()
    return out
