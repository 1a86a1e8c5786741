"""ctypes binding of libstito.so (include/stito.h).

This is the reference-side stub a maintainer of st-ito would add: plain
pointers and sizes, no torch types.  The library is built in-tree (``st_ito_b200/libstito.so``)
by ``st_ito_b200/csrc/Makefile``; there is NO CPU fallback -- if the library is missing or no
B200 is present the product path raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstito.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_FX = 8
MAX_FX_PARAMS = 32

FX_EQ, FX_COMPRESSOR, FX_DISTORTION, FX_DELAY, FX_REVERB, FX_CONV_REVERB, FX_LTI_COMPRESSOR = range(7)


class FxDesc(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("num_channels", c_int32),
        ("num_params", c_int32),
        ("w_index", c_int32 * MAX_FX_PARAMS),
        ("fixed_raw", c_double * MAX_FX_PARAMS),
        ("iopt", c_int32 * 4),
    ]


class ChainDesc(Structure):
    _fields_ = [
        ("num_fx", c_int32),
        ("num_w", c_int32),
        ("normalize_stages", c_int32),
        ("reserved", c_int32),
        ("sample_rate", c_double),
        ("fx", FxDesc * MAX_FX),
    ]


_f32p = POINTER(c_float)


class EncoderWeights(Structure):
    _fields_ = [
        ("n_fft", c_int32), ("hop", c_int32), ("n_mels", c_int32), ("embed_dim", c_int32),
        ("bn_eps", c_float), ("reserved", c_int32),
        ("conv_w", _f32p * 12),
        ("bn_weight", _f32p * 12), ("bn_bias", _f32p * 12), ("bn_mean", _f32p * 12), ("bn_var", _f32p * 12),
        ("fc_mid_w", _f32p), ("fc_mid_b", _f32p), ("fc_side_w", _f32p), ("fc_side_b", _f32p),
        ("mel_w", _f32p),
    ]


class Timing(Structure):
    _fields_ = [
        ("ms_dsp", c_float), ("ms_frontend", c_float), ("ms_encoder", c_float), ("ms_fitness", c_float),
        ("ms_total", c_float),
        ("launches", c_int32), ("precision", c_int32),
        ("encoder_flop", c_double), ("dsp_bytes", c_double), ("frontend_bytes", c_double),
        ("ms_conv", c_float * 12),
        ("comp_fallbacks", c_int32), ("act_overflow", c_int32),
    ]


class StitoError(RuntimeError):
    """A libstito call failed; ``code`` is the negative STITO_E* value."""

    def __init__(self, code, message):
        super().__init__(f"libstito error {code}: {message}")
        self.code = code


EXPORTS = {
    # name: (restype, argtypes)
    "stito_create": (c_int, [POINTER(ChainDesc), POINTER(EncoderWeights), c_int, POINTER(c_void_p)]),
    "stito_destroy": (None, [c_void_p]),
    "stito_set_chain": (c_int, [c_void_p, POINTER(ChainDesc)]),
    "stito_set_precision": (c_int, [c_void_p, c_int]),
    "stito_set_input": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64]),
    "stito_set_target": (c_int, [c_void_p, c_void_p, c_int, c_int64]),
    "stito_set_target_embeds": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "stito_eval_population": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "stito_gather_export": (c_int, [c_void_p, c_int, c_void_p]),
    "stito_gather_attach": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "stito_eval_population_gather": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_int, c_void_p,
                                             c_void_p]),
    "stito_process": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_void_p,
                              c_void_p]),
    "stito_out_channels": (c_int, [c_void_p, c_int]),
    "stito_embed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "stito_logmel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "stito_get_timing": (c_int, [c_void_p, POINTER(Timing)]),
    "stito_crv_host_filterbank": (c_int, [c_double, c_void_p]),
    "stito_crv_host_noise": (c_int, [c_uint64, c_int64, c_void_p]),
    "stito_lticomp_host_design": (c_int, [c_double, c_int64, c_float, c_void_p]),
    "stito_cma_create": (c_int, [c_void_p, c_int, c_double, c_int, c_double, c_double, c_uint64, POINTER(c_void_p)]),
    "stito_cma_destroy": (None, [c_void_p]),
    "stito_cma_eig": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "stito_cma_ask": (c_int, [c_void_p, c_void_p]),
    "stito_cma_geno": (c_int, [c_void_p, c_void_p]),
    "stito_cma_tell": (c_int, [c_void_p, c_void_p, c_void_p]),
    "stito_cma_result": (c_int, [c_void_p] + [c_void_p] * 11),
    "stito_last_error": (c_char_p, []),
    "stito_version": (c_int, []),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile libstito.so for sm_100a with nvcc (in-tree; idempotent via make)."""
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """Load libstito.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make -C {CSRC}` (or `python -c 'import "
                "__graft_entry__ as g; g.build()'`).  st_ito_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().stito_last_error()
        raise StitoError(rc, msg.decode() if msg else "")
    return rc


def ptr(t):
    """Raw data pointer of a torch tensor / numpy array (host or device), as c_void_p."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)


__all__ = ["lib", "build", "check", "ptr", "ChainDesc", "FxDesc", "EncoderWeights", "Timing", "StitoError",
           "byref", "EXPORTS", "LIB_PATH"]
