"""Import the reference's own modules so they can be EXECUTED as the ground truth.

Only usable where ``/root/reference`` exists (the build container); used by
``tests/golden/make_golden.py`` to generate committed fixtures and by the
optional ``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).
Nothing is copied: the reference sources are imported from where they lie.

The reference imports packages that are not installed here (pedalboard,
dasp_pytorch, pyloudnorm, cma, torchlibrosa).  They are replaced by inert stubs
in ``sys.modules``; the code paths that run -- ``biqaud``, ``parametric_eq``,
``BasicParametricEQ``, ``Parameter``, ``load_plugins``, ``process_audio``,
``parameters_to_dict`` and the ``Cnn14`` body -- never touch them.  For
``torchlibrosa`` the stub classes are this repo's restatement of the front-end
(oracle/frontend.py), registered under the parameter names torchlibrosa uses.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STITO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "st_ito", "style_transfer.py"))


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    return sys.modules[name]


def install_stubs():
    class _Absent:
        def __init__(self, *a, **k):
            raise RuntimeError("third-party effect not available in the oracle harness")

    _stub("pedalboard", **{n: _Absent for n in (
        "Pedalboard", "Gain", "Chorus", "Reverb", "Compressor", "Phaser", "Delay", "Distortion",
        "Limiter")}, load_plugin=_Absent)
    _stub("dasp_pytorch")
    _stub("pyloudnorm")
    _stub("cma")

    from oracle import frontend

    class SpecAugmentation:  # identity in eval mode; never called by the path
        def __init__(self, *a, **k):
            pass

        def __call__(self, x):
            return x

    tl = _stub("torchlibrosa")
    _stub("torchlibrosa.stft", Spectrogram=frontend.Spectrogram, LogmelFilterBank=frontend.LogmelFilterBank)
    _stub("torchlibrosa.augmentation", SpecAugmentation=SpecAugmentation)
    tl.stft = sys.modules["torchlibrosa.stft"]
    tl.augmentation = sys.modules["torchlibrosa.augmentation"]


def load():
    """Return (effects, style_transfer, panns) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    effects = importlib.import_module("st_ito.effects")
    style_transfer = importlib.import_module("st_ito.style_transfer")
    panns = importlib.import_module("st_ito.models.panns")
    return effects, style_transfer, panns
