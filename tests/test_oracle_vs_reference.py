"""Live pinning of the oracle: where the reference tree is present (the build container; never the GPU box) the
reference's OWN code is executed -- biqaud / parametric_eq / BasicParametricEQ / load_plugins / process_audio /
parameters_to_dict and the Cnn14 body -- and compared with the oracle on fresh seeds (the committed fixtures under
tests/golden/ were produced the same way, tests/golden/make_golden.py).  Skipped when /root/reference is absent."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import ref_import
from tests.signals import eq_corner_vectors, test_signal

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")

SR = 48000


@pytest.fixture(scope="module")
def reference():
    return ref_import.load()


def test_eq_chain_walk_and_waveform_bit_exact(reference):
    effects, st, _ = reference
    from oracle import dsp

    dsp.build()
    with contextlib.redirect_stdout(io.StringIO()):
        rp, D, rinit = st.load_plugins({"ParametricEQ": {"class_path": effects.BasicParametricEQ, "num_params": None,
                                                         "num_channels": 1, "fixed_parameters": {}}})
        op, oD, oinit = dsp.load_plugins(dsp.make_plugins(["eq"]))
    assert D == oD == 19 and list(rinit) == pytest.approx(list(oinit))
    rng = np.random.RandomState(4242)
    for chs, L in ((1, 30011), (2, 65536)):
        x = test_signal(chs, L, seed=L)
        for w in [rng.rand(D) for _ in range(3)] + eq_corner_vectors(D)[:2]:
            want = st.process_audio(x.copy(), w, SR, rp)
            got = dsp.process_audio(x.copy(), w, SR, op)
            assert want.dtype == got.dtype == np.float32
            np.testing.assert_array_equal(got, want)  # same fp64 operation order as scipy.signal.lfilter
            assert st.parameters_to_dict(w, rp)["ParametricEQ"] == pytest.approx(dsp.parameters_to_dict(w, op)["ParametricEQ"])


def test_biquad_coefficients_bit_exact(reference):
    effects, _, _ = reference
    from oracle import dsp

    rng = np.random.RandomState(7)
    for kind in ("low_shelf", "peaking", "high_shelf"):
        for _ in range(50):
            g, fc, q = rng.uniform(-24, 24), rng.uniform(20, 18000), rng.uniform(0.1, 4)
            b, a = effects.biqaud(g, fc, q, SR, kind)
            ob, oa = dsp.biquad_coefs(g, fc, q, SR, kind)
            np.testing.assert_array_equal(np.asarray(ob), np.asarray(b))
            np.testing.assert_array_equal(np.asarray(oa), np.asarray(a))


@pytest.mark.parametrize("chs", [1, 2])
def test_cnn14_body_matches_reference_module(reference, chs):
    _, _, panns = reference
    from oracle import cnn14

    ora = cnn14.make_encoder(seed=21, bn_stats=True)
    ref = panns.Cnn14(**cnn14.AFX_REP_ARGS).eval()
    ref.load_state_dict(ora.state_dict())
    x = torch.from_numpy(np.stack([test_signal(chs, 36000, seed=9 + b) for b in range(2)]))
    x = x / x.abs().amax(dim=(1, 2), keepdim=True)
    with torch.no_grad():
        rm, rs = ref(x)
        om, os_ = ora(x)
    assert torch.allclose(om, rm, rtol=1e-5, atol=1e-6) and torch.allclose(os_, rs, rtol=1e-5, atol=1e-6)
