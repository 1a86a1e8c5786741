"""The host loop (run_es / evaluate / process_audio) on its GENERIC path -- arbitrary plugin objects and
an arbitrary embedding function, no GPU -- (a) against invariants of the reference's loop, and (b) where the reference
tree is present, against the reference's OWN run_es executed live with the same CMA-ES class and seeds: identical
parameter vectors, fitness histories and output audio."""
import contextlib
import io
import sys

import numpy as np
import pytest
import torch

from oracle import ref_import

SR = 48000


class _P:  # the duck-typed parameter object of the plugin protocol (effects.py:784-797)
    def __init__(self, init, lo, hi):
        self.min_value, self.max_value = lo, hi
        self.raw_value = (init - lo) / (hi - lo)

    def get_value(self):
        return self.raw_value * (self.max_value - self.min_value) + self.min_value

    def set_value(self, v):
        assert self.min_value <= v <= self.max_value
        self.raw_value = (v - self.min_value) / (self.max_value - self.min_value)


class ToyTilt:
    """1-channel user plugin: gain + one-pole tilt; cheap, deterministic, sensitive to both parameters."""
    seen_lengths = []

    def __init__(self):
        self.parameters = {"gain_db": _P(0.0, -12.0, 12.0), "tilt": _P(0.5, 0.0, 0.95)}

    def process(self, x, sample_rate):
        ToyTilt.seen_lengths.append(x.shape[-1])
        from scipy.signal import lfilter

        g = 10.0 ** (self.parameters["gain_db"].get_value() / 20.0)
        a = self.parameters["tilt"].get_value()
        return (g * lfilter([1.0 - a], [1.0, -a], x.astype(np.float64), axis=-1)).astype(np.float32)


def toy_embed(x, model, sample_rate):
    """Embedding function of the run_es protocol: {name: [bs, E]}; band energies of a crude spectrum."""
    spec = torch.fft.rfft(x.float(), dim=-1).abs()
    bands = torch.stack([c.mean(dim=-1) for c in torch.chunk(spec, 16, dim=-1)], dim=-1)  # [bs, chs, 16]
    feats = torch.log1p(bands).flatten(1)
    return {"mid": torch.nn.functional.normalize(feats, dim=-1), "side": torch.nn.functional.normalize(feats ** 2, dim=-1)}


def _plugins(cls):
    return {"Tilt": {"class_path": cls, "num_params": None, "num_channels": 1, "fixed_parameters": {}}}


def _signals(L=600):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, L, generator=g)
    t = torch.cumsum(torch.randn(1, 2, L, generator=g), dim=-1) * 0.1
    return x, t


def _run(module, cma_mod, seed, **kw):
    """run_es of `module` with numpy's global RNG and the CMA-ES seeded identically."""
    kw = dict(kw)
    class SeededES(cma_mod.CMAEvolutionStrategy):
        def __init__(self, x0, sigma0, opts=None):
            o = dict(opts or {})
            o.setdefault("seed", seed)
            o.setdefault("verbose", -9)
            super().__init__(x0, sigma0, o)

    shim = type(sys)("cma")
    shim.CMAEvolutionStrategy = SeededES
    module.cma = shim
    np.random.seed(seed)
    ToyTilt.seen_lengths = []
    with contextlib.redirect_stdout(io.StringIO()):
        plugins, D, _ = module.load_plugins(_plugins(ToyTilt))
        x, t = _signals(kw.pop("L", 600))
        res = module.run_es(x, t, SR, plugins, None, toy_embed, **kw)
    return res, x, t, list(ToyTilt.seen_lengths), D


def test_generic_run_es_invariants():
    from st_ito_b200 import cma, style_transfer

    res, x, t, lengths, D = _run(style_transfer, cma, 3, max_iters=14, popsize=6, sigma0=0.33, find_w0=True)
    assert D == 3  # our_bypass + 2
    assert set(res) == {"output_audio", "params", "fopt", "wopt", "fval_history", "wopt_history"}
    assert float(x.abs().max()) == 1.0 and float(t.abs().max()) == pytest.approx(1.0)  # in-place peak normalisation
    n_it = len(res["fval_history"])
    assert 1 <= n_it <= 14 and len(res["wopt_history"]) == n_it
    assert res["wopt_history"][0] is None  # histories are appended BEFORE tell (style_transfer.py:639-640)
    # every evaluate() call zero-pads the 600-sample input to 262144 (:518); L and R are separate mono passes (:98-102);
    # the final render runs on the un-padded input (:675-678)
    assert lengths[-2:] == [600, 600] and set(lengths[:-2]) == {262144}
    assert len(lengths) == 2 * 6 * (n_it + 1) + 2
    assert res["output_audio"].shape == (2, 600) and float(res["output_audio"].abs().max()) == pytest.approx(1.0)
    assert np.all((res["wopt"] >= 0) & (res["wopt"] <= 1)) and -1.0 <= res["fopt"] <= 1.0
    assert list(res["params"]["Tilt"]) == ["our_bypass", "gain_db", "tilt"]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kw", [dict(max_iters=6, popsize=6, sigma0=0.33, find_w0=True),
                                dict(max_iters=16, popsize=4, sigma0=0.05, find_w0=False),
                                dict(max_iters=3, popsize=4, sigma0=0.2, find_w0=True, random_crop=True, L=300000),
                                dict(max_iters=2, popsize=4, sigma0=0.2, find_w0=False, parallel=False, L=270000)])
def test_generic_run_es_matches_the_reference_loop(kw):
    """Same plugins, same embedding function, same CMA-ES class and seeds through the reference's run_es and ours."""
    from st_ito_b200 import cma, style_transfer

    _, ref_st, _ = ref_import.load()
    ours, _, _, lo, _ = _run(style_transfer, cma, 11, **kw)
    ref, _, _, lr, _ = _run(ref_st, cma, 11, **kw)
    assert lo == lr  # identical sequence of plugin.process calls (lengths): same padding / channel policy
    assert len(ours["fval_history"]) == len(ref["fval_history"])  # same early-stopping decision
    np.testing.assert_array_equal(np.asarray(ours["wopt"]), np.asarray(ref["wopt"]))
    assert ours["fopt"] == ref["fopt"]
    assert [float(v) for v in ours["fval_history"][1:]] == [float(v) for v in ref["fval_history"][1:]]
    for a, b in zip(ours["wopt_history"], ref["wopt_history"]):
        assert (a is None and b is None) or np.array_equal(np.asarray(a), np.asarray(b))
    np.testing.assert_array_equal(ours["output_audio"].numpy(), ref["output_audio"].numpy())
    assert ours["params"] == ref["params"]
