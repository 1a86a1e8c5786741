"""Pins oracle/ to the reference: every committed golden vector was produced by
EXECUTING the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn14, dsp, frontend
from tests.signals import test_signal

SR = 48000


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_biquad_coefficients_match_reference(golden_dir):
    tab = _load(golden_dir, "biquad.npz")["table"]
    kinds = ("low_shelf", "peaking", "high_shelf")
    for row in tab:
        b, a = dsp.biquad_coefs(row[0], row[1], row[2], SR, kinds[int(row[3])])
        ref = row[4:10]
        got = np.concatenate([b, a])
        # same formulas, same fp64; numpy's sin/cos and libm's may differ in the last ulp
        np.testing.assert_allclose(got, ref, rtol=4e-15, atol=1e-300)


@pytest.mark.parametrize("chs", [1, 2])
@pytest.mark.parametrize("L", [4096, 262144, 480000])
def test_eq_process_audio_matches_reference(golden_dir, chs, L):
    g = _load(golden_dir, "eq.npz")
    plugins, D, init = dsp.load_plugins(dsp.make_plugins(["eq"]))
    assert D == int(g["D"]) == 19
    np.testing.assert_array_equal(np.array(init), g["init"])
    x = test_signal(chs, L, seed=L + chs)
    for i, w in enumerate(g["W"]):
        y = dsp.process_audio(x, w, SR, plugins)
        assert y.dtype == np.float32 and y.shape == (chs, L)
        assert np.abs(y).max() == 1.0
        if L == 4096:
            ref = g[f"y_{chs}_{L}"][i]
            np.testing.assert_allclose(y, ref, rtol=0, atol=2e-7)
            assert (y == ref).mean() > 0.999
        else:
            np.testing.assert_allclose(y[:, ::997], g[f"ys_{chs}_{L}"][i], rtol=0, atol=2e-7)
            np.testing.assert_allclose(y[:, :512], g[f"yh_{chs}_{L}"][i], rtol=0, atol=2e-7)
            np.testing.assert_allclose((y.astype(np.float64) ** 2).sum(-1), g[f"ye_{chs}_{L}"][i], rtol=1e-9)


def test_parameters_to_dict_matches_reference(golden_dir):
    g = _load(golden_dir, "eq.npz")
    plugins, _, _ = dsp.load_plugins(dsp.make_plugins(["eq"]))
    d = dsp.parameters_to_dict(g["W"][0], plugins)["ParametricEQ"]
    np.testing.assert_array_equal(np.array(list(d.values())), g["param_dict_last"])
    assert list(d.keys())[0] == "our_bypass"


@pytest.mark.parametrize("tag,bn", [("plain", False), ("bnstats", True)])
@pytest.mark.parametrize("chs", [1, 2])
def test_cnn14_matches_reference(golden_dir, tag, bn, chs):
    g = _load(golden_dir, "cnn14.npz")
    model = cnn14.make_encoder(seed=3, bn_stats=bn)
    x = torch.from_numpy(np.stack([test_signal(chs, 40000, seed=100 + b) for b in range(2)]))
    x = x / x.abs().amax(dim=(1, 2), keepdim=True)
    with torch.no_grad():
        mid, side = model(x)
    for got, ref in ((mid, g[f"{tag}_mid_{chs}"]), (side, g[f"{tag}_side_{chs}"])):
        ref = torch.from_numpy(ref)
        assert ((got - ref).norm() / ref.norm()).item() < 1e-6


def test_fitness_and_ranking_match_reference(golden_dir):
    g = _load(golden_dir, "fitness.npz")
    plugins, D, _ = dsp.load_plugins(dsp.make_plugins(["eq"]))
    model = cnn14.make_encoder(seed=3, bn_stats=True, conv_gain=2.0)
    cnn14.centre_heads(model)
    np.testing.assert_allclose(model.fc_mid.bias.detach().numpy(), g["bias_mid"], atol=2e-5)
    assert np.ptp(g["fitness"]) > 0.3 and np.diff(np.sort(g["fitness"])).min() > 1e-3  # well-conditioned ranking
    x = test_signal(2, 40000, seed=5)
    x = x / np.abs(x).max()
    tgt = dsp.process_audio(x, g["w_star"], SR, plugins)
    outs = np.stack([dsp.process_audio(x, w, SR, plugins) for w in g["W"]])
    te = cnn14.get_param_embeds(torch.from_numpy(tgt[None].copy()), model, SR)
    oe = cnn14.get_param_embeds(torch.from_numpy(outs.copy()), model, SR)
    f = cnn14.fitness(oe, te).numpy()
    np.testing.assert_allclose(f, g["fitness"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(np.argsort(f, kind="stable"), g["argsort"])
    assert int(np.argmin(f)) == int(g["argsort"][0])
    np.testing.assert_allclose(oe["mid"].numpy(), g["mid"], atol=2e-5)


def test_frontend_against_independent_stft_and_mel():
    """torchlibrosa restatement vs torch.stft (fp64) and torchaudio's Slaney filterbank."""
    import torchaudio

    m = cnn14.make_encoder(seed=0, bn_stats=False)
    fb = torchaudio.functional.melscale_fbanks(1025, 20.0, 20000.0, 128, SR, norm="slaney", mel_scale="slaney")
    assert (fb - m.logmel_extractor.melW).abs().max().item() < 2e-7
    x = torch.from_numpy(test_signal(2, 40000, seed=1))[None]
    x = x / x.abs().max()
    with torch.no_grad():
        lm = m.logmel(x)
    xs = torch.stack([(x[0, 0] + x[0, 1]) / 2, (x[0, 0] - x[0, 1]) / 2]).double()
    win = torch.from_numpy(frontend.hann_periodic(2048))
    S = torch.stft(xs, 2048, 1024, 2048, win, center=True, pad_mode="reflect", return_complex=True).abs() ** 2
    mel = S.transpose(1, 2) @ m.logmel_extractor.melW.double()
    ref = frontend.minmax_norm(10 * torch.log10(mel.clamp(min=1e-10)))
    assert lm.shape == (2, 1, 40000 // 1024 + 1, 128)
    assert (lm[:, 0] - ref).abs().max().item() < 2e-5


def test_compressor_and_reverb_basic_properties():
    """Unpinned restatements: sanity properties only (unity below threshold, dry-only reverb)."""
    x = test_signal(2, 20000, seed=3)
    comp = dsp.OracleCompressor()
    comp.parameters["threshold_db"].raw_value = 1.0  # 0 dB: peak 0.7 never exceeds it
    np.testing.assert_array_equal(comp.process(x, SR), x)
    comp.parameters["threshold_db"].raw_value = 0.5  # -40 dB
    y = comp.process(x, SR)
    assert np.abs(y).max() < np.abs(x).max() and np.all(np.abs(y) <= np.abs(x) + 1e-7)
    rev = dsp.OracleReverb()
    rev.parameters["wet_dry"].raw_value = 0.0
    np.testing.assert_allclose(rev.process(x, SR), 2.0 * x, atol=1e-6)  # dry gain = 2*dry_level
    rev.parameters["wet_dry"].raw_value = 1.0
    yw = rev.process(x, SR)
    assert np.abs(yw[:, :1000]).max() < 1e-3 and np.abs(yw).max() > 1e-3  # first comb tap at ~1214 samples
