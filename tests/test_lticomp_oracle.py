"""CPU tests of the compressor with LTI gain smoothing (second half): properties of the oracle
restatement (oracle/lticomp.py -- parity unpinned upstream, dasp-pytorch is absent) and agreement of the host-side filter
design libstito restates in C++ with the oracle's."""
import numpy as np
import pytest

from tests.signals import test_signal

SR = 48000


@pytest.mark.parametrize("L,attack_ms", [(300, 250.0), (4096, 250.0), (5000, 100.0), (48000, 0.1), (48000, 250.0)])
def test_frequency_sampling_equals_recursion_plus_wraparound(L, attack_ms):
    """lfilter_via_fsm is a circular convolution: the recursion started from the periodic steady state, exactly."""
    from scipy.signal import lfilter

    from oracle import lticomp as lc

    g = -10.0 * np.abs(np.random.RandomState(L).randn(L)).astype(np.float32)
    a = lc.attack_alpha(attack_ms, SR)
    fsm, rec = lc.smooth_gain_fsm(g, a), lc.smooth_gain_recursive(g, a)
    np.testing.assert_allclose(rec, fsm, rtol=0, atol=1e-10)
    y0 = lfilter([float(np.float32(1) - a)], [1.0, -float(a)], g.astype(np.float64))
    wrap = np.abs(fsm - y0).max()
    assert (wrap > 1.0) if (L <= 4096 and attack_ms == 250.0) else (wrap < 0.05)  # short clip + long attack: not negligible


def test_libstito_filter_design_matches_the_oracle():
    from oracle import lticomp as lc
    from st_ito_b200 import _lib

    lib = _lib.lib()
    out = np.empty(3, dtype=np.float64)
    for sr in (48000.0, 44100.0):
        for L in (300, 4096, 50001, 480000):
            for att in (0.1, 0.37, 1.0, 10.0, 99.9, 250.0):
                assert lib.stito_lticomp_host_design(sr, L, att, out.ctypes.data) == 0
                a = lc.attack_alpha(np.float32(att), sr)
                assert out[0] == float(a) and out[1] == float(np.float32(1.0) - a)
                n_fft = lc.fsm_fft_size(L)
                want = np.exp(np.log(float(a)) * (n_fft - L)) / (-np.expm1(np.log(float(a)) * n_fft))
                assert abs(out[2] - want) <= 1e-12 * max(want, 1e-300) + 1e-300
    assert lib.stito_lticomp_host_design(48000.0, 0, 1.0, out.ctypes.data) < 0


def test_oracle_compressor_properties():
    from oracle import lticomp as lc

    L = 30000
    x = test_signal(2, L, seed=4)
    x = (0.05 * x / np.abs(x).max()).astype(np.float32)
    # far below the threshold: unit gain curve, only make-up gain and the look-ahead delay remain
    y = lc.lti_compressor(x, SR, threshold_db=0.0, ratio=8.0, attack_ms=5.0, release_ms=100.0, knee_db=1.0,
                          makeup_gain_db=6.0, lookahead_samples=512)
    want = np.zeros_like(x)
    want[:, 512:] = x[:, :-512] * np.float32(10.0 ** (6.0 / 20.0))
    np.testing.assert_allclose(y, want, rtol=2e-6, atol=0)
    # ratio 1: the curve is the identity whatever the level
    y1 = lc.lti_compressor(20 * x, SR, -40.0, 1.0, 5.0, 100.0, 6.0, 0.0, 0)
    np.testing.assert_allclose(y1, 20 * x, rtol=2e-6, atol=0)
    # constant level far above the knee: the smoothed gain settles on the static curve T + (x_db - T) / R - x_db
    c = np.full((1, L), 0.5, dtype=np.float32)
    y2 = lc.lti_compressor(c, SR, -30.0, 4.0, 1.0, 100.0, 6.0, 0.0, 0)
    x_db = 20 * np.log10(0.5)
    g = -30.0 + (x_db + 30.0) / 4.0 - x_db
    assert abs(20 * np.log10(y2[0, -1] / 0.5) - g) < 1e-3
    # the side-chain is the SUM of the channels: a stereo pair of identical channels is driven 6 dB harder than mono
    m = lc.lti_compressor(c, SR, -30.0, 4.0, 1.0, 100.0, 6.0, 0.0, 0)
    s = lc.lti_compressor(np.concatenate([c, c]), SR, -30.0, 4.0, 1.0, 100.0, 6.0, 0.0, 0)
    assert np.array_equal(s[0], s[1]) and abs(20 * np.log10(s[0, -1] / m[0, -1]) + 20 * np.log10(2.0) * 0.75) < 1e-3
    # the release time is accepted and ignored (upstream behaviour)
    assert np.array_equal(lc.lti_compressor(x, SR, -40.0, 4.0, 5.0, 10.0, 6.0, 0.0, 0),
                          lc.lti_compressor(x, SR, -40.0, 4.0, 5.0, 2000.0, 6.0, 0.0, 0))


def test_plugin_wrappers_agree_on_the_parameter_schema():
    from oracle.lticomp import OracleLTICompressor
    from st_ito_b200 import effects

    a, b = effects.BasicLTICompressor(), OracleLTICompressor()
    assert list(a.parameters) == list(b.parameters) and a.lookahead_samples == b.lookahead_samples == 512
    for k in a.parameters:
        assert (a.parameters[k].min_value, a.parameters[k].max_value) == (b.parameters[k].min_value, b.parameters[k].max_value)
        assert a.parameters[k].raw_value == pytest.approx(b.parameters[k].raw_value)
    assert effects.is_native_plugin(a) and list(effects.make_chain("mastering-dasp")) == ["ParametricEQ", "LTICompressor", "NoiseShapedReverb"]


@pytest.mark.parametrize("L,attack_ms", [(3000, 250.0), (4096, 37.0), (4097, 250.0), (20000, 0.1), (50001, 120.0)])
def test_chunked_fold_and_apply_equal_the_frequency_sampled_filter(L, attack_ms):
    """The decomposition lticomp.cu uses, restated in numpy: 4096-sample chunks folded into affine maps s -> A s + B
    (A = alpha^valid), entering states from sum_j B_j * prod_{i>j} A_i with alpha^(4096 * count) for the full chunks in
    between, the wrap-around term alpha^n0 * y0[L-1] * wrap, then the recurrence inside every chunk."""
    from oracle import lticomp as lc

    CH = 4096
    g = -12.0 * np.abs(np.random.RandomState(L).randn(L)).astype(np.float32)
    alpha = lc.attack_alpha(attack_ms, SR)
    a, b0, ln_a = float(alpha), float(np.float32(1.0) - alpha), np.log(float(alpha))
    nchunks = (L + CH - 1) // CH
    n_fft = lc.fsm_fft_size(L)
    wrap = np.exp(ln_a * (n_fft - L)) / (-np.expm1(ln_a * n_fft))

    def run(seg, s):  # the recurrence over one chunk from state s; returns the states
        out = np.empty(len(seg))
        for i, v in enumerate(seg.astype(np.float64)):
            s = a * s + b0 * v
            out[i] = s
        return out

    maps = []
    for c in range(nchunks):
        seg = g[c * CH:(c + 1) * CH]
        maps.append((a ** len(seg), run(seg, 0.0)[-1]))
    last_a = maps[-1][0]
    y_end = sum(maps[j][1] * np.exp(ln_a * (nchunks - 2 - j) * CH) * last_a for j in range(nchunks - 1)) + maps[-1][1]
    y_init = y_end * wrap
    got = np.empty(L)
    for c in range(nchunks):
        s_in = sum(maps[j][1] * np.exp(ln_a * (c - 1 - j) * CH) for j in range(c)) + y_init * np.exp(ln_a * c * CH)
        got[c * CH:(c + 1) * CH] = run(g[c * CH:(c + 1) * CH], s_in)
    np.testing.assert_allclose(got, lc.smooth_gain_fsm(g, alpha), rtol=0, atol=1e-9)


def test_float32_torch_restatement_of_the_upstream_function_agrees():
    """dasp_pytorch.compressor as recalled, written the way upstream writes it (torch float32 end to end, including the
    float32 rfft / irfft of lfilter_via_fsm), next to the oracle (float32 gain computer, float64 smoothing): the two differ
    only by the float32 FFT's rounding -- this bounds how far a float32 torch run of the real thing can sit from the
    oracle, i.e. what "parity unpinned" costs numerically for this effect (measured here: 1e-7 ... 2e-5 of the peak)."""
    import torch

    from oracle import lticomp as lc

    def upstream(x, sample_rate, threshold_db, ratio, attack_ms, release_ms, knee_db, makeup_gain_db, eps=1e-8,
                 lookahead_samples=0):
        bs, chs, seq_len = x.size()
        x_side = x.sum(dim=1, keepdim=True).view(-1, 1, seq_len)
        threshold_db, ratio, attack_ms = threshold_db.view(-1, 1, 1), ratio.view(-1, 1, 1), attack_ms.view(-1, 1, 1)
        knee_db, makeup_gain_db = knee_db.view(-1, 1, 1), makeup_gain_db.view(-1, 1, 1)
        normalized_attack_time = sample_rate * (attack_ms / 1e3)
        constant = torch.tensor([9.0]).type_as(attack_ms)
        alpha_A = torch.exp(-torch.log(constant) / normalized_attack_time)
        x_db = 20 * torch.log10(torch.abs(x_side).clamp(eps))
        x_sc = x_db.clone()
        idx = torch.logical_and(x_db >= (threshold_db - (knee_db / 2)), x_db <= (threshold_db + (knee_db / 2)))
        x_sc_below = x_db + ((1 / ratio) - 1) * ((x_db - threshold_db + (knee_db / 2)) ** 2) / (2 * knee_db)
        x_sc[idx] = x_sc_below[idx]
        idx = x_db > (threshold_db + (knee_db / 2))
        x_sc_above = threshold_db + ((x_db - threshold_db) / ratio)
        x_sc[idx] = x_sc_above[idx]
        g_c = x_sc - x_db
        b = torch.cat([(1 - alpha_A), torch.zeros(bs, 1, 1)], dim=-1).squeeze(1)
        a = torch.cat([torch.ones(bs, 1, 1), -alpha_A], dim=-1).squeeze(1)
        n_fft = int(2 ** torch.ceil(torch.log2(torch.tensor(g_c.shape[-1] + g_c.shape[-1] - 1))))
        H = (torch.fft.rfft(b, n_fft) / torch.fft.rfft(a, n_fft)).unsqueeze(1)
        g_c_attack = torch.fft.irfft(torch.fft.rfft(g_c, n_fft) * H, n_fft)[..., :seq_len]
        if lookahead_samples > 0:
            x = torch.roll(x, lookahead_samples, dims=-1)
            x[:, :, :lookahead_samples] = 0
        return x * 10 ** ((g_c_attack + makeup_gain_db) / 20.0)

    rng = np.random.RandomState(5)
    for L, chs in ((20000, 2), (48000, 1)):
        x = test_signal(chs, L, seed=L)
        x = (0.7 * x / np.abs(x).max()).astype(np.float32)
        for _ in range(4):
            thr, ratio, att = -60 * rng.rand(), 1 + 19 * rng.rand(), 0.1 + 249.9 * rng.rand()
            knee, mk = 1 + 23 * rng.rand(), 24 * rng.rand()
            t = lambda v: torch.tensor([v], dtype=torch.float32)
            want = upstream(torch.from_numpy(x[None].copy()), SR, t(thr), t(ratio), t(att), t(100.0), t(knee), t(mk),
                            lookahead_samples=512)[0].numpy()
            got = lc.lti_compressor(x, SR, np.float32(thr), np.float32(ratio), np.float32(att), 100.0, np.float32(knee),
                                    np.float32(mk), 512)
            peak = np.abs(want).max()
            assert np.abs(got - want).max() <= 5e-5 * peak, (L, chs, thr, ratio, att, np.abs(got - want).max() / peak)
