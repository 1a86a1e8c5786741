"""CPU tests of the noise-shaped convolution reverb: properties of the oracle restatement
(oracle/convreverb.py -- parity unpinned upstream, dasp-pytorch is absent) and bit-level agreement of the host-side setup
pieces libstito restates in C++ (seeded white noise, scipy.signal.firwin filter bank) with the oracle's."""
import numpy as np
import pytest

from tests.signals import test_signal


def test_white_noise_is_deterministic_and_gaussian():
    from oracle import convreverb as cr

    a, b, c = cr.white_noise(3, 200000), cr.white_noise(3, 200000), cr.white_noise(4, 200000)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1.0) < 0.01 and 3.5 < np.abs(a).max() < 6.0
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.01 and np.array_equal(cr.white_noise(3, 1000), a[:1000])


def test_libstito_host_setup_matches_the_oracle():
    """The C++ restatements (cma-free, GPU-free): same noise bits, filter bank equal to scipy's firwin after the float32 cast."""
    from oracle import convreverb as cr
    from st_ito_b200 import _lib

    L = _lib.lib()
    n = 50001
    got = np.empty(n, dtype=np.float32)
    assert L.stito_crv_host_noise(7, n, got.ctypes.data) == 0
    want = cr.white_noise(7, n)
    assert (got == want).mean() > 0.9999 and np.abs(got - want).max() < 1e-6  # libm vs numpy log/cos: last-bit ties only
    for sr in (48000.0, 44100.0):
        fb = np.empty((12, 1023), dtype=np.float32)
        assert L.stito_crv_host_filterbank(sr, fb.ctypes.data) == 0
        ref = cr.octave_band_filterbank(1023, sr)
        assert np.abs(fb - ref).max() <= 2e-9 and (fb == ref).mean() > 0.99
        assert abs(fb[0].sum() - 1.0) < 1e-5  # low-pass: unit DC gain (scale=True)
    assert L.stito_crv_host_filterbank(30000.0, fb.ctypes.data) < 0  # the 18 kHz band needs fs > 36.1 kHz


def test_oracle_reverb_properties():
    from oracle import convreverb as cr

    ref = cr.OracleNoiseShapedReverb(num_samples=4096, seed=1)
    assert list(ref.parameters)[:2] == ["band0_gain", "band1_gain"] and list(ref.parameters)[-1] == "mix" and len(ref.parameters) == 25
    x = test_signal(2, 30000, seed=5)
    ref.parameters["mix"].raw_value = 0.0
    np.testing.assert_array_equal(ref.process(x, 48000), x)  # dry only
    mono = ref.process(x[:1], 48000)
    assert mono.shape == (2, 30000) and np.array_equal(mono[0], x[0]) and np.array_equal(mono[1], x[0])  # up-mix
    ref.parameters["mix"].raw_value = 1.0
    y1, y2 = ref.process(x, 48000), ref.process(2.0 * x, 48000)
    np.testing.assert_allclose(y2, 2.0 * y1, rtol=0, atol=2e-6 * np.abs(y2).max())  # linear in the input
    imp = np.zeros((2, 6000), dtype=np.float32)
    imp[:, 0] = 1.0
    ir = cr.impulse_response(ref._bands[48000.0], [q.get_value() for q in list(ref.parameters.values())[:12]],
                             [q.get_value() for q in list(ref.parameters.values())[12:24]])
    out = ref.process(imp, 48000)
    np.testing.assert_allclose(out[:, :4096], ir, rtol=0, atol=1e-6)  # impulse in -> impulse response out
    assert np.abs(out[:, 4096:]).max() < 1e-6
    with pytest.raises(AssertionError):
        ref.parameters["mix"].set_value(1.5)


def test_float32_torch_restatement_of_the_upstream_function_agrees():
    """dasp_pytorch.noise_shaped_reverberation as recalled, written the way upstream writes it (torch float32: grouped
    conv1d of the noise with the filter bank, envelope, mean over bands, direct conv1d with the flipped impulse response),
    fed the SAME white noise as the oracle, next to the oracle (float64 FFT convolutions): bounds what "parity unpinned"
    costs numerically for this effect when the noise is given."""
    import torch

    from oracle import convreverb as cr

    sr, num_samples, taps, L = 48000.0, 6000, cr.NUM_TAPS, 9000
    rng = np.random.RandomState(11)
    x = test_signal(2, L, seed=3)
    x = (x / np.abs(x).max()).astype(np.float32)
    span = num_samples + taps - 1
    wn = torch.from_numpy(cr.white_noise(21, 2 * cr.NUM_BANDS * span).reshape(2, cr.NUM_BANDS, span).copy())
    filters = torch.from_numpy(cr.octave_band_filterbank(taps, sr)).unsqueeze(1)           # [12, 1, taps]
    bands = cr.filtered_noise_bands(sr, num_samples, 21)
    for _ in range(3):
        gains, decays, mix = rng.rand(12).astype(np.float32), rng.rand(12).astype(np.float32), np.float32(rng.rand())
        xt = torch.from_numpy(x[None].copy())                                              # [1, 2, L]
        wn_filt = torch.nn.functional.conv1d(wn, filters, groups=cr.NUM_BANDS).view(1, 2, cr.NUM_BANDS, num_samples)
        t = torch.linspace(0, 1, steps=num_samples)
        band_decays = torch.from_numpy(decays).view(1, 1, -1, 1) * 10.0 + 1.0
        env = torch.exp(-band_decays * t.view(1, 1, 1, -1))
        wn_filt = wn_filt * (env * torch.from_numpy(gains).view(1, 1, -1, 1))
        ir = wn_filt.mean(2, keepdim=True)[0]                                              # [2, 1, num_samples]
        x_pad = torch.nn.functional.pad(xt, (num_samples - 1, 0))
        y = torch.nn.functional.conv1d(x_pad, torch.flip(ir, dims=[-1]), groups=2)
        want = ((1 - float(mix)) * xt + float(mix) * y)[0].numpy()
        got = cr.noise_shaped_reverberation(x, sr, gains, decays, mix, bands)
        peak = np.abs(want).max()
        assert np.abs(got - want).max() <= 2e-5 * peak, np.abs(got - want).max() / peak
