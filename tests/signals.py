"""Seeded synthetic audio shared by the golden generator, the tests and bench.py."""
import numpy as np

SR = 48000


def test_signal(chs: int, L: int, seed: int = 0) -> np.ndarray:
    """0.1*noise mixed 50/50 with a 5-tone chord (110*2^k Hz) plus a decaying click every 0.5 s.

    Stereo channels differ (decorrelated noise, detuned chord) so mid AND side carry signal.
    Returns float32 [chs, L] with peak < 1.
    """
    rng = np.random.RandomState(seed)
    t = np.arange(L) / SR
    out = np.zeros((chs, L))
    for c in range(chs):
        noise = 0.1 * rng.randn(L)
        chord = sum(np.sin(2 * np.pi * 110.0 * (2 ** k) * (1 + 0.003 * c) * t + 0.7 * k + c) for k in range(5)) / 5
        clicks = np.zeros(L)
        for s in range(0, L, SR // 2):
            n = min(200, L - s)
            clicks[s:s + n] += 0.8 * np.exp(-np.arange(n) / 20.0) * (1 if (s // (SR // 2) + c) % 2 == 0 else -1)
        out[c] = 0.5 * noise + 0.5 * (0.6 * chord + clicks)
    out *= 0.7 / np.abs(out).max()
    return out.astype(np.float32)


test_signal.__test__ = False  # not a pytest test


def eq_corner_vectors(D: int):
    """Parameter-box corners for the 18-parameter EQ (+ optional leading our_bypass slot).

    Includes SURVEY Appendix E's worst case: every cutoff at its minimum, +24 dB, Q = 4.
    """
    lead = [0.0] * (D - 18)
    worst = lead + [1.0, 0.0, 1.0] * 6
    all0 = lead + [0.0] * 18
    all1 = [1.0] * D
    cut_hi = lead + [0.0, 1.0, 0.0] * 6
    return [np.array(v, dtype=np.float64) for v in (worst, all0, all1, cut_hi)]
