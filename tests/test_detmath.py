"""tor_detmath.h (the sin/cos/pow both the kernels and the oracle's `det` mode use) against mpmath
and glibc.  The host build of the header is what is exercised here; the device build of the same
source is covered by the bit-exact GPU parity tests."""
import mpmath
import numpy as np

mpmath.mp.prec = 200


def _ulp_err(got, want_mp):
    want = float(want_mp)
    if want == 0.0:
        return abs(got)
    ulp = np.spacing(abs(want))
    return float(abs(mpmath.mpf(got) - want_mp) / mpmath.mpf(float(ulp)))


def test_sincos_below_one_ulp(oracle):
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(0, 2 * np.pi, 4000), [0.0, np.pi / 2, np.pi, 1.5 * np.pi, 6.283185307179586,
                                                          1e-300, 1e-9, np.nextafter(2 * np.pi, 0)]])
    s, c = oracle.det_sincos(a)
    worst = 0.0
    for ai, si, ci in zip(a, s, c):
        worst = max(worst, _ulp_err(si, mpmath.sin(mpmath.mpf(float(ai)))), _ulp_err(ci, mpmath.cos(mpmath.mpf(float(ai)))))
    assert worst < 1.0, worst


def test_sincos_close_to_glibc(oracle):
    a = np.random.default_rng(2).uniform(0, 2 * np.pi, 100000)
    s, c = oracle.det_sincos(a)
    ls, lc = oracle.libm_sincos(a)
    assert np.max(np.abs(s - ls)) <= 2.3e-16 and np.max(np.abs(c - lc)) <= 2.3e-16


def test_pow5_correctly_rounded(oracle):  # schlick, materials.nim:60
    x = np.random.default_rng(3).uniform(0, 1, 3000)
    got = oracle.det_pow(x, 5.0)
    for xi, gi in zip(x, got):
        assert _ulp_err(gi, mpmath.mpf(float(xi)) ** 5) <= 0.5000001


def test_pow_gamma(oracle):  # canvas.nim:50-54
    g = 1.0 / float(np.float32(2.2))
    x = np.concatenate([np.random.default_rng(4).uniform(0, 1.2, 3000), [0.0, 1.0, 1e-300, 1e-12]])
    got = oracle.det_pow(x, g)
    assert got[-4] == 0.0 and got[-3] == 1.0
    worst = 0.0
    for xi, gi in zip(x, got):
        if xi > 0:
            worst = max(worst, _ulp_err(gi, mpmath.mpf(float(xi)) ** mpmath.mpf(g)))
    assert worst <= 0.5000001, worst
    assert np.max(np.abs(got - oracle.libm_pow(x, g)) / np.maximum(got, 1e-300)) < 3e-16


def test_pow_special_cases(oracle):
    nan = float("nan")
    inf = float("inf")
    x = np.array([nan, -1.0, 0.0, inf, 2.0, 0.5, 1.0, 3.0])
    y = np.array([0.45, 0.45, 0.45, 0.45, inf, inf, nan, 0.0])
    got = oracle.det_pow_general(x, y)
    assert np.isnan(got[0]) and np.isnan(got[1])
    assert got[2] == 0.0 and got[3] == inf and got[4] == inf and got[5] == 0.0 and got[6] == 1.0 and got[7] == 1.0
