"""CPU test: the deterministic libm (hnm_detmath.h) against glibc on the ranges the renderer uses."""
import numpy as np
import pytest

N = 200000


def _ulp_diff(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


@pytest.mark.parametrize("fn,name,gen,min_equal", [
    (0, "sin", lambda r: (r.random(N) * 2 * np.pi, None), 0.95),
    (1, "cos", lambda r: (r.random(N) * 2 * np.pi, None), 0.95),
    (2, "exp", lambda r: (-r.random(N) * 0.5, None), 0.97),
    (2, "exp_wide", lambda r: ((r.random(N) - 0.5) * 1400, None), 0.97),
    (3, "pow_2.2", lambda r: (r.random(N), np.full(N, 2.2)), 0.97),
    (3, "pow_texel", lambda r: (r.integers(0, 256, N) / 255.0, np.full(N, 2.2)), 0.97),
    (3, "pow_inv", lambda r: (r.random(N), np.full(N, 1 / 2.2)), 0.97),
    (4, "acos", lambda r: (r.random(N) * 2 - 1, None), 0.90),
])
def test_detmath_within_one_ulp_of_glibc(oracle, oracle_glibc, fn, name, gen, min_equal):
    x, y = gen(np.random.default_rng(7))
    det = oracle.math(fn, x, y)
    ref = oracle_glibc.math(fn, x, y)
    d = _ulp_diff(det, ref)
    assert d.max() <= 1, (name, d.max())
    assert (d == 0).mean() >= min_equal, (name, (d == 0).mean())


def test_detmath_special_values(oracle):
    assert oracle.math(3, [0.0, 1.0, 0.5, 0.0, 2.0], [2.2, 2.2, 0.0, 1 / 2.2, 3.0]).tolist() == [0.0, 1.0, 1.0, 0.0, 8.0]
    assert np.isnan(oracle.math(3, [-0.5], [2.2])[0])
    e = oracle.math(2, [0.0, -1e4, 1e4, -745.2])
    assert e[0] == 1.0 and e[1] == 0.0 and np.isinf(e[2]) and e[3] == 0.0
    a = oracle.math(4, [1.0, -1.0, 0.0, 1.5])
    assert a[0] == 0.0 and a[1] == np.pi and a[2] == np.pi / 2 and np.isnan(a[3])
    s = oracle.math(0, [0.0, np.pi / 2])
    assert s[0] == 0.0 and s[1] == 1.0
    c = oracle.math(1, [0.0])
    assert c[0] == 1.0
