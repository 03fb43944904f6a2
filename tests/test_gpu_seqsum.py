"""chr3d_seq_sum on the device: the parallel evaluation of `for v in a: s += v` must return the bits of the loop
(the Perl's `$sum += ...`, C3D:132-139 mean of IF**alpha, C3D:590-598 sum of deviations)."""
import numpy as np
import pytest

from chromosome3d_b200 import api, datasets

pytestmark = pytest.mark.gpu


def scalar_sum(a):
    s = 0.0
    for v in a.tolist():
        s += v
    return s


def same_bits(x, y):
    return np.float64(x).tobytes() == np.float64(y).tobytes() or (x != x and y != y)


def cases():
    rng = np.random.default_rng(20261019)
    n = 150_001
    out = {}
    out["uniform"] = rng.random(n)
    L = 455
    cnt = np.floor(5000.0 * rng.random((L, L)) / (1.0 + np.abs(np.subtract.outer(np.arange(L), np.arange(L)))))
    out["hic_pow_1.1"] = (cnt ** 1.1).ravel()
    out["hic_pow_0.5"] = (cnt ** 0.5).ravel()
    out["wide_range"] = np.ldexp(rng.random(n) + 0.5, rng.integers(-60, 60, n))
    out["dyadic_ties"] = np.ldexp(rng.integers(1, 17, n).astype(np.float64), rng.integers(-50, 10, n))
    dev = np.abs(rng.integers(0, 100000, n) / 1000.0 - rng.integers(0, 1000, n) / 10.0)
    dev[rng.random(n) < 0.75] = 0.0
    out["deviations"] = dev
    z = np.zeros(n); z[n // 3:2 * n // 3] = rng.integers(0, 1000, 2 * n // 3 - n // 3); z[2 * n // 3:] = rng.integers(0, 1000, n - 2 * n // 3) + 0.5
    out["zeros_ints_halves"] = z
    m = rng.random(n); m[rng.random(n) < 0.005] *= -1.0; m[rng.random(n) < 0.003] = 5e-324 * 7; m[rng.random(n) < 0.002] = -0.0
    m[rng.random(n) < 0.002] = 2.0 ** 80
    out["mixed_signs_denormals"] = m
    out["powers_of_two"] = np.ldexp(1.0, rng.integers(-4, 4, n))
    out["tiny"] = np.ldexp(rng.random(20_000) + 0.5, rng.integers(-1020, -1000, 20_000))
    inf = rng.random(5000); inf[1234] = np.inf
    out["inf"] = inf
    nan = rng.random(5000); nan[1234] = np.nan
    out["nan"] = nan
    out["one"] = np.array([0.1])
    out["empty"] = np.zeros(0)
    return out


def test_seq_sum_has_the_bits_of_the_scalar_loop():
    api.init(0)
    for name, a in cases().items():
        ref = scalar_sum(a)
        got = api.seq_sum(a)
        assert same_bits(ref, got), (name, float.hex(ref), float.hex(got))
        api.set_option("seqsum", "serial")
        try:
            ser = api.seq_sum(a)
        finally:
            api.set_option("seqsum", None)
        assert same_bits(ref, ser), (name, "serial", float.hex(ref), float.hex(ser))


def test_seq_sum_segments_are_independent():
    api.init(0)
    rng = np.random.default_rng(5)
    a = rng.random((20, 30_011)) * np.ldexp(1.0, rng.integers(-8, 8, (20, 1)))
    a[3, :5000] = 0.0
    got = api.seq_sum(a)
    for r in range(a.shape[0]):
        assert same_bits(scalar_sum(a[r]), got[r]), r


def test_mean_of_a_shipped_matrix():
    """The sum the restraint derivation takes (C3D:132-139) on a real input, against the scalar loop over libm's pow."""
    api.init(0)
    IF = datasets.load_if_matrix(datasets.input_path("chr1_500kb"))
    P = np.array([float(v) ** 1.1 for v in IF.ravel().tolist()])
    assert same_bits(scalar_sum(P), api.seq_sum(P))
