"""The product's host generator == the oracle's generator, bit for bit (CPU only)."""
import numpy as np
import pytest

import metrovector_b200 as mv
import oracle


@pytest.mark.parametrize("dtype", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [1, 2, 7, 8, 16, 33, 128, 769])
def test_host_synth_matches_oracle(dtype, d):
    for seed, row0, n in ((0, 0, 5), (0x4D564631, 123456789012, 4), (99, 2 ** 40, 3)):
        a = mv.synth_fill_host(dtype, seed, row0, n, d)
        b = oracle.synth(dtype, seed, row0, n, d)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_rows_are_position_independent():
    a = oracle.synth(oracle.F16, 7, 0, 100, 64)
    b = oracle.synth(oracle.F16, 7, 37, 10, 64)
    assert np.array_equal(a[37:47].view(np.uint16), b.view(np.uint16))
    c = oracle.synth(oracle.F32, 7, 0, 64, 32, threads=4)
    assert np.array_equal(c, oracle.synth(oracle.F32, 7, 0, 64, 32))


def test_distribution_sanity():
    x = oracle.synth(oracle.F32, 1, 0, 2000, 128)
    assert abs(float(x.mean())) < 0.01 and 0.78 < float(x.std()) < 0.85 and float(np.abs(x).max()) < 2.0
    u = oracle.synth(oracle.U8, 1, 0, 2000, 128)
    assert 120 < float(u.mean()) < 135
