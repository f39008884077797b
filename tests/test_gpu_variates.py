"""GPU: device generator known answers (Random123 Philox4x32-10 vectors) and distribution moments of the six manusim
distributions, scalar draws (clamped at 0) and batch sums, through the C-ABI test hooks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from manusim_b200 import _lib

    return _lib, _lib.load()


def test_philox4x32_10_known_answers():
    L, lib = _lib()
    ctr = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], dtype=np.uint32)
    key = np.array([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0xA4093822, 0x299F31D0]], dtype=np.uint32)
    out = np.zeros((3, 4), dtype=np.uint32)
    L.check(lib.msim_test_philox(ctr.ctypes.data, key.ctypes.data, out.ctypes.data, 3))
    want = np.array([[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8],
                     [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
                     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]], dtype=np.uint32)
    assert np.array_equal(out, want)


def _draw(dist, params, n=400000, batch=-1, seed=99):
    from manusim_b200.compiler import compile_dist

    L, lib = _lib()
    d = compile_dist({"dist": dist, "params": params})
    cd = L.Dist(kind=d.kind, raw0=d.raw0, raw1=d.raw1, d0=d.d0, d1=d.d1)
    out = np.zeros(n, dtype=np.float64)
    L.check(lib.msim_test_variates(C.byref(cd), seed, 2, batch, out.ctypes.data, n))
    return out


@pytest.mark.parametrize("dist,params,mean,var", [
    ("gamma", [0.78, 0.45], 0.78, 0.45 ** 2), ("erlang", [10, 5.77], 10.0, 5.77 ** 2), ("gamma", [1.0, 1.6], 1.0, 1.6 ** 2),
    ("expo", [4000], 4000.0, 4000.0 ** 2), ("uniform", [10, 2.0], 10.0, 4.0), ("constant", [63], 63.0, 0.0)])
def test_scalar_moments(dist, params, mean, var):
    x = _draw(dist, params)
    assert (x >= 0).all()
    se = np.sqrt(max(var, 1e-12) / len(x))
    assert abs(x.mean() - mean) < 6 * se + 1e-9
    if var > 0:
        assert abs(x.var() - var) / var < 0.03


def test_normal_is_clamped_like_the_reference():
    x = _draw("normal", [0.04, 0.05])                       # toc_dbr r02 setup: 21 % of draws clamp to 0
    assert (x >= 0).all()
    from scipy.stats import norm

    p0 = norm.cdf(-0.04 / 0.05)
    assert abs((x == 0).mean() - p0) < 0.004
    want = 0.04 * (1 - p0) + 0.05 * norm.pdf(-0.04 / 0.05)  # E[max(0, N)]
    assert abs(x.mean() - want) < 5 * 0.05 / np.sqrt(len(x))


@pytest.mark.parametrize("dist,params,count,mean,var", [
    ("gamma", [0.78, 0.45], 5, 3.9, 5 * 0.45 ** 2), ("expo", [2.0], 3, 6.0, 12.0), ("uniform", [10, 2.0], 2, 20.0, 8.0),
    ("constant", [0.78], 3, 2.34, 0.0), ("normal", [5.0, 1.0], 4, 20.0, 4.0)])
def test_batch_sum_moments(dist, params, count, mean, var):
    x = _draw(dist, params, n=200000, batch=count)
    se = np.sqrt(max(var, 1e-12) / len(x))
    assert abs(x.mean() - mean) < 6 * se + 1e-9
    if var > 0:
        assert abs(x.var() - var) / var < 0.04
    else:
        assert np.all(x == x[0]) and x[0] == count * params[0]
