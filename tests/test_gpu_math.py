"""Accuracy of the device's branch-free FP64 primitives (csrc/pr_device.cuh) against IEEE results."""

import numpy as np
import pytest

from flow_sim_b200 import abi

pytestmark = pytest.mark.gpu


def probe(x):
    lib = abi.load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty((6, x.size))
    abi.check(lib, lib.pr_math_probe(x.ctypes.data_as(abi.c_double_p), x.size, out.ctypes.data_as(abi.c_double_p)), "probe")
    return out


def test_fast_primitives_are_accurate_to_a_few_ulp():
    rng = np.random.default_rng(5)
    x = np.concatenate([10.0 ** rng.uniform(-12, 12, 200_000), rng.uniform(0.5, 2.0, 100_000), [1.0, 2.0, 3.0, 1e-30, 1e30]])
    out = probe(x)
    eps = np.finfo(float).eps
    rel = lambda got, ref: np.max(np.abs(got - ref) / np.abs(ref))
    assert rel(out[0], 1.0 / x) <= 2 * eps
    assert rel(out[1], np.sqrt(x)) <= 2 * eps
    assert rel(out[2], 1.0 / np.sqrt(x)) <= 3 * eps
    ref_rcbrt = np.array([float(v) for v in np.exp(-np.log(x.astype(np.longdouble)) / 3)])
    assert rel(out[3], ref_rcbrt) <= 4 * eps
    # negative arguments of the reciprocal (determinants can have either sign)
    assert rel(probe(-x)[0], -1.0 / x) <= 2 * eps
    print("seed errors: rcp %.3e rsqrt %.3e" % (rel(out[4], 1.0 / x), rel(out[5], 1.0 / np.sqrt(x))))


def test_sqrt_of_zero_is_zero():
    out = probe(np.array([0.0, 4.0]))
    assert out[1, 0] == 0.0 and out[1, 1] == 2.0
