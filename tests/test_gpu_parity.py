"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs
and against the committed outputs of the live reference (tests/golden, see oracle/make_golden.py)."""
import numpy as np
import pytest

import util
from flow_sim_b200 import abi
from flow_sim_b200.runner import run_flat

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", util.SMALL_CASES + ["gerd_full", "gerd_gated_full"])
def test_cuda_vs_reference_golden(case):
    if not util.has_golden_outputs(case):
        pytest.skip("golden outputs not generated")
    flat = util.golden_inputs(case)
    ref = util.golden_outputs(case)
    out = run_flat(flat, mem=abi.PR_MEM_HOST)
    assert out["status"][0] == abi.PR_STATUS_OK
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], case)
    assert np.array_equal(out["iters"][0], ref["iters"]), f"{case}: Newton iteration counts differ"
    if "storage_stage" in ref.files:
        assert util.max_rel(out["storage_stage"][0], ref["storage_stage"]) <= util.RTOL


@pytest.mark.parametrize("case", util.SMALL_CASES)
def test_cuda_vs_oracle(case):
    import oracle_py

    flat = util.golden_inputs(case)
    ora = oracle_py.run(flat)
    out = run_flat(flat, mem=abi.PR_MEM_HOST)
    util.assert_parity(out["depth"][0], out["flow"][0], ora["depth"][0], ora["flow"][0], case)
    assert np.array_equal(out["iters"], ora["iters"])
    assert np.array_equal(out["status"], ora["status"])
