"""The C-ABI shared library: loads, exports every symbol include/preissmann_b200.h declares, validates its
arguments.  No compute call is made (this file runs without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from flow_sim_b200 import abi
from flow_sim_b200.runner import PreparedCall

HEADER = os.path.join(os.path.dirname(util.GOLD), "..", "include", "preissmann_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = declared_functions()
    assert sorted(names) == sorted(abi.EXPORTED_SYMBOLS), "abi.py and the header disagree on the entry points"
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in the header but not exported"
    assert lib.pr_abi_version() == abi.PR_ABI_VERSION


def test_ctypes_structs_match_header_constants():
    src = open(HEADER).read()
    for name in ("PR_ABI_VERSION", "PR_MAX_POLY", "PR_MAX_GATES"):
        assert int(re.search(rf"#define {name} (\d+)", src).group(1)) == getattr(abi, name)
    # field order of the structs is what the C side reads: spot-check sizes against the documented layout
    assert C.sizeof(abi.pr_config) == 10 * 4 + 5 * 8 + 8   # + member_order (ABI 7)
    assert C.sizeof(abi.pr_geom) == 23 * 8        # 18 per-node / per-member arrays + 5 irregular-section arrays
    assert C.sizeof(abi.pr_state) == 24 and C.sizeof(abi.pr_outputs) == 7 * 8


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(abi.PreissmannLibraryError, match="no CPU fallback"):
        abi.load_library(str(tmp_path / "nope.so"))


def test_argument_validation_without_gpu():
    lib = abi.load_library()
    flat = util.golden_inputs("example")
    call = PreparedCall(flat, 1)
    call.cfg.abi_version = 1
    assert lib.pr_ensemble_run(*call.args(), None) == abi.PR_ERR_ARG
    assert b"abi_version" in lib.pr_last_error()
    call = PreparedCall(flat, 1)
    call.cfg.n_nodes = 1
    assert lib.pr_ensemble_run(*call.args(), None) == abi.PR_ERR_ARG
    call = PreparedCall(flat, 1)
    call.cfg.dt = 0.0
    assert lib.pr_ensemble_run(*call.args(), None) == abi.PR_ERR_ARG
    call = PreparedCall(flat, 1)
    call.cfg.mem = 7
    assert lib.pr_ensemble_run(*call.args(), None) == abi.PR_ERR_ARG
    tf = C.c_double()
    assert lib.pr_fp64_peak(1.0, None) == abi.PR_ERR_ARG
    assert lib.pr_launch_count() == 0


def test_prepared_call_rejects_inconsistent_member_arrays():
    flat = util.golden_inputs("gerd_calib_m0")
    flat.member_n_main = np.array([0.02, 0.03])
    with pytest.raises(ValueError):
        PreparedCall(flat, 3)
    flat.member_n_main = None
    flat.up.series = np.zeros((2, flat.n_levels))
    with pytest.raises(ValueError):
        PreparedCall(flat, 3)


def test_compute_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback: where no CUDA device is visible a well-formed call comes back PR_ERR_CUDA with the runtime's
    message, from the ABI and (as PreissmannLibraryError) from the solver object."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    lib = abi.load_library()
    flat = util.golden_inputs("example")
    call = PreparedCall(flat, 1)
    assert lib.pr_ensemble_run(*call.args(), None) == abi.PR_ERR_CUDA
    assert lib.pr_last_error()
    from flow_sim_b200.cases import build_example

    solver, kw = build_example()
    with pytest.raises(abi.PreissmannLibraryError):
        solver.run(verbose=0, **kw)
    assert lib.pr_long_last_trips() == -1
