"""ctypes mirror of ``include/preissmann_b200.h`` and the loader of the CUDA library.

The product path has exactly one implementation: ``csrc/libpreissmann_b200.so`` (hand-written sm_100a
kernels behind a C ABI).  There is no CPU fallback - if the library is missing or was not built,
:func:`load_library` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

PR_ABI_VERSION = 7
PR_MAX_POLY = 12
PR_MAX_GATES = 8

PR_OK, PR_ERR_ARG, PR_ERR_UNSUPPORTED, PR_ERR_CUDA = 0, 1, 2, 3
PR_STATUS_OK, PR_STATUS_MAX_ITER, PR_STATUS_NAN, PR_STATUS_SUPERCRITICAL = 0, 1, 2, 3

PR_XS_RECT, PR_XS_TRAPEZOID, PR_XS_COMPOUND, PR_XS_IRREGULAR = 0, 1, 2, 3
(PR_BC_FLOW_HYDROGRAPH, PR_BC_FIXED_DEPTH, PR_BC_NORMAL_DEPTH, PR_BC_RATING_CURVE,
 PR_BC_STAGE_HYDROGRAPH, PR_BC_FIXED_DEPTH_STORAGE) = range(6)
PR_RC_NONE, PR_RC_POLY2, PR_RC_POWER, PR_RC_POLYNOMIAL, PR_RC_ROSEIRES = range(5)
PR_OUT_FULL, PR_OUT_UPSTREAM = 0, 1
PR_MEM_HOST, PR_MEM_DEVICE = 0, 1

BC_NAMES = {
    "flow_hydrograph": PR_BC_FLOW_HYDROGRAPH,
    "fixed_depth": PR_BC_FIXED_DEPTH,
    "normal_depth": PR_BC_NORMAL_DEPTH,
    "rating_curve": PR_BC_RATING_CURVE,
    "stage_hydrograph": PR_BC_STAGE_HYDROGRAPH,
}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class pr_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_nodes", C.c_int32), ("n_levels", C.c_int32), ("n_members", C.c_int32),
        ("max_iter", C.c_int32), ("out_mode", C.c_int32), ("mem", C.c_int32), ("device", C.c_int32),
        ("lanes_per_member", C.c_int32), ("reserved0", C.c_int32),
        ("theta", C.c_double), ("dt", C.c_double), ("dx", C.c_double), ("tol", C.c_double), ("g", C.c_double),
        ("member_order", c_int32_p),
    ]


GEOM_FIELDS = ["kind", "z_bed", "b_main", "m_main", "h_bank", "T_bank", "W_bank", "b_fp_l", "b_fp_r", "m_fp",
               "n_l", "n_m", "n_r", "curvature", "w1", "w2"]


class pr_geom(C.Structure):
    _fields_ = ([("kind", c_int32_p)] + [(f, c_double_p) for f in GEOM_FIELDS[1:]] +
                [("member_n_main", c_double_p), ("member_n_fp", c_double_p),
                 ("irr_offset", c_int32_p), ("irr_x", c_double_p), ("irr_z", c_double_p),
                 ("irr_left", c_double_p), ("irr_right", c_double_p)])


class pr_rating(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("n_coef", C.c_int32),
        ("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("stage_shift", C.c_double),
        ("coef", C.c_double * PR_MAX_POLY), ("dcoef", C.c_double * PR_MAX_POLY),
        ("off", C.c_double), ("scl", C.c_double),
        ("spill", C.c_double * 6), ("sluice", C.c_double * 6),
        ("twl", C.c_double),
        ("open_state", C.c_double * PR_MAX_GATES), ("closed_state", C.c_double * PR_MAX_GATES),
        ("n_gates", C.c_int32), ("sluices_open", C.c_int32), ("sluices_closed", C.c_int32), ("gate_control", C.c_int32),
        ("stage0", C.c_double), ("buffer", C.c_double), ("q_hydro", C.c_double), ("dY", C.c_double),
        ("max_cooldown", C.c_double), ("initially_open", C.c_int32), ("reserved", C.c_int32),
    ]


class pr_bc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("reserved", C.c_int32),
        ("bed_level", C.c_double), ("bed_slope", C.c_double), ("fixed_depth", C.c_double),
        ("series", c_double_p), ("series_member_stride", C.c_int64),
        ("rating", pr_rating), ("member_ratings", C.POINTER(pr_rating)),
        ("storage_area", C.c_double), ("storage_min_stage", C.c_double),
        ("storage_ymin", C.c_double), ("storage_ymax", C.c_double),
        ("storage_curve_stage", c_double_p), ("storage_curve_area", c_double_p),
        ("storage_curve_len", C.c_int32), ("storage_capture_losses", C.c_int32),
        ("storage_alpha", C.c_double), ("storage_beta", C.c_double),
        ("storage_reservoir_length", C.c_double), ("storage_Kq", C.c_double),
        ("storage_outflow", pr_rating),
    ]


class pr_state(C.Structure):
    _fields_ = [("depth", c_double_p), ("flow", c_double_p), ("member_stride", C.c_int64)]


class pr_outputs(C.Structure):
    _fields_ = [("depth", c_double_p), ("flow", c_double_p), ("iters", c_int32_p), ("status", c_int32_p),
                ("fail_level", c_int32_p), ("storage_stage", c_double_p), ("final_error", c_double_p)]


# ---------------------------------------------------------------------------------------------

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PR_B200_LIB") or os.path.join(_HERE, "csrc", "libpreissmann_b200.so")   # override: tuning builds

#: every symbol include/preissmann_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = ["pr_abi_version", "pr_last_error", "pr_ensemble_run", "pr_gvf_initial_conditions",
                    "pr_rating_objective", "pr_fp64_peak", "pr_launch_count", "pr_math_probe",
                    "pr_normal_depth_initial_conditions", "pr_derived_results", "pr_release_workspace",
                    "pr_long_last_trips"]

_lib = None
_lock = threading.Lock()


class PreissmannLibraryError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen the CUDA library.  Raises if it has not been built (no fallback of any kind)."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise PreissmannLibraryError(
                f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  flow_sim_b200 has no CPU fallback.")
        lib = C.CDLL(p)
        lib.pr_abi_version.restype = C.c_int
        lib.pr_last_error.restype = C.c_char_p
        lib.pr_ensemble_run.restype = C.c_int
        lib.pr_ensemble_run.argtypes = [C.POINTER(pr_config), C.POINTER(pr_geom), C.POINTER(pr_bc), C.POINTER(pr_bc),
                                        C.POINTER(pr_state), C.POINTER(pr_outputs), C.c_void_p]
        lib.pr_gvf_initial_conditions.restype = C.c_int
        lib.pr_gvf_initial_conditions.argtypes = [C.POINTER(pr_config), C.POINTER(pr_geom), c_double_p, C.c_int64,
                                                  c_double_p, C.c_int64, c_double_p, c_double_p, c_int32_p, C.c_void_p]
        lib.pr_rating_objective.restype = C.c_int
        lib.pr_rating_objective.argtypes = [C.POINTER(pr_config), c_double_p, c_double_p, C.c_double, c_double_p,
                                            c_double_p, C.c_int32, c_double_p, c_double_p, C.c_void_p]
        lib.pr_fp64_peak.restype = C.c_int
        lib.pr_fp64_peak.argtypes = [C.c_double, c_double_p]
        lib.pr_launch_count.restype = C.c_int64
        lib.pr_normal_depth_initial_conditions.restype = C.c_int
        lib.pr_normal_depth_initial_conditions.argtypes = [C.POINTER(pr_config), C.POINTER(pr_geom), c_double_p, c_double_p,
                                                           C.c_int64, c_double_p, c_double_p, C.c_void_p]
        lib.pr_derived_results.restype = C.c_int
        lib.pr_derived_results.argtypes = [C.POINTER(pr_config), C.POINTER(pr_geom)] + [c_double_p] * 8 + [C.c_void_p]
        lib.pr_math_probe.restype = C.c_int
        lib.pr_release_workspace.restype = C.c_int
        lib.pr_long_last_trips.restype = C.c_int64
        lib.pr_math_probe.argtypes = [c_double_p, C.c_int32, c_double_p]
        if lib.pr_abi_version() != PR_ABI_VERSION:
            raise PreissmannLibraryError(f"ABI mismatch: library {lib.pr_abi_version()} != python {PR_ABI_VERSION}")
        if path is None:
            _lib = lib
        return lib


def check(lib, rc: int, what: str) -> None:
    if rc != PR_OK:
        msg = lib.pr_last_error()
        raise PreissmannLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


# ---------------------------------------------------------------------------------------------
# Pointer plumbing: numpy arrays (host) or torch CUDA tensors (device) -> ctypes pointers
# ---------------------------------------------------------------------------------------------

class Arena:
    """Keeps the buffers of one call alive and hands out typed pointers.

    ``mem == PR_MEM_HOST``: buffers are C-contiguous numpy arrays.
    ``mem == PR_MEM_DEVICE``: buffers are torch CUDA tensors (torch is plumbing for device memory only).
    """

    def __init__(self, mem: int = PR_MEM_HOST, device=None):
        self.mem = mem
        self.device = device
        self.keep = []

    def put(self, arr, dtype=np.float64):
        """Register an input; returns (pointer, stored buffer)."""
        if arr is None:
            return (c_double_p() if dtype == np.float64 else c_int32_p()), None
        if self.mem == PR_MEM_HOST:
            a = np.ascontiguousarray(arr, dtype=dtype)
            self.keep.append(a)
            ptr = a.ctypes.data_as(c_double_p if dtype == np.float64 else c_int32_p)
            return ptr, a
        import torch

        if isinstance(arr, torch.Tensor):
            t = arr.to(device=self.device, dtype=torch.float64 if dtype == np.float64 else torch.int32).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(self.device)
        self.keep.append(t)
        return C.cast(t.data_ptr(), c_double_p if dtype == np.float64 else c_int32_p), t

    def empty(self, shape, dtype=np.float64):
        if self.mem == PR_MEM_HOST:
            a = np.empty(shape, dtype=dtype)
            self.keep.append(a)
            return a.ctypes.data_as(c_double_p if dtype == np.float64 else c_int32_p), a
        import torch

        t = torch.empty(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32, device=self.device)
        self.keep.append(t)
        return C.cast(t.data_ptr(), c_double_p if dtype == np.float64 else c_int32_p), t


def make_rating(d: dict | None) -> pr_rating:
    r = pr_rating()
    if not d:
        r.type = PR_RC_NONE
        return r
    r.type = int(d["type"])
    for k in ("a", "b", "c", "stage_shift", "off", "scl", "twl", "stage0", "buffer", "q_hydro", "dY", "max_cooldown"):
        setattr(r, k, float(d.get(k, 0.0)))
    coef = list(np.asarray(d.get("coef", []), dtype=np.float64))
    dcoef = list(np.asarray(d.get("dcoef", []), dtype=np.float64))
    if len(coef) > PR_MAX_POLY:
        raise NotImplementedError(f"fitted Polynomial rating curve with {len(coef)} > {PR_MAX_POLY} coefficients")
    r.n_coef = len(coef)
    for i, v in enumerate(coef):
        r.coef[i] = v
    for i, v in enumerate(dcoef):
        r.dcoef[i] = v
    for name in ("spill", "sluice"):
        vals = np.asarray(d.get(name, np.zeros(6)), dtype=np.float64)
        for i in range(6):
            getattr(r, name)[i] = float(vals[i])
    for name in ("open_state", "closed_state"):
        vals = list(np.asarray(d.get(name, []), dtype=np.float64))
        if len(vals) > PR_MAX_GATES:
            raise NotImplementedError("more than PR_MAX_GATES spillway gates")
        for i, v in enumerate(vals):
            getattr(r, name)[i] = float(v)
    r.n_gates = int(d.get("n_gates", 0))
    r.sluices_open = int(d.get("sluices_open", 0))
    r.sluices_closed = int(d.get("sluices_closed", 0))
    r.gate_control = int(d.get("gate_control", 0))
    r.initially_open = int(d.get("initially_open", 0))
    return r
