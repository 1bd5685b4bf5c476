"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
It reuses the product's struct marshalling (flow_sim_b200.runner.PreparedCall) so that the oracle and
the CUDA library are fed byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from flow_sim_b200 import abi
from flow_sim_b200.runner import PreparedCall

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpreissmann_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "preissmann_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.pr_oracle_run.restype = C.c_int
        L.pr_oracle_run.argtypes = [C.POINTER(abi.pr_config), C.POINTER(abi.pr_geom), C.POINTER(abi.pr_bc),
                                    C.POINTER(abi.pr_bc), C.POINTER(abi.pr_state), C.POINTER(abi.pr_outputs)]
        L.pr_oracle_newton_step.restype = C.c_int
        L.pr_oracle_trace_prev_error.restype = None
        L.pr_oracle_trace_prev_error.argtypes = [abi.c_double_p]
        L.pr_oracle_gvf.restype = C.c_int
        L.pr_oracle_gvf.argtypes = [C.POINTER(abi.pr_config), C.POINTER(abi.pr_geom), abi.c_double_p, C.c_int64,
                                    abi.c_double_p, C.c_int64, abi.c_double_p, abi.c_double_p, abi.c_int32_p]
        L.pr_oracle_objective.restype = C.c_int
        L.pr_oracle_objective.argtypes = [C.POINTER(abi.pr_config), abi.c_double_p, abi.c_double_p, C.c_double,
                                          abi.c_double_p, abi.c_double_p, C.c_int32, abi.c_double_p, abi.c_double_p]
        L.pr_oracle_section_probe.restype = C.c_int
        L.pr_oracle_section_probe.argtypes = [C.POINTER(abi.pr_geom), C.c_int, C.c_int, C.c_double, C.c_double,
                                              C.c_double, abi.c_double_p]
        L.pr_oracle_rating_discharge.restype = C.c_double
        L.pr_oracle_rating_discharge.argtypes = [C.POINTER(abi.pr_rating), C.c_double]
        L.pr_oracle_rating_dQdz.restype = C.c_double
        L.pr_oracle_rating_dQdz.argtypes = [C.POINTER(abi.pr_rating), C.c_double]
        L.pr_oracle_brentq_poly.restype = C.c_double
        L.pr_oracle_brentq_poly.argtypes = [abi.c_double_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


def run(flat, n_members=None, out_mode=abi.PR_OUT_FULL, trace_prev_error: bool = False) -> dict:
    """trace_prev_error: also return res["prev_error"] [M, levels-1], ||R|| of the iteration before the accepted one."""
    call = PreparedCall(flat, n_members, out_mode, abi.PR_MEM_HOST)
    L = lib()
    prev = None
    if trace_prev_error:
        prev = np.full((call.M, max(flat.n_levels - 1, 1)), np.nan)
        L.pr_oracle_trace_prev_error(_dp(prev))
    try:
        rc = L.pr_oracle_run(*call.args())
    finally:
        if trace_prev_error:
            L.pr_oracle_trace_prev_error(None)
    if rc != 0:
        raise RuntimeError(f"pr_oracle_run -> {rc}")
    res = call.results()
    if prev is not None:
        res["prev_error"] = prev
    return res


def newton_step(flat, level, h0, q0, h1, q1, stage_record=None, member=0):
    """One Newton iteration: returns (R[2N], J[8N-4], delta[2N]) in the reference's ordering."""
    call = PreparedCall(flat, None, abi.PR_OUT_FULL, abi.PR_MEM_HOST)
    N = flat.n_nodes
    R = np.empty(2 * N); J = np.empty(8 * N - 4); d = np.empty(2 * N)
    st = np.full(flat.n_levels, np.nan) if stage_record is None else np.ascontiguousarray(stage_record, dtype=np.float64)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (h0, q0, h1, q1)]
    L = lib()
    rc = L.pr_oracle_newton_step(C.byref(call.cfg), C.byref(call.geom), C.byref(call.up), C.byref(call.down),
                                 C.c_int(member), C.c_int(level), *[_dp(a) for a in arrs], _dp(st), _dp(R), _dp(J), _dp(d))
    return rc, R, J, d


def gvf(flat, q0, downstream_depth, n_members=None):
    call = PreparedCall(flat, n_members, abi.PR_OUT_UPSTREAM, abi.PR_MEM_HOST)
    M, N = call.M, call.N
    q0 = np.ascontiguousarray(np.atleast_1d(q0), dtype=np.float64)
    hd = np.ascontiguousarray(np.atleast_1d(downstream_depth), dtype=np.float64)
    h = np.empty((M, N)); q = np.empty((M, N)); st = np.empty(M, dtype=np.int32)
    rc = lib().pr_oracle_gvf(C.byref(call.cfg), C.byref(call.geom), _dp(q0), 0 if q0.size == 1 else 1,
                             _dp(hd), 0 if hd.size == 1 else 1, _dp(h), _dp(q), st.ctypes.data_as(abi.c_int32_p))
    if rc != 0:
        raise RuntimeError(f"pr_oracle_gvf -> {rc}")
    return h, q, st


def objective(n_levels, up_flow, up_depth, z0, q_query, h_target):
    up_flow = np.ascontiguousarray(up_flow, dtype=np.float64)
    up_depth = np.ascontiguousarray(up_depth, dtype=np.float64)
    M = up_flow.shape[0]
    q_query = np.ascontiguousarray(q_query, dtype=np.float64)
    h_target = np.ascontiguousarray(h_target, dtype=np.float64)
    cfg = abi.pr_config(abi_version=abi.PR_ABI_VERSION, n_nodes=2, n_levels=n_levels, n_members=M)
    lv = np.empty((M, q_query.size)); rm = np.empty(M)
    lib().pr_oracle_objective(C.byref(cfg), _dp(up_flow), _dp(up_depth), float(z0), _dp(q_query), _dp(h_target),
                              q_query.size, _dp(lv), _dp(rm))
    return lv, rm


def section_probe(flat, node, h, Q, member=0, n_members=None):
    call = PreparedCall(flat, n_members, abi.PR_OUT_UPSTREAM, abi.PR_MEM_HOST)
    out = np.empty(16)
    lib().pr_oracle_section_probe(C.byref(call.geom), node, member, flat.g, float(h), float(Q), _dp(out))
    names = ["A", "P", "R", "T", "K", "n_eq", "dR_dA", "dK_dA", "Sf", "dSf_dA", "dSf_dQ", "Sc", "dSc_dA", "dSc_dQ", "Se"]
    return dict(zip(names, out))


def rating(rating_dict, stage):
    r = abi.make_rating(rating_dict)
    L = lib()
    return L.pr_oracle_rating_discharge(C.byref(r), float(stage)), L.pr_oracle_rating_dQdz(C.byref(r), float(stage))


def brentq_poly(coef, xa, xb):
    c = np.ascontiguousarray(coef, dtype=np.float64)
    err = C.c_int(0)
    x = lib().pr_oracle_brentq_poly(_dp(c), c.size, float(xa), float(xb), C.byref(err))
    return x, err.value
