"""FlatCase -> C-ABI call.  One marshalling routine used by PreissmannSolver.run(), the ensemble API,
bench.py and (with the oracle's function pointer) the parity tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .flatten import FlatCase


class PreparedCall:
    """ctypes structs + the buffers behind them for one pr_ensemble_run-shaped call."""

    def __init__(self, flat: FlatCase, n_members: int | None = None, out_mode: int = abi.PR_OUT_FULL,
                 mem: int = abi.PR_MEM_HOST, device=None, lanes: int = 0, want_error: bool = True):
        M = int(n_members if n_members is not None else flat.n_members_hint)
        N, L = flat.n_nodes, flat.n_levels
        self.flat, self.M, self.N, self.L, self.out_mode, self.mem = flat, M, N, L, out_mode, mem
        ar = self.arena = abi.Arena(mem, device)

        dev_index = -1
        if mem == abi.PR_MEM_DEVICE:
            import torch

            dev_index = torch.device(device).index if device is not None else torch.cuda.current_device()
            if dev_index is None:
                dev_index = torch.cuda.current_device()
        self.cfg = abi.pr_config(abi_version=abi.PR_ABI_VERSION, n_nodes=N, n_levels=L, n_members=M,
                                 max_iter=flat.max_iter, out_mode=out_mode, mem=mem, device=dev_index,
                                 lanes_per_member=lanes, theta=flat.theta, dt=flat.dt, dx=flat.dx,
                                 tol=flat.tol, g=flat.g)

        g = abi.pr_geom()
        for name in abi.GEOM_FIELDS:
            ptr, _ = ar.put(flat.geom[name], np.int32 if name == "kind" else np.float64)
            setattr(g, name, ptr)
        for name in ("member_n_main", "member_n_fp"):
            v = getattr(flat, name)
            if v is not None and len(v) != M:
                raise ValueError(f"{name} has {len(v)} entries for {M} members")
            setattr(g, name, ar.put(v)[0])
        self.geom = g

        self.up = self._bc(flat.up)
        self.down = self._bc(flat.down)

        st = abi.pr_state()
        icd = flat.ic_depth
        if np.ndim(icd) == 2 or (hasattr(icd, "dim") and icd.dim() == 2):
            if icd.shape[0] != M:
                raise ValueError("per-member initial conditions must have M rows")
            st.member_stride = N
        else:
            st.member_stride = 0
        st.depth = ar.put(flat.ic_depth)[0]
        st.flow = ar.put(flat.ic_flow)[0]
        self.ic = st

        o = abi.pr_outputs()
        shape = (M, L, N) if out_mode == abi.PR_OUT_FULL else (M, L)
        o.depth, self.depth = ar.empty(shape)
        o.flow, self.flow = ar.empty(shape)
        o.iters, self.iters = ar.empty((M, max(L - 1, 1)), np.int32)
        o.status, self.status = ar.empty((M,), np.int32)
        o.fail_level, self.fail_level = ar.empty((M,), np.int32)
        if flat.down.type == abi.PR_BC_FIXED_DEPTH_STORAGE:
            o.storage_stage, self.storage_stage = ar.empty((M, L))
        else:
            self.storage_stage = None
        if want_error:
            o.final_error, self.final_error = ar.empty((M, max(L - 1, 1)))
        else:
            self.final_error = None
        self.out = o

    def _bc(self, b) -> abi.pr_bc:
        s = abi.pr_bc()
        s.type = b.type
        s.bed_level, s.bed_slope, s.fixed_depth = b.bed_level, b.bed_slope, b.fixed_depth
        if b.series is not None:
            ser = b.series
            two_d = (np.ndim(ser) == 2) if not hasattr(ser, "dim") else ser.dim() == 2
            if two_d and ser.shape[0] != self.M:
                raise ValueError("per-member boundary series must have M rows")
            s.series_member_stride = self.L if two_d else 0
            s.series = self.arena.put(ser)[0]
        s.rating = abi.make_rating(b.rating)
        s.storage_area, s.storage_min_stage = b.storage_area, b.storage_min_stage
        s.storage_ymin, s.storage_ymax = b.storage_ymin, b.storage_ymax
        return s

    def args(self):
        return (C.byref(self.cfg), C.byref(self.geom), C.byref(self.up), C.byref(self.down), C.byref(self.ic),
                C.byref(self.out))

    def results(self) -> dict:
        r = dict(depth=self.depth, flow=self.flow, iters=self.iters[:, : self.L - 1], status=self.status,
                 fail_level=self.fail_level)
        if self.storage_stage is not None:
            r["storage_stage"] = self.storage_stage
        if self.final_error is not None:
            r["final_error"] = self.final_error[:, : self.L - 1]
        return r


def run_flat(flat: FlatCase, n_members: int | None = None, out_mode: int = abi.PR_OUT_FULL,
             mem: int = abi.PR_MEM_HOST, device=None, lanes: int = 0, stream=None) -> dict:
    """Run the CUDA solver on a flattened case; returns numpy arrays (HOST) or torch tensors (DEVICE)."""
    lib = abi.load_library()
    call = PreparedCall(flat, n_members, out_mode, mem, device, lanes)
    rc = lib.pr_ensemble_run(*call.args(), C.c_void_p(stream or 0))
    abi.check(lib, rc, "pr_ensemble_run")
    if mem == abi.PR_MEM_DEVICE:
        import torch

        torch.cuda.synchronize()
    return call.results()
