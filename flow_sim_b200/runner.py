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
                 mem: int = abi.PR_MEM_HOST, device=None, lanes: int = 0, want_error: bool = True, member_order=None):
        M = int(n_members if n_members is not None else flat.n_members_hint)
        N, L = flat.n_nodes, flat.n_levels
        self.flat, self.M, self.N, self.L, self.out_mode, self.mem = flat, M, N, L, out_mode, mem
        ar = self.arena = abi.Arena(mem, device)

        dev_index = _device_index(mem, device)
        self.cfg = abi.pr_config(abi_version=abi.PR_ABI_VERSION, n_nodes=N, n_levels=L, n_members=M,
                                 max_iter=flat.max_iter, out_mode=out_mode, mem=mem, device=dev_index,
                                 lanes_per_member=lanes, theta=flat.theta, dt=flat.dt, dx=flat.dx,
                                 tol=flat.tol, g=flat.g)

        if member_order is not None:          # processing order of the members (a permutation; results stay in member order)
            if len(member_order) != M:
                raise ValueError("member_order must have one entry per member")
            self.cfg.member_order = ar.put(member_order, np.int32)[0]

        self.geom = _geom_struct(flat, ar, M)

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
        if b.member_ratings is not None:
            if len(b.member_ratings) != self.M:
                raise ValueError("member_ratings must have one entry per member")
            arr = (abi.pr_rating * self.M)(*[abi.make_rating(d) for d in b.member_ratings])   # host memory, always
            self.arena.keep.append(arr)
            s.member_ratings = C.cast(arr, C.POINTER(abi.pr_rating))
        s.storage_area, s.storage_min_stage = b.storage_area, b.storage_min_stage
        s.storage_ymin, s.storage_ymax = b.storage_ymin, b.storage_ymax
        if b.storage_curve is not None:
            curve = b.storage_curve
            s.storage_curve_stage = self.arena.put(np.ascontiguousarray(np.asarray(curve)[:, 0]))[0]
            s.storage_curve_area = self.arena.put(np.ascontiguousarray(np.asarray(curve)[:, 1]))[0]
            s.storage_curve_len = int(np.asarray(curve).shape[0])
        s.storage_alpha, s.storage_beta = b.storage_alpha, b.storage_beta
        s.storage_capture_losses = int(bool(b.storage_losses))
        s.storage_reservoir_length, s.storage_Kq = b.storage_reservoir_length, b.storage_Kq
        s.storage_outflow = abi.make_rating(b.storage_outflow)
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


def _geom_struct(flat: FlatCase, arena: abi.Arena, M: int) -> abi.pr_geom:
    g = abi.pr_geom()
    for name in abi.GEOM_FIELDS:
        setattr(g, name, arena.put(flat.geom[name], np.int32 if name == "kind" else np.float64)[0])
    for name in ("member_n_main", "member_n_fp"):
        v = getattr(flat, name)
        if v is not None and len(v) != M:
            raise ValueError(f"{name} has {len(v)} entries for {M} members")
        setattr(g, name, arena.put(v)[0])
    if "irr_offset" in flat.geom:          # IrregularSection nodes: CSR polylines + composite-roughness limits
        g.irr_offset = arena.put(flat.geom["irr_offset"], np.int32)[0]
        for name in ("irr_x", "irr_z", "irr_left", "irr_right"):
            setattr(g, name, arena.put(flat.geom[name], np.float64)[0])
    return g


def _device_index(mem: int, device) -> int:
    """CUDA ordinal the call's device buffers live on (-1 = the current device, host-memory calls)."""
    if mem != abi.PR_MEM_DEVICE:
        return -1
    import torch

    idx = torch.device(device).index if device is not None else None
    return torch.cuda.current_device() if idx is None else idx


def _config(flat: FlatCase, M: int, mem: int, device, out_mode: int = abi.PR_OUT_UPSTREAM) -> abi.pr_config:
    dev_index = _device_index(mem, device)
    return abi.pr_config(abi_version=abi.PR_ABI_VERSION, n_nodes=flat.n_nodes, n_levels=flat.n_levels, n_members=M,
                         max_iter=flat.max_iter, out_mode=out_mode, mem=mem, device=dev_index, lanes_per_member=0,
                         theta=flat.theta, dt=flat.dt, dx=flat.dx, tol=flat.tol, g=flat.g)


def gvf_initial_conditions(flat: FlatCase, n_members: int, q0, downstream_depth,
                           mem: int = abi.PR_MEM_HOST, device=None, stream=None):
    """Channel._gvh_conditions (channel.py:307-378) for every member on the device.
    Returns (depth[M,N], flow[M,N], status[M])."""
    lib = abi.load_library()
    ar = abi.Arena(mem, device)
    cfg = _config(flat, n_members, mem, device)
    g = _geom_struct(flat, ar, n_members)
    if mem == abi.PR_MEM_HOST or not hasattr(q0, "data_ptr"):
        q0 = np.atleast_1d(np.asarray(q0, dtype=np.float64))
    if mem == abi.PR_MEM_HOST or not hasattr(downstream_depth, "data_ptr"):
        downstream_depth = np.atleast_1d(np.asarray(downstream_depth, dtype=np.float64))
    n_q0, n_hd = q0.shape[0], downstream_depth.shape[0]
    if n_q0 not in (1, n_members) or n_hd not in (1, n_members):
        raise ValueError("q0 and downstream_depth must have 1 or M entries")
    q0p, _ = ar.put(q0)
    hdp, _ = ar.put(downstream_depth)
    hp, h = ar.empty((n_members, flat.n_nodes))
    qp, q = ar.empty((n_members, flat.n_nodes))
    sp, st = ar.empty((n_members,), np.int32)
    rc = lib.pr_gvf_initial_conditions(C.byref(cfg), C.byref(g), q0p, 0 if n_q0 == 1 else 1, hdp,
                                       0 if n_hd == 1 else 1, hp, qp, sp, C.c_void_p(stream or 0))
    abi.check(lib, rc, "pr_gvf_initial_conditions")
    return h, q, st


def rating_objective(n_levels: int, up_flow, up_depth, z0: float, q_query, h_target,
                     mem: int = abi.PR_MEM_HOST, device=None, stream=None):
    """np.interp(Q, flow[:,0], depth[:,0]+z0) and its RMSE against h_target per member
    (model.py:105-113, n_calibrate.py:55-63).  Returns (levels[M,nq], rmse[M])."""
    lib = abi.load_library()
    ar = abi.Arena(mem, device)
    M = int(up_flow.shape[0])
    cfg = abi.pr_config(abi_version=abi.PR_ABI_VERSION, n_nodes=2, n_levels=n_levels, n_members=M, max_iter=1,
                        out_mode=abi.PR_OUT_UPSTREAM, mem=mem, device=_device_index(mem, device), theta=0.5, dt=1.0,
                        dx=1.0, tol=1.0, g=9.80665)
    fq, _ = ar.put(up_flow)
    fh, _ = ar.put(up_depth)
    qq, qa = ar.put(np.asarray(q_query, dtype=np.float64) if not hasattr(q_query, "data_ptr") else q_query)
    ht, _ = ar.put(np.asarray(h_target, dtype=np.float64) if not hasattr(h_target, "data_ptr") else h_target)
    nq = int(qa.shape[0])
    lp, lv = ar.empty((M, nq))
    rp, rm = ar.empty((M,))
    rc = lib.pr_rating_objective(C.byref(cfg), fq, fh, float(z0), qq, ht, nq, lp, rp, C.c_void_p(stream or 0))
    abi.check(lib, rc, "pr_rating_objective")
    return lv, rm


def normal_depth_initial_conditions(flat: FlatCase, n_members: int, q0, bed_slope=None,
                                    mem: int = abi.PR_MEM_HOST, device=None, stream=None):
    """Channel._steady_conditions (channel.py:296-305) for every member on the device: Brent's method per node.
    Returns (depth[M,N], flow[M,N])."""
    lib = abi.load_library()
    ar = abi.Arena(mem, device)
    cfg = _config(flat, n_members, mem, device)
    g = _geom_struct(flat, ar, n_members)
    slope = flat.meta["bed_slope"] if bed_slope is None else bed_slope
    if mem == abi.PR_MEM_HOST or not hasattr(q0, "data_ptr"):
        q0 = np.atleast_1d(np.asarray(q0, dtype=np.float64))
    n_q0 = q0.shape[0]
    if n_q0 not in (1, n_members):
        raise ValueError("q0 must have 1 or M entries")
    sp, _ = ar.put(slope)
    q0p, _ = ar.put(q0)
    hp, h = ar.empty((n_members, flat.n_nodes))
    qp, q = ar.empty((n_members, flat.n_nodes))
    rc = lib.pr_normal_depth_initial_conditions(C.byref(cfg), C.byref(g), sp, q0p, 0 if n_q0 == 1 else 1, hp, qp,
                                                C.c_void_p(stream or 0))
    abi.check(lib, rc, "pr_normal_depth_initial_conditions")
    return h, q


DERIVED = ("level", "area", "top_width", "froude_number", "velocity", "wave_celerity")


def derived_results(flat: FlatCase, depth, flow, mem: int = abi.PR_MEM_HOST, device=None, stream=None,
                    want=DERIVED) -> dict:
    """The array part of Solver.prepare_results (solver.py:65-98) for results shaped [M, levels, N]."""
    lib = abi.load_library()
    ar = abi.Arena(mem, device)
    M, L, N = depth.shape
    if (L, N) != (flat.n_levels, flat.n_nodes):
        raise ValueError("depth must be [members, levels, nodes]")
    cfg = _config(flat, M, mem, device, abi.PR_OUT_FULL)
    g = _geom_struct(flat, ar, M) if flat.member_n_main is None and flat.member_n_fp is None else None
    if g is None:
        import copy

        f2 = copy.copy(flat); f2.member_n_main = None; f2.member_n_fp = None   # geometry only, roughness is irrelevant here
        g = _geom_struct(f2, ar, M)
    dp, _ = ar.put(depth)
    fp, _ = ar.put(flow)
    ptrs, bufs = [], {}
    for name in DERIVED:
        if name in want:
            p, b = ar.empty((M, L, N))
            bufs[name] = b
        else:
            p = abi.c_double_p()
        ptrs.append(p)
    rc = lib.pr_derived_results(C.byref(cfg), C.byref(g), dp, fp, *ptrs, C.c_void_p(stream or 0))
    abi.check(lib, rc, "pr_derived_results")
    return bufs
