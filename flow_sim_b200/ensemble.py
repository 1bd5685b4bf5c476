"""Ensemble entry point (additive to the reference API): many members of one reach -
Manning-roughness calibration sweeps, inflow and release scenarios - advanced together on one GPU.

The reference runs members one after another (cases/gerd_roseires/n_calibrate.py:55-63 calls model.run
per roughness value).  Here the member axis is the parallel axis: per-member inputs go to the device once,
initial conditions, the whole time loop and the calibration objective run there, and only reduced outputs
come back.  torch is used for device buffers and streams only.
"""
from __future__ import annotations

import copy

import numpy as np

from . import abi
from .flatten import FlatCase, flatten_rating
from .runner import PreparedCall, gvf_initial_conditions, rating_objective


class nvtx_range:
    """NVTX range around a host-side phase (flatten / H2D / GVF / Newton / objective / gather), visible in nsys / ncu
    timelines; a no-op where torch was built without NVTX."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        try:
            import torch

            torch.cuda.nvtx.range_push(self.name)
            self._on = True
        except Exception:
            self._on = False
        return self

    def __exit__(self, *exc):
        if self._on:
            import torch

            torch.cuda.nvtx.range_pop()
        return False


class EnsembleRunner:
    """Holds one reach (geometry + boundary description) resident on a device and runs member batches."""

    def __init__(self, flat: FlatCase, device="cuda:0"):
        import torch

        self.torch = torch
        self.device = torch.device(device)
        self.lib = abi.load_library()
        self.flat = copy.copy(flat)
        # geometry is member-independent: upload once, reuse for every call
        self.flat.geom = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in flat.geom.items()}
        self.flat.up = copy.copy(flat.up)
        self.flat.down = copy.copy(flat.down)
        for b in (self.flat.up, self.flat.down):
            if b.series is not None and not hasattr(b.series, "data_ptr"):
                b.series = torch.from_numpy(np.ascontiguousarray(b.series, dtype=np.float64)).to(self.device)
        self.flat.ic_depth = torch.from_numpy(np.ascontiguousarray(flat.ic_depth)).to(self.device)
        self.flat.ic_flow = torch.from_numpy(np.ascontiguousarray(flat.ic_flow)).to(self.device)

    def _to_device(self, a):
        t = self.torch
        if a is None:
            return None
        if isinstance(a, t.Tensor):
            return a.to(self.device, non_blocking=True)
        return t.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device, non_blocking=True)

    def solve(self, n_members: int, member_n_main=None, member_n_fp=None, up_series=None, ic_depth=None, ic_flow=None,
              out_mode: int = abi.PR_OUT_UPSTREAM, stream=None, want_error: bool = False, member_ratings=None,
              member_order=None) -> dict:
        """One pr_ensemble_run on device buffers; returns torch tensors (no synchronisation).
        member_order: optional permutation of the members (int32) - the order in which the persistent kernel hands
        them to its warps; put the expensive members first."""
        f = copy.copy(self.flat)
        f.member_n_main = self._to_device(member_n_main)
        f.member_n_fp = self._to_device(member_n_fp)
        if member_ratings is not None:
            if f.down.type != abi.PR_BC_RATING_CURVE:
                raise ValueError("member_ratings need a rating_curve downstream boundary")
            f.down = copy.copy(f.down)
            f.down.member_ratings = list(member_ratings)
        if up_series is not None:
            f.up = copy.copy(f.up)
            f.up.series = self._to_device(up_series)
        if ic_depth is not None:
            f.ic_depth, f.ic_flow = self._to_device(ic_depth), self._to_device(ic_flow)
        call = PreparedCall(f, n_members, out_mode, abi.PR_MEM_DEVICE, self.device, want_error=want_error,
                            member_order=member_order)
        import ctypes as C

        rc = self.lib.pr_ensemble_run(*call.args(), C.c_void_p(stream or 0))
        abi.check(self.lib, rc, "pr_ensemble_run")
        return call.results()

    def roughness_sweep(self, n_main, n_fp=None, q_query=None, h_target=None, downstream_depth=None, q0=None,
                        out_mode: int = abi.PR_OUT_UPSTREAM, stream=None) -> dict:
        """The calibration ensemble (config 4): for every n_main[m] recompute the GVF initial profile
        (it depends on the roughness), run the whole simulation and, when q_query / h_target are given,
        evaluate the rating objective.  Inputs may be pinned host tensors; outputs stay on the device."""
        with nvtx_range("h2d: per-member inputs"):
            n_dev = self._to_device(n_main)
            nfp_dev = self._to_device(n_fp)
        M = int(n_dev.shape[0])
        f = copy.copy(self.flat)
        f.member_n_main, f.member_n_fp = n_dev, nfp_dev
        h_dn = float(self.flat.meta["downstream_depth"] if downstream_depth is None else downstream_depth)
        q_init = self.flat.meta["initial_flow"] if q0 is None else q0
        with nvtx_range("gvf initial conditions"):
            ich, icq, ic_status = gvf_initial_conditions(f, M, q_init, h_dn, abi.PR_MEM_DEVICE, self.device, stream)
        # the Newton iteration total of a member grows with its roughness (409 -> 676 across the gerd grid):
        # rough members first, so that the launch ends on the cheap ones
        order = self.torch.argsort(n_dev, descending=True).to(self.torch.int32)
        with nvtx_range("newton time loop"):
            res = self.solve(M, member_n_main=n_dev, member_n_fp=nfp_dev, ic_depth=ich, ic_flow=icq, out_mode=out_mode,
                             stream=stream, member_order=order)
        res["ic_status"] = ic_status
        _fail_rejected_profiles(res, ic_status)
        if q_query is not None:
            if out_mode == abi.PR_OUT_UPSTREAM:
                upq, uph = res["flow"], res["depth"]
            else:
                upq, uph = res["flow"][:, :, 0].contiguous(), res["depth"][:, :, 0].contiguous()
            with nvtx_range("rating objective"):
                lv, rm = rating_objective(self.flat.n_levels, upq, uph, float(self.flat.meta["z0"]),
                                          self._to_device(q_query), self._to_device(h_target), abi.PR_MEM_DEVICE,
                                          self.device, stream)
            res["levels"], res["rmse"] = lv, rm
        return res


    def release_scenarios(self, rating_curves, n_main=None, n_fp=None, up_series=None, downstream_depth=None,
                          q0=None, out_mode: int = abi.PR_OUT_UPSTREAM, stream=None) -> dict:
        """Release-scenario ensemble: member m runs the reach against its own downstream rating curve
        (rating_curves[m]: a rating-curve object - e.g. RoseiresRatingCurve with its own pool level, jammed gates,
        blend buffer - or an already flattened dict), optionally with its own roughness and inflow series
        (up_series [M, levels]).  The GVF initial profile is recomputed per member from the member's downstream
        depth (default: the curve's initial stage above the boundary bed, as model.py:58-66 sets it) and initial
        flow q0 (default: the case's)."""
        ratings = [r if isinstance(r, dict) else flatten_rating(r) for r in rating_curves]
        M = len(ratings)
        if downstream_depth is None:
            if all("stage0" in r for r in ratings):
                downstream_depth = np.array([r["stage0"] - self.flat.down.bed_level for r in ratings])
            else:
                downstream_depth = float(self.flat.meta["downstream_depth"])
        f = copy.copy(self.flat)
        f.member_n_main, f.member_n_fp = self._to_device(n_main), self._to_device(n_fp)
        q_init = self.flat.meta["initial_flow"] if q0 is None else q0
        hd = self._to_device(np.atleast_1d(np.asarray(downstream_depth, dtype=np.float64)))
        qi = q_init if hasattr(q_init, "data_ptr") else self._to_device(np.atleast_1d(np.asarray(q_init, dtype=np.float64)))
        ich, icq, ic_status = gvf_initial_conditions(f, M, qi, hd, abi.PR_MEM_DEVICE, self.device, stream)
        res = self.solve(M, member_n_main=f.member_n_main, member_n_fp=f.member_n_fp, up_series=up_series,
                         ic_depth=ich, ic_flow=icq, out_mode=out_mode, stream=stream, member_ratings=ratings)
        res["ic_status"] = ic_status
        _fail_rejected_profiles(res, ic_status)
        return res


def _fail_rejected_profiles(res: dict, ic_status) -> None:
    """The reference refuses to start from a backwater profile that turned supercritical (RuntimeError,
    channel.py:328-332).  In an ensemble such members are failed instead: status PR_STATUS_SUPERCRITICAL, fail level 0,
    results NaN - never a status-0 member with meaningless numbers."""
    import torch

    bad = ic_status != 0
    res["status"] = torch.where(bad, torch.full_like(res["status"], abi.PR_STATUS_SUPERCRITICAL), res["status"])
    res["fail_level"] = torch.where(bad, torch.zeros_like(res["fail_level"]), res["fail_level"])
    nan = float("nan")
    for k in ("depth", "flow"):
        shape = (-1,) + (1,) * (res[k].dim() - 1)
        res[k] = torch.where(bad.view(shape), torch.full_like(res[k], nan), res[k])
    res["iters"] = torch.where(bad.view(-1, 1), torch.zeros_like(res["iters"]), res["iters"])


def gather_packed(res: dict, total: int, rank: int, world: int, layout: str = "strided") -> dict:
    """The end-of-run gather of SURVEY.md 8e as ONE collective: per member the calibration RMSE, the Newton iteration
    counts per level, the status and the upstream stage / discharge series (8 + 4 (L-1) + 4 + 16 L bytes: 668 B for the
    gerd grid, 43.8 MB for 65,536 members) are packed into one byte row, gathered, and unpacked in member order."""
    import torch

    parts = [("rmse", res["rmse"].reshape(-1, 1)), ("iters", res["iters"]), ("status", res["status"].reshape(-1, 1)),
             ("depth", res["depth"]), ("flow", res["flow"])]
    rows = [t.contiguous().view(torch.uint8).reshape(t.shape[0], -1) for _, t in parts]
    packed = torch.cat(rows, dim=1)
    allp = gather_members(packed, total, rank, world, layout)
    out, col = {}, 0
    for (name, t), r in zip(parts, rows):
        w = r.shape[1]
        out[name] = allp[:, col:col + w].contiguous().view(t.dtype).reshape((total,) + tuple(t.shape[1:]))
        col += w
    out["rmse"], out["status"] = out["rmse"].reshape(-1), out["status"].reshape(-1)
    out["bytes_per_member"] = int(packed.shape[1])
    return out


def to_host(res: dict) -> dict:
    """Device results -> numpy (one synchronising copy per array)."""
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}


class PinnedResults:
    """Page-locked host buffers for the results of repeated ensemble calls of the same shape: ``fetch`` issues one
    asynchronous device-to-host copy per array on the current stream and synchronises once.  Pageable ``.cpu()`` copies
    of the full outputs of a 65,536-member sweep (43.8 MB) take ~16 ms; from pinned buffers ~2 ms."""

    def __init__(self):
        self.buffers = {}

    def fetch(self, res: dict, keys=None) -> dict:
        import torch

        out = {}
        for k, v in res.items():
            if keys is not None and k not in keys:
                continue
            if not hasattr(v, "is_cuda") or not v.is_cuda:
                out[k] = v
                continue
            buf = self.buffers.get(k)
            if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                self.buffers[k] = buf
            buf.copy_(v, non_blocking=True)
            out[k] = buf
        torch.cuda.current_stream().synchronize()
        return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in out.items()}


# -------------------------------------------------------------------------------------------------
# Multi-GPU: members are independent, so the ensemble is cut into contiguous blocks, one per rank
# (one process per GPU), with no traffic inside the time loop and ONE collective at the end
# (SURVEY.md 8e).  The helpers below hold the host-side logic; they are backend-agnostic
# (NCCL on GPUs, gloo in the CPU tests).
# -------------------------------------------------------------------------------------------------

def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) of the contiguous member block owned by `rank`; the first total % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def shard_members(total: int, rank: int, world: int, layout: str = "strided") -> np.ndarray:
    """Member indices owned by `rank`.

    "strided" (default): rank, rank + world, rank + 2*world, ...  A parameter sweep is usually ordered, and the
    Newton iteration count grows with the roughness (409 -> 676 iterations across the gerd grid), so dealing
    members out round-robin gives every GPU the same mix of cheap and expensive members.
    "block": the contiguous block of shard_bounds()."""
    if layout == "block":
        a, b = shard_bounds(total, rank, world)
        return np.arange(a, b)
    if layout != "strided":
        raise ValueError("layout must be 'strided' or 'block'")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, total, world)


def gather_members(local, total: int, rank: int, world: int, layout: str = "strided"):
    """The single end-of-run collective: every rank receives the per-member array of the whole ensemble in member
    order.  `local` is this rank's part ([m_local, ...] torch tensor on the backend's device), ordered as
    shard_members() orders it."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    counts = [len(shard_members(total, r, world, layout)) for r in range(world)]
    if len(set(counts)) == 1:
        flat = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, local.contiguous())
        parts = list(flat.split(counts[0], dim=0))
    else:
        pad = max(counts)
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        parts = [p[:c] for p, c in zip(parts, counts)]
    if layout == "block":
        return torch.cat(parts, dim=0)
    out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, p in enumerate(parts):
        out[r::world] = p
    return out
