"""``PreissmannSolver`` with the reference's constructor and ``run`` signature (preissmann.py:9-163).

``run`` is where this package takes over: instead of the Python Newton loop + scipy ``spsolve`` it
flattens the configured objects and calls ``pr_ensemble_run`` (CUDA, sm_100a) through the C ABI with one
member.  There is no CPU fallback: without the built library, or for a configuration the device path
does not cover, ``run`` raises.
"""
from __future__ import annotations

import numpy as np

from .. import abi
from ..flatten import flatten_solver
from ..runner import run_flat
from .solver import Solver


class PreissmannSolver(Solver):
    def __init__(self, theta, **kwargs):
        super().__init__(**kwargs)
        self.theta = theta
        self.unknowns = None
        self.type = self._type = "preissmann"
        self.iterations = None        # Newton iterations per time level (the reference only prints them)
        self.status = None
        self.initialize_t0()

    def initialize_t0(self):
        super().initialize_t0()
        self.unknowns = self.channel.initial_conditions.flatten()

    def run(self, tolerance=1e-4, verbose=3, max_iter=100, diagnos=False):
        flat = flatten_solver(self, tolerance=tolerance, max_iter=max_iter)
        out = run_flat(flat, n_members=1, out_mode=abi.PR_OUT_FULL, mem=abi.PR_MEM_HOST)
        status, fail_level = int(out["status"][0]), int(out["fail_level"][0])
        self.iterations = out["iters"][0].copy()
        self.final_error = out["final_error"][0].copy()
        self.status = status
        last = self.number_of_time_levels - 1 if status == abi.PR_STATUS_OK else fail_level
        self.depth[: last + 1] = out["depth"][0][: last + 1]
        self.flow[: last + 1] = out["flow"][0][: last + 1]
        if verbose >= 2:
            for k in range(1, last + 1):
                print(f"\n> Time level #{k}\n>> {int(self.iterations[k - 1])} iterations.")
        if status != abi.PR_STATUS_OK:
            self.time_level = fail_level
            if status == abi.PR_STATUS_NAN and diagnos:
                raise ValueError("NaN in system assembly")
            # preissmann.py:124-126
            raise ValueError(f"Convergence within {int(self.iterations[fail_level - 1])} iterations couldn't be achieved.")
        self.time_level = self.number_of_time_levels - 1
        if "storage_stage" in out:
            self._storage_stage = out["storage_stage"][0]
        # the reference leaves the post-update vector of the last level in `unknowns` (preissmann.py:147);
        # it is not part of any result, so the stored last level is reported instead
        self.unknowns = np.column_stack([self.depth[-1], self.flow[-1]]).ravel()
        self._finalize(verbose)

    def run_ensemble(self, members: dict, tolerance=1e-4, max_iter=100, device="cuda:0", full_output=False,
                     q_query=None, h_target=None) -> dict:
        """Additive entry point: run many members of this reach in one launch.

        ``members`` holds per-member overrides, each of length M:
          ``n_main`` / ``n_fp``      Manning roughness (model.run(n_main=, n_fp=), cross_section.py:887-893)
          ``inflow``                 [M, levels] upstream series sampled at t = k*dt, or a list of Hydrograph objects
          ``rating_curves``          one downstream rating-curve object per member (release scenarios:
                                     initial_roseires_level, jammed gates ... of model.run)
          ``downstream_depth``/``initial_flow``  inputs of the GVF initial profile, which is recomputed per member
                                     whenever the channel was set up with the backwater initial conditions
        Returns numpy arrays: ``depth``/``flow`` ([M, levels] at the upstream node, or [M, levels, nodes] with
        ``full_output``), ``iterations`` [M, levels-1], ``status`` [M], and ``levels``/``rmse`` when the
        calibration targets ``q_query`` / ``h_target`` (model.py:105-113, n_calibrate.py:55-63) are given."""
        from ..ensemble import EnsembleRunner, to_host
        from ..runner import rating_objective

        known = {"n_main", "n_fp", "inflow", "rating_curves", "downstream_depth", "initial_flow"}
        if set(members) - known:
            raise ValueError(f"unknown member overrides: {sorted(set(members) - known)}")
        sizes = {len(v) for v in members.values() if v is not None and np.ndim(v) > 0}
        if len(sizes) != 1:
            raise ValueError("member overrides must all have the same length M")
        M = sizes.pop()
        flat = flatten_solver(self, tolerance=tolerance, max_iter=max_iter)
        runner = EnsembleRunner(flat, device)
        series = members.get("inflow")
        if series is not None and not isinstance(series, np.ndarray):
            series = np.array([[float(hy.get_at(k * flat.dt)) for k in range(flat.n_levels)] for hy in series])
        mode = abi.PR_OUT_FULL if full_output else abi.PR_OUT_UPSTREAM
        n_main, n_fp = members.get("n_main"), members.get("n_fp")
        gvf = flat.meta.get("ic_method") == "GVF_equation"
        if members.get("rating_curves") is not None:
            if not gvf:
                raise NotImplementedError("release scenarios need the backwater (GVF) initial conditions")
            res = runner.release_scenarios(members["rating_curves"], n_main=n_main, n_fp=n_fp, up_series=series,
                                           downstream_depth=members.get("downstream_depth"),
                                           q0=members.get("initial_flow"), out_mode=mode)
        elif gvf and (n_main is not None or n_fp is not None or members.get("downstream_depth") is not None
                      or members.get("initial_flow") is not None):
            if series is not None:
                runner.flat.up.series = runner._to_device(series)
            nm = n_main if n_main is not None else np.full(M, np.nan)
            if n_main is None:
                raise NotImplementedError("per-member n_fp / initial state without n_main")
            res = runner.roughness_sweep(nm, n_fp=n_fp, downstream_depth=members.get("downstream_depth"),
                                         q0=members.get("initial_flow"), out_mode=mode)
        else:
            res = runner.solve(M, member_n_main=n_main, member_n_fp=n_fp, up_series=series, out_mode=mode)
        if q_query is not None:
            upq = res["flow"] if not full_output else res["flow"][:, :, 0].contiguous()
            uph = res["depth"] if not full_output else res["depth"][:, :, 0].contiguous()
            lv, rm = rating_objective(flat.n_levels, upq, uph, float(flat.meta["z0"]), runner._to_device(q_query),
                                      runner._to_device(h_target), abi.PR_MEM_DEVICE, runner.device)
            res["levels"], res["rmse"] = lv, rm
        out = to_host(res)
        out["iterations"] = out.pop("iters")
        return out
