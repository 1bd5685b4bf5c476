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
