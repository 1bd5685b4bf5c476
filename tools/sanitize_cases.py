#!/usr/bin/env python
"""A few small runs of every kernel family in one process (fused 32/16/8-lane, rare-boundary, polyline, tiled path,
GVF, objective) - a quick smoke of a fresh build, and the workload to put under compute-sanitizer where that is
available:

    python tools/sanitize_cases.py
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import util  # noqa: E402
from flow_sim_b200 import abi  # noqa: E402
from flow_sim_b200.runner import gvf_initial_conditions, rating_objective, run_flat  # noqa: E402


def main():
    for case, M, lanes in (("example", 5, 0), ("akbari", 3, 0), ("gerd_calib_m0", 2, 0), ("gerd_calib_m0", 2, -1),
                           ("storage_general", 2, 0), ("gerd_gated", 2, 0), ("irregular", 3, 0), ("irregular", 2, -1),
                           ("irregular_curved", 1, 0)):
        flat = util.golden_inputs(case)
        if case.startswith("gerd_calib"):
            flat.n_levels = 5
            flat.up.series = flat.up.series[:5]
        if case in ("example", "akbari", "irregular"):
            flat.member_n_main = np.linspace(0.025, 0.035, M)
        out = run_flat(flat, n_members=M, lanes=lanes)
        print(case, M, lanes, "status", out["status"].tolist(), "iterations", int(out["iters"].sum()))
    flat = util.golden_inputs("gerd_calib_m0")
    flat.member_n_main = np.array([0.02, 0.04])
    h, q, st = gvf_initial_conditions(flat, 2, flat.meta["initial_flow"], flat.meta["downstream_depth"])
    lv, rm = rating_objective(flat.n_levels, q[:, :flat.n_levels].copy(), h[:, :flat.n_levels].copy(), flat.meta["z0"],
                              [1562.5, 3850.0], [497.5, 500.0])
    print("gvf", st.tolist(), "objective", rm.tolist())


if __name__ == "__main__":
    main()
