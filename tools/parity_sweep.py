#!/usr/bin/env python
"""Large parity sweep (not part of the timed bench): K members of the 65,536-member gerd roughness grid on the GPU
vs the CPU oracle on all host cores.  Reports max relative error and the number of members / level-steps whose
Newton iteration count differs (expected: rare near-tolerance flips).

    python tools/parity_sweep.py --members 2048
"""
import argparse
import json
import os
import sys
from multiprocessing import get_context

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


def _work(n_values):
    import oracle_py
    from bench import load_case

    flat = load_case()
    M = len(n_values)
    flat.member_n_main = np.asarray(n_values)
    h, q, _ = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=M)
    flat.ic_depth, flat.ic_flow = h, q
    o = oracle_py.run(flat, n_members=M, out_mode=1)
    return o["depth"], o["flow"], o["iters"], o["status"], o["final_error"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=2048)
    a = ap.parse_args()
    from bench import load_case, member_roughness
    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner, to_host

    total = 65536
    idx = np.linspace(0, total - 1, a.members).round().astype(np.int64)
    n = member_roughness(idx, total)
    flat = load_case()
    res = to_host(EnsembleRunner(flat, "cuda:0").roughness_sweep(n, out_mode=abi.PR_OUT_UPSTREAM))
    cores = os.cpu_count() or 1
    chunks = [n[c::cores] for c in range(cores)]
    with get_context("fork").Pool(cores) as pool:
        parts = pool.map(_work, chunks)
    depth = np.empty_like(res["depth"]); flow = np.empty_like(res["flow"]); iters = np.empty_like(res["iters"])
    status = np.empty_like(res["status"]); ferr = np.empty(res["iters"].shape)
    for c, (d, f, it, st, fe) in enumerate(parts):
        depth[c::cores], flow[c::cores], iters[c::cores], status[c::cores], ferr[c::cores] = d, f, it, st, fe
    diff = res["iters"] != iters
    same = ~diff.any(axis=1)
    out = {
        "members": a.members, "level_steps": int(iters.size), "status_equal": bool(np.array_equal(status, res["status"])),
        "iteration_count_mismatches": int(diff.sum()), "members_with_mismatch": int((~same).sum()),
        "max_rel_depth_on_matching_members": float(np.max(np.abs(res["depth"][same] - depth[same]) / np.abs(depth[same]))),
        "max_rel_flow_on_matching_members": float(np.max(np.abs(res["flow"][same] - flow[same]) / np.abs(flow[same]))),
        "max_rel_depth_all": float(np.max(np.abs(res["depth"] - depth) / np.abs(depth))),
    }
    if diff.any():
        m, k = np.argwhere(diff)[0]
        out["first_mismatch"] = {"member": int(m), "level": int(k + 1), "gpu_iters": int(res["iters"][m, k]),
                                 "oracle_iters": int(iters[m, k]), "oracle_final_error": float(ferr[m, k]), "tol": flat.tol}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
