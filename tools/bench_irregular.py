#!/usr/bin/env python
"""Throughput of the IrregularSection (polyline) node pass: a roughness ensemble of the 13-node companion reach
(tests/golden/irregular.in.npz) in the fused kernel and on the tile kernels.

    python tools/bench_irregular.py [--members 16384]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=16384)
    ap.add_argument("--case", default="irregular")
    a = ap.parse_args()
    import torch

    from flow_sim_b200 import abi
    from flow_sim_b200.flatten import load_flat
    from flow_sim_b200.runner import PreparedCall

    flat = load_flat(os.path.join(REPO, "tests", "golden", f"{a.case}.in.npz"))
    M = a.members
    flat.member_n_main = np.linspace(0.031, 0.04, M)          # (n = 0.0306-0.0309 diverges in the reference too)
    lib = abi.load_library()
    import ctypes as C

    for lanes, name in ((0, "fused kernel"), (-1, "tile kernels")):
        best = None
        for _ in range(3):
            call = PreparedCall(flat, M, abi.PR_OUT_UPSTREAM, abi.PR_MEM_DEVICE, "cuda:0", lanes=lanes, want_error=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            abi.check(lib, lib.pr_ensemble_run(*call.args(), C.c_void_p(0)), "pr_ensemble_run")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        r = call.results()
        its = int(r["iters"].sum().item())
        print(json.dumps({"case": a.case, "path": name, "members": M, "nodes": flat.n_nodes, "ms": best, "newton_iterations": its,
                          "failed": int((r["status"] != 0).sum().item()), "node_iterations_per_s": its * flat.n_nodes / (best * 1e-3)}))


if __name__ == "__main__":
    main()
