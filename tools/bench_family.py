#!/usr/bin/env python
"""Throughput of the fused kernel families (nodes per lane M = 1, 2, 4, 8) on a synthetic compound reach:
node-iterations/s for a given reach length.  A tuning tool, not a bench line.

    python tools/bench_family.py --nodes 121 241 --members 16384
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, nargs="+", default=[121, 241])
    ap.add_argument("--members", type=int, default=16384)
    ap.add_argument("--levels", type=int, default=9)
    ap.add_argument("--lanes", type=int, default=0)
    a = ap.parse_args()
    import torch

    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner
    from test_gpu_ensemble import _prismatic

    for n in a.nodes:
        flat = _prismatic(kind="compound", n_nodes=n, levels=a.levels)
        runner = EnsembleRunner(flat, "cuda:0")
        nm = torch.from_numpy(np.linspace(0.025, 0.035, a.members)).cuda()
        best = 1e9
        for r in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            res = runner.solve(a.members, member_n_main=nm, out_mode=abi.PR_OUT_UPSTREAM)
            e1.record()
            torch.cuda.synchronize()
            if r:
                best = min(best, e0.elapsed_time(e1) * 1e-3)
        iters = int(res["iters"].sum().item())
        print(json.dumps(dict(nodes=n, members=a.members, seconds=best, newton_iterations=iters,
                              node_iterations_per_s=iters * n / best, ok=int((res["status"] == 0).sum().item()))))


if __name__ == "__main__":
    main()
