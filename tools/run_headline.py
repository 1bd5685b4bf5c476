#!/usr/bin/env python
"""One launch of the headline kernel on `--members` members of the gerd roughness grid (for ncu captures / quick A-B
timings of tuning builds via PR_B200_LIB).  Prints the Newton kernel's time (CUDA events, best / mean of --reps)."""
import argparse
import copy
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=8192)
    ap.add_argument("--total", type=int, default=65536, help="grid the members are drawn from (evenly spaced)")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--tag", default="")
    ap.add_argument("--order", default="none", choices=["none", "desc"], help="desc: hand out rough (expensive) members first")
    a = ap.parse_args()
    import torch

    from bench import load_case, member_roughness
    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner
    from flow_sim_b200.runner import gvf_initial_conditions

    flat = load_case()
    M = a.members
    idx = np.linspace(0, a.total - 1, M).round().astype(np.int64)
    runner = EnsembleRunner(flat, "cuda:0")
    dev = runner.device
    n_dev = torch.from_numpy(member_roughness(idx, a.total)).to(dev)
    f = copy.copy(runner.flat)
    f.member_n_main = n_dev
    ich, icq, _ = gvf_initial_conditions(f, M, flat.meta["initial_flow"], flat.meta["downstream_depth"], abi.PR_MEM_DEVICE, dev, None)
    order = torch.argsort(n_dev, descending=True).to(torch.int32) if a.order == "desc" else None
    times = []
    for r in range(a.reps + (1 if a.reps > 1 else 0)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = runner.solve(M, member_n_main=n_dev, ic_depth=ich, ic_flow=icq, out_mode=abi.PR_OUT_UPSTREAM, member_order=order)
        e1.record()
        torch.cuda.synchronize()
        if r > 0 or a.reps == 1:
            times.append(e0.elapsed_time(e1))
    it = int(res["iters"].sum().item())
    print(json.dumps({"tag": a.tag, "order": a.order, "lanes": os.environ.get("PR_FORCE_LANES", "auto"), "lib": os.environ.get("PR_B200_LIB", "default"), "members": M, "ms_best": min(times),
                      "ms_mean": float(np.mean(times)), "iterations": it, "failed": int((res["status"] != 0).sum().item()),
                      "node_iterations_per_s": it * flat.n_nodes / (min(times) * 1e-3),
                      "iters_checksum": int((res["iters"].long() * torch.arange(1, res["iters"].numel() + 1, device=dev).view_as(res["iters"]) % 1000003).sum().item())}))


if __name__ == "__main__":
    main()
