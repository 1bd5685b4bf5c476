#!/usr/bin/env python
"""The GPU path over the whole 65,536-member gerd roughness grid; dumps what tools/grid_flips.py compares with the
oracle's run of the same grid (tools/oracle_grid.py): iteration counts, status, RMSE, upstream series, ||R|| at acceptance.

    python tools/grid_gpu.py [--members 65536] [--out gpurun_out/grid_gpu.npz]
"""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=65536)
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "grid_gpu.npz"))
    a = ap.parse_args()
    import torch

    from bench import H_TARGET, Q_QUERY, load_case, member_roughness
    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner, to_host
    from flow_sim_b200.runner import gvf_initial_conditions, rating_objective
    import copy

    flat = load_case()
    T = a.members
    n = member_roughness(np.arange(T), T)
    runner = EnsembleRunner(flat, "cuda:0")
    dev = runner.device
    n_dev = torch.from_numpy(n).to(dev)
    f = copy.copy(runner.flat)
    f.member_n_main = n_dev
    t0 = time.time()
    ich, icq, _ = gvf_initial_conditions(f, T, flat.meta["initial_flow"], flat.meta["downstream_depth"], abi.PR_MEM_DEVICE, dev, None)
    res = runner.solve(T, member_n_main=n_dev, ic_depth=ich, ic_flow=icq, out_mode=abi.PR_OUT_UPSTREAM, want_error=True)
    _, rm = rating_objective(flat.n_levels, res["flow"], res["depth"], flat.meta["z0"], torch.from_numpy(Q_QUERY).to(dev),
                             torch.from_numpy(H_TARGET).to(dev), abi.PR_MEM_DEVICE, dev, None)
    res["rmse"] = rm
    r = to_host(res)
    print(f"{T} members on the GPU in {time.time() - t0:.2f} s (incl. copies); iterations {int(r['iters'].sum())}; "
          f"failed {int((r['status'] != 0).sum())}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    np.savez_compressed(a.out, iters=r["iters"].astype(np.int8), status=r["status"].astype(np.int8), rmse=r["rmse"],
                        depth=r["depth"], flow=r["flow"], final_error=r["final_error"])


if __name__ == "__main__":
    main()
