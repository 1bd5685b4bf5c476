#!/usr/bin/env python
"""Newton iteration-count parity of the GPU path on the WHOLE 65,536-member headline grid.

Joins the GPU dump (tools/grid_gpu.py -> gpurun_out/grid_gpu.npz) with the oracle's run of the same grid
(tools/oracle_grid.py -> oracle/_build/gerd_grid65536.full.npz) and reports every level-step whose iteration count
differs, with the oracle's ||R|| at the decision that flipped (the accepted iterate's norm when the GPU needed one
more iteration, the norm one iteration earlier when it needed one fewer), the relative distance of that norm from
tol, and how far the flipped members' hydrographs are from the oracle's.

    python tools/grid_flips.py [--gpu gpurun_out/grid_gpu.npz] [--out profiles/r02_parity_sweep_65536.json]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", default=os.path.join(REPO, "gpurun_out", "grid_gpu.npz"))
    ap.add_argument("--oracle", default=os.path.join(REPO, "oracle", "_build", "gerd_grid65536.full.npz"))
    ap.add_argument("--out", default=None)
    ap.add_argument("--kernel", default="")
    a = ap.parse_args()
    g = np.load(a.gpu)
    o = np.load(a.oracle)
    tol = 1e-6
    gi, oi = g["iters"].astype(np.int32), o["iters"].astype(np.int32)
    T, K = oi.shape
    assert gi.shape == oi.shape
    diff = gi != oi
    members = np.nonzero(diff.any(axis=1))[0]
    flips = []
    for m in members:
        k = int(np.argmax(diff[m]))                 # first differing level-step of the member: the flip itself
        d = int(gi[m, k] - oi[m, k])
        # GPU one more: the oracle accepted an iterate whose norm the GPU saw above tol -> oracle final_error just below tol
        # GPU one fewer: the GPU accepted one iteration earlier -> oracle prev_error just above tol
        deciding = float(o["final_error"][m, k] if d > 0 else o["prev_error"][m, k])
        rel_h = float(np.max(np.abs(g["depth"][m] - o["depth"][m]) / np.abs(o["depth"][m])))
        rel_q = float(np.max(np.abs(g["flow"][m] - o["flow"][m]) / np.abs(o["flow"][m])))
        flips.append({"member": int(m), "level": k + 1, "gpu_iters": int(gi[m, k]), "oracle_iters": int(oi[m, k]),
                      "oracle_norm_at_decision": deciding, "rel_distance_from_tol": abs(deciding - tol) / tol,
                      "later_level_steps_differing": int(diff[m, k + 1:].sum()),
                      "max_rel_depth": rel_h, "max_rel_flow": rel_q,
                      "rmse_rel": float(abs(g["rmse"][m] - o["rmse"][m]) / abs(o["rmse"][m]))})
    same = ~diff.any(axis=1)
    rel_h = np.abs(g["depth"][same] - o["depth"][same]) / np.abs(o["depth"][same])
    rel_q = np.abs(g["flow"][same] - o["flow"][same]) / np.abs(o["flow"][same])
    # how far apart are the two implementations' norms where they took the same decision?
    ge, oe = g["final_error"][same], o["final_error"][same]
    rel_norm = np.abs(ge - oe) / oe
    # near ties in the oracle's own record (a different rounding of the iterate can move these decisions)
    fe, pe = o["final_error"], o["prev_error"]
    ties = {f"{w:g}": int((((fe >= tol * (1 - w)) & (fe < tol)) | ((pe >= tol) & (pe <= tol * (1 + w)))).sum())
            for w in (1e-3, 1e-4, 1e-5, 1e-6, 1e-7)}
    out = {
        "kernel": a.kernel, "members": int(T), "level_steps": int(T * K), "tol": tol,
        "status_equal": bool(np.array_equal(g["status"], o["status"])),
        "failed_members": int((o["status"] != 0).sum()),
        "iteration_total_gpu": int(gi.sum()), "iteration_total_oracle": int(oi.sum()),
        "members_with_a_flip": int(len(members)), "level_steps_differing": int(diff.sum()),
        "flips": flips,
        "max_rel_distance_from_tol_of_a_flip": max([f["rel_distance_from_tol"] for f in flips], default=0.0),
        "max_rel_depth_flipped_members": max([f["max_rel_depth"] for f in flips], default=0.0),
        "max_rel_flow_flipped_members": max([f["max_rel_flow"] for f in flips], default=0.0),
        "matching_members": {"count": int(same.sum()), "max_rel_depth": float(rel_h.max()), "max_rel_flow": float(rel_q.max()),
                             "max_rel_rmse": float(np.max(np.abs(g["rmse"][same] - o["rmse"][same]) / np.abs(o["rmse"][same])))},
        "norm_deviation_gpu_vs_oracle": {"median": float(np.median(rel_norm)), "p99": float(np.quantile(rel_norm, 0.99)),
                                         "p99.99": float(np.quantile(rel_norm, 0.9999)), "max": float(rel_norm.max())},
        "oracle_near_ties_by_relative_width": ties,
    }
    txt = json.dumps(out, indent=1)
    print(txt if len(flips) < 12 else json.dumps({k: v for k, v in out.items() if k != "flips"}, indent=1))
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
