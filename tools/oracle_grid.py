#!/usr/bin/env python
"""The CPU oracle over the WHOLE 65,536-member gerd roughness grid (BASELINE configs[3]) - test infrastructure.

Writes the per-member, per-level Newton iteration counts, the calibration RMSE and the status of every member of the
headline grid to tests/golden/gerd_grid65536.oracle.npz (small: committed; the GPU parity test holds the device path
to it), and the bulky by-products (upstream depth / flow series, ||R|| at acceptance) to oracle/_build/ (git-ignored)
for tools/grid_flips.py.  About 50 minutes on 8 cores.

    python tools/oracle_grid.py [--procs 8] [--members 65536]
"""
import argparse
import os
import sys
import time
from multiprocessing import get_context

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

CHUNK = 128
NEAR_TIE = 1e-3


def _work(args):
    lo, hi, total = args
    import oracle_py
    from bench import H_TARGET, Q_QUERY, load_case, member_roughness

    flat = load_case()
    M = hi - lo
    flat.member_n_main = member_roughness(np.arange(lo, hi), total)
    h, q, ic_status = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=M)
    flat.ic_depth, flat.ic_flow = h, q
    o = oracle_py.run(flat, n_members=M, out_mode=1, trace_prev_error=True)
    _, rm = oracle_py.objective(flat.n_levels, o["flow"], o["depth"], flat.meta["z0"], Q_QUERY, H_TARGET)
    return lo, hi, o["depth"], o["flow"], o["iters"], o["status"], o["final_error"], rm, o["prev_error"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--members", type=int, default=65536)
    a = ap.parse_args()
    import oracle_py
    from bench import load_case

    oracle_py.build()
    flat = load_case()
    L = flat.n_levels
    T = a.members
    depth = np.empty((T, L)); flow = np.empty((T, L)); ferr = np.empty((T, L - 1)); perr = np.empty((T, L - 1))
    iters = np.empty((T, L - 1), dtype=np.int32); status = np.empty(T, dtype=np.int32); rmse = np.empty(T)
    tasks = [(lo, min(lo + CHUNK, T), T) for lo in range(0, T, CHUNK)]
    t0 = time.time()
    done = 0
    with get_context("fork").Pool(a.procs) as pool:
        for lo, hi, d, f, it, st, fe, rm, pe in pool.imap_unordered(_work, tasks):
            depth[lo:hi], flow[lo:hi], iters[lo:hi], status[lo:hi], ferr[lo:hi], rmse[lo:hi], perr[lo:hi] = d, f, it, st, fe, rm, pe
            done += hi - lo
            if (done // CHUNK) % 32 == 0:
                print(f"{done}/{T} members, {time.time() - t0:.0f} s", flush=True)
    assert iters.max() < 127
    suffix = "" if T == 65536 else f"_{T}"
    # near ties of the convergence test ||R|| < tol: level-steps where the accepted iterate's norm lies within
    # NEAR_TIE (relative) below tol, or the norm one iteration earlier lies within NEAR_TIE above it.  A different
    # rounding of the iterate can move such a decision by one iteration.
    tol = flat.tol
    tie = ((ferr >= tol * (1 - NEAR_TIE)) & (ferr < tol)) | ((perr >= tol) & (perr <= tol * (1 + NEAR_TIE)))
    tm, tk = np.nonzero(tie)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", f"gerd_grid65536{suffix}.oracle.npz"),
                        iters=iters.astype(np.int8), status=status.astype(np.int8), rmse=rmse, tol=tol,
                        near_tie=NEAR_TIE, tie_member=tm.astype(np.int32), tie_level=(tk + 1).astype(np.int8),
                        tie_final_error=ferr[tm, tk], tie_prev_error=perr[tm, tk])
    os.makedirs(os.path.join(REPO, "oracle", "_build"), exist_ok=True)
    np.savez(os.path.join(REPO, "oracle", "_build", f"gerd_grid65536{suffix}.full.npz"), depth=depth, flow=flow,
             final_error=ferr, prev_error=perr, iters=iters, status=status, rmse=rmse)
    print(f"done: {T} members in {time.time() - t0:.0f} s; iterations {int(iters.sum())}; failed {int((status != 0).sum())}")


if __name__ == "__main__":
    main()
