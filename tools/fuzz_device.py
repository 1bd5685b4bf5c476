#!/usr/bin/env python
"""Exploratory differential fuzzing of the device path against the oracle on the GPU box (no reference there; the oracle
is pinned to the reference by tests/golden/ and oracle/fuzz_reference.py).  Replays `tests/fuzz_cases.random_case` seeds
beyond the committed corpus, single runs in every lane packing plus a 13-member roughness ensemble, and prints every
disagreement (fate, failure level, values beyond 1e-9, iteration counts beyond a near tie).

    python tools/fuzz_device.py --seeds 320:900 [--out gpurun_out/fuzz_device.json]     (seeds >= 1000 are long reaches: slow)
"""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="320:900")
    ap.add_argument("--gerd", default="", help="seed range of random members / scenarios of the headline reach")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import fuzz_cases
    import oracle_py
    import util
    from flow_sim_b200.abi import PreissmannLibraryError
    from flow_sim_b200.flatten import flatten_solver
    from flow_sim_b200.runner import run_flat

    seeds = [s for part in a.seeds.split(",") if part for s in range(*(int(v) for v in part.split(":")))]
    problems, n_run, n_refused, n_members = [], 0, 0, 0
    for seed in seeds:
        if a.out and seed % 50 == 0:        # progress survives a time-out
            json.dump(dict(seeds=a.seeds, reached=seed, launches=n_run, members=n_members, problems=problems), open(a.out, "w"), indent=1)
        d = fuzz_cases.describe(seed)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                solver, kw, _ = fuzz_cases.random_case(fuzz_cases.mirror_namespace(), seed)
                flat = flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw["max_iter"])
        except Exception:
            n_refused += 1
            continue
        rng = np.random.default_rng(30_000 + seed)
        for M in (1, 13):
            if M > 1:
                flat.member_n_main = d["n_main"] * rng.uniform(0.7, 1.4, M)
                flat.member_n_fp = d["n_fp"] * rng.uniform(0.7, 1.4, M) if seed % 2 else None
            ora = oracle_py.run(flat, M, trace_prev_error=True)
            for lanes in ((0, 8, 16, 32) if M == 1 else ((0, 8, 16, 32)[seed % 4],)):
                try:
                    out = run_flat(flat, n_members=M, lanes=lanes)
                except PreissmannLibraryError as e:
                    if "instantiation holds" in str(e):
                        continue
                    problems.append(dict(seed=seed, M=M, lanes=lanes, what=f"library error: {e}"))
                    continue
                n_run += 1
                n_members += M
                tag = dict(seed=seed, M=M, lanes=lanes, family=d["family"], up=d["up"], down=d["down"], ic=d["ic"], N=flat.n_nodes)
                if not np.array_equal(out["status"] != 0, ora["status"] != 0) or not np.array_equal(out["fail_level"], ora["fail_level"]):
                    problems.append(dict(tag, what="fate differs", got=[out["status"].tolist(), out["fail_level"].tolist()],
                                         oracle=[ora["status"].tolist(), ora["fail_level"].tolist()]))
                    continue
                ok = np.nonzero(ora["status"] == 0)[0]
                if len(ok):
                    try:
                        util.assert_iteration_parity(out, ora, flat.tol, "x", members=ok)
                    except AssertionError as e:
                        problems.append(dict(tag, what=str(e)[:300]))
    # random members / scenarios of the headline reach (fuzz_cases.describe_gerd): single runs on the fused and on the
    # tiled path, and a 16-member roughness ensemble on each scenario
    from flow_sim_b200.cases import build_gerd

    for seed in [s for part in a.gerd.split(",") if part for s in range(*(int(v) for v in part.split(":")))]:
        kwargs = fuzz_cases.describe_gerd(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            solver, kw = build_gerd(**kwargs)
            flat = flatten_solver(solver, tolerance=kw["tolerance"])
        rng = np.random.default_rng(40_000 + seed)
        for M, lanes in ((1, 0), (1, -1), (16, 0)):
            if M > 1:
                flat.member_n_main = rng.uniform(0.018, 0.065, M)
                flat.member_n_fp = rng.uniform(0.03, 0.12, M) if seed % 2 else None
            ora = oracle_py.run(flat, M, trace_prev_error=True)
            out = run_flat(flat, n_members=M, lanes=lanes)
            n_run += 1
            n_members += M
            tag = dict(seed=f"g{seed}", M=M, lanes=lanes, scenario={k: v for k, v in kwargs.items() if k != "n_main"})
            if not np.array_equal(out["status"], ora["status"]) or not np.array_equal(out["fail_level"], ora["fail_level"]):
                problems.append(dict(tag, what="fate differs", got=out["status"].tolist(), oracle=ora["status"].tolist()))
                continue
            ok = np.nonzero(ora["status"] == 0)[0]
            try:
                util.assert_iteration_parity(out, ora, flat.tol, "x", members=ok)
            except AssertionError as e:
                problems.append(dict(tag, what=str(e)[:300]))
    rep = dict(seeds=a.seeds, gerd=a.gerd, launches=n_run, members=n_members, refused_at_setup=n_refused, problems=problems)
    print(json.dumps(rep, indent=1)[:6000])
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
