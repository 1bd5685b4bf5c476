"""Generate the committed golden fixtures under tests/golden/ from the LIVE reference.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py outputs [case ...]   # slow: runs the reference (config 3 = ~11 min)
    python oracle/make_golden.py inputs  [case ...]   # fast: flattens the reference objects

``outputs`` writes ``tests/golden/<case>.ref.npz``  (depth, flow, iters, [storage_stage], seconds and
the first per-iteration (J.data, R, delta, x) captures from an ``spsolve`` hook).
``inputs``  writes ``tests/golden/<case>.in.npz``   (the flattened SoA description of the same case,
produced by ``flow_sim_b200.flatten`` from the *reference's own objects*), so the GPU box - where the
reference does not exist - can replay exactly the inputs the reference saw.
"""
from __future__ import annotations

import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

# config-4 ensemble grid: n_main_m = 0.020 + 0.040*m/65535
CALIB_MEMBERS = [0, 8191, 21845, 30000, 36408, 43690, 54321, 65535]


def calib_n(m: int) -> float:
    return 0.020 + 0.040 * m / 65535


# one release scenario (a member of a gate-state ensemble): lower pool, jammed gates, narrower blend buffer
RELEASE_SCENARIO = dict(initial_roseires_level=486.2, rating_kwargs=dict(jammed_spillways=2, jammed_sluice_gates=1, buffer=0.3))

CASES = ["example", "akbari", "gerd_full", "akbari_long", "storage_general", "gerd_release", "gerd_gated", "gerd_gated_full", "irregular", "irregular_curved", "irregular_pocket", "mixed_sections"] + [f"gerd_calib_m{m}" for m in CALIB_MEMBERS]


def build(case: str):
    import ref_harness as rh

    if case == "example":
        return rh.build_example()
    if case == "akbari":
        return rh.build_akbari()
    if case == "storage_general":
        return rh.build_storage_general()
    if case == "akbari_long":
        # reduced clone of config 5: prismatic akbari channel, N=2001, dx=100, dt=600, theta=0.6
        return rh.build_akbari(length=200000, spatial_step=100, time_step=600, duration=16 * 600, theta=0.6)
    if case == "gerd_release":
        return rh.build_gerd(n_main=0.03, calibration=True, **RELEASE_SCENARIO)
    if case == "gerd_gated":     # RoseiresRatingCurve(smooth=False): stateful gate control; starting with the gates
        # open and a 2 h cool-down makes them close (level 5) and re-open (level 25) in the middle of Newton loops
        return rh.build_gerd(n_main=0.03, calibration=True,
                             rating_kwargs=dict(smooth=False, initially_open=True, max_cooldown=7200))
    if case == "irregular":          # IrregularSection polylines with composite roughness
        return rh.build_irregular()
    if case == "irregular_curved":   # polyline sections on a curved centre line (curvature slope Sc with FD derivatives)
        return rh.build_irregular(curved=True)
    if case == "irregular_levee":    # inputs only: a bar splits low flows into two equal sub-channels; the reference's
        # own Newton iteration diverges on it (negative depths, ZeroDivisionError in level 1)
        return rh.build_irregular(levee=True)
    if case == "irregular_pocket":   # split flow the reference survives: a side pocket behind a ridge (75 of 117
        # node-levels run on the multi-sub-channel conveyance, cross_section.py:374-439)
        return rh.build_irregular(pocket=True)
    if case == "mixed_sections":     # compound trapezoid upstream, polyline downstream, blended node by node (cross_section.py:933-969)
        return rh.build_mixed()
    if case == "gerd_gated_full":   # config 3 (curvature, 16 days) with gate control: dozens of open/close cycles
        return rh.build_gerd(calibration=False, rating_kwargs=dict(smooth=False))
    if case == "gerd_full":
        return rh.build_gerd(calibration=False)
    if case.startswith("gerd_calib_m"):
        return rh.build_gerd(n_main=calib_n(int(case[len("gerd_calib_m"):])), calibration=True)
    raise KeyError(case)


def do_outputs(case: str):
    import ref_harness as rh

    solver, kw = build(case)
    res = rh.run_and_record(solver, kw, capture_iterations=6)
    out = dict(depth=res["depth"], flow=res["flow"], iters=res["iters"], seconds=np.float64(res["seconds"]))
    if "storage_stage" in res:
        out["storage_stage"] = res["storage_stage"]
    for j, c in enumerate(res.get("captures", [])):
        if c["J"].size > 4000:      # keep fixtures small: no captures for the long reach
            break
        for key in ("J", "R", "delta", "x"):
            out[f"cap{j}_{key}"] = c[key]
        out[f"cap{j}_level"] = np.int32(c["level"])
    if case.startswith("gerd_calib"):
        out["calib_levels"] = rh.gerd_calibration_levels(solver, res, rh.CALIB_Q)
        out["calib_rmse"] = np.float64(np.mean((out["calib_levels"] - rh.CALIB_H_TARGET) ** 2) ** 0.5)
    if case == "akbari_long":        # 2001 nodes x 17 levels: keep as is (544 KB) -> trim to float64 arrays only
        pass
    path = os.path.join(GOLD, f"{case}.ref.npz")
    np.savez_compressed(path, **out)
    return case, float(res["seconds"]), int(res["iters"].sum())


def do_inputs(case: str):
    from flow_sim_b200.flatten import flatten_solver, save_flat

    solver, kw = build(case)
    flat = flatten_solver(solver, tolerance=kw.get("tolerance", 1e-4), max_iter=kw.get("max_iter", 100))
    save_flat(os.path.join(GOLD, f"{case}.in.npz"), flat)
    return case


if __name__ == "__main__":
    mode = sys.argv[1]
    cases = sys.argv[2:] or CASES
    os.makedirs(GOLD, exist_ok=True)
    fn = do_outputs if mode == "outputs" else do_inputs
    with Pool(min(len(cases), os.cpu_count() or 1)) as pool:
        for r in pool.imap_unordered(fn, cases):
            print(r, flush=True)
