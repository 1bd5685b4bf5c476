#!/usr/bin/env python
"""bench.py - Preissmann node-steps/s on the gerd_roseires Manning-n calibration ensemble (BASELINE.json
configs[3]: 65,536 members x 121 nodes x 32 time steps, FP64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

One "step" = one pass of the hot path over one batch of members: GVF initial profile per member
(it depends on the roughness), the whole implicit time loop (32 levels, Newton + block-tridiagonal solve per
level), and the calibration objective.  N > 1: one process per GPU under torchrun, members sharded across
ranks (weak scaling: --members is per GPU), no traffic inside the time loop, one NCCL all_gather of the
per-member RMSE at the end of each step.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)



def load_case():
    """The gerd_roseires calibration set-up (cases/gerd_roseires/n_calibrate.py:5-17) built on the mirror API and
    flattened; bit-identical to the inputs flattened from the reference's own objects (tests/test_mirror_api.py)."""
    from flow_sim_b200.cases import build_gerd
    from flow_sim_b200.flatten import flatten_solver

    solver, kw = build_gerd(n_main=0.020, calibration=True)
    return flatten_solver(solver, tolerance=kw["tolerance"])

Q_QUERY = np.array([1562.5, 3850, 6000, 10000, 14000, 21000.0])     # cases/gerd_roseires/n_calibrate.py:30
H_TARGET = np.array([497.5, 500, 502, 505, 507, 510.0])             # cases/gerd_roseires/n_calibrate.py:29
METRIC = "Preissmann node-steps/s (members x nodes x steps)"
UNIT = "node-steps/s"
# algorithmic FP64 flops per node per Newton iteration (FMA = 2, every other op = 1)
F_ITER_INBANK, F_ITER_OVERBANK = 136.0, 162.0


def member_roughness(members: np.ndarray, total: int) -> np.ndarray:
    """n_main_m = 0.020 + 0.040*m/(total-1): the deterministic calibration grid."""
    return 0.020 + 0.040 * np.asarray(members, dtype=np.float64) / max(total - 1, 1)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[6]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (or, in the build container, the live Python reference) on the host cores
# ----------------------------------------------------------------------------------------------

def _cpu_worker(args):
    n_values, = args
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_py
    flat = load_case()
    M = len(n_values)
    flat.member_n_main = np.asarray(n_values, dtype=np.float64)
    t0 = time.perf_counter()
    ich, icq, _ = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=M)
    flat.ic_depth, flat.ic_flow = ich, icq
    out = oracle_py.run(flat, n_members=M)
    oracle_py.objective(flat.n_levels, out["flow"][:, :, 0], out["depth"][:, :, 0], flat.meta["z0"], Q_QUERY, H_TARGET)
    dt = time.perf_counter() - t0
    return dt, int(out["iters"].sum())


def cpu_sample(total_members: int, per_core: int, cores: int):
    """Times `cores*per_core` evenly spaced members of the ensemble, one process per core.
    Returns (node-steps/s aggregate, wall seconds, members, iterations)."""
    from multiprocessing import get_context

    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_py

    oracle_py.build()
    flat = load_case()
    n_sample = cores * per_core
    idx = np.linspace(0, total_members - 1, n_sample).round().astype(np.int64)
    n_all = 0.020 + 0.040 * idx / max(total_members - 1, 1)
    chunks = [(n_all[c::cores],) for c in range(cores)]
    ctx = get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    node_steps = n_sample * flat.n_nodes * (flat.n_levels - 1)
    return node_steps / wall, wall, n_sample, sum(r[1] for r in res)


def run_reference_arm(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    cores = os.cpu_count() or 1
    total = args.members * args.gpus
    per_core = max(1, args.cpu_members_per_core)
    vals, walls, iters = [], [], 0
    for s in range(args.warmup + args.steps):
        v, w, n_sample, it = cpu_sample(total, per_core, cores)
        if s >= args.warmup:
            vals.append(v); walls.append(w); iters += it
    value = float(np.mean(vals))
    flat = load_case()
    sample = (f"{cores * per_core} evenly spaced members of the {total}-member ensemble per step "
              f"({per_core} per core), C port of the reference algorithm (oracle/preissmann_oracle.c); the "
              "reference itself is pure Python and is not present on the GPU box")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(walls) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"gerd_roseires Manning-n calibration ensemble: {total} members x {flat.n_nodes} nodes x "
                               f"{flat.n_levels - 1} steps (bounded sample, see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_gpu_arm(args) -> dict | None:
    import torch
    import torch.distributed as dist

    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner, gather_members, shard_members
    from flow_sim_b200.runner import gvf_initial_conditions, rating_objective

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = abi.load_library()
    flat = load_case()
    N, L = flat.n_nodes, flat.n_levels
    M = args.members                      # per GPU (weak scaling)
    total = M * world
    runner = EnsembleRunner(flat, dev)
    n_host = torch.from_numpy(member_roughness(shard_members(total, rank, world), total)).pin_memory()   # round-robin deal
    n_dev = n_host.to(dev)
    q_dev = torch.from_numpy(Q_QUERY).to(dev)
    h_dev = torch.from_numpy(H_TARGET).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step_device(timed: dict | None):
        """Inputs resident in HBM.  Returns the result dict (device tensors)."""
        e = [ev() for _ in range(5)]
        e[0].record()
        f = runner.flat
        import copy

        f = copy.copy(f)
        f.member_n_main = n_dev
        ich, icq, _ = gvf_initial_conditions(f, M, flat.meta["initial_flow"], flat.meta["downstream_depth"],
                                             abi.PR_MEM_DEVICE, dev, stream)
        e[1].record()
        res = runner.solve(M, member_n_main=n_dev, ic_depth=ich, ic_flow=icq, out_mode=abi.PR_OUT_UPSTREAM, stream=stream)
        e[2].record()
        lv, rm = rating_objective(L, res["flow"], res["depth"], flat.meta["z0"], q_dev, h_dev, abi.PR_MEM_DEVICE, dev, stream)
        e[3].record()
        if world > 1:
            res["rmse_all"] = gather_members(rm, total, rank, world)      # the one collective of the run
        e[4].record()
        res["rmse"] = rm
        if timed is not None:
            timed["events"] = e
        return res

    def step_e2e():
        """Through the public API with HOST buffers: pinned H2D of the per-member inputs, D2H of the results."""
        res = runner.roughness_sweep(n_host, q_query=q_dev, h_target=h_dev, out_mode=abi.PR_OUT_UPSTREAM, stream=stream)
        rmse = res["rmse"].to("cpu", non_blocking=False)
        status = res["status"].to("cpu", non_blocking=False)
        return rmse, status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        res = step_device(None)
    barrier()

    # ---- timed: device-resident ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = lib.pr_launch_count()
    t_wall0 = time.time()
    step_ms, gvf_ms, solve_ms, obj_ms = [], [], [], []
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        barrier()
        timed = {}
        res = step_device(timed)
        barrier()
        e = timed["events"]
        step_ms.append(e[0].elapsed_time(e[4]))
        gvf_ms.append(e[0].elapsed_time(e[1]))
        solve_ms.append(e[1].elapsed_time(e[2]))
        obj_ms.append(e[2].elapsed_time(e[3]))
    t_wall1 = time.time()
    launches = lib.pr_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)      # max over ranks
    total_s = float(total_ms.item()) * 1e-3
    node_steps_per_step = total * N * (L - 1)
    value = node_steps_per_step * args.steps / total_s

    # ---- timed: end to end through the public API (host buffers) ----
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        a, b = ev(), ev()
        a.record()
        rmse_host, status_host = step_e2e()
        b.record()
        barrier()
        e2e_ms.append(a.elapsed_time(b))
    e2e_total = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_value = node_steps_per_step * args.steps / (float(e2e_total.item()) * 1e-3)

    # ---- work model for the roofline ----
    iters_sum = int(res["iters"].sum().item())                # Newton iterations of this rank's members
    n_bad = int((res["status"] != 0).sum().item())
    # over-bank share of node evaluations from a 32-member full-output sample
    sidx = torch.linspace(0, M - 1, min(32, M)).round().long()
    samp = runner.roughness_sweep(n_dev[sidx.to(dev)], out_mode=abi.PR_OUT_FULL, stream=stream)
    hb = torch.from_numpy(flat.geom["h_bank"]).to(dev)
    over = ((samp["depth"] > hb) & (torch.from_numpy(flat.geom["kind"]).to(dev) == abi.PR_XS_COMPOUND)).double().mean().item()
    f_iter = F_ITER_OVERBANK * over + F_ITER_INBANK * (1.0 - over)
    flops_per_launch = iters_sum * N * f_iter
    solve_s = float(np.mean(solve_ms)) * 1e-3
    tf = abi.C.c_double(0.0)
    abi.check(lib, lib.pr_fp64_peak(200.0, abi.C.byref(tf)), "pr_fp64_peak")
    achieved_tf = flops_per_launch / solve_s / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out_bytes = M * L * 16.0 + M * (L - 1) * 4.0 + M * 8.0    # boundary series + iteration counts + status
    traffic, pipe_pct = None, None
    try:
        prof = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        traffic = prof.get("ensemble_kernel_dram_bytes_per_launch")
        pipe_pct = prof.get("ensemble_kernel_fp64_pipe_pct_ncu")
    except Exception:
        pass

    # ---- parity spot check of this very run against the CPU oracle (outside every timed region) ----
    parity = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import oracle_py

        pick = np.linspace(0, M - 1, 8).round().astype(int)
        f2 = load_case()
        f2.member_n_main = n_host.numpy()[pick]
        ich, icq, _ = oracle_py.gvf(f2, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=len(pick))
        f2.ic_depth, f2.ic_flow = ich, icq
        ora = oracle_py.run(f2, n_members=len(pick), out_mode=abi.PR_OUT_UPSTREAM)
        got_h = res["depth"][pick].cpu().numpy(); got_q = res["flow"][pick].cpu().numpy()
        parity = {
            "members_checked": len(pick),
            "max_rel_depth": float(np.max(np.abs(got_h - ora["depth"]) / np.abs(ora["depth"]))),
            "max_rel_flow": float(np.max(np.abs(got_q - ora["flow"]) / np.abs(ora["flow"]))),
            "iterations_equal": bool(np.array_equal(res["iters"][pick].cpu().numpy(), ora["iters"])),
            "against": "oracle/preissmann_oracle.c (C port pinned to the live reference)",
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, w, n_sample, _ = cpu_sample(total, args.cpu_members_per_core, cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_sample} evenly spaced members of the ensemble ({w:.1f} s wall on {cores} processes); "
                                  "C port of the reference algorithm - the Python reference itself measured 79-129 "
                                  "node-steps/s per core on this case (BASELINE.md)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"gerd_roseires Manning-n calibration ensemble (BASELINE configs[3]): {M} members per GPU x "
                               f"{N} nodes x {L - 1} steps, n_main = 0.020..0.060, GVF initial profile per member",
                   "members_per_gpu": M, "members_total": total, "nodes": N, "time_steps": L - 1,
                   "parallelism": f"members dealt round-robin over {world} GPU(s), no traffic in the time loop, one all_gather of RMSE" if world > 1 else "1 GPU",
                   "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events summed",
                   "failed_members": n_bad},
        "kernel_ms": {"gvf_initial_conditions": float(np.mean(gvf_ms)), "ensemble_newton": float(np.mean(solve_ms)),
                      "objective": float(np.mean(obj_ms))},
        "newton_iterations_per_step": iters_sum / (M * (L - 1)),
        "node_iterations_per_s": iters_sum * N * world / solve_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(M * 8), "d2h_bytes_per_step": int(M * 8 + M * 4),
                "ms_per_step": float(e2e_total.item()) / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s",
                     "frac": achieved_tf / tf.value if tf.value > 0 else None, "traffic": traffic,
                     "kernel": "pr_ensemble_kernel<G=32, M=4, W=16, CURV=0, RM=1, EXACT=1>",
                     "flops_per_node_iteration": f_iter, "overbank_share": over,
                     "fp64_pipe_pct_ncu": pipe_pct,     # from the committed ncu capture (profiles/), not this run

                     "peak_source": "pr_fp64_peak: register-resident DFMA microbenchmark measured in this run "
                                    "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2)",
                     "hbm_view": {"algorithmic_bytes_per_launch": out_bytes, "achieved_gbs": out_bytes / solve_s / 1e9,
                                  "peak_gbs": hbm_peak, "note": "state stays on chip; HBM is not the bound"}},
        "cpu_baseline": cpu_baseline,
        "parity": parity,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=65536, help="ensemble members per GPU")
    ap.add_argument("--cpu-members-per-core", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        line = run_reference_arm(args)
    else:
        line = run_gpu_arm(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
