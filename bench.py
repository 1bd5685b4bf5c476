#!/usr/bin/env python
"""bench.py - Preissmann node-steps/s on the gerd_roseires Manning-n calibration ensemble (BASELINE.json
configs[3]: 65,536 members x 121 nodes x 32 time steps, FP64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

One "step" = one pass of the hot path over one batch of members: GVF initial profile per member (it depends on the
roughness), the whole implicit time loop (32 levels, Newton + block-tridiagonal solve per level), and the calibration
objective.  N > 1: one process per GPU under torchrun, members dealt round-robin over the ranks, no traffic inside the
time loop, one NCCL all_gather at the end of each step.

The JSON line (rank 0) carries, beside the contract's keys:
  value / e2e     weak scaling: --members (65,536) per GPU; e2e = the same through the public API with host buffers
  e2e_full        e2e with everything the API hands back copied to the host (iteration counts, upstream series)
  strong          the north star's own configuration: 65,536 members IN TOTAL over the N GPUs, gather of RMSE +
                  iteration counts + status + upstream series (43.8 MB)
  config5         BASELINE configs[4]: 100,000-node prismatic reach x 1,024 inflow scenarios (long-reach path), with its
                  HBM roofline (48 algorithmic bytes per node-iteration)
  single_runs     configs 1-3 as single-member runs on the GPU next to the reference's own wall times
  cpu_baseline    the unmodified Python reference on the host cores when it is staged (oracle/_ref), and the C port
  parity          members of this very run against the oracle; iteration counts of the WHOLE grid against the committed
                  oracle record (tests/golden/gerd_grid65536.oracle.npz)
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
GRID = 65536                                                        # members of the headline calibration grid
# algorithmic FP64 flops per node per Newton iteration (FMA = 2, every other op = 1)
F_ITER_INBANK, F_ITER_OVERBANK = 136.0, 162.0
# the reference's own wall times of configs 1-3 (one core of the build container)
REFERENCE_SECONDS = {"example": 0.347, "akbari_firoozi": 0.383, "gerd_roseires": 674.0}


def member_roughness(members: np.ndarray, total: int) -> np.ndarray:
    """n_main_m = 0.020 + 0.040*m/(total-1): the deterministic calibration grid."""
    return 0.020 + 0.040 * np.asarray(members, dtype=np.float64) / max(total - 1, 1)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled DURING the timed region: NVML queries from a thread of this
    process (nvidia_ml_py) - an `nvidia-smi -lms` child was seen to stall the running kernel by 3-40 ms at every sample
    (it enumerates every GPU of the box per query); nvidia-smi is only the fall-back where NVML does not import."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int, period_s: float = 0.1):
        self.index = index
        self.period = period_s
        self.rows = []            # (time, sm_mhz, max_mhz, reasons set, power_w)
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.source = None

    def _nvml_loop(self, nv, handle):
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        mx = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                pw = float(nv.nvmlDeviceGetPowerUsage(handle)) / 1000.0
                self.rows.append((time.time(), sm, mx, {k for k, bit in names.items() if mask & bit}, pw))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            handle = None
            try:                                   # the CUDA device's own UUID: right whatever CUDA_VISIBLE_DEVICES says
                import torch

                handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)).encode())
            except Exception:
                handle = None
            if handle is None:
                visible = (os.environ.get("CUDA_VISIBLE_DEVICES") or "").split(",")
                entry = visible[self.index].strip() if self.index < len(visible) else ""
                if entry.startswith("GPU-"):
                    handle = nv.nvmlDeviceGetHandleByUUID(entry.encode())
                else:
                    handle = nv.nvmlDeviceGetHandleByIndex(int(entry) if entry.isdigit() else self.index)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            try:
                reasons = {name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6])
                           if val.lower().startswith("active")}
                self.rows.append((time.time(), float(parts[0]), float(parts[1]), reasons, float(parts[6])))
            except Exception:
                continue

    def stop(self, t0: float, t1: float) -> dict:
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"]}
        time.sleep(0.15)
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.2]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "samples": 0, "source": self.source}
        reasons = set()
        for r in rows:
            reasons |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": float(max(r[4] for r in rows)), "source": self.source}


# ----------------------------------------------------------------------------------------------
# CPU arm: the C port of the reference algorithm (oracle/), and - where oracle/stage_reference.py has staged it - the
# unmodified Python reference itself, on the host cores
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
    """Times `cores*per_core` evenly spaced members of the ensemble on the C port, one process per core.
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


def _pyref_worker(n_main):
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import ref_harness

    secs, iters = ref_harness.time_member(float(n_main))
    return secs, iters


def python_reference_sample(total_members: int, cores: int, nodes: int, steps: int) -> dict | None:
    """One member per host core through the UNMODIFIED reference (PreissmannSolver.run of cve-mohd/flow-sim, staged by
    oracle/stage_reference.py; numpy/scipy/pandas/scikit-learn from the image).  None when it is not staged or does
    not import.  About 30-50 s of wall time."""
    from multiprocessing import get_context

    sys.path.insert(0, os.path.join(REPO, "oracle"))
    try:
        import ref_harness

        if not ref_harness.reference_available():
            return None
        idx = np.linspace(0, total_members - 1, cores).round().astype(np.int64)
        n_all = member_roughness(idx, total_members)
        t0 = time.perf_counter()
        with get_context("spawn").Pool(cores) as pool:      # spawn: the harness changes directory and patches pandas
            res = pool.map(_pyref_worker, list(n_all))
        wall = time.perf_counter() - t0
    except Exception as exc:                                 # missing scipy / sklearn on the box, ...
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    run_s = [r[0] for r in res]
    value = cores * nodes * steps / max(run_s)              # solver.run() only, as SURVEY.md 8d prescribes
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "per_core": value / cores,
            "sample": f"{cores} evenly spaced members of the {total_members}-member ensemble, one per core, through the "
                      f"unmodified Python reference (PreissmannSolver.run only; slowest member {max(run_s):.1f} s, pool wall "
                      f"{wall:.1f} s incl. imports and set-up); Newton iterations {sum(r[1] for r in res)}"}


def run_reference_arm(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    cores = os.cpu_count() or 1
    total = args.members * args.gpus
    per_core = max(1, args.cpu_members_per_core)
    flat = load_case()
    pyref = None if args.no_python_reference else python_reference_sample(total, cores, flat.n_nodes, flat.n_levels - 1)
    vals, walls, iters = [], [], 0
    for s in range(args.warmup + args.steps):
        v, w, n_sample, it = cpu_sample(total, per_core, cores)
        if s >= args.warmup:
            vals.append(v); walls.append(w); iters += it
    value = float(np.mean(vals))
    sample = (f"{cores * per_core} evenly spaced members of the {total}-member ensemble per step "
              f"({per_core} per core), C port of the reference algorithm (oracle/preissmann_oracle.c) - about 100x faster "
              "per core than the Python reference, so a ratio against this arm understates the speed-up over the reference")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(walls) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"gerd_roseires Manning-n calibration ensemble: {total} members x {flat.n_nodes} nodes x "
                               f"{flat.n_levels - 1} steps (bounded sample, see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "per_core": value / cores, "sample": sample},
        "python_reference": pyref,
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
    from flow_sim_b200.ensemble import EnsembleRunner, PinnedResults, gather_members, gather_packed, nvtx_range, shard_members
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
    with nvtx_range("flatten: case -> SoA"):
        flat = load_case()
    N, L = flat.n_nodes, flat.n_levels
    runner = EnsembleRunner(flat, dev)
    q_dev = torch.from_numpy(Q_QUERY).to(dev)
    h_dev = torch.from_numpy(H_TARGET).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_step(members_local: int, total: int, full_gather: bool):
        """One step with the per-member inputs resident in HBM.  Returns (step(timed) -> result dict, n_host)."""
        import copy

        n_host = torch.from_numpy(member_roughness(shard_members(total, rank, world), total)).pin_memory()   # round-robin deal
        n_dev = n_host.to(dev)

        def step(timed: dict | None):
            e = [ev() for _ in range(5)]
            e[0].record()
            f = copy.copy(runner.flat)
            f.member_n_main = n_dev
            with nvtx_range("gvf initial conditions"):
                ich, icq, _ = gvf_initial_conditions(f, members_local, flat.meta["initial_flow"], flat.meta["downstream_depth"],
                                                     abi.PR_MEM_DEVICE, dev, stream)
            e[1].record()
            with nvtx_range("newton time loop"):
                order = torch.argsort(n_dev, descending=True).to(torch.int32)     # rough (expensive) members first
                res = runner.solve(members_local, member_n_main=n_dev, ic_depth=ich, ic_flow=icq, out_mode=abi.PR_OUT_UPSTREAM,
                                   stream=stream, member_order=order)
            e[2].record()
            with nvtx_range("rating objective"):
                lv, rm = rating_objective(L, res["flow"], res["depth"], flat.meta["z0"], q_dev, h_dev, abi.PR_MEM_DEVICE, dev, stream)
            e[3].record()
            res["rmse"] = rm
            if world > 1:
                with nvtx_range("gather"):                                     # the one collective of the run
                    if full_gather:
                        res["gathered"] = gather_packed(res, total, rank, world)
                    else:
                        res["rmse_all"] = gather_members(rm, total, rank, world)
            e[4].record()
            if timed is not None:
                timed["events"] = e
            return res

        return step, n_host

    def time_steps(step, n_steps: int, warm: int):
        import gc

        for _ in range(warm):
            res = step(None)
        barrier()
        per = {"step": [], "gvf": [], "solve": [], "obj": [], "gather": []}
        # a generational garbage collection of the interpreter between two launches of a step shows up as GPU idle time
        # inside the step's events (seen: +40 ms on one step in ten): collect now, not in the timed steps
        gc.collect()
        gc.disable()
        try:
            return _timed(step, n_steps, per)
        finally:
            gc.enable()

    def _timed(step, n_steps, per):
        res = None
        for _ in range(n_steps):
            res = None                          # drop the previous step's result tensors first: the caching allocator then
            flush.zero_()                       # reuses their blocks instead of calling cudaMalloc between two launches
            barrier()                           # (flush: L2 between timed iterations, not timed)
            timed = {}
            res = step(timed)
            barrier()
            e = timed["events"]
            per["step"].append(e[0].elapsed_time(e[4])); per["gvf"].append(e[0].elapsed_time(e[1]))
            per["solve"].append(e[1].elapsed_time(e[2])); per["obj"].append(e[2].elapsed_time(e[3]))
            per["gather"].append(e[3].elapsed_time(e[4]))
        return res, per

    # ---- headline: weak scaling, --members per GPU, inputs resident in HBM ----
    M = args.members
    total = M * world
    step_device, n_host = make_step(M, total, full_gather=False)
    n_dev = n_host.to(dev)
    sampler = ClockSampler(local, period_s=float(os.environ.get("PR_BENCH_CLOCK_PERIOD", "0.1")))
    for _ in range(max(args.warmup, 3)):
        res = step_device(None)
    barrier()
    sampler.start()
    time.sleep(1.0)                         # nvidia-smi's own start-up (NVML init) can stall a running kernel: keep it out of the timed steps
    for _ in range(2):
        step_device(None)                   # the GPU idled during that second: back to load clocks
    barrier()
    launches0 = lib.pr_launch_count()
    t_wall0 = time.time()
    res, per = time_steps(step_device, args.steps, 0)
    t_wall1 = time.time()
    launches = lib.pr_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    total_s = max_over_ranks(sum(per["step"])) * 1e-3
    node_steps_per_step = total * N * (L - 1)
    value = node_steps_per_step * args.steps / total_s
    solve_ms = per["solve"]

    # ---- end to end through the public API (host buffers) ----
    pinned = PinnedResults()
    def step_e2e(full: bool):
        """Pinned H2D of the per-member inputs, the sweep, D2H of the results: RMSE + status, or (full) everything the
        API hands back - iteration counts and the upstream stage / discharge series as well."""
        r = runner.roughness_sweep(n_host, q_query=q_dev, h_target=h_dev, out_mode=abi.PR_OUT_UPSTREAM, stream=stream)
        with nvtx_range("d2h: results"):            # into page-locked buffers, one synchronisation
            out = pinned.fetch(r, keys=("rmse", "status", "iters", "depth", "flow") if full else ("rmse", "status"))
        return out

    def time_e2e(full: bool, n_steps: int):
        import gc

        for _ in range(2):
            step_e2e(full)
        barrier()
        ms = []
        gc.collect()
        gc.disable()                     # (see time_steps)
        try:
            for _ in range(n_steps):
                flush.zero_()
                barrier()
                a, b = ev(), ev()
                a.record()
                step_e2e(full)
                b.record()
                barrier()
                ms.append(a.elapsed_time(b))
        finally:
            gc.enable()
        tot = max_over_ranks(sum(ms))
        return node_steps_per_step * n_steps / (tot * 1e-3), tot / n_steps

    e2e_value, e2e_ms = time_e2e(False, args.steps)
    e2e_full_value, e2e_full_ms = time_e2e(True, max(2, min(args.steps, 5)))
    d2h_small = int(M * 8 + M * 4)
    d2h_full = int(d2h_small + M * (L - 1) * 4 + 2 * M * L * 8)

    # ---- strong scaling: the north star's configuration, 65,536 members in total over the N GPUs ----
    strong = None
    if args.no_strong:
        pass
    elif total != GRID or world > 1:
        if GRID % world == 0:
            s_local = GRID // world
            step_strong, _ = make_step(s_local, GRID, full_gather=True)
            res_s, per_s = time_steps(step_strong, args.steps, 3)
            s_total = max_over_ranks(sum(per_s["step"])) * 1e-3
            gathered = res_s.get("gathered")
            strong = {"members_total": GRID, "members_per_gpu": s_local, "n_gpus": world,
                      "value": GRID * N * (L - 1) * args.steps / s_total, "unit": UNIT, "ms_per_step": s_total * 1e3 / args.steps,
                      "kernel_ms": {"gvf_initial_conditions": float(np.mean(per_s["gvf"])), "ensemble_newton": float(np.mean(per_s["solve"])),
                                    "objective": float(np.mean(per_s["obj"])), "gather": float(np.mean(per_s["gather"]))},
                      "gather": "one all_gather of RMSE + iteration counts + status + upstream stage / discharge series",
                      "gather_bytes": int(gathered["bytes_per_member"]) * GRID if gathered else 0}
    else:
        strong = {"members_total": GRID, "members_per_gpu": M, "n_gpus": 1, "value": value, "unit": UNIT,
                  "ms_per_step": total_s * 1e3 / args.steps, "note": "at one GPU the strong and the weak configuration coincide"}

    # ---- work model for the roofline ----
    iters_sum = int(res["iters"].sum().item())                # Newton iterations of this rank's members
    n_bad = int((res["status"] != 0).sum().item())
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
    traffic, prof_note = None, None
    try:
        prof = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        traffic = prof.get("ensemble_kernel_dram_bytes_per_launch")
        prof_note = prof.get("note")
    except Exception:
        pass

    # ---- parity of this very run (outside every timed region) ----
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
        # every member of this rank against the oracle's record of the whole 65,536-member grid
        gold_path = os.path.join(REPO, "tests", "golden", "gerd_grid65536.oracle.npz")
        if total == GRID and os.path.exists(gold_path):
            gold = np.load(gold_path)
            mine = shard_members(total, rank, world)
            gi, oi = res["iters"].cpu().numpy(), gold["iters"][mine].astype(np.int32)
            diff = gi != oi
            flipped = np.nonzero(diff.any(axis=1))[0]
            tol = float(gold["tol"])
            ties = {(int(m), int(k)): (float(fe), float(pe)) for m, k, fe, pe in
                    zip(gold["tie_member"], gold["tie_level"], gold["tie_final_error"], gold["tie_prev_error"])}
            flips = []
            for j in flipped:
                k = int(np.argmax(diff[j])) + 1
                d = int(gi[j, k - 1]) - int(oi[j, k - 1])
                fe, pe = ties.get((int(mine[j]), k), (float("nan"), float("nan")))
                norm = fe if d > 0 else pe
                flips.append({"member": int(mine[j]), "level": k, "gpu_iters": int(gi[j, k - 1]), "oracle_iters": int(oi[j, k - 1]),
                              "oracle_norm_at_decision": norm, "rel_distance_from_tol": abs(norm - tol) / tol})
            rel_rmse = np.abs(res["rmse"].cpu().numpy() - gold["rmse"][mine]) / np.abs(gold["rmse"][mine])
            parity["whole_grid"] = {"members": int(len(mine)), "level_steps": int(gi.size), "iteration_flips": len(flips), "flips": flips,
                                    "max_rel_rmse_other_members": float(np.max(np.delete(rel_rmse, flipped))) if len(flipped) < len(mine) else None,
                                    "against": "tests/golden/gerd_grid65536.oracle.npz (tools/oracle_grid.py: the oracle over the whole grid)"}

    # ---- BASELINE configs[4]: the long-reach path ----
    config5 = None
    if not args.no_config5:
        config5 = bench_config5(args, torch, dist, dev, rank, world, hbm_peak, max_over_ranks, barrier)

    # ---- configs 1-3 as single runs; CPU baselines (rank 0, one GPU) ----
    single = None
    cpu_baseline, cpu_port = None, None
    if rank == 0 and world == 1:
        if not args.no_single_runs:
            single = bench_single_runs(torch)
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, w, n_sample, _ = cpu_sample(total, args.cpu_members_per_core, cores)
            cpu_port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "per_core": v / cores,
                        "sample": f"{n_sample} evenly spaced members of the ensemble ({w:.1f} s wall on {cores} processes), "
                                  "C port of the reference algorithm (oracle/preissmann_oracle.c)"}
            pyref = None if args.no_python_reference else python_reference_sample(total, cores, N, L - 1)
            if pyref and "value" in pyref:
                cpu_baseline = pyref
            else:
                cpu_baseline = dict(cpu_port)
                if pyref:
                    cpu_baseline["python_reference"] = pyref

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
        "kernel_ms": {"gvf_initial_conditions": float(np.mean(per["gvf"])), "ensemble_newton": float(np.mean(solve_ms)),
                      "objective": float(np.mean(per["obj"]))},
        "step_ms": [round(v, 3) for v in per["step"]],
        "newton_iterations_per_step": iters_sum / (M * (L - 1)),
        "node_iterations_per_s": iters_sum * N * world / solve_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(M * 8), "d2h_bytes_per_step": d2h_small,
                "ms_per_step": e2e_ms, "copies_back": "calibration RMSE and status per member (what the calibration loop consumes)"},
        "e2e_full": {"value": e2e_full_value, "unit": UNIT, "h2d_bytes_per_step": int(M * 8), "d2h_bytes_per_step": d2h_full,
                     "ms_per_step": e2e_full_ms,
                     "copies_back": "everything the API returns: RMSE, status, Newton iteration counts per level, upstream stage and discharge series"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s",
                     "frac": achieved_tf / tf.value if tf.value > 0 else None, "traffic": traffic,
                     "kernel": "pr_ensemble_kernel<G=32, M=4, W=16, CURV=0, RM=1, EXACT=1> (persistent warps)",
                     "flops_per_node_iteration": f_iter, "overbank_share": over,
                     "traffic_source": prof_note,
                     "peak_source": "pr_fp64_peak: register-resident DFMA microbenchmark measured in this run "
                                    "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2)",
                     "hbm_view": {"algorithmic_bytes_per_launch": out_bytes, "achieved_gbs": out_bytes / solve_s / 1e9,
                                  "peak_gbs": hbm_peak, "note": "state stays on chip; HBM is not the bound"}},
        "strong": strong,
        "config5": config5,
        "single_runs": single,
        "cpu_baseline": cpu_baseline,
        "cpu_port": cpu_port,
        "parity": parity,
    }


def bench_config5(args, torch, dist, dev, rank, world, hbm_peak, max_over_ranks, barrier) -> dict | None:
    """BASELINE configs[4]: prismatic channel, 100,000 nodes x 1,024 inflow scenarios x 16 steps through the long-reach
    path (tile kernels + chain kernel, trip loop as a CUDA-graph WHILE node).  The 1,024 scenarios are dealt over the
    ranks (strong scaling); one gather of the upstream series at the end."""
    from flow_sim_b200 import abi
    from flow_sim_b200.cases.akbari_firoozi import build_long_reach_flat, flood_wave_series
    from flow_sim_b200.ensemble import EnsembleRunner, gather_members, shard_members
    from flow_sim_b200.runner import normal_depth_initial_conditions

    lib = abi.load_library()
    nodes, total, steps = args.config5_nodes, args.config5_members, 16
    if total % world:
        return None
    torch.cuda.synchronize()
    t0 = time.time()
    flat = build_long_reach_flat(n_nodes=nodes, n_steps=steps)           # array arithmetic, no per-node objects
    L = flat.n_levels
    mine = shard_members(total, rank, world)
    series = flood_wave_series(100.0 + 200.0 * mine / max(total - 1, 1), L, flat.dt)       # Q_p,m = 100 + 200 m/(M-1)
    ich, icq = normal_depth_initial_conditions(flat, 1, flat.meta["initial_flow"], mem=abi.PR_MEM_DEVICE, device=dev)
    flat.ic_depth, flat.ic_flow = ich[0].cpu().numpy(), icq[0].cpu().numpy()
    setup_s = time.time() - t0
    runner = EnsembleRunner(flat, dev)
    ser_dev = torch.from_numpy(series).to(dev)
    m_local = len(mine)
    ms = []
    res = None
    for r in range(1 + max(2, min(args.steps, 3))):
        res = None                       # (free the previous result before the next call allocates its own)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = runner.solve(m_local, up_series=ser_dev, out_mode=abi.PR_OUT_UPSTREAM)
        if world > 1:
            gather_members(res["depth"], total, rank, world)
        e1.record()
        barrier()
        if r > 0:
            ms.append(e0.elapsed_time(e1))
    trips = int(lib.pr_long_last_trips())
    secs = max_over_ranks(float(np.mean(ms))) * 1e-3
    it_t = res["iters"].sum().double().reshape(1)
    bad_t = (res["status"] != 0).sum().double().reshape(1)
    if world > 1:
        dist.all_reduce(it_t); dist.all_reduce(bad_t)
    iters = int(it_t.item())
    node_iters = iters * nodes
    out = {
        "workload": f"prismatic channel (BASELINE configs[4]): {nodes} nodes x {total} inflow scenarios x {L - 1} steps, "
                    f"dx = 100 m, dt = 600 s, theta = 0.6; {m_local} scenarios per GPU",
        "n_gpus": world, "ms": secs * 1e3, "value": total * nodes * (L - 1) / secs, "unit": UNIT,
        "node_iterations_per_s": node_iters / secs, "newton_iterations_per_step": iters / (total * (L - 1)),
        "newton_trips": trips, "kernel_launches_per_trip": 5, "failed_members": int(bad_t.item()),
        "roofline": {"bound": "hbm", "algorithmic_bytes_per_node_iteration": 48,
                     "achieved": node_iters * 48 / secs / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                     "frac": node_iters * 48 / secs / 1e9 / hbm_peak / world,
                     "algorithmic_bytes_per_launch": 48.0 * nodes * m_local,
                     "traffic": 74.0 * nodes * m_local, "moved_bytes_per_node_iteration": 64,
                     "note": "48 B = read level-k state + read and write the iterate (SURVEY.md 8d); the fused tile kernel moves 64 B "
                             "(level-k state kept as four constants per cell) plus 32 B where a level is accepted - measured "
                             "1.24 GB read + 0.43-0.83 GB written per trip of 25.6 M nodes (profiles/r02d_ncu_long_fused_launches.csv); "
                             "traffic = 74 B per node per launch of the tile kernel, from that capture scaled to this launch; "
                             "the kernel is bound by FP64 latency, not by HBM"},
        "host_setup_s": setup_s,
    }
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import copy

        import oracle_py

        pick = np.array([0])
        f2 = copy.copy(flat); f2.up = copy.copy(flat.up); f2.up.series = series[pick]
        ora = oracle_py.run(f2, n_members=1, out_mode=abi.PR_OUT_UPSTREAM)
        gh = res["depth"][pick].cpu().numpy()
        out["parity"] = {"members_checked": 1, "max_rel_depth": float(np.max(np.abs(gh - ora["depth"]) / np.abs(ora["depth"]))),
                         "iterations_equal": bool(np.array_equal(res["iters"][pick].cpu().numpy(), ora["iters"])),
                         "against": "oracle/preissmann_oracle.c on the same member"}
    lib.pr_release_workspace()
    return out


def bench_single_runs(torch) -> dict:
    """BASELINE configs[0..2] - the three shipped cases as single-member runs through the mirror API's
    PreissmannSolver.run() (host buffers, one member = one warp of the GPU): latency, not throughput."""
    from flow_sim_b200.cases import build_akbari, build_example, build_gerd

    out = {}
    for name, build in (("example", build_example), ("akbari_firoozi", build_akbari), ("gerd_roseires", lambda: build_gerd())):
        try:
            best = None
            for _ in range(3):
                solver, kw = build()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                solver.run(verbose=0, **kw)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            steps = solver.number_of_time_levels - 1
            out[name] = {"nodes": int(solver.number_of_nodes), "steps": int(steps), "newton_iterations": int(np.sum(solver.iterations)),
                         "run_seconds": best, "node_steps_per_s": solver.number_of_nodes * steps / best,
                         "reference_seconds": REFERENCE_SECONDS[name], "speedup_vs_reference": REFERENCE_SECONDS[name] / best}
        except Exception as exc:
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    out["note"] = ("wall time of PreissmannSolver.run() incl. flattening, H2D / D2H and one kernel launch; reference_seconds = "
                   "the reference's own run() on one core of the build container (BASELINE.md)")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=GRID, help="ensemble members per GPU (weak scaling)")
    ap.add_argument("--cpu-members-per-core", type=int, default=24)
    ap.add_argument("--config5-nodes", type=int, default=100_000)
    ap.add_argument("--config5-members", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-single-runs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        line = run_reference_arm(args)
    else:
        line = run_gpu_arm(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
