#!/usr/bin/env python
"""HBM roofline of the derived-results kernel (Solver.prepare_results arrays): 16 B in + 48 B out per (member, level, node)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from bench import load_case
from flow_sim_b200 import abi
from flow_sim_b200.ensemble import EnsembleRunner
from flow_sim_b200.runner import derived_results

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
flat = load_case()
dev = torch.device("cuda", 0)
runner = EnsembleRunner(flat, dev)
L, N = flat.n_levels, flat.n_nodes
depth = torch.rand((M, L, N), dtype=torch.float64, device=dev) * 10 + 5
flow = torch.rand((M, L, N), dtype=torch.float64, device=dev) * 5000 + 1000
times = []
for r in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = derived_results(runner.flat, depth, flow, abi.PR_MEM_DEVICE, dev)
    b.record(); torch.cuda.synchronize()
    if r: times.append(a.elapsed_time(b) * 1e-3)
secs = min(times)
bytes_ = M * L * N * 64.0
peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"kernel": "pr_derived_kernel", "elements": M * L * N, "seconds_incl_alloc": secs, "algorithmic_bytes": bytes_,
                  "achieved_gbs": bytes_ / secs / 1e9, "peak_gbs": peak, "frac": bytes_ / secs / 1e9 / peak,
                  "note": "time includes the output allocations and the geometry-table kernel of the call"}))
