"""Seeded random configurations through the C ABI against the CPU oracle: reach length (every kernel family and the
tiled path), section kind, boundary types, member count, per-member roughness / inflow, forced lane widths."""
import numpy as np
import pytest

import util
from flow_sim_b200.runner import run_flat

pytestmark = pytest.mark.gpu


def _case(seed):
    from test_gpu_ensemble import _prismatic

    rng = np.random.default_rng(seed)
    n_nodes = int(rng.choice([rng.integers(2, 9), rng.integers(9, 33), rng.integers(33, 64), rng.integers(64, 126),
                              rng.integers(126, 250), rng.integers(250, 700)]))
    kind = str(rng.choice(["rect", "trapezoid", "compound"]))
    down = str(rng.choice(["normal_depth", "fixed_depth"]))
    up = str(rng.choice(["flow_hydrograph", "flow_hydrograph", "stage_hydrograph"]))
    flat = _prismatic(kind=kind, down=down, up=up, n_nodes=n_nodes, levels=int(rng.integers(2, 6)))
    M = int(rng.integers(1, 14))
    if rng.random() < 0.7:
        flat.member_n_main = rng.uniform(0.02, 0.045, M)
        if kind == "compound" and rng.random() < 0.5:
            flat.member_n_fp = rng.uniform(0.04, 0.08, M)
    if rng.random() < 0.6:
        base = np.array(flat.up.series)
        scale = rng.uniform(0.2, 1.5, M)
        flat.up.series = np.stack([base[0] + (base - base[0]) * s for s in scale])
    lanes = 0
    if n_nodes <= 8 and rng.random() < 0.5:
        lanes = int(rng.choice([16, 32]))
    elif n_nodes <= 16 and rng.random() < 0.5:
        lanes = int(rng.choice([16, 32]))
    elif n_nodes < 250 and rng.random() < 0.2:
        lanes = -1
    return flat, M, lanes, f"seed {seed}: N={n_nodes} {kind} {up}->{down} M={M} lanes={lanes}"


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_vs_oracle(seed):
    import oracle_py

    flat, M, lanes, what = _case(seed)
    ora = oracle_py.run(flat, n_members=M)
    out = run_flat(flat, n_members=M, lanes=lanes)
    assert np.array_equal(out["status"], ora["status"]), what
    assert np.array_equal(np.isnan(out["depth"]), np.isnan(ora["depth"])), what
    fin = ~np.isnan(ora["depth"])
    util.assert_parity(out["depth"][fin], out["flow"][fin], ora["depth"][fin], ora["flow"][fin], what)
    assert np.array_equal(out["iters"], ora["iters"]), what
