"""GPU tests of the long-reach path (N > 249: tiled block cyclic reduction over global state, BASELINE configs[4])
against the CPU oracle, the reference golden of the 2001-node clone, and the fused kernel on the same inputs."""
import copy

import numpy as np
import pytest

import util
from flow_sim_b200 import abi
from flow_sim_b200.runner import run_flat

pytestmark = pytest.mark.gpu


def _vs_oracle(flat, M, what, lanes=0):
    import oracle_py

    ora = oracle_py.run(flat, n_members=M)
    out = run_flat(flat, n_members=M, lanes=lanes)
    assert np.array_equal(out["status"], ora["status"]), what
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], what)
    assert np.array_equal(out["iters"], ora["iters"]), f"{what}: Newton iteration counts differ"
    return out, ora


def test_akbari_long_clone_matches_reference_golden():
    """N = 2001, dx = 100 m, dt = 600 s, theta = 0.6, 16 steps: the reduced clone of config 5."""
    flat = util.golden_inputs("akbari_long")
    ref = util.golden_outputs("akbari_long")
    out = run_flat(flat)
    assert out["status"][0] == 0
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], "akbari_long")
    assert np.array_equal(out["iters"][0], ref["iters"])


@pytest.mark.parametrize("case", ["example", "akbari", "gerd_calib_m0", "gerd_calib_m65535", "gerd_full", "gerd_gated_full"])
def test_long_path_equals_fused_kernel_on_short_reaches(case):
    """Forcing the tiled path on the shipped cases (storage, normal-depth and Roseires boundaries, compound
    sections, centre-line curvature over 384 levels, gate control): same results as the fused kernel and as the
    reference."""
    flat = util.golden_inputs(case)
    ref = util.golden_outputs(case)
    out = run_flat(flat, lanes=-1)
    assert out["status"][0] == 0
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], case)
    assert np.array_equal(out["iters"][0], ref["iters"])
    if "storage_stage" in ref.files:
        assert util.max_rel(out["storage_stage"][0], ref["storage_stage"]) <= util.RTOL


@pytest.mark.parametrize("n_nodes", [250, 257, 385, 1000, 3969, 4098, 8066])
def test_tile_and_chain_boundaries(n_nodes):
    """Sizes around the tile (128 cells) and chain (31 lanes) boundaries, compound sections (slow linear
    convergence: ~20 iterations per level)."""
    from test_gpu_ensemble import _prismatic

    _vs_oracle(_prismatic(kind="compound", n_nodes=n_nodes, levels=3), 1, f"N={n_nodes}")


def test_inflow_scenarios_on_a_long_reach():
    flat = util.golden_inputs("akbari_long")
    M = 5
    base = flat.up.series
    peak = np.linspace(0.6, 1.4, M)[:, None]
    flat.up.series = base[None, 0] + (base[None, :] - base[0]) * peak
    out, ora = _vs_oracle(flat, M, "long-reach scenarios")
    # members are independent: a member alone gives the same bits as inside the batch
    one = copy.copy(flat); one.up = copy.copy(flat.up); one.up.series = flat.up.series[3:4]
    solo = run_flat(one, n_members=1)
    assert np.array_equal(solo["depth"][0], out["depth"][3]) and np.array_equal(solo["iters"][0], out["iters"][3])


def test_general_storage_on_the_long_path():
    """Area curve + outflow curve + head losses through the tiled path: the reference's own run of that case."""
    full = util.golden_inputs("storage_general")
    ref = util.golden_outputs("storage_general")
    out, _ = _vs_oracle(full, 1, "general storage with losses, long path", lanes=-1)
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], "storage_general, long path")
    assert util.max_rel(out["storage_stage"][0], ref["storage_stage"]) <= util.RTOL
    a = copy.copy(full); a.down = copy.copy(full.down); a.down.storage_losses = False
    _vs_oracle(a, 1, "general storage, long path", lanes=-1)


def test_non_convergence_on_the_long_path():
    flat = util.golden_inputs("gerd_calib_m0")
    flat.max_iter = 11
    out, ora = None, None
    import oracle_py

    ora = oracle_py.run(flat)
    out = run_flat(flat, lanes=-1)
    assert out["status"][0] == abi.PR_STATUS_MAX_ITER and out["fail_level"][0] == ora["fail_level"][0] == 6
    assert np.array_equal(np.isnan(out["depth"]), np.isnan(ora["depth"]))
    assert np.array_equal(out["iters"], ora["iters"])


def test_config5_size_two_steps():
    """100 000 nodes (BASELINE configs[4]); 2 members x 2 steps, one of them against the oracle."""
    import oracle_py
    from flow_sim_b200.cases.akbari_firoozi import build_long_reach
    from flow_sim_b200.flatten import flatten_solver

    solver, kw = build_long_reach(n_nodes=100_000, n_steps=2)
    flat = flatten_solver(solver, tolerance=kw["tolerance"])
    s = flat.up.series
    flat.up.series = np.stack([s, s[0] + (s - s[0]) * 1.5])
    out = run_flat(flat, n_members=2, out_mode=abi.PR_OUT_FULL)
    assert not out["status"].any()
    one = copy.copy(flat); one.up = copy.copy(flat.up); one.up.series = flat.up.series[1]
    ora = oracle_py.run(one, n_members=1)
    util.assert_parity(out["depth"][1], out["flow"][1], ora["depth"][0], ora["flow"][0], "config 5")
    assert np.array_equal(out["iters"][1], ora["iters"][0])


def test_roughness_members_on_the_tiled_path():
    """Per-member n_main / n_fp on the long-reach kernels: the config-4 members forced onto the tiled path equal the
    reference goldens, and a 600-node compound reach with both overrides equals the oracle."""
    from test_gpu_ensemble import _prismatic

    flat = util.golden_inputs("gerd_calib_m0")
    ms = [0, 36408, 65535]
    flat.member_n_main = np.array([util.calib_n(m) for m in ms])
    flat.ic_depth = np.stack([util.golden_inputs(f"gerd_calib_m{m}").ic_depth for m in ms])
    flat.ic_flow = np.stack([util.golden_inputs(f"gerd_calib_m{m}").ic_flow for m in ms])
    out = run_flat(flat, n_members=3, lanes=-1)
    for i, m in enumerate(ms):
        ref = util.golden_outputs(f"gerd_calib_m{m}")
        util.assert_parity(out["depth"][i], out["flow"][i], ref["depth"], ref["flow"], f"tiled path, member {m}")
        assert np.array_equal(out["iters"][i], ref["iters"])
    flat = _prismatic(kind="compound", n_nodes=600, levels=3)
    flat.member_n_main = np.array([0.025, 0.03, 0.035, 0.04, 0.045])
    flat.member_n_fp = np.array([0.05, 0.06, 0.07, 0.08, 0.09])
    _vs_oracle(flat, 5, "N=600 with per-member n_main and n_fp")


def test_gvf_profile_on_a_reach_too_long_for_the_shared_memory_stage():
    """N = 3000 (> 1219 nodes): the GVF kernel reads the derived geometry from a global table; per-member roughness,
    flow and downstream depth; against the oracle's restatement of Channel._gvh_conditions."""
    import oracle_py
    from flow_sim_b200.runner import gvf_initial_conditions
    from test_gpu_ensemble import _prismatic

    flat = _prismatic(kind="compound", n_nodes=3000, levels=2)
    M = 5
    flat.member_n_main = np.array([0.025, 0.03, 0.035, 0.04, 0.045])
    q0 = np.array([40.0, 60.0, 80.0, 100.0, 120.0])
    hd = np.array([1.5, 2.0, 2.5, 3.0, 3.5])
    h, q, st = gvf_initial_conditions(flat, M, q0, hd)
    ho, qo, sto = oracle_py.gvf(flat, q0, hd, n_members=M)
    assert np.array_equal(st, sto)
    assert util.max_rel(h, ho) <= 1e-12 and np.array_equal(q, qo)


def test_long_path_with_more_members_than_a_grid_dimension_holds():
    """the tiled path put the members on gridDim.y (limit 65,535), so 65,536 members failed after the whole
    time loop.  Members now sit on gridDim.x: 65,536 + 3 members of the akbari reach forced onto the tiled path, a few
    of them against the fused kernel."""
    flat = util.golden_inputs("akbari")
    M = 65536 + 3
    base = np.array(flat.up.series)
    scale = np.linspace(0.6, 1.4, M)
    flat.up.series = base[0] + (base - base[0])[None, :] * scale[:, None]
    out = run_flat(flat, n_members=M, out_mode=abi.PR_OUT_UPSTREAM, lanes=-1)
    assert not out["status"].any()
    pick = np.array([0, 1, 65534, 65535, 65536, M - 1])
    f2 = util.golden_inputs("akbari")
    f2.up.series = flat.up.series[pick]
    ref = run_flat(f2, n_members=len(pick), out_mode=abi.PR_OUT_UPSTREAM)
    util.assert_parity(out["depth"][pick], out["flow"][pick], ref["depth"], ref["flow"], "tiled path, 65,539 members")
    assert np.array_equal(out["iters"][pick], ref["iters"])


def test_two_streams_run_long_reaches_side_by_side():
    """The long-reach path hands out pooled workspaces instead of holding a process-wide lock: two device-memory runs
    enqueued on two streams (neither call waits for the GPU) both come out right, and a third run reuses a workspace."""
    import torch

    from flow_sim_b200.ensemble import EnsembleRunner, to_host

    flat = util.golden_inputs("akbari_long")                 # 2 001 nodes: the tiled path
    ref = util.golden_outputs("akbari_long")
    runner = EnsembleRunner(flat, "cuda:0")
    L = flat.n_levels
    series = np.tile(np.array(flat.up.series)[None, :], (4, 1))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        a = runner.solve(4, up_series=series, out_mode=abi.PR_OUT_FULL, stream=s1.cuda_stream)
    with torch.cuda.stream(s2):
        b = runner.solve(4, up_series=series, out_mode=abi.PR_OUT_FULL, stream=s2.cuda_stream)
    torch.cuda.synchronize()
    c = runner.solve(4, up_series=series, out_mode=abi.PR_OUT_FULL)
    torch.cuda.synchronize()
    for res in (to_host(a), to_host(b), to_host(c)):
        assert not res["status"].any()
        for m in range(4):
            util.assert_parity(res["depth"][m], res["flow"][m], ref["depth"], ref["flow"], "two streams")
            assert np.array_equal(res["iters"][m], ref["iters"])
    assert abi.load_library().pr_long_last_trips() > 0


def test_tiled_path_on_a_reach_whose_node_count_is_a_multiple_of_four():
    """N % 4 == 0 takes the 256-bit state accesses of the tile kernel (and N - 1 = 1999 cells leave a ragged last tile):
    the vectorised prismatic builder + device normal depth against the oracle, scenarios with different flood peaks."""
    import oracle_py
    from flow_sim_b200.cases.akbari_firoozi import build_long_reach_flat, flood_wave_series
    from flow_sim_b200.runner import normal_depth_initial_conditions

    flat = build_long_reach_flat(n_nodes=2000, n_steps=6)
    h, q = normal_depth_initial_conditions(flat, 1, flat.meta["initial_flow"])
    flat.ic_depth, flat.ic_flow = h[0], q[0]
    flat.up.series = flood_wave_series([120.0, 200.0, 260.0, 300.0, 175.0], flat.n_levels, flat.dt)
    ora = oracle_py.run(flat, n_members=5, out_mode=abi.PR_OUT_FULL)
    out = run_flat(flat, n_members=5, out_mode=abi.PR_OUT_FULL)
    assert not out["status"].any() and not ora["status"].any()
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], "2000-node reach")
    assert np.array_equal(out["iters"], ora["iters"])
