"""IrregularSection nodes through the C ABI: polyline sections with composite roughness and the
reference's finite-difference derivatives, against the live reference's run of the synthetic companion case
(tests/golden/irregular.*, oracle/ref_harness.build_irregular) and against the oracle."""
import copy

import numpy as np
import pytest

import util
from flow_sim_b200 import abi
from flow_sim_b200.runner import gvf_initial_conditions, run_flat

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lanes", [0, -1])       # fused kernel (N <= 249) / tiled long-reach kernels
def test_irregular_reach_matches_the_reference_run(lanes):
    flat = util.golden_inputs("irregular")
    ref = util.golden_outputs("irregular")
    assert (flat.geom["kind"] == abi.PR_XS_IRREGULAR).all() and flat.geom["irr_offset"][-1] == 192
    out = run_flat(flat, lanes=lanes)
    assert out["status"][0] == abi.PR_STATUS_OK
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], "irregular")
    assert np.array_equal(out["iters"][0], ref["iters"])


def test_polyline_sections_on_a_curved_centre_line():
    """Curvature slope Sc at polyline nodes (geometric top width, composite n, finite-difference dR/dA): the reference's
    run of the curved companion reach; the curvature moves the stages by 6e-6 m, four orders above the parity bar."""
    flat = util.golden_inputs("irregular_curved")
    ref = util.golden_outputs("irregular_curved")
    assert np.count_nonzero(flat.geom["curvature"]) == 10
    for lanes in (0, -1):
        out = run_flat(flat, lanes=lanes)
        util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], f"irregular, curved, lanes={lanes}")
        assert np.array_equal(out["iters"][0], ref["iters"])


def test_irregular_ensemble_roughness_and_inflow_members():
    import oracle_py

    flat = util.golden_inputs("irregular")
    M = 7
    flat.member_n_main = np.linspace(0.025, 0.04, M)
    flat.member_n_fp = np.linspace(0.045, 0.07, M)
    base = np.array(flat.up.series)
    flat.up.series = np.stack([base[0] + (base - base[0]) * (0.5 + 0.25 * m) for m in range(M)])
    ora = oracle_py.run(flat, n_members=M)
    for lanes in (0, -1):
        out = run_flat(flat, n_members=M, lanes=lanes)
        assert np.array_equal(out["status"], ora["status"]) and not out["status"].any()
        util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], f"irregular ensemble, lanes={lanes}")
        assert np.array_equal(out["iters"], ora["iters"])


def test_mixed_reach_and_other_boundaries():
    """Trapezoid nodes next to polyline nodes (what interpolating a surveyed section with a design section gives at the
    input stations), with a normal-depth downstream boundary evaluated on a polyline node."""
    import oracle_py

    flat = util.golden_inputs("irregular")
    N = flat.n_nodes
    g = {k: np.array(v) for k, v in flat.geom.items()}
    # node 0 becomes a compound trapezoid with the same invert
    off = g["irr_offset"]
    n0 = off[1] - off[0]
    g["irr_x"], g["irr_z"] = g["irr_x"][n0:], g["irr_z"][n0:]
    g["irr_offset"] = np.concatenate([[0], off[1:] - n0]).astype(np.int32)
    g["kind"][0] = abi.PR_XS_COMPOUND
    g["b_main"][0], g["m_main"][0], g["h_bank"][0] = 12.0, 1.5, 1.6
    g["T_bank"][0] = 12.0 + 2 * 1.5 * 1.6
    g["b_fp_l"][0], g["b_fp_r"][0], g["m_fp"][0] = 8.0, 6.0, 3.0
    g["W_bank"][0] = g["T_bank"][0] + 14.0
    flat.geom = g
    flat.down = copy.copy(flat.down)
    flat.down.type = abi.PR_BC_NORMAL_DEPTH
    flat.down.bed_slope = 0.0005
    ora = oracle_py.run(flat)
    out = run_flat(flat)
    assert np.array_equal(out["status"], ora["status"]) and out["status"][0] == 0
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], "mixed reach, normal depth")
    assert np.array_equal(out["iters"], ora["iters"])


def test_split_flow_is_refused_per_member_not_emulated():
    """A mid-channel bar that splits low flows into two equal sub-channels.  The reference's own Newton iteration
    diverges on this case (negative depths and a ZeroDivisionError in level 1); with the split-flow conveyance built,
    oracle and device follow it there and stop the member in the same level with PR_STATUS_NAN."""
    import oracle_py

    flat = util.golden_inputs("irregular_levee")        # oracle/ref_harness.build_irregular(levee=True)
    ora = oracle_py.run(flat)
    out = run_flat(flat)
    assert ora["status"][0] == abi.PR_STATUS_NAN and out["status"][0] == abi.PR_STATUS_NAN
    assert out["fail_level"][0] == ora["fail_level"][0] == 1


@pytest.mark.parametrize("lanes", [0, -1])
def test_split_flow_reach_matches_the_reference_run(lanes):
    """Split flow the reference survives (tests/golden/irregular_pocket.*: a side pocket behind a ridge; 75 of its 117
    node-levels run on the multi-sub-channel conveyance, the rest on the single-channel formulas, with the switch in
    the middle of the run): fused kernel and tile kernels against the reference's own run."""
    flat = util.golden_inputs("irregular_pocket")
    ref = util.golden_outputs("irregular_pocket")
    out = run_flat(flat, lanes=lanes)
    assert out["status"][0] == abi.PR_STATUS_OK
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], f"irregular_pocket, lanes={lanes}")
    assert np.array_equal(out["iters"][0], ref["iters"])


def test_split_flow_roughness_ensemble():
    """64 roughness members of the side-pocket reach (GVF profile per member on the device, split and single-channel
    nodes side by side) against the oracle."""
    import oracle_py
    from flow_sim_b200.runner import gvf_initial_conditions

    flat = util.golden_inputs("irregular_pocket")
    M = 64
    flat.member_n_main = np.linspace(0.024, 0.036, M)
    h, q, st = gvf_initial_conditions(flat, M, 60.0, 2.0)
    oh, oq, ost = oracle_py.gvf(flat, 60.0, 2.0, n_members=M)
    assert np.array_equal(st, ost) and util.max_rel(h, oh) <= 1e-11
    flat.ic_depth, flat.ic_flow = oh, oq
    ora = oracle_py.run(flat, n_members=M, out_mode=abi.PR_OUT_UPSTREAM, trace_prev_error=True)
    out = run_flat(flat, n_members=M, out_mode=abi.PR_OUT_UPSTREAM)
    assert np.array_equal(out["status"], ora["status"])
    ok = np.nonzero(ora["status"] == 0)[0]
    assert len(ok) > M // 2
    util.assert_iteration_parity(out, ora, flat.tol, "irregular_pocket ensemble", members=ok)


def test_derived_arrays_and_normal_depth_on_polyline_sections():
    """The set-up and post-processing kernels take IrregularSection nodes: derived result arrays (area, top width,
    Froude number, velocity, celerity - Solver.prepare_results) and the steady normal-depth initial state
    (Channel._steady_conditions) against the mirror's host-side section objects."""
    from flow_sim_b200.cases import build_irregular
    from flow_sim_b200.flatten import flatten_solver
    from flow_sim_b200.runner import derived_results, normal_depth_initial_conditions

    solver, kw = build_irregular(pocket=True)
    flat = flatten_solver(solver, **kw)
    xs = solver.channel.xs_at_node
    rng = np.random.default_rng(4)
    M, L, N = 3, flat.n_levels, flat.n_nodes
    depth = rng.uniform(0.6, 4.5, (M, L, N))
    flow = rng.uniform(20.0, 120.0, (M, L, N))
    got = derived_results(flat, depth, flow)
    area = np.array([[[xs[i].area(depth[m, k, i] + xs[i].z_min) for i in range(N)] for k in range(L)] for m in range(M)])
    top = np.array([[[xs[i].top_width(depth[m, k, i] + xs[i].z_min) for i in range(N)] for k in range(L)] for m in range(M)])
    assert util.max_rel(got["area"], area) <= 1e-13 and util.max_rel(got["top_width"], top) <= 1e-13
    assert util.max_rel(got["velocity"], flow / area) <= 1e-13
    assert util.max_rel(got["wave_celerity"], flow / area + np.sqrt(flat.g * area / top)) <= 1e-13
    # normal depth: the root of Q - K(hw) sqrt(S0) per node, brentq on the host / Brent's method on the device
    h, q = normal_depth_initial_conditions(flat, 1, 60.0)
    host = np.array([s.normal_depth(Q_target=60.0) for s in xs])
    assert np.max(np.abs(h[0] - host)) <= 1e-10 and np.all(q == 60.0)


def test_backwater_profile_and_roughness_sweep_on_polyline_sections():
    """The calibration workflow on surveyed sections: device GVF profile per member (geometric top width, composite
    roughness), then the ensemble solve from those profiles - against the oracle."""
    import oracle_py
    from flow_sim_b200.ensemble import EnsembleRunner, to_host

    flat = util.golden_inputs("irregular")
    M = 6
    n_main = np.linspace(0.025, 0.04, M)
    n_fp = np.linspace(0.05, 0.08, M)
    flat.member_n_main, flat.member_n_fp = n_main, n_fp
    q0, hd = np.linspace(50.0, 70.0, M), np.linspace(1.8, 2.4, M)
    h, q, st = gvf_initial_conditions(flat, M, q0, hd)
    ho, qo, sto = oracle_py.gvf(flat, q0, hd, n_members=M)
    assert np.array_equal(st, sto) and not st.any()
    assert util.max_rel(h, ho) <= 1e-12 and np.array_equal(q, qo)
    flat.member_n_main = flat.member_n_fp = None
    res = to_host(EnsembleRunner(flat, "cuda:0").roughness_sweep(n_main, n_fp=n_fp, downstream_depth=2.0, q0=60.0,
                                                                  out_mode=abi.PR_OUT_FULL))
    flat.member_n_main, flat.member_n_fp = n_main, n_fp
    ho, qo, _ = oracle_py.gvf(flat, 60.0, 2.0, n_members=M)
    flat.ic_depth, flat.ic_flow = ho, qo
    ora = oracle_py.run(flat, n_members=M)
    assert not res["status"].any() and np.array_equal(res["iters"], ora["iters"])
    util.assert_parity(res["depth"], res["flow"], ora["depth"], ora["flow"], "roughness sweep on polyline sections")


def test_mirror_solver_run_with_polyline_sections():
    """The drop-in surface: IrregularSection objects on the mirror API, solver.run() on the device, the reference's
    result arrays (derived arrays come from the host-side section objects, as in the reference)."""
    from flow_sim_b200.cases import build_irregular

    solver, kw = build_irregular()
    solver.run(verbose=0, **kw)
    ref = util.golden_outputs("irregular")
    util.assert_parity(solver.depth, solver.flow, ref["depth"], ref["flow"], "mirror run, irregular")
    assert np.array_equal(solver.iterations, ref["iters"])
    assert solver.area.shape == solver.depth.shape and np.all(solver.top_width > 0) and np.all(solver.froude_number < 1)
    with pytest.raises(ValueError, match="Convergence|NaN"):      # the bar case: the reference itself diverges in level 1
        build_irregular(bar=True)[0].run(verbose=0, **kw)
    pocket, kw = build_irregular(pocket=True)                     # split flow through the mirror API
    pocket.run(verbose=0, **kw)
    ref = util.golden_outputs("irregular_pocket")
    util.assert_parity(pocket.depth, pocket.flow, ref["depth"], ref["flow"], "mirror run, irregular_pocket")
    assert np.array_equal(pocket.iterations, ref["iters"])


@pytest.mark.parametrize("lanes", [0, 8, 16, 32])
def test_reach_from_a_trapezoid_to_a_polyline(lanes):
    """Node 0 a compound trapezoid, the other nodes the reference's blend of it with a polyline (cross_section.py:795-849,
    933-969): the flat inputs against the reference run (tests/golden/mixed_sections.*), in every packing; then through
    the mirror API, derived arrays from the device for both node kinds."""
    flat = util.golden_inputs("mixed_sections")
    ref = util.golden_outputs("mixed_sections")
    assert flat.geom["kind"][0] != flat.geom["kind"][1]
    out = run_flat(flat, lanes=lanes)
    assert np.array_equal(out["iters"][0], ref["iters"])
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], f"mixed_sections, lanes={lanes}")
    if lanes == 0:
        from flow_sim_b200.cases import build_mixed

        solver, kw = build_mixed()
        solver.run(verbose=0, **kw)
        util.assert_parity(solver.depth, solver.flow, ref["depth"], ref["flow"], "mirror run, mixed_sections")
        assert np.array_equal(solver.iterations, ref["iters"])
        sections = solver.channel.xs_at_node
        area = np.array([[sections[i].area(solver.depth[k, i] + sections[i].z_min) for i in range(len(sections))]
                         for k in range(solver.depth.shape[0])])
        assert np.allclose(solver.area, area, rtol=1e-12, atol=0)


def test_general_storage_with_head_losses_behind_a_polyline_node():
    """Area curve + outflow curve + head losses (friction over the reservoir length uses the polyline node's
    conveyance, composite n and dA/dh) - the downstream boundary of the storage_general case on the polyline reach."""
    import oracle_py

    flat = util.golden_inputs("irregular")
    flat.down = copy.copy(util.golden_inputs("storage_general").down)
    flat.down.fixed_depth = 2.0
    flat.down.storage_ymax = 200.0
    ora = oracle_py.run(flat)
    out = run_flat(flat)
    assert np.array_equal(out["status"], ora["status"]) and out["status"][0] == 0
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], "storage with losses, polyline reach")
    assert np.array_equal(out["iters"], ora["iters"])
    assert util.max_rel(out["storage_stage"], ora["storage_stage"]) <= util.RTOL


def test_long_polyline_reach_on_the_tile_kernels():
    """N = 301 > 249: the polyline reach stretched to 300 km goes to the tiled path on its own."""
    import oracle_py

    flat = util.golden_inputs("irregular")
    N = 301
    g = flat.geom
    src = np.minimum(np.arange(N) * flat.n_nodes // N, flat.n_nodes - 1)       # repeat the 13 sections along the reach
    off = g["irr_offset"]
    geom = {k: np.array(v)[src] for k, v in g.items() if not k.startswith("irr_") }
    geom["z_bed"] = g["z_bed"][0] * (1.0 - np.arange(N) / (N - 1))              # uniform slope
    xs, zs, offs = [], [], [0]
    for i, s_ in enumerate(src):
        seg = slice(off[s_], off[s_ + 1])
        xs.append(g["irr_x"][seg]); zs.append(g["irr_z"][seg] - g["z_bed"][s_] + geom["z_bed"][i])
        offs.append(offs[-1] + len(xs[-1]))
    geom["irr_x"], geom["irr_z"] = np.concatenate(xs), np.concatenate(zs)
    geom["irr_offset"] = np.array(offs, dtype=np.int32)
    geom["irr_left"], geom["irr_right"] = g["irr_left"][src], g["irr_right"][src]
    flat.geom, flat.n_nodes = geom, N
    flat.ic_depth, flat.ic_flow = np.full(N, 2.0), np.full(N, 60.0)
    flat.up.bed_level = float(geom["z_bed"][0])
    flat.n_levels = 4
    flat.up.series = flat.up.series[:4]
    ora = oracle_py.run(flat)
    out = run_flat(flat)
    assert np.array_equal(out["status"], ora["status"]) and out["status"][0] == 0
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], "N=301 polyline reach")
    assert np.array_equal(out["iters"], ora["iters"])


@pytest.mark.parametrize("k,lanes", [(6, 0), (6, -1), (20, 0), (20, -1)])
def test_dense_polylines_with_split_flow(k, lanes):
    """The side-pocket reach with every polyline segment cut into k pieces: k = 6 keeps the sections within the stage
    tables (<= 128 points), k = 20 gives 221-381 points per section - no table, sums of more than 128 terms (numpy's
    recursive pairwise order) and wetted sub-channels of ~100 points (views of the parent's points, no size limit).
    Fused and tiled path against the oracle, which the live reference pins on such a section
    (tests/golden/irregular_dense_probe.npz)."""
    import oracle_py

    flat = util.golden_inputs("irregular_pocket")
    util.densify_polylines(flat, k)
    assert int(np.diff(flat.geom["irr_offset"]).max()) > (300 if k == 20 else 100)
    M = 5
    flat.member_n_main = np.linspace(0.026, 0.034, M)
    ora = oracle_py.run(flat, M, trace_prev_error=True)
    out = run_flat(flat, n_members=M, lanes=lanes)
    assert np.array_equal(out["status"], ora["status"]) and not ora["status"].any()
    util.assert_iteration_parity(out, ora, flat.tol, f"dense polylines x{k}, lanes={lanes}")
