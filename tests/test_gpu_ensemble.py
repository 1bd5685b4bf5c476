"""GPU tests of the ensemble path through the C ABI: per-member roughness / inflow / initial conditions,
device-side GVF profile and objective, independence of members, boundary-condition and section coverage,
edge cases (smallest grids, single level, non-convergence) - each against the CPU oracle on identical inputs."""
import copy

import numpy as np
import pytest

import util
from flow_sim_b200 import abi
from flow_sim_b200.runner import gvf_initial_conditions, rating_objective, run_flat

pytestmark = pytest.mark.gpu

Q_QUERY = [1562.5, 3850, 6000, 10000, 14000, 21000]
H_TARGET = [497.5, 500, 502, 505, 507, 510]


def _check(flat, M, what, out_mode=abi.PR_OUT_FULL, rtol=util.RTOL):
    import oracle_py

    ora = oracle_py.run(flat, n_members=M, out_mode=out_mode)
    out = run_flat(flat, n_members=M, out_mode=out_mode)
    assert np.array_equal(out["status"], ora["status"]), what
    assert np.array_equal(np.isnan(out["depth"]), np.isnan(ora["depth"])), f"{what}: NaN fill differs"
    fin = ~np.isnan(ora["depth"])
    util.assert_parity(out["depth"][fin], out["flow"][fin], ora["depth"][fin], ora["flow"][fin], what, rtol)
    assert np.array_equal(out["iters"], ora["iters"]), f"{what}: Newton iteration counts differ"
    assert np.array_equal(out["fail_level"], ora["fail_level"])
    return out, ora


def test_roughness_ensemble_with_device_gvf_and_objective():
    """256 members of the config-4 grid: GVF kernel, Newton kernel and objective kernel against the oracle."""
    import oracle_py

    flat = util.golden_inputs("gerd_calib_m0")
    M = 256
    flat.member_n_main = 0.020 + 0.040 * np.arange(M) / (M - 1)
    h, q, st = gvf_initial_conditions(flat, M, flat.meta["initial_flow"], flat.meta["downstream_depth"])
    ho, qo, sto = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=M)
    assert np.array_equal(st, sto) and not st.any()
    assert util.max_rel(h, ho) <= 1e-12 and np.array_equal(q, qo)
    flat.ic_depth, flat.ic_flow = ho, qo            # identical initial state for both solvers
    out, ora = _check(flat, M, "roughness ensemble", out_mode=abi.PR_OUT_UPSTREAM)
    lv, rm = rating_objective(flat.n_levels, out["flow"], out["depth"], flat.meta["z0"], Q_QUERY, H_TARGET)
    lvo, rmo = oracle_py.objective(flat.n_levels, ora["flow"], ora["depth"], flat.meta["z0"], Q_QUERY, H_TARGET)
    assert util.max_rel(lv, lvo) <= util.RTOL and util.max_rel(rm, rmo) <= util.RTOL
    # the reference's own numbers for the two ends of the grid
    assert abs(rm[0] - 5.847778566618879) <= 1e-9 * 5.85 and abs(rm[-1] - 1.7408210264943833) <= 1e-9 * 1.75


def test_members_are_independent_of_batch_composition():
    """A member's result must not depend on which other members share its launch, CTA or batch position:
    this is what makes 1/2/4/8-GPU sharding bit-identical."""
    flat = util.golden_inputs("gerd_calib_m0")
    n = 0.020 + 0.040 * np.arange(64) / 63
    f = copy.copy(flat); f.member_n_main = n
    h, q, _ = gvf_initial_conditions(f, 64, flat.meta["initial_flow"], flat.meta["downstream_depth"])
    f.ic_depth, f.ic_flow = h, q
    full = run_flat(f, n_members=64, out_mode=abi.PR_OUT_UPSTREAM)
    perm = np.random.default_rng(0).permutation(64)[:13]
    g = copy.copy(flat); g.member_n_main = n[perm]; g.ic_depth, g.ic_flow = h[perm], q[perm]
    part = run_flat(g, n_members=13, out_mode=abi.PR_OUT_UPSTREAM)
    for key in ("depth", "flow", "iters", "status"):
        assert np.array_equal(part[key], full[key][perm]), key


def test_host_and_device_memory_paths_agree():
    import torch

    flat = util.golden_inputs("gerd_calib_m21845")
    host = run_flat(flat, mem=abi.PR_MEM_HOST)
    dev = run_flat(flat, mem=abi.PR_MEM_DEVICE, device="cuda:0")
    for key in ("depth", "flow", "iters", "status"):
        assert np.array_equal(host[key], dev[key].cpu().numpy()), key
    assert isinstance(dev["depth"], torch.Tensor) and dev["depth"].is_cuda


def test_inflow_scenarios_and_floodplain_roughness_per_member():
    """Per-member upstream hydrographs [M, levels] and an n_fp override."""
    flat = util.golden_inputs("gerd_calib_m30000")
    M = 6
    scale = np.linspace(0.8, 1.1, M)[:, None]
    base = flat.up.series[None, :]
    flat.up.series = base[:, :1] + (base - base[:, :1]) * scale        # same initial flow, scaled wave
    flat.member_n_fp = np.linspace(0.04, 0.07, M)
    _check(flat, M, "inflow scenarios + n_fp")


def _prismatic(kind="rect", down="normal_depth", up="flow_hydrograph", n_nodes=12, levels=8, rating=None):
    """Small synthetic reach built on the mirror API."""
    from math import pi, sin

    from flow_sim_b200.flatten import flatten_solver
    from flow_sim_b200.hydromodel import Boundary, Channel, Hydrograph, PreissmannSolver, TrapezoidalSection

    L, S0, dt = 1000.0 * (n_nodes - 1), 0.0005, 1800
    wave = lambda t: 60 + 40 * sin(pi * min(t, 6 * dt) / (6 * dt)) ** 2
    if up == "flow_hydrograph":
        us = Boundary("flow_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=2.0, hydrograph=Hydrograph(wave))
    else:
        us = Boundary("stage_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=2.0,
                      hydrograph=Hydrograph(lambda t: S0 * L + 2.0 + 0.3 * sin(pi * min(t, 6 * dt) / (6 * dt)) ** 2))
    kw = dict(rating_curve=rating) if down == "rating_curve" else {}
    ds = Boundary(down, chainage=L, bed_level=0.0, initial_depth=2.0, **kw)
    ch = Channel(us, ds, initial_flow=60.0, roughness=0.03, width=30.0, interpolation_method="linear")
    if kind != "rect":
        mk = lambda z: TrapezoidalSection(z_bed=z, b_main=20.0, m_main=1.5, n_main=0.03, bed_slope=S0,
                                          **(dict(z_bank=z + 1.6, b_fp_left=15.0, b_fp_right=5.0, m_fp=3.0, n_left=0.05,
                                                  n_right=0.06) if kind == "compound" else {}))
        ch.set_cross_sections([0.0, L], [mk(S0 * L), mk(0.0)])
    s = PreissmannSolver(channel=ch, theta=0.6, time_step=dt, spatial_step=1000.0, simulation_time=(levels - 1) * dt)
    return flatten_solver(s, tolerance=1e-6, max_iter=60)


@pytest.mark.parametrize("kind", ["rect", "trapezoid", "compound"])
@pytest.mark.parametrize("down", ["normal_depth", "fixed_depth"])
def test_section_kinds_and_simple_boundaries(kind, down):
    _check(_prismatic(kind=kind, down=down), 1, f"{kind}/{down}")


def test_stage_hydrograph_upstream():
    _check(_prismatic(kind="trapezoid", up="stage_hydrograph"), 1, "stage hydrograph")


@pytest.mark.parametrize("form", ["polynomial", "power", "fitted"])
def test_rating_curve_forms(form):
    from flow_sim_b200.hydromodel import RatingCurve

    rc = RatingCurve()
    if form == "polynomial":
        rc.set("polynomial", a=6.0, b=9.0, c=18.0, stage_shift=0.0)
    elif form == "power":
        rc.set("power", a=21.0, b=1.55, stage_shift=0.0)
    else:
        st = np.linspace(0.5, 4.0, 9)
        rc.fit(discharges=21.0 * st ** 1.55, stages=st, stage_shift=0.0, type="polynomial", scale=True, degree=3)
    _check(_prismatic(kind="rect", down="rating_curve", rating=rc), 1, f"rating {form}")


@pytest.mark.parametrize("n_nodes", [2, 3, 33, 34, 63, 125, 126, 249])
def test_grid_sizes_across_nodes_per_lane_families(n_nodes):
    """N = 2 (one cell) up to 249 (the largest reach the fused kernel takes: 8 nodes per lane), including sizes
    where the last node is interior to a lane."""
    _check(_prismatic(kind="compound", n_nodes=n_nodes, levels=4), 1, f"N={n_nodes}")


def test_single_level_run_returns_initial_state():
    flat = _prismatic(levels=1)
    out = run_flat(flat)
    assert np.array_equal(out["depth"][0, 0], flat.ic_depth) and out["status"][0] == 0


def test_non_convergence_is_reported_per_member_not_raised():
    """max_iter exhausted -> status 1 with the failing level, NaN-filled remainder; the healthy member next to it
    is unaffected (preissmann.py:124-126 raises for a single run)."""
    flat = util.golden_inputs("gerd_calib_m0")
    flat.max_iter = 11                      # level 3 needs 11, level 6 needs 12 iterations for n = 0.020
    flat.member_n_main = np.array([0.020, 0.020])
    out, ora = _check(flat, 2, "max_iter")
    assert out["status"].tolist() == [abi.PR_STATUS_MAX_ITER] * 2 and out["fail_level"][0] == 6
    assert np.isnan(out["depth"][0, 6:]).all() and not np.isnan(out["depth"][0, :6]).any()
    assert out["iters"][0, 5] == 11 and not out["iters"][0, 6:].any()


def test_mirror_solver_run_matches_reference_golden():
    """The drop-in surface: PreissmannSolver.run() on the mirror objects, result attributes as in the reference."""
    from flow_sim_b200.cases import build_example

    ref = util.golden_outputs("example")
    solver, kw = build_example()
    solver.run(verbose=0, **kw)
    util.assert_parity(solver.depth, solver.flow, ref["depth"], ref["flow"], "mirror example")
    assert np.array_equal(solver.iterations, ref["iters"])
    assert util.max_rel(solver.storage_stage, ref["storage_stage"]) <= util.RTOL
    assert solver.level.shape == solver.depth.shape == (25, 21)
    for name in ("area", "top_width", "froude_number", "velocity", "wave_celerity", "amplitude", "peak_amplitude",
                 "storage_outflow"):
        assert np.all(np.isfinite(getattr(solver, name)))
    solver2, kw = build_example()
    with pytest.raises(ValueError, match="Convergence within 5 iterations couldn't be achieved."):
        solver2.run(verbose=0, tolerance=1e-4, max_iter=5)


def test_general_lumped_storage_variants_vs_oracle():
    """Area curve only / outflow curve only / everything, on the fused kernel; the full variant is
    also pinned to the reference by test_cuda_vs_reference_golden[storage_general]."""
    import copy

    full = util.golden_inputs("storage_general")
    a = copy.copy(full); a.down = copy.copy(full.down); a.down.storage_outflow = None; a.down.storage_losses = False
    b = copy.copy(full); b.down = copy.copy(full.down); b.down.storage_curve = None; b.down.storage_area = 1.5e6
    b.down.storage_losses = False
    c = copy.copy(full); c.down = copy.copy(full.down); c.down.storage_curve = None; c.down.storage_area = 1.25e6
    c.down.storage_outflow = None                      # constant area + losses only
    a.down.storage_ymax = c.down.storage_ymax = 200.0   # without an outflow the pool rises past the 40 m bracket
    for name, flat in (("area curve", a), ("outflow", b), ("losses", c), ("all", full)):
        out, ora = _check(flat, 1, f"storage: {name}")
        assert out["status"][0] == 0
        assert util.max_rel(out["storage_stage"], ora["storage_stage"]) <= util.RTOL


def test_general_storage_and_gate_control_with_both_roughness_overrides():
    """The rare-boundary kernels decide the roughness overrides at run time: per-member n_main AND n_fp."""
    flat = util.golden_inputs("gerd_gated")
    flat.member_n_main = np.array([0.028, 0.030, 0.034])
    flat.member_n_fp = np.array([0.04, 0.05, 0.06])
    _check(flat, 3, "gate control + n_main + n_fp")
    flat = util.golden_inputs("storage_general")
    flat.member_n_main = np.array([0.03, 0.035])
    _check(flat, 2, "general storage + n_main")


def test_storage_root_outside_solution_boundaries_fails_the_member():
    """scipy's brentq raises when the mass balance has no sign change on solution_boundaries; the member stops at that
    iteration with status NaN (the oracle does the same), for the closed-form and for the Brent variant."""
    import copy

    full = util.golden_inputs("storage_general")
    a = copy.copy(full); a.down = copy.copy(full.down); a.down.storage_outflow = None; a.down.storage_losses = False
    c = copy.copy(a); c.down = copy.copy(a.down); c.down.storage_curve = None; c.down.storage_area = 1.25e6
    for flat in (a, c):
        out, ora = _check(flat, 1, "storage bracket failure")
        assert out["status"][0] == abi.PR_STATUS_NAN and out["fail_level"][0] > 1


def test_mirror_general_storage_run_matches_reference():
    from test_mirror_api import _storage_general_solver

    ref = util.golden_outputs("storage_general")
    solver, kw = _storage_general_solver()
    solver.run(verbose=0, **kw)
    util.assert_parity(solver.depth, solver.flow, ref["depth"], ref["flow"], "mirror storage_general")
    assert np.array_equal(solver.iterations, ref["iters"])
    assert util.max_rel(solver.storage_stage, ref["storage_stage"]) <= util.RTOL
    assert np.all(np.isfinite(solver.storage_outflow))


def test_release_scenarios_per_member_rating_curves_vs_reference_goldens():
    """pr_bc.member_ratings: two members with different Roseires gate scenarios, pool levels and roughness in one
    launch; each equals the reference run of its own scenario."""
    flat, refs = util.release_ensemble()
    out, _ = _check(flat, 2, "release scenarios")
    for m, ref in enumerate(refs):
        util.assert_parity(out["depth"][m], out["flow"][m], ref["depth"], ref["flow"], f"release member {m}")
        assert np.array_equal(out["iters"][m], ref["iters"])


def test_release_scenario_runner_builds_profiles_and_curves_per_member():
    """EnsembleRunner.release_scenarios: 48 gate/pool-level scenarios from rating-curve objects, device GVF profile
    from each member's own pool level, against the oracle fed the same flattened curves."""
    import oracle_py
    from flow_sim_b200.cases import gerd_roseires as gr
    from flow_sim_b200.ensemble import EnsembleRunner, to_host
    from flow_sim_b200.flatten import flatten_rating

    flat = util.golden_inputs("gerd_release")
    q0 = float(flat.meta["initial_flow"])
    curves = [gr.RoseiresRatingCurve(initial_stage=485.5 + 0.05 * m, initial_flow=q0, jammed_spillways=m % 4,
                                     jammed_sluice_gates=(m // 4) % 3, buffer=0.25 + 0.05 * (m % 5))
              for m in range(48)]
    res = to_host(EnsembleRunner(flat, "cuda:0").release_scenarios(curves, out_mode=abi.PR_OUT_FULL))
    assert not res["ic_status"].any() and not res["status"].any()
    ratings = [flatten_rating(c) for c in curves]
    depth = np.array([r["stage0"] - flat.down.bed_level for r in ratings])
    ho, qo, _ = oracle_py.gvf(flat, q0, depth, n_members=48)
    flat.down.member_ratings = ratings
    flat.ic_depth, flat.ic_flow = ho, qo
    ora = oracle_py.run(flat, n_members=48)
    util.assert_parity(res["depth"], res["flow"], ora["depth"], ora["flow"], "release scenarios (runner)")
    assert np.array_equal(res["iters"], ora["iters"])
    assert len({tuple(r) for r in res["iters"]}) > 8          # the scenarios really differ


def test_gate_controlled_rating_curve_state_per_member():
    """RoseiresRatingCurve(smooth=False): every member carries its own gate state (open flag, cool-down, last stage).
    Members differ in initial gate position, cool-down, roughness and inflow scale; one of them stops converging
    when its gates slam shut - all against the oracle, which reproduces the reference's gated run (gerd_gated)."""
    flat = util.golden_inputs("gerd_gated")
    base = np.array(flat.up.series)
    variants = [(1.0, 7200.0, 1, 0.030), (1.0, 18000.0, 0, 0.030), (4.0, 7200.0, 1, 0.030), (3.0, 3600.0, 0, 0.040),
                (6.0, 3600.0, 1, 0.030), (2.0, 0.0, 1, 0.035)]
    M = len(variants)
    flat.up.series = np.stack([base[0] + (base - base[0]) * v[0] for v in variants])
    flat.down.member_ratings = [dict(flat.down.rating, max_cooldown=v[1], initially_open=v[2]) for v in variants]
    flat.member_n_main = np.array([v[3] for v in variants])
    out, ora = _check(flat, M, "gate-controlled ensemble")
    assert out["status"].tolist().count(abi.PR_STATUS_MAX_ITER) >= 1 and out["status"][0] == abi.PR_STATUS_OK
    ref = util.golden_outputs("gerd_gated")
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], "gated member 0 vs reference")
    # mixed launch: gated and smooth members side by side
    flat.down.member_ratings[1] = dict(flat.down.rating, gate_control=0)
    _check(flat, M, "gated + smooth members")


def test_gate_controlled_curve_on_the_long_reach_path():
    flat = util.golden_inputs("gerd_gated")
    ref = util.golden_outputs("gerd_gated")
    out = run_flat(flat, n_members=3, lanes=-1)
    for m in range(3):
        util.assert_parity(out["depth"][m], out["flow"][m], ref["depth"], ref["flow"], "gated, tiled path")
        assert np.array_equal(out["iters"][m], ref["iters"])


def test_unsupported_combinations_are_refused_with_a_message_not_emulated():
    """Every configuration the kernels do not implement comes back as PR_ERR_UNSUPPORTED / PR_ERR_ARG with a reason in
    pr_last_error - never a silent CPU path or a wrong answer."""
    from flow_sim_b200.abi import PreissmannLibraryError

    def refused(flat, M, pattern, **kw):
        with pytest.raises(PreissmannLibraryError, match=pattern):
            run_flat(flat, n_members=M, **kw)

    # gate-controlled curve: not upstream
    flat = util.golden_inputs("gerd_gated")
    flat.up, flat.down = copy.copy(flat.down), copy.copy(flat.up)
    refused(flat, 1, "upstream")
    # malformed inputs
    flat = util.golden_inputs("example")
    flat.down.storage_area = 0.0
    refused(flat, 1, "surface area")
    flat = util.golden_inputs("gerd_calib_m0")
    flat.down.rating = dict(flat.down.rating, buffer=0.0)
    refused(flat, 1, "buffer")
    flat = util.golden_inputs("gerd_calib_m0")
    flat.down.member_ratings = [flat.down.rating, dict(flat.down.rating, n_gates=99)]
    refused(flat, 2, "n_gates")


# ---- short reaches: 2 or 4 members per warp (8 / 16 lanes per member) --------------------------------------------

@pytest.mark.parametrize("n_nodes,lanes", [(2, 0), (8, 0), (9, 0), (15, 0), (16, 0), (21, 0), (29, 0), (30, 0), (31, 0),
                                           (32, 0), (45, 0), (61, 0), (62, 0), (21, 16), (21, 32), (8, 16), (61, 32)])
def test_packed_warp_families_vs_oracle(n_nodes, lanes):
    """Members that share a warp run in lockstep but converge on their own: 11 members with different roughness and
    inflow (so different iteration counts and convergence moments inside one warp), every size class of the
    8- and 16-lane kernels, and the same sizes forced onto wider groups."""
    import oracle_py

    M = 11                                                    # not a multiple of the members per warp
    flat = _prismatic(kind="compound", n_nodes=n_nodes, levels=5)
    flat.member_n_main = 0.022 + 0.004 * np.arange(M)
    base = np.array(flat.up.series)
    flat.up.series = np.stack([base[0] + (base - base[0]) * (0.3 + 0.25 * m) for m in range(M)])
    ora = oracle_py.run(flat, n_members=M)
    out = run_flat(flat, n_members=M, lanes=lanes)
    assert np.array_equal(out["status"], ora["status"]) and not out["status"].any()
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], f"N={n_nodes} lanes={lanes}")
    assert np.array_equal(out["iters"], ora["iters"])
    assert len({tuple(r) for r in out["iters"]}) > 1 or n_nodes <= 3     # the members really differ


def test_packed_warp_one_member_failing_leaves_its_warp_mates_alone():
    """max_iter cut short for everybody, but only the members with the big flood need that many iterations: the
    failing members of a warp stop (NaN-filled), the others finish, exactly as the oracle says."""
    import oracle_py

    M = 8
    flat = _prismatic(kind="compound", n_nodes=21, levels=6)
    base = np.array(flat.up.series)
    flat.up.series = np.stack([base[0] + (base - base[0]) * (0.05 if m % 2 else 3.0) for m in range(M)])
    flat.max_iter = 24
    ora = oracle_py.run(flat, n_members=M)
    assert 0 < (ora["status"] != 0).sum() < M                 # a mix of failing and surviving members in each warp
    out = run_flat(flat, n_members=M)
    assert np.array_equal(out["status"], ora["status"]) and np.array_equal(out["fail_level"], ora["fail_level"])
    assert np.array_equal(np.isnan(out["depth"]), np.isnan(ora["depth"]))
    fin = ~np.isnan(ora["depth"])
    util.assert_parity(out["depth"][fin], out["flow"][fin], ora["depth"][fin], ora["flow"][fin], "mixed failure")
    assert np.array_equal(out["iters"], ora["iters"])


@pytest.mark.parametrize("case", ["example", "akbari", "storage_general"])
def test_shipped_short_cases_as_packed_ensembles(case):
    """Configs 1 and 2 (N = 21 / 30: lumped storage, normal depth) as 9-member inflow ensembles on the packed
    kernels, member 0 being the reference's own run."""
    import oracle_py

    flat = util.golden_inputs(case)
    ref = util.golden_outputs(case)
    M = 9
    base = np.array(flat.up.series)
    flat.up.series = np.stack([base[0] + (base - base[0]) * (1.0 + 0.1 * m) for m in range(M)])
    ora = oracle_py.run(flat, n_members=M)
    out = run_flat(flat, n_members=M)
    assert np.array_equal(out["status"], ora["status"])
    fin = ~np.isnan(ora["depth"])
    util.assert_parity(out["depth"][fin], out["flow"][fin], ora["depth"][fin], ora["flow"][fin], case)
    assert np.array_equal(out["iters"], ora["iters"])
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], f"{case} member 0 vs reference")
    if "storage_stage" in ref.files:
        assert util.max_rel(out["storage_stage"][0], ref["storage_stage"]) <= util.RTOL


def test_solver_run_ensemble_entry_point():
    """PreissmannSolver.run_ensemble (the additive API of SURVEY.md 8b) on the mirror objects: roughness members of the
    gerd calibration case must equal single runs of the reference (goldens), release scenarios its gerd_release run."""
    from flow_sim_b200.cases import build_gerd
    from flow_sim_b200.cases import gerd_roseires as gr

    solver, kw = build_gerd(n_main=0.03, calibration=True)
    ms = [0, 36408, 65535]
    res = solver.run_ensemble({"n_main": [util.calib_n(m) for m in ms]}, tolerance=kw["tolerance"], full_output=True,
                              q_query=Q_QUERY, h_target=H_TARGET)
    for i, m in enumerate(ms):
        ref = util.golden_outputs(f"gerd_calib_m{m}")
        util.assert_parity(res["depth"][i], res["flow"][i], ref["depth"], ref["flow"], f"run_ensemble member {m}")
        assert np.array_equal(res["iterations"][i], ref["iters"])
        assert abs(res["rmse"][i] - float(ref["calib_rmse"])) <= util.RTOL * float(ref["calib_rmse"])
    q0 = float(solver.channel.initial_flow_rate)
    curves = [gr.RoseiresRatingCurve(initial_stage=486.2, initial_flow=q0, jammed_spillways=2, jammed_sluice_gates=1, buffer=0.3),
              gr.RoseiresRatingCurve(initial_stage=487.0, initial_flow=q0)]
    res = solver.run_ensemble({"rating_curves": curves, "n_main": [0.03, 0.03]}, tolerance=kw["tolerance"], full_output=True)
    ref = util.golden_outputs("gerd_release")
    util.assert_parity(res["depth"][0], res["flow"][0], ref["depth"], ref["flow"], "run_ensemble release scenario")
    assert np.array_equal(res["iterations"][0], ref["iters"]) and not res["status"].any()
    with pytest.raises(ValueError, match="unknown member overrides"):
        solver.run_ensemble({"n_bank": [0.1]})


def test_member_order_changes_the_schedule_not_the_results():
    """pr_config.member_order (ABI 7): the persistent kernel hands the members to its warps in the given order; every
    member's result is bit-identical whatever the order, outputs stay in member order.  Also a launch far smaller than
    the grid (3 members) and one that leaves a partial last ticket."""
    from flow_sim_b200.ensemble import EnsembleRunner, to_host

    flat = util.golden_inputs("gerd_calib_m0")
    M = 1000
    n_main = 0.020 + 0.040 * np.arange(M) / (M - 1)
    runner = EnsembleRunner(flat, "cuda:0")
    base = to_host(runner.roughness_sweep(n_main))                       # descending-n order inside
    rng = np.random.default_rng(9)
    import torch

    f_ic = base                                                          # reuse: explicit solve with other orders
    from flow_sim_b200.runner import gvf_initial_conditions
    import copy

    f = copy.copy(runner.flat)
    f.member_n_main = torch.from_numpy(n_main).to("cuda:0")
    ich, icq, _ = gvf_initial_conditions(f, M, flat.meta["initial_flow"], flat.meta["downstream_depth"], abi.PR_MEM_DEVICE, "cuda:0", None)
    for order in (None, np.arange(M)[::-1].copy(), rng.permutation(M)):
        res = to_host(runner.solve(M, member_n_main=n_main, ic_depth=ich, ic_flow=icq,
                                   member_order=None if order is None else order.astype(np.int32)))
        for k in ("depth", "flow", "iters", "status"):
            assert np.array_equal(res[k], base[k]), k
    small = to_host(runner.roughness_sweep(n_main[:3]))
    for k in ("depth", "flow", "iters"):
        assert np.array_equal(small[k], base[k][:3])


def test_pinned_result_buffers_equal_plain_copies():
    from flow_sim_b200.ensemble import EnsembleRunner, PinnedResults, to_host

    flat = util.golden_inputs("gerd_calib_m0")
    runner = EnsembleRunner(flat, "cuda:0")
    res = runner.roughness_sweep(np.linspace(0.02, 0.06, 40))
    plain = to_host(res)
    pinned = PinnedResults()
    for _ in range(2):                      # second call reuses the buffers
        got = pinned.fetch(res, keys=("depth", "flow", "iters", "status"))
        assert set(got) == {"depth", "flow", "iters", "status"}
        for k in got:
            assert np.array_equal(got[k], plain[k]), k
