"""Pins the CPU oracle (oracle/preissmann_oracle.c) to the reference: committed outputs of the live reference
(tests/golden/*.ref.npz, made by oracle/make_golden.py), per-iteration (J.data, R, delta) captures, the two
rating-curve CSVs the reference ships, and scipy's brentq."""
import json
import os

import numpy as np
import pytest

import oracle_py
import util
from flow_sim_b200 import abi


@pytest.mark.parametrize("case", util.SMALL_CASES + ["gerd_full", "akbari_long", "gerd_gated_full"])
def test_oracle_reproduces_reference_run(case):
    flat = util.golden_inputs(case)
    ref = util.golden_outputs(case)
    out = oracle_py.run(flat)
    assert out["status"][0] == abi.PR_STATUS_OK
    util.assert_parity(out["depth"][0], out["flow"][0], ref["depth"], ref["flow"], case)
    assert np.array_equal(out["iters"][0], ref["iters"]), "Newton iteration counts differ from the reference"
    if "storage_stage" in ref.files:
        assert util.max_rel(out["storage_stage"][0], ref["storage_stage"]) <= util.RTOL


@pytest.mark.parametrize("case", ["example", "akbari", "gerd_calib_m0", "gerd_calib_m36408", "gerd_full"])
def test_oracle_assembly_matches_reference_captures(case):
    """R agrees to rounding of its largest term (bit-identical except the Roseires row, where sklearn's dot product
    sums in a different order), J.data to ~1e-11 (the finite-difference dQ/dz amplifies that rounding),
    delta agrees with SuperLU's to 1e-9 relative."""
    flat = util.golden_inputs(case)
    ref = util.golden_outputs(case)
    depth, flow = ref["depth"], ref["flow"]
    stage = ref["storage_stage"] if "storage_stage" in ref.files else None
    n_caps = len([k for k in ref.files if k.endswith("_R")])
    assert n_caps >= 1
    for j in range(n_caps):
        level = int(ref[f"cap{j}_level"])
        x = ref[f"cap{j}_x"]
        rc, R, J, delta = oracle_py.newton_step(flat, level, depth[level - 1], flow[level - 1], x[0::2], x[1::2],
                                                stage_record=stage)
        assert rc == 0
        Rr, Jr, dr = ref[f"cap{j}_R"], ref[f"cap{j}_J"], ref[f"cap{j}_delta"]
        assert np.max(np.abs(R - Rr)) <= 1e-14 * np.max(np.abs(x)) + 1e-12 * np.max(np.abs(Rr))
        assert np.max(np.abs(J - Jr) / np.maximum(np.abs(Jr), 1e-300)) <= 1e-9
        assert np.max(np.abs(delta - dr)) <= 1e-9 * max(np.max(np.abs(dr)), 1e-12)


def test_oracle_gvf_and_objective_match_reference():
    """The GVF profile (channel.py:307-378) and the calibration objective (model.py:105-113) for the members
    whose reference inputs/outputs are committed."""
    base = util.golden_inputs("gerd_calib_m0")
    for m in util.CALIB_MEMBERS:
        ref_in = util.golden_inputs(f"gerd_calib_m{m}")
        ref = util.golden_outputs(f"gerd_calib_m{m}")
        base.member_n_main = np.array([util.calib_n(m)])
        h, q, st = oracle_py.gvf(base, base.meta["initial_flow"], base.meta["downstream_depth"], n_members=1)
        assert st[0] == 0
        assert np.array_equal(h[0], ref_in.ic_depth), f"member {m}: GVF profile differs"
        lv, rm = oracle_py.objective(base.n_levels, ref["flow"][None, :, 0], ref["depth"][None, :, 0], base.meta["z0"],
                                     [1562.5, 3850, 6000, 10000, 14000, 21000], [497.5, 500, 502, 505, 507, 510])
        assert np.allclose(lv[0], ref["calib_levels"], rtol=1e-14, atol=0)
        assert abs(rm[0] - float(ref["calib_rmse"])) <= 1e-13


def test_member_roughness_override_equals_reference_sections():
    """n*w1 + n*w2 from the override reproduces the per-node roughness the reference interpolates (quirk 11)."""
    a = util.golden_inputs("gerd_calib_m30000")
    base = util.golden_inputs("gerd_calib_m0")
    base.member_n_main = np.array([util.calib_n(30000)])
    for node in (0, 1, 17, 60, 120):
        pa = oracle_py.section_probe(a, node, 9.0, 3000.0)
        pb = oracle_py.section_probe(base, node, 9.0, 3000.0, n_members=1)
        assert pa == pb


def test_release_scenario_ensemble_reproduces_both_reference_runs():
    """Per-member rating curves (pr_bc.member_ratings): each member must equal the reference run of its own scenario."""
    flat, refs = util.release_ensemble()
    out = oracle_py.run(flat, n_members=2)
    for m, ref in enumerate(refs):
        assert out["status"][m] == abi.PR_STATUS_OK
        util.assert_parity(out["depth"][m], out["flow"][m], ref["depth"], ref["flow"], f"release member {m}")
        assert np.array_equal(out["iters"][m], ref["iters"])
    # and the initial profiles come from the per-member downstream depth
    depth = [flat.down.member_ratings[m]["stage0"] - flat.down.bed_level for m in range(2)]
    h, q, st = oracle_py.gvf(flat, flat.meta["initial_flow"], depth, n_members=2)
    assert not st.any() and util.max_rel(h, flat.ic_depth) <= 1e-12


def test_roseires_release_curves_known_answers():
    """The reference's low/high_release_rating_curve.csv (10 printed digits)."""
    from flow_sim_b200.cases.gerd_roseires import RoseiresRatingCurve
    from flow_sim_b200.flatten import flatten_rating

    kat = json.load(open(os.path.join(util.GOLD, "roseires_release_kat.json")))
    rc = RoseiresRatingCurve(initial_stage=487, initial_flow=2094.106301)
    d = flatten_rating(rc)
    for name, state_key, sl_key in (("low", "closed_state", "sluices_closed"), ("high", "open_state", "sluices_open")):
        one = dict(d)
        one["open_state"] = one["closed_state"] = d[state_key]
        one["sluices_open"] = one["sluices_closed"] = d[sl_key]
        for stage, q_ref in kat[name]:
            q, _ = oracle_py.rating(one, stage)
            assert abs(q - q_ref) <= 5e-8 * abs(q_ref), (name, stage, q, q_ref)   # CSV prints 10 significant digits
            state = rc.closed_state if name == "low" else rc.open_state
            assert abs(rc.release(stage, state) - q_ref) <= 5e-8 * abs(q_ref)


def test_rating_curve_forms_against_numpy():
    from flow_sim_b200.flatten import flatten_rating
    from flow_sim_b200.hydromodel import RatingCurve

    rng = np.random.default_rng(3)
    rc = RatingCurve(); rc.set("polynomial", a=2.5, b=-3.0, c=11.0, stage_shift=1.5)
    rp = RatingCurve(); rp.set("power", a=4.2, b=1.6, stage_shift=0.3)
    stages = np.linspace(1.0, 9.0, 12)
    rf = RatingCurve(); rf.fit(discharges=3.0 * (stages + 2.0) ** 1.7 + rng.normal(0, .1, 12), stages=stages, stage_shift=2.0,
                              type="polynomial", scale=True, degree=3)
    for curve in (rc, rp, rf):
        d = flatten_rating(curve)
        for s in (1.3, 4.0, 8.8):
            q, dq = oracle_py.rating(d, s)
            assert abs(q - curve.discharge(s)) <= 1e-13 * abs(curve.discharge(s))
            assert abs(dq - curve.dQ_dz(s)) <= 1e-13 * abs(curve.dQ_dz(s))


def test_brentq_restatement_is_bitwise_scipy():
    from scipy.optimize import brentq

    rng = np.random.default_rng(7)
    checked = 0
    for _ in range(400):
        c = rng.normal(size=4) * [5, 3, 1, 0.5]
        f = lambda x: ((c[3] * x + c[2]) * x + c[1]) * x + c[0]      # Horner, same order as the oracle's polyval
        a, b = -3.0, 4.0
        if f(a) * f(b) >= 0:
            continue
        x_ref = brentq(f, a, b)
        x, err = oracle_py.brentq_poly(c, a, b)
        assert err == 0 and x == x_ref
        checked += 1
    assert checked > 50
