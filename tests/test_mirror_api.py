"""Host-side mirror of the reference API (flow_sim_b200.hydromodel) and the case builders: the flattened
inputs they produce must be bit-identical to those flattened from the reference's own objects
(tests/golden/*.in.npz, made in the build container by oracle/make_golden.py)."""
import sys

import numpy as np
import pytest

import util
from flow_sim_b200 import abi, hydromodel
from flow_sim_b200.cases import build_akbari, build_example, build_gerd, build_irregular, build_mixed
from flow_sim_b200.cases import akbari_firoozi, gerd_roseires
from flow_sim_b200.flatten import flatten_solver, load_flat, save_flat


def _same_flat(a, b):
    assert (a.n_nodes, a.n_levels, a.theta, a.dt, a.dx, a.tol, a.max_iter, a.g) == \
           (b.n_nodes, b.n_levels, b.theta, b.dt, b.dx, b.tol, b.max_iter, b.g)
    for k in a.geom:
        assert np.array_equal(a.geom[k], b.geom[k]), f"geometry field {k}"
    assert np.array_equal(a.ic_depth, b.ic_depth) and np.array_equal(a.ic_flow, b.ic_flow)
    for x, y in ((a.up, b.up), (a.down, b.down)):
        assert x.type == y.type
        for f in ("bed_level", "bed_slope", "fixed_depth", "storage_area", "storage_min_stage", "storage_ymin", "storage_ymax"):
            u, v = getattr(x, f), getattr(y, f)
            assert u == v or (u != u and v != v), f
        assert (x.series is None) == (y.series is None)
        if x.series is not None:
            assert np.array_equal(x.series, y.series)
        assert (x.rating is None) == (y.rating is None)
        if x.rating:
            for k, v in x.rating.items():
                assert np.array_equal(np.asarray(v, float), np.asarray(y.rating[k], float)), f"rating {k}"


@pytest.mark.parametrize("case,builder", [
    ("example", lambda: build_example()),
    ("akbari", lambda: build_akbari()),
    ("akbari_long", lambda: build_akbari(length=200000, spatial_step=100, time_step=600, duration=16 * 600, theta=0.6)),
    ("gerd_calib_m0", lambda: build_gerd(n_main=util.calib_n(0), calibration=True)),
    ("gerd_calib_m36408", lambda: build_gerd(n_main=util.calib_n(36408), calibration=True)),
    ("gerd_full", lambda: build_gerd()),
    ("irregular", lambda: build_irregular()),
    ("irregular_levee", lambda: build_irregular(bar=True)),
    ("irregular_curved", lambda: build_irregular(curved=True)),
    ("irregular_pocket", lambda: build_irregular(pocket=True)),
    ("mixed_sections", lambda: build_mixed()),      # trapezoid blended with a polyline (cross_section.py:795-849, 933-969)
])
def test_case_builders_reproduce_reference_inputs(case, builder):
    solver, kw = builder()
    flat = flatten_solver(solver, tolerance=kw.get("tolerance", 1e-4), max_iter=kw.get("max_iter", 100))
    _same_flat(util.golden_inputs(case), flat)


def _storage_general_solver():
    """The general lumped-storage set-up of oracle/ref_harness.build_storage_general on the mirror API."""
    solver, kw = build_example()
    ls = solver.channel.downstream_boundary.lumped_storage
    stages = np.arange(0.0, 42.0, 2.0)
    ls.set_area_curve(np.column_stack([stages, 1.25e6 * (1.0 + 0.05 * stages)]), alpha=1.0, beta=0.0)
    rc = hydromodel.RatingCurve()
    rc.set("polynomial", a=20.0, b=10.0, c=0.0)
    ls.rating_curve, ls.capture_losses, ls.reservoir_length, ls.K_q = rc, True, 2000.0, 0.3
    return solver, kw


def test_general_storage_flattens_like_the_reference(tmp_path):
    solver, kw = _storage_general_solver()
    flat = flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw["max_iter"])
    ref = util.golden_inputs("storage_general")
    _same_flat(ref, flat)
    for f in ("storage_alpha", "storage_beta", "storage_losses", "storage_reservoir_length", "storage_Kq"):
        assert getattr(ref.down, f) == getattr(flat.down, f)
    assert np.array_equal(ref.down.storage_curve, flat.down.storage_curve)
    assert ref.down.storage_outflow["type"] == flat.down.storage_outflow["type"] == abi.PR_RC_POLY2
    p = str(tmp_path / "s.npz")
    save_flat(p, flat)
    assert np.array_equal(load_flat(p).down.storage_curve, flat.down.storage_curve)


def test_irregular_section_mirror_equals_the_oracle_restatement():
    """IrregularSection on the mirror API (host-side set-up code) against the oracle's restatement, which is
    bit-identical with the live reference (oracle/ref_harness, checked when the goldens were made)."""
    import oracle_py

    solver, kw = build_irregular()
    flat = flatten_solver(solver, **kw)
    assert (flat.geom["kind"] == abi.PR_XS_IRREGULAR).all()
    rng = np.random.default_rng(2)
    for node in (0, 5, 12):
        xs = solver.channel.xs_at_node[node]
        for _ in range(25):
            h, Q = rng.uniform(0.3, 5.5), rng.uniform(20, 200)
            hw = h + xs.z_min
            got = oracle_py.section_probe(flat, node, h, Q)
            mine = dict(A=xs.area(hw), P=xs.wetted_perimeter(hw), T=xs.top_width(hw), K=xs.conveyance(hw),
                        n_eq=xs.get_equivalent_n(hw), dR_dA=xs.dR_dA(hw), dK_dA=xs.dK_dA(hw), Sf=xs.friction_slope(h, Q),
                        dSf_dA=xs.dSf_dA(h, Q), dSf_dQ=xs.dSf_dQ(h, Q))
            for k, v in mine.items():
                assert abs(got[k] - v) <= 1e-13 * abs(v), k
    bar = build_irregular(bar=True)[0].channel.xs_at_node[4]
    assert bar.sub_channels(bar.z_min + 1.0) == 2 and bar.sub_channels(bar.z_min + 4.0) == 1


def test_split_flow_conveyance_equals_the_reference():
    """Several wetted sub-channels (cross_section.py:329-439): friction slope and its derivatives of the side-pocket
    sections, on the oracle and on the mirror, against values computed by the live reference when the fixture was made
    (tests/golden/irregular_pocket_probe.npz: 252 probes, 114 of them split) - bit for bit."""
    import oracle_py

    fx = np.load(f"{util.GOLD}/irregular_pocket_probe.npz")
    rows = fx["rows"]
    assert (rows[:, 3] > 1).sum() > 100
    flat = util.golden_inputs("irregular_pocket")
    sections = build_irregular(pocket=True)[0].channel.xs_at_node
    for node, h, Q, nsub, Sf, dSfA, dSfQ, K, dKA, A, dAdh in rows:
        got = oracle_py.section_probe(flat, int(node), h, Q)
        assert (got["Sf"], got["dSf_dA"], got["dSf_dQ"]) == (Sf, dSfA, dSfQ), (node, h)
        xs = sections[int(node)]
        assert len(xs.get_subchannels(h + xs.z_min)) == int(nsub)
        assert (xs.friction_slope(h, Q), xs.dSf_dA(h, Q), xs.dSf_dQ(h, Q)) == (Sf, dSfA, dSfQ), (node, h)
        assert (xs.conveyance(h + xs.z_min), xs.dK_dA(h + xs.z_min)) == (K, dKA)


def _dense_pocket_case(k):
    """The side-pocket reach with every polyline segment cut into k pieces (util.densify_polylines)."""
    flat = util.golden_inputs("irregular_pocket")
    util.densify_polylines(flat, k)
    return flat


def test_split_flow_on_dense_polylines_equals_the_reference():
    """A 381-point version of a side-pocket section (tests/golden/irregular_dense_probe.npz, oracle/make_probes.py):
    sums of more than 128 terms take numpy's recursive pairwise order, sub-channels hold ~100 points.  Oracle and mirror
    against the live reference's values - bit for bit."""
    import oracle_py

    fx = np.load(f"{util.GOLD}/irregular_dense_probe.npz")
    rows = fx["rows"]
    assert (rows[:, 3] > 1).sum() >= 5 and len(fx["x"]) > 300
    par = fx["roughness"]
    xs = hydromodel.IrregularSection(x=fx["x"], z=fx["z"], n=par[1], bed_slope=5e-4)
    xs.set_roughness_para(tuple(par))
    flat = util.golden_inputs("irregular_pocket")
    util.replace_polyline(flat, 7, fx["x"], fx["z"], par[3], par[4])
    for node, h, Q, nsub, Sf, dSfA, dSfQ, K, dKA, A, dAdh in rows:
        got = oracle_py.section_probe(flat, 7, h, Q)
        assert (got["Sf"], got["dSf_dA"], got["dSf_dQ"]) == (Sf, dSfA, dSfQ), h
        assert (xs.friction_slope(h, Q), xs.dSf_dA(h, Q), xs.dSf_dQ(h, Q)) == (Sf, dSfA, dSfQ), h
        assert len(xs.get_subchannels(h + xs.z_min)) == int(nsub)


def test_trapezoid_bed_profile_and_mixed_blend_equal_the_reference():
    """`TrapezoidalSection.z_at` (cross_section.py:795-849) and the blend of a trapezoid with a polyline
    (`interpolate_cross_section`, :933-969, either order, three weights) against the live reference's values
    (tests/golden/trapezoid_z_at_probe.npz, oracle/make_probes.py) - bit for bit."""
    fx = np.load(f"{util.GOLD}/trapezoid_z_at_probe.npz")
    T, I = hydromodel.TrapezoidalSection, hydromodel.IrregularSection
    shapes = dict(rect=T(z_bed=1.5, b_main=12.0, m_main=0.0, n_main=0.03), simple=T(z_bed=1.5, b_main=12.0, m_main=2.0, n_main=0.03),
                  compound=T(z_bed=1.5, b_main=12.0, m_main=2.0, n_main=0.03, z_bank=3.9, b_fp_left=6.0, b_fp_right=9.0, m_fp=3.0,
                             n_left=0.05, n_right=0.06, bed_slope=5e-4))
    for name, xs in shapes.items():
        assert np.array_equal(np.array([xs.z_at(v) for v in fx["grid"]]), fx[f"z_{name}"]), name
    assert np.isinf(fx["z_rect"]).any() and (fx["z_compound"] == 3.9).any()
    poly = I(x=np.array([-25, -15, -11, -5, 5, 11, 15, 25.0]), z=np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]), n=0.03, bed_slope=5e-4)
    poly.set_roughness_para((0.05, 0.03, 0.06, -11.0, 11.0))
    for j, (d1, d2) in enumerate(fx["blend_dists"]):
        for tag, (a, b) in (("tp", (shapes["compound"], poly)), ("pt", (poly, shapes["compound"]))):
            s = hydromodel.interpolate_cross_section(a, b, d1, d2)
            assert isinstance(s, I)
            assert np.array_equal(s.x, fx[f"blend_{tag}{j}_x"]) and np.array_equal(s.z, fx[f"blend_{tag}{j}_z"]), (tag, j)
            par = [s.n_left, s.n_main, s.n_right, s.left_fp_limit, s.right_fp_limit, s.bed_slope, s.curvature]
            assert np.array_equal(np.array(par, dtype=np.float64), fx[f"blend_{tag}{j}_par"]), (tag, j)


def test_flat_roundtrip(tmp_path):
    solver, kw = build_gerd(n_main=0.03, calibration=True)
    flat = flatten_solver(solver, tolerance=kw["tolerance"])
    p = str(tmp_path / "c.npz")
    save_flat(p, flat)
    _same_flat(flat, load_flat(p))
    assert load_flat(p).meta["downstream_depth"] == flat.meta["downstream_depth"]
    solver, kw = build_irregular()
    flat = flatten_solver(solver, **kw)
    save_flat(p, flat)
    back = load_flat(p)
    _same_flat(flat, back)
    assert back.geom["irr_offset"].dtype == np.int32 and np.array_equal(back.geom["irr_x"], flat.geom["irr_x"])


def test_grid_sizing_matches_reference_rules():
    """N = round(L/dx)+1, dx = L/(N-1), levels = T//dt + 1 (solver.py:34-35,53-55)."""
    s, _ = build_gerd(calibration=True)
    assert s.number_of_nodes == 121 and s.number_of_time_levels == 33
    assert s.spatial_step == s.channel.length / 120
    s, _ = build_example()
    assert (s.number_of_nodes, s.number_of_time_levels) == (21, 25)
    assert s.depth.shape == (25, 21) and s.depth.flags["C_CONTIGUOUS"] and s.depth.dtype == np.float64


def test_unsupported_configurations_are_rejected_not_emulated():
    s, _ = build_example()
    s.regularization = True
    with pytest.raises(NotImplementedError):
        flatten_solver(s)
    s, _ = build_example()
    s.channel.downstream_boundary.lumped_storage.capture_losses = True      # losses need a reservoir length
    with pytest.raises(ValueError):
        flatten_solver(s)

    class Irregular:          # stands for the reference's IrregularSection (no trapezoid attributes)
        z_min = 0.0

    s, _ = build_example()
    s.channel.xs_at_node[3] = Irregular()
    with pytest.raises(NotImplementedError, match="TrapezoidalSection"):
        flatten_solver(s)
    s, _ = build_gerd(calibration=True)
    rc = s.channel.downstream_boundary.rating_curve
    rc.smooth = False
    assert flatten_solver(s).down.rating["gate_control"] == 1          # a fresh gate-controlled curve is supported ...
    rc.discharge(stage=rc.initial_stage + 0.7, time=3600)               # ... one that has already been stepped is not
    with pytest.raises(NotImplementedError, match="already been stepped"):
        flatten_solver(s)
    with pytest.raises(ValueError, match="Invalid boundary condition"):
        hydromodel.Boundary(condition="weir", chainage=0)
    with pytest.raises(ValueError, match="Invalid interpolation method"):
        hydromodel.Channel(hydromodel.Boundary("fixed_depth", 0, 0, 1), hydromodel.Boundary("fixed_depth", 10, 0, 1), 1.0,
                           interpolation_method="spline")


def test_install_registers_reference_import_names():
    hydromodel.install()
    from src.hydromodel.channel import Channel            # the import lines of cases/example/main.py:1-5
    from src.hydromodel.preissmann import PreissmannSolver
    import hydromodel as hm

    assert Channel is hydromodel.Channel and PreissmannSolver is hydromodel.PreissmannSolver and hm is hydromodel
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k.startswith("hydromodel")]:
        del sys.modules[k]


def test_roseires_fit_and_gate_state():
    rc = gerd_roseires.RoseiresRatingCurve(initial_stage=487.0, initial_flow=2094.106301)
    # one fully open spillway, a partial opening that rounds to 0.0 and therefore contributes nothing
    # (quirk 10, roseires_rating_curve.py:167-173), no sluice
    assert rc.closed_state == ([13, 0.0, 0, 0, 0, 0, 0], 0)
    assert rc.open_state == ([13] * 7, 5)
    assert abs(rc.discharge(487.0) - 2090.155184) < 1e-5      # low_release_rating_curve.csv, Y = 487
    assert abs(rc.discharge(487.5) - rc.release(487.5, rc.open_state)) < 1e-9


def test_long_reach_builder_shapes():
    s, _ = akbari_firoozi.build_long_reach(n_nodes=501, n_steps=4)
    assert s.number_of_nodes == 501 and s.number_of_time_levels == 5
    assert np.allclose(s.channel.initial_conditions[:, 0], s.channel.initial_conditions[0, 0])


def test_vectorised_long_reach_builder_is_bit_identical_with_the_object_model():
    """Config 5's set-up without 100 000 section objects: every geometry column, the boundaries and the hydrograph
    table equal what flattening the object model gives (2 001-node clone), bit for bit."""
    from flow_sim_b200.cases.akbari_firoozi import (build_long_reach, build_long_reach_flat, flood_wave,
                                                    flood_wave_series)
    from flow_sim_b200.flatten import flatten_solver

    solver, kw = build_long_reach(n_nodes=2001, n_steps=16)
    a = flatten_solver(solver, tolerance=kw["tolerance"])
    b = build_long_reach_flat(n_nodes=2001, n_steps=16)
    assert set(a.geom) == set(b.geom)
    for k in a.geom:
        assert np.array_equal(a.geom[k], b.geom[k]) and a.geom[k].dtype == b.geom[k].dtype, k
    assert (a.n_nodes, a.n_levels, a.theta, a.dt, a.dx, a.tol, a.max_iter, a.g) == (b.n_nodes, b.n_levels, b.theta, b.dt, b.dx, b.tol, b.max_iter, b.g)
    for ba, bb in ((a.up, b.up), (a.down, b.down)):
        assert ba.type == bb.type and ba.bed_level == bb.bed_level
        assert (ba.series is None and bb.series is None) or np.array_equal(ba.series, bb.series)
    assert a.down.bed_slope == b.down.bed_slope
    assert np.array_equal(a.meta["bed_slope"], b.meta["bed_slope"]) and a.meta["z0"] == b.meta["z0"]
    peaks = [100.0, 237.5, 300.0]
    ref = np.array([[flood_wave(peak_flow=p)(k * a.dt) for k in range(a.n_levels)] for p in peaks])
    assert np.array_equal(flood_wave_series(peaks, a.n_levels, a.dt), ref)
