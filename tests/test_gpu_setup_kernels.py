"""GPU tests of the kernels either side of the time loop that SURVEY.md 8f ranks next: steady normal-depth
initial state per member (Brent on the device) and the derived result arrays."""
import numpy as np
import pytest

import util
from flow_sim_b200.runner import derived_results, normal_depth_initial_conditions, run_flat

pytestmark = pytest.mark.gpu


def test_normal_depth_profile_matches_host_brentq():
    """akbari: uniform normal depth 0.8637978579... from scipy.optimize.brentq on the host mirror."""
    from flow_sim_b200.cases import build_akbari
    from flow_sim_b200.flatten import flatten_solver

    solver, kw = build_akbari()
    flat = flatten_solver(solver, tolerance=kw["tolerance"])
    h, q = normal_depth_initial_conditions(flat, 1, flat.meta["initial_flow"])
    assert np.max(np.abs(h[0] - flat.ic_depth)) <= 5e-12          # brentq's own xtol is 2e-12
    assert np.array_equal(q[0], flat.ic_flow)
    assert abs(h[0, 0] - 0.863797857935) < 1e-10


def test_normal_depth_per_member_roughness_and_flow():
    """Compound sections, per-member n_main and per-member flow, against the mirror's host implementation."""
    from test_gpu_ensemble import _prismatic
    from flow_sim_b200.hydromodel import TrapezoidalSection

    flat = _prismatic(kind="compound", n_nodes=24, levels=2)
    n = np.array([0.02, 0.03, 0.045])
    q0 = np.array([40.0, 90.0, 260.0])           # the last one is over bank
    flat.member_n_main = n
    h, q = normal_depth_initial_conditions(flat, 3, q0)
    g = flat.geom
    for m in range(3):
        for nd in (0, 11, 23):
            xs = TrapezoidalSection(z_bed=g["z_bed"][nd], b_main=g["b_main"][nd], m_main=g["m_main"][nd], n_main=n[m],
                                    z_bank=g["z_bed"][nd] + g["h_bank"][nd], b_fp_left=g["b_fp_l"][nd],
                                    b_fp_right=g["b_fp_r"][nd], m_fp=g["m_fp"][nd], n_left=g["n_l"][nd],
                                    n_right=g["n_r"][nd], bed_slope=flat.meta["bed_slope"][nd])
            assert abs(h[m, nd] - xs.normal_depth(q0[m])) <= 5e-11, (m, nd)
    assert h[2, 0] > g["h_bank"][0] > h[0, 0]
    assert np.array_equal(q, np.repeat(q0[:, None], 24, axis=1))


def test_derived_results_match_mirror_prepare_results():
    from flow_sim_b200.cases import build_example, build_gerd
    from flow_sim_b200.flatten import flatten_solver

    for builder in (build_example, lambda: build_gerd(n_main=0.03, calibration=True)):
        solver, kw = builder()
        flat = flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw.get("max_iter", 100))
        solver.run(verbose=0, **kw)
        out = derived_results(flat, solver.depth[None], solver.flow[None])
        for name in ("level", "area", "top_width", "froude_number", "velocity", "wave_celerity"):
            assert util.max_rel(out[name][0], getattr(solver, name)) <= 1e-13, name


def test_steady_profile_feeds_the_solver():
    """A roughness ensemble on the akbari channel: device normal-depth profile per member -> solver -> oracle."""
    import oracle_py
    from flow_sim_b200.cases import build_akbari
    from flow_sim_b200.flatten import flatten_solver

    solver, kw = build_akbari()
    flat = flatten_solver(solver, tolerance=kw["tolerance"])
    flat.member_n_main = np.array([0.018, 0.023, 0.035])
    h, q = normal_depth_initial_conditions(flat, 3, flat.meta["initial_flow"])
    flat.ic_depth, flat.ic_flow = h, q
    out = run_flat(flat, n_members=3)
    ora = oracle_py.run(flat, n_members=3)
    util.assert_parity(out["depth"], out["flow"], ora["depth"], ora["flow"], "akbari roughness ensemble")
    assert np.array_equal(out["iters"], ora["iters"])
