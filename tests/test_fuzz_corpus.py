"""Differential fuzz corpus: 320 random small reaches and 32 of 274..484 nodes - the tiled long-reach path -
and 16 random members / scenarios of the headline reach (tests/fuzz_cases.py) that the LIVE reference ran in the build
container (oracle/fuzz_reference.py --seeds 0:320,1000:1032 --gerd 0:16 --write -> tests/golden/fuzz_corpus.npz: its depth / flow / Newton
iteration counts, or the level it raised in).  The inputs are rebuilt here from the seed on the mirror API and must
hash to the digest of the inputs flattened from the reference's own objects; then the oracle (CPU) and the device path
(GPU, through the C ABI) replay every case."""
import contextlib
import io

import numpy as np
import pytest

import fuzz_cases
import util
from flow_sim_b200.flatten import flatten_solver

CORPUS = np.load(f"{util.GOLD}/fuzz_corpus.npz")
SEEDS = [int(s) for s in CORPUS["seeds"]]


def _inputs(seed):
    with contextlib.redirect_stdout(io.StringIO()):
        solver, kw, d = fuzz_cases.random_case(fuzz_cases.mirror_namespace(), seed)
        flat = flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw["max_iter"])
    assert fuzz_cases.flat_digest(flat) == str(CORPUS[f"s{seed}_digest"]), f"seed {seed}: inputs differ from the reference's"
    return flat, d


GERD_SEEDS = [int(s) for s in CORPUS["gerd_seeds"]]


def _gerd_inputs(seed):
    """A random member / scenario of the headline reach (fuzz_cases.describe_gerd) on the mirror API."""
    from flow_sim_b200.cases import build_gerd

    kwargs = fuzz_cases.describe_gerd(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        solver, kw = build_gerd(**kwargs)
        flat = flatten_solver(solver, tolerance=kw["tolerance"])
    assert fuzz_cases.flat_digest(flat) == str(CORPUS[f"sg{seed}_digest"]), f"gerd seed {seed}: inputs differ from the reference's"
    gated = not kwargs["rating_kwargs"].get("smooth", True)
    return flat, dict(family="gerd" + ("" if kwargs["calibration"] else ", curved"), up="gerd release",
                      down="roseires" + (", gate control" if gated else ""), ic="GVF_equation")


def _check(seed, d, out, what, rtol, iters_exact=True):
    """One replayed case against the reference's record: the same fate (finished / died in the same level), depth and
    flow of every level the reference completed, identical iteration counts."""
    fail_level = int(CORPUS[f"s{seed}_fail_level"])
    tag = f"{what}, seed {seed} ({d['family']}, up {d['up']}, down {d['down']}, ic {d['ic']})"
    depth, flow = CORPUS[f"s{seed}_depth"], CORPUS[f"s{seed}_flow"]
    if fail_level:
        assert out["status"][0] != 0 and int(out["fail_level"][0]) == fail_level, \
            f"{tag}: the reference raised in level {fail_level}, got status {out['status'][0]} level {out['fail_level'][0]}"
    else:
        assert out["status"][0] == 0, f"{tag}: the reference finished, got status {out['status'][0]} in level {out['fail_level'][0]}"
        if iters_exact:
            assert np.array_equal(out["iters"][0], CORPUS[f"s{seed}_iters"]), f"{tag}: iteration counts"
    good = depth.shape[0]
    util.assert_parity(out["depth"][0][:good], out["flow"][0][:good], depth, flow, tag, rtol=rtol)


def test_corpus_covers_the_configuration_space():
    fam, down, up, ic, fate = set(), set(), set(), set(), [0, 0]
    for s in SEEDS:
        d = fuzz_cases.describe(s)
        fam.add(d["family"]); down.add(d["down"]); up.add(d["up"]); ic.add(d["ic"])
        fate[int(CORPUS[f"s{s}_fail_level"]) != 0] += 1
    assert fam == set(fuzz_cases.FAMILIES) and down == set(fuzz_cases.DOWNSTREAM) | {"storage_general"}
    assert up == set(fuzz_cases.UPSTREAM) | {"fixed_depth", "normal_depth"}
    assert ic == {"linear", "GVF_equation", "steady-state"}
    assert fate[0] >= 100 and fate[1] >= 50          # runs the reference finishes, and runs it dies in


def test_mirror_refuses_what_the_reference_refuses():
    for seed in CORPUS["refused"]:
        with pytest.raises(Exception), contextlib.redirect_stdout(io.StringIO()):
            solver, kw, _ = fuzz_cases.random_case(fuzz_cases.mirror_namespace(), int(seed))
            flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw["max_iter"])


def test_oracle_replays_the_corpus():
    import oracle_py

    for seed in SEEDS:
        flat, d = _inputs(seed)
        _check(seed, d, oracle_py.run(flat, 1), "oracle", rtol=1e-10)
    assert len(GERD_SEEDS) >= 12
    for seed in GERD_SEEDS:
        flat, d = _gerd_inputs(seed)
        _check(f"g{seed}", d, oracle_py.run(flat, 1), "oracle", rtol=1e-10)


@pytest.mark.gpu
def test_device_replays_the_corpus():
    """Every case through the C ABI.  Iteration counts must be the reference's except at a near tie of the convergence
    test in the oracle's own record (util.assert_iteration_parity)."""
    import oracle_py
    from flow_sim_b200.abi import PreissmannLibraryError
    from flow_sim_b200.runner import run_flat

    flips = 0
    for seed in SEEDS:
        flat, d = _inputs(seed)
        try:        # the family the library picks, and every packing that holds the reach (others are refused loudly)
            out = run_flat(flat, lanes=(0, 8, 16, 32)[seed % 4])
        except PreissmannLibraryError as e:
            assert "instantiation holds" in str(e)
            out = run_flat(flat)
        if int(CORPUS[f"s{seed}_fail_level"]) == 0 and out["status"][0] == 0 and \
                not np.array_equal(out["iters"][0], CORPUS[f"s{seed}_iters"]):
            ora = oracle_py.run(flat, 1, trace_prev_error=True)
            flips += util.assert_iteration_parity(out, ora, flat.tol, f"device, seed {seed}")
            continue
        _check(seed, d, out, "device", rtol=util.RTOL)
    for seed in GERD_SEEDS:             # the headline reach: roughness off the grid, floodplain overrides, gate control
        flat, d = _gerd_inputs(seed)
        out = run_flat(flat)
        if not np.array_equal(out["iters"][0], CORPUS[f"sg{seed}_iters"]):
            ora = oracle_py.run(flat, 1, trace_prev_error=True)
            flips += util.assert_iteration_parity(out, ora, flat.tol, f"device, gerd seed {seed}")
            continue
        _check(f"g{seed}", d, out, "device", rtol=util.RTOL)
    assert flips <= 2


@pytest.mark.gpu
def test_device_ensembles_on_corpus_configurations():
    """Ensembles on every third short configuration of the corpus: 21 members with their own main-channel and floodplain
    roughness and their own upstream series, so that one warp of a packed family holds members of different fate (some
    die in level 1, some finish) and different iteration counts.  Against the oracle's run of the same members."""
    import oracle_py
    from flow_sim_b200.abi import PreissmannLibraryError
    from flow_sim_b200.runner import run_flat

    M, flips, died, finished = 21, 0, 0, 0
    for seed in [s for s in SEEDS if s < fuzz_cases.LONG_SEEDS and s % 3 == 0]:
        flat, d = _inputs(seed)
        rng = np.random.default_rng(10_000 + seed)
        flat.member_n_main = d["n_main"] * rng.uniform(0.7, 1.4, M)
        if seed % 2:
            flat.member_n_fp = d["n_fp"] * rng.uniform(0.7, 1.4, M)
        if flat.up.series is not None:          # (a fixed-depth / normal-depth upstream end has no series)
            flat.up.series = flat.up.series[None, :] * (1.0 + (0.1 if d["up"] == "flow_hydrograph" else 0.01) * rng.uniform(-1, 1, (M, 1)))
        ora = oracle_py.run(flat, M, trace_prev_error=True)
        try:
            out = run_flat(flat, n_members=M, lanes=(0, 8, 16, 32)[(seed // 3) % 4])
        except PreissmannLibraryError as e:
            assert "instantiation holds" in str(e)
            out = run_flat(flat, n_members=M)
        what = f"ensemble on seed {seed} ({d['family']}, up {d['up']}, down {d['down']})"
        assert np.array_equal(out["status"] != 0, ora["status"] != 0), what
        assert np.array_equal(out["fail_level"], ora["fail_level"]), what
        ok = np.nonzero(ora["status"] == 0)[0]
        died += M - len(ok); finished += len(ok)
        if len(ok):
            flips += util.assert_iteration_parity(out, ora, flat.tol, what, members=ok)
        # the levels a failed member completed: a member on its way to blowing up is ill-conditioned already (seed 237:
        # 6e-9 one level before it dies), so these are held to FLIP_RTOL, not to the 1e-9 of runs that finish
        for m in np.nonzero(ora["status"] != 0)[0]:
            k = int(ora["fail_level"][m])
            util.assert_parity(out["depth"][m][:k], out["flow"][m][:k], ora["depth"][m][:k], ora["flow"][m][:k], what, rtol=util.FLIP_RTOL)
    assert died > 500 and finished > 1000 and flips <= 4


@pytest.mark.gpu
def test_device_initial_conditions_on_the_corpus():
    """The device's GVF march (channel.py:307-378) and normal-depth solve (cross_section.py:184-205) against the initial
    state the REFERENCE computed for the corpus configurations (it is part of the digested inputs): every section
    family, 5-484 nodes."""
    from flow_sim_b200.runner import gvf_initial_conditions, normal_depth_initial_conditions

    n_gvf = n_normal = 0
    for seed in SEEDS:
        d = fuzz_cases.describe(seed)
        if d["ic"] == "linear":
            continue
        flat, d = _inputs(seed)
        if d["ic"] == "GVF_equation":
            h, q, st = gvf_initial_conditions(flat, 1, flat.meta["initial_flow"], flat.meta["downstream_depth"])
            assert st[0] == 0, f"seed {seed}: GVF status {st[0]}"
            # 1e-16 typical; on a grid too coarse for the explicit predictor-corrector the march itself is unstable and
            # amplifies rounding (seed 312: dx = 2 km, depths jumping between 2.5 and 10.8 m, 2e-10)
            tol, n_gvf = util.RTOL, n_gvf + 1
        else:
            h, q = normal_depth_initial_conditions(flat, 1, flat.meta["initial_flow"])
            tol, n_normal = 5e-11, n_normal + 1          # brentq's xtol is 2e-12 absolute
        err = float(np.max(np.abs(h[0] - flat.ic_depth) / flat.ic_depth))
        assert err <= tol, f"seed {seed} ({d['family']}, {d['ic']}): initial depth off by {err:.3g}"
        assert np.array_equal(q[0], flat.ic_flow)
    assert n_gvf >= 60 and n_normal >= 70


@pytest.mark.gpu
def test_device_derived_arrays_on_the_corpus():
    """pr_derived_results on the reference's own depth / flow against the arrays its Solver.prepare_results made
    (solver.py:65-98; stored for every fourth finished run): area, geometric top width, Froude number, wave celerity."""
    from flow_sim_b200.runner import derived_results

    n = 0
    for seed in SEEDS:
        if f"s{seed}_area" not in CORPUS.files:
            continue
        flat, d = _inputs(seed)
        got = derived_results(flat, CORPUS[f"s{seed}_depth"][None], CORPUS[f"s{seed}_flow"][None])
        for k in ("area", "top_width", "froude_number", "wave_celerity"):
            ref = CORPUS[f"s{seed}_{k}"]
            err = float(np.max(np.abs(got[k][0] - ref) / np.maximum(np.abs(ref), 1e-12)))
            assert err <= 1e-12, f"seed {seed} ({d['family']}): {k} off by {err:.3g}"
        n += 1
    assert n >= 40


@pytest.mark.gpu
def test_device_roughness_sweeps_on_corpus_configurations():
    """The calibration workflow (EnsembleRunner.roughness_sweep: device GVF profile per member -> Newton run, all on the
    device, members launched in descending-roughness order) on the corpus configurations that start from a GVF profile,
    against the oracle's GVF + run of the same members."""
    import oracle_py
    from flow_sim_b200 import abi
    from flow_sim_b200.ensemble import EnsembleRunner, to_host

    M, n, flips = 9, 0, 0
    for seed in [s for s in SEEDS if fuzz_cases.describe(s)["ic"] == "GVF_equation" and s < fuzz_cases.LONG_SEEDS]:
        flat, d = _inputs(seed)
        rng = np.random.default_rng(20_000 + seed)
        n_main = d["n_main"] * rng.uniform(0.8, 1.3, M)
        n_fp = d["n_fp"] * rng.uniform(0.8, 1.3, M) if seed % 2 else None
        res = to_host(EnsembleRunner(flat, "cuda:0").roughness_sweep(n_main, n_fp=n_fp, out_mode=abi.PR_OUT_FULL))
        flat.member_n_main, flat.member_n_fp = n_main, n_fp
        ho, qo, sto = oracle_py.gvf(flat, flat.meta["initial_flow"], flat.meta["downstream_depth"], n_members=M)
        what = f"roughness sweep on seed {seed} ({d['family']}, up {d['up']}, down {d['down']})"
        assert np.array_equal(res["ic_status"] != 0, sto != 0), what
        # members whose profile the GVF march accepts (subcritical) and marches stably: on a grid too coarse for the
        # explicit predictor-corrector the profile zigzags by metres and amplifies rounding (seed 148, 2e-9; the
        # profiles themselves are held to the reference's in test_device_initial_conditions_on_the_corpus)
        keep = np.nonzero((sto == 0) & (np.abs(np.diff(ho, axis=1)).max(axis=1) < 1.0))[0]
        flat.ic_depth, flat.ic_flow = ho, qo
        ora = oracle_py.run(flat, n_members=M, trace_prev_error=True)
        assert np.all(res["status"][sto != 0] != 0), what                 # rejected profiles come back failed
        assert np.array_equal(res["status"][keep] != 0, ora["status"][keep] != 0), what
        ok = keep[ora["status"][keep] == 0]
        if len(ok):
            flips += util.assert_iteration_parity(res, ora, flat.tol, what, members=ok)
            n += len(ok)
    assert n >= 200 and flips <= 3


@pytest.mark.gpu
@pytest.mark.parametrize("seed,lanes", [(411, 32), (544, 0)])
def test_members_the_reference_loses_stay_lost(seed, lanes):
    """Two ensembles of tools/fuzz_device.py: every member leaves a node dry in level 1 - the reference (and the
    oracle) clamps the depth to zero and dies of the zero conveyance - yet one member used to run on to the end on the
    device, because a depth below -b / (2 sqrt(1 + m^2)) makes A and P both negative and their ratio positive
    (`poison_dry`, pr_device.cuh)."""
    import oracle_py
    from flow_sim_b200.runner import run_flat

    d = fuzz_cases.describe(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        solver, kw, _ = fuzz_cases.random_case(fuzz_cases.mirror_namespace(), seed)
        flat = flatten_solver(solver, tolerance=kw["tolerance"], max_iter=kw["max_iter"])
    rng = np.random.default_rng(30_000 + seed)
    M = 13
    flat.member_n_main = d["n_main"] * rng.uniform(0.7, 1.4, M)
    flat.member_n_fp = d["n_fp"] * rng.uniform(0.7, 1.4, M) if seed % 2 else None
    ora = oracle_py.run(flat, M)
    assert np.all(ora["status"] != 0) and np.all(ora["fail_level"] == 1)
    out = run_flat(flat, n_members=M, lanes=lanes)
    assert np.array_equal(out["status"], ora["status"]) and np.array_equal(out["fail_level"], ora["fail_level"])
    assert np.all(np.isnan(out["depth"][:, 1:]))
