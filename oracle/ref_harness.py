"""Live-reference harness (TEST INFRASTRUCTURE - never imported by the product path).

Imports the unmodified reference (`/root/reference`, package ``src.hydromodel`` and the
``cases`` drivers) in THIS container, builds the shipped cases exactly as the reference's
own drivers do, runs ``PreissmannSolver.run`` and records what the parity tests need:

* ``depth`` / ``flow``            - result arrays  (reference ``solver.py:43-44``)
* ``iters``                       - Newton iterations per time level (``preissmann.py:122-161``;
                                    the reference only prints them, so we count ``spsolve`` calls)
* ``storage_stage``               - reservoir stage record (``boundary.py:126-131``)
* optional per-iteration ``(J.data, R, delta)`` captures via an ``spsolve`` hook.

The reference cannot travel to the GPU box, so everything it produces is committed as small
fixtures under ``tests/golden/`` by ``oracle/make_golden.py``.

Caveats handled here: matplotlib/openpyxl/geopandas are absent, so
``matplotlib.pyplot`` is stubbed before the gerd modules are imported; the gerd case uses
Windows path literals, so ``pandas.read_csv`` is wrapped to normalise ``\\`` and the process
changes directory to the reference root.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")     # oracle/stage_reference.py
REFERENCE_ROOT = os.environ.get("PR_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "hydromodel"))


_ready = False


def setup_reference() -> None:
    """Make ``src.hydromodel`` and ``cases.*`` importable, with the stubs described above."""
    global _ready
    if _ready:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)

    # matplotlib is not installed; the gerd helper module imports pyplot at module scope.
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    import pandas

    if not getattr(pandas.read_csv, "_pr_wrapped", False):
        _orig = pandas.read_csv

        def read_csv(path, *a, **k):
            if isinstance(path, str):
                path = path.replace("\\", "/")
            return _orig(path, *a, **k)

        read_csv._pr_wrapped = True
        pandas.read_csv = read_csv
    _ready = True


# --------------------------------------------------------------------------------------
# Case builders: each returns an un-run reference PreissmannSolver plus run() kwargs.
# They restate the reference drivers' *configuration* (not algorithms) because the drivers
# execute at import time and write result files.
# --------------------------------------------------------------------------------------

def build_example():
    """cases/example/main.py:8-57 (config 1)."""
    setup_reference()
    from src.hydromodel.boundary import Boundary
    from src.hydromodel.channel import Channel
    from src.hydromodel.hydrograph import Hydrograph
    from src.hydromodel.lumped_storage import LumpedStorage
    from src.hydromodel.preissmann import PreissmannSolver

    def inflow(t):
        base, peak = 1000, 10000
        rise, hold, fall = 3 * 3600, 6 * 3600, 4 * 3600
        if t <= 0:
            return base
        if t < rise:
            return base + (peak - base) * t / rise
        if t - rise < hold:
            return peak
        if t - rise - hold < fall:
            return peak - (peak - base) * (t - rise - hold) / fall
        return base

    us = Boundary(condition="flow_hydrograph", bed_level=5, chainage=0, hydrograph=Hydrograph(function=inflow))
    ds = Boundary(condition="fixed_depth", initial_depth=5, bed_level=0, chainage=20000)
    ds.set_lumped_storage(LumpedStorage(surface_area=5000 * 250, min_stage=5, solution_boundaries=(0, 200)))
    ch = Channel(width=250, initial_flow=us.hydrograph.get_at(0), roughness=0.027,
                 upstream_boundary=us, downstream_boundary=ds)
    solver = PreissmannSolver(channel=ch, theta=0.8, time_step=3600, spatial_step=1000, simulation_time=24 * 3600)
    return solver, dict(tolerance=1e-4, max_iter=100)


def build_storage_general():
    """Synthetic companion of config 1 for the general lumped-storage boundary: tabulated area
    curve, polynomial outflow rating curve and head losses.  No shipped case exercises these branches
    (lumped_storage.py:24-35,47-143,145-179), so the reference is run on this set-up to pin them."""
    solver, kw = build_example()
    from src.hydromodel.rating_curve import RatingCurve

    ls = solver.channel.downstream_boundary.lumped_storage
    stages = np.arange(0.0, 42.0, 2.0)
    ls.set_area_curve(np.column_stack([stages, 1.25e6 * (1.0 + 0.05 * stages)]), alpha=1.0, beta=0.0)
    rc = RatingCurve()
    rc.set("polynomial", a=20.0, b=10.0, c=0.0)
    ls.rating_curve = rc
    ls.capture_losses = True
    ls.reservoir_length = 2000.0
    ls.K_q = 0.3
    return solver, kw


def build_akbari(peak_flow: float = 200.0, n_nodes_override: int | None = None, **over):
    """cases/akbari_firoozi/settings.py + main_preissmann.py:5-32 (config 2).

    ``over`` lets the config-5 style clones change length / steps / theta for spot checks.
    """
    setup_reference()
    from math import cos, pi, sin

    from src.hydromodel.boundary import Boundary
    from src.hydromodel.channel import Channel
    from src.hydromodel.hydrograph import Hydrograph
    from src.hydromodel.preissmann import PreissmannSolver

    width = over.get("width", 120)
    length = over.get("length", 29000)
    roughness = over.get("roughness", 0.023)
    S_0 = over.get("S_0", 0.00061)
    dx = over.get("spatial_step", 1000)
    duration = over.get("duration", 20 * 3600)
    theta = over.get("theta", 0.5)
    dt = over.get("time_step", 3600)
    tol = over.get("tolerance", 1e-4)
    Q_b = 100

    def hyd(t):
        t_b, t_p = 15 * 3600, 5 * 3600
        if t <= t_p:
            return peak_flow / 2 * sin(pi * t / t_p - pi / 2) + peak_flow / 2 + Q_b
        if t <= t_b:
            return peak_flow / 2 * cos(pi * (t - t_p) / (t_b - t_p)) + peak_flow / 2 + Q_b
        return Q_b

    us = Boundary(condition="flow_hydrograph", bed_level=S_0 * length, chainage=0, hydrograph=Hydrograph(hyd))
    ds = Boundary(condition="normal_depth", bed_level=0, chainage=length)
    ch = Channel(width=width, initial_flow=Q_b, roughness=roughness, upstream_boundary=us,
                 downstream_boundary=ds, interpolation_method="steady-state")
    solver = PreissmannSolver(channel=ch, theta=theta, time_step=dt, spatial_step=dx,
                              simulation_time=duration, regularization=False)
    return solver, dict(tolerance=tol)


def build_gerd(n_main=None, n_fp=None, calibration: bool = False, **over):
    """cases/gerd_roseires/model.py:10-92 (config 3) / n_calibrate.py:5-17 (config 4 member).

    Restates model.run() up to (not including) ``solver.run`` so the solver object can be
    flattened, hooked and timed separately.
    """
    setup_reference()
    from cases.gerd_roseires import settings
    from cases.gerd_roseires.custom_functions import import_hydrograph, import_table, load_trapzoid_xs
    from cases.gerd_roseires.gerd_discharge import GerdHydrograph
    from cases.gerd_roseires.roseires_rating_curve import RoseiresRatingCurve
    from src.hydromodel.boundary import Boundary
    from src.hydromodel.channel import Channel
    from src.hydromodel.hydrograph import Hydrograph
    from src.hydromodel.preissmann import PreissmannSolver

    if calibration:
        inflow_path = "cases\\gerd_roseires\\data\\inflow_hydrograph_small.csv"
        coords_path = None
        sim_duration = None
    else:
        inflow_path = settings.inflow_hyd_path
        coords_path = settings.coords_path
        sim_duration = settings.sim_duration
    sim_duration = over.get("sim_duration", sim_duration)
    coords_path = over.get("coords_path", coords_path)
    time_step = over.get("time_step", settings.time_step)
    level0 = over.get("initial_roseires_level", settings.initial_roseires_level)

    inflow = Hydrograph(table=import_hydrograph(inflow_path))
    duration = int(inflow.table[-1, 0]) if sim_duration is None else int(sim_duration)
    gerd = GerdHydrograph()
    gerd.build(inflow_hydrograph=inflow, time_step=time_step, duration=duration,
               initial_stage=settings.initial_gerd_level)
    q0 = gerd.get_at(time=0)
    chain, sections = load_trapzoid_xs(file_path=settings.cross_sections_path, n_fp=n_fp, n_main=n_main)
    bed = sections[-1].z_min
    up = Boundary(condition="flow_hydrograph", hydrograph=gerd, chainage=chain[0])
    down = Boundary(initial_depth=level0 - bed, bed_level=bed, condition="rating_curve",
                    rating_curve=RoseiresRatingCurve(initial_stage=level0, initial_flow=q0,
                                                     **{**dict(jammed_sluice_gates=settings.JAMMED_SLUICEGATES,
                                                               jammed_spillways=settings.JAMMED_SPILLWAYS),
                                                        **over.get("rating_kwargs", {})}),
                    chainage=chain[-1])
    ch = Channel(initial_flow=q0, upstream_boundary=up, downstream_boundary=down)
    if coords_path is not None:
        coords = import_table(coords_path, sort_by="chainage")
        ch.set_coords(coords=coords[:, 1:], chainages=coords[:, 0])
    ch.set_cross_sections(chainages=chain, sections=sections)
    solver = PreissmannSolver(channel=ch, theta=settings.theta, time_step=time_step,
                              spatial_step=settings.spatial_step, simulation_time=duration)
    solver._pr_first_section_z = sections[0].z_min
    return solver, dict(tolerance=settings.tolerance)


# --------------------------------------------------------------------------------------

def run_and_record(solver, run_kwargs, capture_iterations: int = 0):
    """Run the reference solver; returns dict(depth, flow, iters, seconds, [captures])."""
    from src.hydromodel import preissmann as ref_pr

    counts: dict[int, int] = {}
    captures = []
    orig = ref_pr.spla.spsolve

    def hooked(J, rhs, *a, **k):
        lvl = int(solver.time_level)
        counts[lvl] = counts.get(lvl, 0) + 1
        delta = orig(J, rhs, *a, **k)
        if len(captures) < capture_iterations:
            captures.append(dict(level=lvl, J=np.array(J.data, dtype=np.float64),
                                 R=-np.array(rhs, dtype=np.float64), delta=np.array(delta, dtype=np.float64),
                                 x=np.array(solver.unknowns, dtype=np.float64)))
        return delta

    ref_pr.spla.spsolve = hooked
    t0 = time.perf_counter()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            solver.run(verbose=0, **run_kwargs)
    finally:
        ref_pr.spla.spsolve = orig
    secs = time.perf_counter() - t0
    levels = solver.depth.shape[0]
    iters = np.array([counts.get(k, 0) for k in range(1, levels)], dtype=np.int32)
    out = dict(depth=np.array(solver.depth), flow=np.array(solver.flow), iters=iters, seconds=secs)
    st = getattr(solver, "storage_stage", None)
    if st is not None:
        out["storage_stage"] = np.array(st, dtype=np.float64)
    if captures:
        out["captures"] = captures
    return out


def gerd_calibration_levels(solver, result, Q):
    """cases/gerd_roseires/model.py:105-113: stage at the upstream node interpolated at Q."""
    return np.interp(Q, result["flow"][:, 0], result["depth"][:, 0] + solver._pr_first_section_z)


CALIB_Q = np.array([1562.5, 3850, 6000, 10000, 14000, 21000], dtype=np.float64)        # n_calibrate.py:30
CALIB_H_TARGET = np.array([497.5, 500, 502, 505, 507, 510], dtype=np.float64)           # n_calibrate.py:29


def time_member(n_main: float):
    """One config-4 member through the unmodified reference (used by bench.py --impl reference
    only in this container; on the GPU box the reference is absent and the C port is timed)."""
    solver, kw = build_gerd(n_main=n_main, calibration=True)
    res = run_and_record(solver, kw)
    return res["seconds"], int(res["iters"].sum())


if __name__ == "__main__":
    s, kw = build_example()
    r = run_and_record(s, kw)
    print("example iters", r["iters"].tolist(), r["depth"].sum(), r["flow"].sum(), r["seconds"])


def build_irregular(levee: bool = False, curved: bool = False, pocket: bool = False):
    """Synthetic companion case for IrregularSection (no shipped case instantiates it): a 12 km
    reach between two surveyed-style polylines with composite roughness, interpolated node by node
    (cross_section.py:933-969), flow hydrograph upstream, fixed depth downstream."""
    setup_reference()
    from math import pi, sin

    from src.hydromodel.boundary import Boundary
    from src.hydromodel.channel import Channel
    from src.hydromodel.cross_section import IrregularSection
    from src.hydromodel.hydrograph import Hydrograph
    from src.hydromodel.preissmann import PreissmannSolver

    L, S0, dt = 12000.0, 0.0005, 1800
    wave = lambda t: 60 + 40 * sin(pi * min(t, 6 * dt) / (6 * dt)) ** 2
    us = Boundary("flow_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=2.0, hydrograph=Hydrograph(function=wave))
    ds = Boundary("fixed_depth", chainage=L, bed_level=0.0, initial_depth=2.0)
    ch = Channel(upstream_boundary=us, downstream_boundary=ds, initial_flow=60.0, roughness=0.03, width=30.0,
                 interpolation_method="linear")

    def sec(z0, shift):
        x = np.array([0, 10, 14, 20, 30, 36, 40, 50.0]) + shift
        z = np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]) + z0
        if levee:       # a mid-channel bar that splits low flows into two wetted sub-channels
            x = np.array([0, 10, 14, 20, 24, 26, 30, 36, 40, 50.0]) + shift
            z = np.array([6, 3.0, 1.2, 0.0, 2.6, 2.6, 0.1, 1.5, 3.2, 6.0]) + z0
        if pocket:      # a side pocket behind a ridge on the right bank: a second, small wetted sub-channel at low stages
            # (split-flow conveyance, cross_section.py:329-439) that joins the main channel once the ridge is overtopped;
            # unlike the bar of `levee`, the reference's Newton iteration survives this one
            x = np.array([0, 10, 14, 20, 30, 36, 38, 39, 41, 42, 50.0]) + shift
            z = np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 1.9, 1.9, 3.2, 6.0]) + z0
        s = IrregularSection(x=x, z=z, n=0.03, bed_slope=S0)
        s.set_roughness_para((0.05, 0.03, 0.06, 14.0 + shift, 36.0 + shift))
        return s

    if curved:          # a gentle S-bend: centre-line coordinates -> per-node curvature (channel.py:243-277)
        sx = np.linspace(0.0, L, 25)
        ch.set_coords(coords=np.column_stack([sx, 600.0 * np.sin(2 * np.pi * sx / L)]), chainages=sx * 1.0)
    if curved:          # curvature is computed at the interior input sections only (channel.py:243-277)
        stations = [0.0, 4000.0, 8000.0, L]
        ch.set_cross_sections(stations, [sec(S0 * (L - c), c / L) for c in stations])
    else:
        ch.set_cross_sections([0.0, L], [sec(S0 * L, 0.0), sec(0.0, 1.0)])
    solver = PreissmannSolver(channel=ch, theta=0.6, time_step=dt, spatial_step=1000.0, simulation_time=8 * dt)
    return solver, dict(tolerance=1e-6, max_iter=60)


def build_mixed():
    """A reach that starts on a compound TrapezoidalSection and ends on a surveyed polyline: every interior node is the
    reference's blend of the two (cross_section.py:933-969 - the trapezoid sampled through z_at, :795-849, on the
    polyline's stations), so node 0 is a trapezoid and nodes 1.. are IrregularSections."""
    setup_reference()
    from math import pi, sin

    from src.hydromodel.boundary import Boundary
    from src.hydromodel.channel import Channel
    from src.hydromodel.cross_section import IrregularSection, TrapezoidalSection
    from src.hydromodel.hydrograph import Hydrograph
    from src.hydromodel.preissmann import PreissmannSolver

    L, S0, dt = 12000.0, 0.0005, 1800
    wave = lambda t: 60 + 40 * sin(pi * min(t, 6 * dt) / (6 * dt)) ** 2
    us = Boundary("flow_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=2.0, hydrograph=Hydrograph(function=wave))
    ds = Boundary("fixed_depth", chainage=L, bed_level=0.0, initial_depth=2.0)
    ch = Channel(upstream_boundary=us, downstream_boundary=ds, initial_flow=60.0, roughness=0.03, width=30.0,
                 interpolation_method="linear")
    head = TrapezoidalSection(z_bed=S0 * L, b_main=12.0, m_main=2.0, n_main=0.03, z_bank=S0 * L + 2.4, b_fp_left=6.0,
                              b_fp_right=9.0, m_fp=3.0, n_left=0.05, n_right=0.06, bed_slope=S0)
    tail = IrregularSection(x=np.array([-25, -15, -11, -5, 5, 11, 15, 25.0]), z=np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]),
                            n=0.03, bed_slope=S0)
    tail.set_roughness_para((0.05, 0.03, 0.06, -11.0, 11.0))
    ch.set_cross_sections([0.0, L], [head, tail])
    solver = PreissmannSolver(channel=ch, theta=0.6, time_step=dt, spatial_step=1000.0, simulation_time=8 * dt)
    return solver, dict(tolerance=1e-6, max_iter=60)
