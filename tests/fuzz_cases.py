"""Random small reaches for differential testing.  ``random_case(ns, seed)`` builds the SAME configuration on any
implementation of the hydromodel API handed in as a namespace: the live reference (oracle/fuzz_reference.py, in the build
container) or the mirror (tests, on the GPU box).  Every draw comes from ``numpy.random.default_rng(seed)``; nothing
else is random.  The generator aims at the branches the shipped cases visit rarely: every section family and their
blends, over-bank flow, every boundary-condition type on either end, all three initial-condition methods, theta and
step sizes over their useful range."""
import hashlib
from math import pi, sin
from types import SimpleNamespace

import numpy as np


def mirror_namespace():
    from flow_sim_b200 import hydromodel as hm

    return SimpleNamespace(Boundary=hm.Boundary, Channel=hm.Channel, Hydrograph=hm.Hydrograph, LumpedStorage=hm.LumpedStorage,
                           PreissmannSolver=hm.PreissmannSolver, RatingCurve=hm.RatingCurve,
                           TrapezoidalSection=hm.TrapezoidalSection, IrregularSection=hm.IrregularSection)


FAMILIES = ["width", "simple", "rect_sections", "compound", "polyline", "trapezoid_to_polyline", "polyline_to_trapezoid",
            "compound_to_simple"]
DOWNSTREAM = ["fixed_depth", "normal_depth", "rating_curve", "stage_hydrograph", "storage"]       # + "storage_general", drawn late
UPSTREAM = ["flow_hydrograph", "flow_hydrograph", "flow_hydrograph", "stage_hydrograph"]
LONG_SEEDS = 1000      # seeds from here on build reaches of 274..484 nodes


def describe(seed):
    """The draws of a seed as a dict (what random_case builds) - also the label of a test case."""
    rng = np.random.default_rng(seed)
    d = dict(seed=int(seed))
    d["family"] = FAMILIES[int(rng.integers(len(FAMILIES)))]
    d["down"] = DOWNSTREAM[int(rng.integers(len(DOWNSTREAM)))]
    d["up"] = UPSTREAM[int(rng.integers(len(UPSTREAM)))]
    d["n_cells"] = int(rng.integers(4, 40))
    if seed >= LONG_SEEDS:          # reaches beyond the fused kernel's 249 nodes: the tiled long-reach path
        d["n_cells"] = int(250 + 6 * d["n_cells"])
    d["dx"] = float(rng.choice([250.0, 500.0, 1000.0, 2000.0]))
    d["dt"] = float(rng.choice([300.0, 600.0, 1800.0, 3600.0]))
    d["theta"] = float(rng.choice([0.55, 0.6, 0.75, 0.9, 1.0]))
    d["levels"] = int(rng.integers(3, 9))
    d["slope"] = float(rng.choice([1e-4, 2e-4, 5e-4, 1e-3]))
    d["ic"] = ["linear", "GVF_equation", "steady-state"][int(rng.integers(3))]
    d["q_base"] = float(rng.uniform(40.0, 120.0))
    d["q_peak"] = float(d["q_base"] * rng.uniform(1.2, 3.0))
    d["depth0"] = float(rng.choice([1.5, 2.0, 2.5, 3.0]))          # round values on purpose (vertex-level ties)
    d["n_main"] = float(rng.uniform(0.02, 0.04))
    d["n_fp"] = float(rng.uniform(0.04, 0.08))
    d["tol"] = float(rng.choice([1e-4, 1e-6]))
    d["shape"] = [float(v) for v in rng.uniform(0.0, 1.0, 8)]
    if d["family"] in ("polyline", "trapezoid_to_polyline", "polyline_to_trapezoid") and d["ic"] != "linear":
        d["ic"] = "linear" if rng.uniform() < 0.5 else d["ic"]
    d["curved"] = bool(rng.uniform() < 0.3) and d["family"] != "width" and d["n_cells"] >= 12     # centre-line curvature slope
    d["rating"] = ["power", "polynomial"][int(rng.integers(2))]
    # later additions draw last, so that the earlier draws of a seed stay what they were
    v = rng.uniform(0.0, 1.0, 4)
    if d["down"] == "storage" and v[0] < 0.6:        # general lumped storage: area curve, outflow curve, head losses
        d["down"] = "storage_general"
    d["storage"] = dict(slope=float(0.02 + 0.08 * v[1]), losses=bool(v[2] < 0.6), k_q=float(0.5 * v[3]))
    if d["up"] == "flow_hydrograph" and v[3] > 0.85:
        d["up"] = ["fixed_depth", "normal_depth"][int(v[2] < 0.5)]
    # a side pocket behind a ridge on the right floodplain of the surveyed sections: a second wetted sub-channel at low
    # stages, i.e. the split-flow conveyance (cross_section.py:329-439)
    d["pocket"] = bool(rng.uniform() < 0.4) and d["family"] in ("polyline", "trapezoid_to_polyline", "polyline_to_trapezoid")
    return d


def _polyline(ns, d, invert, k):
    """A surveyed-style section centred on 0: main channel between the banks at +-(9..13), floodplains beyond."""
    u = d["shape"]
    half = 9.0 + 4.0 * u[(k + 1) % 8]
    x = np.array([-30.0, -half - 8.0, -half, -half + 4.0, -1.0 - 2 * u[k % 8], 2.0 + 2 * u[(k + 2) % 8], half - 3.0, half, half + 7.0, 30.0])
    # no station elevation equals a depth a boundary is HELD at (initial depths, the stage hydrographs' tenths): such a
    # node sits on the jump of the reference's properties() (see irr_tab_tie) and the reference's own run then depends
    # on the last bit of the boundary depth (seeds 150 and 178 of an earlier generator)
    z = np.array([7.0, 3.0 + u[(k + 3) % 8], 2.42, 1.0, 0.0, 0.1 * round(3 * u[(k + 4) % 8]), 1.0 + 0.47 * round(2 * u[(k + 5) % 8]), 2.42, 3.5, 7.5])
    if d["pocket"]:
        x = np.concatenate([x[:8], [half + 1.5, half + 2.5, half + 4.5, half + 5.5], x[8:]])
        z = np.concatenate([z[:8], [3.27, 1.93, 1.93, 3.37], z[8:]])
    s = ns.IrregularSection(x=x, z=z + invert, n=d["n_main"], bed_slope=d["slope"])
    s.set_roughness_para((d["n_fp"], d["n_main"], d["n_fp"] * 1.1, -half, half))
    return s


def _trapezoid(ns, d, invert, k, kind):
    u = d["shape"]
    b = 14.0 + 10.0 * u[k % 8]
    if kind == "rect":
        return ns.TrapezoidalSection(z_bed=invert, b_main=b + 10.0, m_main=0.0, n_main=d["n_main"], bed_slope=d["slope"])
    if kind == "simple":
        return ns.TrapezoidalSection(z_bed=invert, b_main=b, m_main=1.0 + u[(k + 1) % 8], n_main=d["n_main"], bed_slope=d["slope"])
    return ns.TrapezoidalSection(z_bed=invert, b_main=b, m_main=1.0 + u[(k + 1) % 8], n_main=d["n_main"],
                                 z_bank=invert + 1.6 + u[(k + 2) % 8], b_fp_left=5.0 + 10 * u[(k + 3) % 8],
                                 b_fp_right=5.0 + 10 * u[(k + 4) % 8], m_fp=2.0 + u[(k + 5) % 8],
                                 n_left=d["n_fp"], n_right=d["n_fp"] * 1.1, bed_slope=d["slope"])


def random_case(ns, seed):
    """-> (solver, run kwargs, description).  Raises whatever the implementation raises for the configuration."""
    d = describe(seed)
    L = d["n_cells"] * d["dx"]
    S0, dt = d["slope"], d["dt"]
    rise = 3 * dt

    def wave(t):
        return d["q_base"] + (d["q_peak"] - d["q_base"]) * sin(pi * min(t, 2 * rise) / (2 * rise)) ** 2

    def stage_up(t):
        return S0 * L + d["depth0"] + 0.6 * sin(pi * min(t, 2 * rise) / (2 * rise)) ** 2

    def stage_down(t):
        return d["depth0"] + 0.4 * sin(pi * min(t, 2 * rise) / (2 * rise)) ** 2

    if d["up"] == "flow_hydrograph":
        up = ns.Boundary("flow_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=d["depth0"], hydrograph=ns.Hydrograph(function=wave))
    elif d["up"] in ("fixed_depth", "normal_depth"):
        up = ns.Boundary(d["up"], chainage=0, bed_level=S0 * L, initial_depth=d["depth0"])
    else:
        up = ns.Boundary("stage_hydrograph", chainage=0, bed_level=S0 * L, initial_depth=d["depth0"], hydrograph=ns.Hydrograph(function=stage_up))
    if d["down"] == "rating_curve":
        rc = ns.RatingCurve()
        if d["rating"] == "power":
            rc.set("power", a=d["q_base"] / d["depth0"] ** 1.6, b=1.6)
        else:
            rc.set("polynomial", a=0.2 * d["q_base"] / d["depth0"] ** 2, b=0.8 * d["q_base"] / d["depth0"], c=0.0)
        down = ns.Boundary("rating_curve", chainage=L, bed_level=0.0, initial_depth=d["depth0"], rating_curve=rc)
    elif d["down"] == "stage_hydrograph":
        down = ns.Boundary("stage_hydrograph", chainage=L, bed_level=0.0, initial_depth=d["depth0"], hydrograph=ns.Hydrograph(function=stage_down))
    elif d["down"] == "storage_general":
        down = ns.Boundary("fixed_depth", chainage=L, bed_level=0.0, initial_depth=d["depth0"])
        ls = ns.LumpedStorage(surface_area=4.0e5, min_stage=d["depth0"], solution_boundaries=(0, 100))
        stages = np.arange(0.0, 42.0, 2.0)
        ls.set_area_curve(np.column_stack([stages, 4.0e5 * (1.0 + d["storage"]["slope"] * stages)]), alpha=1.0, beta=0.0)
        rc = ns.RatingCurve()
        rc.set("polynomial", a=0.1 * d["q_base"] / d["depth0"] ** 2, b=0.6 * d["q_base"] / d["depth0"], c=0.0)
        ls.rating_curve, ls.capture_losses, ls.reservoir_length, ls.K_q = rc, d["storage"]["losses"], 1500.0, d["storage"]["k_q"]
        down.set_lumped_storage(ls)
    elif d["down"] == "storage":
        down = ns.Boundary("fixed_depth", chainage=L, bed_level=0.0, initial_depth=d["depth0"])
        down.set_lumped_storage(ns.LumpedStorage(surface_area=4.0e5, min_stage=d["depth0"], solution_boundaries=(0, 100)))
    else:
        down = ns.Boundary(d["down"], chainage=L, bed_level=0.0, initial_depth=d["depth0"])
    ch = ns.Channel(upstream_boundary=up, downstream_boundary=down, initial_flow=d["q_base"], roughness=d["n_main"], width=40.0,
                    interpolation_method=d["ic"])
    fam = d["family"]
    if fam != "width":
        stations = [0.0, L] if d["n_cells"] < 12 else [0.0, float(d["dx"] * (d["n_cells"] // 2)), L]
        kinds = {"simple": ["simple"] * 3, "rect_sections": ["rect"] * 3, "compound": ["compound"] * 3, "polyline": ["poly"] * 3,
                 "trapezoid_to_polyline": ["compound", "poly", "poly"], "polyline_to_trapezoid": ["poly", "poly", "compound"],
                 "compound_to_simple": ["compound", "compound", "simple"]}[fam]
        if len(stations) == 2:
            kinds = [kinds[0], kinds[-1]]
        secs = [(_polyline(ns, d, S0 * (L - c), k) if kind == "poly" else _trapezoid(ns, d, S0 * (L - c), k, kind))
                for k, (c, kind) in enumerate(zip(stations, kinds))]
        if d["curved"]:     # an S-bend: curvature is taken at the interior input sections (channel.py:243-277)
            sx = np.linspace(0.0, L, 25)
            ch.set_coords(coords=np.column_stack([sx, 0.05 * L * np.sin(2 * np.pi * sx / L)]), chainages=sx * 1.0)
        ch.set_cross_sections(stations, secs)
    solver = ns.PreissmannSolver(channel=ch, theta=d["theta"], time_step=int(dt), spatial_step=int(d["dx"]),
                                 simulation_time=d["levels"] * int(dt))       # the reference wants integers here
    return solver, dict(tolerance=d["tol"], max_iter=60), d


def flat_digest(flat):
    """sha1 over every array / scalar of the flattened inputs (what the device and the oracle consume)."""
    h = hashlib.sha1()

    def put(v):
        if v is None:
            h.update(b"none")
        elif isinstance(v, dict):
            for k in sorted(v):
                h.update(k.encode()); put(v[k])
        elif isinstance(v, (str, bytes)):
            h.update(v if isinstance(v, bytes) else v.encode())
        else:
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
            h.update(str(a.shape).encode()); h.update(a.tobytes())

    put(flat.geom); put(flat.ic_depth); put(flat.ic_flow)
    put([flat.n_nodes, flat.n_levels, flat.dt, flat.dx, flat.theta, flat.g, flat.tol, flat.max_iter])
    for bc in (flat.up, flat.down):
        put(int(bc.type)); put(bc.series); put(bc.rating)
        for f in ("bed_level", "bed_slope", "fixed_depth", "storage_area", "storage_min_stage", "storage_ymin", "storage_ymax"):
            put(getattr(bc, f))
    return h.hexdigest()


# ---- the headline reach (cases/gerd_roseires): random members and scenarios -------------------------------------------

def describe_gerd(seed):
    """Roughness beyond the calibration grid, floodplain overrides, pool levels, jammed gates, blend widths, gate control
    and (every fourth) the curved centre line: keyword arguments for `build_gerd` on either implementation."""
    rng = np.random.default_rng(50_000 + seed)
    kw = dict(n_main=float(rng.uniform(0.018, 0.065)), calibration=True)
    if rng.uniform() < 0.5:
        kw["n_fp"] = float(rng.uniform(0.03, 0.12))
    kw["initial_roseires_level"] = float(np.round(rng.uniform(484.5, 489.0), 2))
    rk = dict(jammed_spillways=int(rng.integers(0, 3)), jammed_sluice_gates=int(rng.integers(0, 3)),
              buffer=float(np.round(rng.uniform(0.2, 1.0), 2)))
    if rng.uniform() < 0.3:
        rk.update(smooth=False, initially_open=bool(rng.uniform() < 0.5), max_cooldown=int(rng.choice([3600, 7200, 14400])))
    kw["rating_kwargs"] = rk
    if seed % 4 == 3:           # config 3's curved centre line, 12 hourly levels
        kw.update(calibration=False, sim_duration=12 * 3600)
    return kw
