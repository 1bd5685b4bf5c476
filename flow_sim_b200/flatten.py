"""Flatten ``hydromodel`` objects (the reference's or this package's mirror) into the SoA arrays the
C ABI consumes (SURVEY.md 8b "one flatten routine").

Works by duck typing on the attribute names of the reference classes:

* ``PreissmannSolver``: ``theta, time_step, spatial_step, number_of_nodes, number_of_time_levels, channel``
  (solver.py:31-44, preissmann.py:41)
* ``Channel``: ``xs_at_node, ch_at_node, xs_chainages, initial_conditions, upstream_boundary,
  downstream_boundary`` (channel.py:30-51)
* ``TrapezoidalSection``: ``z_bed, b_main, m_main, _is_compound, _is_rect, bankfull_depth, T_main_at_bank,
  _width_at_bank, b_fp_left, b_fp_right, m_fp, n_left, n_main, n_right, curvature, bed_slope``
  (cross_section.py:569-613)
* ``Boundary`` / ``RatingCurve`` / ``LumpedStorage`` / ``Hydrograph`` as cited below.

Anything the device path does not implement raises ``NotImplementedError`` - there is no CPU fallback.
"""
from __future__ import annotations

import inspect
import sys
from dataclasses import dataclass, field

import numpy as np

from . import abi

G_STANDARD = 9.80665  # scipy.constants.g (preissmann.py:2)


@dataclass
class FlatBoundary:
    type: int
    bed_level: float = float("nan")
    bed_slope: float = float("nan")
    fixed_depth: float = float("nan")
    series: np.ndarray | None = None          # [levels] or [M, levels]
    rating: dict | None = None
    member_ratings: list | None = None        # release scenarios: one flatten_rating() dict per member
    storage_area: float = 0.0
    storage_min_stage: float = 0.0
    storage_ymin: float = 0.0
    storage_ymax: float = 0.0
    storage_curve: np.ndarray | None = None   # [K, 2] (stage, area) of LumpedStorage.set_area_curve
    storage_alpha: float = 1.0
    storage_beta: float = 0.0
    storage_losses: bool = False
    storage_reservoir_length: float = 0.0
    storage_Kq: float = 0.0
    storage_outflow: dict | None = None       # flatten_rating() of LumpedStorage.rating_curve


@dataclass
class FlatCase:
    n_nodes: int
    n_levels: int
    theta: float
    dt: float
    dx: float
    tol: float
    max_iter: int
    g: float
    geom: dict                                  # name -> [N] array (abi.GEOM_FIELDS)
    up: FlatBoundary
    down: FlatBoundary
    ic_depth: np.ndarray                        # [N] or [M, N]
    ic_flow: np.ndarray
    member_n_main: np.ndarray | None = None     # [M]
    member_n_fp: np.ndarray | None = None       # [M]
    meta: dict = field(default_factory=dict)    # z0 (first section bed), gvf inputs, chainages ...

    @property
    def n_members_hint(self) -> int:
        for a in (self.member_n_main, self.member_n_fp):
            if a is not None:
                return int(np.asarray(a).shape[0])
        if self.up.series is not None and np.ndim(self.up.series) == 2:
            return int(self.up.series.shape[0])
        if np.ndim(self.ic_depth) == 2:
            return int(self.ic_depth.shape[0])
        return 1


# ------------------------------------------------------------------------------------------

def _section_row(xs) -> dict:
    if hasattr(xs, "x") and hasattr(xs, "z") and hasattr(xs, "left_fp_limit") and not hasattr(xs, "b_main"):
        # IrregularSection (cross_section.py:207-543): polyline + composite roughness; the trapezoid columns are unused
        return dict(kind=abi.PR_XS_IRREGULAR, z_bed=float(xs.z_min), b_main=0.0, m_main=0.0, h_bank=0.0, T_bank=0.0,
                    W_bank=0.0, b_fp_l=0.0, b_fp_r=0.0, m_fp=0.0, n_l=float(xs.n_left), n_m=float(xs.n_main),
                    n_r=float(xs.n_right), curvature=float(xs.curvature),
                    _poly=(np.asarray(xs.x, dtype=np.float64), np.asarray(xs.z, dtype=np.float64),
                           float(xs.left_fp_limit), float(xs.right_fp_limit)))
    if not hasattr(xs, "_is_compound") or not hasattr(xs, "b_main"):
        raise NotImplementedError(
            f"{type(xs).__name__}: only TrapezoidalSection (rect / simple / compound) and IrregularSection run on the device")
    compound = bool(xs._is_compound)
    rect = bool(getattr(xs, "_is_rect", (not compound) and xs.m_main == 0.0))
    kind = abi.PR_XS_COMPOUND if compound else (abi.PR_XS_RECT if rect else abi.PR_XS_TRAPEZOID)
    return dict(
        kind=kind, z_bed=float(xs.z_bed), b_main=float(xs.b_main), m_main=float(xs.m_main),
        h_bank=float(xs.bankfull_depth) if compound else 0.0,
        T_bank=float(xs.T_main_at_bank) if compound else 0.0,
        W_bank=float(xs._width_at_bank) if compound else 0.0,
        b_fp_l=float(xs.b_fp_left), b_fp_r=float(xs.b_fp_right), m_fp=float(xs.m_fp),
        n_l=float(xs.n_left), n_m=float(xs.n_main), n_r=float(xs.n_right),
        curvature=float(xs.curvature),
    )


def interpolation_weights(ch_at_node, xs_chainages):
    """(w1, w2) per node as channel.py:218-238 + cross_section.py:874-888 compute them.

    Nodes that coincide with an input section get (1, 0) / (0, 1) so that v*w1 + v*w2 == v exactly,
    matching the reference, which reuses the input section object there.
    """
    ch = np.asarray(ch_at_node, dtype=np.float64)
    xc = np.asarray(xs_chainages, dtype=np.float64)
    w1 = np.ones_like(ch)
    w2 = np.zeros_like(ch)
    for i, s in enumerate(ch):
        if s <= xc[0] or s >= xc[-1]:
            continue
        j = int(np.searchsorted(xc, s)) - 1
        d1 = s - xc[j]
        d2 = xc[j + 1] - s
        tot = d1 + d2
        if tot < 1e-9 or d1 < 1e-9:
            continue
        if d2 < 1e-9:
            w1[i], w2[i] = 0.0, 1.0
            continue
        w1[i] = d2 / tot
        w2[i] = d1 / tot
    return w1, w2


def flatten_geometry(channel) -> dict:
    rows = [_section_row(xs) for xs in channel.xs_at_node]
    geom = {k: np.array([r[k] for r in rows], dtype=np.int32 if k == "kind" else np.float64)
            for k in rows[0] if not k.startswith("_")}
    w1, w2 = interpolation_weights(channel.ch_at_node, channel.xs_chainages)
    geom["w1"], geom["w2"] = w1, w2
    if any("_poly" in r for r in rows):
        empty = (np.empty(0), np.empty(0), 0.0, 0.0)
        polys = [r.get("_poly", empty) for r in rows]
        geom["irr_offset"] = np.concatenate([[0], np.cumsum([len(p[0]) for p in polys])]).astype(np.int32)
        geom["irr_x"] = np.concatenate([p[0] for p in polys])
        geom["irr_z"] = np.concatenate([p[1] for p in polys])
        geom["irr_left"] = np.array([p[2] for p in polys], dtype=np.float64)
        geom["irr_right"] = np.array([p[3] for p in polys], dtype=np.float64)
    return geom


def _linreg_coefficients(pipeline):
    """[intercept, s, o, s^2, s*o, o^2] from sklearn Pipeline(PolynomialFeatures(2, include_bias=False),
    LinearRegression) - roseires_rating_curve.py:229-257."""
    poly = pipeline.named_steps["poly"]
    lin = pipeline.named_steps["linreg"]
    if poly.degree != 2 or poly.include_bias or getattr(poly, "interaction_only", False):
        raise NotImplementedError("Roseires pipeline is not PolynomialFeatures(degree=2, include_bias=False)")
    coef = np.asarray(lin.coef_, dtype=np.float64).ravel()
    if coef.size != 5:
        raise NotImplementedError("expected 5 polynomial features [s, o, s^2, s*o, o^2]")
    return np.concatenate([[float(lin.intercept_)], coef])


def flatten_rating(rc) -> dict | None:
    if rc is None:
        return None
    # Roseires-style gate curve (reference class or this package's mirror)
    if hasattr(rc, "spillway_model") or hasattr(rc, "spill_coef"):
        gated = not getattr(rc, "smooth", True)
        if gated and (getattr(rc, "prev_time", None) is not None or rc.cooldown != 0
                      or rc.current_stage != rc.initial_stage):
            raise NotImplementedError("gate-controlled rating curve that has already been stepped (only a fresh one is supported)")
        if hasattr(rc, "spill_coef"):
            spill, sluice = np.asarray(rc.spill_coef, float), np.asarray(rc.sluice_coef, float)
            q_hydro = float(rc.hydropower_q)
            dY = float(rc.dY)
        else:
            spill, sluice = _linreg_coefficients(rc.spillway_model), _linreg_coefficients(rc.sluice_model)
            q_hydro = float(sys.modules[type(rc).__module__].HYDROPOWER_Q)
            dY = float(inspect.signature(rc.dQ_dz).parameters["dY"].default)
        open_gates, open_sl = rc.open_state
        closed_gates, closed_sl = rc.closed_state
        n_gates = max(len(open_gates), len(closed_gates))
        pad = lambda v: list(map(float, v)) + [0.0] * (n_gates - len(v))
        return dict(type=abi.PR_RC_ROSEIRES, spill=spill, sluice=sluice, twl=float(rc.tail_water_level),
                    open_state=pad(open_gates), closed_state=pad(closed_gates), n_gates=n_gates,
                    sluices_open=int(open_sl), sluices_closed=int(closed_sl),
                    stage0=float(rc.initial_stage), buffer=float(rc.buffer), q_hydro=q_hydro, dY=dY,
                    gate_control=int(gated), initially_open=int(bool(rc.open)) if gated else 0,
                    max_cooldown=float(getattr(rc, "max_cooldown", 0.0)))
    if not getattr(rc, "defined", False):
        raise ValueError("Rating curve is undefined.")
    shift = float(getattr(rc, "stage_shift", 0) or 0)
    if getattr(rc, "function", None) is not None:           # RatingCurve.fit(scale=True): numpy Polynomial
        if rc.type != "polynomial":
            raise NotImplementedError("fitted function with a non-polynomial rating type")
        off, scl = rc.function.mapparms()
        return dict(type=abi.PR_RC_POLYNOMIAL, coef=np.asarray(rc.function.coef, float),
                    dcoef=np.asarray(rc.derivative.coef, float), off=float(off), scl=float(scl), stage_shift=shift)
    if rc.type == "polynomial":
        return dict(type=abi.PR_RC_POLY2, a=float(rc.a), b=float(rc.b), c=float(rc.c), stage_shift=shift)
    if rc.type == "power":
        return dict(type=abi.PR_RC_POWER, a=float(rc.a), b=float(rc.b), stage_shift=shift)
    raise NotImplementedError(f"rating curve type {rc.type!r}")


def flatten_boundary(b, n_levels: int, dt, downstream: bool) -> FlatBoundary:
    if b.condition not in abi.BC_NAMES:
        raise ValueError("Invalid boundary condition.")
    t = abi.BC_NAMES[b.condition]
    fb = FlatBoundary(type=t)
    fb.bed_level = float("nan") if b.bed_level is None else float(b.bed_level)
    slope = getattr(b.cross_section, "bed_slope", None)
    fb.bed_slope = float("nan") if slope is None else float(slope)
    if t in (abi.PR_BC_FLOW_HYDROGRAPH, abi.PR_BC_STAGE_HYDROGRAPH):
        # the reference only evaluates hydrographs at time = time_level * time_step (preissmann.py:215,313)
        fb.series = np.array([float(b.hydrograph.get_at(k * dt)) for k in range(n_levels)], dtype=np.float64)
    if t == abi.PR_BC_NORMAL_DEPTH and slope is None:
        raise ValueError("normal_depth boundary needs a cross-section bed slope")
    if t == abi.PR_BC_RATING_CURVE:
        fb.rating = flatten_rating(b.rating_curve)
    if t == abi.PR_BC_FIXED_DEPTH:
        ls = getattr(b, "lumped_storage", None)
        if ls is None:
            fb.fixed_depth = float(b.initial_depth)
        else:
            if not downstream:
                raise NotImplementedError("lumped storage at the upstream boundary")
            fb.type = abi.PR_BC_FIXED_DEPTH_STORAGE
            fb.storage_area = float(ls.surface_area) if ls.surface_area is not None else 0.0
            fb.storage_min_stage = float(ls.min_stage)
            fb.storage_ymin, fb.storage_ymax = float(ls.Y_min), float(ls.Y_max)
            if ls.area_curve is not None:
                curve = np.asarray(ls.area_curve, dtype=np.float64)
                if curve.ndim != 2 or curve.shape[0] < 2 or not np.all(np.diff(curve[:, 0]) > 0):
                    raise ValueError("area curve must be a [K, 2] table with increasing stages")
                fb.storage_curve = curve[:, :2].copy()
                fb.storage_alpha, fb.storage_beta = float(ls.alpha), float(ls.beta)
            elif ls.surface_area is None:
                raise ValueError("lumped storage needs a surface area or an area curve")
            if ls.rating_curve is not None:
                fb.storage_outflow = flatten_rating(ls.rating_curve)
                if fb.storage_outflow["type"] == abi.PR_RC_ROSEIRES:
                    raise NotImplementedError("gate-blend rating curve as reservoir outflow")
            if ls.capture_losses:
                if ls.reservoir_length is None:
                    raise ValueError("capture_losses needs reservoir_length")
                fb.storage_losses = True
                fb.storage_reservoir_length, fb.storage_Kq = float(ls.reservoir_length), float(ls.K_q)
    return fb


def flatten_solver(solver, tolerance: float = 1e-4, max_iter: int = 100) -> FlatCase:
    """Reduce a constructed (not yet run) PreissmannSolver to a :class:`FlatCase`."""
    if getattr(solver, "regularization", False):
        raise NotImplementedError("regularization=True is broken in the reference (solver.py:283 vs :298) "
                                  "and is out of scope")
    ch = solver.channel
    N = int(solver.number_of_nodes)
    L = int(solver.number_of_time_levels)
    dt = solver.time_step
    ic = np.asarray(ch.initial_conditions, dtype=np.float64)
    flat = FlatCase(
        n_nodes=N, n_levels=L, theta=float(solver.theta), dt=float(dt), dx=float(solver.spatial_step),
        tol=float(tolerance), max_iter=int(max_iter), g=G_STANDARD,
        geom=flatten_geometry(ch),
        up=flatten_boundary(ch.upstream_boundary, L, dt, downstream=False),
        down=flatten_boundary(ch.downstream_boundary, L, dt, downstream=True),
        ic_depth=ic[:, 0].copy(), ic_flow=ic[:, 1].copy(),
    )
    if flat.up.type == abi.PR_BC_FIXED_DEPTH_STORAGE:
        raise NotImplementedError("lumped storage at the upstream boundary")
    flat.meta["z0"] = float(ch.xs_at_node[0].z_min)
    flat.meta["chainage"] = np.asarray(ch.ch_at_node, dtype=np.float64)
    flat.meta["initial_flow"] = float(ch.initial_flow_rate)
    ds = ch.downstream_boundary
    if getattr(ds, "initial_depth", None) is not None:
        flat.meta["downstream_depth"] = float(ds.initial_depth)
    flat.meta["ic_method"] = str(getattr(ch, "interpolation_method", ""))
    flat.meta["bed_slope"] = np.array([float("nan") if getattr(xs, "bed_slope", None) is None else float(xs.bed_slope)
                                       for xs in ch.xs_at_node], dtype=np.float64)
    return flat


# ------------------------------------------------------------------------------------------
# (de)serialisation of a FlatCase to a flat npz (used for the committed golden inputs)
# ------------------------------------------------------------------------------------------

def _bc_to_npz(prefix: str, b: FlatBoundary, out: dict) -> None:
    out[f"{prefix}_scalars"] = np.array([b.type, b.bed_level, b.bed_slope, b.fixed_depth, b.storage_area,
                                         b.storage_min_stage, b.storage_ymin, b.storage_ymax], dtype=np.float64)
    out[f"{prefix}_storagex"] = np.array([b.storage_alpha, b.storage_beta, float(b.storage_losses),
                                          b.storage_reservoir_length, b.storage_Kq], dtype=np.float64)
    if b.storage_curve is not None:
        out[f"{prefix}_storage_curve"] = np.asarray(b.storage_curve, dtype=np.float64)
    if b.storage_outflow:
        for k, v in b.storage_outflow.items():
            out[f"{prefix}_outflow_{k}"] = np.asarray(v, dtype=np.float64)
    if b.series is not None:
        out[f"{prefix}_series"] = np.asarray(b.series, dtype=np.float64)
    if b.rating:
        for k, v in b.rating.items():
            out[f"{prefix}_rating_{k}"] = np.asarray(v, dtype=np.float64)


def _bc_from_npz(prefix: str, z) -> FlatBoundary:
    s = z[f"{prefix}_scalars"]
    b = FlatBoundary(type=int(s[0]), bed_level=float(s[1]), bed_slope=float(s[2]), fixed_depth=float(s[3]),
                     storage_area=float(s[4]), storage_min_stage=float(s[5]), storage_ymin=float(s[6]),
                     storage_ymax=float(s[7]))
    if f"{prefix}_series" in z:
        b.series = np.array(z[f"{prefix}_series"])
    if f"{prefix}_storagex" in z:
        x = z[f"{prefix}_storagex"]
        b.storage_alpha, b.storage_beta, b.storage_losses = float(x[0]), float(x[1]), bool(x[2])
        b.storage_reservoir_length, b.storage_Kq = float(x[3]), float(x[4])
    if f"{prefix}_storage_curve" in z:
        b.storage_curve = np.array(z[f"{prefix}_storage_curve"])

    def rating_from(tag):
        keys = [k for k in z.files if k.startswith(f"{prefix}_{tag}_")]
        if not keys:
            return None
        d = {}
        for k in keys:
            v = np.array(z[k])
            name = k[len(prefix) + len(tag) + 2:]
            d[name] = v if v.ndim else (int(v) if name in ("type", "n_gates", "sluices_open", "sluices_closed", "gate_control", "initially_open") else float(v))
        return d

    b.rating = rating_from("rating")
    b.storage_outflow = rating_from("outflow")
    return b


def save_flat(path: str, flat: FlatCase) -> None:
    out = {"scalars": np.array([flat.n_nodes, flat.n_levels, flat.theta, flat.dt, flat.dx, flat.tol,
                                flat.max_iter, flat.g], dtype=np.float64)}
    for k, v in flat.geom.items():
        out[f"geom_{k}"] = v
    _bc_to_npz("up", flat.up, out)
    _bc_to_npz("down", flat.down, out)
    out["ic_depth"], out["ic_flow"] = flat.ic_depth, flat.ic_flow
    if flat.member_n_main is not None:
        out["member_n_main"] = flat.member_n_main
    if flat.member_n_fp is not None:
        out["member_n_fp"] = flat.member_n_fp
    for k, v in flat.meta.items():
        if isinstance(v, str):
            out[f"metas_{k}"] = np.array(v)
        else:
            out[f"meta_{k}"] = np.asarray(v)
    np.savez_compressed(path, **out)


def load_flat(path: str) -> FlatCase:
    z = np.load(path, allow_pickle=False)
    s = z["scalars"]
    geom = {k[5:]: np.array(z[k]) for k in z.files if k.startswith("geom_")}
    geom["kind"] = geom["kind"].astype(np.int32)
    if "irr_offset" in geom:
        geom["irr_offset"] = geom["irr_offset"].astype(np.int32)
    flat = FlatCase(n_nodes=int(s[0]), n_levels=int(s[1]), theta=float(s[2]), dt=float(s[3]), dx=float(s[4]),
                    tol=float(s[5]), max_iter=int(s[6]), g=float(s[7]), geom=geom,
                    up=_bc_from_npz("up", z), down=_bc_from_npz("down", z),
                    ic_depth=np.array(z["ic_depth"]), ic_flow=np.array(z["ic_flow"]))
    if "member_n_main" in z:
        flat.member_n_main = np.array(z["member_n_main"])
    if "member_n_fp" in z:
        flat.member_n_fp = np.array(z["member_n_fp"])
    for k in z.files:
        if k.startswith("meta_"):
            v = np.array(z[k])
            flat.meta[k[5:]] = v if v.ndim else float(v)
        elif k.startswith("metas_"):
            flat.meta[k[6:]] = str(z[k])
    return flat
