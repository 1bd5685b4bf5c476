"""cases/gerd_roseires (BASELINE configs 3 and 4): 120 km compound-trapezoid reach from GERD to Roseires.

Host-side setup of the case on the mirror API, from the packed input data (``data/gerd_roseires.json``, made by
tools/make_gerd_bundle.py from the reference's CSV files):

* ``GerdHydrograph``      upstream discharge = inflow routed through the GERD reservoir
                          (reference: cases/gerd_roseires/gerd_discharge.py:6-124)
* ``RoseiresRatingCurve`` downstream stage-discharge relation of the Roseires gates: two degree-2 bivariate
                          least-squares fits + gate states + smooth blending
                          (reference: cases/gerd_roseires/roseires_rating_curve.py:18-257)
* ``build``               the model set-up of cases/gerd_roseires/model.py:10-92 and n_calibrate.py:5-17

None of this is on the hot path: it produces the per-step inflow table and the rating-curve parameters the
device consumes.
"""
from __future__ import annotations

import json
import os

import numpy as np
from scipy.optimize import brentq

from ..hydromodel import Boundary, Channel, Hydrograph, PreissmannSolver, RatingCurve, TrapezoidalSection

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "gerd_roseires.json")
_bundle = None

# cases/gerd_roseires/settings.py
SPATIAL_STEP, TIME_STEP, THETA = 1000, 3600, 0.6
SIM_DURATION, TOLERANCE = 3600 * 384, 1e-6
INITIAL_ROSEIRES_LEVEL, INITIAL_GERD_LEVEL = 487.0, 637.0
# cases/gerd_roseires/n_calibrate.py:28-30
CALIB_Q = np.array([1562.5, 3850, 6000, 10000, 14000, 21000], dtype=np.float64)
CALIB_H_TARGET = np.array([497.5, 500, 502, 505, 507, 510], dtype=np.float64)


def bundle() -> dict:
    global _bundle
    if _bundle is None:
        with open(_DATA) as f:
            _bundle = json.load(f)
    return _bundle


# ------------------------------------------------------------------------------------------------

class GerdHydrograph(Hydrograph):
    """Outflow of the GERD reservoir for a given inflow hydrograph (level-pool routing, one brentq per step)."""
    CREST, MAX_OPERATING = 624.9, 640.0

    def __init__(self):
        super().__init__(function=None, table=None)
        self.turbine_flow = 1562.5

    def capacity(self, wl):
        ramp = 0.0 if wl <= self.CREST else 1.0 if wl >= self.MAX_OPERATING else (wl - self.CREST) / (self.MAX_OPERATING - self.CREST)
        gated = 196.4017 * max(0, wl - 624.9) ** (3 / 2) * ramp
        stepped = 447.3594 * max(0, wl - 640.0) ** (3 / 2)
        emergency = 654.6723 * max(0, wl - 642.0) ** (3 / 2)
        return gated + stepped + emergency + 0 + self.turbine_flow

    def release(self, inflow, stage, initial_stage):
        cap = self.capacity(stage)
        if stage > initial_stage:
            return cap
        return max(min(inflow, cap), self.turbine_flow)

    def build(self, inflow_hydrograph, time_step, duration, initial_stage):
        curve = np.asarray(bundle()["gerd_vol_curve"], dtype=np.float64)
        vols, stages = curve[:, 0], curve[:, 1]
        self.table = np.empty((duration // time_step + 1, 2), dtype=np.float64)
        stage0 = initial_stage
        in0 = inflow_hydrograph.get_at(0)
        out0 = self.release(inflow=in0, stage=stage0, initial_stage=initial_stage)
        self.table[0] = (0, out0)
        for t in range(time_step, duration + time_step, time_step):
            in1 = inflow_hydrograph.get_at(t)
            avg_in = 0.5 * (in1 + in0)
            vol0 = np.interp(x=stage0, xp=stages, fp=vols)

            def balance(stage1):
                out1 = self.release(in1, stage1, initial_stage)
                return (np.interp(x=stage1, xp=stages, fp=vols) - vol0) - (avg_in - 0.5 * (out1 + out0)) * time_step * 1e-6

            stage1 = brentq(f=balance, a=624.9, b=645)
            out1 = self.release(in1, stage1, initial_stage)
            self.table[t // time_step] = (t, out1)
            stage0, in0, out0 = stage1, in1, out1


# ------------------------------------------------------------------------------------------------

def fit_quadratic_surface(table: dict) -> np.ndarray:
    """[intercept, s, o, s^2, s*o, o^2] of the least-squares degree-2 surface through a release table -
    what sklearn's Pipeline(PolynomialFeatures(2, include_bias=False), LinearRegression()) computes:
    centre X and y, solve with scipy.linalg.lstsq, recover the intercept."""
    from scipy import linalg

    rows, targets = [], []
    for s, line in zip(table["stage"], table["discharge"]):
        for o, qv in zip(table["column"], line):
            if qv is not None:
                rows.append([s, o, s * s, s * o, o * o])
                targets.append(qv)
    X = np.asarray(rows, dtype=np.float64)
    y = np.asarray(targets, dtype=np.float64)
    x_mean, y_mean = X.mean(axis=0), y.mean()
    Xc = X - x_mean
    coef, _, _, _ = linalg.lstsq(Xc, y - y_mean, cond=max(Xc.shape) * np.finfo(Xc.dtype).eps)
    return np.concatenate([[y_mean - x_mean @ coef], coef])


class RoseiresRatingCurve(RatingCurve):
    HYDROPOWER_Q = 63.0 * 1e6 / (24 * 3600)
    NUM_SLUICE_GATES, NUM_SPILLWAYS, MAX_SPILLWAY_OPENING = 5, 7, 13
    MIN_STAGE, MAX_STAGE = 466.7, 492
    TAIL_WATER_LEVEL_RANGE = (440, 455)

    def __init__(self, initial_stage=None, initial_flow=None, initially_open=False, jammed_spillways=0,
                 jammed_sluice_gates=0, max_cooldown=3600 * 5, smooth=True, buffer=0.5, deep_sluices_active=True,
                 dY=0.001):
        super().__init__()
        self.defined, self.type = True, "roseires"
        b = bundle()
        self.spill_coef = fit_quadratic_surface(b["spillway_releases"])
        self.sluice_coef = fit_quadratic_surface(b["sluice_releases"])
        self.hydropower_q = self.HYDROPOWER_Q
        self.dY = dY
        if initial_stage > self.MAX_STAGE or initial_stage < self.MIN_STAGE:
            raise ValueError(f"Roseires water stage must be between {self.MIN_STAGE} m and {self.MAX_STAGE} m.")
        self.initial_stage = initial_stage
        self.smooth, self.buffer = smooth, buffer
        self.jammed_spillways = jammed_spillways
        self.jammed_sluice_gates = jammed_sluice_gates if deep_sluices_active else self.NUM_SLUICE_GATES
        n_free = self.NUM_SPILLWAYS - self.jammed_spillways
        self.open_state = ([self.MAX_SPILLWAY_OPENING] * n_free + [0] * self.jammed_spillways,
                           self.NUM_SLUICE_GATES - self.jammed_sluice_gates)
        self.tail_water_level = float(np.average(self.TAIL_WATER_LEVEL_RANGE))
        self.closed_state = self._closed_state(initial_flow)
        self.open = bool(initially_open)
        # gate-control state (smooth=False): the device keeps one copy per member, seeded from these
        self.max_cooldown, self.cooldown, self.prev_time, self.current_stage = max_cooldown, 0, None, initial_stage

    @staticmethod
    def _surface(c, s, o):
        return float(np.dot([s, o, s * s, s * o, o * o], c[1:]) + c[0])

    def release(self, stage, state):
        openings, sluices = state
        spill = sum([self._surface(self.spill_coef, stage, o) if o > 0 else 0 for o in openings])
        return spill + self._surface(self.sluice_coef, stage, self.tail_water_level) * sluices + self.hydropower_q

    def _closed_state(self, initial_flow):
        """Gate setting that passes `initial_flow` at the initial stage (roseires_rating_curve.py:143-178)."""
        s0, full, n_free = self.initial_stage, self.MAX_SPILLWAY_OPENING, self.NUM_SPILLWAYS - self.jammed_spillways
        sluices = None
        for i in range(1, self.NUM_SLUICE_GATES + 1 - self.jammed_sluice_gates):
            sluices = i
            if self.release(s0, ([full] * n_free, i)) > initial_flow:
                sluices = i - 1
                break
        fully = 0
        for i in range(1, self.NUM_SPILLWAYS + 1 - self.jammed_spillways):
            if self.release(s0, ([full] * i + [0] * (self.NUM_SPILLWAYS - i), sluices)) > initial_flow:
                fully = i - 1
                break
        gates = lambda part: [full] * fully + [part] + [0] * (self.NUM_SPILLWAYS - fully - 1)
        partial = round(brentq(lambda p: initial_flow - self.release(s0, (gates(p), sluices)), 0, full), 2)
        if fully + (1 if partial > 0 else 0) > n_free:
            raise ValueError("closed gate state needs more spillways than are free")
        return gates(partial), sluices

    def alpha_smooth(self, stage):
        if stage >= self.initial_stage + self.buffer:
            return 1.0
        if stage <= self.initial_stage:
            return 0.0
        s = (stage - self.initial_stage) / self.buffer
        return 3 * s ** 2 - 2 * s ** 3

    def gate_control(self, time):
        """Open above initial_stage + 0.5, close below initial_stage - 1, at most once per max_cooldown seconds;
        judged on the stage seen by the previous call (roseires_rating_curve.py:111-130)."""
        if self.prev_time is not None:
            self.cooldown = max(0, self.cooldown - (time - self.prev_time))
        self.prev_time = time
        if self.cooldown > 0:
            return
        if self.current_stage >= self.initial_stage + 0.5 and not self.open:
            self.cooldown, self.open = self.max_cooldown, True
        elif self.current_stage <= self.initial_stage - 1 and self.open:
            self.cooldown, self.open = self.max_cooldown, False

    def discharge(self, stage, time=None, update_stage=True, update_gate_state=True, smooth=None):
        if not (self.smooth if smooth is None else smooth):
            if update_gate_state:
                self.gate_control(time)
            q = self.release(stage, self.open_state if self.open else self.closed_state)
            if update_stage:
                self.current_stage = stage
            return q
        a = self.alpha_smooth(stage)
        return (1.0 - a) * self.release(stage, self.closed_state) + a * self.release(stage, self.open_state)

    def dQ_dz(self, stage, time=None, dY=None):
        dY = self.dY if dY is None else dY
        frozen = dict(time=time, update_stage=False, update_gate_state=False)
        return (self.discharge(stage + dY, **frozen) - self.discharge(stage - dY, **frozen)) / (2 * dY)


# ------------------------------------------------------------------------------------------------

def load_sections(n_main=None, n_fp=None):
    """custom_functions.load_trapzoid_xs (:128-157): one compound trapezoid per surveyed section; section 53 is skipped."""
    s = bundle()["sections"]
    chain, secs = [], []
    for i, name in enumerate(s["file"]):
        if name == "53.csv":
            continue
        chain.append(s["chainage"][i])
        secs.append(TrapezoidalSection(
            z_bed=s["z_min"][i], b_main=s["b_main"][i], m_main=s["m_main"][i],
            n_main=s["n_main"][i] if n_main is None else n_main, z_bank=s["z_min"][i] + s["h_bankfull"][i],
            b_fp_left=s["b_fp_left"][i], b_fp_right=s["b_fp_right"][i], m_fp=s["m_fp"][i],
            n_left=s["n_left"][i] if n_fp is None else n_fp, n_right=s["n_right"][i] if n_fp is None else n_fp))
    return chain, secs


def inflow_table(small: bool) -> np.ndarray:
    t = np.asarray(bundle()["inflow_hydrograph_small_hours" if small else "inflow_hydrograph_hours"], dtype=np.float64)
    t = t.copy()
    t[:, 0] *= 3600
    return t


def build(n_main=None, n_fp=None, calibration=False, with_gerd=True, curvature=None, sim_duration="default",
          initial_roseires_level=INITIAL_ROSEIRES_LEVEL, gerd_level=INITIAL_GERD_LEVEL, time_step=TIME_STEP,
          spatial_step=SPATIAL_STEP, theta=THETA, tolerance=TOLERANCE, jammed_spillways=0, jammed_sluice_gates=0,
          inflow_hyd_func=None, rating_kwargs=None):
    """model.run() up to the solver construction.  calibration=True reproduces n_calibrate.run_model:
    small inflow hydrograph, duration from the table (32 h), no centre-line curvature."""
    if inflow_hyd_func is not None:
        inflow = Hydrograph(function=inflow_hyd_func)
    else:
        inflow = Hydrograph(table=inflow_table(small=calibration))
    if sim_duration == "default":
        sim_duration = None if calibration else SIM_DURATION
    if sim_duration is None:
        if inflow.table is None:
            raise ValueError("Simulation duration must be specified.")
        duration = int(inflow.table[-1, 0])
    else:
        duration = int(sim_duration)
    use_curvature = (not calibration) if curvature is None else curvature
    gerd = GerdHydrograph()
    gerd.build(inflow_hydrograph=inflow, time_step=time_step, duration=duration, initial_stage=gerd_level)
    q0 = gerd.get_at(time=0)
    chain, sections = load_sections(n_main=n_main, n_fp=n_fp)
    bed = sections[-1].z_min
    up = Boundary(condition="flow_hydrograph", hydrograph=gerd if with_gerd else inflow, chainage=chain[0])
    down = Boundary(initial_depth=initial_roseires_level - bed, bed_level=bed, condition="rating_curve",
                    rating_curve=RoseiresRatingCurve(initial_stage=initial_roseires_level, initial_flow=q0,
                                                     **{**dict(jammed_sluice_gates=jammed_sluice_gates,
                                                               jammed_spillways=jammed_spillways), **(rating_kwargs or {})}),
                    chainage=chain[-1])
    ch = Channel(initial_flow=q0, upstream_boundary=up, downstream_boundary=down)
    if use_curvature:
        c = np.asarray(bundle()["centerline"], dtype=np.float64)
        ch.set_coords(coords=c[:, 1:], chainages=c[:, 0])
    ch.set_cross_sections(chainages=chain, sections=sections)
    solver = PreissmannSolver(channel=ch, theta=theta, time_step=time_step, spatial_step=spatial_step,
                              simulation_time=duration)
    solver.first_section_bed = sections[0].z_min
    return solver, dict(tolerance=tolerance)


def calibration_levels(solver, Q=CALIB_Q):
    """model.py:105-113: stage at the upstream node read off the simulated rating loop at the flows Q."""
    return np.interp(Q, solver.flow[:, 0], solver.depth[:, 0] + solver.first_section_bed)
