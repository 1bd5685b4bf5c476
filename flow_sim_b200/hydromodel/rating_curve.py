"""``RatingCurve`` (rating_curve.py:3-162): polynomial / power / fitted stage-discharge relations.

``discharge`` and ``dQ_dz`` exist here for host-side use (setup, post-processing, ``stage``); inside the
time loop the same forms are evaluated by ``rating_q`` / ``rating_dq`` in csrc/pr_device.cuh from the
parameters that ``flow_sim_b200.flatten.flatten_rating`` extracts from this object.
"""
from __future__ import annotations

import numpy as np


class RatingCurve:
    def __init__(self):
        self.function = None
        self.derivative = None
        self.defined = False
        self.type = None
        self.stage_shift = 0

    def set(self, type, a, b, c=None, stage_shift=None):
        self.stage_shift = 0 if stage_shift is None else stage_shift
        if type == "polynomial":
            if c is None:
                raise ValueError("Insufficient arguments. c must be specified.")
            self.a, self.b, self.c = a, b, c
        elif type == "power":
            self.a, self.b = a, b
        else:
            raise ValueError("Invalid type.")
        self.function = self.derivative = None
        self.defined, self.type = True, type

    def _require(self):
        if not self.defined:
            raise ValueError("Rating curve is undefined.")

    def discharge(self, stage, time=None):
        self._require()
        if self.function is not None:
            return self.function(stage)
        x = stage + self.stage_shift
        return self.a * x ** 2 + self.b * x + self.c if self.type == "polynomial" else self.a * x ** self.b

    def dQ_dz(self, stage, time=None):
        self._require()
        y = stage + self.stage_shift
        if self.type == "polynomial":
            return self.derivative(y) if self.function is not None else self.a * 2 * y + self.b
        return self.a * self.b * y ** (self.b - 1)

    def stage(self, discharge, trial_stage=None, time=None, tolerance=1e-2, rate=1):
        self._require()
        if trial_stage is None:
            trial_stage = -self.stage_shift * 1.05
        q = self.discharge(stage=trial_stage, time=time)
        while abs(q - discharge) > tolerance:
            trial_stage += -rate * (q - discharge) / self.dQ_dz(stage=trial_stage, time=time)
            q = self.discharge(stage=trial_stage, time=time)
        return trial_stage

    def fit(self, discharges, stages, stage_shift=0, type="polynomial", scale=True, degree=2):
        self.type = type
        q = np.asarray(discharges, dtype=np.float64)
        y = np.asarray(stages, dtype=np.float64)
        if q.size < 3:
            raise ValueError("Need at least 3 points.")
        if q.shape != y.shape:
            raise ValueError("Q and Y lists should have the same lengths.")
        self.stage_shift = stage_shift
        ys = y + stage_shift
        if any(ys <= 0):
            raise ValueError("All (stage - base) values must be positive for power-law fitting.")
        if type == "polynomial":
            if scale:
                self.function = np.polynomial.polynomial.Polynomial.fit(x=ys, y=q, deg=degree)
                self.derivative = self.function.deriv()
            else:
                if degree != 2:
                    print("WARNING: Polynomial degree defaults to 2 for unscaled fitting.")
                a, b, c = np.polyfit(ys, q, deg=2)
                self.a, self.b, self.c = float(a), float(b), float(c)
        elif type == "power":
            b, log_a = np.polyfit(np.log(ys), np.log(q), deg=1)
            self.a, self.b = float(np.exp(log_a)), float(b)
        else:
            raise ValueError("Invalid rating curve type.")
        self.defined = True

    def tostring(self):
        self._require()
        s = f"(Y+{self.stage_shift})"
        if self.type == "polynomial":
            return str(self.function) if self.function is not None else f"{self.a} {s}^2 + {self.b} {s} + {self.c}"
        return f"{self.a} {s}^{self.b}"
