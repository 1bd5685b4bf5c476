"""``Hydrograph`` (hydrograph.py:3-33): an arbitrary callable f(t) or a (time, value) table."""
from __future__ import annotations

import numpy as np


class Hydrograph:
    def __init__(self, function=None, table=None):
        self.table = table
        self.used_function = self.interpolate_hydrograph if function is None else function

    def interpolate_hydrograph(self, time):
        if self.table is None:
            raise ValueError("Hydrograph is not defined.")
        return float(np.interp(time, self.table[:, 0], self.table[:, 1]))

    def get_at(self, time):
        return self.used_function(time)

    def set_table(self, table):
        self.table = table

    def set_function(self, func):
        self.used_function = func

    def sample(self, n_levels: int, dt) -> np.ndarray:
        """Values at t = k*dt, k = 0..n_levels-1: all the solver ever asks for (preissmann.py:215,313)."""
        return np.array([float(self.get_at(k * dt)) for k in range(n_levels)], dtype=np.float64)
