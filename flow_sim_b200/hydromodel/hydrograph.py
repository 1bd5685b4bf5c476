"""``Hydrograph``: a time series given either as a callable f(t) or as a two-column (time [s], value) table
(same constructor keywords and methods as the reference's class, hydrograph.py:3-33).

The solver never interpolates inside the time loop: ``sample`` evaluates the hydrograph once at the grid times
t = k*dt, which is all the reference ever asks of it (preissmann.py:215,313), and the resulting vector is what
crosses the C ABI (``pr_bc.series``).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


class Hydrograph:
    def __init__(self, function: Optional[Callable[[float], float]] = None, table: Optional[np.ndarray] = None):
        self.table = table
        self._fn = function

    # the reference exposes the active evaluator as an attribute; keep the name readable and writable
    @property
    def used_function(self) -> Callable[[float], float]:
        return self._fn if self._fn is not None else self.interpolate_hydrograph

    @used_function.setter
    def used_function(self, fn) -> None:
        self._fn = fn

    def interpolate_hydrograph(self, time: float) -> float:
        """Piecewise-linear look-up in the table, constant beyond its ends (numpy.interp)."""
        if self.table is None:
            raise ValueError("Hydrograph is not defined.")
        t, v = np.asarray(self.table)[:, 0], np.asarray(self.table)[:, 1]
        return float(np.interp(time, t, v))

    def get_at(self, time: float) -> float:
        return self.used_function(time)

    __call__ = get_at

    def set_table(self, table: np.ndarray) -> None:
        self.table = table

    def set_function(self, func: Callable[[float], float]) -> None:
        self._fn = func

    def sample(self, n_levels: int, dt) -> np.ndarray:
        """Values on the time grid 0, dt, 2 dt, ... ((n_levels-1) dt)."""
        return np.fromiter((float(self.get_at(k * dt)) for k in range(n_levels)), dtype=np.float64, count=n_levels)
