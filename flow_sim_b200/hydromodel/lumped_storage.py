"""``LumpedStorage`` (lumped_storage.py:7-179): 0-D reservoir behind the downstream node.

The device path implements the constant-surface-area form (the one the shipped example uses,
cases/example/main.py:42-43) in closed form, and the general form - tabulated area curve, outflow rating curve,
head losses - with Brent's method on the device.
"""
from __future__ import annotations

import numpy as np


class LumpedStorage:
    def __init__(self, solution_boundaries, surface_area=None, min_stage=None, rating_curve=None):
        self.rating_curve = rating_curve
        self.surface_area = surface_area
        self.min_stage = min_stage
        self.stage_hydrograph = []
        self.area_curve = None
        self.reservoir_length = None
        self.capture_losses = False
        self.Cc = 0.5
        self.K_q = 0
        if solution_boundaries is not None:
            self.Y_min, self.Y_max = solution_boundaries[0], solution_boundaries[1]

    def set_area_curve(self, table, alpha=1, beta=0, update_solution_boundaries=True):
        self.alpha, self.beta = alpha, beta
        self.area_curve = np.asarray(table, dtype=np.float64)
        self.area_gradient = np.gradient(self.area_curve[:, 1], self.area_curve[:, 0])
        if update_solution_boundaries:
            self.Y_min = np.min(self.area_curve[:, 0])
            self.Y_max = np.max(self.area_curve[:, 0])

    def area_at(self, stage):
        if self.area_curve is None:
            return self.surface_area
        return self.alpha * np.interp(stage + self.beta, self.area_curve[:, 0], self.area_curve[:, 1])

    def net_vol_change(self, Y1, Y2):
        if self.area_curve is None:
            return (Y2 - Y1) * self.surface_area
        step = np.min(np.abs(np.diff(self.area_curve[:, 0])))
        n = int(abs(Y2 - Y1) / step)
        if n > 2:
            ys = np.linspace(Y1, Y2, n)
            return np.trapezoid([self.area_at(y) for y in ys], ys)
        return 0.5 * (self.area_at(Y2) + self.area_at(Y1)) * (Y2 - Y1)

    def energy_loss(self, entry_area, flow, roughness, hydraulic_radius, A_str=None):
        """Head loss between the last node and the reservoir (lumped_storage.py:47-73): Manning friction over
        reservoir_length + empirical K_q V^2/2g (+ a sudden-expansion term when A_str is given)."""
        if not self.capture_losses:
            return 0
        from . import hydraulics

        V = flow / entry_area
        hf = hydraulics.Sf(A=entry_area, Q=flow, n=roughness, R=hydraulic_radius) * self.reservoir_length
        h_exp = 0 if A_str is None else (1 - entry_area / A_str) ** 2 * V ** 2 / (2 * hydraulics.g)
        return hf + h_exp + self.K_q * V ** 2 / (2 * hydraulics.g)
