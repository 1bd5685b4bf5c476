"""``Boundary`` (boundary.py:7-247): configuration of one end of the reach.

The residual / derivative rows (``condition_residual``, ``df_dh``, ``df_dQ`` in the reference) are evaluated
on the device by ``bc_eval`` (csrc/pr_device.cuh); this object only carries what ``flatten`` needs.
"""
from __future__ import annotations

CONDITIONS = ["flow_hydrograph", "fixed_depth", "normal_depth", "rating_curve", "stage_hydrograph"]


class Boundary:
    def __init__(self, condition, chainage, bed_level=None, initial_depth=None, rating_curve=None, hydrograph=None):
        if condition not in CONDITIONS:
            raise ValueError("Invalid boundary condition.")
        self.condition = condition
        self.cross_section = None
        self.bed_level = bed_level
        if initial_depth is None:
            self.initial_depth = self.initial_stage = None
        else:
            self.initial_depth = initial_depth
            self.initial_stage = bed_level + initial_depth
        self.chainage = chainage
        self.rating_curve = rating_curve
        self.hydrograph = hydrograph
        self.lumped_storage = None

    def set_lumped_storage(self, lumped_storage):
        self.lumped_storage = lumped_storage

    def condition_type(self) -> bool:
        """True if the boundary equation is written on Q (boundary.py:244-247)."""
        return self.condition in ("flow_hydrograph", "normal_depth", "rating_curve")
