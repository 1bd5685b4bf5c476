"""``Solver`` (solver.py:10-329): grid sizing, result arrays and derived results.

The accessors the reference's assembly code calls per scalar (``area_at``, ``Se_at`` ...) are kept for API
compatibility; the derived result arrays are produced with whole-array numpy operations after the run.
"""
from __future__ import annotations

import os

import numpy as np

from .hydraulics import g
from .utility import seconds_to_hms


class Solver:
    def __init__(self, channel, time_step, spatial_step, simulation_time, regularization=False, fit_spatial_step=True):
        self.channel = channel
        self.time_step, self.spatial_step = time_step, spatial_step
        self.time_level = 0
        self.number_of_nodes = self.channel.length // self.spatial_step + 1
        self.number_of_time_levels = simulation_time // self.time_step + 1
        if fit_spatial_step:
            self.fit_spatial_step()
        self.number_of_nodes = int(self.number_of_nodes)
        self.number_of_time_levels = int(self.number_of_time_levels)
        self.channel.initialize_conditions(n_nodes=self.number_of_nodes)
        self.num_celerity = self.spatial_step / self.time_step
        self.flow = np.empty((self.number_of_time_levels, self.number_of_nodes), dtype=np.float64)
        self.depth = np.empty_like(self.flow)
        self._type = None
        self._solved = False
        self.total_sim_duration = 0
        self.regularization = regularization
        self.eps = 1e-4

    def fit_spatial_step(self):
        self.number_of_nodes = round(self.channel.length / self.spatial_step) + 1
        self.spatial_step = self.channel.length / (self.number_of_nodes - 1)

    def initialize_t0(self):
        self.depth[0, :] = self.channel.initial_conditions[:, 0]
        self.flow[0, :] = self.channel.initial_conditions[:, 1]

    # ---- state accessors (solver.py:244-296) ---------------------------------------------------
    def _level(self, k):
        return self.time_level if k is None else self.time_level - 1 if k == -1 else k

    def depth_at(self, k=None, i=None, regularization=None):
        if i is None:
            raise ValueError("Spatial node must be specified.")
        return self.depth[self._level(k), i]

    def flow_at(self, k=None, i=None, chi_scaling=None):
        if i is None:
            raise ValueError("Spatial node must be specified.")
        return self.flow[self._level(k), i]

    def water_level_at(self, k=None, i=None, regularization=None):
        return self.channel.bed_level_at(i=i) + self.depth_at(k=k, i=i)

    def area_at(self, k=None, i=None, regularization=None):
        if i is None:
            raise ValueError("Spatial node must be specified.")
        return self.channel.area_at(i=i, hw=self.water_level_at(k=k, i=i))

    def Se_at(self, k=None, i=None, regularization=None, chi_scaling=None):
        return self.channel.Se(h=self.depth_at(k=k, i=i), Q=self.flow_at(k=k, i=i), i=i)

    def dA_dh(self, k=None, i=None, regularization=None):
        return self.channel.dA_dh(i=i, hw=self.water_level_at(k=k, i=i))

    # ---- derived results (solver.py:65-127) ------------------------------------------------------
    def prepare_results(self):
        if self.time_level + 1 < self.number_of_time_levels:
            self.flow = self.flow[: self.time_level + 1, :]
            self.depth = self.depth[: self.time_level + 1, :]
        xs = self.channel.xs_at_node
        self.bed_profile = np.array([s.z_min for s in xs], dtype=np.float64)
        self.level = self.depth + self.bed_profile
        self.area = np.empty_like(self.flow)
        self.top_width = np.empty_like(self.flow)
        for i, s in enumerate(xs):
            props = [s.properties(hw) for hw in self.level[:, i]]
            self.area[:, i] = [p[0] for p in props]
            self.top_width[:, i] = [p[3] for p in props]
        V = self.flow / np.maximum(self.area, 1e-6)                           # hydraulics.froude_num clamps
        D = self.area / np.maximum(self.top_width, 1e-6)
        self.froude_number = V / np.sqrt(g * np.maximum(D, 1e-6))
        self.velocity = self.flow / self.area
        self.wave_celerity = self.velocity + np.sqrt(g * self.area / self.top_width)
        self.amplitude = self.depth - self.depth[0, :]
        self.peak_amplitude = self.amplitude.max(axis=0)

        storage = self.channel.downstream_boundary.lumped_storage
        if storage is not None and getattr(self, "_storage_stage", None) is not None:
            self.storage_stage = np.array(self._storage_stage[: self.time_level + 1], dtype=np.float64)
            storage.stage_hydrograph = [[k * self.time_step, float(v)] for k, v in enumerate(self.storage_stage)]
            out = np.empty(self.time_level + 1, dtype=np.float64)
            q_end = self.flow[:, -1]
            out[0] = 0 if storage.rating_curve is None else min(q_end[0], storage.rating_curve.discharge(
                stage=self.storage_stage[0], time=0))
            for k in range(1, self.time_level + 1):
                avg_in = 0.5 * (q_end[k - 1] + q_end[k])
                dvol = storage.net_vol_change(Y1=self.storage_stage[k - 1], Y2=self.storage_stage[k])
                out[k] = (avg_in - dvol / self.time_step) * q_end[k] / avg_in
            self.storage_outflow = out

    def summary(self) -> str:
        """The text report the reference writes next to its workbook (solver.py:187-233)."""
        q_in, q_out = self.flow[:, 0], self.flow[:, -1]
        imbalance = np.sum(q_in - q_out) * self.time_step
        lines = [f"Spatial step = {self.spatial_step} m", f"Time step = {self.time_step} s"]
        if self._type == "preissmann":
            lines.append(f"Theta = {self.theta}")
        lines += [f"Simulation duration = {seconds_to_hms(self.total_sim_duration)}",
                  f"Mass imbalance (total inflow - total outflow) = {imbalance:.2f} m^3 = "
                  f"{float(imbalance / self.time_step / np.sum(q_in)) * 100:.4f}% of inflow.",
                  f"Peak inflow = {np.max(q_in):.2f} m^3/s", f"Peak outflow = {np.max(q_out):.2f} m^3/s",
                  f"Attenuation = {(np.max(q_in) - np.max(q_out)) / np.max(q_in) * 100:.2f}%"]

        def median_time(q):
            cum = np.concatenate([[0.0], np.cumsum(q)[:-1]])
            return int(np.argmax(cum >= 0.5 * cum[-1])) * self.time_step

        t_in, t_out = median_time(q_in), median_time(q_out)
        lines += [f"Median volume entry time = {seconds_to_hms(t_in)}",
                  f"Median volume arrival time = {seconds_to_hms(t_out)}",
                  f"Median volume travel time = {seconds_to_hms(t_out - t_in)}"]
        return "\n".join(lines) + "\n"

    def save_results(self, folder_path, file_name=None):
        """Workbook (when pandas + openpyxl are importable) or .npz, plus the text summary."""
        folder_path = folder_path.replace("\\", "/")
        os.makedirs(folder_path, exist_ok=True)
        file_name = "results.xlsx" if file_name is None else file_name
        path = os.path.join(folder_path, file_name)
        sheets = {"Level": self.level, "Flow": self.flow, "Depth": self.depth, "Velocity": self.velocity,
                  "Area": self.area, "Top width": self.top_width, "Wave celerity": self.wave_celerity,
                  "Amplitude": self.amplitude, "Froude number": self.froude_number}
        time = np.arange(self.flow.shape[0]) * self.time_step
        dist = np.asarray(self.channel.ch_at_node, dtype=np.float64)
        try:
            import openpyxl  # noqa: F401
            import pandas as pd

            with pd.ExcelWriter(path, engine="openpyxl") as w:
                for name, arr in sheets.items():
                    pd.DataFrame(arr, index=time, columns=dist).to_excel(w, sheet_name=name)
        except ImportError:
            np.savez_compressed(os.path.splitext(path)[0] + ".npz", time=time, distance=dist,
                                **{k.replace(" ", "_"): v for k, v in sheets.items()})
        with open(os.path.splitext(path)[0] + ".txt", "w") as f:
            f.write(self.summary())

    def _finalize(self, verbose):
        self._solved = True
        self.total_sim_duration = self.time_level * self.time_step
        self.prepare_results()
        if verbose >= 1:
            print("Simulation completed successfully.")
