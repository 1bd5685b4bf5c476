"""``Channel`` (channel.py:7-390): boundaries + per-node cross-sections + initial conditions.

Everything here runs once per run on the host (the reference's setup layer).  For
ensembles whose initial profile depends on the member (roughness sweeps) the GVF profile is computed on
the device instead - ``pr_gvf_initial_conditions`` / ``EnsembleRunner.roughness_sweep``.
"""
from __future__ import annotations

import numpy as np

from . import hydraulics
from .cross_section import TrapezoidalSection, interpolate_cross_section


class Channel:
    def __init__(self, upstream_boundary, downstream_boundary, initial_flow, roughness=None, width=None,
                 interpolation_method="GVF_equation"):
        if interpolation_method not in ("linear", "GVF_equation", "steady-state"):
            raise ValueError("Invalid interpolation method.")
        self.initial_conditions = None
        self.conditions_initialized = False
        self.initial_flow_rate = initial_flow
        self.roughness, self.width = roughness, width
        self.length = downstream_boundary.chainage - upstream_boundary.chainage
        self.upstream_boundary, self.downstream_boundary = upstream_boundary, downstream_boundary
        self.interpolation_method = interpolation_method
        self.xs_chainages = self.input_xs = self.ch_at_node = self.xs_at_node = None
        self.coords_chainages = self.coords = None

    # ---- configuration -----------------------------------------------------------------------
    def set_coords(self, coords, chainages):
        self.coords_chainages = np.asarray(chainages, dtype=np.float64)
        self.coords = np.asarray(coords, dtype=np.float64)
        self.coordinated = True

    def set_cross_sections(self, chainages, sections):
        chainages = np.asarray(chainages, dtype=float)
        if len(chainages) != len(sections):
            raise ValueError("chainages and sections must have same length")
        if not np.all(np.diff(chainages) > 0):
            raise ValueError("chainages must be strictly increasing")
        self.xs_chainages, self.input_xs = chainages, sections

    # ---- per-node accessors ------------------------------------------------------------------
    def area_at(self, i, hw):
        return self.xs_at_node[i].area(hw)

    def hydraulic_radius(self, i, hw):
        return self.xs_at_node[i].hydraulic_radius(hw)

    def top_width(self, i, hw):
        return self.xs_at_node[i].top_width(hw)

    def bed_level_at(self, i):
        return self.xs_at_node[i].z_min

    def dA_dh(self, i, hw):
        return self.xs_at_node[i].dA_dh(hw=hw)

    def Se(self, h, Q, i):
        xs = self.xs_at_node[i]
        return xs.friction_slope(h=h, Q=Q) + xs.curvature_slope(h=h, Q=Q)

    def dSe_dA(self, h, Q, i):
        xs = self.xs_at_node[i]
        return xs.dSf_dA(h=h, Q=Q) + xs.dSc_dA(h=h, Q=Q)

    def dSe_dQ(self, h, Q, i):
        xs = self.xs_at_node[i]
        return xs.dSf_dQ(h=h, Q=Q) + xs.dSc_dQ(h=h, Q=Q)

    # ---- geometry ----------------------------------------------------------------------------
    def _provisional_sections(self):
        """Two rectangular sections from width / roughness / boundary bed levels (channel.py:282-294)."""
        us = TrapezoidalSection(b_main=self.width, m_main=0, z_bed=self.upstream_boundary.bed_level, n_main=self.roughness)
        ds = TrapezoidalSection(b_main=self.width, m_main=0, z_bed=self.downstream_boundary.bed_level, n_main=self.roughness)
        us.bed_slope = ds.bed_slope = (us.z_min - ds.z_min) / self.length
        self.upstream_boundary.cross_section, self.downstream_boundary.cross_section = us, ds
        self.xs_chainages = [self.upstream_boundary.chainage, self.downstream_boundary.chainage]
        self.input_xs = [us, ds]

    def _centreline_curvature(self):
        """Signed curvature at every interior input section from three centre-line points (channel.py:243-277)."""
        cx, cy = self.coords[:, 0], self.coords[:, 1]
        for i in range(1, len(self.input_xs) - 1):
            chs = np.array([self.xs_chainages[i - 1], self.xs_chainages[i], self.xs_chainages[i + 1]])
            pts = np.column_stack([np.interp(chs, self.coords_chainages, cx), np.interp(chs, self.coords_chainages, cy)])
            v1, v2 = pts[1] - pts[0], pts[2] - pts[1]
            n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
            if n1 == 0 or n2 == 0:
                curvature = 0.0
            else:
                turn = np.arccos(np.clip(np.dot(v1, v2) / (n1 * n2), -1.0, 1.0))
                mean_len = 0.5 * (n1 + n2)
                cross = v1[0] * v2[1] - v1[1] * v2[0]
                curvature = 2 * np.sin(turn / 2) / mean_len * np.sign(cross)
            self.input_xs[i].curvature = curvature

    def _initialize_geometry(self, n_nodes):
        if self.xs_chainages is None or self.input_xs is None:
            self._provisional_sections()
        self.ch_at_node = np.linspace(self.upstream_boundary.chainage, self.downstream_boundary.chainage, n_nodes)
        if self.coords_chainages is not None and self.coords is not None:
            self._centreline_curvature()
        nodes = []
        for s in self.ch_at_node:
            if s <= self.xs_chainages[0]:
                nodes.append(self.input_xs[0])
            elif s >= self.xs_chainages[-1]:
                nodes.append(self.input_xs[-1])
            else:
                j = int(np.searchsorted(self.xs_chainages, s)) - 1
                nodes.append(interpolate_cross_section(self.input_xs[j], self.input_xs[j + 1],
                                                       dist1=s - self.xs_chainages[j], dist2=self.xs_chainages[j + 1] - s))
        self.xs_at_node = nodes
        self.upstream_boundary.cross_section = nodes[0]
        self.downstream_boundary.cross_section = nodes[-1]

    # ---- initial conditions ------------------------------------------------------------------
    def initialize_conditions(self, n_nodes):
        self._initialize_geometry(n_nodes=n_nodes)
        self.initial_conditions = np.zeros((n_nodes, 2), dtype=np.float64)
        Q = self.initial_flow_rate
        {"linear": self._linear_conditions, "GVF_equation": self._gvh_conditions,
         "steady-state": self._steady_conditions}[self.interpolation_method](n_nodes, Q)
        self.conditions_initialized = True

    def _linear_conditions(self, n_nodes, Q):
        h0, hN = self.upstream_boundary.initial_depth, self.downstream_boundary.initial_depth
        for i in range(n_nodes):
            distance = self.length * i / (n_nodes - 1)
            self.initial_conditions[i] = (h0 + (hN - h0) * distance / self.length, Q)

    def _steady_conditions(self, n_nodes, Q):
        for i, xs in enumerate(self.xs_at_node):
            if xs.bed_slope is None:
                raise ValueError("Bed slope must be defined.")
            self.initial_conditions[i] = (xs.normal_depth(Q_target=Q), Q)

    def _gvh_conditions(self, n_nodes, Q):
        """Backwater profile marched upstream from the downstream depth with a predictor-corrector
        (channel.py:307-378); the device twin is pr_gvf_kernel."""
        dx = self.length / (n_nodes - 1)

        def slope(h_in, node, S0):
            hw = h_in + self.bed_level_at(node)
            A, T = self.area_at(node, hw), self.top_width(node, hw)
            if T < 1e-6 or A < 1e-6:
                return 0.0
            Fr = hydraulics.froude_num(T=T, A=A, Q=Q)
            if Fr > 1.0:
                raise RuntimeError(f"GVF Error: Flow became supercritical (Fr={Fr:.2f}) at node {node}. "
                                   "Downstream boundary control is not valid for this Q.")
            den = 1 - Fr ** 2
            if den < 0.01:
                print(f"Warning: GVF approaching critical depth at node {node} (Fr={Fr:.2f}). Clamping slope.")
                den = 0.01
            return (S0 - self.Se(h=h_in, Q=Q, i=node)) / den

        h = self.downstream_boundary.initial_depth
        self.initial_conditions[n_nodes - 1] = (h, Q)
        for i in reversed(range(n_nodes - 1)):
            S0 = (self.bed_level_at(i) - self.bed_level_at(i + 1)) / dx      # same S0 for both stages (channel.py:344)
            s_down = slope(h, i + 1, S0)
            h_pred = h - s_down * dx
            if h_pred <= 0:
                h_pred = 0.01
            s_pred = slope(h_pred, i, S0)
            h_up = h - 0.5 * (s_down + s_pred) * dx
            if h_up <= 0:
                print(f"Warning: GVF calculation resulted in h <= 0 at node {i}. Setting to 0.01.")
                h_up = 0.01
            h = h_up
            self.initial_conditions[i] = (h, Q)
