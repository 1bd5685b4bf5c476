"""cases/akbari_firoozi (BASELINE config 2) and its long prismatic clone (config 5): rectangular channel,
sine-cosine flood wave, normal-depth outflow, steady uniform initial state
(reference: cases/akbari_firoozi/settings.py:1-34, main_preissmann.py:5-32)."""
from math import cos, pi, sin

from ..hydromodel import Boundary, Channel, Hydrograph, PreissmannSolver

WIDTH, LENGTH, ROUGHNESS, S_0 = 120, 29000, 0.023, 0.00061
BASE_FLOW = 100


def flood_wave(peak_flow=200, base_flow=BASE_FLOW, t_p=5 * 3600, t_b=15 * 3600):
    def q(t):
        if t <= t_p:
            return peak_flow / 2 * sin(pi * t / t_p - pi / 2) + peak_flow / 2 + base_flow
        if t <= t_b:
            return peak_flow / 2 * cos(pi * (t - t_p) / (t_b - t_p)) + peak_flow / 2 + base_flow
        return base_flow
    return q


def build(peak_flow=200, length=LENGTH, spatial_step=1000, time_step=3600, duration=20 * 3600, theta=0.5,
          tolerance=1e-4, width=WIDTH, roughness=ROUGHNESS, bed_slope=S_0):
    us = Boundary(condition="flow_hydrograph", bed_level=bed_slope * length, chainage=0,
                  hydrograph=Hydrograph(flood_wave(peak_flow)))
    ds = Boundary(condition="normal_depth", bed_level=0, chainage=length)
    ch = Channel(width=width, initial_flow=BASE_FLOW, roughness=roughness, upstream_boundary=us,
                 downstream_boundary=ds, interpolation_method="steady-state")
    solver = PreissmannSolver(channel=ch, theta=theta, time_step=time_step, spatial_step=spatial_step,
                              simulation_time=duration, regularization=False)
    return solver, dict(tolerance=tolerance)


def build_long_reach(n_nodes=100_000, spatial_step=100, time_step=600, n_steps=16, theta=0.6, peak_flow=200):
    """Config 5: the same prismatic channel stretched to n_nodes nodes."""
    return build(peak_flow=peak_flow, length=(n_nodes - 1) * spatial_step, spatial_step=spatial_step,
                 time_step=time_step, duration=n_steps * time_step, theta=theta)
