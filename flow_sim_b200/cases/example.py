"""cases/example (BASELINE config 1): 20 km rectangular channel, trapezoidal inflow wave, fixed-depth outlet
backed by a constant-area lumped storage (reference driver: cases/example/main.py:8-61)."""
from ..hydromodel import Boundary, Channel, Hydrograph, LumpedStorage, PreissmannSolver


def inflow(t, base=1000, peak=10000, rise=3 * 3600, hold=6 * 3600, fall=4 * 3600):
    if t <= 0:
        return base
    if t < rise:
        return base + (peak - base) * t / rise
    if t - rise < hold:
        return peak
    if t - rise - hold < fall:
        return peak - (peak - base) * (t - rise - hold) / fall
    return base


def build(theta=0.8, time_step=3600, spatial_step=1000, simulation_time=24 * 3600, hydrograph=inflow):
    us = Boundary(condition="flow_hydrograph", bed_level=5, chainage=0, hydrograph=Hydrograph(function=hydrograph))
    ds = Boundary(condition="fixed_depth", initial_depth=5, bed_level=0, chainage=20000)
    ds.set_lumped_storage(LumpedStorage(surface_area=5000 * 250, min_stage=5, solution_boundaries=(0, 200)))
    ch = Channel(width=250, initial_flow=us.hydrograph.get_at(0), roughness=0.027, upstream_boundary=us,
                 downstream_boundary=ds)
    solver = PreissmannSolver(channel=ch, theta=theta, time_step=time_step, spatial_step=spatial_step,
                              simulation_time=simulation_time)
    return solver, dict(tolerance=1e-4, max_iter=100)
