"""A synthetic reach with surveyed-style polyline sections (``IrregularSection``), the companion of
``oracle/ref_harness.build_irregular`` on the mirror API.  The reference ships no case of its own with irregular
sections; this one pins the device path to a run of the live reference (tests/golden/irregular.*)."""
from math import pi, sin

import numpy as np

from ..hydromodel import Boundary, Channel, Hydrograph, IrregularSection, PreissmannSolver, TrapezoidalSection

LENGTH, BED_SLOPE, TIME_STEP = 12000.0, 0.0005, 1800


def inflow(t):
    return 60 + 40 * sin(pi * min(t, 6 * TIME_STEP) / (6 * TIME_STEP)) ** 2


def section(invert, shift, bar=False, pocket=False):
    """Main channel between stations 14 and 36, rougher overbanks.  ``bar`` adds a mid-channel bar that splits low
    flows into two equal sub-channels (the reference's Newton iteration diverges on it); ``pocket`` adds a side pocket
    behind a ridge on the right bank - a small second sub-channel that joins the main one once the ridge is overtopped
    (split-flow conveyance, cross_section.py:329-439; pinned by a reference run, tests/golden/irregular_pocket.*)."""
    x = [0, 10, 14, 20, 24, 26, 30, 36, 40, 50.0] if bar else [0, 10, 14, 20, 30, 36, 40, 50.0]
    z = [6, 3.0, 1.2, 0.0, 2.6, 2.6, 0.1, 1.5, 3.2, 6.0] if bar else [6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]
    if pocket:
        x = [0, 10, 14, 20, 30, 36, 38, 39, 41, 42, 50.0]
        z = [6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 1.9, 1.9, 3.2, 6.0]
    s = IrregularSection(x=np.array(x) + shift, z=np.array(z) + invert, n=0.03, bed_slope=BED_SLOPE)
    s.set_roughness_para((0.05, 0.03, 0.06, 14.0 + shift, 36.0 + shift))
    return s


def build(bar=False, levels=8, curved=False, pocket=False):
    up = Boundary("flow_hydrograph", chainage=0, bed_level=BED_SLOPE * LENGTH, initial_depth=2.0,
                  hydrograph=Hydrograph(function=inflow))
    down = Boundary("fixed_depth", chainage=LENGTH, bed_level=0.0, initial_depth=2.0)
    ch = Channel(upstream_boundary=up, downstream_boundary=down, initial_flow=60.0, roughness=0.03, width=30.0,
                 interpolation_method="linear")
    if curved:      # an S-bend; curvature is computed at the interior input sections (channel.py:243-277)
        sx = np.linspace(0.0, LENGTH, 25)
        ch.set_coords(coords=np.column_stack([sx, 600.0 * np.sin(2 * np.pi * sx / LENGTH)]), chainages=sx * 1.0)
        stations = [0.0, 4000.0, 8000.0, LENGTH]
        ch.set_cross_sections(stations, [section(BED_SLOPE * (LENGTH - c), c / LENGTH, bar, pocket) for c in stations])
    else:
        ch.set_cross_sections([0.0, LENGTH], [section(BED_SLOPE * LENGTH, 0.0, bar, pocket), section(0.0, 1.0, bar, pocket)])
    solver = PreissmannSolver(channel=ch, theta=0.6, time_step=TIME_STEP, spatial_step=1000.0,
                              simulation_time=levels * TIME_STEP)
    return solver, dict(tolerance=1e-6, max_iter=60)


def build_mixed(levels=8):
    """Compound trapezoid at the head, surveyed polyline at the tail, every interior node their blend (the trapezoid
    sampled on the polyline's stations, cross_section.py:795-849, 933-969): a reach whose node 0 is a trapezoid and
    whose other nodes are polylines.  Pinned by a reference run, tests/golden/mixed_sections.* (oracle/ref_harness.build_mixed)."""
    up = Boundary("flow_hydrograph", chainage=0, bed_level=BED_SLOPE * LENGTH, initial_depth=2.0,
                  hydrograph=Hydrograph(function=inflow))
    down = Boundary("fixed_depth", chainage=LENGTH, bed_level=0.0, initial_depth=2.0)
    ch = Channel(upstream_boundary=up, downstream_boundary=down, initial_flow=60.0, roughness=0.03, width=30.0,
                 interpolation_method="linear")
    z0 = BED_SLOPE * LENGTH
    head = TrapezoidalSection(z_bed=z0, b_main=12.0, m_main=2.0, n_main=0.03, z_bank=z0 + 2.4, b_fp_left=6.0, b_fp_right=9.0,
                              m_fp=3.0, n_left=0.05, n_right=0.06, bed_slope=BED_SLOPE)
    tail = IrregularSection(x=np.array([-25, -15, -11, -5, 5, 11, 15, 25.0]), z=np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]),
                            n=0.03, bed_slope=BED_SLOPE)
    tail.set_roughness_para((0.05, 0.03, 0.06, -11.0, 11.0))
    ch.set_cross_sections([0.0, LENGTH], [head, tail])
    solver = PreissmannSolver(channel=ch, theta=0.6, time_step=TIME_STEP, spatial_step=1000.0, simulation_time=levels * TIME_STEP)
    return solver, dict(tolerance=1e-6, max_iter=60)
