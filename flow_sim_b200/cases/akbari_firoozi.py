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


def build_long_reach_flat(n_nodes=100_000, spatial_step=100, time_step=600, n_steps=16, theta=0.6, peak_flow=200,
                          tolerance=1e-4):
    """Config 5 without 100 000 Python objects: the FlatCase of ``build_long_reach`` assembled with array arithmetic.

    A prismatic reach is two input sections blended per node (channel.py:213-241, cross_section.py:857-930:
    ``a*w1 + b*w2`` with ``w1 = d2/(d1+d2)``), so every per-node column is one numpy expression - the same IEEE
    operations in the same order, hence bit-identical to flattening the object model (tests/test_mirror_api.py).
    The boundaries are flattened from a three-node twin of the same reach.  The initial state (normal depth per node)
    is left to the device kernel ``pr_normal_depth_initial_conditions`` (flow_sim_b200.runner); ``ic_depth`` is None
    here."""
    import numpy as np

    from .. import abi
    from ..flatten import G_STANDARD, FlatCase, flatten_boundary

    length = (n_nodes - 1) * spatial_step
    twin, _ = build(peak_flow=peak_flow, length=length, spatial_step=length / 2, time_step=time_step,
                    duration=n_steps * time_step, theta=theta)
    ch = twin.channel
    us, ds = ch.input_xs
    L = int(twin.number_of_time_levels)
    s = np.linspace(ch.upstream_boundary.chainage, ch.downstream_boundary.chainage, n_nodes)
    d1, d2 = s - ch.xs_chainages[0], ch.xs_chainages[1] - s
    tot = d1 + d2
    w1, w2 = d2 / tot, d1 / tot
    at_us, at_ds = (tot < 1e-9) | (d1 < 1e-9), d2 < 1e-9           # the input sections themselves (interpolate_cross_section)
    at_us[0], at_ds[-1] = True, True

    def mix(a, b):
        v = a * w1 + b * w2
        v[at_us], v[at_ds & ~at_us] = a, b
        return v

    zeros = np.zeros(n_nodes)
    geom = dict(kind=np.full(n_nodes, abi.PR_XS_RECT, dtype=np.int32), z_bed=mix(us.z_bed, ds.z_bed),
                b_main=mix(us.b_main, ds.b_main), m_main=mix(us.m_main, ds.m_main), h_bank=zeros.copy(),
                T_bank=zeros.copy(), W_bank=zeros.copy(), b_fp_l=zeros.copy(), b_fp_r=zeros.copy(), m_fp=zeros.copy(),
                n_l=mix(us.n_left, ds.n_left), n_m=mix(us.n_main, ds.n_main), n_r=mix(us.n_right, ds.n_right),
                curvature=mix(us.curvature, ds.curvature))
    gw1, gw2 = w1.copy(), w2.copy()
    gw1[at_us], gw2[at_us] = 1.0, 0.0
    gw1[at_ds & ~at_us], gw2[at_ds & ~at_us] = 0.0, 1.0
    gw1[-1], gw2[-1] = 1.0, 0.0                                    # flatten.interpolation_weights: end nodes carry (1, 0)
    geom["w1"], geom["w2"] = gw1, gw2
    flat = FlatCase(n_nodes=n_nodes, n_levels=L, theta=float(theta), dt=float(time_step), dx=float(length / (n_nodes - 1)),
                    tol=float(tolerance), max_iter=100, g=G_STANDARD, geom=geom,
                    up=flatten_boundary(ch.upstream_boundary, L, time_step, downstream=False),
                    down=flatten_boundary(ch.downstream_boundary, L, time_step, downstream=True),
                    ic_depth=None, ic_flow=None)
    flat.meta.update(z0=float(us.z_min), chainage=s, initial_flow=float(ch.initial_flow_rate), ic_method="steady-state",
                     bed_slope=mix(us.bed_slope, ds.bed_slope))
    return flat


def flood_wave_series(peaks, n_levels, dt, base_flow=BASE_FLOW, t_p=5 * 3600, t_b=15 * 3600):
    """``flood_wave(peak)(k*dt)`` for every peak and level as one array [len(peaks), n_levels] (same expressions as the
    scalar function, evaluated with numpy)."""
    import numpy as np

    pk = np.asarray(peaks, dtype=np.float64)[:, None]
    t = (np.arange(n_levels) * dt)[None, :].astype(np.float64)
    rise = pk / 2 * np.sin(pi * t / t_p - pi / 2) + pk / 2 + base_flow
    fall = pk / 2 * np.cos(pi * (t - t_p) / (t_b - t_p)) + pk / 2 + base_flow
    return np.where(t <= t_p, rise, np.where(t <= t_b, fall, float(base_flow)))
