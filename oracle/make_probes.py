#!/usr/bin/env python
"""Section-level known answers from the LIVE reference (run in the build container, where /root/reference or its staged
copy oracle/_ref/pyref exists) -> small fixtures under tests/golden/:

* irregular_pocket_probe.npz - friction slope, its derivatives, conveyance, area of the side-pocket sections of
  `ref_harness.build_irregular(pocket=True)` at (node, depth, flow) probe points, 114 of the 252 with several wetted
  sub-channels (cross_section.py:329-439).  The probe points are kept when the fixture already exists (values are
  re-evaluated, so `--check` proves the committed numbers are the reference's), else drawn with a fixed seed.
* irregular_dense_probe.npz - the same quantities for a 201-point version of one of those sections (sub-channels of
  ~100 points, no stage table on the device).
* trapezoid_z_at_probe.npz - `TrapezoidalSection.z_at` (cross_section.py:795-849) of a rectangle, a simple and a compound
  trapezoid on a lateral grid, and the blend of a compound trapezoid with a polyline (`interpolate_cross_section`,
  :933-969) at three weights.

    python oracle/make_probes.py            # (re)write both fixtures
    python oracle/make_probes.py --check    # re-evaluate and compare with the committed fixtures, bit for bit
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

COLUMNS = ["node", "h", "Q", "subchannels", "Sf", "dSf_dA", "dSf_dQ", "K", "dK_dA", "A", "dA_dh"]


def pocket_rows():
    import ref_harness as rh

    solver, _ = rh.build_irregular(pocket=True)
    path = os.path.join(GOLD, "irregular_pocket_probe.npz")
    if os.path.exists(path):
        points = np.load(path)["rows"][:, :3]
    else:
        rng = np.random.default_rng(20261018)
        points = []
        for node in (0, 3, 7, 12):
            depths = np.concatenate([np.arange(0.3, 4.55, 0.1), rng.uniform(1.9, 3.3, 20)])
            points += [(node, h, rng.uniform(20.0, 120.0)) for h in depths]
        points = np.array(points)
    sections = solver.channel.xs_at_node
    rows = []
    for node, h, Q in points:
        xs = sections[int(node)]
        hw = h + xs.z_min
        rows.append([node, h, Q, len(xs.get_subchannels(hw)), xs.friction_slope(h, Q), xs.dSf_dA(h, Q), xs.dSf_dQ(h, Q),
                     xs.conveyance(hw), xs.dK_dA(hw), xs.area(hw), xs.dA_dh(hw)])
    return dict(rows=np.array(rows, dtype=np.float64), columns=np.array(COLUMNS))


def densify(x, z, k):
    """Every segment of a polyline cut into k equal pieces (the same shape, k times the points)."""
    t = np.linspace(0.0, 1.0, k, endpoint=False)
    xs = np.concatenate([x[:-1, None] + (x[1:] - x[:-1])[:, None] * t[None, :], x[-1:, None]], axis=None)
    zs = np.concatenate([z[:-1, None] + (z[1:] - z[:-1])[:, None] * t[None, :], z[-1:, None]], axis=None)
    return xs, zs


def dense_rows():
    """The side-pocket section of node 7 with 20 points per segment (201 points: beyond the device's stage tables, and
    sub-channels of ~100 points): the same quantities as pocket_rows on a fixed stage / flow grid."""
    import ref_harness as rh

    solver, _ = rh.build_irregular(pocket=True)
    rh.setup_reference()
    from src.hydromodel.cross_section import IrregularSection

    base = solver.channel.xs_at_node[7]
    x, z = densify(np.asarray(base.x, float), np.asarray(base.z, float), 20)
    xs = IrregularSection(x=x, z=z, n=base.n_main, bed_slope=base.bed_slope)
    xs.set_roughness_para(base.get_roughness_para())
    rows = []
    for h in np.arange(0.35, 4.4, 0.11):
        Q = 40.0 + 17.0 * h
        hw = h + xs.z_min
        rows.append([7, h, Q, len(xs.get_subchannels(hw)), xs.friction_slope(h, Q), xs.dSf_dA(h, Q), xs.dSf_dQ(h, Q),
                     xs.conveyance(hw), xs.dK_dA(hw), xs.area(hw), xs.dA_dh(hw)])
    return dict(rows=np.array(rows, dtype=np.float64), columns=np.array(COLUMNS), x=x, z=z,
                roughness=np.array(base.get_roughness_para(), dtype=np.float64))


def z_at_rows():
    import ref_harness as rh

    rh.setup_reference()
    from src.hydromodel.cross_section import IrregularSection, TrapezoidalSection, interpolate_cross_section

    grid = np.concatenate([np.linspace(-40.0, 40.0, 161), [-6.0, 6.0, -10.8, 10.8, -16.8, 19.8]])
    shapes = dict(rect=TrapezoidalSection(z_bed=1.5, b_main=12.0, m_main=0.0, n_main=0.03),
                  simple=TrapezoidalSection(z_bed=1.5, b_main=12.0, m_main=2.0, n_main=0.03),
                  compound=TrapezoidalSection(z_bed=1.5, b_main=12.0, m_main=2.0, n_main=0.03, z_bank=3.9, b_fp_left=6.0,
                                              b_fp_right=9.0, m_fp=3.0, n_left=0.05, n_right=0.06, bed_slope=5e-4))
    out = dict(grid=grid)
    for name, xs in shapes.items():
        out[f"z_{name}"] = np.array([xs.z_at(v) for v in grid], dtype=np.float64)
    poly = IrregularSection(x=np.array([-25, -15, -11, -5, 5, 11, 15, 25.0]), z=np.array([6, 3.0, 1.2, 0.0, 0.1, 1.5, 3.2, 6.0]),
                            n=0.03, bed_slope=5e-4)
    poly.set_roughness_para((0.05, 0.03, 0.06, -11.0, 11.0))
    for j, (d1, d2) in enumerate([(1000.0, 11000.0), (6000.0, 6000.0), (11000.0, 1000.0)]):
        for tag, (a, b) in (("tp", (shapes["compound"], poly)), ("pt", (poly, shapes["compound"]))):
            s = interpolate_cross_section(a, b, d1, d2)
            out[f"blend_{tag}{j}_x"], out[f"blend_{tag}{j}_z"] = np.asarray(s.x, float), np.asarray(s.z, float)
            out[f"blend_{tag}{j}_par"] = np.array([s.n_left, s.n_main, s.n_right, s.left_fp_limit, s.right_fp_limit,
                                                   s.bed_slope, s.curvature], dtype=np.float64)
    out["blend_dists"] = np.array([(1000.0, 11000.0), (6000.0, 6000.0), (11000.0, 1000.0)])
    return out


def main():
    check = "--check" in sys.argv[1:]
    for name, make in (("irregular_pocket_probe.npz", pocket_rows), ("irregular_dense_probe.npz", dense_rows),
                       ("trapezoid_z_at_probe.npz", z_at_rows)):
        data = make()
        path = os.path.join(GOLD, name)
        if check:
            old = np.load(path)
            for k, v in data.items():
                assert np.array_equal(old[k], v, equal_nan=v.dtype.kind == "f"), f"{name}: {k} differs from the live reference"
            print(name, "matches the live reference")
        else:
            np.savez_compressed(path, **data)
            print("wrote", path)


if __name__ == "__main__":
    main()
