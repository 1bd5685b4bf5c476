"""Shared helpers of the parity tests."""
import os

import numpy as np

from flow_sim_b200.flatten import load_flat

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

#: the north star's tolerance: 1e-9 relative on stage and discharge at every node and step
RTOL = 1e-9

CALIB_MEMBERS = [0, 8191, 21845, 30000, 36408, 43690, 54321, 65535]
SMALL_CASES = ["example", "akbari", "storage_general", "gerd_release", "gerd_gated", "irregular", "irregular_curved", "irregular_pocket", "mixed_sections"] + [f"gerd_calib_m{m}" for m in CALIB_MEMBERS]


def calib_n(m):
    return 0.020 + 0.040 * m / 65535


def golden_inputs(case):
    return load_flat(os.path.join(GOLD, f"{case}.in.npz"))


def golden_outputs(case):
    return np.load(os.path.join(GOLD, f"{case}.ref.npz"))


def has_golden_outputs(case):
    return os.path.exists(os.path.join(GOLD, f"{case}.ref.npz"))


def max_rel(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over all entries."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor if floor > 0 else np.finfo(float).tiny)
    return float(np.max(np.abs(a - b) / den))


def assert_parity(got_depth, got_flow, ref_depth, ref_flow, what, rtol=RTOL):
    """Stage = depth (+ constant bed), discharge = flow.  Discharge can pass through zero (example case,
    level 1 at the storage boundary), so its error is taken relative to max(|Q|, 1e-3 * max|Q|)."""
    ed = max_rel(got_depth, ref_depth)
    qfloor = 1e-3 * float(np.max(np.abs(ref_flow)))
    eq = max_rel(got_flow, ref_flow, floor=qfloor)
    assert ed <= rtol, f"{what}: depth rel err {ed:.3e} > {rtol}"
    assert eq <= rtol, f"{what}: flow rel err {eq:.3e} > {rtol}"
    return ed, eq


#: Convergence is the absolute test ||R||_2 < tol (preissmann.py:149-153).  Two implementations whose iterates differ in
#: the last digits (the device's condensed arithmetic and cyclic reduction against the oracle's banded elimination, or
#: the oracle against SuperLU) see norms that differ by ~1e-7 relative (measured on the 65,536-member grid,
#: profiles/r02_parity_sweep_65536.json), so a decision whose norm lies within NEAR_TIE of tol can fall either way.
#: Such a member runs one Newton iteration more or fewer at that level; its hydrograph then differs by about the size
#: of that last update, far below FLIP_RTOL.
NEAR_TIE = 1e-4
FLIP_RTOL = 1e-6


def assert_iteration_parity(got, ora, tol, what, rtol=RTOL, near_tie=NEAR_TIE, members=None):
    """Iteration counts and hydrographs of an ensemble against the oracle's run with trace_prev_error=True.

    Every member whose counts agree is held to `rtol`.  A member whose counts differ must differ FIRST at a level-step
    that the oracle's own record marks as a near tie of the convergence test - by exactly one iteration, in the
    direction the tie allows - and is then held to FLIP_RTOL.  Returns the number of such members."""
    gi, oi = np.asarray(got["iters"]), np.asarray(ora["iters"])
    idx = np.arange(gi.shape[0]) if members is None else np.asarray(members)
    diff = gi[idx] != oi[idx]
    flipped = idx[diff.any(axis=1)]
    same = idx[~diff.any(axis=1)]
    if len(same):
        assert_parity(got["depth"][same], got["flow"][same], ora["depth"][same], ora["flow"][same], what, rtol)
    for m in flipped:
        k = int(np.argmax(gi[m] != oi[m]))
        d = int(gi[m, k]) - int(oi[m, k])
        norm = float(ora["final_error"][m, k] if d > 0 else ora["prev_error"][m, k])
        ok = (d == 1 and tol * (1 - near_tie) <= norm < tol) or (d == -1 and tol <= norm <= tol * (1 + near_tie))
        assert ok, (f"{what}: member {m} level {k + 1}: {gi[m, k]} iterations against {oi[m, k]} and the oracle's norm at "
                    f"that decision is {norm!r} (tol {tol}) - not a near tie")
        assert_parity(got["depth"][m], got["flow"][m], ora["depth"][m], ora["flow"][m], f"{what} (flipped member {m})",
                      FLIP_RTOL)
    return len(flipped)


def release_ensemble():
    """Two-member release-scenario ensemble assembled from two reference goldens: member 0 = gerd_release
    (pool 486.2 m, jammed gates, buffer 0.3, n = 0.03), member 1 = gerd_calib_m0 (the stock curve, n = 0.02).
    Returns (flat, [ref outputs member 0, member 1])."""
    a, b = golden_inputs("gerd_release"), golden_inputs("gerd_calib_m0")
    flat = a
    flat.down.member_ratings = [a.down.rating, b.down.rating]
    flat.member_n_main = np.array([0.03, calib_n(0)])
    flat.ic_depth = np.stack([a.ic_depth, b.ic_depth])
    flat.ic_flow = np.stack([a.ic_flow, b.ic_flow])
    return flat, [golden_outputs("gerd_release"), golden_outputs("gerd_calib_m0")]


def replace_polyline(flat, node, x, z, lim_l, lim_r):
    """Swap the polyline of one IrregularSection node of flattened inputs (CSR arrays of flat.geom)."""
    g = flat.geom
    off = g["irr_offset"]
    xs = [g["irr_x"][off[i]:off[i + 1]] for i in range(flat.n_nodes)]
    zs = [g["irr_z"][off[i]:off[i + 1]] for i in range(flat.n_nodes)]
    xs[node], zs[node] = np.asarray(x, np.float64), np.asarray(z, np.float64)
    g["irr_x"], g["irr_z"] = np.concatenate(xs), np.concatenate(zs)
    g["irr_offset"] = np.concatenate([[0], np.cumsum([len(p) for p in xs])]).astype(np.int32)
    g["irr_left"], g["irr_right"], g["z_bed"] = g["irr_left"].copy(), g["irr_right"].copy(), g["z_bed"].copy()
    g["irr_left"][node], g["irr_right"][node], g["z_bed"][node] = lim_l, lim_r, float(np.min(z))


def densify_polylines(flat, k):
    """Every segment of every polyline of flattened inputs cut into k equal pieces: the same shapes, k times the points."""
    g = flat.geom
    off = g["irr_offset"].copy()
    t = np.linspace(0.0, 1.0, k, endpoint=False)
    for i in range(flat.n_nodes):
        x, z = g["irr_x"][g["irr_offset"][i]:g["irr_offset"][i + 1]], g["irr_z"][g["irr_offset"][i]:g["irr_offset"][i + 1]]
        if len(x) < 2:
            continue
        xd = np.concatenate([x[:-1, None] + (x[1:] - x[:-1])[:, None] * t[None, :], x[-1:, None]], axis=None)
        zd = np.concatenate([z[:-1, None] + (z[1:] - z[:-1])[:, None] * t[None, :], z[-1:, None]], axis=None)
        replace_polyline(flat, i, xd, zd, g["irr_left"][i], g["irr_right"][i])
