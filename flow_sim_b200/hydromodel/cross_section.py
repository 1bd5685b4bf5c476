"""``TrapezoidalSection``, ``IrregularSection`` and ``interpolate_cross_section`` with the reference's constructors,
attributes and method names (cross_section.py:207-543, 549-846, 857-969).  Rectangular, simple and compound
trapezoids are what the shipped cases instantiate; polyline sections are the SURVEY.md 8f-4 widening.

Host-side SETUP code (geometry interpolation, initial conditions, derived results).  The per-iteration
evaluation of these quantities happens in the CUDA node pass, never here.
"""
from __future__ import annotations

import math

import numpy as np

from . import hydraulics


class CrossSection:
    """Common roughness / slope bookkeeping (cross_section.py:6-31,106-112)."""

    def __init__(self, n=None, bed_slope=None, curvature=0.0):
        self.n_left = self.n_main = self.n_right = n
        self.left_fp_limit = 0.0
        self.right_fp_limit = 0.0
        self.curvature = curvature
        self.bed_slope = bed_slope

    def get_roughness_para(self):
        return (self.n_left, self.n_main, self.n_right, self.left_fp_limit, self.right_fp_limit)

    def set_roughness_para(self, parameters):
        self.n_left, self.n_main, self.n_right, self.left_fp_limit, self.right_fp_limit = parameters


class TrapezoidalSection(CrossSection):
    def __init__(self, z_bed, b_main, m_main, n_main, z_bank=None, b_fp_left=0.0, b_fp_right=0.0, m_fp=0.0,
                 n_left=0.03, n_right=0.03, **kwargs):
        super().__init__(n=n_main, **kwargs)
        self.z_bed, self.b_main, self.m_main = float(z_bed), float(b_main), float(m_main)
        self._z_min = self.z_bed
        self._width = math.inf
        self._is_compound = z_bank is not None
        if self._is_compound:
            self.z_bank = float(z_bank)
            if self.z_bank <= self.z_bed:
                raise ValueError("Bank elevation z_bank must be above bed z_bed")
            self.b_fp_left, self.b_fp_right, self.m_fp = float(b_fp_left), float(b_fp_right), float(m_fp)
            self.bankfull_depth = self.z_bank - self.z_bed
            self.T_main_at_bank = self.b_main + 2.0 * self.m_main * self.bankfull_depth
            self.left_fp_limit = -self.T_main_at_bank / 2.0
            self.right_fp_limit = self.T_main_at_bank / 2.0
            self._width_at_bank = self.b_fp_left + self.T_main_at_bank + self.b_fp_right
        else:
            self.z_bank = None
            self.b_fp_left = self.b_fp_right = self.m_fp = 0.0
            self.left_fp_limit, self.right_fp_limit = -math.inf, math.inf
        self._is_rect = (not self._is_compound) and self.m_main == 0.0
        self.set_roughness_para((n_left, n_main, n_right, self.left_fp_limit, self.right_fp_limit))

    # ---- geometry ------------------------------------------------------------------------
    @property
    def z_min(self):
        return self._z_min

    @property
    def width(self):
        return self._width

    def _overbank(self, depth):
        return self._is_compound and depth > self.bankfull_depth

    def properties(self, hw):
        """(A, P, R, T) at water surface elevation hw (cross_section.py:623-679)."""
        depth = max(0.0, float(hw) - self.z_bed)
        if depth <= 0.0:
            return (0.0, 0.0, 0.0, 0.0)
        b, m = self.b_main, self.m_main
        if self._is_rect:
            A, P, T = b * depth, b + 2.0 * depth, b
        elif not self._overbank(depth):
            T = b + 2.0 * m * depth
            A = (b + T) / 2.0 * depth
            P = b + 2.0 * depth * math.sqrt(1.0 + m ** 2)
        else:
            dfp = depth - self.bankfull_depth
            wall = math.sqrt(1.0 + self.m_fp ** 2)
            A = ((b + self.T_main_at_bank) / 2.0 * self.bankfull_depth
                 + (self.b_fp_left + 0.5 * self.m_fp * dfp) * dfp
                 + (self.b_fp_right + 0.5 * self.m_fp * dfp) * dfp)     # quirk: no T_bank*dfp column (:660,672)
            P = ((b + 2.0 * self.bankfull_depth * math.sqrt(1.0 + m ** 2))
                 + (self.b_fp_left + dfp * wall) + (self.b_fp_right + dfp * wall))
            T = self._width_at_bank + 2.0 * self.m_fp * dfp
        return (A, P, A / P if P > 0.0 else 0.0, T)

    def area(self, hw):
        return self.properties(hw)[0]

    def z_at(self, x):
        """Bed elevation at lateral station ``x`` of the section centred on 0 (cross_section.py:795-849) - what the
        blend of a trapezoid with a polyline samples.  Outside a rectangle's bed the elevation is +inf."""
        x = float(x)
        half = self.b_main / 2.0
        if self._is_rect:
            return self.z_bed if -half < x < half else np.inf
        if self._is_compound and not (self.left_fp_limit <= x <= self.right_fp_limit):
            # floodplain bed at bank level, then the outer wall at 1 : m_fp
            beyond = (self.left_fp_limit - self.b_fp_left) - x if x < self.left_fp_limit \
                else x - (self.right_fp_limit + self.b_fp_right)
            return self.z_bank if beyond <= 0 else self.z_bank + beyond / self.m_fp
        if -half <= x <= half:
            return self.z_bed
        return self.z_bed + ((x - half) if x > half else (-x - half)) / self.m_main

    def wetted_perimeter(self, hw):
        return self.properties(hw)[1]

    def hydraulic_radius(self, hw):
        return self.properties(hw)[2]

    def top_width(self, hw):
        return self.properties(hw)[3]

    def dA_dh(self, hw):
        return self.top_width(hw)

    def _subsections(self, hw):
        """((A,P,R) left, main, right) with bed-only perimeters (cross_section.py:681-708)."""
        depth = max(0.0, hw - self.z_bed)
        zero = (0, 0, 0)
        if depth <= 0.0:
            return zero, zero, zero
        if not self._overbank(depth):
            A, P, R, _ = self.properties(hw)
            return zero, (A, P, R), zero
        dfp = depth - self.bankfull_depth
        wall = math.sqrt(1.0 + self.m_fp ** 2)
        Am = (self.b_main + self.T_main_at_bank) / 2.0 * self.bankfull_depth + self.T_main_at_bank * dfp
        Pm = self.b_main + 2.0 * self.bankfull_depth * math.sqrt(1.0 + self.m_main ** 2)
        out = [None, (Am, Pm, Am / Pm if Pm > 0 else 0.0), None]
        for slot, bfp in ((0, self.b_fp_left), (2, self.b_fp_right)):
            A = (bfp + 0.5 * self.m_fp * dfp) * dfp
            P = bfp + dfp * wall
            out[slot] = (A, P, A / P if P > 0 else 0.0)
        return tuple(out)

    # ---- conveyance ----------------------------------------------------------------------
    def conveyance(self, hw):
        if not self._is_compound:
            A, _, R, _ = self.properties(hw)
            return hydraulics.conveyance(A, self.n_main, R)
        (Al, _, Rl), (Am, _, Rm), (Ar, _, Rr) = self._subsections(hw)
        Kl = hydraulics.conveyance(Al, self.n_left, Rl)
        Km = hydraulics.conveyance(Am, self.n_main, Rm)
        Kr = hydraulics.conveyance(Ar, self.n_right, Rr)
        return (Kl ** 1.5 + Km ** 1.5 + Kr ** 1.5) ** (2.0 / 3.0)

    def get_equivalent_n(self, hw):
        if not self._is_compound:
            return self.n_main
        K = self.conveyance(hw)
        A, _, R, _ = self.properties(hw)
        if A <= 0 or R <= 0 or K <= 0.0:
            return self.n_main
        return (A * (R ** (2.0 / 3.0))) / K

    def dR_dA(self, hw):
        A, P, _, T = self.properties(hw)
        if P <= 0.0 or T <= 0.0:
            return 0.0
        depth = max(0.0, hw - self.z_bed)
        if self._is_rect:
            dP_dh = 2.0
        elif self._overbank(depth):
            dP_dh = 2.0 * math.sqrt(1.0 + self.m_fp ** 2)
        else:
            dP_dh = 2.0 * math.sqrt(1.0 + self.m_main ** 2)
        dP_dA = dP_dh * (1.0 / T)
        return (P - A * dP_dA) / (P ** 2)

    def dK_dA(self, hw):
        A, _, R, _ = self.properties(hw)
        if A <= 0.0:
            return 0.0
        return hydraulics.dK_dA_(A=A, n=self.get_equivalent_n(hw), R=R, dR_dA=self.dR_dA(hw))

    # ---- slopes (CrossSection concrete methods, cross_section.py:114-202) --------------------
    def friction_slope(self, h, Q):
        return hydraulics.Sf(Q=Q, K=self.conveyance(hw=h + self.z_min))

    def dSf_dA(self, h, Q):
        hw = h + self.z_min
        return hydraulics.dSf_dA(Q=Q, K=self.conveyance(hw=hw), dK_dA=self.dK_dA(hw=hw))

    def dSf_dQ(self, h, Q):
        return hydraulics.dSf_dQ(Q=Q, K=self.conveyance(hw=h + self.z_min))

    def curvature_slope(self, h, Q):
        if self.curvature == 0:
            return 0.0
        hw = h + self.z_min
        A, _, R, T = self.properties(hw)
        return hydraulics.Sc(h=h, T=T, A=A, Q=Q, n=self.get_equivalent_n(hw), R=R, rc=1.0 / self.curvature)

    def dSc_dA(self, h, Q):
        if abs(self.curvature) <= 1e-12:
            return 0.0
        hw = h + self.z_min
        A, _, R, T = self.properties(hw)
        return hydraulics.dSc_dA(h=h, A=A, Q=Q, n=self.get_equivalent_n(hw), R=R, rc=1.0 / self.curvature,
                                 dR_dA=self.dR_dA(hw), T=T) * self.dA_dh(hw)

    def dSc_dQ(self, h, Q):
        if abs(self.curvature) <= 1e-12:
            return 0.0
        hw = h + self.z_min
        A, _, R, T = self.properties(hw)
        return hydraulics.dSc_dQ(h=h, T=T, A=A, Q=Q, n=self.get_equivalent_n(hw), R=R, rc=1.0 / self.curvature)

    def normal_flow(self, hw):
        if self.bed_slope is None or self.bed_slope <= 0.0:
            return 0.0
        return hydraulics.normal_flow(bed_slope=self.bed_slope, K=self.conveyance(hw=hw))

    def normal_depth(self, Q_target, hw_max=None):
        from scipy.optimize import brentq

        z = self.z_min
        hw_max = z + 100 if hw_max is None else hw_max
        f = lambda hw: Q_target - self.normal_flow(hw=hw)
        try:
            return brentq(f, z, hw_max) - z
        except ValueError:
            if f(z) < 0:
                return 0.0
            if f(hw_max) > 0:
                return hw_max - z
            return 0.0


class IrregularSection(CrossSection):
    """Surveyed section given as a polyline (x, z), with composite roughness between ``left_fp_limit`` and
    ``right_fp_limit`` (cross_section.py:207-543).  Area and hydraulic radius are differentiated by central
    differences with dh = 1e-6, as the reference does."""

    def __init__(self, x, z, **kwargs):
        super().__init__(**kwargs)
        x = np.ascontiguousarray(x, dtype=float)
        z = np.ascontiguousarray(z, dtype=float)
        if x.shape != z.shape:
            raise ValueError("x and z must have the same shape")
        if x.ndim != 1:
            raise ValueError("x and z must be 1-D arrays")
        order = np.argsort(x)
        self.x, self.z = x[order], z[order]
        self.left_fp_limit, self.right_fp_limit = self.x[0], self.x[-1]

    @property
    def z_min(self):
        return float(self.z.min())

    @property
    def width(self):
        return float(self.x.max() - self.x.min())

    def _runs(self, wet):
        """[first, last] index pairs of the maximal runs of True."""
        edges = np.flatnonzero(np.diff(np.concatenate(([0], wet.astype(np.int8), [0]))))
        return [(int(a), int(b) - 1) for a, b in zip(edges[0::2], edges[1::2])]

    def properties(self, hw):
        hw = float(hw)
        x, z = self.x, self.z
        if hw <= self.z_min:
            return 0.0, 0.0, 0.0, 0.0
        area = perimeter = top = 0.0
        for first, last in self._runs(hw - z > 0.0):
            px, pz = x[first:last + 1], z[first:last + 1]
            if first > 0 and z[first - 1] > hw:             # the bank cuts the water surface
                frac = (hw - z[first - 1]) / (z[first] - z[first - 1])
                px = np.insert(px, 0, x[first - 1] + frac * (x[first] - x[first - 1]))
                pz = np.insert(pz, 0, hw)
            if last < x.size - 1 and z[last + 1] > hw:
                frac = (hw - z[last]) / (z[last + 1] - z[last])
                px = np.append(px, x[last] + frac * (x[last + 1] - x[last]))
                pz = np.append(pz, hw)
            depth = np.maximum(hw - pz, 0.0)
            span, rise = np.diff(px), np.diff(pz)
            area += np.sum(0.5 * (depth[:-1] + depth[1:]) * span)
            perimeter += np.sum(np.sqrt(span ** 2 + rise ** 2))
            top += px[-1] - px[0]
        radius = area / perimeter if perimeter > 0.0 else 0.0
        return float(area), float(perimeter), float(radius), float(top)

    def area(self, hw):
        return self.properties(hw)[0]

    def wetted_perimeter(self, hw):
        return self.properties(hw)[1]

    def hydraulic_radius(self, hw):
        return self.properties(hw)[2]

    def top_width(self, hw):
        return self.properties(hw)[3]

    def z_at(self, x):
        return np.interp(x, self.x, self.z, left=self.z[0], right=self.z[-1])

    def sub_channels(self, hw):
        """Number of separately wetted sub-channels with at least two submerged points."""
        return sum(1 for a, b in self._runs(self.z < hw) if b - a + 1 >= 2)

    def _part_conveyance(self, hw, x_from, x_to, n_value):
        inside = (self.x >= x_from) & (self.x <= x_to)
        if inside.sum() < 2:
            return 0.0
        part = IrregularSection(x=self.x[inside], z=self.z[inside])
        A, P, _, _ = part.properties(hw)
        if A <= 0 or P <= 0:
            return 0.0
        return hydraulics.conveyance(A=A, n=n_value, R=A / P)

    def get_equivalent_n(self, hw):
        parts = (self._part_conveyance(hw, self.x[0], self.left_fp_limit, self.n_left),
                 self._part_conveyance(hw, self.left_fp_limit, self.right_fp_limit, self.n_main),
                 self._part_conveyance(hw, self.right_fp_limit, self.x[-1], self.n_right))
        A, P, _, _ = self.properties(hw)
        if A <= 0 or P <= 0:
            return self.n_main
        K = (parts[0] ** 1.5 + parts[1] ** 1.5 + parts[2] ** 1.5) ** (2.0 / 3.0)
        if K <= 0.0:
            return self.n_main
        return (A * (A / P) ** (2.0 / 3.0)) / K

    def conveyance(self, hw):
        A = self.area(hw)
        if A <= 0.0:
            return 0.0
        return hydraulics.conveyance(A=A, n=self.get_equivalent_n(hw), R=self.hydraulic_radius(hw))

    def dR_dA(self, hw, dh=1e-6):
        lo, hi = self.properties(hw - dh), self.properties(hw + dh)
        if hi[0] - lo[0] == 0.0:
            return 0.0
        return (hi[2] - lo[2]) / (hi[0] - lo[0])

    def dA_dh(self, hw, dh=1e-6):
        return (self.area(hw + dh) - self.area(hw - dh)) / (2 * dh)

    def dK_dA(self, hw):
        A = self.area(hw)
        if A <= 0.0:
            return 0.0
        return hydraulics.dK_dA_(A=A, n=self.get_equivalent_n(hw), R=self.hydraulic_radius(hw), dR_dA=self.dR_dA(hw))

    def get_subchannels(self, hw):
        """The separately wetted parts of the section as point lists: every run of at least two submerged points, with
        the points where it meets the water surface added at its ends (cross_section.py:329-372).  The edge stations
        come from ``np.interp`` called exactly as the reference calls it - on the left edge with a decreasing abscissa
        pair, for which numpy returns the first submerged station itself (a vertical wall), not the intersection."""
        x, z = self.x, self.z
        parts = []
        for first, last in self._runs(z < hw):
            if last - first + 1 < 2:
                continue
            px, pz = x[first:last + 1], z[first:last + 1]
            if first > 0 and z[first - 1] > hw:
                px = np.insert(px, 0, np.interp(hw, [z[first - 1], z[first]], [x[first - 1], x[first]]))
                pz = np.insert(pz, 0, hw)
            if last + 1 < x.size and z[last] < hw and z[last + 1] > hw:
                px = np.append(px, np.interp(hw, [z[last], z[last + 1]], [x[last], x[last + 1]]))
                pz = np.append(pz, hw)
            parts.append({"x": px, "z": pz})
        return parts

    def _split_conveyance(self, hw):
        """(sum K_j^1.5, sum 1.5 K_j^0.5 dK_j/dA_j) over the sub-channels, each taken as a section of its own with this
        section's roughness limits and values (cross_section.py:380-392, 406-414); None for a single channel."""
        parts = self.get_subchannels(hw)
        if len(parts) <= 1:
            return None
        k_sum = dk_sum = 0.0
        for part in parts:
            sub = IrregularSection(x=part["x"], z=part["z"])
            sub.set_roughness_para(self.get_roughness_para())
            k = sub.conveyance(hw)
            k_sum += k ** 1.5
            dk_sum += 1.5 * (k ** 0.5) * sub.dK_dA(hw)
        return k_sum, dk_sum

    def friction_slope(self, h, Q):
        hw = h + self.z_min
        split = self._split_conveyance(hw)
        if split is None:
            return hydraulics.Sf(Q=Q, K=self.conveyance(hw))
        return hydraulics.Sf(Q=Q, K=split[0] ** (2.0 / 3.0))

    def dSf_dA(self, h, Q):
        hw = h + self.z_min
        split = self._split_conveyance(hw)
        if split is None:
            return hydraulics.dSf_dA(Q=Q, K=self.conveyance(hw), dK_dA=self.dK_dA(hw))
        k_sum, dk_sum = split
        return hydraulics.dSf_dA(Q=Q, K=k_sum ** (2.0 / 3.0), dK_dA=(2.0 / 3.0) * k_sum ** (-1.0 / 3.0) * dk_sum)

    def dSf_dQ(self, h, Q):
        hw = h + self.z_min
        split = self._split_conveyance(hw)
        if split is None:
            return hydraulics.dSf_dQ(Q=Q, K=self.conveyance(hw))
        return hydraulics.dSf_dQ(Q=Q, K=split[0] ** (2.0 / 3.0))

    # the curvature slope only needs properties / n_eq / dR_dA / dA_dh, which both section families provide
    curvature_slope = TrapezoidalSection.curvature_slope
    dSc_dA = TrapezoidalSection.dSc_dA
    dSc_dQ = TrapezoidalSection.dSc_dQ
    normal_flow = TrapezoidalSection.normal_flow
    normal_depth = TrapezoidalSection.normal_depth


def interpolate_cross_section(xs1, xs2, dist1, dist2):
    """Distance-weighted blend of two sections (cross_section.py:857-969): two trapezoids give a trapezoid, a polyline
    on either side gives a polyline on the union of the lateral stations.  Returns xs1 / xs2
    themselves when the location coincides with one of them."""
    total = dist1 + dist2
    if total < 1e-9 or dist1 < 1e-9:
        return xs1
    if dist2 < 1e-9:
        return xs2
    w1, w2 = dist2 / total, dist1 / total
    mix = lambda a, b: a * w1 + b * w2
    slope = None if (xs1.bed_slope is None or xs2.bed_slope is None) else mix(xs1.bed_slope, xs2.bed_slope)
    if not (isinstance(xs1, TrapezoidalSection) and isinstance(xs2, TrapezoidalSection)):
        # a polyline on either side: blend bed elevations on the union of the lateral stations (cross_section.py:933-969)
        stations = [s.x for s in (xs1, xs2) if isinstance(s, IrregularSection)]
        grid = np.union1d(stations[0], stations[1]) if len(stations) == 2 else stations[0]
        bed = lambda s: np.array([s.z_at(v) for v in grid], dtype=float)
        out = IrregularSection(x=grid, z=bed(xs1) * w1 + bed(xs2) * w2, n=mix(xs1.n_main, xs2.n_main), bed_slope=slope,
                               curvature=mix(xs1.curvature, xs2.curvature))
        out.set_roughness_para((mix(xs1.n_left, xs2.n_left), mix(xs1.n_main, xs2.n_main), mix(xs1.n_right, xs2.n_right),
                                mix(xs1.left_fp_limit, xs2.left_fp_limit), mix(xs1.right_fp_limit, xs2.right_fp_limit)))
        return out
    z_bed = mix(xs1.z_bed, xs2.z_bed)
    y1 = (xs1.z_bank - xs1.z_bed) if xs1._is_compound else 0.0
    y2 = (xs2.z_bank - xs2.z_bed) if xs2._is_compound else 0.0
    y_bank = mix(y1, y2)
    return TrapezoidalSection(
        z_bed=z_bed, b_main=mix(xs1.b_main, xs2.b_main), m_main=mix(xs1.m_main, xs2.m_main),
        z_bank=(z_bed + y_bank) if y_bank > 1e-6 else None,
        b_fp_left=mix(xs1.b_fp_left, xs2.b_fp_left), b_fp_right=mix(xs1.b_fp_right, xs2.b_fp_right),
        m_fp=mix(xs1.m_fp, xs2.m_fp),
        n_main=mix(xs1.n_main, xs2.n_main), n_left=mix(xs1.n_left, xs2.n_left), n_right=mix(xs1.n_right, xs2.n_right),
        bed_slope=slope, curvature=mix(xs1.curvature, xs2.curvature))
