"""``TrapezoidalSection`` and ``interpolate_cross_section`` with the reference's constructor, attributes
and method names (cross_section.py:549-846, 857-930).  Rectangular, simple and compound trapezoids -
the only section family the shipped cases instantiate; ``IrregularSection`` is not provided.

Host-side SETUP code (geometry interpolation, initial conditions, derived results).  The per-iteration
evaluation of these quantities happens in the CUDA node pass, never here.
"""
from __future__ import annotations

import math

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


def interpolate_cross_section(xs1, xs2, dist1, dist2):
    """Distance-weighted blend of two trapezoidal sections (cross_section.py:857-930).  Returns xs1 / xs2
    themselves when the location coincides with one of them."""
    total = dist1 + dist2
    if total < 1e-9 or dist1 < 1e-9:
        return xs1
    if dist2 < 1e-9:
        return xs2
    if not (isinstance(xs1, TrapezoidalSection) and isinstance(xs2, TrapezoidalSection)):
        raise NotImplementedError("only TrapezoidalSection pairs can be interpolated (IrregularSection: SURVEY.md 8f-4)")
    w1, w2 = dist2 / total, dist1 / total
    mix = lambda a, b: a * w1 + b * w2
    slope = None if (xs1.bed_slope is None or xs2.bed_slope is None) else mix(xs1.bed_slope, xs2.bed_slope)
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
