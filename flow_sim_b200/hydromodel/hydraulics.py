"""Free hydraulic functions with the reference's names and argument meaning (hydraulics.py:1-229).

Host-side SETUP code only (initial conditions, result derivation); the time loop never calls these -
the node pass of the CUDA kernel (csrc/pr_device.cuh) carries the same formulas on the device.
Arithmetic is written in the reference's evaluation order so that initial conditions come out
bit-identical (tests/test_mirror_api.py compares them with the reference's flattened inputs).
"""
from __future__ import annotations

import math

g = 9.80665  # scipy.constants.g

_TWO_THIRDS = 2 / 3
_MINUS_THIRD = 2 / 3 - 1


def conveyance(A: float, n: float, R: float) -> float:
    return A * R ** _TWO_THIRDS / n


def dK_dA_(A, n, R, dR_dA):
    return (R ** _TWO_THIRDS + A * 2. / 3. * R ** _MINUS_THIRD * dR_dA) / n


def normal_flow(bed_slope, area=None, roughness=None, hydraulic_radius=None, K=None):
    if K is None:
        K = conveyance(A=area, n=roughness, R=hydraulic_radius)
    Q = K * abs(bed_slope) ** 0.5
    return -Q if bed_slope < 0 else Q


def dQn_dA(S_0, A=None, n=None, R=None, dR_dA=None, dK_dA=None):
    if dK_dA is None:
        dK_dA = dK_dA_(A=A, n=n, R=R, dR_dA=dR_dA)
    d = dK_dA * abs(S_0) ** 0.5
    return -d if S_0 < 0 else d


def Sf(Q, A=None, n=None, R=None, K=None):
    if K is None:
        K = conveyance(A=A, n=n, R=R)
    return Q * abs(Q) / K ** 2


def dSf_dA(Q, A=None, n=None, R=None, dR_dA=None, K=None, dK_dA=None):
    if K is None or dK_dA is None:
        K = conveyance(A=A, n=n, R=R)
        dK_dA = dK_dA_(A=A, n=n, R=R, dR_dA=dR_dA)
    return -2 * Sf(Q=Q, K=K) * (dK_dA / K)


def dSf_dQ(Q, A=None, n=None, R=None, K=None):
    if K is None:
        K = conveyance(A=A, n=n, R=R)
    return 2 * abs(Q) / K ** 2


def froude_num(T, A, Q):
    V = Q / max(A, 1e-6)
    D = A / max(T, 1e-6)
    return V / math.sqrt(g * max(D, 1e-6))


def dFr_dA(T, A, Q):
    V, D = Q / A, A / T
    return -0.5 * V * (g * D) ** (-1.5) * g * (1.0 / T) + (-Q / A ** 2) * (g * D) ** (-0.5)


def dFr_dQ(T, A):
    return (1.0 / A) * (g * (A / T)) ** (-0.5)


def darcey_weisbach_f(n, R):
    C = R ** (1 / 6) / n
    return 8 * g / C ** 2


def Sc(h, T, A, Q, n, R, rc):
    Fr = froude_num(T=T, A=A, Q=Q)
    f = darcey_weisbach_f(n=n, R=R)
    num = (2.86 * math.sqrt(f) + 2.07 * f) * h ** 2 * Fr ** 2
    den = (0.565 + math.sqrt(f)) * rc ** 2
    return num / den


def dSc_dA(h, A, Q, n, R, rc, dR_dA, T):
    Fr = froude_num(T=T, A=A, Q=Q)
    C = R ** (1 / 6) / n
    f = 8 * g / C ** 2
    dh_dA = 1. / T
    dFr = dFr_dA(A=A, Q=Q, T=T)
    df_dA = -(8.0 / 3.0) * g * n ** 2 * R ** (-4.0 / 3.0) * dR_dA
    sq = math.sqrt(f)
    num = (2.86 * sq + 2.07 * f) * h ** 2 * Fr ** 2
    den = (0.565 + sq) * rc ** 2
    dnum = (2.86 / (2 * sq) * df_dA + 2.07 * df_dA) * h ** 2 * Fr ** 2 \
        + (2.86 * sq + 2.07 * f) * (2 * h * dh_dA * Fr ** 2 + h ** 2 * 2 * Fr * dFr)
    dden = (1.0 / (2 * sq) * df_dA) * rc ** 2
    return (dnum * den - num * dden) / (den ** 2)


def dSc_dQ(h, T, A, Q, n, R, rc):
    Fr = froude_num(T=T, A=A, Q=Q)
    C = R ** (1 / 6) / n
    f = 8 * g / C ** 2
    sq = math.sqrt(f)
    num = (2.86 * sq + 2.07 * f) * h ** 2 * Fr ** 2
    den = (0.565 + sq) * rc ** 2
    dnum = (2.86 * sq + 2.07 * f) * h ** 2 * 2 * Fr * dFr_dQ(T=T, A=A)
    return (dnum * den - num * 0.0) / (den ** 2)
