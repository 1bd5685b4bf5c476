// pr_device.cuh - device-side data model and node / boundary math of the Preissmann scheme.
//
// Everything here is FP64 and written for sm_100a.  The formulas follow the reference
// (cve-mohd/flow-sim, src/hydromodel) function by function - citations below - but are algebraically
// condensed for the FP64 pipe: one reciprocal cube root and at most three square roots per node instead
// of up to nine pow(), reciprocals shared between terms, Manning factors n^-1.5 hoisted out of the loop.  The
// condensation changes results at the few-ulp level only (tests hold the device path to 1e-9 relative
// and identical Newton iteration counts against the CPU oracle and the reference's golden outputs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/preissmann_b200.h"

namespace pr {

// ---- kernel parameter block (host builds it from the ABI structs) ---------------------------------

struct DevRating {
  int type, n_coef;
  double a, b, c, shift;
  double coef[PR_MAX_POLY], dcoef[PR_MAX_POLY];
  double off, scl;
  // Roseires, collapsed on the host (long double) to two quadratics centred on stage0:
  //   Q_state(s) = q[0] + u*(q[1] + u*q[2]),  u = s - stage0
  double lo[3], hi[3], dlt[3];   // dlt = hi - lo
  double stage0, buffer, inv_buffer, dY, inv_2dY;
  // gate_control = 1: RoseiresRatingCurve(smooth=False), per-member GateState
  int gate_control, initially_open;
  double max_cooldown;
};

// Per-member gate state of a gate-controlled Roseires curve (roseires_rating_curve.py:38-55); lives as long as the run.
struct GateState {
  double cooldown, prev_time, cur_stage;
  int open, have_prev;
};

struct DevBC {
  int type;
  double bed_level, slope_factor /* sign(S0)*sqrt|S0| */, fixed_depth;
  const double* series;
  long long series_stride;
  DevRating rc;
  int gated;                      // some member's curve is gate-controlled (stateful): needs the GST kernels
  const DevRating* member_rc;     // release scenarios: one reduced curve per member (device), or nullptr
  double st_area, st_inv_area, st_min_stage;
  // general lumped storage (lumped_storage.py:24-179): tabulated area curve, outflow rating curve, head losses
  int st_general;                 // 0 = constant area, no outflow, no losses (closed form)
  int st_curve_len, st_losses;
  const double *st_curve_stage, *st_curve_area;
  double st_alpha, st_beta, st_step, st_ymin, st_ymax, st_length, st_kq;
  DevRating st_out;
};

struct DevGeom {
  const int* kind;
  const double *z, *b, *m, *hb, *Tb, *Wb, *bl, *br, *mfp, *nl, *nm, *nr, *curv;
  const double *member_nm, *member_nfp;
  // IrregularSection nodes (pr_irregular.cuh): CSR polylines and composite-roughness limits, or nullptr
  const int* irr_offset;
  const double *irr_x, *irr_z, *irr_left, *irr_right;
  // stage tables of the IrregularSection nodes, built on the device once per call (pr_irregular.cuh, pr_irr_build_tables):
  // breakpoints = the distinct vertex elevations of a section; per interval the polynomials of area, wetted perimeter
  // and top width of the whole section and of its three roughness sub-sections.  Stored at the node's irr_offset.
  const int* irr_tab_n;        // [N] number of breakpoints (0: no table - the node pass scans the polyline)
  const double* irr_tab_z;     // [points] breakpoints, ascending
  const double* irr_tab_c;     // [points][kIrrTabCols] coefficients of the interval above breakpoint k
  const int* irr_tab_runs;     // [points] wetted sub-channels (runs of >= 2 submerged points) in that interval
};

struct DevParams {
  int N, L, M, max_iter, out_mode;
  double theta, dt, dx, tol, tol2, g;       // tol2 = tol^2
  double i2dt, th_dx, hth, omt_dx, homt;   // 1/(2dt), theta/dx, theta/2, (1-theta)/dx, (1-theta)/2
  double ghth, th_dx2, mtheta;             // g*theta/2, 2*theta/dx, -theta
  DevGeom geo;
  DevBC up, dn;
  const double *ic_h, *ic_q;
  long long ic_stride;
  double *out_h, *out_q;
  int *iters, *status, *fail_level;
  double *storage_stage, *final_error;
  // persistent scheduling of the fused kernel (pr_ensemble_kernel.cuh): zeroed ticket counter of this launch, number
  // of tickets, optional processing order of the members (a permutation of 0..M-1), SM count of the device
  unsigned int* ticket;
  int n_tickets, sm_count;
  const int* member_order;
};

// ---- FP64 primitives without slow-path branches ----------------------------------------------------
// CUDA's IEEE division / sqrt / cbrt carry special-case branches (BSSY/BSYNC/CALL) that cost issue slots
// and diverge.  The scheme only needs ~1e-15 relative accuracy on positive, normal arguments, so these use
// the SFU seed (MUFU.RCP64H / RSQ64H, fp32 LG2/EX2) plus one second-order correction, all on the FMA pipe.

// Seeds are accurate to 2^-20 (measured, tests/test_gpu_math.py); ONE correction step that keeps the quadratic
// term of the error series brings them to < 2^-58, i.e. full double precision up to the final rounding.
__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);          // 1/a = r / (1 - e) = r (1 + e + e^2 + O(e^3))
  return fma(r, fma(e, e, e), r);
}

__device__ __forceinline__ double fast_sqrt(double a) {   // a >= 0; sqrt(0) = 0 without a select:
  double y;                                               // the seed of a + tiny is finite, and 0 * finite = 0
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a + 1e-300));
  const double g = a * y;
  const double e = fma(-g, y, 1.0);          // sqrt(a) = g (1 - e)^(-1/2) = g (1 + e/2 + 3e^2/8 + O(e^3))
  return fma(g, e * fma(e, 0.375, 0.5), g);
}

__device__ __forceinline__ double fast_sqrt_pos(double a) {   // a > 0: no guard against the infinite seed of 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = a * y;
  const double e = fma(-g, y, 1.0);
  return fma(g, e * fma(e, 0.375, 0.5), g);
}

__device__ __forceinline__ double fast_rsqrt(double a) {  // a > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a * y, y, 1.0);
  return fma(y, e * fma(e, 0.375, 0.5), y);
}

__device__ __forceinline__ double fast_rcbrt(double x) {  // x^(-1/3), x > 0 within float range
  const float xf = __double2float_rn(x);
  float lg, sf;                                                 // raw SFU ops: no denormal fix-up code around them
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sf) : "f"(-0.33333334f * lg));
  const double r = (double)sf;                                  // ~2^-20 relative
  const double e = fma(-x * r, r * r, 1.0);  // x^(-1/3) = r (1 - e)^(-1/3) = r (1 + e/3 + 2e^2/9 + O(e^3))
  return fma(r, e * fma(e, 2.0 / 9.0, 1.0 / 3.0), r);
}

// A node that an update has left dry (depth <= 0) ends the member at its next residual evaluation: the reference clamps
// the depth to zero (cross_section.py:628-632: A = P = R = T = 0) and dies of the zero conveyance that follows.  The
// node pass leaves the depth unclamped - and a depth below -b / (2 sqrt(1 + m^2)) makes A and P both negative, their
// ratio positive, so that a member the reference has lost could run on (1 member in ~8,000 of the device fuzz,
// tools/fuzz_device.py) - so the UPDATE turns such a depth into NaN: an integer test on the high word (depths below
// 2^-1022 * 2^20 count as dry), no FP64-pipe work (a clamp in the node pass cost the headline kernel 1.8 %).
__device__ __forceinline__ double poison_dry(double h) {
  return __double2hiint(h) <= 0 ? __longlong_as_double(0x7ff8000000000000LL) : h;
}

// ---- per-node geometry staged in shared memory (SoA: field f of slot idx at sg[f*NP + idx]) -----------

enum GeoField {
  F_KIND = 0, F_Z, F_B, F_M, F_SQM, F_HB, F_TB, F_WB, F_BL, F_BR, F_MFP, F_SQFP, F_AMF, F_PM, F_INVPM,
  F_NM, F_INVNM, F_CNL, F_CNM, F_CNR, F_CURV, F_COUNT
};

__device__ __forceinline__ double inv_n15(double n) { return 1.0 / (n * sqrt(n)); }  // n^-1.5

// Fills sg[F_COUNT][NP].  Slot `idx` holds node map(idx) (identity for the per-thread kernels; lane-major
// for the cooperative kernel so that a warp reads consecutive doubles).  Nodes >= N replicate node N-1 so
// padded lanes compute finite throw-away values.
template <class Map>
__device__ inline void stage_geometry(const DevGeom& g, int N, int NP, double* sg, int tid, int nthreads, Map map) {
  for (int idx = tid; idx < NP; idx += nthreads) {
    const int node = map(idx);
    const int s = node < N ? node : N - 1;
    const double b = g.b[s], m = g.m[s], hb = g.hb[s], Tb = g.Tb[s], mfp = g.mfp[s];
    const double sqm = sqrt(1.0 + m * m);
    const double Pm = b + 2.0 * hb * sqm;               // cross_section.py:661,695
    sg[F_KIND * NP + idx] = (double)g.kind[s];
    sg[F_Z * NP + idx] = g.z[s];
    sg[F_B * NP + idx] = b;
    sg[F_M * NP + idx] = m;
    sg[F_SQM * NP + idx] = sqm;
    sg[F_HB * NP + idx] = g.kind[s] == PR_XS_COMPOUND ? hb : 1e300;
    sg[F_TB * NP + idx] = Tb;
    sg[F_WB * NP + idx] = g.Wb[s];
    sg[F_BL * NP + idx] = g.bl[s];
    sg[F_BR * NP + idx] = g.br[s];
    sg[F_MFP * NP + idx] = mfp;
    sg[F_SQFP * NP + idx] = sqrt(1.0 + mfp * mfp);
    sg[F_AMF * NP + idx] = (b + Tb) / 2.0 * hb;         // cross_section.py:660
    sg[F_PM * NP + idx] = Pm;
    sg[F_INVPM * NP + idx] = Pm > 0.0 ? 1.0 / Pm : 0.0;
    sg[F_NM * NP + idx] = g.nm[s];
    sg[F_INVNM * NP + idx] = 1.0 / g.nm[s];
    sg[F_CNL * NP + idx] = inv_n15(g.nl[s]);
    sg[F_CNM * NP + idx] = inv_n15(g.nm[s]);
    sg[F_CNR * NP + idx] = inv_n15(g.nr[s]);
    sg[F_CURV * NP + idx] = g.curv[s];
  }
}

// Roughness mode (template parameter RM of the kernels): bit 0 = per-member n_main override,
// bit 1 = per-member n_fp override (model.run(n_main=, n_fp=)); 0 = the node's own values;
// 4 = decided at run time from the two flags below (the long-reach kernels, which are not built per mode).
struct Rough {
  double nm, inm, cnm, cnfp;    // n_main, 1/n_main, n_main^-1.5, n_fp^-1.5
  double nfp;                   // n_fp itself (irregular sections)
  bool om, ofp;                 // RM = 4: which overrides are present
};

template <int RM>
__device__ __forceinline__ bool rough_main(const Rough& r) { return RM == 4 ? r.om : (RM & 1) != 0; }
template <int RM>
__device__ __forceinline__ bool rough_fp(const Rough& r) { return RM == 4 ? r.ofp : (RM & 2) != 0; }

template <int RM>
__device__ __forceinline__ Rough load_rough(const DevGeom& g, long long member) {
  Rough rg;
  rg.om = RM == 4 ? g.member_nm != nullptr : (RM & 1) != 0;
  rg.ofp = RM == 4 ? g.member_nfp != nullptr : (RM & 2) != 0;
  rg.nm = rg.om ? g.member_nm[member] : 1.0;
  rg.inm = 1.0 / rg.nm;
  rg.cnm = inv_n15(rg.nm);
  rg.nfp = rg.ofp ? g.member_nfp[member] : 1.0;
  rg.cnfp = inv_n15(rg.nfp);
  return rg;
}

__host__ __device__ inline int rough_mode(const DevGeom& g) { return (g.member_nm ? 1 : 0) | (g.member_nfp ? 2 : 0); }

// Everything the two adjacent cells need from one node at the current iterate.  The Jacobian pieces that
// depend on this node alone are formed here once instead of once per adjacent cell.
struct NodeVals {
  double Q;     // discharge
  double A;     // wetted area                       (TrapezoidalSection.properties, cross_section.py:623-679)
  double T;     // top width = dA/dh                 (:792-793)
  double Y;     // water level z_min + h             (Solver.water_level_at, solver.py:287-288)
  double Se;    // energy slope Sf + Sc              (Channel.Se, channel.py:53-69)
  double F;     // Q^2 / A
  double QA;    // Q / A
  double w1;    // (theta/dx) (Q/A)^2 T              |d_dQ2Adx_dA * dA_dh|       (preissmann.py:540,546)
  double w2;    // (theta/2) dSe_dA T                d_avgSe_dA * dA_dh          (:543,547); dSe_dA = dSf_dA + dSc_dA,
                //                                   the latter already x dA/dh (quirk 7, channel.py:71-87)
  double w3;    // (theta/2) dSe_dQ                  d_avgSe_dQ                  (:665; channel.py:89-105)
  double w4;    // 2 (theta/dx) Q/A                  |d_dQ2Adx_dQ|               (:662)
};

struct NodeConv {
  double K;     // conveyance                        (cross_section.py:741-754)
  double dKA;   // dK/dA                             (cross_section.py:756-764)
  double A, Sf, dSfA, dSfQ;   // area and friction slope with its derivatives (hydraulics.py:42-92), for the
                              // head losses of the lumped storage (lumped_storage.py:47-143)
};

// Node pass.  h = depth unknown, Q = discharge unknown, idx = node slot in shared memory.
//
// Conveyance without pow():  with r = X^(-1/3),
//   simple / in-bank : X = R = A/P,  K = A R^(2/3)/n = A (R r)/n,      1/K^2 = (n/A)^2 r^4
//   over-bank        : X = S = sum_j A_j^1.5 R_j n_j^-1.5 (= sum K_j^1.5),  K = S^(2/3) = S r,  1/K^2 = r^4
// and dK/dA = K (5/(3A) - (2/3) dP_dh/(T P)) in every case (n_eq frozen at A R^(2/3)/K, quirk 6).
// One reciprocal of A*P*T yields 1/A and 1/(T P); the over-bank branch needs one more for the
// floodplain hydraulic radii.  The expensive tail (reciprocal, cube root) is common to all branches, so
// lanes of a warp that sit on different branches re-converge before it.
template <bool CURV, int RM, bool WANT_K = false, class KP = DevParams, bool CMP = true>
__device__ __forceinline__ void node_eval(const double* __restrict__ sg, const int NP, const int idx, const double h,
                                          const double Q, const Rough& rg, const KP& k, NodeVals& o,
                                          NodeConv* kc = nullptr) {
#define GEO(f) sg[(f)*NP + idx]
  const double z = GEO(F_Z), b = GEO(F_B);
  const double hw = z + h;        // Solver.water_level_at
  const double d = hw - z;        // depth = max(0, hw - z_bed) in the reference; a dry node ends the member: poison_dry()
  // Branch-free section geometry: the in-bank (simple trapezoid; a rectangle is m = 0) and the over-bank
  // expressions are both evaluated and selected, so that the whole node pass is straight-line code the
  // scheduler can interleave with its neighbours.  Non-compound sections are staged with h_bank = 1e300.
  // CMP = false: the reach has no compound section, only the in-bank expressions are compiled.
  const double hb = CMP ? GEO(F_HB) : 0.0;
  const bool over = CMP && d > hb;
  double A, P, T, dPdh, X = 0.0;
  {
    const double m = GEO(F_M), sqm = GEO(F_SQM);
    const double Ti = b + 2.0 * m * d;
    const double Ai = (b + Ti) * 0.5 * d;
    const double Pi = b + 2.0 * d * sqm;
    if (CMP) {
      const double dfp = over ? d - hb : 1.0, mfp = GEO(F_MFP), sqfp = GEO(F_SQFP);
      const double bl = GEO(F_BL), br = GEO(F_BR);
      const double hm = 0.5 * mfp * dfp;
      const double Al = (bl + hm) * dfp, Pl = bl + dfp * sqfp;
      const double Ar = (br + hm) * dfp, Pr = br + dfp * sqfp;
      const double Amf = GEO(F_AMF);
      const double Ao = Amf + Al + Ar;                       // quirk 4: total area omits T_bank*dfp
      const double Po = GEO(F_PM) + Pl + Pr;
      const double To = GEO(F_WB) + 2.0 * mfp * dfp;
      // K_j^1.5 = A_j*sqrt(A_j) * (A_j/P_j) * n_j^-1.5   (cross_section.py:681-754, hydraulics.py:15-26)
      const double Am = Amf + GEO(F_TB) * dfp;               // conveyance area includes the column (:694)
      const double cnm = rough_main<RM>(rg) ? rg.cnm : GEO(F_CNM);
      const double cnl = rough_fp<RM>(rg) ? rg.cnfp : GEO(F_CNL);
      const double cnr = rough_fp<RM>(rg) ? rg.cnfp : GEO(F_CNR);
      const double w = fast_rcp(Pl * Pr);
      double Xo = (Am * Am) * fast_sqrt_pos(Am) * (GEO(F_INVPM) * cnm);   // Am >= bankfull area > 0
      Xo = fma((Al * Al) * fast_sqrt(Al), (Pr * w) * cnl, Xo);
      Xo = fma((Ar * Ar) * fast_sqrt(Ar), (Pl * w) * cnr, Xo);
      A = over ? Ao : Ai;
      P = over ? Po : Pi;
      T = over ? To : Ti;
      dPdh = over ? sqfp : sqm;               // dP/dh / 2
      X = Xo;
    } else {
      A = Ai; P = Pi; T = Ti; dPdh = sqm;
    }
  }
  const double PT = P * T;
  const double u = fast_rcp(A * PT);
  const double invA = u * PT, invTP = u * A;
  const double nm = rough_main<RM>(rg) ? rg.nm : GEO(F_NM);
  if (!over) X = A * (invTP * T);                          // R = A/P
  const double r = fast_rcbrt(X);
  const double r2 = r * r, r4 = r2 * r2;
  double invK2;
  if (over) {
    invK2 = r4;
  } else {
    const double na = nm * invA;
    invK2 = (na * na) * r4;
  }
  const double dKA_over_K = fma(-4.0 / 3.0, dPdh * invTP, (5.0 / 3.0) * invA);     // dPdh holds dP/dh / 2
  const double absQ = fabs(Q);
  const double aq = absQ * invK2;
  const double Sf = Q * aq;                                 // hydraulics.py:42-57
  double Se = Sf;
  double dSeA = -2.0 * Sf * dKA_over_K;                     // hydraulics.py:59-75   (dead code without curvature, see w2)
  double dSeQ = 2.0 * aq;                                   // hydraulics.py:77-92
  double K = 0.0;
  if (WANT_K || CURV) {
    const double inm = rough_main<RM>(rg) ? rg.inm : GEO(F_INVNM);
    K = over ? X * r : A * (X * r) * inm;
  }
  if (CURV) {
    const double curv = GEO(F_CURV);
    if (curv != 0.0) {
      // hydraulics.Sc / dSc_dA / dSc_dQ (hydraulics.py:94-229), CrossSection wrappers cross_section.py:143-175
      const double invT = invTP * P, invP = invTP * T;
      const double R = A * invP;
      const double rR = over ? fast_rcbrt(R) : r;           // R^(-1/3)
      const double cR = R * rR * rR;                        // R^(1/3)
      const double n_eq = (GEO(F_KIND) == (double)PR_XS_COMPOUND) ? A * (cR * cR) / K : nm;    // cross_section.py:710-739
      const double rc = 1.0 / curv;
      const double V = Q / fmax(A, 1e-6), D = A / fmax(T, 1e-6);
      const double Fr = V / sqrt(k.g * fmax(D, 1e-6));
      const double f = 8.0 * k.g * (n_eq * n_eq) / cR;      // C = R^(1/6)/n, f = 8g/C^2
      const double sqf = sqrt(f);
      const double lead = 2.86 * sqf + 2.07 * f;
      const double num = lead * (h * h) * (Fr * Fr);
      const double den = (0.565 + sqf) * (rc * rc);
      Se += num / den;
      if (fabs(curv) > 1e-12) {
        const double dRA = (P - A * (2.0 * dPdh) * invT) * (invP * invP);
        const double gD = k.g * (A * invT);
        const double rs = rsqrt(gD);                        // (gD)^-0.5, unclamped (quirk 7)
        const double dFrA = -0.5 * (Q * invA) * (rs * rs * rs) * k.g * invT + (-Q * invA * invA) * rs;
        const double dFrQ = invA * rs;
        const double dfA = -(8.0 / 3.0) * k.g * (n_eq * n_eq) / (R * cR) * dRA;
        const double dnumA = (2.86 / (2.0 * sqf) * dfA + 2.07 * dfA) * (h * h) * (Fr * Fr) +
                             lead * (2.0 * h * invT * (Fr * Fr) + (h * h) * 2.0 * Fr * dFrA);
        const double ddenA = (1.0 / (2.0 * sqf) * dfA) * (rc * rc);
        const double dScA = (dnumA * den - num * ddenA) / (den * den);
        dSeA += dScA * T;                                   // cross_section.py:164 (x dA_dh), multiplied again by the caller
        const double dnumQ = lead * (h * h) * 2.0 * Fr * dFrQ;
        dSeQ += (dnumQ * den) / (den * den);
      }
    }
  }
  const double QA = Q * invA;
  o.Q = Q;
  o.A = A;
  o.T = T;
  o.Y = hw;
  o.Se = Se;
  o.F = Q * QA;
  o.QA = QA;
  o.w1 = (k.th_dx * QA) * (QA * T);
  if (CURV) {
    o.w2 = (k.hth * dSeA) * T;
    o.w3 = k.hth * dSeQ;
  } else {                                                  // (theta/2) * (-2 Sf dK/K) * T and (theta/2) * 2 |Q| / K^2
    o.w2 = (k.mtheta * (Sf * dKA_over_K)) * T;
    o.w3 = k.theta * aq;
  }
  o.w4 = k.th_dx2 * QA;
  if (WANT_K) {
    kc->K = K;
    kc->dKA = K * dKA_over_K;
    kc->A = A;
    kc->Sf = Sf;
    kc->dSfA = -2.0 * Sf * dKA_over_K;
    kc->dSfQ = 2.0 * aq;
  }
#undef GEO
}

// ---- rating curves ----------------------------------------------------------------------------------

__device__ __forceinline__ double horner(const double* c, int n, double x) {
  double acc = c[n - 1];
  for (int i = n - 2; i >= 0; --i) acc = c[i] + acc * x;
  return acc;
}

__device__ __forceinline__ double roseires_q(const DevRating& r, double stage) {
  // RoseiresRatingCurve.alpha_smooth + effective_release (roseires_rating_curve.py:87-109)
  // alpha = 0 below stage0, 1 above stage0 + buffer, smoothstep between: clamping s does all three
  const double u = stage - r.stage0;
  double s = u * r.inv_buffer;
  s = s > 0.0 ? s : 0.0;                                    // (cheaper than fmin/fmax; a NaN stage still ends in NaN)
  s = s < 1.0 ? s : 1.0;
  const double alpha = s * s * (3.0 - 2.0 * s);
  const double lo = r.lo[0] + u * (r.lo[1] + u * r.lo[2]);
  const double dl = r.dlt[0] + u * (r.dlt[1] + u * r.dlt[2]);
  return fma(alpha, dl, lo);                                // (1 - alpha) lo + alpha hi
}

__device__ __forceinline__ void gate_init(GateState& g, const DevRating& r) {
  g.cooldown = 0.0; g.prev_time = 0.0; g.cur_stage = r.stage0; g.open = r.initially_open; g.have_prev = 0;
}

// RoseiresRatingCurve.gate_control (roseires_rating_curve.py:111-130): the decision uses the stage stored by the
// PREVIOUS residual evaluation; 0.5 / 1.0 are the reference's hard-coded thresholds.
__device__ __forceinline__ void gate_control(GateState& g, const DevRating& r, double time) {
  if (g.have_prev) g.cooldown = fmax(0.0, g.cooldown - (time - g.prev_time));
  g.prev_time = time; g.have_prev = 1;
  if (g.cooldown > 0.0) return;
  if (g.cur_stage >= r.stage0 + 0.5 && !g.open) { g.cooldown = r.max_cooldown; g.open = 1; }
  else if (g.cur_stage <= r.stage0 - 1.0 && g.open) { g.cooldown = r.max_cooldown; g.open = 0; }
}

// total_release with the gates in their current position
__device__ __forceinline__ double roseires_gated_q(const DevRating& r, int open, double stage) {
  const double u = stage - r.stage0;
  const double* c = open ? r.hi : r.lo;
  return c[0] + u * (c[1] + u * c[2]);
}

// RatingCurve.discharge (rating_curve.py:32-63)
__device__ __forceinline__ double rating_q(const DevRating& r, double stage) {
  switch (r.type) {
    case PR_RC_POLY2: { const double x = stage + r.shift; return r.a * (x * x) + r.b * x + r.c; }
    case PR_RC_POWER: return r.a * pow(stage + r.shift, r.b);
    case PR_RC_POLYNOMIAL: return horner(r.coef, r.n_coef, r.off + r.scl * stage);   // NB: no shift (:48-49)
    case PR_RC_ROSEIRES: return roseires_q(r, stage);
    default: return nan("");
  }
}

// RatingCurve.dQ_dz (rating_curve.py:132-147); Roseires: central difference, dY = 1e-3 (quirk 10)
__device__ __forceinline__ double rating_dq(const DevRating& r, double stage) {
  switch (r.type) {
    case PR_RC_POLY2: return r.a * 2.0 * (stage + r.shift) + r.b;
    case PR_RC_POWER: return r.a * r.b * pow(stage + r.shift, r.b - 1.0);
    case PR_RC_POLYNOMIAL: return horner(r.dcoef, r.n_coef - 1, r.off + r.scl * (stage + r.shift));
    case PR_RC_ROSEIRES: return (roseires_q(r, stage + r.dY) - roseires_q(r, stage - r.dY)) * r.inv_2dY;
    default: return nan("");
  }
}

// ---- Brent's method (scipy/optimize/Zeros/brentq.c: xtol = 2e-12, rtol = 4 eps, maxiter = 100) -----------------
// sign_error mirrors brentq's ValueError when f(xa) and f(xb) have the same sign.
template <class F>
__device__ __forceinline__ double brent_root(F f, const double xa, const double xb, bool& sign_error) {
  const double xtol = 2e-12, rtol = 8.881784197001252e-16;
  double xpre = xa, xcur = xb, xblk = 0.0, fpre = f(xpre), fcur = f(xcur), fblk = 0.0, spre = 0.0, scur = 0.0;
  sign_error = false;
  if (fpre == 0.0) return xpre;
  if (fcur == 0.0) return xcur;
  if (signbit(fpre) == signbit(fcur)) { sign_error = true; return nan(""); }
  for (int it = 0; it < 100; ++it) {
    if (fpre != 0.0 && fcur != 0.0 && signbit(fpre) != signbit(fcur)) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
    if (fabs(fblk) < fabs(fcur)) { xpre = xcur; xcur = xblk; xblk = xpre; fpre = fcur; fcur = fblk; fblk = fpre; }
    const double delta = (xtol + rtol * fabs(xcur)) / 2.0, sbis = (xblk - xcur) / 2.0;
    if (fcur == 0.0 || fabs(sbis) < delta) return xcur;
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      double stry;
      if (xpre == xblk) stry = -fcur * (xcur - xpre) / (fcur - fpre);
      else {
        const double dpre = (fpre - fcur) / (xpre - xcur), dblk = (fblk - fcur) / (xblk - xcur);
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
      }
      if (2.0 * fabs(stry) < fmin(fabs(spre), 3.0 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
      else { spre = sbis; scur = sbis; }
    } else { spre = sbis; scur = sbis; }
    xpre = xcur; fpre = fcur;
    if (fabs(scur) > delta) xcur += scur;
    else xcur += (sbis > 0.0 ? delta : -delta);
    fcur = f(xcur);
  }
  return xcur;
}

// numpy.interp on an increasing table
__device__ __forceinline__ double table_interp(double x, const double* xp, const double* fp, int n) {
  if (x != x) return x;
  if (x > xp[n - 1]) return fp[n - 1];
  if (x < xp[0]) return fp[0];
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (x >= xp[mid]) lo = mid + 1; else hi = mid;
  }
  const int j = lo - 1;
  if (j >= n - 1) return fp[n - 1];
  if (xp[j] == x) return fp[j];
  return (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]) * (x - xp[j]) + fp[j];
}

// forward declaration (rating curves are defined above bc_eval)
__device__ __forceinline__ double rating_q(const DevRating& r, double stage);

// LumpedStorage.area_at / net_vol_change / mass_balance (lumped_storage.py:24-45,155-179), general form
__device__ __forceinline__ double storage_area_at(const DevBC& b, double stage) {
  if (b.st_curve_len <= 0) return b.st_area;
  return b.st_alpha * table_interp(stage + b.st_beta, b.st_curve_stage, b.st_curve_area, b.st_curve_len);
}

__device__ inline double storage_net_vol_change(const DevBC& b, double Y1, double Y2) {
  if (b.st_curve_len <= 0) return (Y2 - Y1) * b.st_area;
  const int n = (int)(fabs(Y2 - Y1) / b.st_step);
  if (n > 2) {      // np.trapezoid over np.linspace(Y1, Y2, n)
    const double dstep = (Y2 - Y1) / (n - 1);
    double sum = 0.0, y_prev = Y1, a_prev = storage_area_at(b, Y1);
    for (int i = 1; i < n; ++i) {
      const double y = (i == n - 1) ? Y2 : i * dstep + Y1;
      const double a = storage_area_at(b, y);
      sum += (y - y_prev) * (a + a_prev) / 2.0;
      y_prev = y; a_prev = a;
    }
    return sum;
  }
  return 0.5 * (storage_area_at(b, Y2) + storage_area_at(b, Y1)) * (Y2 - Y1);
}

__device__ inline double storage_mass_balance(const DevBC& b, double vol_in, double Y_old, double dt, bool& failed) {
  const double q_old = b.st_out.type != PR_RC_NONE ? rating_q(b.st_out, Y_old) : 0.0;
  auto f = [&](double Y_new) {
    const double Q_out = b.st_out.type != PR_RC_NONE ? 0.5 * (q_old + rating_q(b.st_out, Y_new)) : 0.0;
    return storage_net_vol_change(b, Y_old, Y_new) - (vol_in - Q_out * dt);
  };
  double Y = brent_root(f, b.st_ymin, b.st_ymax, failed);
  if (Y < b.st_min_stage) Y = b.st_min_stage;
  return Y;
}

// ---- boundary rows (Boundary.condition_residual / df_dh / df_dQ, boundary.py:56-242) --------------------

struct BcRow {
  double res, dh, dq;
  double stage_rec;   // storage: reservoir stage recorded by this evaluation (boundary.py:126-131)
  bool fail;          // the reference would raise here (brentq: "f(a) and f(b) must have different signs")
};

// level = time // dt; hyd = series sample at this level; q_prev = stored Q of the previous level at the node;
// stage_prev = reservoir stage recorded for level-1.
// GST: compile the general lumped-storage branch (Brent solve, area curve, losses).  It is rare and register-hungry,
// so the kernels that do not need it (every shipped case) are instantiated without it.
template <bool GST>
__device__ __forceinline__ BcRow bc_eval(const DevBC& bc, const int member, const int level, const double hyd, const double h,
                                         const double Q, const double q_prev, const double stage_prev,
                                         const double dt, const double g, const NodeConv& kc, const double T,
                                         GateState* gate = nullptr) {
  const double K = kc.K, dKA = kc.dKA;
  BcRow o;
  o.stage_rec = 0.0;
  o.fail = false;
  switch (bc.type) {
    case PR_BC_FLOW_HYDROGRAPH:
      o.res = Q - hyd; o.dh = 0.0; o.dq = 1.0;
      break;
    case PR_BC_STAGE_HYDROGRAPH:
      o.res = h - (hyd - bc.bed_level); o.dh = 1.0; o.dq = 0.0;
      break;
    case PR_BC_FIXED_DEPTH:
      o.res = h - bc.fixed_depth; o.dh = 1.0; o.dq = 0.0;
      break;
    case PR_BC_NORMAL_DEPTH:
      // hydraulics.normal_flow / dQn_dA (hydraulics.py:4-13, 206-215); host checks bed_level == z_min
      o.res = Q - K * bc.slope_factor;
      o.dh = 0.0 - dKA * bc.slope_factor * T;
      o.dq = 1.0;
      break;
    case PR_BC_RATING_CURVE: {
      const double stage = bc.bed_level + h;
      // two instantiations on purpose: the shared curve is read straight from the kernel-parameter constant bank,
      // the per-member one from global memory
      auto rows = [&](const DevRating& rc) {
        if (GST && rc.type == PR_RC_ROSEIRES && rc.gate_control) {
          // discharge(update_gate_state=True, update_stage=True) then dQ_dz with the state frozen (:65-81, :202-208)
          gate_control(*gate, rc, level * dt);
          o.res = Q - roseires_gated_q(rc, gate->open, stage);
          gate->cur_stage = stage;
          o.dh = 0.0 - (roseires_gated_q(rc, gate->open, stage + rc.dY) - roseires_gated_q(rc, gate->open, stage - rc.dY)) * rc.inv_2dY;
        } else {
          o.res = Q - rating_q(rc, stage);
          o.dh = 0.0 - rating_dq(rc, stage);
        }
        o.dq = 1.0;
      };
      if (bc.member_rc) rows(bc.member_rc[member]);
      else rows(bc.rc);
      break;
    }
    case PR_BC_FIXED_DEPTH_STORAGE: {
      const double vol_in = 0.5 * (q_prev + Q) * dt;                       // preissmann.py:314
      const double y_old = (level == 1) ? h + bc.bed_level : stage_prev;    // quirk 9
      double y_new, dy_dvol;
      if (!GST || !bc.st_general) {
        // LumpedStorage.mass_balance with constant area: the brentq root of
        // (Y - Y_old)*A_s - vol_in is Y_old + vol_in/A_s, clamped at min_stage (lumped_storage.py:24-45)
        y_new = y_old + vol_in * bc.st_inv_area;
        dy_dvol = bc.st_inv_area;
        // brentq brackets the root in solution_boundaries and raises when it lies outside
        o.fail = !(y_new >= bc.st_ymin && y_new <= bc.st_ymax);
        if (y_new < bc.st_min_stage) y_new = bc.st_min_stage;
      } else {
        bool failed;
        y_new = storage_mass_balance(bc, vol_in, y_old, dt, failed);       // brentq, as the reference
        o.fail = failed;
        dy_dvol = 1.0 / storage_area_at(bc, y_new);                        // dY_new_dvol_in, :37-45
      }
      if (y_new <= bc.st_min_stage) dy_dvol = 0.0;
      double hl = 0.0, dhl_dA = 0.0, dhl_dQ = 0.0;
      if (GST && bc.st_losses) {
        // friction over the reservoir length + empirical K_q V^2/2g (lumped_storage.py:47-143)
        const double V = Q / kc.A, i2g = 1.0 / (2.0 * g);
        hl = kc.Sf * bc.st_length + bc.st_kq * (V * V) * i2g;
        dhl_dA = kc.dSfA * bc.st_length + bc.st_kq * 2.0 * V * (-Q / (kc.A * kc.A)) * i2g;
        dhl_dQ = kc.dSfQ * bc.st_length + bc.st_kq * 2.0 * V * (1.0 / kc.A) * i2g;
      }
      o.stage_rec = y_new;
      o.res = h - ((y_new + hl) - bc.bed_level);
      o.dh = 1.0 - dhl_dA * T;
      o.dq = 0.0 - (dy_dvol * (0.5 * dt) + dhl_dQ);
      break;
    }
    default:
      o.res = nan(""); o.dh = 1.0; o.dq = 0.0;
  }
  return o;
}

}  // namespace pr
