/*
 * preissmann_oracle.c - CPU restatement of the reference's Preissmann hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this; the product path never does.
 *
 * Parity status: PINNED.  oracle/make_golden.py runs the live reference (cve-mohd/flow-sim,
 * /root/reference) and tests/test_oracle_golden.py checks this file against those outputs for
 * configs 1-4 (depth, flow <= 1e-9 relative, identical Newton iteration counts), against per-iteration
 * (J.data, R, delta) captures, and against the reference's two rating-curve CSVs.
 *
 * Every function cites the reference file:line it follows; arithmetic is written in the reference's
 * own evaluation order (compile with -ffp-contract=off so no FMA is formed).  Third-party pieces that
 * are not under /root/reference are restated from their published algorithms:
 *   - scipy.sparse.linalg.spsolve (SuperLU, scipy unpinned, 1.18.1 here; call site preissmann.py:146)
 *       -> Gaussian elimination with partial pivoting on the (kl=ku=2) band; same solution up to rounding.
 *   - scipy.optimize.brentq (call site lumped_storage.py:30) -> Brent's method as in scipy/optimize/Zeros.
 *   - sklearn Pipeline(PolynomialFeatures(2), LinearRegression).predict (roseires_rating_curve.py:185,200)
 *       -> intercept + dot(coef, [s, o, s^2, s*o, o^2]).
 */
#include "preissmann_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TWO_THIRDS (2.0 / 3.0)           /* Python: 2/3 and 2.0/3.0 -> 0.6666666666666666 */
#define M_ONE_THIRD (2.0 / 3.0 - 1.0)    /* Python: 2/3-1 -> -0.33333333333333337 (hydraulics.py:40) */

typedef struct {
  int kind;
  double z, b, m, hb, Tb, Wb, bl, br, mfp, nl, nm, nr, curv;
  /* IrregularSection (cross_section.py:207-543): polyline sorted by x, composite-roughness limits */
  const double *px, *pz;
  int np;
  double lim_l, lim_r;
} xs_t;

static xs_t xs_load(const pr_geom* g, int i, int member) {
  xs_t s;
  s.kind = g->kind[i];
  s.z = g->z_bed[i]; s.b = g->b_main[i]; s.m = g->m_main[i];
  s.hb = g->h_bank[i]; s.Tb = g->T_bank[i]; s.Wb = g->W_bank[i];
  s.bl = g->b_fp_l[i]; s.br = g->b_fp_r[i]; s.mfp = g->m_fp[i];
  s.nl = g->n_l[i]; s.nm = g->n_m[i]; s.nr = g->n_r[i];
  s.curv = g->curvature[i];
  s.px = s.pz = NULL; s.np = 0; s.lim_l = s.lim_r = 0.0;
  if (s.kind == PR_XS_IRREGULAR) {
    s.px = g->irr_x + g->irr_offset[i]; s.pz = g->irr_z + g->irr_offset[i];
    s.np = g->irr_offset[i + 1] - g->irr_offset[i];
    s.lim_l = g->irr_left[i]; s.lim_r = g->irr_right[i];
  }
  if (g->member_n_main) {            /* cross_section.py:892 with xs1.n_main == xs2.n_main == v */
    double v = g->member_n_main[member];
    s.nm = v * g->w1[i] + v * g->w2[i];
  }
  if (g->member_n_fp) {              /* cross_section.py:891,893 */
    double v = g->member_n_fp[member];
    s.nl = v * g->w1[i] + v * g->w2[i];
    s.nr = s.nl;
  }
  return s;
}


static double hy_conveyance(double A, double n, double R);
static double xs_dA_dh(const xs_t* s, double hw);

/* ---------------- IrregularSection (cross_section.py:207-543) ---------------- */

/* numpy's pairwise summation of a contiguous double array (np.sum, loops_utils.h.src pairwise_sum) */
static double np_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

/* IrregularSection.properties (:247-327) on the sub-polyline [lo, hi] taken as a section of its own (what
 * get_equivalent_n's subsection_props builds with IrregularSection(x[mask], z[mask]), :448-470). */
static void irr_properties(const double* x, const double* z, int lo, int hi, double hw, double* A_out, double* P_out,
                           double* T_out) {
  *A_out = *P_out = *T_out = 0.0;
  int n = hi - lo + 1;
  if (n <= 0) return;
  x += lo; z += lo;
  double z_min = z[0];
  for (int i = 1; i < n; ++i) if (z[i] < z_min) z_min = z[i];
  if (hw <= z_min) return;
  double A_total = 0.0, P_total = 0.0, T_total = 0.0;
  double* xs = (double*)malloc(sizeof(double) * (n + 2) * 4);
  double *zs = xs + (n + 2), *ta = zs + (n + 2), *tp = ta + (n + 2);
  int i = 0;
  while (i < n) {
    if (hw - z[i] > 0.0) {
      int i0 = i;
      while (i + 1 < n && hw - z[i + 1] > 0.0) i += 1;
      int iN = i, m = 0;
      if (i0 > 0 && z[i0 - 1] > hw) {                          /* intersection with the water surface on the left */
        double z0 = z[i0 - 1], z1 = z[i0], x0 = x[i0 - 1], x1 = x[i0];
        double t = (hw - z0) / (z1 - z0);
        xs[m] = x0 + t * (x1 - x0); zs[m] = hw; ++m;
      }
      for (int k = i0; k <= iN; ++k) { xs[m] = x[k]; zs[m] = z[k]; ++m; }
      if (iN < n - 1 && z[iN + 1] > hw) {
        double z0 = z[iN], z1 = z[iN + 1], x0 = x[iN], x1 = x[iN + 1];
        double t = (hw - z0) / (z1 - z0);
        xs[m] = x0 + t * (x1 - x0); zs[m] = hw; ++m;
      }
      for (int k = 0; k + 1 < m; ++k) {
        double d0 = fmax(hw - zs[k], 0.0), d1 = fmax(hw - zs[k + 1], 0.0);
        double dx = xs[k + 1] - xs[k], dz = zs[k + 1] - zs[k];
        ta[k] = 0.5 * (d0 + d1) * dx;
        tp[k] = sqrt(dx * dx + dz * dz);
      }
      A_total += np_sum(ta, m - 1);
      P_total += np_sum(tp, m - 1);
      T_total += xs[m - 1] - xs[0];
    }
    i += 1;
  }
  free(xs);
  *A_out = A_total; *P_out = P_total; *T_out = T_total;
}

static void irr_props_full(const xs_t* s, double hw, double* A, double* P, double* R, double* T) {
  irr_properties(s->px, s->pz, 0, s->np - 1, hw, A, P, T);
  *R = (*P > 0.0) ? *A / *P : 0.0;
}

/* number of wetted sub-channels with at least two submerged points (get_subchannels, :329-372) */
static int irr_subchannels(const xs_t* s, double hw) {
  int count = 0, i = 0, n = s->np;
  while (i < n) {
    if (!(s->pz[i] < hw)) { i += 1; continue; }
    int start = i;
    while (i < n && s->pz[i] < hw) i += 1;
    if (i - start >= 2) count += 1;
  }
  return count;
}

/* get_equivalent_n.subsection_props (:448-470): conveyance of the points with x_min <= x <= x_max */
static double irr_sub_K(const xs_t* s, double hw, double x_min, double x_max, double n_value) {
  int lo = 0, hi = s->np - 1;
  while (lo < s->np && !(s->px[lo] >= x_min)) ++lo;
  while (hi >= 0 && !(s->px[hi] <= x_max)) --hi;
  if (hi - lo + 1 < 2) return 0.0;
  double A, P, T;
  irr_properties(s->px, s->pz, lo, hi, hw, &A, &P, &T);
  if (A <= 0) return 0.0;
  if (P <= 0) return 0.0;
  return hy_conveyance(A, n_value, A / P);
}

/* IrregularSection.get_equivalent_n (:441-500) */
static double irr_equivalent_n(const xs_t* s, double hw) {
  double left_K = irr_sub_K(s, hw, s->px[0], s->lim_l, s->nl);
  double main_K = irr_sub_K(s, hw, s->lim_l, s->lim_r, s->nm);
  double right_K = irr_sub_K(s, hw, s->lim_r, s->px[s->np - 1], s->nr);
  double A, P, R, T;
  irr_props_full(s, hw, &A, &P, &R, &T);
  if (A <= 0 || P <= 0) return s->nm;
  double R_total = A / P;
  double K_total = pow(pow(left_K, 1.5) + pow(main_K, 1.5) + pow(right_K, 1.5), 2.0 / 3.0);
  if (K_total <= 0.0) return s->nm;
  return (A * pow(R_total, 2.0 / 3.0)) / K_total;
}

/* IrregularSection.conveyance (:502-510), dR_dA (:524-532), dA_dh (:534-539), dK_dA (:512-522) */
static double irr_conveyance(const xs_t* s, double hw) {
  double A, P, R, T;
  irr_props_full(s, hw, &A, &P, &R, &T);
  if (A <= 0.0) return 0.0;
  return hy_conveyance(A, irr_equivalent_n(s, hw), R);
}
static double irr_dR_dA(const xs_t* s, double hw) {
  const double dh = 1e-6;
  double A1, A2, P1, P2, R1, R2, T;
  irr_props_full(s, hw - dh, &A1, &P1, &R1, &T);
  irr_props_full(s, hw + dh, &A2, &P2, &R2, &T);
  if ((A2 - A1) == 0.0) return 0.0;
  return (R2 - R1) / (A2 - A1);
}
static double irr_dA_dh(const xs_t* s, double hw) {
  const double dh = 1e-6;
  double A1, A2, P, R, T;
  irr_props_full(s, hw - dh, &A1, &P, &R, &T);
  irr_props_full(s, hw + dh, &A2, &P, &R, &T);
  return (A2 - A1) / (2 * dh);
}
static double irr_dK_dA(const xs_t* s, double hw) {
  double A, P, R, T;
  irr_props_full(s, hw, &A, &P, &R, &T);
  if (A <= 0.0) return 0.0;
  double n = irr_equivalent_n(s, hw);
  double dR_dA = irr_dR_dA(s, hw);
  return (pow(R, TWO_THIRDS) + A * 2. / 3. * pow(R, M_ONE_THIRD) * dR_dA) / n;
}

/* TrapezoidalSection.properties, cross_section.py:623-679 */
static void xs_properties(const xs_t* s, double hw, double* A, double* P, double* R, double* T) {
  if (s->kind == PR_XS_IRREGULAR) { irr_props_full(s, hw, A, P, R, T); return; }
  double depth = fmax(0.0, hw - s->z);
  if (depth <= 0.0) { *A = *P = *R = *T = 0.0; return; }
  if (s->kind == PR_XS_RECT) {
    *A = s->b * depth;
    *P = s->b + 2.0 * depth;
    *T = s->b;
  } else if (s->kind == PR_XS_TRAPEZOID || depth <= s->hb) {
    *T = s->b + 2.0 * s->m * depth;
    *A = (s->b + *T) / 2.0 * depth;
    *P = s->b + 2.0 * depth * sqrt(1.0 + s->m * s->m);
  } else {
    double dfp = depth - s->hb;
    double A_main = (s->b + s->Tb) / 2.0 * s->hb;
    double P_main = s->b + 2.0 * s->hb * sqrt(1.0 + s->m * s->m);
    double A_left = (s->bl + 0.5 * s->mfp * dfp) * dfp;
    double P_left = s->bl + dfp * sqrt(1.0 + s->mfp * s->mfp);
    double A_right = (s->br + 0.5 * s->mfp * dfp) * dfp;
    double P_right = s->br + dfp * sqrt(1.0 + s->mfp * s->mfp);
    *A = A_main + A_left + A_right;          /* quirk 4: omits the T_bank*dfp column */
    *P = P_main + P_left + P_right;
    *T = s->Wb + 2.0 * s->mfp * dfp;
  }
  *R = (*P > 0.0) ? *A / *P : 0.0;
}

/* hydraulics.conveyance, hydraulics.py:15-26 */
static double hy_conveyance(double A, double n, double R) { return A * pow(R, TWO_THIRDS) / n; }

/* TrapezoidalSection._get_subsection_props, cross_section.py:681-708: (A,P,R) x (left, main, right) */
static void xs_subsections(const xs_t* s, double hw, double sub[3][3]) {
  memset(sub, 0, 9 * sizeof(double));
  double depth = fmax(0.0, hw - s->z);
  if (depth <= 0.0) return;
  if (s->kind != PR_XS_COMPOUND || depth <= s->hb) {
    double A, P, R, T;
    xs_properties(s, hw, &A, &P, &R, &T);
    sub[1][0] = A; sub[1][1] = P; sub[1][2] = R;
    return;
  }
  double dfp = depth - s->hb;
  double A_main = (s->b + s->Tb) / 2.0 * s->hb + s->Tb * dfp;
  double P_main = s->b + 2.0 * s->hb * sqrt(1.0 + s->m * s->m);
  double A_left = (s->bl + 0.5 * s->mfp * dfp) * dfp;
  double P_left = s->bl + dfp * sqrt(1.0 + s->mfp * s->mfp);
  double A_right = (s->br + 0.5 * s->mfp * dfp) * dfp;
  double P_right = s->br + dfp * sqrt(1.0 + s->mfp * s->mfp);
  sub[0][0] = A_left;  sub[0][1] = P_left;  sub[0][2] = (P_left > 0) ? A_left / P_left : 0.0;
  sub[1][0] = A_main;  sub[1][1] = P_main;  sub[1][2] = (P_main > 0) ? A_main / P_main : 0.0;
  sub[2][0] = A_right; sub[2][1] = P_right; sub[2][2] = (P_right > 0) ? A_right / P_right : 0.0;
}

/* TrapezoidalSection.conveyance, cross_section.py:741-754 */
static double xs_conveyance(const xs_t* s, double hw) {
  if (s->kind == PR_XS_IRREGULAR) return irr_conveyance(s, hw);
  if (s->kind != PR_XS_COMPOUND) {
    double A, P, R, T;
    xs_properties(s, hw, &A, &P, &R, &T);
    return hy_conveyance(A, s->nm, R);
  }
  double sub[3][3];
  xs_subsections(s, hw, sub);
  double Kl = hy_conveyance(sub[0][0], s->nl, sub[0][2]);
  double Km = hy_conveyance(sub[1][0], s->nm, sub[1][2]);
  double Kr = hy_conveyance(sub[2][0], s->nr, sub[2][2]);
  return pow(pow(Kl, 1.5) + pow(Km, 1.5) + pow(Kr, 1.5), TWO_THIRDS);
}

/* TrapezoidalSection.get_equivalent_n, cross_section.py:710-739 */
static double xs_equivalent_n(const xs_t* s, double hw) {
  if (s->kind == PR_XS_IRREGULAR) return irr_equivalent_n(s, hw);
  if (s->kind != PR_XS_COMPOUND) return s->nm;
  double K = xs_conveyance(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  if (A <= 0 || R <= 0) return s->nm;
  if (K <= 0.0) return s->nm;
  return (A * pow(R, TWO_THIRDS)) / K;
}

/* TrapezoidalSection.dR_dA, cross_section.py:766-790 */
static double xs_dR_dA(const xs_t* s, double hw) {
  if (s->kind == PR_XS_IRREGULAR) return irr_dR_dA(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  if (P <= 0.0 || T <= 0.0) return 0.0;
  double depth = fmax(0.0, hw - s->z);
  double dP_dh;
  if (s->kind == PR_XS_RECT) dP_dh = 2.0;
  else if (s->kind == PR_XS_TRAPEZOID) dP_dh = 2.0 * sqrt(1.0 + s->m * s->m);
  else if (depth <= s->hb) dP_dh = 2.0 * sqrt(1.0 + s->m * s->m);
  else dP_dh = 2.0 * sqrt(1.0 + s->mfp * s->mfp);
  double dh_dA = 1.0 / T;
  double dP_dA = dP_dh * dh_dA;
  return (P - A * dP_dA) / (P * P);
}

/* TrapezoidalSection.dK_dA + hydraulics.dK_dA_, cross_section.py:756-764, hydraulics.py:28-40 */
static double xs_dK_dA(const xs_t* s, double hw) {
  if (s->kind == PR_XS_IRREGULAR) return irr_dK_dA(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  if (A <= 0.0) return 0.0;
  double n = xs_equivalent_n(s, hw);
  double dR_dA = xs_dR_dA(s, hw);
  return (pow(R, TWO_THIRDS) + A * 2. / 3. * pow(R, M_ONE_THIRD) * dR_dA) / n;
}

/* hydraulics.Sf / dSf_dA / dSf_dQ, hydraulics.py:42-92 (K given) */
static double hy_Sf(double Q, double K) { return Q * fabs(Q) / (K * K); }

/* CrossSection.friction_slope / dSf_dA / dSf_dQ, cross_section.py:114-141 */
/* IrregularSection overrides (:374-439): one wetted sub-channel -> the base-class formulas.  Two or more (a bar or a
 * ridge splitting the flow): every sub-channel is made a section of its own - its submerged points plus the points
 * where it meets the water surface, with the PARENT's roughness limits and values (:385-387) - and the conveyances are
 * combined as K = (sum K_j^1.5)^(2/3) (:389-392), dK/dA = 2/3 (sum K_j^1.5)^(-1/3) sum 1.5 K_j^0.5 dK_j/dA_j (:408-418). */
static int xs_split(const xs_t* s, double hw) { return s->kind == PR_XS_IRREGULAR && irr_subchannels(s, hw) > 1; }

/* numpy.interp(x, [xp0, xp1], [fp0, fp1]) as get_subchannels calls it (:357,361).  On the LEFT edge the reference passes
 * xp = [z[start-1], z[start]], which DEcreases; numpy's interp then takes its "x beyond the last abscissa" exit and
 * returns fp1 - the first submerged point itself, not the intersection: the sub-channel gets a vertical wall there.
 * Reproduced as is (compiled_base.c: arr_interp). */
static double np_interp2(double x, double xp0, double xp1, double fp0, double fp1) {
  if (x != x) return x;
  if (x > xp1) return fp1;
  if (x < xp0) return fp0;
  if (x == xp1) return fp1;
  const double slope = (fp1 - fp0) / (xp1 - xp0);
  double r = slope * (x - xp0) + fp0;
  if (r != r) {                       /* numpy retries from the right end, then gives up */
    r = slope * (x - xp1) + fp1;
    if (r != r && fp0 == fp1) r = fp0;
  }
  return r;
}

/* Sum over the sub-channels of K_j^1.5 and of 1.5 K_j^0.5 dK_j/dA_j (:380-392, :406-414). */
static void irr_split_sums(const xs_t* s, double hw, double* K_sum, double* dK_sum) {
  *K_sum = 0.0; *dK_sum = 0.0;
  const int n = s->np;
  double* buf = (double*)malloc(sizeof(double) * 2 * (n + 2));
  double *xs = buf, *zs = buf + (n + 2);
  int i = 0;
  while (i < n) {
    if (!(s->pz[i] < hw)) { i += 1; continue; }
    const int start = i;
    while (i < n && s->pz[i] < hw) i += 1;
    const int end = i;                                  /* one past the last submerged point */
    if (end - start < 2) continue;
    int m = 0;
    if (start > 0 && s->pz[start - 1] > hw) {
      xs[m] = np_interp2(hw, s->pz[start - 1], s->pz[start], s->px[start - 1], s->px[start]); zs[m] = hw; ++m;
    }
    for (int k = start; k < end; ++k) { xs[m] = s->px[k]; zs[m] = s->pz[k]; ++m; }
    if (end < n && s->pz[end - 1] < hw && s->pz[end] > hw) {
      xs[m] = np_interp2(hw, s->pz[end - 1], s->pz[end], s->px[end - 1], s->px[end]); zs[m] = hw; ++m;
    }
    /* IrregularSection(x=..., z=...) sorts by x (:226-228); the points are increasing already, and numpy's argsort
     * keeps ties in place for arrays this short (insertion sort), which is the vertical wall of the left edge */
    xs_t sub = *s;
    sub.px = xs; sub.pz = zs; sub.np = m;
    sub.z = zs[0];
    for (int k = 1; k < m; ++k) if (zs[k] < sub.z) sub.z = zs[k];
    const double K_j = irr_conveyance(&sub, hw);
    *K_sum += pow(K_j, 1.5);
    *dK_sum += 1.5 * pow(K_j, 0.5) * irr_dK_dA(&sub, hw);
  }
  free(buf);
}

/* CrossSection.dA_dh: top width for the trapezoids (:792-793), central difference for polylines (:534-539) */
static double xs_dA_dh(const xs_t* s, double hw) {
  if (s->kind == PR_XS_IRREGULAR) return irr_dA_dh(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  return T;
}

static double xs_friction_slope(const xs_t* s, double h, double Q) {
  const double hw = h + s->z;
  if (xs_split(s, hw)) {
    double K_sum, dK_sum;
    irr_split_sums(s, hw, &K_sum, &dK_sum);
    return hy_Sf(Q, pow(K_sum, 2.0 / 3.0));
  }
  return hy_Sf(Q, xs_conveyance(s, hw));
}
static double xs_dSf_dA(const xs_t* s, double h, double Q) {
  double hw = h + s->z;
  if (xs_split(s, hw)) {
    double K_sum, dK_sum;
    irr_split_sums(s, hw, &K_sum, &dK_sum);
    const double K_eq = pow(K_sum, 2.0 / 3.0);
    const double dK_dA_eq = (2.0 / 3.0) * pow(K_sum, -1.0 / 3.0) * dK_sum;
    return -2 * hy_Sf(Q, K_eq) * (dK_dA_eq / K_eq);
  }
  double K = xs_conveyance(s, hw);
  double dK = xs_dK_dA(s, hw);
  return -2 * hy_Sf(Q, K) * (dK / K);
}
static double xs_dSf_dQ(const xs_t* s, double h, double Q) {
  const double hw = h + s->z;
  if (xs_split(s, hw)) {
    double K_sum, dK_sum;
    irr_split_sums(s, hw, &K_sum, &dK_sum);
    const double K_eq = pow(K_sum, 2.0 / 3.0);
    return 2 * fabs(Q) / (K_eq * K_eq);
  }
  double K = xs_conveyance(s, hw);
  return 2 * fabs(Q) / (K * K);
}

/* hydraulics.froude_num / dFr_dA / dFr_dQ / darcey_weisbach_f, hydraulics.py:155-229 */
static double hy_froude(double g, double T, double A, double Q) {
  double V = Q / fmax(A, 1e-6);
  double D = A / fmax(T, 1e-6);
  return V / sqrt(g * fmax(D, 1e-6));
}
static double hy_dFr_dA(double g, double T, double A, double Q) {
  double V = Q / A, D = A / T;
  double dV_dA = -Q / (A * A);
  double dD_dA = 1.0 / T;
  return -0.5 * V * pow(g * D, -1.5) * g * dD_dA + dV_dA * pow(g * D, -0.5);
}
static double hy_dFr_dQ(double g, double T, double A) {
  double D = A / T;
  double dV_dQ = 1.0 / A;
  return dV_dQ * pow(g * D, -0.5);
}
static double hy_darcy_f(double g, double n, double R) {
  double C = pow(R, 1.0 / 6.0) / n;
  return 8 * g / (C * C);
}

/* hydraulics.Sc, hydraulics.py:94-117 */
static double hy_Sc(double g, double h, double T, double A, double Q, double n, double R, double rc) {
  double Fr = hy_froude(g, T, A, Q);
  double f = hy_darcy_f(g, n, R);
  double numerator = (2.86 * sqrt(f) + 2.07 * f) * (h * h) * (Fr * Fr);
  double denominator = (0.565 + sqrt(f)) * (rc * rc);
  return numerator / denominator;
}
/* hydraulics.dSc_dA, hydraulics.py:119-137 */
static double hy_dSc_dA(double g, double h, double A, double Q, double n, double R, double rc, double dR_dA, double T) {
  double Fr = hy_froude(g, T, A, Q);
  double C = pow(R, 1.0 / 6.0) / n;
  double f = 8 * g / (C * C);
  double dh_dA = 1. / T;
  double dFr = hy_dFr_dA(g, T, A, Q);
  double df_dA = -(8.0 / 3.0) * g * (n * n) * pow(R, -4.0 / 3.0) * dR_dA;
  double sqrtf = sqrt(f);
  double num = (2.86 * sqrtf + 2.07 * f) * (h * h) * (Fr * Fr);
  double den = (0.565 + sqrtf) * (rc * rc);
  double dnum = (2.86 / (2 * sqrtf) * df_dA + 2.07 * df_dA) * (h * h) * (Fr * Fr)
              + (2.86 * sqrtf + 2.07 * f) * (2 * h * dh_dA * (Fr * Fr) + (h * h) * 2 * Fr * dFr);
  double dden = (1.0 / (2 * sqrtf) * df_dA) * (rc * rc);
  return (dnum * den - num * dden) / (den * den);
}
/* hydraulics.dSc_dQ, hydraulics.py:139-153 */
static double hy_dSc_dQ(double g, double h, double T, double A, double Q, double n, double R, double rc) {
  double Fr = hy_froude(g, T, A, Q);
  double C = pow(R, 1.0 / 6.0) / n;
  double f = 8 * g / (C * C);
  double dFr = hy_dFr_dQ(g, T, A);
  double sqrtf = sqrt(f);
  double num = (2.86 * sqrtf + 2.07 * f) * (h * h) * (Fr * Fr);
  double den = (0.565 + sqrtf) * (rc * rc);
  double dnum = (2.86 * sqrtf + 2.07 * f) * (h * h) * 2 * Fr * dFr;
  double dden = 0.0;
  return (dnum * den - num * dden) / (den * den);
}

/* CrossSection.curvature_slope / dSc_dA / dSc_dQ, cross_section.py:143-175 */
static double xs_curvature_slope(const xs_t* s, double g, double h, double Q) {
  if (s->curv == 0) return 0.0;
  double hw = h + s->z;
  double n = xs_equivalent_n(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  return hy_Sc(g, h, T, A, Q, n, R, 1.0 / s->curv);
}
static double xs_dSc_dA(const xs_t* s, double g, double h, double Q) {
  if (fabs(s->curv) <= 1e-12) return 0.0;
  double hw = h + s->z;
  double n = xs_equivalent_n(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  double dR = xs_dR_dA(s, hw);
  return hy_dSc_dA(g, h, A, Q, n, R, 1.0 / s->curv, dR, T) * xs_dA_dh(s, hw);   /* quirk 7: already x dA_dh (= T for trapezoids) */
}
static double xs_dSc_dQ(const xs_t* s, double g, double h, double Q) {
  if (fabs(s->curv) <= 1e-12) return 0.0;
  double hw = h + s->z;
  double n = xs_equivalent_n(s, hw);
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  return hy_dSc_dQ(g, h, T, A, Q, n, R, 1.0 / s->curv);
}

/* Channel.Se / dSe_dA / dSe_dQ, channel.py:53-105 */
static double ch_Se(const xs_t* s, double g, double h, double Q) {
  return xs_friction_slope(s, h, Q) + xs_curvature_slope(s, g, h, Q);
}
static double ch_dSe_dA(const xs_t* s, double g, double h, double Q) {
  return xs_dSf_dA(s, h, Q) + xs_dSc_dA(s, g, h, Q);
}
static double ch_dSe_dQ(const xs_t* s, double g, double h, double Q) {
  return xs_dSf_dQ(s, h, Q) + xs_dSc_dQ(s, g, h, Q);
}

/* ------------------------------- rating curves --------------------------------------------- */

static double sk_predict(const double c[6], double s, double o) {
  /* sklearn LinearRegression.predict on PolynomialFeatures(2, include_bias=False): X @ coef_ + intercept_ */
  return (c[1] * s + c[2] * o + c[3] * (s * s) + c[4] * (s * o) + c[5] * (o * o)) + c[0];
}

/* RoseiresRatingCurve.total_release, roseires_rating_curve.py:82-85 */
static double ro_total_release(const pr_rating* r, double stage, const double* openings, int sluices) {
  double sluice_releases = sk_predict(r->sluice, stage, r->twl) * sluices;
  double spill = 0;
  for (int j = 0; j < r->n_gates; ++j)
    if (openings[j] > 0) spill = spill + sk_predict(r->spill, stage, openings[j]);
  return spill + sluice_releases + r->q_hydro;
}

/* RoseiresRatingCurve.discharge (smooth=True): alpha_smooth + effective_release, :65-109 */
static double ro_discharge(const pr_rating* r, double stage) {
  double alpha;
  if (stage >= r->stage0 + r->buffer) alpha = 1.0;
  else if (stage <= r->stage0) alpha = 0.0;
  else {
    double s = (stage - r->stage0) / r->buffer;
    alpha = 3 * (s * s) - 2 * (s * s * s);
  }
  double high_Q = ro_total_release(r, stage, r->open_state, r->sluices_open);
  double low_Q = ro_total_release(r, stage, r->closed_state, r->sluices_closed);
  return (1.0 - alpha) * low_Q + alpha * high_Q;
}

/* RoseiresRatingCurve(smooth=False): stateful gate control, roseires_rating_curve.py:65-81,111-142.
 * `time` is the solver's time_level*time_step (boundary.py:95). */
typedef struct { int open, have_prev; double cooldown, prev_time, current_stage; } gate_t;

static void gate_init(gate_t* g, const pr_rating* r) {
  g->open = r->initially_open ? 1 : 0; g->have_prev = 0; g->cooldown = 0; g->prev_time = 0; g->current_stage = r->stage0;
}

static double ro_discharge_gated(const pr_rating* r, gate_t* g, double stage, double time, int update) {
  if (update) {                                   /* gate_control(time), :111-124 */
    if (g->have_prev) { double c = g->cooldown - (time - g->prev_time); g->cooldown = c > 0 ? c : 0; }
    g->prev_time = time; g->have_prev = 1;
    if (!(g->cooldown > 0)) {
      if (g->current_stage >= r->stage0 + 0.5 && !g->open) { g->cooldown = r->max_cooldown; g->open = 1; }
      else if (g->current_stage <= r->stage0 - 1 && g->open) { g->cooldown = r->max_cooldown; g->open = 0; }
    }
  }
  double Q = g->open ? ro_total_release(r, stage, r->open_state, r->sluices_open)
                     : ro_total_release(r, stage, r->closed_state, r->sluices_closed);
  if (update) g->current_stage = stage;
  return Q;
}

static double polyval(const double* c, int n, double x) {   /* numpy.polynomial.polynomial.polyval (Horner) */
  double c0 = c[n - 1];
  for (int i = 2; i <= n; ++i) c0 = c[n - i] + c0 * x;
  return c0;
}

/* RatingCurve.discharge, rating_curve.py:32-63 ; RoseiresRatingCurve.discharge */
double pr_oracle_rating_discharge(const pr_rating* r, double stage) {
  switch (r->type) {
    case PR_RC_POLY2: { double x = stage + r->stage_shift; return r->a * (x * x) + r->b * x + r->c; }
    case PR_RC_POWER: { double x = stage + r->stage_shift; return r->a * pow(x, r->b); }
    case PR_RC_POLYNOMIAL: return polyval(r->coef, r->n_coef, r->off + r->scl * stage);  /* NB: no shift (:48-49) */
    case PR_RC_ROSEIRES: return ro_discharge(r, stage);
    default: return NAN;
  }
}

/* RatingCurve.dQ_dz, rating_curve.py:132-147 ; RoseiresRatingCurve.dQ_dz, roseires_rating_curve.py:202-208 */
double pr_oracle_rating_dQdz(const pr_rating* r, double stage) {
  switch (r->type) {
    case PR_RC_POLY2: { double Y = stage + r->stage_shift; return r->a * 2 * Y + r->b; }
    case PR_RC_POWER: { double Y = stage + r->stage_shift; return r->a * r->b * pow(Y, r->b - 1); }
    case PR_RC_POLYNOMIAL: { double Y = stage + r->stage_shift; return polyval(r->dcoef, r->n_coef - 1, r->off + r->scl * Y); }
    case PR_RC_ROSEIRES: {
      double fp = ro_discharge(r, stage + r->dY), fm = ro_discharge(r, stage - r->dY);
      return (fp - fm) / (2 * r->dY);
    }
    default: return NAN;
  }
}

/* ------------------------------- scipy.optimize.brentq ------------------------------------- */
/* Brent's method as implemented by scipy/optimize/Zeros/brentq.c (xtol=2e-12, rtol=4*eps, maxiter=100). */
typedef double (*fn1)(double, void*);
static double brentq(fn1 f, void* ctx, double xa, double xb, int* err) {
  const double xtol = 2e-12, rtol = 8.881784197001252e-16;
  double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis, delta, stry, dpre, dblk;
  *err = 0;
  fpre = f(xpre, ctx);
  fcur = f(xcur, ctx);
  if (fpre == 0) return xpre;
  if (fcur == 0) return xcur;
  if (signbit(fpre) == signbit(fcur)) { *err = 1; return 0.; }
  for (int i = 0; i < 100; ++i) {
    if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
      xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
    }
    if (fabs(fblk) < fabs(fcur)) {
      xpre = xcur; xcur = xblk; xblk = xpre;
      fpre = fcur; fcur = fblk; fblk = fpre;
    }
    delta = (xtol + rtol * fabs(xcur)) / 2;
    sbis = (xblk - xcur) / 2;
    if (fcur == 0 || fabs(sbis) < delta) return xcur;
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      if (xpre == xblk) {
        stry = -fcur * (xcur - xpre) / (fcur - fpre);
      } else {
        dpre = (fpre - fcur) / (xpre - xcur);
        dblk = (fblk - fcur) / (xblk - xcur);
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
      }
      double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
      if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
      else { spre = sbis; scur = sbis; }
    } else { spre = sbis; scur = sbis; }
    xpre = xcur; fpre = fcur;
    if (fabs(scur) > delta) xcur += scur;
    else xcur += (sbis > 0 ? delta : -delta);
    fcur = f(xcur, ctx);
  }
  *err = 2;
  return xcur;
}

double pr_oracle_brentq_poly(const double* c, int n, double xa, double xb, int* err);
typedef struct { const double* c; int n; } polyctx;
static double poly_f(double x, void* p) { polyctx* q = (polyctx*)p; return polyval(q->c, q->n, x); }
/* exported so tests can check the brentq restatement against scipy.optimize.brentq bit for bit */
double pr_oracle_brentq_poly(const double* c, int n, double xa, double xb, int* err) {
  polyctx q = {c, n};
  return brentq(poly_f, &q, xa, xb, err);
}

static double np_interp(double x, const double* xp, const double* fp, int n);

/* LumpedStorage.mass_balance, lumped_storage.py:24-35 */
/* LumpedStorage.area_at, lumped_storage.py:155-160 */
static double st_area_at(const pr_bc* b, double stage) {
  if (b->storage_curve_len <= 0) return b->storage_area;
  return b->storage_alpha * np_interp(stage + b->storage_beta, b->storage_curve_stage, b->storage_curve_area, b->storage_curve_len);
}

/* LumpedStorage.net_vol_change, lumped_storage.py:171-179 */
static double st_net_vol_change(const pr_bc* b, double Y1, double Y2) {
  if (b->storage_curve_len <= 0) return (Y2 - Y1) * b->storage_area;
  double step = INFINITY;
  for (int i = 1; i < b->storage_curve_len; ++i) step = fmin(step, fabs(b->storage_curve_stage[i] - b->storage_curve_stage[i - 1]));
  int n = (int)(fabs(Y2 - Y1) / step);
  if (n > 2) {
    /* ys = np.linspace(Y1, Y2, n); np.trapezoid([area_at(y) for y in ys], ys) */
    double dstep = (Y2 - Y1) / (n - 1), sum = 0.0;
    double y_prev = Y1, a_prev = st_area_at(b, Y1);
    for (int i = 1; i < n; ++i) {
      double y = (i == n - 1) ? Y2 : i * dstep + Y1;
      double a = st_area_at(b, y);
      sum += (y - y_prev) * (a + a_prev) / 2.0;
      y_prev = y; a_prev = a;
    }
    return sum;
  }
  return 0.5 * (st_area_at(b, Y2) + st_area_at(b, Y1)) * (Y2 - Y1);
}

typedef struct { const pr_bc* b; double Y_old, vol_in, duration; } mbctx;
static double mb_f(double Y_new, void* p) {
  mbctx* c = (mbctx*)p;
  double Q_out = 0.0;
  if (c->b->storage_outflow.type != PR_RC_NONE)
    Q_out = 0.5 * (pr_oracle_rating_discharge(&c->b->storage_outflow, c->Y_old) + pr_oracle_rating_discharge(&c->b->storage_outflow, Y_new));
  double target_vol = c->vol_in - Q_out * c->duration;
  return st_net_vol_change(c->b, c->Y_old, Y_new) - target_vol;
}
static double storage_mass_balance(const pr_bc* b, double vol_in, double Y_old, double duration, int* err) {
  mbctx c = {b, Y_old, vol_in, duration};
  double Y = brentq(mb_f, &c, b->storage_ymin, b->storage_ymax, err);
  if (Y < b->storage_min_stage) Y = b->storage_min_stage;
  return Y;
}

/* LumpedStorage.energy_loss / dhl_dA / dhl_dQ, lumped_storage.py:47-143 (A_str is never passed: no expansion term) */
static double st_energy_loss(const pr_bc* b, double g, double A, double Q, double n, double R) {
  if (!b->storage_capture_losses) return 0;
  double K = hy_conveyance(A, n, R);
  double hf = hy_Sf(Q, K) * b->storage_reservoir_length;
  double V = Q / A;
  double h_emp = b->storage_Kq * (V * V) / (2 * g);
  return hf + 0 + h_emp;
}
static double st_dhl_dA(const pr_bc* b, double g, double A, double Q, double n, double R, double dR_dA) {
  if (!b->storage_capture_losses) return 0;
  double K = hy_conveyance(A, n, R);
  double dK = (pow(R, TWO_THIRDS) + A * 2. / 3. * pow(R, M_ONE_THIRD) * dR_dA) / n;
  double dhf = (-2 * hy_Sf(Q, K) * (dK / K)) * b->storage_reservoir_length;
  double V = Q / A, dV_dA = -Q / (A * A);
  double demp = b->storage_Kq * 2 * V * dV_dA / (2 * g);
  return dhf + 0 + demp;
}
static double st_dhl_dQ(const pr_bc* b, double g, double A, double Q, double n, double R) {
  if (!b->storage_capture_losses) return 0;
  double K = hy_conveyance(A, n, R);
  double dhf = (2 * fabs(Q) / (K * K)) * b->storage_reservoir_length;
  double V = Q / A, dV_dQ = 1. / A;
  double demp = b->storage_Kq * 2 * V * dV_dQ / (2 * g);
  return dhf + 0 + demp;
}

/* ------------------------------- boundaries ------------------------------------------------ */

typedef struct {
  const pr_bc* bc;
  const xs_t* xs;
  int member, level;
  double dt, g;
  double* stage_record;   /* storage: [levels] of this member, entry k = stage recorded for level k */
  gate_t* gate;           /* Roseires gate state of this member (gate_control = 1), persists over the whole run */
} bc_ctx;

static const pr_rating* bc_rating(const bc_ctx* c) {
  return c->bc->member_ratings ? &c->bc->member_ratings[c->member] : &c->bc->rating;
}
static int bc_gated(const bc_ctx* c) {
  const pr_rating* r = bc_rating(c);
  return r->type == PR_RC_ROSEIRES && r->gate_control && c->gate;
}

static double bc_series(const bc_ctx* c) {
  return c->bc->series[(int64_t)c->member * c->bc->series_member_stride + c->level];
}

/* Boundary.condition_residual, boundary.py:56-141 */
static double bc_residual(const bc_ctx* c, double depth, double flow, double vol_in, int* err) {
  const pr_bc* b = c->bc;
  double hw = c->xs->z + depth;
  switch (b->type) {
    case PR_BC_FLOW_HYDROGRAPH: return flow - bc_series(c);
    case PR_BC_NORMAL_DEPTH: {
      double K = xs_conveyance(c->xs, hw);
      double Qn = K * pow(fabs(b->bed_slope), 0.5);        /* hydraulics.normal_flow, :4-13 */
      if (b->bed_slope < 0) Qn = -Qn;
      return flow - Qn;
    }
    case PR_BC_RATING_CURVE:
      if (bc_gated(c)) return flow - ro_discharge_gated(bc_rating(c), c->gate, b->bed_level + depth, c->level * c->dt, 1);
      return flow - pr_oracle_rating_discharge(bc_rating(c), b->bed_level + depth);
    case PR_BC_FIXED_DEPTH: return depth - b->fixed_depth;
    case PR_BC_STAGE_HYDROGRAPH: return depth - (bc_series(c) - b->bed_level);
    case PR_BC_FIXED_DEPTH_STORAGE: {
      int k = c->level;                                     /* time // duration */
      double Y_old = (k == 1) ? depth + b->bed_level : c->stage_record[k - 1];   /* quirk 9 */
      double reservoir_stage = storage_mass_balance(b, vol_in, Y_old, c->dt, err);
      double A, P, R, T;
      xs_properties(c->xs, hw, &A, &P, &R, &T);
      double head_loss = st_energy_loss(b, c->g, A, flow, xs_equivalent_n(c->xs, hw), R);   /* boundary.py:117-122 */
      double interface_stage = reservoir_stage + head_loss;
      c->stage_record[k] = reservoir_stage;                  /* boundary.py:126-131 (overwritten each evaluation) */
      return depth - (interface_stage - b->bed_level);
    }
    default: *err = 3; return NAN;
  }
}

/* Boundary.df_dh, boundary.py:143-187 */
static double bc_df_dh(const bc_ctx* c, double depth, double flow) {
  const pr_bc* b = c->bc;
  if (b->type == PR_BC_FLOW_HYDROGRAPH) return 0;
  double hw = depth + b->bed_level;
  double A, P, R, T;
  xs_properties(c->xs, hw, &A, &P, &R, &T);
  double dA_dh = xs_dA_dh(c->xs, hw);
  switch (b->type) {
    case PR_BC_FIXED_DEPTH: return 1 - 0 * dA_dh;
    case PR_BC_FIXED_DEPTH_STORAGE: {                        /* boundary.py:168-177 */
      double dhl_dA = st_dhl_dA(b, c->g, A, flow, xs_equivalent_n(c->xs, hw), R, xs_dR_dA(c->xs, hw));
      return 1 - dhl_dA * dA_dh;
    }
    case PR_BC_NORMAL_DEPTH: {
      double dQ = xs_dK_dA(c->xs, hw) * pow(fabs(b->bed_slope), 0.5);   /* hydraulics.dQn_dA, :206-215 */
      if (b->bed_slope < 0) dQ = -dQ;
      return 0 - dQ * dA_dh;
    }
    case PR_BC_RATING_CURVE: {
      const pr_rating* r = bc_rating(c);
      double stage = b->bed_level + depth;
      if (bc_gated(c)) {                          /* dQ_dz with update_stage = update_gate_state = False, :202-208 */
        double fp = ro_discharge_gated(r, c->gate, stage + r->dY, 0, 0), fm = ro_discharge_gated(r, c->gate, stage - r->dY, 0, 0);
        return 0 - (fp - fm) / (2 * r->dY);
      }
      return 0 - pr_oracle_rating_dQdz(r, stage);
    }
    case PR_BC_STAGE_HYDROGRAPH: return 1;
    default: return NAN;
  }
}

/* Boundary.df_dQ, boundary.py:189-242 */
static double bc_df_dQ(const bc_ctx* c, double depth, double flow, double vol_in, int* err) {
  const pr_bc* b = c->bc;
  switch (b->type) {
    case PR_BC_FLOW_HYDROGRAPH: case PR_BC_NORMAL_DEPTH: case PR_BC_RATING_CURVE: return 1;
    case PR_BC_FIXED_DEPTH: return 0;
    case PR_BC_STAGE_HYDROGRAPH: return 0;
    case PR_BC_FIXED_DEPTH_STORAGE: {
      int k = c->level;
      /* NB: the residual of this iteration has already overwritten stage_record[k]; [k-1] is untouched */
      double Y_old = (k == 1) ? depth + b->bed_level : c->stage_record[k - 1];
      double Y_new = storage_mass_balance(b, vol_in, Y_old, c->dt, err);       /* dY_new_dvol_in, :37-45 */
      double dY_new_dvol = (Y_new <= b->storage_min_stage) ? 0.0 : 1 / st_area_at(b, Y_new);
      double dvol_dQ = 0.5 * c->dt;
      double hw = depth + b->bed_level;
      double A, P, R, T;
      xs_properties(c->xs, hw, &A, &P, &R, &T);
      double dhl_dQ = st_dhl_dQ(b, c->g, A, flow, xs_equivalent_n(c->xs, hw), R);    /* boundary.py:232-235 */
      return 0 - (dY_new_dvol * dvol_dQ + dhl_dQ);
    }
    default: return NAN;
  }
}

/* ------------------------------- banded solve (stands in for SuperLU) ---------------------- */
/* Gaussian elimination with partial pivoting on a band with kl = ku = 2 (fill to ku = 4).
 * B[r][c - r + 4] holds A[r][c].  Returns 0, or -1 if a zero pivot column is met. */
static int banded_solve(int n, double (*B)[9], double* rhs, double* x) {
  for (int j = 0; j < n; ++j) {
    int p = j;
    double best = fabs(B[j][4]);
    int rmax = (j + 2 < n - 1) ? j + 2 : n - 1;
    for (int r = j + 1; r <= rmax; ++r) {
      double v = fabs(B[r][j - r + 4]);
      if (v > best) { best = v; p = r; }
    }
    if (best == 0.0 || best != best) return -1;
    int cmax = (j + 4 < n - 1) ? j + 4 : n - 1;
    if (p != j) {
      for (int c = j; c <= cmax; ++c) {
        double t = B[j][c - j + 4]; B[j][c - j + 4] = B[p][c - p + 4]; B[p][c - p + 4] = t;
      }
      double t = rhs[j]; rhs[j] = rhs[p]; rhs[p] = t;
    }
    for (int r = j + 1; r <= rmax; ++r) {
      double f = B[r][j - r + 4] / B[j][4];
      if (f != 0.0) {
        for (int c = j + 1; c <= cmax; ++c) B[r][c - r + 4] -= f * B[j][c - j + 4];
        rhs[r] -= f * rhs[j];
      }
      B[r][j - r + 4] = 0.0;
    }
  }
  for (int j = n - 1; j >= 0; --j) {
    double s = rhs[j];
    int cmax = (j + 4 < n - 1) ? j + 4 : n - 1;
    for (int c = j + 1; c <= cmax; ++c) s -= B[j][c - j + 4] * x[c];
    x[j] = s / B[j][4];
  }
  return 0;
}

/* ------------------------------- the scheme ------------------------------------------------ */

typedef struct {
  int N;
  double theta, dt, dx, g;
  const xs_t* xs;
  const double* h0; const double* q0;   /* level k-1 (stored) */
  const double* h1; const double* q1;   /* level k   (current iterate) */
} sch_t;

/* PreissmannSolver.time_diff / spatial_diff / cell_avg, preissmann.py:899-910 */
static double time_diff(const sch_t* s, double k1_i1, double k1_i, double k_i1, double k_i) {
  return (k1_i1 + k1_i - k_i1 - k_i) / (2 * s->dt);
}
static double spatial_diff(const sch_t* s, double k1_i1, double k1_i, double k_i1, double k_i) {
  double dx_k1 = (k1_i1 - k1_i) / s->dx;
  double dx_k = (k_i1 - k_i) / s->dx;
  return s->theta * dx_k1 + (1 - s->theta) * dx_k;
}
static double cell_avg(const sch_t* s, double k1_i1, double k1_i, double k_i1, double k_i) {
  double k1 = 0.5 * s->theta * (k1_i1 + k1_i);
  double k2 = 0.5 * (1 - s->theta) * (k_i1 + k_i);
  return k1 + k2;
}

/* Solver.area_at / water_level_at / Se_at / dA_dh, solver.py:271-296 */
static double area_at(const sch_t* s, int lvl, int i) {
  double h = lvl ? s->h1[i] : s->h0[i];
  double A, P, R, T;
  xs_properties(&s->xs[i], s->xs[i].z + h, &A, &P, &R, &T);
  return A;
}
static double topw_at(const sch_t* s, int i) {      /* Solver.dA_dh (solver.py:295-296) */
  return xs_dA_dh(&s->xs[i], s->xs[i].z + s->h1[i]);
}
static double wl_at(const sch_t* s, int lvl, int i) { return s->xs[i].z + (lvl ? s->h1[i] : s->h0[i]); }
static double flow_at(const sch_t* s, int lvl, int i) { return lvl ? s->q1[i] : s->q0[i]; }
static double Se_at(const sch_t* s, int lvl, int i) {
  return ch_Se(&s->xs[i], s->g, lvl ? s->h1[i] : s->h0[i], flow_at(s, lvl, i));
}

/* continuity_residual, preissmann.py:220-249 */
static double continuity_residual(const sch_t* s, int i) {
  double dA_dt = time_diff(s, area_at(s, 1, i + 1), area_at(s, 1, i), area_at(s, 0, i + 1), area_at(s, 0, i));
  double dQ_dx = spatial_diff(s, flow_at(s, 1, i + 1), flow_at(s, 1, i), flow_at(s, 0, i + 1), flow_at(s, 0, i));
  return dA_dt + dQ_dx;
}

static void cell_terms(const sch_t* s, int i, double* avg_A, double* dY_dx, double* avg_Se) {
  *avg_A = cell_avg(s, area_at(s, 1, i + 1), area_at(s, 1, i), area_at(s, 0, i + 1), area_at(s, 0, i));
  *dY_dx = spatial_diff(s, wl_at(s, 1, i + 1), wl_at(s, 1, i), wl_at(s, 0, i + 1), wl_at(s, 0, i));
  *avg_Se = cell_avg(s, Se_at(s, 1, i + 1), Se_at(s, 1, i), Se_at(s, 0, i + 1), Se_at(s, 0, i));
}

/* momentum_residual, preissmann.py:251-301 */
static double momentum_residual(const sch_t* s, int i) {
  double dQ_dt = time_diff(s, flow_at(s, 1, i + 1), flow_at(s, 1, i), flow_at(s, 0, i + 1), flow_at(s, 0, i));
  double f11 = flow_at(s, 1, i + 1) * flow_at(s, 1, i + 1) / area_at(s, 1, i + 1);
  double f10 = flow_at(s, 1, i) * flow_at(s, 1, i) / area_at(s, 1, i);
  double f01 = flow_at(s, 0, i + 1) * flow_at(s, 0, i + 1) / area_at(s, 0, i + 1);
  double f00 = flow_at(s, 0, i) * flow_at(s, 0, i) / area_at(s, 0, i);
  double dQ2A_dx = spatial_diff(s, f11, f10, f01, f00);
  double avg_A, dY_dx, avg_Se;
  cell_terms(s, i, &avg_A, &dY_dx, &avg_Se);
  return dQ_dt + dQ2A_dx + s->g * avg_A * (dY_dx + avg_Se);
}

/* dM_dh_i / dM_dh_ip1, preissmann.py:496-612 ; node = i + side */
static double dM_dh(const sch_t* s, int i, int side) {
  int nd = i + side;
  double A = area_at(s, 1, nd), Q = s->q1[nd], h = s->h1[nd];
  double dA_dh = topw_at(s, nd);
  double dSe_dA = ch_dSe_dA(&s->xs[nd], s->g, h, Q);
  double avg_A, dY_dx, avg_Se;
  cell_terms(s, i, &avg_A, &dY_dx, &avg_Se);
  double unit_sd = side ? spatial_diff(s, 1, 0, 0, 0) : spatial_diff(s, 0, 1, 0, 0);
  double unit_ca = side ? cell_avg(s, 1, 0, 0, 0) : cell_avg(s, 0, 1, 0, 0);
  double d_dQdt_dA = 0;
  double d_dQ2Adx_dA = -unit_sd * ((Q / A) * (Q / A));
  double d_avgA_dA = unit_ca;
  double d_dYdx_dh = unit_sd;
  double d_avgSe_dA = unit_ca * dSe_dA;
  return d_dQdt_dA * dA_dh + d_dQ2Adx_dA * dA_dh +
         s->g * (avg_A * (d_dYdx_dh + d_avgSe_dA * dA_dh) + d_avgA_dA * dA_dh * (dY_dx + avg_Se));
}

/* dM_dQ_i / dM_dQ_ip1, preissmann.py:619-733 */
static double dM_dQ(const sch_t* s, int i, int side) {
  int nd = i + side;
  double A = area_at(s, 1, nd), Q = s->q1[nd], h = s->h1[nd];
  double dSe_dQ = ch_dSe_dQ(&s->xs[nd], s->g, h, Q);
  double avg_A, dY_dx, avg_Se;
  cell_terms(s, i, &avg_A, &dY_dx, &avg_Se);
  double unit_td = side ? time_diff(s, 1, 0, 0, 0) : time_diff(s, 0, 1, 0, 0);
  double unit_sd = side ? spatial_diff(s, 1, 0, 0, 0) : spatial_diff(s, 0, 1, 0, 0);
  double unit_ca = side ? cell_avg(s, 1, 0, 0, 0) : cell_avg(s, 0, 1, 0, 0);
  double d_dQ2Adx_dQ = unit_sd * 2 * Q / A;
  double d_avgSe_dQ = unit_ca * dSe_dQ;
  return unit_td + d_dQ2Adx_dQ + s->g * (avg_A * (0 + d_avgSe_dQ) + 0 * (dY_dx + avg_Se));
}

/* compute_residual_vector + compute_jacobian_data, preissmann.py:61-81, 322-344.
 * R has 2N entries; J has 8N-4 entries in the reference's row-major order. */
static int assemble(const sch_t* s, bc_ctx* up, bc_ctx* dn, double* R, double* J) {
  int N = s->N, err = 0;
  double vol_dn = 0.5 * (s->q0[N - 1] + s->q1[N - 1]) * s->dt;            /* preissmann.py:314 */
  R[0] = bc_residual(up, s->h1[0], s->q1[0], 0.0, &err);
  R[2 * N - 1] = bc_residual(dn, s->h1[N - 1], s->q1[N - 1], vol_dn, &err);
  for (int i = 0; i < N - 1; ++i) {
    R[1 + 2 * i] = continuity_residual(s, i);
    R[2 + 2 * i] = momentum_residual(s, i);
  }
  double vol_up = 0.5 * (s->q1[0] + s->q0[0]);                              /* preissmann.py:391 */
  int p = 0;
  J[p++] = bc_df_dh(up, s->h1[0], s->q1[0]);
  J[p++] = bc_df_dQ(up, s->h1[0], s->q1[0], vol_up, &err);
  for (int i = 0; i < N - 1; ++i) {
    J[p++] = time_diff(s, 0, 1, 0, 0) * topw_at(s, i);       /* dC_dh_i   :431-447 */
    J[p++] = spatial_diff(s, 0, 1, 0, 0);                    /* dC_dQ_i   :476-494 */
    J[p++] = time_diff(s, 1, 0, 0, 0) * topw_at(s, i + 1);   /* dC_dh_ip1 :407-422 */
    J[p++] = spatial_diff(s, 1, 0, 0, 0);                    /* dC_dQ_ip1 :456-474 */
    J[p++] = dM_dh(s, i, 0);
    J[p++] = dM_dQ(s, i, 0);
    J[p++] = dM_dh(s, i, 1);
    J[p++] = dM_dQ(s, i, 1);
  }
  J[p++] = bc_df_dh(dn, s->h1[N - 1], s->q1[N - 1]);
  J[p++] = bc_df_dQ(dn, s->h1[N - 1], s->q1[N - 1], vol_dn, &err);
  return err;
}

/* One reference Newton iteration for tests: given the stored level (h0,q0) and iterate (h1,q1) of one
 * member at `level`, return R[2N], J[8N-4] and delta[2N]. */
int pr_oracle_newton_step(const pr_config* cfg, const pr_geom* geom, const pr_bc* up_bc, const pr_bc* dn_bc,
                          int member, int level, const double* h0, const double* q0, const double* h1,
                          const double* q1, double* stage_record, double* R, double* J, double* delta) {
  int N = cfg->n_nodes, n2 = 2 * N;
  xs_t* xs = (xs_t*)malloc(sizeof(xs_t) * N);
  for (int i = 0; i < N; ++i) xs[i] = xs_load(geom, i, member);
  sch_t s = {N, cfg->theta, cfg->dt, cfg->dx, cfg->g, xs, h0, q0, h1, q1};
  bc_ctx up = {up_bc, &xs[0], member, level, cfg->dt, cfg->g, stage_record, NULL};
  bc_ctx dn = {dn_bc, &xs[N - 1], member, level, cfg->dt, cfg->g, stage_record, NULL};
  int err = assemble(&s, &up, &dn, R, J);
  double(*B)[9] = (double(*)[9])calloc(n2, sizeof(double[9]));
  double* rhs = (double*)malloc(sizeof(double) * n2);
  int p = 0;
  B[0][4] = J[p++]; B[0][5] = J[p++];
  for (int i = 0; i < N - 1; ++i)
    for (int e = 0; e < 2; ++e) {
      int r = 1 + 2 * i + e;
      for (int c = 2 * i; c < 2 * i + 4; ++c) B[r][c - r + 4] = J[p++];
    }
  B[n2 - 1][3] = J[p++]; B[n2 - 1][4] = J[p++];
  for (int r = 0; r < n2; ++r) rhs[r] = -R[r];
  if (banded_solve(n2, B, rhs, delta) != 0) err = err ? err : 4;
  free(B); free(rhs); free(xs);
  return err;
}

/* Diagnostic trace for the near-tie analysis of the convergence test (tools/oracle_grid.py): when set, receives
 * [M][levels-1] ||R||_2 of the iteration BEFORE the accepted one (NaN when the first iterate was accepted). */
static double* g_trace_prev_error = NULL;
void pr_oracle_trace_prev_error(double* buf) { g_trace_prev_error = buf; }

/* PreissmannSolver.run, preissmann.py:101-163, for every member. */
int pr_oracle_run(const pr_config* cfg, const pr_geom* geom, const pr_bc* up_bc, const pr_bc* dn_bc,
                  const pr_state* ic, const pr_outputs* out) {
  int N = cfg->n_nodes, L = cfg->n_levels, M = cfg->n_members, n2 = 2 * N;
  if (N < 2 || L < 1 || M < 1) return PR_ERR_ARG;
  xs_t* xs = (xs_t*)malloc(sizeof(xs_t) * N);
  double* depth = (double*)malloc(sizeof(double) * (size_t)L * N);
  double* flow = (double*)malloc(sizeof(double) * (size_t)L * N);
  double* x = (double*)malloc(sizeof(double) * n2);
  double* h1 = (double*)malloc(sizeof(double) * N);
  double* q1 = (double*)malloc(sizeof(double) * N);
  double* R = (double*)malloc(sizeof(double) * n2);
  double* J = (double*)malloc(sizeof(double) * (8 * N));
  double* delta = (double*)malloc(sizeof(double) * n2);
  double* rhs = (double*)malloc(sizeof(double) * n2);
  double* stage = (double*)malloc(sizeof(double) * L);
  double(*B)[9] = (double(*)[9])malloc(sizeof(double[9]) * n2);

  for (int m = 0; m < M; ++m) {
    for (int i = 0; i < N; ++i) xs[i] = xs_load(geom, i, m);
    const double* h_ic = ic->depth + (int64_t)m * ic->member_stride;
    const double* q_ic = ic->flow + (int64_t)m * ic->member_stride;
    for (int i = 0; i < N; ++i) {                          /* initialize_t0, solver.py:61-63, preissmann.py:59 */
      depth[i] = h_ic[i]; flow[i] = q_ic[i];
      x[2 * i] = h_ic[i]; x[2 * i + 1] = q_ic[i];
    }
    for (int k = 0; k < L; ++k) stage[k] = NAN;
    stage[0] = xs[N - 1].z + h_ic[N - 1];                  /* solver.py:101-108 */
    if (dn_bc->type == PR_BC_FIXED_DEPTH_STORAGE && dn_bc->storage_capture_losses) {
      double A, P, R, T, Y = stage[0];
      xs_properties(&xs[N - 1], Y, &A, &P, &R, &T);
      stage[0] = Y - st_energy_loss(dn_bc, cfg->g, A, q_ic[N - 1], xs_equivalent_n(&xs[N - 1], Y), R);
    }
    int status = PR_STATUS_OK, fail_level = 0;
    gate_t gate_up, gate_dn;
    gate_init(&gate_up, up_bc->member_ratings ? &up_bc->member_ratings[m] : &up_bc->rating);
    gate_init(&gate_dn, dn_bc->member_ratings ? &dn_bc->member_ratings[m] : &dn_bc->rating);
    for (int k = 1; k < L && status == PR_STATUS_OK; ++k) {
      double* hk0 = depth + (size_t)(k - 1) * N; double* qk0 = flow + (size_t)(k - 1) * N;
      double* hk1 = depth + (size_t)k * N;       double* qk1 = flow + (size_t)k * N;
      sch_t s = {N, cfg->theta, cfg->dt, cfg->dx, cfg->g, xs, hk0, qk0, hk1, qk1};
      bc_ctx up = {up_bc, &xs[0], m, k, cfg->dt, cfg->g, stage, &gate_up};
      bc_ctx dn = {dn_bc, &xs[N - 1], m, k, cfg->dt, cfg->g, stage, &gate_dn};
      int iteration = 0, converged = 0;
      double error = NAN, prev_error = NAN;
      while (!converged) {
        iteration += 1;
        prev_error = error;
        if (iteration - 1 >= cfg->max_iter) { status = PR_STATUS_MAX_ITER; iteration -= 1; break; }
        for (int i = 0; i < N; ++i) { hk1[i] = x[2 * i]; qk1[i] = x[2 * i + 1]; }   /* update_guesses */
        int err = assemble(&s, &up, &dn, R, J);
        memset(B, 0, sizeof(double[9]) * n2);
        int p = 0;
        B[0][4] = J[p++]; B[0][5] = J[p++];
        for (int i = 0; i < N - 1; ++i)
          for (int e = 0; e < 2; ++e) {
            int r = 1 + 2 * i + e;
            for (int c = 2 * i; c < 2 * i + 4; ++c) B[r][c - r + 4] = J[p++];
          }
        B[n2 - 1][3] = J[p++]; B[n2 - 1][4] = J[p++];
        for (int r = 0; r < n2; ++r) rhs[r] = -R[r];
        if (err || banded_solve(n2, B, rhs, delta) != 0) { status = PR_STATUS_NAN; break; }
        for (int r = 0; r < n2; ++r) x[r] += delta[r];
        double ss = 0;
        for (int r = 0; r < n2; ++r) ss += R[r] * R[r];
        error = pow(ss, 0.5);                               /* utility.euclidean_norm, :20-22 */
        if (error < cfg->tol) converged = 1;
      }
      if (status == PR_STATUS_MAX_ITER && !(error == error)) status = PR_STATUS_NAN;
      if (out->iters) out->iters[(size_t)m * (L - 1) + (k - 1)] = iteration;
      if (out->final_error) out->final_error[(size_t)m * (L - 1) + (k - 1)] = error;
      if (g_trace_prev_error) g_trace_prev_error[(size_t)m * (L - 1) + (k - 1)] = prev_error;
      if (status != PR_STATUS_OK) {
        fail_level = k;
        for (int kk = k; kk < L; ++kk) {
          for (int i = 0; i < N; ++i) { depth[(size_t)kk * N + i] = NAN; flow[(size_t)kk * N + i] = NAN; }
          if (kk > k && out->iters) out->iters[(size_t)m * (L - 1) + (kk - 1)] = 0;
          if (kk > k && out->final_error) out->final_error[(size_t)m * (L - 1) + (kk - 1)] = NAN;
        }
      }
    }
    if (out->status) out->status[m] = status;
    if (out->fail_level) out->fail_level[m] = fail_level;
    if (out->storage_stage) memcpy(out->storage_stage + (size_t)m * L, stage, sizeof(double) * L);
    if (cfg->out_mode == PR_OUT_FULL) {
      if (out->depth) memcpy(out->depth + (size_t)m * L * N, depth, sizeof(double) * (size_t)L * N);
      if (out->flow) memcpy(out->flow + (size_t)m * L * N, flow, sizeof(double) * (size_t)L * N);
    } else {
      for (int k = 0; k < L; ++k) {
        if (out->depth) out->depth[(size_t)m * L + k] = depth[(size_t)k * N];
        if (out->flow) out->flow[(size_t)m * L + k] = flow[(size_t)k * N];
      }
    }
  }
  free(xs); free(depth); free(flow); free(x); free(h1); free(q1); free(R); free(J); free(delta); free(rhs);
  free(stage); free(B);
  return PR_OK;
}

/* Channel._gvh_conditions, channel.py:307-378.  status: PR_STATUS_SUPERCRITICAL mirrors the RuntimeError. */
static double gvf_dh_dx(const xs_t* xs, double g, double Q, double S0, double h_in, int node, int* status) {
  const xs_t* s = &xs[node];
  double hw = h_in + s->z;
  double A, P, R, T;
  xs_properties(s, hw, &A, &P, &R, &T);
  if (T < 1e-6 || A < 1e-6) return 0.0;
  double Fr = hy_froude(g, T, A, Q);
  if (Fr > 1.0) { *status = PR_STATUS_SUPERCRITICAL; return NAN; }
  double Fr_sq = Fr * Fr;
  double denominator = 1 - Fr_sq;
  if (denominator < 0.01) denominator = 0.01;
  double Sf = ch_Se(s, g, h_in, Q);
  return (S0 - Sf) / denominator;
}

int pr_oracle_gvf(const pr_config* cfg, const pr_geom* geom, const double* q0, int64_t q0_stride,
                  const double* downstream_depth, int64_t hd_stride, double* ic_depth, double* ic_flow,
                  int32_t* status_out) {
  int N = cfg->n_nodes, M = cfg->n_members;
  xs_t* xs = (xs_t*)malloc(sizeof(xs_t) * N);
  double dx = cfg->dx;                                    /* channel.py:308 == fitted spatial step */
  for (int m = 0; m < M; ++m) {
    for (int i = 0; i < N; ++i) xs[i] = xs_load(geom, i, m);
    double Q = q0[(int64_t)m * q0_stride];
    double* hd = ic_depth + (size_t)m * N; double* qd = ic_flow + (size_t)m * N;
    int status = PR_STATUS_OK;
    double h = downstream_depth[(int64_t)m * hd_stride];
    hd[N - 1] = h; qd[N - 1] = Q;
    for (int i = N - 2; i >= 0; --i) {
      double S0 = (xs[i].z - xs[i + 1].z) / dx;           /* channel.py:344 (closure's loop index i) */
      double h_down = h;
      double dh_down = gvf_dh_dx(xs, cfg->g, Q, S0, h_down, i + 1, &status);
      double h_pred = h_down - dh_down * dx;
      if (h_pred <= 0) h_pred = 0.01;
      double dh_pred = gvf_dh_dx(xs, cfg->g, Q, S0, h_pred, i, &status);
      double dh_avg = 0.5 * (dh_down + dh_pred);
      double h_up = h_down - dh_avg * dx;
      if (h_up <= 0) h_up = 0.01;
      h = h_up;
      hd[i] = h; qd[i] = Q;
    }
    if (status_out) status_out[m] = status;
  }
  free(xs);
  return PR_OK;
}

/* numpy.interp + n_calibrate.calc_rmse_curve: model.py:105-113, n_calibrate.py:55-63 */
static double np_interp(double x, const double* xp, const double* fp, int n) {
  if (x != x) return x;
  if (x > xp[n - 1]) return fp[n - 1];
  if (x < xp[0]) return fp[0];
  /* binary search for j with xp[j] <= x < xp[j+1] (numpy's binary_search_with_guess semantics) */
  int lo = 0, hi = n;
  while (lo < hi) { int mid = lo + ((hi - lo) >> 1); if (x >= xp[mid]) lo = mid + 1; else hi = mid; }
  int j = lo - 1;
  if (j >= n - 1) return fp[n - 1];
  if (xp[j] == x) return fp[j];
  double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  double r = slope * (x - xp[j]) + fp[j];
  if (r != r) { r = slope * (x - xp[j + 1]) + fp[j + 1]; if (r != r && fp[j] == fp[j + 1]) r = fp[j]; }
  return r;
}

int pr_oracle_objective(const pr_config* cfg, const double* up_flow, const double* up_depth, double z0,
                        const double* q_query, const double* h_target, int32_t n_query, double* levels_out,
                        double* rmse_out) {
  int L = cfg->n_levels, M = cfg->n_members;
  double* fp = (double*)malloc(sizeof(double) * L);
  for (int m = 0; m < M; ++m) {
    const double* xp = up_flow + (size_t)m * L;
    for (int k = 0; k < L; ++k) fp[k] = up_depth[(size_t)m * L + k] + z0;
    double ss = 0;
    for (int j = 0; j < n_query; ++j) {
      double v = np_interp(q_query[j], xp, fp, L);
      if (levels_out) levels_out[(size_t)m * n_query + j] = v;
      double d = v - h_target[j];
      ss += d * d;
    }
    if (rmse_out) rmse_out[m] = pow(ss / n_query, 0.5);
  }
  free(fp);
  return PR_OK;
}

/* Unit-level probe for KATs against the imported reference functions: everything the node pass
 * produces for one section at (h, Q).  out[16] = A,P,R,T,K,n_eq,dR_dA,dK_dA,Sf,dSf_dA,dSf_dQ,Sc,dSc_dA,dSc_dQ,Se,0 */
int pr_oracle_section_probe(const pr_geom* geom, int node, int member, double g, double h, double Q, double* out) {
  xs_t s = xs_load(geom, node, member);
  double hw = h + s.z;
  xs_properties(&s, hw, &out[0], &out[1], &out[2], &out[3]);
  out[4] = xs_conveyance(&s, hw);
  out[5] = xs_equivalent_n(&s, hw);
  out[6] = xs_dR_dA(&s, hw);
  out[7] = xs_dK_dA(&s, hw);
  out[8] = xs_friction_slope(&s, h, Q);
  out[9] = xs_dSf_dA(&s, h, Q);
  out[10] = xs_dSf_dQ(&s, h, Q);
  out[11] = xs_curvature_slope(&s, g, h, Q);
  out[12] = xs_dSc_dA(&s, g, h, Q);
  out[13] = xs_dSc_dQ(&s, g, h, Q);
  out[14] = ch_Se(&s, g, h, Q);
  out[15] = 0;
  return PR_OK;
}
