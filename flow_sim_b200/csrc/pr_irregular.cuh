// pr_irregular.cuh - node pass for IrregularSection nodes (cross_section.py:207-543): polyline sections with
// composite roughness and finite-difference derivatives.  Used by the long-reach kernels only (a reach with an
// irregular section is routed there whatever its length): the work per node is a handful of scans over the
// section's points in global memory, not register arithmetic.
//
// One wetted sub-channel (get_subchannels, :329-372) takes the base-class formulas the reference falls back to
// (:374-439 -> :114-141).  Two or more (a bar or a ridge splitting the flow): every sub-channel is evaluated as a
// section of its own and the conveyances are combined, K = (sum K_j^1.5)^(2/3), exactly as the reference does
// (irr_split_K below; the sub-channels are views of the parent's points, PolySub - no copy, no size limit).
#pragma once
#include "pr_device.cuh"

namespace pr {

// IrregularSection.properties (:247-327) of the sub-polyline [lo, hi] taken as a section of its own (what
// get_equivalent_n's subsection_props builds, :448-470).  The two np.sum reductions are reproduced in numpy's
// pairwise order (8 interleaved accumulators up to 128 terms): the reference differentiates A and R by central
// differences with dh = 1e-6, which amplifies a last-bit difference in A a million times.
// The scans read a polyline through a view: the arrays themselves, or one wetted sub-channel of them - its submerged
// points [0, n_in) preceded / followed by a point on the water surface (what get_subchannels builds, :340-370) -
// without copying it.  One view type for both, so that the scans are compiled once.
struct PolySub {
  const double* x;          // first submerged point of the sub-channel
  const double* z;
  int n_in;                 // submerged points
  bool has_l, has_r;
  double xl, xr, hw;        // stations of the water-surface points, their elevation
  __device__ __forceinline__ double X(int k) const { const int j = k - (has_l ? 1 : 0); return j < 0 ? xl : (j >= n_in ? xr : x[j]); }
  __device__ __forceinline__ double Z(int k) const { const int j = k - (has_l ? 1 : 0); return (j < 0 || j >= n_in) ? hw : z[j]; }
  __device__ __forceinline__ int size() const { return n_in + (has_l ? 1 : 0) + (has_r ? 1 : 0); }
};

static __device__ __noinline__ void irr_properties(const PolySub& pl, int lo, int hi, double hw, double& A_out, double& P_out, double& T_out) {
  A_out = 0.0; P_out = 0.0; T_out = 0.0;
  const int n = hi - lo + 1;
  if (n <= 0) return;
  auto x = [&](int i) { return pl.X(lo + i); };
  auto z = [&](int i) { return pl.Z(lo + i); };
  double z_min = z(0);
  for (int i = 1; i < n; ++i) { const double zi = z(i); z_min = zi < z_min ? zi : z_min; }
  if (hw <= z_min) return;
  double A_total = 0.0, P_total = 0.0, T_total = 0.0;
  int i = 0;
  while (i < n) {
    if (hw - z(i) > 0.0) {
      const int i0 = i;
      while (i + 1 < n && hw - z(i + 1) > 0.0) i += 1;
      const int iN = i;
      const bool has_l = i0 > 0 && z(i0 - 1) > hw, has_r = iN < n - 1 && z(iN + 1) > hw;
      double xl = 0.0, xr = 0.0;
      if (has_l) { const double z0 = z(i0 - 1), z1 = z(i0), x0 = x(i0 - 1), x1 = x(i0); xl = x0 + (hw - z0) / (z1 - z0) * (x1 - x0); }
      if (has_r) { const double z0 = z(iN), z1 = z(iN + 1), x0 = x(iN), x1 = x(iN + 1); xr = x0 + (hw - z0) / (z1 - z0) * (x1 - x0); }
      const int m = (iN - i0 + 1) + (has_l ? 1 : 0) + (has_r ? 1 : 0), nt = m - 1;
      // point k of the wetted polyline (intersection points carry z = hw)
      auto px = [&](int k) { const int j = k - (has_l ? 1 : 0); return j < 0 ? xl : (j > iN - i0 ? xr : x(i0 + j)); };
      auto pz = [&](int k) { const int j = k - (has_l ? 1 : 0); return (j < 0 || j > iN - i0) ? hw : z(i0 + j); };
      auto term = [&](int k, double& ta, double& tp) {
        const double xa = px(k), xb = px(k + 1), za = pz(k), zb = pz(k + 1);
        const double d0 = fmax(hw - za, 0.0), d1 = fmax(hw - zb, 0.0), dx = xb - xa, dz = zb - za;
        ta = 0.5 * (d0 + d1) * dx;
        tp = sqrt(dx * dx + dz * dz);
      };
      // np.sum of terms [k0, k0 + cnt), cnt <= 128: numpy's pairwise_sum block (8 interleaved accumulators from 8 terms on)
      auto block = [&](int k0, int cnt, double& sa, double& sp) {
        sa = 0.0; sp = 0.0;
        if (cnt < 8) {
          for (int k = 0; k < cnt; ++k) { double ta, tp; term(k0 + k, ta, tp); sa += ta; sp += tp; }
        } else {
          double ra[8], rp[8];
          for (int k = 0; k < 8; ++k) term(k0 + k, ra[k], rp[k]);
          const int nb = cnt - (cnt % 8);
          for (int k = 8; k < nb; ++k) { double ta, tp; term(k0 + k, ta, tp); ra[k & 7] += ta; rp[k & 7] += tp; }
          sa = ((ra[0] + ra[1]) + (ra[2] + ra[3])) + ((ra[4] + ra[5]) + (ra[6] + ra[7]));
          sp = ((rp[0] + rp[1]) + (rp[2] + rp[3])) + ((rp[4] + rp[5]) + (rp[6] + rp[7]));
          for (int k = nb; k < cnt; ++k) { double ta, tp; term(k0 + k, ta, tp); sa += ta; sp += tp; }
        }
      };
      double sa = 0.0, sp = 0.0;
      if (nt <= 128) {
        block(0, nt, sa, sp);
      } else {
        // beyond 128 terms numpy halves recursively: sum(a[:n2]) + sum(a[n2:]), n2 = n/2 rounded down to a multiple of 8.
        // The same tree, walked with a small explicit stack (depth <= log2(nt / 64)).
        constexpr int kDepth = 10;
        int lo_s[kDepth], n_s[kDepth], st_s[kDepth];
        double acc_a[kDepth], acc_p[kDepth];
        int top = 0;
        lo_s[0] = 0; n_s[0] = nt; st_s[0] = 0;
        while (top >= 0) {
          int n2 = n_s[top] / 2;
          n2 -= n2 % 8;
          if (st_s[top] == 0) {
            if (n_s[top] <= 128 || top + 1 >= kDepth) { block(lo_s[top], n_s[top], sa, sp); --top; }   // (depth 10: 64k terms)
            else { st_s[top] = 1; lo_s[top + 1] = lo_s[top]; n_s[top + 1] = n2; st_s[top + 1] = 0; ++top; }
          } else if (st_s[top] == 1) {            // the left half is in (sa, sp)
            acc_a[top] = sa; acc_p[top] = sp;
            st_s[top] = 2; lo_s[top + 1] = lo_s[top] + n2; n_s[top + 1] = n_s[top] - n2; st_s[top + 1] = 0; ++top;
          } else { sa = acc_a[top] + sa; sp = acc_p[top] + sp; --top; }
        }
      }
      A_total += sa; P_total += sp; T_total += px(m - 1) - px(0);
    }
    i += 1;
  }
  A_out = A_total; P_out = P_total; T_out = T_total;
}

// conveyance of the points with x_min <= x <= x_max (subsection_props, :448-470)
__device__ inline double irr_sub_K(const PolySub& pl, int n, double hw, double x_min, double x_max, double n_value) {
  int lo = 0, hi = n - 1;
  while (lo < n && !(pl.X(lo) >= x_min)) ++lo;
  while (hi >= 0 && !(pl.X(hi) <= x_max)) --hi;
  if (hi - lo + 1 < 2) return 0.0;
  double A, P, T;
  irr_properties(pl, lo, hi, hw, A, P, T);
  if (A <= 0.0 || P <= 0.0) return 0.0;
  return A * pow(A / P, 2.0 / 3.0) / n_value;                 // hydraulics.conveyance (hydraulics.py:15-26)
}

// IrregularSection.properties / get_equivalent_n / conveyance / dR_dA / dK_dA of the polyline (x, z)[0..n) with
// composite-roughness limits and values (cross_section.py:247-327, 441-532).
struct IrrSec {
  double A, P, T, R;        // properties(hw)
  double A1, A2;            // area(hw -/+ 1e-6): dA_dh = (A2 - A1) / 2e-6 (:534-539)
  double n_eq, dRA, K, dKA;
};

static __device__ __noinline__ void irr_section(const PolySub& pl, int n, double hw, double lim_l, double lim_r,
                                   double nl, double nm, double nr, IrrSec& s) {
  const double dh = 1e-6;
  double P1, T1, P2, T2;
  irr_properties(pl, 0, n - 1, hw, s.A, s.P, s.T);
  irr_properties(pl, 0, n - 1, hw - dh, s.A1, P1, T1);
  irr_properties(pl, 0, n - 1, hw + dh, s.A2, P2, T2);
  s.R = s.P > 0.0 ? s.A / s.P : 0.0;
  const double R1 = P1 > 0.0 ? s.A1 / P1 : 0.0, R2 = P2 > 0.0 ? s.A2 / P2 : 0.0;
  // get_equivalent_n (:441-500)
  s.n_eq = nm;
  if (s.A > 0.0 && s.P > 0.0) {
    const double Kl = irr_sub_K(pl, n, hw, pl.X(0), lim_l, nl);
    const double Km = irr_sub_K(pl, n, hw, lim_l, lim_r, nm);
    const double Kr = irr_sub_K(pl, n, hw, lim_r, pl.X(n - 1), nr);
    const double K_total = pow(pow(Kl, 1.5) + pow(Km, 1.5) + pow(Kr, 1.5), 2.0 / 3.0);
    if (K_total > 0.0) s.n_eq = (s.A * pow(s.R, 2.0 / 3.0)) / K_total;
  }
  // conveyance (:502-510), dR_dA (:524-532), dK_dA (:512-522 with hydraulics.dK_dA_, hydraulics.py:28-40)
  s.K = 0.0; s.dKA = 0.0;
  s.dRA = (s.A2 - s.A1) == 0.0 ? 0.0 : (R2 - R1) / (s.A2 - s.A1);
  if (s.A > 0.0) {
    s.K = s.A * pow(s.R, 2.0 / 3.0) / s.n_eq;
    s.dKA = (pow(s.R, 2.0 / 3.0) + s.A * 2. / 3. * pow(s.R, 2.0 / 3.0 - 1.0) * s.dRA) / s.n_eq;
  }
}

__device__ inline void irr_section(const double* x, const double* z, int n, double hw, double lim_l, double lim_r,
                                   double nl, double nm, double nr, IrrSec& s) {
  irr_section(PolySub{x, z, n, false, false, 0.0, 0.0, hw}, n, hw, lim_l, lim_r, nl, nm, nr, s);
}
__device__ inline void irr_properties(const double* x, const double* z, int lo, int hi, double hw, double& A, double& P, double& T) {
  irr_properties(PolySub{x, z, hi + 1, false, false, 0.0, 0.0, hw}, lo, hi, hw, A, P, T);
}

// numpy.interp(x, [xp0, xp1], [fp0, fp1]) as get_subchannels calls it (:357,361).  On the LEFT edge of a sub-channel
// the reference passes xp = [z[start-1], z[start]], which decreases; numpy then leaves through its "beyond the last
// abscissa" exit and returns fp1 - the first submerged point itself, not the intersection, so the sub-channel gets a
// vertical wall there.  Reproduced as is.
__device__ inline double irr_np_interp2(double x, double xp0, double xp1, double fp0, double fp1) {
  if (x != x) return x;
  if (x > xp1) return fp1;
  if (x < xp0) return fp0;
  if (x == xp1) return fp1;
  const double slope = (fp1 - fp0) / (xp1 - xp0);
  double r = slope * (x - xp0) + fp0;
  if (r != r) {
    r = slope * (x - xp1) + fp1;
    if (r != r && fp0 == fp1) r = fp0;
  }
  return r;
}

// Split flow (IrregularSection.friction_slope / dSf_dA / dSf_dQ with several sub-channels, :374-439): each run of
// >= 2 submerged points plus the points where it meets the water surface becomes a section of its own with the
// parent's roughness limits and values; K_eq = (sum K_j^1.5)^(2/3), dK_eq/dA = 2/3 (sum K_j^1.5)^(-1/3) sum 1.5 K_j^0.5 dK_j/dA_j.
// The sub-channel is a view of the parent's points (PolySub), whatever its size.
__device__ inline void irr_split_K(const double* __restrict__ x, const double* __restrict__ z, int n, double hw,
                                   double lim_l, double lim_r, double nl, double nm, double nr, double& K_eq,
                                   double& dKA_eq) {
  double K_sum = 0.0, dK_sum = 0.0;
  for (int i = 0; i < n;) {
    if (!(z[i] < hw)) { ++i; continue; }
    const int start = i;
    while (i < n && z[i] < hw) ++i;
    const int end = i;
    if (end - start < 2) continue;
    PolySub sub_pl{x + start, z + start, end - start, false, false, 0.0, 0.0, hw};
    if (start > 0 && z[start - 1] > hw) { sub_pl.has_l = true; sub_pl.xl = irr_np_interp2(hw, z[start - 1], z[start], x[start - 1], x[start]); }
    if (end < n && z[end - 1] < hw && z[end] > hw) { sub_pl.has_r = true; sub_pl.xr = irr_np_interp2(hw, z[end - 1], z[end], x[end - 1], x[end]); }
    IrrSec sub;
    irr_section(sub_pl, sub_pl.size(), hw, lim_l, lim_r, nl, nm, nr, sub);
    K_sum += pow(sub.K, 1.5);
    dK_sum += 1.5 * pow(sub.K, 0.5) * sub.dKA;
  }
  K_eq = pow(K_sum, 2.0 / 3.0);
  dKA_eq = (2.0 / 3.0) * pow(K_sum, -1.0 / 3.0) * dK_sum;
}

// ---- stage tables ----------------------------------------------------------------------------------------------
// Area, wetted perimeter and top width of a polyline are piecewise polynomials of the stage: between two consecutive
// vertex elevations every segment is either submerged (area linear in hw), cut by the water surface (area quadratic,
// perimeter and width linear) or dry.  One table per node turns the six polyline scans of a node evaluation
// (properties at hw and hw -/+ 1e-6, three roughness sub-sections) into one interval search and a few Horner steps.
// The values agree with the scans to rounding (~1e-16 relative); the reference's finite-difference derivatives are
// formed from them exactly as from the scanned values.
constexpr int kIrrTabCols = 22;     // whole section: a0 a1 a2 p0 p1 t0 t1;  left, main, right sub-section: a0 a1 a2 p0 p1
constexpr int kIrrTabMaxPts = 128;  // larger sections get no table (tab_n = 0) and keep the scanning node pass

// coefficients of the points [lo, hi] (a section of its own, as subsection_props builds it) for zk < hw <= next breakpoint
__device__ inline void irr_tab_accumulate(const double* x, const double* z, int lo, int hi, double zk, double* c, bool want_t) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, p0 = 0.0, p1 = 0.0, t0 = 0.0, t1 = 0.0;
  for (int j = lo; j < hi; ++j) {
    const double za = z[j], zb = z[j + 1], dx = x[j + 1] - x[j], dz = zb - za;
    const double zl = za < zb ? za : zb, zh = za < zb ? zb : za;
    if (zh <= zk) {                         // both ends submerged: 0.5 (d_a + d_b) dx, the whole segment wetted
      a0 += dx * (zk - 0.5 * (za + zb)); a1 += dx;
      p0 += sqrt(dx * dx + dz * dz);
      t0 += dx;
    } else if (zl <= zk) {                  // cut by the water surface at the fraction (hw - zl) / (zh - zl)
      const double inv = 1.0 / (zh - zl), cc = zk - zl, len = sqrt(dx * dx + dz * dz);
      const double gq = 0.5 * dx * inv;
      a0 += gq * cc * cc; a1 += 2.0 * gq * cc; a2 += gq;
      p0 += len * cc * inv; p1 += len * inv;
      t0 += dx * cc * inv; t1 += dx * inv;
    }
  }
  c[0] = a0; c[1] = a1; c[2] = a2; c[3] = p0; c[4] = p1;
  if (want_t) { c[5] = t0; c[6] = t1; }
}

static __global__ void pr_irr_build_tables(DevGeom g, int N, int* tab_n, double* tab_z, double* tab_c, int* tab_runs) {
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= N) return;
  tab_n[node] = 0;
  if (g.kind[node] != PR_XS_IRREGULAR) return;
  const int off = g.irr_offset[node], n = g.irr_offset[node + 1] - off;
  if (n < 2 || n > kIrrTabMaxPts) return;
  const double* x = g.irr_x + off;
  const double* z = g.irr_z + off;
  double zz[kIrrTabMaxPts];
  int nb = 0;
  for (int i = 0; i < n; ++i) {             // insertion sort of the distinct elevations
    const double v = z[i];
    int k = nb;
    while (k > 0 && zz[k - 1] > v) --k;
    if (k > 0 && zz[k - 1] == v) continue;
    for (int m = nb; m > k; --m) zz[m] = zz[m - 1];
    zz[k] = v; ++nb;
  }
  // the three roughness sub-sections: points with x_min <= x <= x_max (get_equivalent_n, cross_section.py:448-470)
  const double lim_l = g.irr_left[node], lim_r = g.irr_right[node];
  int lo[3], hi[3];
  const double xmin[3] = {x[0], lim_l, lim_r}, xmax[3] = {lim_l, lim_r, x[n - 1]};
  for (int s = 0; s < 3; ++s) {
    lo[s] = 0; hi[s] = n - 1;
    while (lo[s] < n && !(x[lo[s]] >= xmin[s])) ++lo[s];
    while (hi[s] >= 0 && !(x[hi[s]] <= xmax[s])) --hi[s];
  }
  for (int k = 0; k < nb; ++k) {
    const double zk = zz[k];
    double* c = tab_c + (size_t)(off + k) * kIrrTabCols;
    tab_z[off + k] = zk;
    irr_tab_accumulate(x, z, 0, n - 1, zk, c, true);
    for (int s = 0; s < 3; ++s) {
      double* cs = c + 7 + 5 * s;
      if (hi[s] - lo[s] + 1 < 2) { cs[0] = cs[1] = cs[2] = cs[3] = cs[4] = 0.0; }
      else irr_tab_accumulate(x, z, lo[s], hi[s], zk, cs, false);
    }
    int runs = 0;                           // get_subchannels: runs of >= 2 points with z < hw, i.e. z <= zk here
    for (int i = 0; i < n;) {
      if (!(z[i] <= zk)) { ++i; continue; }
      const int st = i;
      while (i < n && z[i] <= zk) ++i;
      runs += (i - st >= 2) ? 1 : 0;
    }
    tab_runs[off + k] = runs;
  }
  tab_n[node] = nb;
}

struct IrrTab {
  const double *z, *c;
  const int* runs;
  int nb;
};

__device__ __forceinline__ bool irr_tab_get(const DevGeom& g, int node, IrrTab& t) {
  t.nb = g.irr_tab_n ? g.irr_tab_n[node] : 0;
  if (t.nb == 0) return false;
  const int off = g.irr_offset[node];
  t.z = g.irr_tab_z + off; t.c = g.irr_tab_c + (size_t)off * kIrrTabCols; t.runs = g.irr_tab_runs + off;
  return true;
}

// interval of a stage: the largest k with z_k < hw (a point AT the water level is dry, as in properties), -1 = dry section
__device__ __forceinline__ int irr_tab_interval(const IrrTab& t, double hw) {
  int lo = 0, hi = t.nb;                    // first k with z_k >= hw
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.z[mid] < hw) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

// A vertex exactly AT the stage: the reference's properties() adds an intersection point only beside a vertex strictly
// above the water (cross_section.py:292,300), so with z == hw the segments between that vertex and its wet neighbours
// drop out of area, perimeter and width - a jump at one stage value that the polynomial pieces do not have.  It is hit
// in practice (an initial depth of 2.0 over a survey with round elevations), so such stages go to the scans.
// k = irr_tab_interval(t, hw).
__device__ __forceinline__ bool irr_tab_tie(const IrrTab& t, int k, double hw) { return k + 1 < t.nb && t.z[k + 1] == hw; }

// sec: 0 whole section, 1 / 2 / 3 left / main / right roughness sub-section
__device__ __forceinline__ void irr_tab_eval(const IrrTab& t, int k, double hw, int sec, double& A, double& P, double* T) {
  if (k < 0) { A = 0.0; P = 0.0; if (T) *T = 0.0; return; }
  const double u = hw - t.z[k];
  const double* c = t.c + (size_t)k * kIrrTabCols + (sec == 0 ? 0 : 7 + 5 * (sec - 1));
  A = c[0] + u * (c[1] + u * c[2]);
  P = c[3] + u * c[4];
  if (T) *T = c[5] + u * c[6];
}

__device__ __forceinline__ double irr_pow23(double v) { const double c = cbrt(v); return c * c; }     // v^(2/3), v >= 0

// irr_section from the node's stage table; false = one of the three stages sits exactly on a vertex elevation (see
// irr_tab_tie): nothing is written, the caller scans
__device__ inline bool irr_section_tab(const IrrTab& t, double hw, double nl, double nm, double nr, IrrSec& s, int& runs) {
  const double dh = 1e-6;
  const int k = irr_tab_interval(t, hw);
  const int k1 = (k >= 0 && hw - dh > t.z[k]) ? k : irr_tab_interval(t, hw - dh);
  const int k2 = (k >= 0 && (k + 1 >= t.nb || !(hw + dh > t.z[k + 1]))) ? k : irr_tab_interval(t, hw + dh);
  if (irr_tab_tie(t, k, hw) || irr_tab_tie(t, k1, hw - dh) || irr_tab_tie(t, k2, hw + dh)) return false;
  double P1, P2;
  irr_tab_eval(t, k, hw, 0, s.A, s.P, &s.T);
  irr_tab_eval(t, k1, hw - dh, 0, s.A1, P1, nullptr);
  irr_tab_eval(t, k2, hw + dh, 0, s.A2, P2, nullptr);
  runs = k >= 0 ? t.runs[k] : 0;
  s.R = s.P > 0.0 ? s.A / s.P : 0.0;
  const double R1 = P1 > 0.0 ? s.A1 / P1 : 0.0, R2 = P2 > 0.0 ? s.A2 / P2 : 0.0;
  const double R23 = irr_pow23(s.R);
  s.n_eq = nm;
  if (s.A > 0.0 && s.P > 0.0) {             // get_equivalent_n (:441-500)
    double sum15 = 0.0;
    const double nv[3] = {nl, nm, nr};
#pragma unroll
    for (int sec = 1; sec <= 3; ++sec) {
      double As, Ps;
      irr_tab_eval(t, k, hw, sec, As, Ps, nullptr);
      if (As > 0.0 && Ps > 0.0) {
        const double Ks = As * irr_pow23(As / Ps) / nv[sec - 1];
        sum15 += Ks * sqrt(Ks);
      }
    }
    const double K_total = irr_pow23(sum15);
    if (K_total > 0.0) s.n_eq = (s.A * R23) / K_total;
  }
  s.K = 0.0; s.dKA = 0.0;
  s.dRA = (s.A2 - s.A1) == 0.0 ? 0.0 : (R2 - R1) / (s.A2 - s.A1);
  if (s.A > 0.0) {
    s.K = s.A * R23 / s.n_eq;
    s.dKA = (R23 + s.A * (2. / 3.) * (R23 / s.R) * s.dRA) / s.n_eq;
  }
  return true;
}

// Everything the scheme needs from an irregular node: the counterpart of node_eval.
//   T is Solver.dA_dh = the central difference of the area (:534-539), which is what the Jacobian uses.
//   top_width (optional): the geometric top width of `properties`, which the GVF initial profile uses (channel.py:320).
template <class KP, bool CURV = false>
__device__ inline void node_eval_irregular(const DevGeom& g, int node, double h, double Q, const Rough& rg, const KP& k,
                                           NodeVals& o, NodeConv* kc, double* top_width = nullptr) {
  const int off = g.irr_offset[node], n = g.irr_offset[node + 1] - off;
  const double* x = g.irr_x + off;
  const double* z = g.irr_z + off;
  const double z_min = g.z[node];
  const double hw = h + z_min;
  const double nm = rg.om ? rg.nm : g.nm[node];
  const double nl = rg.ofp ? rg.nfp : g.nl[node], nr = rg.ofp ? rg.nfp : g.nr[node];
  const double dh = 1e-6;
  const double lim_l = g.irr_left[node], lim_r = g.irr_right[node];
  IrrSec sec;
  // wetted sub-channels (z < hw runs of >= 2 points): with more than one the friction slope and its derivatives take
  // the combined conveyance of the sub-channels; everything else stays with the whole section
  int runs = 0;
  IrrTab tab;
  if (!(irr_tab_get(g, node, tab) && irr_section_tab(tab, hw, nl, nm, nr, sec, runs))) {
    irr_section(x, z, n, hw, lim_l, lim_r, nl, nm, nr, sec);
    for (int i = 0; i < n;) {
      if (!(z[i] < hw)) { ++i; continue; }
      const int s = i;
      while (i < n && z[i] < hw) ++i;
      runs += (i - s >= 2) ? 1 : 0;
    }
  }
  const double A = sec.A, T = sec.T, R = sec.R, A1 = sec.A1, A2 = sec.A2, n_eq = sec.n_eq, dRA = sec.dRA;
  const double K = sec.K, dKA = sec.dKA;
  double Kf = K, dKAf = dKA;
  if (runs > 1) irr_split_K(x, z, n, hw, lim_l, lim_r, nl, nm, nr, Kf, dKAf);
  const double absQ = fabs(Q);
  const double Sf = Q * absQ / (Kf * Kf);                      // hydraulics.Sf (hydraulics.py:42-57)
  const double dSfA = -2 * Sf * (dKAf / Kf), dSfQ = 2 * absQ / (Kf * Kf);
  const double dAdh = (A2 - A1) / (2 * dh);
  double Se = Sf, dSeA = dSfA, dSeQ = dSfQ;
  if (CURV) {
    // CrossSection.curvature_slope / dSc_dA / dSc_dQ (cross_section.py:143-175) with hydraulics.Sc, dSc_dA, dSc_dQ,
    // froude_num, dFr_dA, dFr_dQ (hydraulics.py:94-204): geometric top width T, composite n, finite-difference dR/dA
    const double curv = g.curv[node];
    if (curv != 0.0) {
      const double rc = 1.0 / curv, gg = k.g;
      const double V = Q / fmax(A, 1e-6), D = A / fmax(T, 1e-6);
      const double Fr = V / sqrt(gg * fmax(D, 1e-6));
      const double C = pow(R, 1.0 / 6.0) / n_eq;
      const double f = 8 * gg / (C * C), sqf = sqrt(f);
      const double lead = 2.86 * sqf + 2.07 * f;
      const double num = lead * (h * h) * (Fr * Fr), den = (0.565 + sqf) * (rc * rc);
      Se += num / den;
      if (fabs(curv) > 1e-12) {
        const double Vu = Q / A, Du = A / T;                       // unclamped in the derivatives (quirk 7)
        const double dFrA = -0.5 * Vu * pow(gg * Du, -1.5) * gg * (1.0 / T) + (-Q / (A * A)) * pow(gg * Du, -0.5);
        const double dFrQ = (1.0 / A) * pow(gg * Du, -0.5);
        const double dfA = -(8.0 / 3.0) * gg * (n_eq * n_eq) * pow(R, -4.0 / 3.0) * dRA;
        const double dnumA = (2.86 / (2 * sqf) * dfA + 2.07 * dfA) * (h * h) * (Fr * Fr) +
                             lead * (2 * h * (1. / T) * (Fr * Fr) + (h * h) * 2 * Fr * dFrA);
        const double ddenA = (1.0 / (2 * sqf) * dfA) * (rc * rc);
        dSeA += (dnumA * den - num * ddenA) / (den * den) * dAdh;  // already x dA_dh, multiplied again below (quirk 7)
        const double dnumQ = lead * (h * h) * 2 * Fr * dFrQ;
        dSeQ += (dnumQ * den - num * 0.0) / (den * den);
      }
    }
  }
  const double QA = Q / A;
  o.Q = Q; o.A = A; o.T = dAdh; o.Y = hw; o.Se = Se; o.F = Q * QA; o.QA = QA;
  o.w1 = (k.th_dx * QA) * (QA * dAdh);
  o.w2 = (k.hth * dSeA) * dAdh;
  o.w3 = k.hth * dSeQ;
  o.w4 = k.th_dx2 * QA;
  if (kc) {
    // the boundary rows (normal depth; head losses of a lumped storage, lumped_storage.py:58-116 through
    // hydraulics.Sf(A, Q, n, R)) take the conveyance of the WHOLE section even where the scheme's friction slope splits
    const double SfW = Q * absQ / (K * K);
    kc->K = K; kc->dKA = dKA; kc->A = A; kc->Sf = SfW; kc->dSfA = -2 * SfW * (dKA / K); kc->dSfQ = 2 * absQ / (K * K);
  }
  if (top_width) *top_width = T;
}

// Out-of-line entry for kernels with many node-pass call sites (the fused ensemble kernel): one copy of the scans.
template <bool CURV, class KP = DevParams>
__device__ __noinline__ void node_eval_irregular_call(const DevGeom& g, int node, double h, double Q, const Rough& rg,
                                                      const KP& k, NodeVals& o, NodeConv* kc, double* top_width = nullptr) {
  node_eval_irregular<KP, CURV>(g, node, h, Q, rg, k, o, kc, top_width);
}

}  // namespace pr
