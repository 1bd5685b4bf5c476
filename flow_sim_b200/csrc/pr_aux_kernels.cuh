// pr_aux_kernels.cuh - the small kernels either side of the Newton loop:
//   * GVF backwater initial conditions, one member per thread (Channel._gvh_conditions, channel.py:307-378)
//   * calibration objective np.interp + RMSE, one member per thread (model.py:105-113, n_calibrate.py:55-63)
//   * DFMA peak probe (roofline denominator)
#pragma once
#include "pr_device.cuh"
#include "pr_irregular.cuh"

namespace pr {

struct GvfParams {
  int N, M;
  double dx, g;
  const double* h_down;
  long long h_down_stride;
  double th_dx, hth, th_dx2, theta, mtheta;   // unused scheme constants node_eval reads (zero)
  DevGeom geo;
  const double* q0;
  long long q0_stride;
  double *ic_h, *ic_q;
  int* status;
  const double* table;         // [F_COUNT][N] derived geometry in global memory for reaches too long to stage, else NULL
};

// dh/dx of the gradually-varied-flow equation at one node (get_dh_dx, channel.py:316-347)
template <bool CURV, int RM>
__device__ __forceinline__ double gvf_slope(const double* sg, int NP, int node, double h_in, double Q, double S0,
                                            const Rough& rg, const GvfParams& p, int& status) {
  const double g = p.g;
  NodeVals nv;
  if (p.geo.irr_offset && sg[F_KIND * NP + node] == (double)PR_XS_IRREGULAR) {
    double top;
    node_eval_irregular_call<CURV, GvfParams>(p.geo, node, h_in, Q, rg, p, nv, nullptr, &top);      // out of line
    nv.T = top;                                   // the profile uses the geometric top width, not dA/dh
  } else {
    node_eval<CURV, RM, false, GvfParams>(sg, NP, node, h_in, Q, rg, p, nv);
  }
  if (nv.T < 1e-6 || nv.A < 1e-6 || !(h_in > 0.0)) return 0.0;
  const double V = Q / fmax(nv.A, 1e-6), D = nv.A / fmax(nv.T, 1e-6);    // hydraulics.froude_num (:155-168)
  const double Fr = V / sqrt(g * fmax(D, 1e-6));
  if (Fr > 1.0) status = PR_STATUS_SUPERCRITICAL;                        // channel.py:328-332 raises
  double den = 1.0 - Fr * Fr;
  if (den < 0.01) den = 0.01;
  return (S0 - nv.Se) / den;
}

template <bool CURV, int RM>
__global__ void __launch_bounds__(128) pr_gvf_kernel(const __grid_constant__ GvfParams p) {
  extern __shared__ double gvf_smem[];
  const int NP = p.N;
  const double* smem = p.table ? p.table : gvf_smem;
  if (!p.table) {
    stage_geometry(p.geo, p.N, NP, gvf_smem, threadIdx.x, blockDim.x, [](int idx) { return idx; });
    __syncthreads();
  }
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const Rough rg = load_rough<RM>(p.geo, m);
  const double Q = p.q0[m * p.q0_stride];
  const int N = p.N;
  double* oh = p.ic_h + (size_t)m * N;
  double* oq = p.ic_q + (size_t)m * N;
  int status = PR_STATUS_OK;
  double h = p.h_down[m * p.h_down_stride];
  oh[N - 1] = h;
  oq[N - 1] = Q;
  for (int i = N - 2; i >= 0; --i) {
    const double S0 = (smem[F_Z * NP + i] - smem[F_Z * NP + i + 1]) / p.dx;   // channel.py:344
    const double h_down = h;
    const double s_down = gvf_slope<CURV, RM>(smem, NP, i + 1, h_down, Q, S0, rg, p, status);   // predictor
    double h_pred = h_down - s_down * p.dx;
    if (h_pred <= 0.0) h_pred = 0.01;
    const double s_pred = gvf_slope<CURV, RM>(smem, NP, i, h_pred, Q, S0, rg, p, status);       // corrector
    double h_up = h_down - 0.5 * (s_down + s_pred) * p.dx;
    if (h_up <= 0.0) h_up = 0.01;
    h = h_up;
    oh[i] = h;
    oq[i] = Q;
  }
  if (p.status) p.status[m] = status;
}

struct ObjParams {
  int L, M, nq;
  double z0;
  const double *up_q, *up_h, *q_query, *h_target;
  double *levels, *rmse;
};

// numpy.interp semantics for one query (xp assumed increasing, as np.interp assumes)
__device__ __forceinline__ double interp_np(double x, const double* xp, const double* fp, double z0, int n) {
  if (x != x) return x;
  if (x > xp[n - 1]) return fp[n - 1] + z0;
  if (x < xp[0]) return fp[0] + z0;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (x >= xp[mid]) lo = mid + 1; else hi = mid;
  }
  const int j = lo - 1;
  if (j >= n - 1) return fp[n - 1] + z0;
  const double f0 = fp[j] + z0, f1 = fp[j + 1] + z0;
  if (xp[j] == x) return f0;
  const double slope = (f1 - f0) / (xp[j + 1] - xp[j]);
  return slope * (x - xp[j]) + f0;
}

__global__ void __launch_bounds__(128) pr_objective_kernel(const __grid_constant__ ObjParams p) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const double* xp = p.up_q + (size_t)m * p.L;
  const double* fp = p.up_h + (size_t)m * p.L;
  double ss = 0.0;
  for (int j = 0; j < p.nq; ++j) {
    const double v = interp_np(p.q_query[j], xp, fp, p.z0, p.L);
    if (p.levels) p.levels[(size_t)m * p.nq + j] = v;
    const double d = v - p.h_target[j];
    ss += d * d;
  }
  if (p.rmse) p.rmse[m] = sqrt(ss / p.nq);
}

// 8 independent DFMA chains per thread; 2*8*iters flops per thread.
__global__ void __launch_bounds__(256) pr_dfma_kernel(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) sink[0] = s;
}

// Derived result arrays for every (member, level, node): Solver.prepare_results (solver.py:65-98).
//   level = depth + z_min, area, top_width (TrapezoidalSection.properties, cross_section.py:623-679),
//   froude_number (hydraulics.froude_num with its 1e-6 clamps), velocity = Q/A, wave_celerity = V + sqrt(g A / T)
// Pure streaming kernel: 16 B in, up to 48 B out per element, geometry table [F_COUNT][N] from L2.
struct DerivedParams {
  long long rows;    // M * L
  int N;
  double g;
  const double* geo;
  DevGeom raw;       // polylines of the IrregularSection nodes
  const double *depth, *flow;
  double *level, *area, *top_width, *froude, *velocity, *celerity;
};

// Thread = one node: its geometry is loaded once into registers, then the thread walks down the (member, level)
// rows, so every load and store of a warp is a contiguous 256-byte segment.
__global__ void __launch_bounds__(128) pr_derived_kernel(const __grid_constant__ DerivedParams p) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= p.N) return;
#define GEO(f) p.geo[(size_t)(f) * p.N + nd]
  const double z = GEO(F_Z), b = GEO(F_B), hb = GEO(F_HB), m2 = 2.0 * GEO(F_M);
  const double mfp = GEO(F_MFP), bl = GEO(F_BL), br = GEO(F_BR), amf = GEO(F_AMF), wb = GEO(F_WB);
  const bool irregular = GEO(F_KIND) == (double)PR_XS_IRREGULAR;
#undef GEO
  const int poly_off = irregular ? p.raw.irr_offset[nd] : 0, poly_n = irregular ? p.raw.irr_offset[nd + 1] - poly_off : 0;
  for (long long r = blockIdx.y; r < p.rows; r += gridDim.y) {
    const size_t i = (size_t)r * p.N + nd;
    const double h = p.depth[i], Q = p.flow[i];
    const double hw = z + h;
    const double d = fmax(0.0, hw - z);
    double A, T;
    if (irregular) {             // IrregularSection.properties (cross_section.py:247-327): area and geometric top width
      double P;
      IrrTab tab;
      int k = 0;
      if (irr_tab_get(p.raw, nd, tab) && !irr_tab_tie(tab, k = irr_tab_interval(tab, hw), hw)) irr_tab_eval(tab, k, hw, 0, A, P, &T);
      else irr_properties(p.raw.irr_x + poly_off, p.raw.irr_z + poly_off, 0, poly_n - 1, hw, A, P, T);
    } else if (d <= hb) {        // rectangle / simple trapezoid / compound in bank (h_bank staged as 1e300 otherwise)
      T = b + m2 * d;
      A = (b + T) / 2.0 * d;
      if (d <= 0.0) { A = 0.0; T = 0.0; }
    } else {
      const double dfp = d - hb;
      A = amf + (bl + 0.5 * mfp * dfp) * dfp + (br + 0.5 * mfp * dfp) * dfp;
      T = wb + 2.0 * mfp * dfp;
    }
    const double V = Q / fmax(A, 1e-6), D = A / fmax(T, 1e-6);
    if (p.level) p.level[i] = hw;
    if (p.area) p.area[i] = A;
    if (p.top_width) p.top_width[i] = T;
    if (p.froude) p.froude[i] = V / sqrt(p.g * fmax(D, 1e-6));
    const double vel = Q / A;
    if (p.velocity) p.velocity[i] = vel;
    if (p.celerity) p.celerity[i] = vel + sqrt(p.g * A / T);
  }
}

// Steady uniform-flow initial state: Channel._steady_conditions (channel.py:296-305) = per node the root of
// Q - K(hw) sqrt(S0) on [z_min, z_min + 100] by Brent's method (CrossSection.normal_depth, cross_section.py:184-202;
// scipy.optimize.brentq, xtol = 2e-12, rtol = 4 eps, maxiter = 100).  One (member, node) per thread.
struct NormalDepthParams {
  int N, M;
  double g;
  const double* geo;        // derived geometry table [F_COUNT][N]
  DevGeom raw;              // member roughness overrides
  const double* bed_slope;  // [N]
  const double* q0;
  long long q0_stride;
  double *ic_h, *ic_q;
  double th_dx, hth, th_dx2, theta, mtheta;   // unused scheme constants node_eval reads (zero)
};

template <int RM>
__device__ __forceinline__ double normal_flow_residual(const NormalDepthParams& p, int nd, double hw, double Qt,
                                                       double S0, const Rough& rg) {
  if (!(S0 > 0.0)) return Qt;                       // CrossSection.normal_flow returns 0 for a non-positive slope
  const double z = p.geo[(size_t)F_Z * p.N + nd];
  const double depth = hw - z;
  if (!(depth > 0.0)) return Qt;                    // K(0) = 0
  if (p.geo[(size_t)F_KIND * p.N + nd] == (double)PR_XS_IRREGULAR) {
    // IrregularSection.conveyance of the whole section (normal_flow does not split, cross_section.py:177-182, 502-510)
    const int off = p.raw.irr_offset[nd], n = p.raw.irr_offset[nd + 1] - off;
    const double nm = rg.om ? rg.nm : p.raw.nm[nd];
    const double nl = rg.ofp ? rg.nfp : p.raw.nl[nd], nr = rg.ofp ? rg.nfp : p.raw.nr[nd];
    IrrSec sec;
    IrrTab tab;
    int runs;
    if (!(irr_tab_get(p.raw, nd, tab) && irr_section_tab(tab, hw, nl, nm, nr, sec, runs)))
      irr_section(p.raw.irr_x + off, p.raw.irr_z + off, n, hw, p.raw.irr_left[nd], p.raw.irr_right[nd], nl, nm, nr, sec);
    return Qt - sec.K * sqrt(S0);
  }
  NodeVals nv;
  NodeConv kc;
  node_eval<false, RM, true, NormalDepthParams>(p.geo, p.N, nd, depth, 0.0, rg, p, nv, &kc);
  return Qt - kc.K * sqrt(S0);
}

template <int RM>
__global__ void __launch_bounds__(128) pr_normal_depth_kernel(const __grid_constant__ NormalDepthParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.M * p.N) return;
  const int m = (int)(i / p.N), nd = (int)(i % p.N);
  const Rough rg = load_rough<RM>(p.raw, m);
  const double Qt = p.q0[m * p.q0_stride], S0 = p.bed_slope[nd];
  const double z = p.geo[(size_t)F_Z * p.N + nd];
  auto f = [&](double hw) { return normal_flow_residual<RM>(p, nd, hw, Qt, S0, rg); };
  // Brent (scipy/optimize/Zeros/brentq.c)
  const double xtol = 2e-12, rtol = 8.881784197001252e-16;
  double xpre = z, xcur = z + 100.0, xblk = 0.0, fpre = f(xpre), fcur = f(xcur), fblk = 0.0, spre = 0.0, scur = 0.0;
  double root;
  bool done = false;
  if (fpre == 0.0) { root = xpre; done = true; }
  else if (fcur == 0.0) { root = xcur; done = true; }
  else if (signbit(fpre) == signbit(fcur)) {
    // brentq raises ValueError -> normal_depth's fallbacks (cross_section.py:196-202)
    root = (fpre < 0.0) ? z : (fcur > 0.0 ? z + 100.0 : z);
    done = true;
  }
  for (int it = 0; it < 100 && !done; ++it) {
    if (fpre != 0.0 && fcur != 0.0 && signbit(fpre) != signbit(fcur)) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
    if (fabs(fblk) < fabs(fcur)) { xpre = xcur; xcur = xblk; xblk = xpre; fpre = fcur; fcur = fblk; fblk = fpre; }
    const double delta = (xtol + rtol * fabs(xcur)) / 2.0, sbis = (xblk - xcur) / 2.0;
    if (fcur == 0.0 || fabs(sbis) < delta) { root = xcur; done = true; break; }
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
  if (!done) root = xcur;
  p.ic_h[i] = root - z;
  p.ic_q[i] = Qt;
}

// Accuracy probe of the FP64 primitives (tests/test_gpu_math.py): out[0..5][i] = fast_rcp, fast_sqrt, fast_rsqrt,
// fast_rcbrt of x[i] and the raw SFU seeds of 1/x and x^-1/2.
__global__ void pr_math_probe_kernel(const double* x, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = x[i];
  double s0, s1;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s0) : "d"(a));
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s1) : "d"(a));
  out[i] = fast_rcp(a);
  out[n + i] = fast_sqrt(a);
  out[2 * n + i] = fast_rsqrt(a);
  out[3 * n + i] = fast_rcbrt(a);
  out[4 * n + i] = s0;
  out[5 * n + i] = s1;
}

}  // namespace pr
