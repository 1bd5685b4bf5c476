// pr_aux_kernels.cuh - the small kernels either side of the Newton loop:
//   * GVF backwater initial conditions, one member per thread (Channel._gvh_conditions, channel.py:307-378)
//   * calibration objective np.interp + RMSE, one member per thread (model.py:105-113, n_calibrate.py:55-63)
//   * DFMA peak probe (roofline denominator)
#pragma once
#include "pr_device.cuh"

namespace pr {

struct GvfParams {
  int N, M;
  double dx, g, h_down;
  double th_dx, hth, th_dx2;   // unused scheme constants node_eval reads (zero)
  DevGeom geo;
  const double* q0;
  long long q0_stride;
  double *ic_h, *ic_q;
  int* status;
};

// dh/dx of the gradually-varied-flow equation at one node (get_dh_dx, channel.py:316-347)
template <bool CURV, int RM>
__device__ __forceinline__ double gvf_slope(const double* sg, int NP, int node, double h_in, double Q, double S0,
                                            const Rough& rg, const GvfParams& p, int& status) {
  const double g = p.g;
  NodeVals nv;
  node_eval<CURV, RM, false, GvfParams>(sg, NP, node, h_in, Q, rg, p, nv);
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
  extern __shared__ double smem[];
  const int NP = p.N;
  stage_geometry(p.geo, p.N, NP, smem, threadIdx.x, blockDim.x, [](int idx) { return idx; });
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const Rough rg = load_rough<RM>(p.geo, m);
  const double Q = p.q0[m * p.q0_stride];
  const int N = p.N;
  double* oh = p.ic_h + (size_t)m * N;
  double* oq = p.ic_q + (size_t)m * N;
  int status = PR_STATUS_OK;
  double h = p.h_down;
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
