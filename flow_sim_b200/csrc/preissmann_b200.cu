// preissmann_b200.cu - C ABI (include/preissmann_b200.h) over the sm_100a kernels.
//
// Host side only: argument validation, staging of PR_MEM_HOST arrays through the device, collapsing the
// Roseires gate tables to two quadratics, choosing the lanes-per-member / nodes-per-lane instantiation,
// and launching.  No CPU implementation of the scheme lives here: if CUDA is unavailable every entry
// point fails with PR_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pr_aux_kernels.cuh"
#include "pr_ensemble_kernel.cuh"
#define PR_LONG_DECLARE_ONLY   // the long-reach kernels are compiled in pr_long_v*.cu
#include "pr_long_kernels.cuh"

#ifndef PR_W4
#define PR_W4 16
#endif
namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) return fail(PR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// Owns the device copies of one PR_MEM_HOST call.
struct Stage {
  bool host;
  cudaStream_t stream;
  std::vector<void*> allocs;
  struct Back { void* host; void* dev; size_t bytes; };
  std::vector<Back> backs;
  cudaError_t err = cudaSuccess;

  Stage(bool host_mem, cudaStream_t s) : host(host_mem), stream(s) {}
  ~Stage() {
    for (void* p : allocs) cudaFree(p);
  }
  template <class T>
  const T* in(const T* p, size_t n) {
    if (!p || !host) return p;
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, n * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    allocs.push_back(d);
    err = cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, stream);
    return static_cast<const T*>(d);
  }
  // source is host memory whatever the call's mem mode; the copy is synchronous (the source may be a temporary)
  template <class T>
  const T* in_host(const T* p, size_t n) {
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, n * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    allocs.push_back(d);
    err = cudaMemcpy(d, p, n * sizeof(T), cudaMemcpyHostToDevice);
    return static_cast<const T*>(d);
  }
  // device scratch that lives as long as the call's staging (freed with it)
  template <class T>
  T* scratch(size_t n) {
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, (n ? n : 1) * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    allocs.push_back(d);
    return static_cast<T*>(d);
  }
  template <class T>
  T* out(T* p, size_t n) {
    if (!p || !host) return p;
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, n * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    allocs.push_back(d);
    backs.push_back({p, d, n * sizeof(T)});
    return static_cast<T*>(d);
  }
  cudaError_t finish() {
    if (err != cudaSuccess) return err;
    for (auto& b : backs) {
      err = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, stream);
      if (err != cudaSuccess) return err;
    }
    if (host) err = cudaStreamSynchronize(stream);
    return err;
  }
};

// copy n doubles that live in `mem` space to the host
int fetch(const double* p, size_t n, int mem, std::vector<double>& out) {
  out.resize(n);
  if (mem == PR_MEM_HOST) {
    std::memcpy(out.data(), p, n * sizeof(double));
    return PR_OK;
  }
  CUDA_TRY(cudaMemcpy(out.data(), p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return PR_OK;
}

// Roseires: Q_state(s) = sum_{o_j>0} spill(s,o_j) + n_sl*sluice(s,twl) + q_hydro, as a quadratic in u = s - stage0
void collapse_roseires(const pr_rating& r, const double* openings, int sluices, long double out[3]) {
  long double k0 = r.q_hydro, k1 = 0, k2 = 0;
  const long double s0 = r.stage0;
  auto add = [&](const double c[6], long double o, long double w) {
    // c0 + c1 s + c2 o + c3 s^2 + c4 s o + c5 o^2 with s = s0 + u
    const long double a0 = c[0] + c[2] * o + c[5] * o * o, a1 = c[1] + c[4] * o, a2 = c[3];
    k0 += w * (a0 + a1 * s0 + a2 * s0 * s0);
    k1 += w * (a1 + 2 * a2 * s0);
    k2 += w * a2;
  };
  for (int j = 0; j < r.n_gates; ++j)
    if (openings[j] > 0) add(r.spill, openings[j], 1.0L);
  add(r.sluice, r.twl, (long double)sluices);
  out[0] = k0; out[1] = k1; out[2] = k2;
}

int make_rating(const pr_rating& r, pr::DevRating& d) {
  std::memset(&d, 0, sizeof d);
  d.type = r.type;
  d.n_coef = r.n_coef;
  d.a = r.a; d.b = r.b; d.c = r.c; d.shift = r.stage_shift;
  d.off = r.off; d.scl = r.scl;
  if (r.type == PR_RC_POLYNOMIAL) {
    if (r.n_coef < 1 || r.n_coef > PR_MAX_POLY) return fail(PR_ERR_ARG, "rating: n_coef=%d out of range", r.n_coef);
    for (int i = 0; i < r.n_coef; ++i) { d.coef[i] = r.coef[i]; d.dcoef[i] = r.dcoef[i]; }
  }
  if (r.type == PR_RC_ROSEIRES) {
    if (r.n_gates < 0 || r.n_gates > PR_MAX_GATES) return fail(PR_ERR_ARG, "rating: n_gates=%d out of range", r.n_gates);
    if (!(r.buffer > 0) || !(r.dY > 0)) return fail(PR_ERR_ARG, "rating: Roseires buffer and dY must be positive");
    long double lo[3], hi[3];
    collapse_roseires(r, r.closed_state, r.sluices_closed, lo);
    collapse_roseires(r, r.open_state, r.sluices_open, hi);
    for (int i = 0; i < 3; ++i) { d.lo[i] = (double)lo[i]; d.hi[i] = (double)hi[i]; d.dlt[i] = (double)(hi[i] - lo[i]); }
    d.stage0 = r.stage0; d.buffer = r.buffer; d.inv_buffer = 1.0 / r.buffer;
    d.dY = r.dY; d.inv_2dY = 1.0 / (2 * r.dY);
    d.gate_control = r.gate_control ? 1 : 0;
    d.initially_open = r.initially_open ? 1 : 0;
    d.max_cooldown = r.max_cooldown;
    if (d.gate_control && !(r.max_cooldown >= 0)) return fail(PR_ERR_ARG, "rating: max_cooldown must be >= 0");
  }
  if (r.type < PR_RC_NONE || r.type > PR_RC_ROSEIRES) return fail(PR_ERR_ARG, "rating: unknown type %d", r.type);
  return PR_OK;
}

int make_bc(const pr_bc& b, const char* which, bool downstream, const pr_config& cfg, double z_node, Stage& st,
            pr::DevBC& d) {
  std::memset(&d, 0, sizeof d);
  d.type = b.type;
  d.bed_level = b.bed_level;
  d.fixed_depth = b.fixed_depth;
  switch (b.type) {
    case PR_BC_FLOW_HYDROGRAPH:
    case PR_BC_STAGE_HYDROGRAPH: {
      if (!b.series) return fail(PR_ERR_ARG, "%s boundary: hydrograph series is NULL", which);
      if (b.series_member_stride != 0 && b.series_member_stride < cfg.n_levels)
        return fail(PR_ERR_ARG, "%s boundary: series_member_stride < n_levels", which);
      const size_t n = b.series_member_stride ? (size_t)b.series_member_stride * cfg.n_members : (size_t)cfg.n_levels;
      d.series = st.in(b.series, n);
      d.series_stride = b.series_member_stride;
      break;
    }
    case PR_BC_FIXED_DEPTH:
      break;
    case PR_BC_NORMAL_DEPTH:
      if (!(b.bed_slope == b.bed_slope)) return fail(PR_ERR_ARG, "%s boundary: normal_depth needs bed_slope", which);
      if (b.bed_level != z_node)
        return fail(PR_ERR_UNSUPPORTED, "%s boundary: normal_depth with bed_level (%g) != cross-section z_min (%g)",
                    which, b.bed_level, z_node);
      d.slope_factor = (b.bed_slope < 0 ? -1.0 : 1.0) * std::sqrt(std::fabs(b.bed_slope));
      break;
    case PR_BC_RATING_CURVE:
      if (b.rating.type == PR_RC_NONE) return fail(PR_ERR_ARG, "%s boundary: rating_curve without a curve", which);
      if (int rc = make_rating(b.rating, d.rc)) return rc;
      d.gated = d.rc.gate_control;
      if (b.member_ratings) {          // release scenarios: reduce every member's curve on the host, ship the array
        std::vector<pr::DevRating> all((size_t)cfg.n_members);
        for (int64_t m = 0; m < cfg.n_members; ++m)
          if (int rc = make_rating(b.member_ratings[m], all[(size_t)m])) return rc;
        d.gated = 0;
        for (const auto& r : all) d.gated |= r.gate_control;
        d.member_rc = st.in_host(all.data(), all.size());
      }
      if (d.gated && !downstream) return fail(PR_ERR_UNSUPPORTED, "gate-controlled rating curve at the upstream boundary");
      break;
    case PR_BC_FIXED_DEPTH_STORAGE:
      if (!downstream) return fail(PR_ERR_UNSUPPORTED, "lumped storage at the upstream boundary");
      d.st_min_stage = b.storage_min_stage;
      d.st_ymin = b.storage_ymin; d.st_ymax = b.storage_ymax;
      d.st_curve_len = b.storage_curve_len;
      if (b.storage_curve_len > 0) {
        if (b.storage_curve_len < 2 || !b.storage_curve_stage || !b.storage_curve_area)
          return fail(PR_ERR_ARG, "storage: area curve needs >= 2 rows and both columns");
        std::vector<double> stg;
        if (int rc = fetch(b.storage_curve_stage, (size_t)b.storage_curve_len, cfg.mem, stg)) return rc;
        double step = INFINITY;
        for (int i = 1; i < b.storage_curve_len; ++i) {
          if (!(stg[i] > stg[i - 1])) return fail(PR_ERR_ARG, "storage: area curve stages must be increasing");
          step = std::fmin(step, std::fabs(stg[i] - stg[i - 1]));     // lumped_storage.py:173
        }
        d.st_step = step;
        d.st_curve_stage = st.in(b.storage_curve_stage, (size_t)b.storage_curve_len);
        d.st_curve_area = st.in(b.storage_curve_area, (size_t)b.storage_curve_len);
        d.st_alpha = b.storage_alpha; d.st_beta = b.storage_beta;
      } else {
        if (!(b.storage_area > 0)) return fail(PR_ERR_ARG, "storage: surface area must be positive");
        d.st_area = b.storage_area;
        d.st_inv_area = 1.0 / b.storage_area;
      }
      if (b.storage_outflow.type != PR_RC_NONE) {
        if (b.storage_outflow.type == PR_RC_ROSEIRES) return fail(PR_ERR_UNSUPPORTED, "storage: gate-blend outflow curve");
        if (int rc = make_rating(b.storage_outflow, d.st_out)) return rc;
      }
      d.st_general = (b.storage_curve_len > 0 || b.storage_outflow.type != PR_RC_NONE) ? 1 : 0;
      if (d.st_general && !(b.storage_ymax > b.storage_ymin))
        return fail(PR_ERR_ARG, "storage: solution_boundaries (ymin < ymax) are required");
      d.st_losses = b.storage_capture_losses ? 1 : 0;
      if (d.st_losses) {
        if (b.bed_level != z_node)
          return fail(PR_ERR_UNSUPPORTED, "storage head losses with bed_level (%g) != cross-section z_min (%g)", b.bed_level, z_node);
        d.st_length = b.storage_reservoir_length; d.st_kq = b.storage_Kq;
      }
      break;
    default:
      return fail(PR_ERR_ARG, "%s boundary: unknown type %d", which, b.type);
  }
  return PR_OK;
}

// Every entry point runs on cfg->device and leaves the calling thread's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
  cudaError_t enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess || device < 0 || device == prev) return e;
    e = cudaSetDevice(device);
    changed = (e == cudaSuccess);
    return e;
  }
};

// Per-device facts and the ticket counters of the persistent fused kernel.  A launch takes the next counter of a ring
// and zeroes it on its stream, so launches on different streams never share one (the ring is far longer than any
// plausible number of launches in flight).
struct DeviceInfo {
  int sm_count = 0;
  unsigned int* tickets = nullptr;
  unsigned next = 0;
};
constexpr unsigned kTicketRing = 4096;
std::mutex g_dev_mu;
DeviceInfo g_dev[64];

int device_info(cudaStream_t s, int& sm_count, unsigned int*& ticket) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(PR_ERR_CUDA, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.tickets) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaMalloc(&d.tickets, kTicketRing * sizeof(unsigned int)));
  }
  sm_count = d.sm_count;
  ticket = d.tickets + (d.next++ % kTicketRing);
  CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s));
  return PR_OK;
}

int check_config(const pr_config* cfg, DeviceGuard& guard) {
  if (!cfg) return fail(PR_ERR_ARG, "cfg is NULL");
  if (cfg->abi_version != PR_ABI_VERSION)
    return fail(PR_ERR_ARG, "abi_version %d != library %d", cfg->abi_version, PR_ABI_VERSION);
  if (cfg->n_nodes < 2) return fail(PR_ERR_ARG, "n_nodes must be >= 2");
  if (cfg->n_levels < 1) return fail(PR_ERR_ARG, "n_levels must be >= 1");
  if (cfg->n_members < 1) return fail(PR_ERR_ARG, "n_members must be >= 1");
  if (cfg->mem != PR_MEM_HOST && cfg->mem != PR_MEM_DEVICE) return fail(PR_ERR_ARG, "mem must be HOST or DEVICE");
  if (!(cfg->dt > 0) || !(cfg->dx > 0)) return fail(PR_ERR_ARG, "dt and dx must be positive");
  CUDA_TRY(guard.enter(cfg->device));
  return PR_OK;
}

int stage_geom(const pr_config& cfg, const pr_geom* g, Stage& st, pr::DevGeom& d) {
  if (!g) return fail(PR_ERR_ARG, "geom is NULL");
  const void* req[] = {g->kind, g->z_bed, g->b_main, g->m_main, g->h_bank, g->T_bank, g->W_bank,
                       g->b_fp_l, g->b_fp_r, g->m_fp, g->n_l, g->n_m, g->n_r, g->curvature};
  for (const void* p : req)
    if (!p) return fail(PR_ERR_ARG, "geom: a per-node array is NULL");
  const size_t N = cfg.n_nodes, M = cfg.n_members;
  d.kind = st.in(g->kind, N);
  d.z = st.in(g->z_bed, N); d.b = st.in(g->b_main, N); d.m = st.in(g->m_main, N);
  d.hb = st.in(g->h_bank, N); d.Tb = st.in(g->T_bank, N); d.Wb = st.in(g->W_bank, N);
  d.bl = st.in(g->b_fp_l, N); d.br = st.in(g->b_fp_r, N); d.mfp = st.in(g->m_fp, N);
  d.nl = st.in(g->n_l, N); d.nm = st.in(g->n_m, N); d.nr = st.in(g->n_r, N);
  d.curv = st.in(g->curvature, N);
  d.member_nm = st.in(g->member_n_main, M);
  d.member_nfp = st.in(g->member_n_fp, M);
  d.irr_offset = nullptr; d.irr_x = d.irr_z = d.irr_left = d.irr_right = nullptr;
  d.irr_tab_n = nullptr; d.irr_tab_z = d.irr_tab_c = nullptr; d.irr_tab_runs = nullptr;
  if (g->irr_offset) {                 // IrregularSection polylines (CSR)
    if (!g->irr_x || !g->irr_z || !g->irr_left || !g->irr_right) return fail(PR_ERR_ARG, "geom: irregular-section arrays are incomplete");
    int32_t total = 0;
    if (cfg.mem == PR_MEM_HOST) total = g->irr_offset[N];
    else CUDA_TRY(cudaMemcpy(&total, g->irr_offset + N, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (total < 0) return fail(PR_ERR_ARG, "geom: irr_offset is not a prefix sum");
    d.irr_offset = st.in(g->irr_offset, N + 1);
    d.irr_x = st.in(g->irr_x, (size_t)total); d.irr_z = st.in(g->irr_z, (size_t)total);
    d.irr_left = st.in(g->irr_left, N); d.irr_right = st.in(g->irr_right, N);
    // stage tables of the polyline sections (one interval search instead of six polyline scans per node evaluation);
    // PR_IRR_NO_TABLES=1 keeps the scanning node pass (A/B and debugging)
    static const bool no_tables = std::getenv("PR_IRR_NO_TABLES") != nullptr;
    if (no_tables) return PR_OK;
    int* tab_n = st.scratch<int>(N);
    double* tab_z = st.scratch<double>((size_t)total);
    double* tab_c = st.scratch<double>((size_t)total * pr::kIrrTabCols);
    int* tab_runs = st.scratch<int>((size_t)total);
    if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
    pr::pr_irr_build_tables<<<(unsigned)((N + 63) / 64), 64, 0, st.stream>>>(d, (int)N, tab_n, tab_z, tab_c, tab_runs);
    g_launches.fetch_add(1);
    d.irr_tab_n = tab_n; d.irr_tab_z = tab_z; d.irr_tab_c = tab_c; d.irr_tab_runs = tab_runs;
  }
  return PR_OK;
}

// Tuning hook (tools/ab_build.sh): PR_M4_VARIANT=<variant .so> replaces the 32-lane x 4-node family of this library by
// the one in that file, so that A/B builds of the headline kernel need not carry the whole library.
using m4_variant_fn = int (*)(const pr::DevParams*, int, cudaStream_t);
m4_variant_fn m4_variant() {
  static m4_variant_fn fn = []() -> m4_variant_fn {
    const char* path = std::getenv("PR_M4_VARIANT");
    if (!path || !*path) return nullptr;
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "PR_M4_VARIANT: %s\n", dlerror()); std::abort(); }
    void* f = dlsym(h, "pr_variant_launch_m4");
    if (!f) { std::fprintf(stderr, "PR_M4_VARIANT: no pr_variant_launch_m4 in %s\n", path); std::abort(); }
    return reinterpret_cast<m4_variant_fn>(f);
  }();
  return fn;
}

int launch_family(int rc_cuda) {
  if (rc_cuda != 0) return fail(PR_ERR_CUDA, "ensemble kernel launch: %s", cudaGetErrorString((cudaError_t)rc_cuda));
  g_launches.fetch_add(1);
  return PR_OK;
}

template <bool CURV, int RM>
int launch_gvf(const pr::GvfParams& p, unsigned grid, size_t smem, cudaStream_t s) {
  CUDA_TRY(cudaFuncSetAttribute(pr::pr_gvf_kernel<CURV, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pr::pr_gvf_kernel<CURV, RM><<<grid, 128, smem, s>>>(p);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PR_OK;
}

template <bool CURV>
int launch_gvf_rm(const pr::GvfParams& p, unsigned grid, size_t smem, cudaStream_t s) {
  switch (pr::rough_mode(p.geo)) {
    case 0: return launch_gvf<CURV, 0>(p, grid, smem, s);
    case 1: return launch_gvf<CURV, 1>(p, grid, smem, s);
    case 2: return launch_gvf<CURV, 2>(p, grid, smem, s);
    default: return launch_gvf<CURV, 3>(p, grid, smem, s);
  }
}

template <int RM>
int launch_normal_depth(const pr::NormalDepthParams& p, cudaStream_t s) {
  const long long total = (long long)p.M * p.N;
  pr::pr_normal_depth_kernel<RM><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(p);
  CUDA_TRY(cudaGetLastError());
  return PR_OK;
}

}  // namespace

extern "C" {

int pr_abi_version(void) { return PR_ABI_VERSION; }

const char* pr_last_error(void) { return g_err.c_str(); }

int64_t pr_launch_count(void) { return g_launches.load(); }

int pr_ensemble_run(const pr_config* cfg, const pr_geom* geom, const pr_bc* upstream, const pr_bc* downstream,
                    const pr_state* initial, const pr_outputs* out, void* cuda_stream) {
  DeviceGuard guard;
  if (int rc = check_config(cfg, guard)) return rc;
  if (!upstream || !downstream || !initial || !out) return fail(PR_ERR_ARG, "a struct pointer is NULL");
  if (!initial->depth || !initial->flow) return fail(PR_ERR_ARG, "initial conditions are NULL");
  if (initial->member_stride != 0 && initial->member_stride < cfg->n_nodes)
    return fail(PR_ERR_ARG, "initial.member_stride < n_nodes");
  if (cfg->out_mode != PR_OUT_FULL && cfg->out_mode != PR_OUT_UPSTREAM) return fail(PR_ERR_ARG, "bad out_mode");
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t N = cfg->n_nodes, L = cfg->n_levels, M = cfg->n_members;
  Stage st(cfg->mem == PR_MEM_HOST, s);

  pr::DevParams p;
  std::memset(&p, 0, sizeof p);
  p.N = (int)N; p.L = (int)L; p.M = (int)M; p.max_iter = cfg->max_iter; p.out_mode = cfg->out_mode;
  p.theta = cfg->theta; p.dt = cfg->dt; p.dx = cfg->dx; p.tol = cfg->tol; p.tol2 = cfg->tol * cfg->tol; p.g = cfg->g;
  p.i2dt = 1.0 / (2.0 * cfg->dt);
  p.th_dx = cfg->theta / cfg->dx;
  p.hth = 0.5 * cfg->theta;
  p.omt_dx = (1.0 - cfg->theta) / cfg->dx;
  p.homt = 0.5 * (1.0 - cfg->theta);
  p.ghth = cfg->g * 0.5 * cfg->theta;
  p.th_dx2 = 2.0 * cfg->theta / cfg->dx;
  p.mtheta = -cfg->theta;

  // host-side look at the few geometry values the dispatch needs
  std::vector<double> curv, zb;
  if (int rc = (geom && geom->curvature && geom->z_bed) ? PR_OK : fail(PR_ERR_ARG, "geom is incomplete")) return rc;
  if (int rc = fetch(geom->curvature, N, cfg->mem, curv)) return rc;
  if (int rc = fetch(geom->z_bed, N, cfg->mem, zb)) return rc;
  bool has_curv = false;
  for (double c : curv) has_curv |= (c != 0.0);
  bool has_compound = false, has_irregular = false;
  {
    std::vector<int32_t> kinds(N);
    if (cfg->mem == PR_MEM_HOST) std::memcpy(kinds.data(), geom->kind, N * sizeof(int32_t));
    else CUDA_TRY(cudaMemcpy(kinds.data(), geom->kind, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int32_t kd : kinds) {
      if (kd < PR_XS_RECT || kd > PR_XS_IRREGULAR) return fail(PR_ERR_ARG, "geom.kind holds an unknown section kind %d", kd);
      has_compound |= (kd == PR_XS_COMPOUND);
      has_irregular |= (kd == PR_XS_IRREGULAR);
    }
    if (has_irregular && !geom->irr_offset) return fail(PR_ERR_ARG, "geom: irregular sections without their polylines");
  }

  if (int rc = stage_geom(*cfg, geom, st, p.geo)) return rc;
  if (int rc = make_bc(*upstream, "upstream", false, *cfg, zb[0], st, p.up)) return rc;
  if (int rc = make_bc(*downstream, "downstream", true, *cfg, zb[N - 1], st, p.dn)) return rc;

  const size_t icn = initial->member_stride ? (size_t)initial->member_stride * M : N;
  p.ic_h = st.in(initial->depth, icn);
  p.ic_q = st.in(initial->flow, icn);
  p.ic_stride = initial->member_stride;
  const size_t on = (cfg->out_mode == PR_OUT_FULL) ? M * L * N : M * L;
  p.out_h = st.out(out->depth, on);
  p.out_q = st.out(out->flow, on);
  p.iters = st.out(out->iters, M * (L > 1 ? L - 1 : 1));
  p.status = st.out(out->status, M);
  p.fail_level = st.out(out->fail_level, M);
  p.storage_stage = st.out(out->storage_stage, M * L);
  p.final_error = st.out(out->final_error, M * (L > 1 ? L - 1 : 1));
  p.member_order = st.in(cfg->member_order, M);
  if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
  if (int rc = device_info(s, p.sm_count, p.ticket)) return rc;

  // Lanes per member G and nodes per lane M: the chain has ceil(cells / M) + 1 <= G block rows.  Short reaches pack
  // 4 or 2 members into a warp (G = 8, 16) so that neither lanes nor parallel-cyclic-reduction steps are wasted.
  const int cells = (int)N - 1;
  int lpm = cfg->lanes_per_member;
  if (lpm == 0)                         // tuning: PR_FORCE_LANES=8|16|32 overrides the automatic choice
    if (const char* e = std::getenv("PR_FORCE_LANES")) lpm = std::atoi(e);
  if (lpm != 0 && lpm != -1 && lpm != 8 && lpm != 16 && lpm != 32)
    return fail(PR_ERR_ARG, "lanes_per_member=%d: must be 0 (auto), 8, 16, 32 or -1 (long-reach path)", lpm);
  auto fits = [&](int G, int M) { return (lpm == 0 || lpm == G) && cells <= (G - 1) * M; };
  int rc;
  // warps per CTA: one CTA per SM, as many warps as registers (65536 / (32 * regs)) and shared memory allow
  if (has_irregular && lpm != -1 && cells <= 31 * 8) {        // polyline node pass inside the fused kernel
    // short polyline reaches pack 4 / 2 members into a warp (8 / 16 lanes x 2 nodes: up to 15 / 31 nodes)
    if (fits(8, 2)) rc = launch_family(pr::launch_ensemble_irregular<8, 2>(p, has_curv, s));
    else if (fits(16, 2)) rc = launch_family(pr::launch_ensemble_irregular<16, 2>(p, has_curv, s));
    else if (cells <= 31 * 2 && (lpm == 0 || lpm == 32)) rc = launch_family(pr::launch_ensemble_irregular<32, 2>(p, has_curv, s));
    else if (cells <= 31 * 4 && (lpm == 0 || lpm == 32)) rc = launch_family(pr::launch_ensemble_irregular<32, 4>(p, has_curv, s));
    else if (lpm == 0 || lpm == 32) rc = launch_family(pr::launch_ensemble_irregular<32, 8>(p, has_curv, s));
    else return fail(PR_ERR_UNSUPPORTED, "lanes_per_member=%d: no polyline instantiation holds %d nodes", lpm, (int)N);
  } else if (has_irregular) rc = pr::long_reach_run(p, has_curv, true, true, s, g_launches, g_err);   // ... or the tile kernels
  else if (lpm == -1) rc = pr::long_reach_run(p, has_curv, has_compound, false, s, g_launches, g_err);   // forced long-reach path
  else if (fits(8, 1)) rc = launch_family(pr::launch_ensemble_family<8, 1, 16>(p, has_curv, s));
  else if (fits(8, 2)) rc = launch_family(pr::launch_ensemble_family<8, 2, 16>(p, has_curv, s));
  else if (fits(8, 4)) rc = launch_family(pr::launch_ensemble_family<8, 4, 16>(p, has_curv, s));
  else if (fits(16, 2)) rc = launch_family(pr::launch_ensemble_family<16, 2, 16>(p, has_curv, s));
  else if (fits(32, 1)) rc = launch_family(pr::launch_ensemble_family<32, 1, 16>(p, has_curv, s));
  else if (fits(16, 4)) rc = launch_family(pr::launch_ensemble_family<16, 4, 16>(p, has_curv, s));
  else if (fits(32, 2)) rc = launch_family(pr::launch_ensemble_family<32, 2, 16>(p, has_curv, s));
  else if (fits(32, 4)) rc = launch_family(m4_variant() ? m4_variant()(&p, has_curv, s) : pr::launch_ensemble_family<32, 4, PR_W4>(p, has_curv, s));
  else if (fits(32, 8)) rc = launch_family(pr::launch_ensemble_family<32, 8, 7>(p, has_curv, s));
  else if (lpm != 0) return fail(PR_ERR_UNSUPPORTED, "lanes_per_member=%d: no instantiation holds %d nodes", lpm, (int)N);
  else rc = pr::long_reach_run(p, has_curv, has_compound, false, s, g_launches, g_err);
  if (rc) return rc;
  cudaError_t e = st.finish();
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_ensemble_run: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_gvf_initial_conditions(const pr_config* cfg, const pr_geom* geom, const double* q0, int64_t q0_member_stride,
                              const double* downstream_depth, int64_t downstream_depth_member_stride,
                              double* ic_depth, double* ic_flow, int32_t* status, void* cuda_stream) {
  DeviceGuard guard;
  if (int rc = check_config(cfg, guard)) return rc;
  if (!q0 || !downstream_depth || !ic_depth || !ic_flow) return fail(PR_ERR_ARG, "q0 / downstream_depth / ic buffers are NULL");

  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t N = cfg->n_nodes, M = cfg->n_members;
  Stage st(cfg->mem == PR_MEM_HOST, s);
  pr::GvfParams p;
  std::memset(&p, 0, sizeof p);
  p.N = (int)N; p.M = (int)M; p.dx = cfg->dx; p.g = cfg->g;
  std::vector<double> curv;
  if (!geom || !geom->curvature) return fail(PR_ERR_ARG, "geom is incomplete");
  if (int rc = fetch(geom->curvature, N, cfg->mem, curv)) return rc;
  bool has_curv = false;
  for (double c : curv) has_curv |= (c != 0.0);
  if (int rc = stage_geom(*cfg, geom, st, p.geo)) return rc;
  p.q0 = st.in(q0, q0_member_stride ? M * (size_t)q0_member_stride : 1);
  p.q0_stride = q0_member_stride;
  p.h_down = st.in(downstream_depth, downstream_depth_member_stride ? M * (size_t)downstream_depth_member_stride : 1);
  p.h_down_stride = downstream_depth_member_stride;
  p.ic_h = st.out(ic_depth, M * N);
  p.ic_q = st.out(ic_flow, M * N);
  p.status = st.out(status, M);
  if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
  const size_t smem = sizeof(double) * pr::F_COUNT * N;
  if (smem > 200 * 1024) {          // a reach too long to stage per CTA: derived geometry table in global memory
    double* table = nullptr;
    CUDA_TRY(cudaMalloc(&table, smem));
    st.allocs.push_back(table);
    pr::pr_long_geometry<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(p.geo, (int)N, table);
    g_launches.fetch_add(1);
    p.table = table;
  }
  const unsigned grid = (unsigned)((M + 127) / 128);
  const size_t gvf_smem = p.table ? 0 : smem;
  if (int rc = has_curv ? launch_gvf_rm<true>(p, grid, gvf_smem, s) : launch_gvf_rm<false>(p, grid, gvf_smem, s)) return rc;
  cudaError_t e = st.finish();
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_gvf_initial_conditions: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_rating_objective(const pr_config* cfg, const double* up_flow, const double* up_depth, double z0,
                        const double* q_query, const double* h_target, int32_t n_query, double* levels_out,
                        double* rmse_out, void* cuda_stream) {
  DeviceGuard guard;
  if (int rc = check_config(cfg, guard)) return rc;
  if (!up_flow || !up_depth || !q_query || !h_target || n_query < 1) return fail(PR_ERR_ARG, "objective: NULL input");
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t L = cfg->n_levels, M = cfg->n_members;
  Stage st(cfg->mem == PR_MEM_HOST, s);
  pr::ObjParams p;
  p.L = (int)L; p.M = (int)M; p.nq = n_query; p.z0 = z0;
  p.up_q = st.in(up_flow, M * L);
  p.up_h = st.in(up_depth, M * L);
  p.q_query = st.in(q_query, (size_t)n_query);
  p.h_target = st.in(h_target, (size_t)n_query);
  p.levels = st.out(levels_out, M * (size_t)n_query);
  p.rmse = st.out(rmse_out, M);
  if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
  pr::pr_objective_kernel<<<(unsigned)((M + 127) / 128), 128, 0, s>>>(p);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  cudaError_t e = st.finish();
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_rating_objective: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_derived_results(const pr_config* cfg, const pr_geom* geom, const double* depth, const double* flow,
                       double* level, double* area, double* top_width, double* froude, double* velocity,
                       double* celerity, void* cuda_stream) {
  DeviceGuard guard;
  if (int rc = check_config(cfg, guard)) return rc;
  if (!depth || !flow) return fail(PR_ERR_ARG, "derived results: depth / flow are NULL");
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t N = cfg->n_nodes, total = (size_t)cfg->n_members * cfg->n_levels * N;
  Stage st(cfg->mem == PR_MEM_HOST, s);
  pr::DevGeom dg;
  if (int rc = stage_geom(*cfg, geom, st, dg)) return rc;
  pr::DerivedParams p;
  p.rows = (long long)cfg->n_members * cfg->n_levels; p.N = (int)N; p.g = cfg->g;
  p.depth = st.in(depth, total); p.flow = st.in(flow, total);
  p.level = st.out(level, total); p.area = st.out(area, total); p.top_width = st.out(top_width, total);
  p.froude = st.out(froude, total); p.velocity = st.out(velocity, total); p.celerity = st.out(celerity, total);
  double* table = nullptr;
  if (st.err == cudaSuccess) st.err = cudaMalloc(&table, sizeof(double) * pr::F_COUNT * N);
  if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
  st.allocs.push_back(table);
  p.geo = table;
  p.raw = dg;
  pr::pr_long_geometry<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(dg, (int)N, table);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned gx = (unsigned)((N + 127) / 128);
  long long gy = ((long long)sms * 16 + gx - 1) / gx;          // ~16 CTAs of 4 warps per SM in total
  if (gy > p.rows) gy = p.rows;
  if (gy > 65535) gy = 65535;
  pr::pr_derived_kernel<<<dim3(gx, (unsigned)gy), 128, 0, s>>>(p);
  g_launches.fetch_add(2);
  CUDA_TRY(cudaGetLastError());
  cudaError_t e = st.finish();
  if (cfg->mem == PR_MEM_DEVICE && e == cudaSuccess) e = cudaStreamSynchronize(s);   // the table is freed on return
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_derived_results: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_normal_depth_initial_conditions(const pr_config* cfg, const pr_geom* geom, const double* bed_slope,
                                       const double* q0, int64_t q0_member_stride, double* ic_depth,
                                       double* ic_flow, void* cuda_stream) {
  DeviceGuard guard;
  if (int rc = check_config(cfg, guard)) return rc;
  if (!bed_slope || !q0 || !ic_depth || !ic_flow) return fail(PR_ERR_ARG, "normal depth: NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t N = cfg->n_nodes, M = cfg->n_members;
  Stage st(cfg->mem == PR_MEM_HOST, s);
  pr::NormalDepthParams p;
  std::memset(&p, 0, sizeof p);
  p.N = (int)N; p.M = (int)M; p.g = cfg->g;
  if (int rc = stage_geom(*cfg, geom, st, p.raw)) return rc;
  p.bed_slope = st.in(bed_slope, N);
  p.q0 = st.in(q0, q0_member_stride ? M * (size_t)q0_member_stride : 1);
  p.q0_stride = q0_member_stride;
  p.ic_h = st.out(ic_depth, M * N);
  p.ic_q = st.out(ic_flow, M * N);
  double* table = nullptr;
  if (st.err == cudaSuccess) st.err = cudaMalloc(&table, sizeof(double) * pr::F_COUNT * N);
  if (st.err != cudaSuccess) return fail(PR_ERR_CUDA, "staging: %s", cudaGetErrorString(st.err));
  st.allocs.push_back(table);
  p.geo = table;
  pr::pr_long_geometry<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(p.raw, (int)N, table);
  int rc;
  switch (pr::rough_mode(p.raw)) {
    case 0: rc = launch_normal_depth<0>(p, s); break;
    case 1: rc = launch_normal_depth<1>(p, s); break;
    case 2: rc = launch_normal_depth<2>(p, s); break;
    default: rc = launch_normal_depth<3>(p, s);
  }
  if (rc) return rc;
  g_launches.fetch_add(2);
  cudaError_t e = st.finish();
  if (cfg->mem == PR_MEM_DEVICE && e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_normal_depth_initial_conditions: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_release_workspace(void) {
  pr::long_pool().release_all();
  return PR_OK;
}

int64_t pr_long_last_trips(void) {
  pr::LongPool& pool = pr::long_pool();
  const int* ctr = nullptr;
  int dev = -1;
  {
    std::lock_guard<std::mutex> lock(pool.mu);
    ctr = pool.last_trips; dev = pool.last_trips_device;
  }
  if (!ctr) return -1;
  DeviceGuard guard;
  if (guard.enter(dev) != cudaSuccess) return -1;
  int trips = 0;
  if (cudaMemcpy(&trips, ctr, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;   // waits for the run
  return trips;
}

int pr_math_probe(const double* x_host, int32_t n, double* out_host) {
  if (!x_host || !out_host || n < 1) return fail(PR_ERR_ARG, "pr_math_probe: NULL / empty input");
  double *dx = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc(&dx, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&dout, sizeof(double) * 6 * (size_t)n));
  CUDA_TRY(cudaMemcpy(dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice));
  pr::pr_math_probe_kernel<<<(n + 255) / 256, 256>>>(dx, n, dout);
  g_launches.fetch_add(1);
  cudaError_t e = cudaMemcpy(out_host, dout, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, "pr_math_probe: %s", cudaGetErrorString(e));
  return PR_OK;
}

int pr_fp64_peak(double millis, double* tflops_out) {
  if (!tflops_out) return fail(PR_ERR_ARG, "tflops_out is NULL");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int grid = sms * 8, block = 256;
  int iters = 1 << 14;
  double best = 0.0, spent = 0.0;
  pr::pr_dfma_kernel<<<grid, block>>>(sink, 1024, 1.0000001, 1e-9);   // warm-up
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  while (spent < millis) {
    CUDA_TRY(cudaEventRecord(e0));
    pr::pr_dfma_kernel<<<grid, block>>>(sink, iters, 1.0000001, 1e-9);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    spent += ms;
    const double flops = 2.0 * 8.0 * (double)iters * grid * block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    if (ms < 5.0) iters *= 2;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops_out = best;
  return PR_OK;
}

}  // extern "C"
