/*
 * preissmann_b200.h - C ABI of the B200-native Preissmann (implicit Saint-Venant) ensemble solver.
 *
 * The reference (cve-mohd/flow-sim, package `hydromodel`) is pure Python and has NO FFI / plugin
 * interface; its narrowest operator seam is the body of PreissmannSolver.run()
 * (src/hydromodel/preissmann.py:101-163).  This ABI takes over that whole method for an ENSEMBLE of
 * members that share a reach geometry: plain pointers and sizes, no torch types, caller owns every
 * buffer.  Each entry point cites the reference interface it replaces.  The reference-side binding
 * (a ctypes stub inside PreissmannSolver.run) is shown in INTEGRATION.md.
 *
 * All floating-point data is IEEE double (the reference computes in Python float / numpy float64).
 * Arrays may live in host memory (PR_MEM_HOST: the library stages them through the device itself,
 * H2D/D2H included) or in device memory (PR_MEM_DEVICE: zero-copy, kernels only).
 */
#ifndef PREISSMANN_B200_H
#define PREISSMANN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PR_ABI_VERSION 7
#define PR_MAX_POLY 12   /* max coefficients of a fitted numpy Polynomial rating curve */
#define PR_MAX_GATES 8   /* Roseires: 7 spillway gates (roseires_rating_curve.py:11) */

/* ---- return codes (argument / CUDA errors; numerical failures are reported per member) ---- */
#define PR_OK 0
#define PR_ERR_ARG 1          /* bad argument; pr_last_error() has the text */
#define PR_ERR_UNSUPPORTED 2  /* a configuration the device path does not implement (no CPU fallback) */
#define PR_ERR_CUDA 3         /* CUDA runtime error; pr_last_error() has cudaGetErrorString */

/* ---- per-member status[] (reference: preissmann.py:124-126 raises ValueError; :133-137 NaN) ---- */
#define PR_STATUS_OK 0
#define PR_STATUS_MAX_ITER 1      /* "Convergence within N iterations couldn't be achieved." */
#define PR_STATUS_NAN 2           /* residual norm became NaN/Inf before max_iter */
#define PR_STATUS_SUPERCRITICAL 3 /* GVF initial condition: flow became supercritical (channel.py:328-332) */

/* cross_section.py:636-674 - which branch of TrapezoidalSection.properties applies */
enum pr_section_kind { PR_XS_RECT = 0, PR_XS_TRAPEZOID = 1, PR_XS_COMPOUND = 2, PR_XS_IRREGULAR = 3 };

/* boundary.py:32 - condition names; FIXED_DEPTH_STORAGE = 'fixed_depth' + set_lumped_storage() */
enum pr_bc_type {
  PR_BC_FLOW_HYDROGRAPH = 0,
  PR_BC_FIXED_DEPTH = 1,
  PR_BC_NORMAL_DEPTH = 2,
  PR_BC_RATING_CURVE = 3,
  PR_BC_STAGE_HYDROGRAPH = 4,
  PR_BC_FIXED_DEPTH_STORAGE = 5
};

/* rating_curve.py:10-63,132-147 and cases/gerd_roseires/roseires_rating_curve.py:65-109,180-208 */
enum pr_rc_type {
  PR_RC_NONE = 0,
  PR_RC_POLY2 = 1,      /* a x^2 + b x + c,   x = stage + stage_shift   (RatingCurve.set('polynomial')) */
  PR_RC_POWER = 2,      /* a x^b                                         (RatingCurve.set('power'))      */
  PR_RC_POLYNOMIAL = 3, /* numpy Polynomial.fit: sum coef[i] (off + scl x)^i (RatingCurve.fit(scale=True)) */
  PR_RC_ROSEIRES = 4    /* smooth gate-blend of two sklearn degree-2 bivariate fits, FD derivative        */
};

enum pr_out_mode {
  PR_OUT_FULL = 0,     /* depth/flow are [M][levels][N]  (solver.py:43-44 for every member)              */
  PR_OUT_UPSTREAM = 1  /* depth/flow are [M][levels]: node 0 only (what model.py:105-113 consumes)       */
};

enum pr_mem { PR_MEM_HOST = 0, PR_MEM_DEVICE = 1 };

/* Solver configuration.  Reference: Solver.__init__ (solver.py:11-55), PreissmannSolver.__init__
 * (preissmann.py:23-46) and the kwargs of run() (preissmann.py:101). */
typedef struct pr_config {
  int32_t abi_version;  /* must be PR_ABI_VERSION */
  int32_t n_nodes;      /* N  = round(L/dx)+1            (solver.py:53-55) */
  int32_t n_levels;     /* T//dt + 1                     (solver.py:35)    */
  int32_t n_members;    /* ensemble size M (1 = the reference's single run) */
  int32_t max_iter;     /* run(max_iter=100) */
  int32_t out_mode;     /* enum pr_out_mode */
  int32_t mem;          /* enum pr_mem: where EVERY array of this call lives */
  int32_t device;       /* CUDA device ordinal, -1 = current device */
  int32_t lanes_per_member; /* 0 = auto: fused kernel with 8 / 16 / 32 lanes per member (4 / 2 / 1 members per warp)
                               for reaches of up to 29 / 61 / 249 nodes, tiled long-reach path above;
                               8, 16, 32 = force that group width; -1 = force the long-reach path (testing) */
  int32_t reserved0;
  double theta;         /* Preissmann weighting factor */
  double dt;            /* time_step [s] */
  double dx;            /* fitted spatial_step = L/(N-1) [m] */
  double tol;           /* run(tolerance): absolute bound on ||R||_2 (preissmann.py:149-153) */
  double g;             /* scipy.constants.g = 9.80665 */
  /* Optional processing order of the members in pr_ensemble_run: a permutation of 0..M-1 ([M], in `mem` space), or
   * NULL = 0, 1, 2, ...  The fused kernel is persistent - warps draw members from a counter until the ensemble is
   * used up - so handing out the expensive members first (e.g. a roughness sweep in descending n: the Newton
   * iteration total grows with n) shortens the tail of the launch.  Results are always written in member order. */
  const int32_t* member_order;
} pr_config;

/* Per-node cross-sections in SoA form, i.e. Channel.xs_at_node after interpolation
 * (channel.py:213-241, cross_section.py:857-930).  Every pointer has length n_nodes. */
typedef struct pr_geom {
  const int32_t* kind;   /* enum pr_section_kind */
  const double* z_bed;   /* z_bed == z_min */
  const double* b_main;
  const double* m_main;
  const double* h_bank;  /* bankfull_depth = z_bank - z_bed  (cross_section.py:591); 0 if not compound */
  const double* T_bank;  /* T_main_at_bank                   (cross_section.py:592) */
  const double* W_bank;  /* _width_at_bank                   (cross_section.py:597) */
  const double* b_fp_l;
  const double* b_fp_r;
  const double* m_fp;
  const double* n_l;     /* n_left  */
  const double* n_m;     /* n_main  */
  const double* n_r;     /* n_right */
  const double* curvature;
  /* Ensemble roughness overrides (model.run(n_main=, n_fp=), custom_functions.py:128-157): when
   * member_n_main != NULL, member m uses n_m[i] = v*w1[i] + v*w2[i] with v = member_n_main[m], which
   * is what interpolate_cross_section computes from two sections sharing v (cross_section.py:891-893).
   * Likewise member_n_fp overrides n_l and n_r. */
  const double* w1;
  const double* w2;
  const double* member_n_main; /* [M] or NULL */
  const double* member_n_fp;   /* [M] or NULL */
  /* IrregularSection nodes (kind = PR_XS_IRREGULAR, cross_section.py:207-543): the polyline of node i is
   * irr_x/irr_z[irr_offset[i] .. irr_offset[i+1]) sorted by x (empty for the other kinds); z_bed[i] = min z;
   * irr_left/irr_right[i] = left_fp_limit / right_fp_limit of the composite roughness (n_l, n_m, n_r as above).
   * All NULL when the reach has no irregular section. */
  const int32_t* irr_offset;   /* [N+1] */
  const double* irr_x;
  const double* irr_z;
  const double* irr_left;      /* [N] */
  const double* irr_right;     /* [N] */
} pr_geom;

typedef struct pr_rating {
  int32_t type;    /* enum pr_rc_type */
  int32_t n_coef;  /* POLYNOMIAL: number of coefficients */
  double a, b, c, stage_shift;                 /* POLY2 / POWER */
  double coef[PR_MAX_POLY], dcoef[PR_MAX_POLY]; /* POLYNOMIAL: p and p' series in the mapped variable */
  double off, scl;                              /* POLYNOMIAL: u = off + scl*x (Polynomial.mapparms) */
  /* ROSEIRES, gate_control = 0 (smooth=True, the default): Q = (1-a) Q_closed + a Q_open,
   *   a = smoothstep((stage-stage0)/buffer).
   * gate_control = 1 (smooth=False, roseires_rating_curve.py:65-81,111-142): Q = Q_open or Q_closed by a per-member
   *   gate state: at every residual evaluation the cool-down runs down by the time since the last one, then the
   *   gates open when the stage seen at the previous evaluation is >= stage0 + 0.5 (close when <= stage0 - 1) and
   *   the cool-down restarts at max_cooldown; the state lives for the whole run.
   * Q_state = sum_{opening_j>0} spill(stage, opening_j) + n_sluices*sluice(stage, twl) + q_hydro
   * spill/sluice(s, o) = c[0] + c[1] s + c[2] o + c[3] s^2 + c[4] s o + c[5] o^2
   *   (sklearn PolynomialFeatures(2, include_bias=False) + LinearRegression: intercept_, coef_) */
  double spill[6], sluice[6];
  double twl;
  double open_state[PR_MAX_GATES], closed_state[PR_MAX_GATES];
  int32_t n_gates, sluices_open, sluices_closed, gate_control;
  double stage0, buffer, q_hydro, dY;           /* dY = 1e-3: central-difference step of dQ_dz */
  double max_cooldown;                           /* gate_control = 1: seconds (reference default 18000) */
  int32_t initially_open, reserved;
} pr_rating;

/* One boundary (boundary.py:7-247).  `series` = Hydrograph.get_at(k*dt) sampled for k = 0..levels-1
 * (hydrograph.py:21; the reference only ever evaluates hydrographs at t = k*dt, preissmann.py:215,313). */
typedef struct pr_bc {
  int32_t type;         /* enum pr_bc_type */
  int32_t reserved;
  double bed_level;     /* Boundary.bed_level (derivatives and rating stage use it, boundary.py:95,162,211) */
  double bed_slope;     /* cross_section.bed_slope of the boundary node (normal_depth) */
  double fixed_depth;   /* Boundary.initial_depth (fixed_depth target) */
  const double* series; /* [levels] or [M][levels]; flow or stage hydrograph samples */
  int64_t series_member_stride; /* 0 = shared by all members, else element stride between members */
  pr_rating rating;
  /* Release scenarios: one rating curve per ensemble member (different gate states, jammed gates, initial pool
   * levels ... for the same geometry).  [n_members] array in HOST memory whatever cfg->mem is (the library
   * reduces each curve to its device form before the launch), or NULL = `rating` for every member. */
  const pr_rating* member_ratings;
  /* lumped storage behind the boundary (lumped_storage.py:7-179).  Basic form: constant surface area, no outflow
   * rating curve, capture_losses = False) */
  double storage_area, storage_min_stage, storage_ymin, storage_ymax;
  /* General form (all optional): tabulated area curve (LumpedStorage.set_area_curve, :145-179), outflow rating
   * curve of the reservoir (mass_balance, :24-35) and head losses between the last node and the reservoir
   * (capture_losses, :47-143: friction over reservoir_length + empirical K_q V^2/2g; the expansion term needs
   * A_str, which the reference never passes).  The mass balance is then solved by Brent's method on
   * [storage_ymin, storage_ymax] exactly as the reference does (scipy.optimize.brentq). */
  const double* storage_curve_stage;  /* [storage_curve_len], increasing; NULL = constant area */
  const double* storage_curve_area;   /* [storage_curve_len] */
  int32_t storage_curve_len;
  int32_t storage_capture_losses;
  double storage_alpha, storage_beta; /* area_at(Y) = alpha * interp(Y + beta, stage, area) */
  double storage_reservoir_length, storage_Kq;
  pr_rating storage_outflow;          /* PR_RC_NONE = no outflow */
} pr_bc;

/* Initial conditions = Channel.initial_conditions (channel.py:123-138), split into two arrays. */
typedef struct pr_state {
  const double* depth;    /* [N] or [M][N] */
  const double* flow;     /* [N] or [M][N] */
  int64_t member_stride;  /* 0 = shared, else N */
} pr_state;

/* Result buffers (caller-allocated).  NULL pointers are skipped. */
typedef struct pr_outputs {
  double* depth;          /* per out_mode; level k holds the iterate BEFORE the last update (preissmann.py:166-177) */
  double* flow;
  int32_t* iters;         /* [M][levels-1] Newton iterations per level (preissmann.py:122-161) */
  int32_t* status;        /* [M] PR_STATUS_* */
  int32_t* fail_level;    /* [M] level at which a member stopped (0 if none) */
  double* storage_stage;  /* [M][levels] reservoir stage record (boundary.py:126-131), storage BC only;
                             entry 0 is the initial stage inserted by prepare_results (solver.py:101-108) */
  double* final_error;    /* [M][levels-1] ||R||_2 of the accepted iterate (diagnostic) */
} pr_outputs;

/* ------------------------------------------------------------------------------------------ */

/* Library identification; returns PR_ABI_VERSION. */
int pr_abi_version(void);

/* Text of the last error on the calling thread. */
const char* pr_last_error(void);

/* Replaces PreissmannSolver.run() (preissmann.py:101-163) for n_members members at once:
 * update_guesses -> compute_residual_vector -> compute_jacobian -> spsolve -> unknowns += delta ->
 * euclidean_norm(R) < tolerance, for every time level.  Synchronous w.r.t. host buffers when
 * cfg->mem == PR_MEM_HOST; with PR_MEM_DEVICE the work is only enqueued on `cuda_stream`
 * (a cudaStream_t, NULL = default stream) - also on the long-reach path, whose data-dependent number of Newton trips
 * runs as a CUDA-graph WHILE loop on the device. */
int pr_ensemble_run(const pr_config* cfg, const pr_geom* geom, const pr_bc* upstream,
                    const pr_bc* downstream, const pr_state* initial, const pr_outputs* out,
                    void* cuda_stream);

/* Replaces Channel._gvh_conditions (channel.py:307-378) per member: backwater predictor-corrector
 * from the downstream depth.  q0 / downstream_depth: [1] or [M] (member stride 0/1; the downstream pool level is a
 * per-member quantity in release-scenario ensembles).  Writes
 * ic_depth/ic_flow [M][N] and status [M] (PR_STATUS_SUPERCRITICAL mirrors the RuntimeError). */
int pr_gvf_initial_conditions(const pr_config* cfg, const pr_geom* geom, const double* q0,
                              int64_t q0_member_stride, const double* downstream_depth,
                              int64_t downstream_depth_member_stride, double* ic_depth,
                              double* ic_flow, int32_t* status, void* cuda_stream);

/* Replaces Channel._steady_conditions (channel.py:296-305): per member and node the normal depth for q0, i.e.
 * the root of Q - K(hw) sqrt(S0) on [z_min, z_min+100] by Brent's method (CrossSection.normal_depth,
 * cross_section.py:184-202 -> scipy.optimize.brentq).  bed_slope: [N] cross-section bed slopes.
 * Writes ic_depth / ic_flow [M][N]. */
int pr_normal_depth_initial_conditions(const pr_config* cfg, const pr_geom* geom, const double* bed_slope,
                                       const double* q0, int64_t q0_member_stride, double* ic_depth,
                                       double* ic_flow, void* cuda_stream);

/* Replaces the array part of Solver.prepare_results (solver.py:65-98) for [M][levels][N] results:
 * level, area, top_width, froude_number, velocity, wave_celerity (any output pointer may be NULL). */
int pr_derived_results(const pr_config* cfg, const pr_geom* geom, const double* depth, const double* flow,
                       double* level, double* area, double* top_width, double* froude, double* velocity,
                       double* celerity, void* cuda_stream);

/* Replaces the calibration objective of cases/gerd_roseires (model.py:105-113 + n_calibrate.py:55-63):
 * levels[m][j] = np.interp(Q[j], flow[m][:,0], depth[m][:,0] + z0);  rmse[m] = mean((levels-H)^2)^0.5.
 * up_flow/up_depth are [M][levels] (PR_OUT_UPSTREAM layout). */
int pr_rating_objective(const pr_config* cfg, const double* up_flow, const double* up_depth, double z0,
                        const double* q_query, const double* h_target, int32_t n_query,
                        double* levels_out, double* rmse_out, void* cuda_stream);

/* Measured FP64 FMA peak of the current device (TFLOP/s), the roofline denominator SURVEY.md 8(d)
 * asks for.  Runs a register-resident DFMA kernel for about `millis` ms. */
int pr_fp64_peak(double millis, double* tflops_out);

/* The long-reach path (n_nodes > 249) keeps device workspaces (iterate, level constants, tile cells: ~7 GB for
 * 100 001 nodes x 1 024 members) in a pool, up to two per device, so that two streams can run long reaches side by
 * side; a third concurrent run waits on the GPU for the first workspace that frees up.  This call waits for the idle
 * ones and frees them.  Every entry point may be called from several host threads (pr_last_error is per thread). */
int pr_release_workspace(void);

/* Newton trips (one trip = one iteration of every unfinished member) of the most recent long-reach run; waits for
 * that run.  Each trip is 5 kernel launches issued from a CUDA-graph loop on the device, which pr_launch_count cannot
 * see.  -1 when no long-reach run has been made. */
int64_t pr_long_last_trips(void);

/* Diagnostics: evaluates the device's branch-free FP64 primitives (reciprocal, square root, reciprocal square
 * root, reciprocal cube root - csrc/pr_device.cuh) and their raw SFU seeds on n HOST values;
 * out_host is [6][n].  Used by tests/test_gpu_math.py to bound their error against IEEE results. */
int pr_math_probe(const double* x_host, int32_t n, double* out_host);

/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
int64_t pr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PREISSMANN_B200_H */
