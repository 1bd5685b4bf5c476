// pr_ensemble_kernel.cuh - the fused Preissmann ensemble kernel (short / medium reaches, N <= 249).
//
// One ensemble member is advanced through ALL time levels by a group of G lanes of one warp
// (G = 32: one member per warp).  Lane l owns M consecutive nodes (slots) and the M cells to their
// right; the state lives in registers, per-warp scratch in shared memory:
//
//   per Newton iteration (preissmann.py:122-156)
//     node pass      each lane evaluates its M nodes (area, top width, conveyance, slopes, derivatives)
//     cell pass      continuity + momentum residuals and their 8 Jacobian entries per cell
//                    (preissmann.py:220-301, 407-733); ||R||^2 accumulated on the fly
//     local solve    the lane condenses its M cells into ONE cell between its first node and the next
//                    lane's first node by 2x2 Schur complements (block elimination of interior nodes)
//     chain solve    the condensed cells + both boundary rows form a 2x2-block tridiagonal system of
//                    <= G block rows, one per lane: parallel cyclic reduction over warp shuffles
//                    (log2 G steps; off-diagonal blocks are rank-1 and stay rank-1)
//     back-subst.    interior nodes recovered locally; x += delta; convergence vote = warp-uniform flag
//
// This replaces scipy's SuperLU call (preissmann.py:146): pairing the rows as (U,C0),(M0,C1),...,(M_{N-2},D)
// gives non-singular diagonal blocks (hard part 5).
//
// Level-k quantities enter the residuals only through four per-cell constants (cC, cM, cA, cS); they
// are refreshed at convergence from partial sums the cell pass has anyway, and live in a shared-memory
// ping-pong so that an iteration that turns out to be the converged one costs no re-evaluation.
#pragma once
#include "pr_device.cuh"
#include "pr_irregular.cuh"

// tuning switches (A/B builds: -DPR_...=0/1)
#ifndef PR_PCR_V2
#define PR_PCR_V2 1     // parallel cyclic reduction with two short exchanges per step (14 values) instead of one of 20
#endif

namespace pr {

constexpr unsigned kFull = 0xffffffffu;

// A (possibly condensed) cell between nodes a (left) and b (right), right-hand sides already negated:
//   C: c1 dh_a + c2 dQ_a + c3 dh_b + c4 dQ_b = rc        M: m1 dh_a + m2 dQ_a + m3 dh_b + m4 dQ_b = rm
struct Cell {
  double c1, c2, c3, c4, rc;
  double m1, m2, m3, m4, rm;
};

// Back-substitution record of an eliminated node b between a and c:
//   x_b = Dinv * ( [rm - m1 dh_a - m2 dQ_a ; rc - c3 dh_c - c4 dQ_c] )
struct Elim {
  double i11, i12, i21, i22;
  double m1, m2, rm;
  double c3, rc;          // c4 of a raw cell is the constant theta/dx
};

// Schur-complement merge of S (a..b) and E (b..c) eliminating node b; pivot rows are S.M and E.C.
__device__ __forceinline__ void merge_cells(Cell& S, const Cell& E, Elim& el) {
  const double det = S.m3 * E.c2 - S.m4 * E.c1;
  const double idet = fast_rcp(det);
  const double i11 = E.c2 * idet, i12 = -S.m4 * idet, i21 = -E.c1 * idet, i22 = S.m3 * idet;
  el.i11 = i11; el.i12 = i12; el.i21 = i21; el.i22 = i22;
  el.m1 = S.m1; el.m2 = S.m2; el.rm = S.rm;
  el.c3 = E.c3; el.rc = E.rc;
  // w = [S.c3 S.c4] * Dinv ; v = [E.m1 E.m2] * Dinv
  const double w1 = S.c3 * i11 + S.c4 * i21, w2 = S.c3 * i12 + S.c4 * i22;
  const double v1 = E.m1 * i11 + E.m2 * i21, v2 = E.m1 * i12 + E.m2 * i22;
  Cell n;
  n.c1 = S.c1 - w1 * S.m1;  n.c2 = S.c2 - w1 * S.m2;
  n.c3 = -w2 * E.c3;        n.c4 = -w2 * E.c4;
  n.rc = S.rc - w1 * S.rm - w2 * E.rc;
  n.m1 = -v1 * S.m1;        n.m2 = -v1 * S.m2;
  n.m3 = E.m3 - v2 * E.c3;  n.m4 = E.m4 - v2 * E.c4;
  n.rm = E.rm - v1 * S.rm - v2 * E.rc;
  S = n;
}

// Level-k parts of the residuals of one cell, from the node values of the stored level:
//   cC = -(A_i+A_i+1)/(2dt) + (1-theta)(Q_i+1-Q_i)/dx          cM = -(Q_i+Q_i+1)/(2dt) + (1-theta)(F_i+1-F_i)/dx
//   cA = (1-theta)/2 (A_i+A_i+1)                               cS = (1-theta)(Y_i+1-Y_i)/dx + (1-theta)/2 (Se_i+Se_i+1)
__device__ __forceinline__ void level_constants(const NodeVals& a, const NodeVals& b, const DevParams& k, double& nC,
                                                double& nM, double& nA, double& nS) {
  const double sA = b.A + a.A, dQ = b.Q - a.Q, sQ = b.Q + a.Q, dF = b.F - a.F, dY = b.Y - a.Y, sSe = b.Se + a.Se;
  nC = fma(-sA, k.i2dt, k.omt_dx * dQ);
  nM = fma(-sQ, k.i2dt, k.omt_dx * dF);
  nA = k.homt * sA;
  nS = fma(k.omt_dx, dY, k.homt * sSe);
}

// Residuals + Jacobian of one Preissmann cell (left node a, right node b).  (cC, cM, cA, cS) are the stored
// level's contributions; returns R_C^2 + R_M^2.
__device__ __forceinline__ double cell_assemble(const NodeVals& a, const NodeVals& b, const DevParams& k,
                                                const double cC, const double cM, const double cA, const double cS,
                                                Cell& e) {
  const double sA = b.A + a.A, dQ = b.Q - a.Q, sQ = b.Q + a.Q, dF = b.F - a.F, dY = b.Y - a.Y, sSe = b.Se + a.Se;
  // continuity_residual (preissmann.py:220-249): R_C = time_diff(A) + spatial_diff(Q);  e.rc = -R_C
  e.rc = fma(-sA, k.i2dt, fma(-k.th_dx, dQ, -cC));
  // momentum_residual (preissmann.py:251-301): R_M = time_diff(Q) + spatial_diff(Q^2/A) + g avgA (dYdx + avgSe)
  const double avgA = fma(k.hth, sA, cA);                              // cell_avg(A)
  const double slope = fma(k.th_dx, dY, fma(k.hth, sSe, cS));          // spatial_diff(z+h) + cell_avg(Se)
  const double ga = k.g * avgA;
  e.rm = fma(-ga, slope, fma(-sQ, k.i2dt, fma(-k.th_dx, dF, -cM)));
  // dC_* (preissmann.py:407-494)
  e.c1 = a.T * k.i2dt;  e.c2 = -k.th_dx;  e.c3 = b.T * k.i2dt;  e.c4 = k.th_dx;
  // dM_dh_i / dM_dQ_i / dM_dh_ip1 / dM_dQ_ip1 (preissmann.py:496-733); spatial_diff(unit) = -/+ theta/dx
  const double gs = k.ghth * slope;
  e.m1 = fma(gs, a.T, fma(ga, a.w2 - k.th_dx, a.w1));
  e.m2 = fma(ga, a.w3, k.i2dt - a.w4);
  e.m3 = fma(gs, b.T, fma(ga, b.w2 + k.th_dx, -b.w1));
  e.m4 = fma(ga, b.w3, k.i2dt + b.w4);
  return fma(e.rc, e.rc, e.rm * e.rm);
}

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int s = G / 2; s > 0; s >>= 1) v += __shfl_xor_sync(kFull, v, s, G);
  return v;
}

// Shared memory (doubles): [geometry F_COUNT x NP] then per warp
//   level constants  4 x M x 32       (of the stored level; rebuilt by one extra node pass when a level is accepted)
//   elimination recs (M-1) x 9 x 32
//   neighbour exchange 8 x 32         (first-node values handed to the lane on the left)
// every per-warp array is indexed [..][lane]: consecutive lanes, consecutive doubles, no bank conflicts.
constexpr int kXch = 8;
template <int M>
__host__ __device__ constexpr int warp_smem_doubles() { return 32 * (4 * M + 9 * (M - 1) + kXch); }

template <int G, int M, int W>
__host__ __device__ constexpr size_t ensemble_smem_bytes() {
  return sizeof(double) * ((size_t)F_COUNT * G * M + (size_t)W * warp_smem_doubles<M>());
}

template <int G, int M, int W, bool CURV, int RM, bool EXACT, bool GST, bool IRR = false>
__global__ void __launch_bounds__(W * 32, 1)
pr_ensemble_kernel(const __grid_constant__ DevParams p) {
  extern __shared__ double smem[];
  constexpr int NP = G * M;           // padded node slots per member
  constexpr int MPW = 32 / G;         // members per warp
  double* sg = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* pcw = smem + (size_t)F_COUNT * NP + (size_t)warp * warp_smem_doubles<M>();   // level constants
  double* elw = pcw + 4 * M * 32;                                                    // elimination records
  double* xw = elw + (M - 1) * 9 * 32;                                                   // neighbour exchange
  // slot (j, gl) at index j*G + gl holds node gl*M + j
  stage_geometry(p.geo, p.N, NP, sg, threadIdx.x, blockDim.x, [](int idx) { return (idx % G) * M + idx / G; });
  __syncthreads();

  const int gl = lane % G;            // lane within the member's group
  const int N = p.N, L = p.L;

  // chain topology (uniform): lanes 0..Lc hold one block row each
  const int ncells_total = N - 1;
  const int Lc = (ncells_total + M - 1) / M;           // lanes that own at least one cell
  const int my_first = gl * M;                         // first node owned by this lane
  int nc = ncells_total - my_first;                    // cells owned by this lane
  nc = nc < 0 ? 0 : (nc > M ? M : nc);
  // where node N-1 lives (EXACT: (N-1) % M == 0, so it is slot 0 of lane Lc - known at compile time)
  const int slot_last = EXACT ? 0 : (N - 1) % M;
  const int owner_last = EXACT ? Lc : (N - 1) / M;
  const bool is_first = (gl == 0);
  const bool owns_last = (gl == owner_last);

  // Persistent warps: the grid holds one CTA per SM, and every warp draws its next member (group of 32/G members)
  // from a ticket counter until the ensemble is used up.  Members differ in their Newton iteration totals (409 ... 676
  // across the gerd roughness grid), so static member -> CTA assignment leaves SMs idle behind the slowest warp of a
  // CTA and in the last partial wave; p.member_order (optional) lets the caller hand out expensive members first.
  for (;;) {
  // PHASE: prologue
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(p.ticket, 1u);
  ticket = __shfl_sync(kFull, ticket, 0);
  if (ticket >= (unsigned)p.n_tickets) break;
  int member = (int)ticket * MPW + lane / G;
  const bool member_valid = member < p.M;
  if (!member_valid) member = p.M - 1;
  if (p.member_order) member = p.member_order[member];

  const DevParams& k = p;
  const Rough rg = load_rough<RM>(p.geo, member);
  // EXACT: (N-1) % M == 0, i.e. every lane owns either M cells or none.  Lanes without cells then run the
  // cell pass on padding (finite copies of the last node, zeroed scratch) instead of branching around it,
  // which leaves the node/cell/merge loop as straight-line code.
  static_assert(G == 8 || G == 16 || G == 32, "a member occupies 8, 16 or 32 lanes");
  // G < 32: 32/G members share a warp and run in lockstep; each keeps its own level / iteration / status, and
  // everything that is decided per member (convergence, failure, the level-constant refresh) is per lane group.

  // ---- state: initial conditions (Solver.initialize_t0, solver.py:61-63) ----
  double h[M], q[M];
  {
    const double* ih = p.ic_h + (long long)member * p.ic_stride;
    const double* iq = p.ic_q + (long long)member * p.ic_stride;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const int nd = my_first + j < N ? my_first + j : N - 1;
      h[j] = ih[nd];
      q[j] = iq[nd];
    }
  }
  // the downstream boundary node's unknowns
  auto last_node = [&](double& hl, double& ql) {
    hl = h[0]; ql = q[0];
#pragma unroll
    for (int j = 1; j < M; ++j)
      if (slot_last == j) { hl = h[j]; ql = q[j]; }
  };
  const size_t out_row = (p.out_mode == PR_OUT_FULL) ? (size_t)N : 1;
  auto store_level = [&](int level, bool nanfill) {
    if (!member_valid) return;
    const size_t base = ((size_t)member * L + level) * out_row;
    if (p.out_mode == PR_OUT_FULL) {
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const int nd = my_first + j;
        if (nd < N) {
          if (p.out_h) p.out_h[base + nd] = nanfill ? nan("") : h[j];
          if (p.out_q) p.out_q[base + nd] = nanfill ? nan("") : q[j];
        }
      }
    } else if (is_first) {
      if (p.out_h) p.out_h[base] = nanfill ? nan("") : h[0];
      if (p.out_q) p.out_q[base] = nanfill ? nan("") : q[0];
    }
  };
  store_level(0, false);

  // boundary bookkeeping held by the lane that owns node N-1
  double q_prev_last, stage_prev;
  {
    double hl, ql;
    last_node(hl, ql);
    q_prev_last = ql;                                                   // flow_at(k=-1, i=-1)
    stage_prev = sg[F_Z * NP + slot_last * G + owner_last] + hl;        // solver.py:101-108
    if (GST && p.dn.type == PR_BC_FIXED_DEPTH_STORAGE && p.dn.st_losses && owns_last) {
      NodeVals t;
      NodeConv kc;
      if (IRR && sg[F_KIND * NP + slot_last * G + gl] == (double)PR_XS_IRREGULAR) node_eval_irregular_call<CURV>(p.geo, N - 1, hl, ql, rg, k, t, &kc);
      else node_eval<CURV, RM, true>(sg, NP, slot_last * G + gl, hl, ql, rg, k, t, &kc);
      const double V = ql / kc.A;
      stage_prev -= kc.Sf * p.dn.st_length + p.dn.st_kq * (V * V) / (2.0 * p.g);     // initial stage = Y - energy_loss
    }
  }
  if (member_valid && owns_last && p.storage_stage) p.storage_stage[(size_t)member * L] = stage_prev;
  GateState gate;
  if (GST) gate_init(gate, p.dn.member_rc ? p.dn.member_rc[member] : p.dn.rc);
  // conveyance / friction slope of a boundary node are only needed by the normal-depth condition and by the head
  // losses of a lumped storage
  const bool up_normal = p.up.type == PR_BC_NORMAL_DEPTH;
  const bool dn_normal = p.dn.type == PR_BC_NORMAL_DEPTH || (GST && p.dn.type == PR_BC_FIXED_DEPTH_STORAGE && p.dn.st_losses);

  const bool roseires3 = !GST && G == 32 && p.dn.type == PR_BC_RATING_CURVE && !p.dn.member_rc && p.dn.rc.type == PR_RC_ROSEIRES;

  // Short reaches: a normal-depth (or head-loss) boundary needs the conveyance of node N-1.  Evaluating that node a
  // second time on its owner lane costs a whole node pass at one active thread - a third of the iteration with one
  // node per lane - so these builds take K and dK/dA from the regular pass instead (WANT_K there, a few flops).
  constexpr bool DNK = (G < 32 || M <= 2) && !IRR;
  NodeConv kc_last = {0.0, 0.0, 1.0, 0.0, 0.0, 0.0};
  double T_last = 0.0;

  int level = 1, it = 0;
  bool active = member_valid && L > 1;
  bool failed = false;
  double hyd_up = 0.0, hyd_dn = 0.0;

#define PC(c, j) pcw[((c)*M + (j)) * 32 + lane]
#define EL(j, c) elw[((j)*9 + (c)) * 32 + lane]
#define XW(c, l) xw[(c)*32 + (l)]
// Node pass of slot (j, gl) = node gl*M + j.  IRR kernels send IrregularSection nodes to the polyline scans.
#define PR_NODE(WANTK, j, hh, qq_, out, kcp)                                                                       \
  do {                                                                                                            \
    if (IRR && sg[F_KIND * NP + (j)*G + gl] == (double)PR_XS_IRREGULAR)                                            \
      node_eval_irregular_call<CURV>(p.geo, (gl * M + (j) < N ? gl * M + (j) : N - 1), (hh), (qq_), rg, k, (out), (kcp)); \
    else                                                                                                          \
      node_eval<CURV, RM, WANTK>(sg, NP, (j)*G + gl, (hh), (qq_), rg, k, (out), (kcp));                            \
  } while (0)

  // Node pass at the current state that only (re)builds the level constants: once for the initial state and once
  // per accepted level.  Cheaper than producing candidates in every Newton iteration (one extra node pass per
  // level against 4 stores + 9 flops per cell per iteration), and it halves the constants' shared memory.
  bool commit = true;      // G < 32: which lanes take the refreshed constants
  // PHASE: level refresh
  auto refresh_level_constants = [&]() {
    NodeVals left, right;
    PR_NODE(false, 0, h[0], q[0], left, nullptr);
    __syncwarp();
    XW(0, lane) = left.Q;  XW(1, lane) = left.A;  XW(3, lane) = left.Y;  XW(4, lane) = left.Se;  XW(5, lane) = left.QA;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (j + 1 < M) {
        PR_NODE(false, j + 1, h[j + 1], q[j + 1], right, nullptr);
      } else {
        const int nl = (lane + 1) & 31;
        right.Q = XW(0, nl);  right.A = XW(1, nl);  right.Y = XW(3, nl);  right.Se = XW(4, nl);
        right.F = right.Q * XW(5, nl);
      }
      double nC, nM, nA, nS;
      level_constants(left, right, k, nC, nM, nA, nS);
      if (G == 32 || commit) { PC(0, j) = nC; PC(1, j) = nM; PC(2, j) = nA; PC(3, j) = nS; }
      left = right;
    }
    __syncwarp();
  };
  // PHASE: prologue
  refresh_level_constants();

  // PHASE: loop control
  while (__any_sync(kFull, active)) {
    // hydrograph samples at t = level*dt (preissmann.py:215,313)
#define PR_HYD(bc) ((bc).series ? (bc).series[(long long)member * (bc).series_stride + (level < L ? level : L - 1)] : 0.0)
    if (it == 0) { hyd_up = PR_HYD(p.up); hyd_dn = PR_HYD(p.dn); }
    if (active) it += 1;

    // ------------------------------ node + cell pass ------------------------------
    // PHASE: node pass
    NodeVals left, right;
    if (DNK && dn_normal) {
      NodeConv kct;
      PR_NODE(true, 0, h[0], q[0], left, &kct);
      if (slot_last == 0) { kc_last = kct; T_last = left.T; }
    } else {
      PR_NODE(false, 0, h[0], q[0], left, nullptr);
    }
    // hand the first node to the lane on the left: it closes that lane's last cell
    __syncwarp();
    XW(0, lane) = left.Q;  XW(1, lane) = left.A;  XW(2, lane) = left.T;  XW(3, lane) = left.Y;  XW(4, lane) = left.Se;
    XW(5, lane) = left.QA; XW(6, lane) = left.w2; XW(7, lane) = left.w3;
    __syncwarp();

    double ss = 0.0;
    Cell S;                     // condensed cell of this lane
#pragma unroll
    for (int j = 0; j < M; ++j) {
      // PHASE: node pass
      if (j + 1 < M) {
        if (DNK && dn_normal) {
          NodeConv kct;
          PR_NODE(true, j + 1, h[j + 1], q[j + 1], right, &kct);
          if (slot_last == j + 1) { kc_last = kct; T_last = right.T; }
        } else {
          PR_NODE(false, j + 1, h[j + 1], q[j + 1], right, nullptr);
        }
      } else {
        const int nl = (lane + 1) & 31;
        right.Q = XW(0, nl);  right.A = XW(1, nl);  right.T = XW(2, nl);  right.Y = XW(3, nl);  right.Se = XW(4, nl);
        right.QA = XW(5, nl); right.w2 = XW(6, nl); right.w3 = XW(7, nl);
        right.F = right.Q * right.QA;
        right.w1 = (k.th_dx * right.QA) * (right.QA * right.T);
        right.w4 = k.th_dx2 * right.QA;
      }
      // PHASE: cell pass
      if (EXACT || j < nc) {
        Cell e;
        const double r2 = cell_assemble(left, right, k, PC(0, j), PC(1, j), PC(2, j), PC(3, j), e);
        ss += (!EXACT || nc > 0) ? r2 : 0.0;
        // PHASE: merge
        if (j == 0) S = e;
        else {
          Elim el;
          merge_cells(S, e, el);
          EL(j - 1, 0) = el.i11; EL(j - 1, 1) = el.i12; EL(j - 1, 2) = el.i21; EL(j - 1, 3) = el.i22;
          EL(j - 1, 4) = el.m1;  EL(j - 1, 5) = el.m2;  EL(j - 1, 6) = el.rm;
          EL(j - 1, 7) = el.c3;  EL(j - 1, 8) = el.rc;
        }
      }
      left = right;
    }

    // ------------------------------ boundary rows ------------------------------
    // PHASE: boundary rows
    BcRow U, D;
    U.res = 0.0; U.dh = 1.0; U.dq = 0.0; U.stage_rec = 0.0; U.fail = false;
    D = U;
    if (p.up.type == PR_BC_FLOW_HYDROGRAPH) {
      // the usual upstream condition, Q - hyd(t) (boundary.py:81-84): every lane forms it (no divergent one-lane stretch),
      // lane 0's copy is the row
      U.res = q[0] - hyd_up; U.dh = 0.0; U.dq = 1.0;
    } else if (is_first) {
      NodeConv kc = {0.0, 0.0, 1.0, 0.0, 0.0, 0.0};
      double T0 = 0.0;
      if (up_normal) {
        NodeVals t;
        PR_NODE(true, 0, h[0], q[0], t, &kc);
        T0 = t.T;
      }
      U = bc_eval<false>(p.up, member, level, hyd_up, h[0], q[0], 0.0, 0.0, p.dt, p.g, kc, T0);
    }
    if (roseires3) {
      // Shared Roseires curve: dQ/dz is a central difference (quirk 10), i.e. three curve evaluations per iteration on
      // a single lane.  Run them on three lanes at once instead: owner (stage), owner+1 (stage + dY), owner+2 (- dY).
      double hl, ql;
      last_node(hl, ql);
      const double stage = p.dn.bed_level + __shfl_sync(kFull, hl, owner_last, G);
      const int role = (gl - owner_last) & (G - 1);
      const double qv = roseires_q(p.dn.rc, stage + (role == 1 ? p.dn.rc.dY : role == 2 ? -p.dn.rc.dY : 0.0));
      const double qp = __shfl_sync(kFull, qv, (owner_last + 1) & (G - 1), G);
      const double qm = __shfl_sync(kFull, qv, (owner_last + 2) & (G - 1), G);
      if (owns_last) { D.res = ql - qv; D.dh = 0.0 - (qp - qm) * p.dn.rc.inv_2dY; D.dq = 1.0; }
    } else if (owns_last) {
      double hl, ql;
      last_node(hl, ql);
      NodeConv kc = {0.0, 0.0, 1.0, 0.0, 0.0, 0.0};
      double T0 = 0.0;
      if (dn_normal) {
        if (DNK) {
          kc = kc_last; T0 = T_last;
        } else {
          NodeVals t;
          PR_NODE(true, slot_last, hl, ql, t, &kc);
          T0 = t.T;
        }
      }
      D = bc_eval<GST>(p.dn, member, level, hyd_dn, hl, ql, q_prev_last, stage_prev, p.dt, p.g, kc, T0, GST ? &gate : nullptr);
    }
    if (is_first) ss = fma(U.res, U.res, ss);
    if (owns_last) ss = fma(D.res, D.res, ss);
    // utility.euclidean_norm (utility.py:20-22) < tol, decided on the squared norm; the root is only taken when the
    // norm itself is asked for (final_error)
    const double err2 = group_sum<G>(ss);

    // ------------------------------ chain rows ------------------------------
    // PHASE: chain rows
    // row a (top): lane 0 -> U ; lane 1..Lc -> M~ of the previous lane's condensed cell
    // row b (bottom): lane < Lc -> own C~ ; lane Lc -> D ; beyond -> identity
    double l1, l2, d11, d12, d21, d22, u1, u2, ra, rb;
    {
      const double pm1 = __shfl_up_sync(kFull, S.m1, 1, G), pm2 = __shfl_up_sync(kFull, S.m2, 1, G);
      const double pm3 = __shfl_up_sync(kFull, S.m3, 1, G), pm4 = __shfl_up_sync(kFull, S.m4, 1, G);
      const double prm = __shfl_up_sync(kFull, S.rm, 1, G);
      // D row travels one lane up when node N-1 is interior to lane Lc-1
      const double Dh = (owner_last == Lc) ? D.dh : __shfl_up_sync(kFull, D.dh, 1, G);
      const double Dq = (owner_last == Lc) ? D.dq : __shfl_up_sync(kFull, D.dq, 1, G);
      const double Dr = (owner_last == Lc) ? D.res : __shfl_up_sync(kFull, D.res, 1, G);
      if (gl == 0) { l1 = 0.0; l2 = 0.0; d11 = U.dh; d12 = U.dq; ra = -U.res; }
      else if (gl <= Lc) { l1 = pm1; l2 = pm2; d11 = pm3; d12 = pm4; ra = prm; }
      else { l1 = 0.0; l2 = 0.0; d11 = 1.0; d12 = 0.0; ra = 0.0; }
      if (gl < Lc) { d21 = S.c1; d22 = S.c2; u1 = S.c3; u2 = S.c4; rb = S.rc; }
      else if (gl == Lc) { d21 = Dh; d22 = Dq; u1 = 0.0; u2 = 0.0; rb = -Dr; }
      else { d21 = 0.0; d22 = 1.0; u1 = 0.0; u2 = 0.0; rb = 0.0; }
    }
    // ------------------------------ parallel cyclic reduction ------------------------------
    // PHASE: PCR
    // Rows whose partner lies outside the chain have a zero coupling (l = 0 / u = 0) by construction, and
    // every lane (rows beyond the chain are identity rows) holds finite data, so no guards are needed.
#pragma unroll
    for (int s = 1; s < G; s <<= 1) {
      if (s > Lc) continue;   // uniform: the chain has Lc+1 rows (no break: the steps stay unrolled, no loop-carried moves)
#if PR_PCR_V2
      // Elimination of the couplings to lanes gl-s (through l) and gl+s (through u).  The work for a neighbour's row
      // is done HERE, on the lane that owns the pivot block: the neighbours send their coupling vectors (2 + 2 values,
      // independent of the reciprocal below, so the exchange overlaps it) and get back the five numbers their row
      // needs - 14 values per lane and step instead of the 20 of shipping whole block rows.
      const int up_src = (gl - s) & (G - 1), dn_src = (gl + s) & (G - 1);
      const double Dl1 = __shfl_sync(kFull, l1, dn_src, G), Dl2 = __shfl_sync(kFull, l2, dn_src, G);   // l of lane gl+s
      const double Uu1 = __shfl_sync(kFull, u1, up_src, G), Uu2 = __shfl_sync(kFull, u2, up_src, G);   // u of lane gl-s
      const double idet = fast_rcp(d11 * d22 - d12 * d21);
      const double i11 = d22 * idet, i12 = -d12 * idet, i21 = -d21 * idet, i22 = d11 * idet;
      // for lane gl+s (this lane is its P): [l1 l2] * Dinv, then its new l, the change of its (d11, d12) and of its ra
      const double a1 = Dl1 * i11 + Dl2 * i21, a2 = Dl1 * i12 + Dl2 * i22;
      const double o_l1 = -a1 * l1, o_l2 = -a1 * l2, o_d1 = a2 * u1, o_d2 = a2 * u2, o_r = a1 * ra + a2 * rb;
      // for lane gl-s (this lane is its N): [u1 u2] * Dinv, then its new u, the change of its (d21, d22) and of its rb
      const double b1 = Uu1 * i11 + Uu2 * i21, b2 = Uu1 * i12 + Uu2 * i22;
      const double o_u1 = -b2 * u1, o_u2 = -b2 * u2, o_e1 = b1 * l1, o_e2 = b1 * l2, o_s = b1 * ra + b2 * rb;
      l1 = __shfl_sync(kFull, o_l1, up_src, G);  l2 = __shfl_sync(kFull, o_l2, up_src, G);
      d11 -= __shfl_sync(kFull, o_d1, up_src, G); d12 -= __shfl_sync(kFull, o_d2, up_src, G);
      ra -= __shfl_sync(kFull, o_r, up_src, G);
      u1 = __shfl_sync(kFull, o_u1, dn_src, G);  u2 = __shfl_sync(kFull, o_u2, dn_src, G);
      d21 -= __shfl_sync(kFull, o_e1, dn_src, G); d22 -= __shfl_sync(kFull, o_e2, dn_src, G);
      rb -= __shfl_sync(kFull, o_s, dn_src, G);
#else
      const double idet = fast_rcp(d11 * d22 - d12 * d21);
      const double i11 = d22 * idet, i12 = -d12 * idet, i21 = -d21 * idet, i22 = d11 * idet;
      const int up_src = (gl - s) & (G - 1), dn_src = (gl + s) & (G - 1);
      // rows of lane gl-s
      const double P_i11 = __shfl_sync(kFull, i11, up_src, G), P_i12 = __shfl_sync(kFull, i12, up_src, G);
      const double P_i21 = __shfl_sync(kFull, i21, up_src, G), P_i22 = __shfl_sync(kFull, i22, up_src, G);
      const double P_l1 = __shfl_sync(kFull, l1, up_src, G), P_l2 = __shfl_sync(kFull, l2, up_src, G);
      const double P_u1 = __shfl_sync(kFull, u1, up_src, G), P_u2 = __shfl_sync(kFull, u2, up_src, G);
      const double P_ra = __shfl_sync(kFull, ra, up_src, G), P_rb = __shfl_sync(kFull, rb, up_src, G);
      // rows of lane gl+s
      const double N_i11 = __shfl_sync(kFull, i11, dn_src, G), N_i12 = __shfl_sync(kFull, i12, dn_src, G);
      const double N_i21 = __shfl_sync(kFull, i21, dn_src, G), N_i22 = __shfl_sync(kFull, i22, dn_src, G);
      const double N_l1 = __shfl_sync(kFull, l1, dn_src, G), N_l2 = __shfl_sync(kFull, l2, dn_src, G);
      const double N_u1 = __shfl_sync(kFull, u1, dn_src, G), N_u2 = __shfl_sync(kFull, u2, dn_src, G);
      const double N_ra = __shfl_sync(kFull, ra, dn_src, G), N_rb = __shfl_sync(kFull, rb, dn_src, G);
      const double a1 = l1 * P_i11 + l2 * P_i21, a2 = l1 * P_i12 + l2 * P_i22;   // [l1 l2] * Dinv(P)
      const double b1 = u1 * N_i11 + u2 * N_i21, b2 = u1 * N_i12 + u2 * N_i22;   // [u1 u2] * Dinv(N)
      // row a -= a1*rowa(P) + a2*rowb(P) ; row b -= b1*rowa(N) + b2*rowb(N)
      l1 = -a1 * P_l1;  l2 = -a1 * P_l2;
      d11 -= a2 * P_u1; d12 -= a2 * P_u2;
      ra -= a1 * P_ra + a2 * P_rb;
      u1 = -b2 * N_u1;  u2 = -b2 * N_u2;
      d21 -= b1 * N_l1; d22 -= b1 * N_l2;
      rb -= b1 * N_ra + b2 * N_rb;
#endif
    }
    double dh0, dq0;    // update of this lane's chain node
    {
      const double idet = fast_rcp(d11 * d22 - d12 * d21);
      dh0 = (d22 * ra - d12 * rb) * idet;
      dq0 = (d11 * rb - d21 * ra) * idet;
    }
    // ------------------------------ back-substitution ------------------------------
    // PHASE: back-substitution
    const double dhR = __shfl_down_sync(kFull, dh0, 1, G), dqR = __shfl_down_sync(kFull, dq0, 1, G);
    double dh[M], dq[M];
    dh[0] = dh0; dq[0] = dq0;
    {
      double rh = dhR, rq = dqR;       // right end of the span being unwound
#pragma unroll
      for (int j = M - 1; j >= 1; --j) {
        if (EXACT) {                   // every lane holds M cells or none: no branch, padding lanes keep their copies
          const double t1 = EL(j - 1, 6) - EL(j - 1, 4) * dh0 - EL(j - 1, 5) * dq0;
          const double t2 = EL(j - 1, 8) - EL(j - 1, 7) * rh - p.th_dx * rq;
          rh = EL(j - 1, 0) * t1 + EL(j - 1, 1) * t2;
          rq = EL(j - 1, 2) * t1 + EL(j - 1, 3) * t2;
          dh[j] = nc > 0 ? rh : 0.0;
          dq[j] = nc > 0 ? rq : 0.0;
        } else if (j < nc) {           // node slot j was eliminated by merge j-1
          const double t1 = EL(j - 1, 6) - EL(j - 1, 4) * dh0 - EL(j - 1, 5) * dq0;
          const double t2 = EL(j - 1, 8) - EL(j - 1, 7) * rh - p.th_dx * rq;
          dh[j] = EL(j - 1, 0) * t1 + EL(j - 1, 1) * t2;
          dq[j] = EL(j - 1, 2) * t1 + EL(j - 1, 3) * t2;
          rh = dh[j]; rq = dq[j];
        } else if (j == nc && nc > 0) {  // right end node of a short last lane: it is the chain's last node
          dh[j] = dhR; dq[j] = dqR;
        } else {
          dh[j] = 0.0; dq[j] = 0.0;
        }
      }
    }

    // ------------------------------ accept / update (preissmann.py:146-156) ------------------------------
    // PHASE: accept/update
    // a boundary evaluation the reference would abort on (brentq without a sign change) fails the member at once
    bool bc_failed;
    if (G == 32) bc_failed = __any_sync(kFull, (is_first && U.fail) || (owns_last && D.fail));
    else bc_failed = (__ballot_sync(kFull, (is_first && U.fail) || (owns_last && D.fail)) & ((((1u << (G & 31)) - 1u) << (lane - gl)))) != 0u;
    const bool converged = !bc_failed && err2 < p.tol2;
    if (active) {
      if (converged) {
        store_level(level, false);                     // stored level = iterate BEFORE the update
        if (is_first) {
          if (p.iters) p.iters[(size_t)member * (L - 1) + (level - 1)] = it;
          if (p.final_error) p.final_error[(size_t)member * (L - 1) + (level - 1)] = sqrt(err2);
        }
        if (owns_last) {
          double hl, ql;
          last_node(hl, ql);
          q_prev_last = ql;
          if (p.dn.type == PR_BC_FIXED_DEPTH_STORAGE) {
            stage_prev = D.stage_rec;
            if (p.storage_stage) p.storage_stage[(size_t)member * L + level] = D.stage_rec;
          }
        }
      }
    }
    // the accepted iterate becomes the stored level: rebuild its constants before the update is applied
    // (every lane runs the pass as soon as one member of the warp needs it; only that member's lanes commit)
    if (G < 32) commit = active && converged;
    if (__any_sync(kFull, active && converged)) refresh_level_constants();
    if (active) {
#pragma unroll
      for (int j = 0; j < M; ++j) { h[j] = poison_dry(h[j] + dh[j]); q[j] += dq[j]; }     // unknowns += delta
      if (converged) {
        level += 1; it = 0;
        if (level >= L) active = false;
      } else if (it >= p.max_iter || bc_failed) {
        failed = true;
        if (member_valid && is_first) {
          if (p.status) p.status[member] = (err2 == err2 && !bc_failed) ? PR_STATUS_MAX_ITER : PR_STATUS_NAN;
          if (p.fail_level) p.fail_level[member] = level;
        }
        if (is_first) {
          if (p.iters) p.iters[(size_t)member * (L - 1) + (level - 1)] = it;
          if (p.final_error) p.final_error[(size_t)member * (L - 1) + (level - 1)] = sqrt(err2);
        }
        for (int kk = level; kk < L; ++kk) {
          store_level(kk, true);
          if (kk > level && is_first) {
            if (p.iters) p.iters[(size_t)member * (L - 1) + (kk - 1)] = 0;
            if (p.final_error) p.final_error[(size_t)member * (L - 1) + (kk - 1)] = nan("");
          }
          if (owns_last && p.storage_stage) p.storage_stage[(size_t)member * L + kk] = nan("");
        }
        active = false;
      }
    }
  }
  // PHASE: epilogue
  if (member_valid && is_first && !failed) {
    if (p.status) p.status[member] = PR_STATUS_OK;
    if (p.fail_level) p.fail_level[member] = 0;
  }
  __syncwarp();                      // the per-warp scratch is reused by the next member
  }                                  // next ticket
#undef PR_HYD
#undef PC
#undef EL
#undef XW
#undef PR_NODE
}

// Tickets and grid of a persistent launch: one ticket = the members of one warp pass (32/G of them); one CTA per SM
// (the kernels are built for one resident CTA), fewer when the ensemble is smaller than that.
inline unsigned persistent_grid(DevParams& q, int members_per_warp, int warps_per_cta) {
  q.n_tickets = (int)((q.M + members_per_warp - 1) / members_per_warp);
  const int ctas = (q.n_tickets + warps_per_cta - 1) / warps_per_cta;
  const int sms = q.sm_count > 0 ? q.sm_count : 148;
  return (unsigned)(ctas < sms ? ctas : sms);
}

// Launch of one (lanes per member, nodes per lane) family (all CURV x RM variants); defined in pr_ensemble_*.cu so
// that the instantiations compile in parallel.  Returns a cudaError_t as int.
template <int G, int M, int W>
int launch_ensemble_family(const DevParams& p, bool curv, cudaStream_t s);

// STATIC_RM_ = 1: one build per roughness-override mode (the headline family: the mode's selects and loads are
// compiled away); 0: one build that decides at run time (RM = 4) - a third of the instantiations, a few per cent slower.
#define PR_DEFINE_ENSEMBLE_FAMILY(G_, M_, W_, STATIC_RM_)                                                    \
  namespace pr {                                                                                             \
  template <bool CURV, int RM, bool EXACT, bool GST>                                                         \
  static int launch_y_##G_##_##M_(const DevParams& p, cudaStream_t s) {                                      \
    constexpr size_t smem = ensemble_smem_bytes<G_, M_, W_>();                                               \
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");                                      \
    auto kern = pr_ensemble_kernel<G_, M_, W_, CURV, RM, EXACT, GST>;                                        \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                     \
    DevParams q = p;                                                                                         \
    const unsigned grid = persistent_grid(q, 32 / G_, W_);                                                   \
    kern<<<grid, W_ * 32, smem, s>>>(q);                                                                     \
    return (int)cudaGetLastError();                                                                          \
  }                                                                                                          \
  template <bool CURV, int RM, bool EXACT>                                                                   \
  static int launch_x_##G_##_##M_(const DevParams& p, cudaStream_t s) {                                      \
    /* general lumped storage (Brent solve / losses) and gate-controlled rating curves: one build per       */ \
    /* curvature setting, roughness overrides decided at run time (RM = 4), no straight-line variant        */ \
    const bool gst = (p.dn.type == PR_BC_FIXED_DEPTH_STORAGE && (p.dn.st_general || p.dn.st_losses)) ||      \
                     (p.dn.type == PR_BC_RATING_CURVE && p.dn.gated);                                        \
    if (!gst) return launch_y_##G_##_##M_<CURV, RM, EXACT, false>(p, s);                                     \
    return launch_y_##G_##_##M_<CURV, 4, false, true>(p, s);                                                 \
  }                                                                                                          \
  template <bool CURV, int RM>                                                                               \
  static int launch_one_##G_##_##M_(const DevParams& p, cudaStream_t s) {                                    \
    /* the straight-line variant is built for the curvature-free kernels (the ensemble workloads) */         \
    if (!CURV && (p.N - 1) % M_ == 0) return launch_x_##G_##_##M_<CURV, RM, !CURV>(p, s);                     \
    return launch_x_##G_##_##M_<CURV, RM, false>(p, s);                                                      \
  }                                                                                                          \
  template <bool CURV>                                                                                       \
  static int launch_rm_##G_##_##M_(const DevParams& p, cudaStream_t s) {                                     \
    if constexpr (!(STATIC_RM_)) {                                                                           \
      return launch_one_##G_##_##M_<CURV, 4>(p, s);                                                          \
    } else {                                                                                                 \
      switch (rough_mode(p.geo)) {                                                                           \
        case 0: return launch_one_##G_##_##M_<CURV, 0>(p, s);                                                \
        case 1: return launch_one_##G_##_##M_<CURV, 1>(p, s);                                                \
        case 2: return launch_one_##G_##_##M_<CURV, 2>(p, s);                                                \
        default: return launch_one_##G_##_##M_<CURV, 3>(p, s);                                               \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  template <>                                                                                                \
  int launch_ensemble_family<G_, M_, W_>(const DevParams& p, bool curv, cudaStream_t s) {                    \
    return curv ? launch_rm_##G_##_##M_<true>(p, s) : launch_rm_##G_##_##M_<false>(p, s);                    \
  }                                                                                                          \
  }

// IrregularSection reaches of up to 249 nodes: one build per (lanes per member, nodes per lane) and curvature setting
// (run-time roughness mode, every boundary type); 8 and 16 lanes per member pack 4 / 2 short reaches into a warp.
template <int G, int M>
int launch_ensemble_irregular(const DevParams& p, bool curv, cudaStream_t s);

#define PR_DEFINE_ENSEMBLE_IRREGULAR(G_, M_, W_)                                                             \
  namespace pr {                                                                                             \
  template <bool CURV>                                                                                       \
  static int launch_irr_##G_##_##M_(const DevParams& p, cudaStream_t s) {                                    \
    constexpr size_t smem = ensemble_smem_bytes<G_, M_, W_>();                                               \
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");                                      \
    auto kern = pr_ensemble_kernel<G_, M_, W_, CURV, 4, false, true, true>;                                  \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                     \
    DevParams q = p;                                                                                         \
    const unsigned grid = persistent_grid(q, 32 / G_, W_);                                                   \
    kern<<<grid, W_ * 32, smem, s>>>(q);                                                                     \
    return (int)cudaGetLastError();                                                                          \
  }                                                                                                          \
  template <>                                                                                                \
  int launch_ensemble_irregular<G_, M_>(const DevParams& p, bool curv, cudaStream_t s) {                     \
    return curv ? launch_irr_##G_##_##M_<true>(p, s) : launch_irr_##G_##_##M_<false>(p, s);                  \
  }                                                                                                          \
  }

}  // namespace pr
