// pr_long_kernels.cuh - long-reach path (N > 249 nodes: one member no longer fits one warp's registers).
//
// State (current iterate, level constants) lives in HBM; a reach is cut into TILES of 128 cells, one warp per
// (member, tile), lane l owning 4 consecutive cells exactly as in the fused kernel.  One Newton trip of ALL members:
//
//   pr_long_chain   per member (one warp): the <= 32*Kc condensed tile cells + both boundary rows: serial condensation
//                   per lane, parallel cyclic reduction, back-substitution -> update of every tile-boundary node;
//                   convergence test, level / iteration bookkeeping (the member's state machine lives here)
//   pr_long_retire, pr_long_loop_control   member retirement; one more trip? (the WHILE condition of the graph loop)
//   pr_long_fused   per tile: finish the trip (re-assemble - cheaper than storing 9 doubles per node -, solve the tile
//                   interior with its two end nodes known, x += delta; when the level was accepted write it out and
//                   refresh the level constants) and start the next one on the new iterate (node pass, cell pass,
//                   per-lane Schur condensation, warp tree-merge -> ONE condensed cell per tile + partial ||R||^2)
//
// Same algebra as pr_ensemble_kernel.cuh (cell_assemble / merge_cells / PCR); block cyclic reduction is
// applied hierarchically: lane (4 cells) -> tile (32 lanes) -> chain (<= 32 x Kc tiles).
// The iterate is double-buffered (a tile reads its right neighbour's first node while that tile rewrites it).
// Algorithmic HBM traffic is 48 B per node per Newton iteration; the fused tile kernel moves 64 B
// (iterate read, 4 level constants per cell read, iterate written) + the constants rewritten at accepted levels.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "pr_ensemble_kernel.cuh"
#include "pr_irregular.cuh"

namespace pr {

constexpr int kTileCells = 128;    // 32 lanes x 4 cells
constexpr int kLongM = 4;
constexpr int kChainMaxK = 64;     // tile cells per lane in the chain kernel -> N <= 32*64*128 + 1

struct LongParams {
  DevParams p;
  const double* geo;   // derived geometry table [F_COUNT][N] (stage_geometry layout, NP = N)
  const double* geo_tile;  // the same per tile in lane-major order, [T][F_COUNT][kLongM + 1][32]: slot (j, lane) of tile t is
                           // node t*128 + 4*lane + j, so a warp's load of one field is 256 contiguous bytes
  double *xh, *xq;     // current iterate [M][N] (read)
  double *xh_out, *xq_out;   // next iterate (written by the tile kernel): tiles read their right neighbour's first node, so the
                             // update cannot be done in place
  double* pc;          // level constants [M][4][N]
  double* tcell;       // condensed tile cells [M][T][kTileRec]: 10 cell entries + the tile's partial ||R||^2
  double* dchain;      // updates of the tile-boundary nodes [M][T+1][2]
  int *level, *it, *active, *conv, *out_level;   // [M] per-member state machine
  int *status, *fail_level;                      // [M] the run's own record (copied to the caller's arrays when given)
  double *qprev_last, *stage_prev;               // [M] boundary bookkeeping
  GateState* gate;                               // [M] gate-controlled rating curve state
  int* n_done;         // [0] members finished so far, [1] Newton trips made
  int T, Kc;
  int Np;              // row stride of the iterate and of the level constants: N rounded up to a multiple of 4, so that a
                       // lane's four nodes / cells are always one aligned 32-byte group
  long long max_trips;
  cudaGraphConditionalHandle loop;               // WHILE node of the trip loop (graph-driven runs)
};
constexpr int kTileRec = 11;

static __global__ void pr_long_geometry(DevGeom g, int N, double* table) {
  stage_geometry(g, N, N, table, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, [](int i) { return i; });
}

constexpr int kTileSlots = (kLongM + 1) * 32;     // node slots of a tile in the lane-major table

static __global__ void pr_long_geometry_tiles(const double* __restrict__ table, int N, int T, double* __restrict__ tiles) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)T * F_COUNT * kTileSlots;
  if (i >= total) return;
  const int slot = (int)(i % kTileSlots), f = (int)((i / kTileSlots) % F_COUNT), t = (int)(i / ((long long)kTileSlots * F_COUNT));
  const int lane = slot & 31, j = slot >> 5;
  int nd = t * kTileCells + lane * kLongM + j;
  nd = nd < N ? nd : N - 1;
  tiles[i] = table[(size_t)f * N + nd];
}

// Device workspace of one long-reach run (iterate, level constants, tile cells: ~7 GB for config 5).  Workspaces are
// kept in a pool and handed to one run at a time; a run returns its workspace with an event recorded behind its last
// kernel, so the call itself need not wait for the GPU and two streams can each hold a workspace.
struct LongWorkspace {
  static constexpr int kSlots = 32;
  void* ptr[kSlots] = {};
  size_t bytes[kSlots] = {};
  int device = -1;
  bool taken = false;               // a host thread is enqueuing into it
  cudaEvent_t done = nullptr;       // behind the last kernel of the run that used it
  cudaStream_t capture = nullptr;   // private stream the trip loop is captured on
  cudaGraph_t graph = nullptr;      // trip loop of the last run (kept until the workspace is taken again)
  cudaGraphExec_t exec = nullptr;
  int* host_flags = nullptr;        // pinned: (members done, trips) read back by polling runs
  std::vector<std::pair<cudaGraph_t, cudaGraphExec_t>> retired;    // graphs of earlier runs, destroyed once the GPU is past them
  void drop_graph(bool idle = true) {
    if (graph || exec) retired.emplace_back(graph, exec);
    exec = nullptr; graph = nullptr;
    if (!idle) return;
    for (auto& ge : retired) {
      if (ge.second) cudaGraphExecDestroy(ge.second);
      if (ge.first) cudaGraphDestroy(ge.first);
    }
    retired.clear();
  }
  void release() {
    drop_graph();
    for (int i = 0; i < kSlots; ++i) { if (ptr[i]) cudaFree(ptr[i]); ptr[i] = nullptr; bytes[i] = 0; }
  }
  void* get(int slot, size_t n, cudaError_t& e) {
    if (e != cudaSuccess) return nullptr;
    if (slot < 0 || slot >= kSlots) { e = cudaErrorInvalidValue; return nullptr; }
    if (bytes[slot] < n) {
      if (ptr[slot]) cudaFree(ptr[slot]);
      ptr[slot] = nullptr; bytes[slot] = 0;
      e = cudaMalloc(&ptr[slot], n);
      if (e == cudaSuccess) bytes[slot] = n;
    }
    return ptr[slot];
  }
};

struct LongPool {
  static constexpr int kPerDevice = 2;      // concurrent long-reach runs per device (each holds a full workspace)
  std::mutex mu;
  std::condition_variable cv;
  std::vector<LongWorkspace*> all;
  int last_trips_device = -1;
  const int* last_trips = nullptr;          // device counter of the most recent run (pr_long_last_trips)

  // A workspace for a run on the current device whose kernels go to stream s: an idle one if there is one, a new one
  // while fewer than kPerDevice exist, else the stream waits (on the GPU) for the one that frees up first.
  LongWorkspace* acquire(cudaStream_t s, cudaError_t& e) {
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return nullptr;
    std::unique_lock<std::mutex> lock(mu);
    for (;;) {
      LongWorkspace* pick = nullptr;
      int mine = 0;
      for (LongWorkspace* w : all) {
        if (w->device != dev) continue;
        ++mine;
        if (w->taken) continue;
        if (cudaEventQuery(w->done) == cudaSuccess) { pick = w; break; }     // idle
        if (!pick) pick = w;                                                  // busy on the GPU only
      }
      (void)cudaGetLastError();                  // cudaErrorNotReady of the query is not an error
      if (pick && cudaEventQuery(pick->done) != cudaSuccess && mine < kPerDevice) pick = nullptr;   // rather a fresh one
      (void)cudaGetLastError();
      if (!pick && mine < kPerDevice) {
        pick = new LongWorkspace();
        pick->device = dev;
        e = cudaEventCreateWithFlags(&pick->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pick->capture, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&pick->host_flags, 4 * sizeof(int));
        if (e != cudaSuccess) { delete pick; return nullptr; }
        all.push_back(pick);
      }
      if (pick) {
        pick->taken = true;
        const bool idle = cudaEventQuery(pick->done) == cudaSuccess;
        (void)cudaGetLastError();
        e = cudaStreamWaitEvent(s, pick->done, 0);      // orders our kernels behind the previous run (no-op when idle)
        pick->drop_graph(idle);
        return pick;
      }
      cv.wait(lock);                                     // every workspace of this device is being enqueued into
    }
  }
  void give_back(LongWorkspace* w, cudaStream_t s) {
    cudaEventRecord(w->done, s);
    std::lock_guard<std::mutex> lock(mu);
    w->taken = false;
    cv.notify_all();
  }
  void release_all() {
    std::unique_lock<std::mutex> lock(mu);
    for (LongWorkspace* w : all) {
      if (w->taken) continue;
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(w->device);
      cudaEventSynchronize(w->done);
      w->release();
      cudaSetDevice(cur);
    }
    last_trips = nullptr;
  }
};
inline LongPool& long_pool() { static LongPool p; return p; }

template <bool CMP, bool CURV, bool IRR>
int long_reach_run_t(const DevParams& p, cudaStream_t s, std::atomic<long long>& launches, std::string& err);

#ifndef PR_LONG_DECLARE_ONLY   // the kernels and the driver loop: compiled once per variant in pr_long_v*.cu


// PCR over the 32 lanes of a warp for block rows  [l | d | u] y = r  with rank-1 couplings (see the fused kernel:
// the lane that owns a pivot block does the elimination work for its two neighbours - 14 values exchanged per step).
__device__ __forceinline__ void pcr32(double& l1, double& l2, double& d11, double& d12, double& d21, double& d22,
                                      double& u1, double& u2, double& ra, double& rb, const int rows, const int lane,
                                      double& y1, double& y2) {
#define SH(v, src) __shfl_sync(kFull, v, src)
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    if (s >= rows) continue;
    const int up = (lane - s) & 31, dn = (lane + s) & 31;
    const double Dl1 = SH(l1, dn), Dl2 = SH(l2, dn);       // l of lane + s
    const double Uu1 = SH(u1, up), Uu2 = SH(u2, up);       // u of lane - s
    const double idet = fast_rcp(d11 * d22 - d12 * d21);
    const double i11 = d22 * idet, i12 = -d12 * idet, i21 = -d21 * idet, i22 = d11 * idet;
    const double a1 = Dl1 * i11 + Dl2 * i21, a2 = Dl1 * i12 + Dl2 * i22;
    const double o_l1 = -a1 * l1, o_l2 = -a1 * l2, o_d1 = a2 * u1, o_d2 = a2 * u2, o_r = a1 * ra + a2 * rb;
    const double b1 = Uu1 * i11 + Uu2 * i21, b2 = Uu1 * i12 + Uu2 * i22;
    const double o_u1 = -b2 * u1, o_u2 = -b2 * u2, o_e1 = b1 * l1, o_e2 = b1 * l2, o_s = b1 * ra + b2 * rb;
    l1 = SH(o_l1, up);  l2 = SH(o_l2, up);
    d11 -= SH(o_d1, up); d12 -= SH(o_d2, up);
    ra -= SH(o_r, up);
    u1 = SH(o_u1, dn);  u2 = SH(o_u2, dn);
    d21 -= SH(o_e1, dn); d22 -= SH(o_e2, dn);
    rb -= SH(o_s, dn);
  }
#undef SH
  const double idet = fast_rcp(d11 * d22 - d12 * d21);
  y1 = (d22 * ra - d12 * rb) * idet;
  y2 = (d11 * rb - d21 * ra) * idet;
}

__device__ __forceinline__ Cell shfl_cell(const Cell& c, int src) {
  Cell o;
  o.c1 = __shfl_sync(kFull, c.c1, src); o.c2 = __shfl_sync(kFull, c.c2, src); o.c3 = __shfl_sync(kFull, c.c3, src);
  o.c4 = __shfl_sync(kFull, c.c4, src); o.rc = __shfl_sync(kFull, c.rc, src);
  o.m1 = __shfl_sync(kFull, c.m1, src); o.m2 = __shfl_sync(kFull, c.m2, src); o.m3 = __shfl_sync(kFull, c.m3, src);
  o.m4 = __shfl_sync(kFull, c.m4, src); o.rm = __shfl_sync(kFull, c.rm, src);
  return o;
}

// The tile kernel of a Newton trip.  For every (member, tile) one warp:
//   U  finishes trip k-1 - re-assembles the tile on the iterate x_{k-1} (recomputing is cheaper than storing 9 doubles
//      per node), solves the tile interior with its two end nodes known from the chain kernel (PCR), x_k = x_{k-1} +
//      delta; when the chain kernel accepted the level, x_{k-1} is written out as that level and its level constants
//      replace the old ones;
//   C  starts trip k - node pass and cell pass on x_k, per-lane Schur condensation, warp tree-merge -> ONE condensed
//      cell for the tile and the tile's part of ||R||^2, for the chain kernel.
// U and C used to be two launches (K3 of one trip, K1 of the next), each reading the iterate and the level constants
// from HBM; fused, a trip reads them once and writes the iterate once: 64 B per node-iteration (16 read + 32 read +
// 16 written) instead of 112, plus the constants rewritten when a level is accepted.
// FIRST (trip 0): no U part; the level constants of the initial state are formed instead.
#ifndef PR_LONG_CTAS
#define PR_LONG_CTAS 4      // resident CTAs per SM the tile kernel is compiled for (registers = 65536 / (128 * this)): 4 x 128
                            // registers with 72 bytes of spills ran 6 % faster than 3 x 168 without (config 5)
#endif
template <bool FIRST, bool CMP, bool CURV, bool IRR>
__global__ void __launch_bounds__(128, (IRR || CURV || CMP) ? 2 : PR_LONG_CTAS) pr_long_fused(const __grid_constant__ LongParams q) {
  const DevParams& p = q.p;
  const int lane = threadIdx.x & 31;
  // grid = (ceil(M / 4), tiles) - members on x, which has no 65 535 limit: the four warps of a CTA work on the SAME
  // tile of four members, so the tile's geometry lines are fetched into L1 once per CTA
  const int t = blockIdx.y;
  const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= p.M) return;
  const int N = p.N;
  const int c0 = t * kTileCells + lane * kLongM;        // first cell / node of this lane
  int nc = (N - 1) - c0;
  nc = nc < 0 ? 0 : (nc > kLongM ? kLongM : nc);
  const int Np = q.Np;
  const double* xh = q.xh + (size_t)m * Np;
  const double* xq = q.xq + (size_t)m * Np;
  double* xh_out = q.xh_out + (size_t)m * Np;
  double* xq_out = q.xq_out + (size_t)m * Np;
  double* pc = q.pc + (size_t)m * 4 * Np;
  // Every global load of this warp's state is issued here, before anything depends on one of them, so the warp pays
  // the DRAM latency once instead of once per dependent group of loads.
  const int act = q.active[m];             // 1 running, 3 last level accepted (its output is still to be written)
  const int conv = FIRST ? 0 : q.conv[m];
  double h[kLongM + 1], qq[kLongM + 1], pcv[kLongM][4];
  // a lane's four nodes / cells are 32 contiguous, 32-byte aligned bytes (rows are padded to a multiple of 4): vector
  // accesses, and a warp's access is 1 KB contiguous instead of 32 strided doubles (lanes at the end of the reach take
  // the scalar path)
  static_assert(kLongM == 4, "vector access assumes four cells per lane");
  const bool vec = c0 + kLongM < N;
  if (vec) {
    const double4 a = *reinterpret_cast<const double4*>(xh + c0), b = *reinterpret_cast<const double4*>(xq + c0);
    h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = xh[c0 + 4];
    qq[0] = b.x; qq[1] = b.y; qq[2] = b.z; qq[3] = b.w; qq[4] = xq[c0 + 4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
      if (!FIRST) v = *reinterpret_cast<const double4*>(pc + (size_t)f * Np + c0);
      pcv[0][f] = v.x; pcv[1][f] = v.y; pcv[2][f] = v.z; pcv[3][f] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j <= kLongM; ++j) {
      const int nd = c0 + j < N ? c0 + j : N - 1;
      h[j] = xh[nd];
      qq[j] = xq[nd];
    }
#pragma unroll
    for (int j = 0; j < kLongM; ++j) {
      const int c = c0 + j < N - 1 ? c0 + j : N - 2;
#pragma unroll
      for (int f = 0; f < 4; ++f) pcv[j][f] = FIRST ? 0.0 : pc[(size_t)f * Np + c];
    }
  }
  double yL1 = 0.0, yL2 = 0.0, yR1 = 0.0, yR2 = 0.0;      // updates of the tile's first node / of the next tile's
  if (!FIRST) {
    const double* dc = q.dchain + ((size_t)m * (q.T + 1) + t) * 2;
    yL1 = dc[0]; yL2 = dc[1]; yR1 = dc[2]; yR2 = dc[3];
  }
  if (act != 1 && act != 3) return;
  const Rough rg = load_rough<4>(p.geo, m);
  // elimination records of the U part live in shared memory ([record][field][lane] per warp)
  extern __shared__ double long_smem[];
  double* elw = long_smem + (size_t)(threadIdx.x >> 5) * ((kLongM - 1) * 9 * 32);
#define LEL(j, c) elw[((j)*9 + (c)) * 32 + lane]
  NodeVals nv[2];
  // node pass of one node: arithmetic on the derived table, or polyline scans for an IrregularSection node
  const double* tg = q.geo_tile + (size_t)t * F_COUNT * kTileSlots;       // this tile's geometry, slot (j, lane) at j*32 + lane
  auto eval_node = [&](int j, double hh, double qv, NodeVals& out) {
    const int nd = c0 + j < N ? c0 + j : N - 1;
    if (IRR && tg[F_KIND * kTileSlots + j * 32 + lane] == (double)PR_XS_IRREGULAR) node_eval_irregular_call<CURV>(p.geo, nd, hh, qv, rg, p, out, nullptr);     // out of line: ten call sites
    else node_eval<CURV, 4, false, DevParams, CMP>(tg, kTileSlots, j * 32 + lane, hh, qv, rg, p, out);
  };
  // One pass over the lane's cells on the state (h, qq): residuals + Jacobian per cell and the Schur condensation into S.
  //   RECORDS: keep the elimination records (U part);  REFRESH: this state is (becomes) the stored level - its level
  //   constants replace the old ones, in memory and in pcv, once the old ones have been used for this cell
  // the refreshed level constants go back to memory
  auto store_constants = [&]() {
    if (vec) {
#pragma unroll
      for (int f = 0; f < 4; ++f)
        *reinterpret_cast<double4*>(pc + (size_t)f * Np + c0) = make_double4(pcv[0][f], pcv[1][f], pcv[2][f], pcv[3][f]);
    } else {
#pragma unroll
      for (int j = 0; j < kLongM; ++j)
        if (j < nc) {
#pragma unroll
          for (int f = 0; f < 4; ++f) pc[(size_t)f * Np + c0 + j] = pcv[j][f];
        }
    }
  };
  auto pass = [&](const bool records, const bool refresh, Cell& S, double& ss) {
    eval_node(0, h[0], qq[0], nv[0]);
#pragma unroll
    for (int j = 0; j < kLongM; ++j) {
      eval_node(j + 1, h[j + 1], qq[j + 1], nv[(j + 1) & 1]);
      if (j < nc) {
        Cell e;
        ss += cell_assemble(nv[j & 1], nv[(j + 1) & 1], p, pcv[j][0], pcv[j][1], pcv[j][2], pcv[j][3], e);
        if (refresh) level_constants(nv[j & 1], nv[(j + 1) & 1], p, pcv[j][0], pcv[j][1], pcv[j][2], pcv[j][3]);
        if (j == 0) S = e;
        else {
          Elim el;
          merge_cells(S, e, el);
          if (records) {
            LEL(j - 1, 0) = el.i11; LEL(j - 1, 1) = el.i12; LEL(j - 1, 2) = el.i21; LEL(j - 1, 3) = el.i22;
            LEL(j - 1, 4) = el.m1;  LEL(j - 1, 5) = el.m2;  LEL(j - 1, 6) = el.rm;
            LEL(j - 1, 7) = el.c3;  LEL(j - 1, 8) = el.rc;
          }
        }
      }
    }
  };
  Cell S;
  double ss = 0.0;
  if (FIRST) {
    // level 0: the level constants of the initial state (the cells assembled on the way are not used)
    pass(false, true, S, ss);
    store_constants();
  } else {
    // ---------------- U: tile interior of trip k-1 with both end nodes known ----------------
    pass(true, conv != 0, S, ss);
    if (conv) store_constants();
    const int lanes = (N - 1 - t * kTileCells + kLongM - 1) / kLongM;  // lanes of this tile that own cells
    const int Lt = lanes < 32 ? lanes : 32;
    double l1, l2, d11, d12, d21, d22, u1, u2, ra, rb;
    {
      const Cell Pv = shfl_cell(S, (lane - 1) & 31);
      if (lane == 0 || lane >= Lt) { l1 = l2 = 0.0; d11 = 1.0; d12 = 0.0; ra = (lane == 0) ? yL1 : 0.0; }
      else { l1 = Pv.m1; l2 = Pv.m2; d11 = Pv.m3; d12 = Pv.m4; ra = Pv.rm; }
      if (lane == 0 || lane >= Lt) { d21 = 0.0; d22 = 1.0; u1 = u2 = 0.0; rb = (lane == 0) ? yL2 : 0.0; }
      else if (lane == Lt - 1) { d21 = S.c1; d22 = S.c2; u1 = u2 = 0.0; rb = S.rc - S.c3 * yR1 - S.c4 * yR2; }
      else { d21 = S.c1; d22 = S.c2; u1 = S.c3; u2 = S.c4; rb = S.rc; }
    }
    double dh0, dq0;
    pcr32(l1, l2, d11, d12, d21, d22, u1, u2, ra, rb, Lt, lane, dh0, dq0);
    double dhR = __shfl_sync(kFull, dh0, (lane + 1) & 31), dqR = __shfl_sync(kFull, dq0, (lane + 1) & 31);
    if (lane == Lt - 1 || lane == 31) { dhR = yR1; dqR = yR2; }
    double dh[kLongM + 1], dq[kLongM + 1];
    dh[0] = dh0; dq[0] = dq0;
    {
      double rh = dhR, rq = dqR;
#pragma unroll
      for (int j = kLongM - 1; j >= 1; --j) {
        if (j < nc) {
          const double t1 = LEL(j - 1, 6) - LEL(j - 1, 4) * dh0 - LEL(j - 1, 5) * dq0;
          const double t2 = LEL(j - 1, 8) - LEL(j - 1, 7) * rh - p.th_dx * rq;
          dh[j] = LEL(j - 1, 0) * t1 + LEL(j - 1, 1) * t2;
          dq[j] = LEL(j - 1, 2) * t1 + LEL(j - 1, 3) * t2;
          rh = dh[j]; rq = dq[j];
        } else { dh[j] = 0.0; dq[j] = 0.0; }
      }
    }
    // the node right of the lane's cell nc-1 (slot nc) is the next lane's / next tile's first node
#pragma unroll
    for (int j = 1; j <= kLongM; ++j)
      if (j == nc) { dh[j] = dhR; dq[j] = dqR; }
      else if (j > nc) { dh[j] = 0.0; dq[j] = 0.0; }
    // nodes written by this lane: its own cells' left nodes; the very last node N-1 is written by the lane whose
    // last cell ends there
    const int lvl = q.out_level[m];
    const size_t orow = ((size_t)m * p.L + lvl) * (size_t)N;
    const bool last_lane = nc > 0 && c0 + nc == N - 1;       // this lane's last cell ends at the downstream boundary node
#pragma unroll
    for (int j = 0; j <= kLongM; ++j) {
      if (j < nc || (j == nc && last_lane)) {
        const int nd = c0 + j;
        if (conv && p.out_mode == PR_OUT_FULL) { if (p.out_h) p.out_h[orow + nd] = h[j]; if (p.out_q) p.out_q[orow + nd] = qq[j]; }
      }
    }
    if (conv && p.out_mode == PR_OUT_UPSTREAM && t == 0 && lane == 0) {
      if (p.out_h) p.out_h[(size_t)m * p.L + lvl] = h[0];
      if (p.out_q) p.out_q[(size_t)m * p.L + lvl] = qq[0];
    }
    if (act == 3) return;            // the member's last level: nothing left to solve
#pragma unroll
    for (int j = 0; j <= kLongM; ++j) { h[j] = poison_dry(h[j] + dh[j]); qq[j] += dq[j]; }       // x_k
    if (vec) {                 // nc == 4
      *reinterpret_cast<double4*>(xh_out + c0) = make_double4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<double4*>(xq_out + c0) = make_double4(qq[0], qq[1], qq[2], qq[3]);
      if (last_lane) { xh_out[c0 + kLongM] = h[kLongM]; xq_out[c0 + kLongM] = qq[kLongM]; }     // its cell 3 ends at node N-1
    } else {
#pragma unroll
      for (int j = 0; j <= kLongM; ++j) {
        if (j < nc || (j == nc && last_lane)) {
          xh_out[c0 + j] = h[j];
          xq_out[c0 + j] = qq[j];
        }
      }
    }
    __syncwarp();
  }
#undef LEL
  // ---------------- C: condensed cell of the tile on x_k ----------------
  ss = 0.0;
  pass(false, false, S, ss);
  int cnt = nc;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const Cell R = shfl_cell(S, (lane + s) & 31);
    const int rcnt = __shfl_sync(kFull, cnt, (lane + s) & 31);
    if ((lane & (2 * s - 1)) == 0 && lane + s < 32 && rcnt > 0) {
      if (cnt > 0) { Elim dummy; merge_cells(S, R, dummy); }
      else S = R;
      cnt += rcnt;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(kFull, ss, s);
  if (lane == 0) {
    double* tc = q.tcell + ((size_t)m * q.T + t) * kTileRec;
    tc[0] = S.c1; tc[1] = S.c2; tc[2] = S.c3; tc[3] = S.c4; tc[4] = S.rc;
    tc[5] = S.m1; tc[6] = S.m2; tc[7] = S.m3; tc[8] = S.m4; tc[9] = S.rm;
    tc[10] = ss;      // summed in tile order by the chain kernel: the norm does not depend on the launch's timing
  }
}

// One warp per member: chain of tile cells + boundary rows, convergence and bookkeeping.
// Shared memory per warp: Kc-1 elimination records x 10 doubles x 32 lanes.
template <bool IRR>
__global__ void __launch_bounds__(32) pr_long_chain(const __grid_constant__ LongParams q) {
  extern __shared__ double rec[];
  const DevParams& p = q.p;
  const int m = blockIdx.x, lane = threadIdx.x;
  if (q.active[m] != 1) return;
  const int N = p.N, L = p.L, T = q.T, Kc = q.Kc;
  const int level = q.level[m];
  const int it = q.it[m] + 1;
  const double* xh = q.xh + (size_t)m * q.Np;
  const double* xq = q.xq + (size_t)m * q.Np;
  const Rough rg = load_rough<4>(p.geo, m);
#define REC(j, c) rec[((j)*10 + (c)) * 32 + lane]
  // ---- per-lane serial condensation of Kc tile cells ----
  const int t0 = lane * Kc;
  int nt = T - t0;
  nt = nt < 0 ? 0 : (nt > Kc ? Kc : nt);
  Cell S;
  double ss_tiles = 0.0;
  for (int j = 0; j < nt; ++j) {
    const double* tc = q.tcell + ((size_t)m * T + t0 + j) * kTileRec;
    ss_tiles += tc[10];
    Cell e;
    e.c1 = tc[0]; e.c2 = tc[1]; e.c3 = tc[2]; e.c4 = tc[3]; e.rc = tc[4];
    e.m1 = tc[5]; e.m2 = tc[6]; e.m3 = tc[7]; e.m4 = tc[8]; e.rm = tc[9];
    if (j == 0) S = e;
    else {
      // generic merge record: the eliminated node sits between two CONDENSED cells, so c4 is not a constant
      const double c4 = e.c4;
      Elim el;
      merge_cells(S, e, el);
      REC(j - 1, 0) = el.i11; REC(j - 1, 1) = el.i12; REC(j - 1, 2) = el.i21; REC(j - 1, 3) = el.i22;
      REC(j - 1, 4) = el.m1; REC(j - 1, 5) = el.m2; REC(j - 1, 6) = el.rm;
      REC(j - 1, 7) = el.c3; REC(j - 1, 8) = el.rc; REC(j - 1, 9) = c4;
    }
  }
  const int Lc = (T + Kc - 1) / Kc;      // lanes with tile cells; chain rows 0..Lc
  // ---- boundary rows ----
  BcRow U, D;
  U.res = 0.0; U.dh = 1.0; U.dq = 0.0; U.stage_rec = 0.0; U.fail = false;
  D = U;
  const double hyd_up = p.up.series ? p.up.series[(size_t)m * p.up.series_stride + level] : 0.0;
  const double hyd_dn = p.dn.series ? p.dn.series[(size_t)m * p.dn.series_stride + level] : 0.0;
  if (lane == 0) {
    NodeVals nvb; NodeConv kc;
    if (IRR && q.geo[F_KIND * N] == (double)PR_XS_IRREGULAR) node_eval_irregular_call<false>(p.geo, 0, xh[0], xq[0], rg, p, nvb, &kc);
    else node_eval<false, 4, true>(q.geo, N, 0, xh[0], xq[0], rg, p, nvb, &kc);
    U = bc_eval<false>(p.up, m, level, hyd_up, xh[0], xq[0], 0.0, 0.0, p.dt, p.g, kc, nvb.T);
  }
  if (lane == Lc) {
    NodeVals nvb; NodeConv kc;
    if (IRR && q.geo[F_KIND * N + N - 1] == (double)PR_XS_IRREGULAR) node_eval_irregular_call<false>(p.geo, N - 1, xh[N - 1], xq[N - 1], rg, p, nvb, &kc);
    else node_eval<false, 4, true>(q.geo, N, N - 1, xh[N - 1], xq[N - 1], rg, p, nvb, &kc);
    D = bc_eval<true>(p.dn, m, level, hyd_dn, xh[N - 1], xq[N - 1], q.qprev_last[m], q.stage_prev[m], p.dt, p.g, kc, nvb.T, &q.gate[m]);
  }
  const double Ures = __shfl_sync(kFull, U.res, 0), Dres = __shfl_sync(kFull, D.res, Lc);
  const bool bc_failed = __any_sync(kFull, U.fail || D.fail);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss_tiles += __shfl_xor_sync(kFull, ss_tiles, s);      // fixed order
  const double err = sqrt(ss_tiles + Ures * Ures + Dres * Dres);
  // ---- chain rows + PCR ----
  double l1, l2, d11, d12, d21, d22, u1, u2, ra, rb;
  {
    const Cell Pv = shfl_cell(S, (lane - 1) & 31);
    if (lane == 0) { l1 = l2 = 0.0; d11 = U.dh; d12 = U.dq; ra = -U.res; }
    else if (lane <= Lc) { l1 = Pv.m1; l2 = Pv.m2; d11 = Pv.m3; d12 = Pv.m4; ra = Pv.rm; }
    else { l1 = l2 = 0.0; d11 = 1.0; d12 = 0.0; ra = 0.0; }
    if (lane < Lc) { d21 = S.c1; d22 = S.c2; u1 = S.c3; u2 = S.c4; rb = S.rc; }
    else if (lane == Lc) { d21 = D.dh; d22 = D.dq; u1 = u2 = 0.0; rb = -D.res; }
    else { d21 = 0.0; d22 = 1.0; u1 = u2 = 0.0; rb = 0.0; }
  }
  double y1, y2;
  pcr32(l1, l2, d11, d12, d21, d22, u1, u2, ra, rb, Lc + 1, lane, y1, y2);
  const double yR1 = __shfl_sync(kFull, y1, (lane + 1) & 31), yR2 = __shfl_sync(kFull, y2, (lane + 1) & 31);
  // ---- back-substitution over the lane's tile boundaries; write the chain-node updates ----
  double* dc = q.dchain + (size_t)m * (T + 1) * 2;
  if (lane <= Lc) {
    const int node0 = lane < Lc ? t0 : T;      // chain node index of this lane's row (row Lc = last node)
    dc[2 * node0] = y1; dc[2 * node0 + 1] = y2;
  }
  {
    double rh = yR1, rq = yR2;
    for (int j = nt - 1; j >= 1; --j) {
      const double t1 = REC(j - 1, 6) - REC(j - 1, 4) * y1 - REC(j - 1, 5) * y2;
      const double t2 = REC(j - 1, 8) - REC(j - 1, 7) * rh - REC(j - 1, 9) * rq;
      const double a = REC(j - 1, 0) * t1 + REC(j - 1, 1) * t2;
      const double b = REC(j - 1, 2) * t1 + REC(j - 1, 3) * t2;
      dc[2 * (t0 + j)] = a; dc[2 * (t0 + j) + 1] = b;
      rh = a; rq = b;
    }
  }
#undef REC
  // ---- the member's state machine (preissmann.py:122-161) ----
  if (lane == 0) {
    const bool converged = !bc_failed && err < p.tol;
    q.conv[m] = converged ? 1 : 0;
    q.out_level[m] = level;
    if (converged) {
      if (p.iters) p.iters[(size_t)m * (L - 1) + (level - 1)] = it;
      if (p.final_error) p.final_error[(size_t)m * (L - 1) + (level - 1)] = err;
      q.it[m] = 0;
      q.level[m] = level + 1;
    } else {
      q.it[m] = it;
      if (it >= p.max_iter || bc_failed) {
        if (p.iters) p.iters[(size_t)m * (L - 1) + (level - 1)] = it;
        if (p.final_error) p.final_error[(size_t)m * (L - 1) + (level - 1)] = err;
        q.status[m] = (err == err && !bc_failed) ? PR_STATUS_MAX_ITER : PR_STATUS_NAN;
        q.fail_level[m] = level;
        q.active[m] = 2;           // failed: the update pass of this trip skips the member, pr_long_retire clears it
        atomicAdd(q.n_done, 1);
      }
    }
  }
  if (lane == Lc) {
    const bool converged = !bc_failed && err < p.tol;
    if (converged) {
      q.qprev_last[m] = xq[N - 1];
      if (p.dn.type == PR_BC_FIXED_DEPTH_STORAGE) {
        q.stage_prev[m] = D.stage_rec;
        if (p.storage_stage) p.storage_stage[(size_t)m * L + level] = D.stage_rec;
      }
    }
  }
}

// After the chain kernel: retire members whose last level was accepted (one trip later: the tile pass in between writes that level out).
static __global__ void pr_long_retire(const __grid_constant__ LongParams q) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= q.p.M) return;
  const int a = q.active[m];
  if (a == 1 && q.level[m] >= q.p.L) q.active[m] = 3;                 // last level accepted: the next tile pass writes it out
  else if (a == 3) { q.active[m] = 0; atomicAdd(q.n_done, 1); }       // ... which has happened
  else if (a == 2) q.active[m] = 0;                                   // failed in this trip (counted by the chain kernel)
}

// End of a trip in a graph-driven run: one more trip while members are left (and the bound on the trips holds).
static __global__ void pr_long_loop_control(const __grid_constant__ LongParams q, const int graph_driven) {
  const int trips = q.n_done[1] + 1;
  q.n_done[1] = trips;
  if (graph_driven) cudaGraphSetConditional(q.loop, (q.n_done[0] < q.p.M && trips < q.max_trips) ? 1u : 0u);
}

static __global__ void pr_long_init_state(const __grid_constant__ LongParams q) {
  const DevParams& p = q.p;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.M * p.N;
  if (i >= total) return;
  const int m = (int)(i / p.N), nd = (int)(i % p.N);
  const double h = p.ic_h[(size_t)m * p.ic_stride + nd], qv = p.ic_q[(size_t)m * p.ic_stride + nd];
  q.xh[(size_t)m * q.Np + nd] = h; q.xq[(size_t)m * q.Np + nd] = qv;
  if (p.out_mode == PR_OUT_FULL) {
    if (p.out_h) p.out_h[(size_t)m * p.L * p.N + nd] = h;
    if (p.out_q) p.out_q[(size_t)m * p.L * p.N + nd] = qv;
  } else if (nd == 0) {
    if (p.out_h) p.out_h[(size_t)m * p.L] = h;
    if (p.out_q) p.out_q[(size_t)m * p.L] = qv;
  }
  if (nd == p.N - 1) {
    q.qprev_last[m] = qv;
    double st = q.geo[F_Z * p.N + nd] + h;
    if (p.dn.type == PR_BC_FIXED_DEPTH_STORAGE && p.dn.st_losses) {      // initial stage = Y - energy_loss (solver.py:101-108)
      const Rough rg = load_rough<4>(p.geo, m);
      NodeVals t;
      NodeConv kc;
      if (p.geo.irr_offset && q.geo[F_KIND * p.N + nd] == (double)PR_XS_IRREGULAR) node_eval_irregular(p.geo, nd, h, qv, rg, p, t, &kc);
      else node_eval<false, 4, true>(q.geo, p.N, nd, h, qv, rg, p, t, &kc);
      const double V = qv / kc.A;
      st -= kc.Sf * p.dn.st_length + p.dn.st_kq * (V * V) / (2.0 * p.g);
    }
    q.stage_prev[m] = st;
    gate_init(q.gate[m], p.dn.member_rc ? p.dn.member_rc[m] : p.dn.rc);
    if (p.storage_stage) p.storage_stage[(size_t)m * p.L] = st;
    q.level[m] = 1; q.it[m] = 0; q.conv[m] = 0; q.out_level[m] = 0;
    q.active[m] = p.L > 1 ? 1 : 0;
    if (p.L <= 1) atomicAdd(q.n_done, 1);
    q.status[m] = PR_STATUS_OK;
    q.fail_level[m] = 0;
  }
}

// NaN-fill the levels a failed member never reached.
static __global__ void pr_long_nanfill(const __grid_constant__ LongParams q) {
  const DevParams& p = q.p;
  const int m = blockIdx.x;          // members on x (no 65 535 limit), chunks of the member's rows on y
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    if (p.status) p.status[m] = q.status[m];
    if (p.fail_level) p.fail_level[m] = q.fail_level[m];
  }
  if (q.status[m] == PR_STATUS_OK) return;
  const int fl = q.fail_level[m];
  const size_t row = (p.out_mode == PR_OUT_FULL) ? (size_t)p.N : 1;
  const size_t n = (size_t)(p.L - fl) * row;
  for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) {
    const size_t o = ((size_t)m * p.L + fl) * row + i;
    if (p.out_h) p.out_h[o] = nan("");
    if (p.out_q) p.out_q[o] = nan("");
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    for (int k = fl + 1; k < p.L; ++k) {
      if (p.iters) p.iters[(size_t)m * (p.L - 1) + (k - 1)] = 0;
      if (p.final_error) p.final_error[(size_t)m * (p.L - 1) + (k - 1)] = nan("");
    }
    if (p.storage_stage) for (int k = fl; k < p.L; ++k) p.storage_stage[(size_t)m * p.L + k] = nan("");
  }
}

// The run: set-up kernels, then Newton trips (chain, retire, loop control, fused tile kernel) until every member has finished.  The number of
// trips is data dependent, so the trip loop is a CUDA-graph WHILE node whose condition the last kernel of a trip sets
// on the device: the host enqueues one graph launch and returns - no per-trip synchronisation, no read-back.  (Two
// trips per loop body because the iterate is double-buffered.)  PR_LONG_POLL=1, or a driver without conditional
// nodes, falls back to a host loop that feeds trips in chunks and polls the done-counter one chunk behind.
template <bool CMP, bool CURV, bool IRR>
int long_reach_run_t(const DevParams& p, cudaStream_t s, std::atomic<long long>& launches,
                            std::string& err) {
  auto fail = [&](int code, const std::string& msg) { err = msg; return code; };
  const int N = p.N, M = p.M;
  const int T = (N - 1 + kTileCells - 1) / kTileCells;
  const int Kc = (T + 30) / 31;           // tile cells per lane so that the chain has <= 32 rows
  if (Kc > kChainMaxK) return fail(PR_ERR_UNSUPPORTED, "long-reach path: n_nodes exceeds 32*64*128");
  if (T > 65535) return fail(PR_ERR_UNSUPPORTED, "long-reach path: more than 65535 tiles");
  cudaError_t e = cudaSuccess;
  LongPool& pool = long_pool();
  LongWorkspace* wsp = pool.acquire(s, e);
  if (!wsp) return fail(PR_ERR_CUDA, std::string("long-reach workspace: ") + cudaGetErrorString(e));
  LongWorkspace& ws = *wsp;
  LongParams q;
  q.p = p;
  q.T = T; q.Kc = Kc;
  q.max_trips = (long long)(p.L - 1) * (p.max_iter > 0 ? p.max_iter : 1) + 2;
  int slot = 0;
  auto dalloc = [&](size_t bytes) -> void* { return ws.get(slot++, bytes, e); };
  double* geo = (double*)dalloc(sizeof(double) * F_COUNT * (size_t)N);
  q.geo = geo;
  double* geo_tile = (double*)dalloc(sizeof(double) * F_COUNT * kTileSlots * (size_t)T);
  q.geo_tile = geo_tile;
  const int Np = (N + 3) & ~3;
  q.Np = Np;
  q.xh = (double*)dalloc(sizeof(double) * (size_t)M * Np);
  q.xq = (double*)dalloc(sizeof(double) * (size_t)M * Np);
  q.xh_out = (double*)dalloc(sizeof(double) * (size_t)M * Np);
  q.xq_out = (double*)dalloc(sizeof(double) * (size_t)M * Np);
  q.pc = (double*)dalloc(sizeof(double) * (size_t)M * 4 * Np);
  q.tcell = (double*)dalloc(sizeof(double) * (size_t)M * T * kTileRec);
  q.dchain = (double*)dalloc(sizeof(double) * (size_t)M * (T + 1) * 2);
  q.qprev_last = (double*)dalloc(sizeof(double) * M);
  q.stage_prev = (double*)dalloc(sizeof(double) * M);
  q.gate = (GateState*)dalloc(sizeof(GateState) * M);
  q.level = (int*)dalloc(sizeof(int) * M); q.it = (int*)dalloc(sizeof(int) * M);
  q.active = (int*)dalloc(sizeof(int) * M); q.conv = (int*)dalloc(sizeof(int) * M);
  q.out_level = (int*)dalloc(sizeof(int) * M);
  q.status = (int*)dalloc(sizeof(int) * M); q.fail_level = (int*)dalloc(sizeof(int) * M);
  q.n_done = (int*)dalloc(sizeof(int) * 2);
  auto bail = [&](const std::string& what) {
    pool.give_back(wsp, s);
    return fail(PR_ERR_CUDA, what + ": " + cudaGetErrorString(e));
  };
  if (e != cudaSuccess) return bail("long-reach workspace");
  cudaMemsetAsync(q.n_done, 0, 2 * sizeof(int), s);
  cudaMemsetAsync(q.active, 0, sizeof(int) * M, s);

  pr_long_geometry<<<(N + 255) / 256, 256, 0, s>>>(p.geo, N, geo);
  {
    const long long cells = (long long)T * F_COUNT * kTileSlots;
    pr_long_geometry_tiles<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(geo, N, T, geo_tile);
  }
  const long long total = (long long)M * N;
  pr_long_init_state<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(q);
  const dim3 tile_grid((unsigned)((M + 3) / 4), (unsigned)T);
  const size_t tile_smem = sizeof(double) * 4 * (kLongM - 1) * 9 * 32;   // elimination records, 4 warps per CTA
  pr_long_fused<true, CMP, CURV, IRR><<<tile_grid, 128, tile_smem, s>>>(q);      // level constants of the initial state + first condensation
  launches.fetch_add(4);
  const size_t chain_smem = sizeof(double) * (size_t)(Kc > 1 ? Kc - 1 : 1) * 10 * 32;
  e = cudaFuncSetAttribute(pr_long_chain<IRR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem);
  if (e != cudaSuccess) return bail("long-reach chain kernel");
  // one Newton trip of all members on stream cs; the iterate buffers swap roles after it
  auto enqueue_trip = [&](cudaStream_t cs, int graph_driven) {
    pr_long_chain<IRR><<<M, 32, chain_smem, cs>>>(q);                               // solve the chain of tile cells, decide
    pr_long_retire<<<(M + 127) / 128, 128, 0, cs>>>(q);
    pr_long_loop_control<<<1, 1, 0, cs>>>(q, graph_driven);
    pr_long_fused<false, CMP, CURV, IRR><<<tile_grid, 128, tile_smem, cs>>>(q);     // finish this trip, start the next
    std::swap(q.xh, q.xh_out);
    std::swap(q.xq, q.xq_out);
  };
  bool graph_done = false;
  if (p.L > 1 && !std::getenv("PR_LONG_POLL")) {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    cudaError_t ge = cudaGraphCreate(&g, 0);
    if (ge == cudaSuccess) ge = cudaGraphConditionalHandleCreate(&q.loop, g, 1, cudaGraphCondAssignDefault);
    cudaGraph_t body = nullptr;
    if (ge == cudaSuccess) {
      cudaGraphNodeParams np = {};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = q.loop;
      np.conditional.type = cudaGraphCondTypeWhile;
      np.conditional.size = 1;
      cudaGraphNode_t node;
      ge = cudaGraphAddNode(&node, g, nullptr, 0, &np);
      if (ge == cudaSuccess) body = np.conditional.phGraph_out[0];
    }
    if (ge == cudaSuccess) ge = cudaStreamBeginCaptureToGraph(ws.capture, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (ge == cudaSuccess) {
      enqueue_trip(ws.capture, 1);
      enqueue_trip(ws.capture, 1);
      cudaGraph_t same = nullptr;
      ge = cudaStreamEndCapture(ws.capture, &same);
    }
    if (ge == cudaSuccess) ge = cudaGraphInstantiate(&ex, g, 0);
    if (ge == cudaSuccess) ge = cudaGraphLaunch(ex, s);
    if (ge == cudaSuccess) {
      ws.graph = g; ws.exec = ex;
      launches.fetch_add(1);             // the trips are counted on the device (pr_long_last_trips)
      graph_done = true;
    } else {                             // no conditional nodes on this driver: clean up and poll instead
      (void)cudaGetLastError();
      if (ex) cudaGraphExecDestroy(ex);
      if (g) cudaGraphDestroy(g);
    }
  }
  if (!graph_done && p.L > 1) {
    // host-fed trips, two per chunk; the done-counter of chunk c is read while chunk c+1 runs
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (auto& v : ev) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v, cudaEventDisableTiming);
    volatile int* flags = ws.host_flags;
    flags[0] = flags[2] = 0;
    for (long long chunk = 0; e == cudaSuccess && chunk * 2 < q.max_trips; ++chunk) {
      enqueue_trip(s, 0);
      enqueue_trip(s, 0);
      launches.fetch_add(8);
      const int k = (int)(chunk & 1);
      e = cudaMemcpyAsync(ws.host_flags + 2 * k, q.n_done, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaEventRecord(ev[k], s);
      if (chunk > 0 && e == cudaSuccess) {
        e = cudaEventSynchronize(ev[1 - k]);
        if (flags[2 * (1 - k)] >= M) break;
      }
    }
    for (auto& v : ev) if (v) cudaEventDestroy(v);
    if (e != cudaSuccess) return bail("long-reach trips");
  }
  pr_long_nanfill<<<dim3((unsigned)M, 16), 256, 0, s>>>(q);
  launches.fetch_add(1);
  e = cudaGetLastError();
  {
    std::lock_guard<std::mutex> lock(pool.mu);
    pool.last_trips = q.n_done + 1;
    pool.last_trips_device = ws.device;
  }
  pool.give_back(wsp, s);
  if (e != cudaSuccess) return fail(PR_ERR_CUDA, std::string("long-reach path: ") + cudaGetErrorString(e));
  return PR_OK;
}

#endif  // PR_LONG_DECLARE_ONLY

#ifdef PR_LONG_DECLARE_ONLY   // the dispatcher lives in the library's main unit only: in a pr_long_v*.cu unit it would
                              // instantiate all five variants next to the one that unit is for
inline int long_reach_run(const DevParams& p, bool has_curv, bool has_compound, bool has_irregular, cudaStream_t s,
                          std::atomic<long long>& launches, std::string& err) {
  // IrregularSection nodes: one more set of kernels (compound arithmetic for the trapezoid nodes of a mixed reach).
  if (has_irregular) {
    return has_curv ? long_reach_run_t<true, true, true>(p, s, launches, err)
                    : long_reach_run_t<true, false, true>(p, s, launches, err);
  }
  // centre-line curvature: compiled with the compound-section node pass only (a curved reach with floodplains is the
  // shipped gerd case; a curved prismatic reach takes the same kernels, the floodplain terms select to nothing)
  if (has_curv) return long_reach_run_t<true, true, false>(p, s, launches, err);
  return has_compound ? long_reach_run_t<true, false, false>(p, s, launches, err)
                      : long_reach_run_t<false, false, false>(p, s, launches, err);
}

#endif

}  // namespace pr
