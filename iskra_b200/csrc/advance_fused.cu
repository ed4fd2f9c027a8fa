// Fused advance! + density for one species: the per-timestep particle hot path in ONE pass over
// the particle columns (88 B per particle: read x,y,vx,vy,vz,wg, write x,y,vx,vy,vz).
//   gather E (cloud_in_cell.jl:20-36) -> push (pushers.jl:37-50) -> after_push (wrap.jl:1-33)
//   -> CIC deposit of wg (cloud_in_cell.jl:1-18)
//
// Design (measurements behind it: profiles/r1b_ncu_advance_tiled.md, r1_microbench_warp_ops_b200.txt):
//   * rows are kept sorted by 8x8-cell tile and, inside a tile, round-robin over its cells
//     (sort.cu), so the 32 rows of a batch sit in 32 different cells of one small patch;
//   * every WARP owns private windows of 16x16 nodes in shared memory: the E field of the patch
//     (double2 per node, loaded when the window moves) and the rho accumulator;
//   * gather reads the four corner nodes from the shared E window (4 x LDS.128);
//   * deposit adds the four CIC products into the shared rho window with atomicAdd(double)
//     (a CAS loop on sm_100a: 6 SM-cycles per warp instruction x conflict degree, and the
//     interleaved row order keeps the conflict degree near 1) -- no lane sort, no scan;
//   * the window is flushed with one red.global.add.f64 per non-zero node when the warp moves to
//     another patch or finishes its rows;
//   * rows whose cell is outside the E window (fast particles between sorts, rows appended by
//     ionisation) are NOT handled in line -- one such lane would stall the whole warp on an L2 round
//     trip in about half of all batches -- but queued (row index, 64 per warp in shared memory) and
//     advanced 32 at a time by drain_rows() with global gathers / global REDs;
//   * every warp walks one contiguous range of rows, so its window moves tile by tile;
//   * cell indices use the exact three-instruction division (pic_device.cuh).
// The particle arithmetic (gather, push, boundary, cell index) is bit-identical to the oracle;
// the deposit differs from the reference's sequential sum only in summation order.
#include <cstdlib>
#include <cstring>

#include "surfaces_device.cuh"

namespace {

constexpr int MISS_LIMIT = 8;   // re-anchor when more rows of a batch miss the window

// Per-warp shared windows anchored on the 8x8-cell tile the warp is working on:
//   E window   WE x WE nodes (double2), origin tile*8 - (WE-9)/2   -> gather
//   rho window WR x WR nodes (double),  origin tile*8 - (WR-9)/2   -> deposit
// The E window is the larger one: a gather miss stalls the whole warp on an L2 round trip,
// a deposit miss only issues fire-and-forget global REDs.
// The rho window is the central WR x WR part of the E window (origin + (WE-WR)/2).
template <int WR>
__device__ __forceinline__ void flush_rho(double *rho, int ri0, int rj0, const GridDev &g, double *u, int lane) {
#pragma unroll
  for (int k = 0; k < (WR * WR + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WR * WR) break;
    const double v = rho[e];
    if (v != 0.0) {
      atomicAdd(&u[(int64_t)(ri0 + (e % WR)) + (int64_t)(rj0 + (e / WR)) * g.nx], v);
      rho[e] = 0.0;
    }
  }
}

template <int WE>
__device__ __forceinline__ void load_E(double2 *sE, int ei0, int ej0, const GridDev &g, const double2 *__restrict__ E2,
                                       int lane) {
#pragma unroll
  for (int k = 0; k < (WE * WE + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WE * WE) break;
    sE[e] = __ldg(&E2[(int64_t)(ei0 + (e % WE)) + (int64_t)(ej0 + (e / WE)) * g.nx]);
  }
}

__device__ __forceinline__ int clamp_origin(int o, int n, int wn) {
  if (o > n - wn) o = n - wn;
  return o < 0 ? 0 : o;
}

struct DrainOut {
  double vm2;
  unsigned dead;
  bool too_fast;   // TRACK: check!'s message condition seen on one of this lane's rows
};

// Advance `count` (<= 32) queued rows straight from / to global memory: same arithmetic as the
// windowed path, E from the global field, deposit with global REDs.  Not inlined so that its
// registers do not count against the main loop.
// TRACK: rows in cells next to a surface (track!, track.jl:42-52) are not advanced here either: their row
// index goes to the species' tracked list, and k_advance_tracked (surfaces.cu) advances exactly those rows
// with the cell walk of check! right after this kernel.  An out-of-line walk inside this kernel would cost
// the hot loop its spill-free register allocation (measured: 100 B of spills, 1.67 vs 1.31 ms per launch).
template <int MX, int MY, bool TRACK>
__device__ __forceinline__ DrainOut drain_rows(const uint32_t *q, int count, int lane, double *__restrict__ X,
                                            double *__restrict__ Y, double *__restrict__ VX, double *__restrict__ VY,
                                            double *__restrict__ VZ, const double *__restrict__ WG, const GridDev &g,
                                            const double2 *__restrict__ E2, double qm, double dt, double c1, double *u,
                                            int *status, const uint8_t *__restrict__ trk_cells, uint32_t *trk_list,
                                            unsigned *trk_n, double v_too_fast) {
  DrainOut out{0.0, 0u, false};
  bool dead_now = false;
  if (lane < count) {
    const uint32_t p = q[lane];
    double px = X[p], py = Y[p], vx = VX[p], vy = VY[p], vz = VZ[p];
    const double wq = WG[p];
    int i, j;
    double hx, hy;
    cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
    cell1(py, g.dy, g.rdy, g.fast_div, j, hy);   // in the grid: checked before the row was queued
    // track!: dx == dy on this path, so (i, j) is the cell particle_cell(px, p, st.dh) gives
    bool trk = false;
    if (TRACK) trk = trk_cells[i + j * (g.nx + 1)] != 0;
    if (trk) {
      trk_list[atomicAdd(trk_n, 1u)] = p;
    } else {
    {
      const CicW gw = cic_weights(hx, hy);
      const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
      const double2 e00 = __ldg(&E2[n00]), e10 = __ldg(&E2[n00 + 1]);
      const double2 e01 = __ldg(&E2[n00 + g.nx]), e11 = __ldg(&E2[n00 + g.nx + 1]);
      const double ex = cic_gather(gw, e00.x, e10.x, e01.x, e11.x);
      const double ey = cic_gather(gw, e00.y, e10.y, e01.y, e11.y);
      vx = push_v(vx, ex, c1, qm, dt);
      vy = push_v(vy, ey, c1, qm, dt);
      vz = push_v(vz, 0.0, c1, qm, dt);
    }
    px = push_x(px, vx, dt);
    py = push_x(py, vy, dt);
    if (TRACK) out.too_fast = fmax(fmax(fabs(vx), fabs(vy)), fabs(vz)) > v_too_fast;   // check.jl:41-46
    out.vm2 = fma(vz, vz, fma(vy, vy, vx * vx));
    bool dead = (MX == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, MX);
    if (!dead) dead = (MY == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, MY);
    if (!dead) {
      if (MX == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, MX);
      if (MY == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, MY);
    }
    VX[p] = vx; VY[p] = vy; VZ[p] = vz; Y[p] = py;
    if (dead) {
      X[p] = __longlong_as_double(0x7ff8000000000000LL);
      dead_now = true;
    } else {
      X[p] = px;
      cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
      cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
      if (cell_in_grid(i, j, g.nx, g.ny)) {
        const CicW cw = cic_weights(hx, hy);
        const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
        atomicAdd(&u[n00], __dmul_rn(cw.w00, wq));
        atomicAdd(&u[n00 + 1], __dmul_rn(cw.w10, wq));
        atomicAdd(&u[n00 + g.nx], __dmul_rn(cw.w01, wq));
        atomicAdd(&u[n00 + g.nx + 1], __dmul_rn(cw.w11, wq));
      } else {
        atomicOr(status, ISKB_ST_OOB);
      }
    }
    }   // !trk
  }
  if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD) out.dead = __popc(__ballot_sync(0xffffffffu, dead_now));
  return out;
}

constexpr int QCAP = 64;   // queued rows per warp: at most 31 left over + 32 new ones

// Per-warp shared memory, one contiguous block per warp so that a single base register addresses
// all of it: E window | rho window | miss queue | claim bytes | statistics.
template <int WE, int WR>
struct WarpSmem {
  double2 E[WE * WE];
  double rho[WR * WR];
  uint32_t queue[QCAP];
  unsigned char claim[WR * WR];
  unsigned stats[4];   // gather misses, deposit misses, window moves, discards (lane 0 updates them)
};
// TRACK adds one flag per cell of the E window: the cell is next to a surface (st.cells, build.jl:86-93)
template <int WE, int WR>
struct WarpSmemTrk : WarpSmem<WE, WR> {
  unsigned char trk[WE * WE];
  uint32_t tq[64];     // tracked rows waiting to be appended to the global list, 32 at a time
  unsigned tcnt;       // (one global atomic per 32 rows: same-address atomics cost ~2 ns each on B200)
};
template <int WE, int WR, bool TRACK>
struct SmemOf { typedef WarpSmem<WE, WR> type; };
template <int WE, int WR>
struct SmemOf<WE, WR, true> { typedef WarpSmemTrk<WE, WR> type; };

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

constexpr int NOT_ANCHORED = -(1 << 20);   // window origin that no cell can fit

template <int WE, int WR, int WARPS, int MINB, int MX, int MY, bool CLAIM, bool TRACK>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_advance_tiled(double *__restrict__ X, double *__restrict__ Y, double *__restrict__ VX, double *__restrict__ VY,
                double *__restrict__ VZ, const double *__restrict__ WG, int64_t *cnt, GridDev g,
                const double2 *__restrict__ E2, double qm, double dt, double c1, double *u, int *status,
                unsigned long long *vmax2, int64_t n_sorted, const uint8_t *__restrict__ trk_cells, uint32_t *trk_list,
                unsigned *trk_n, double v_too_fast) {
  typedef typename SmemOf<WE, WR, TRACK>::type Smem;
  constexpr int mode_x = MX, mode_y = MY;   // after_push modes are compile-time: no mode branches per row
  constexpr int OFF = (WE - WR) / 2;        // the rho window is the central part of the E window
  extern __shared__ double2 s_dyn[];
  const int lane = threadIdx.x & 31;
  Smem &sm = ((Smem *)s_dyn)[threadIdx.x >> 5];
  if (lane < 4) sm.stats[lane] = 0;
  for (int e = lane; e < WR * WR; e += 32) sm.rho[e] = 0.0;
  __syncwarp();
  // The loop-carried state is kept deliberately small (32-bit row index, packed window, float
  // velocity bound, statistics in shared memory): the body must fit 80 registers WITHOUT spills,
  // because a spill reload in the body shares its scoreboard with the row prefetch and makes its
  // consumer wait for the prefetch as well (profiles/r1h_ncu_advance_tiled.md).
  bool too_fast = false;             // TRACK: check!'s message condition (check.jl:41-46), reported once per warp
  if constexpr (TRACK) {
    if (lane == 0) sm.tcnt = 0;
    __syncwarp();
  }
  int qn = 0;                        // rows waiting in the queue (warp-uniform)
  int ei0 = NOT_ANCHORED, ej0 = 0;   // node coordinates of the E window's lower-left corner
  float vm2 = 0.0f;                  // upper bound of max |v|^2 of this lane's rows (MCC pruning)

  // Every warp walks one contiguous range of the SORTED rows [0, ns) (its window follows the tiles)
  // and then one slice of the unsorted tail [ns, n) (rows appended by ionisation since the last sort;
  // they never move the window).  Range starts are multiples of 32 rows; rows are addressed by their
  // 32-bit global index.
  for (int rg = 0; rg < 2; ++rg) {
    const bool windowed = rg == 0;
    unsigned row, rend;
    {
      const int64_t n = cnt[CNT_NSLOTS];
      const int64_t ns = n_sorted < n ? n_sorted : n;
      const int64_t nwarps = (int64_t)gridDim.x * WARPS, gwarp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
      const int64_t lo = windowed ? 0 : ns, hi = windowed ? ns : n;
      const int64_t per = ((hi - lo + nwarps - 1) / nwarps + 31) / 32 * 32;
      int64_t beg = lo + gwarp * per;
      if (beg > hi) beg = hi;
      const int64_t end = beg + per < hi ? beg + per : hi;
      if (end <= beg) continue;
      row = (unsigned)beg + lane;
      rend = (unsigned)end;
    }
    const unsigned rlast = rend - 1;
    double px, py, vx, vy, vz, wq;
    {
      const unsigned r = row < rlast ? row : rlast;
      px = X[r]; py = Y[r]; vx = VX[r]; vy = VY[r]; vz = VZ[r]; wq = WG[r];
    }
    for (; row - lane < rend; row += 32) {
      const bool in_range = row < rend;
      // register prefetch of the next batch (index clamped instead of branching)
      const unsigned rn = row + 32 < rlast ? row + 32 : rlast;
      const double nx_ = X[rn], ny_ = Y[rn], nvx_ = VX[rn], nvy_ = VY[rn], nvz_ = VZ[rn], nwq_ = WG[rn];

      const bool live = in_range && !is_dead(px);
      // ---- cell of the old position; window management on it ----
      int i = 0, j = 0;
      double hx = 0, hy = 0;
      bool ing = false;
      if (live) {
        cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
        cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
        ing = cell_in_grid(i, j, g.nx, g.ny);
        if (!ing) atomicOr(status, ISKB_ST_OOB);
      }
      const unsigned gm = __ballot_sync(0xffffffffu, ing);
      bool fit = ing && (unsigned)(i - 1 - ei0) < (unsigned)(WE - 1) && (unsigned)(j - 1 - ej0) < (unsigned)(WE - 1);
      if (gm) {
        unsigned dm = gm & ~__ballot_sync(0xffffffffu, fit);
        if (windowed && (ei0 == NOT_ANCHORED || __popc(dm) > MISS_LIMIT)) {
          // move the window to the tile of the first row that missed
          if (ei0 != NOT_ANCHORED) flush_rho<WR>(sm.rho, ei0 + OFF, ej0 + OFF, g, u, lane);
          const int src = __ffs(dm) - 1;
          const int ti = (__shfl_sync(0xffffffffu, i, src) - 1) >> 3, tj = (__shfl_sync(0xffffffffu, j, src) - 1) >> 3;
          ei0 = clamp_origin(ti * 8 - (WE - 9) / 2, g.nx, WE);
          ej0 = clamp_origin(tj * 8 - (WE - 9) / 2, g.ny, WE);
          if (lane == 0) sm.stats[2] += 1;
          __syncwarp();
          load_E<WE>(sm.E, ei0, ej0, g, E2, lane);
          if constexpr (TRACK) {
            // cell (ei0+1+a, ej0+1+b) <-> window offset b*WE + a (same offset as its lower-left node)
#pragma unroll
            for (int k = 0; k < (WE * WE + 31) / 32; ++k) {
              const int e = k * 32 + lane;
              if (e < WE * WE) sm.trk[e] = trk_cells[(ei0 + 1 + (e % WE)) + (ej0 + 1 + (e / WE)) * (g.nx + 1)];
            }
          }
          __syncwarp();
          fit = ing && (unsigned)(i - 1 - ei0) < (unsigned)(WE - 1) && (unsigned)(j - 1 - ej0) < (unsigned)(WE - 1);
          dm = gm & ~__ballot_sync(0xffffffffu, fit);
        }
        bool trkrow = false;
        if constexpr (TRACK) {
          // track!: rows in cells next to a surface go to the tracked list (warp-aggregated append) and are
          // advanced by k_advance_tracked after this kernel
          trkrow = fit && sm.trk[(j - 1 - ej0) * WE + (i - 1 - ei0)] != 0;
          const unsigned tm = __ballot_sync(0xffffffffu, trkrow);
          if (tm) {
            unsigned nt = sm.tcnt;
            if (trkrow) {
              sm.tq[nt + __popc(tm & lanemask_lt())] = row;
              fit = false;
            }
            nt += __popc(tm);
            __syncwarp();
            if (nt >= 32) {
              unsigned g0 = 0;
              if (lane == 0) g0 = atomicAdd(trk_n, 32u);
              g0 = __shfl_sync(0xffffffffu, g0, 0);
              trk_list[g0 + lane] = sm.tq[lane];
              nt -= 32;
              __syncwarp();
              if (lane < (int)nt) sm.tq[lane] = sm.tq[lane + 32];
            }
            if (lane == 0) sm.tcnt = nt;
            __syncwarp();
          }
        }
        // rows outside the window are queued and advanced later, 32 at a time (drain_rows)
        if (dm) {
          if (ing && !fit && !trkrow) sm.queue[qn + __popc(dm & lanemask_lt())] = row;
          qn += __popc(dm);
          if (lane == 0) sm.stats[0] += __popc(dm);   // (tracked rows had fit == true when dm was formed: not in dm)
        }
      }
      bool dead_now = false;
      bool dep_win = false;            // CLAIM: this lane still has a deposit for the shared window
      double d00 = 0, d10 = 0, d01 = 0, d11 = 0;
      int ci = 0;
      if (live && (fit || !ing)) {
        // ---- gather ----
        double ex = 0.0, ey = 0.0;
        if (ing) {
          const CicW gw = cic_weights(hx, hy);
          const int o = (j - 1 - ej0) * WE + (i - 1 - ei0);
          const double2 e00 = sm.E[o], e10 = sm.E[o + 1], e01 = sm.E[o + WE], e11 = sm.E[o + WE + 1];
          ex = cic_gather(gw, e00.x, e10.x, e01.x, e11.x);
          ey = cic_gather(gw, e00.y, e10.y, e01.y, e11.y);
        }
        // ---- push ----
        vx = push_v(vx, ex, c1, qm, dt);
        vy = push_v(vy, ey, c1, qm, dt);
        vz = push_v(vz, 0.0, c1, qm, dt);
        px = push_x(px, vx, dt);
        py = push_x(py, vy, dt);
        vm2 = fmaxf(vm2, __double2float_ru(fma(vz, vz, fma(vy, vy, vx * vx))));
        if constexpr (TRACK) {   // check!'s message condition, check.jl:41-46 (queued rows: in drain_rows)
          too_fast |= fmax(fmax(fabs(vx), fabs(vy)), fabs(vz)) > v_too_fast;
        }
        // ---- after_push: discards first, then wraps ----
        bool dead = (mode_x == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, mode_x);
        if (!dead) dead = (mode_y == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, mode_y);
        if (!dead) {
          if (mode_x == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, mode_x);
          if (mode_y == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, mode_y);
        }
        VX[row] = vx; VY[row] = vy; VZ[row] = vz; Y[row] = py;
        if (dead) {
          X[row] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
        } else {
          X[row] = px;
          // ---- deposit (new position) ----
          cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
          cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
          if (cell_in_grid(i, j, g.nx, g.ny)) {
            const CicW cw = cic_weights(hx, hy);
            d00 = __dmul_rn(cw.w00, wq); d10 = __dmul_rn(cw.w10, wq);
            d01 = __dmul_rn(cw.w01, wq); d11 = __dmul_rn(cw.w11, wq);
            const int ri = i - 1 - ei0 - OFF, rj = j - 1 - ej0 - OFF;   // fails for NOT_ANCHORED as well
            if ((unsigned)ri < (unsigned)(WR - 1) && (unsigned)rj < (unsigned)(WR - 1)) {
              ci = rj * WR + ri;
              if (CLAIM) {
                dep_win = true;
              } else {
                // shared CAS adds (ATOMS.CAST.SPIN loops)
                double *r0 = sm.rho + ci;
                atomicAdd(r0, d00);
                atomicAdd(r0 + 1, d10);
                atomicAdd(r0 + WR, d01);
                atomicAdd(r0 + WR + 1, d11);
              }
            } else {
              const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
              atomicAdd(&u[n00], d00);
              atomicAdd(&u[n00 + 1], d10);
              atomicAdd(&u[n00 + g.nx], d01);
              atomicAdd(&u[n00 + g.nx + 1], d11);
              atomicAdd(&sm.stats[1], 1u);
            }
          } else {
            atomicOr(status, ISKB_ST_OOB);
          }
        }
      }
      if (CLAIM) {
        // Deposit rounds without atomics.  Each round the lanes write their lane number into the claim
        // byte of their CELL; the lanes that read their own number back sit in pairwise different
        // cells, so for each of the four corners their nodes are pairwise different and a plain
        // load-add-store is safe.  The rest retry.  (compute-sanitizer racecheck reports the claim
        // bytes as write-after-write hazards: several lanes storing to one byte is the point, any one
        // of them may win; it reports no read-after-write or write-after-read hazard on the windows.)  (The shared CAS loops this replaces kept the
        // LSU data pipe at 85 % of its wavefront peak: profiles/r1e_ncu_advance_tiled.md.)
        unsigned pend = __ballot_sync(0xffffffffu, dep_win);
        double *r0 = sm.rho + ci;
        while (pend) {
          if (dep_win) sm.claim[ci] = (unsigned char)lane;
          __syncwarp();
          const bool win = dep_win && sm.claim[ci] == (unsigned char)lane;
          if (win) r0[0] = __dadd_rn(r0[0], d00);
          __syncwarp();
          if (win) r0[1] = __dadd_rn(r0[1], d10);
          __syncwarp();
          if (win) r0[WR] = __dadd_rn(r0[WR], d01);
          __syncwarp();
          if (win) r0[WR + 1] = __dadd_rn(r0[WR + 1], d11);
          __syncwarp();
          if (win) dep_win = false;
          pend = __ballot_sync(0xffffffffu, dep_win);
        }
      }
      if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD) {
        const unsigned dmask = __ballot_sync(0xffffffffu, dead_now);
        if (dmask && lane == 0) sm.stats[3] += __popc(dmask);
      }
      __syncwarp();   // rho window updates of this batch are ordered before a possible flush
      if (qn >= 32) {
        const DrainOut o = drain_rows<MX, MY, TRACK>(sm.queue, 32, lane, X, Y, VX, VY, VZ, WG, g, E2, qm, dt, c1, u, status,
                                                     trk_cells, trk_list, trk_n, v_too_fast);
        vm2 = fmaxf(vm2, __double2float_ru(o.vm2));
        if constexpr (TRACK) too_fast |= o.too_fast;
        if (lane == 0) { sm.stats[3] += o.dead; sm.stats[1] += 32; }
        qn -= 32;
        __syncwarp();
        if (lane < qn) sm.queue[lane] = sm.queue[lane + 32];
        __syncwarp();
      }
      px = nx_; py = ny_; vx = nvx_; vy = nvy_; vz = nvz_; wq = nwq_;
    }
  }
  __syncwarp();
  if (qn > 0) {
    const DrainOut o = drain_rows<MX, MY, TRACK>(sm.queue, qn, lane, X, Y, VX, VY, VZ, WG, g, E2, qm, dt, c1, u, status,
                                                 trk_cells, trk_list, trk_n, v_too_fast);
    vm2 = fmaxf(vm2, __double2float_ru(o.vm2));
    if constexpr (TRACK) too_fast |= o.too_fast;
    if (lane == 0) { sm.stats[3] += o.dead; sm.stats[1] += qn; }
  }
  __syncwarp();
  if constexpr (TRACK) {
    const unsigned nt = sm.tcnt;
    if (nt) {
      unsigned g0 = 0;
      if (lane == 0) g0 = atomicAdd(trk_n, nt);
      g0 = __shfl_sync(0xffffffffu, g0, 0);
      if (lane < (int)nt) trk_list[g0 + lane] = sm.tq[lane];
    }
    if (__any_sync(0xffffffffu, too_fast) && lane == 0) atomicOr(status, ISKB_ST_TOO_FAST);
  }
  if (ei0 != NOT_ANCHORED) flush_rho<WR>(sm.rho, ei0 + OFF, ej0 + OFF, g, u, lane);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmaxf(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if (lane == 0) {
    // float -> double is exact, and the float was rounded up, so this stays an upper bound
    if (vm2 > 0.0f) atomicMax(vmax2, (unsigned long long)__double_as_longlong((double)vm2));
    if (sm.stats[3]) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)sm.stats[3]);
    // window statistics (diagnostics + adaptive re-sort): gather misses, deposit misses, window moves
    atomicAdd((unsigned long long *)&cnt[3], (unsigned long long)sm.stats[0]);
    atomicAdd((unsigned long long *)&cnt[4], (unsigned long long)sm.stats[1]);
    atomicAdd((unsigned long long *)&cnt[5], (unsigned long long)sm.stats[2]);
  }
}

}  // namespace

int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit, bool from_begin);

template <int WE, int WR, int WARPS, int MINB, int MX, int MY, bool CLAIM, bool TRACK = false>
static int32_t launch_modes(iskb_species *sp, double dt) {
  iskb_ctx *c = sp->ctx;
  constexpr int SMEM = WARPS * (int)sizeof(typename SmemOf<WE, WR, TRACK>::type);
  // per device, not per process: set on every launch (a second context may live on another GPU)
  CU_TRY(cudaFuncSetAttribute(k_advance_tiled<WE, WR, WARPS, MINB, MX, MY, CLAIM, TRACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const uint8_t *trk_cells = nullptr;
  double v_too_fast = 0.0;
  if (TRACK) {
    TrackerDev t;
    memset(&t, 0, sizeof(t));
    ISKB_TRY(tracker_prepare(c->tracker, &t));
    trk_cells = t.tracked;
    v_too_fast = t.dh / dt;
    if (!sp->d_trk_list) {
      CU_TRY(cudaMalloc(&sp->d_trk_list, sp->cap * sizeof(uint32_t)));
      CU_TRY(cudaMalloc(&sp->d_trk_n, sizeof(unsigned)));
    }
    CU_TRY(cudaMemsetAsync(sp->d_trk_n, 0, sizeof(unsigned), c->stream));
  }
  const double qm = sp->q / sp->m;
  const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
  int64_t blocks = (bound + (int64_t)1024 * WARPS - 1) / ((int64_t)1024 * WARPS);   // >= 1024 rows per warp
  const int64_t maxb = (int64_t)c->n_sm * MINB;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  ISKB_TRY(sp_vmax_reset(sp));
  sp_touch(sp);
  ISKB_TRY(prof_begin(c));
  k_advance_tiled<WE, WR, WARPS, MINB, MX, MY, CLAIM, TRACK><<<(int)blocks, WARPS * 32, SMEM, c->stream>>>(
      sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5], sp->d_cnt, c->g, c->d_E2, qm, dt,
      0.5 * dt * qm, sp->d_u, c->d_status, sp->d_vmax2, sp->h_nsorted, trk_cells, sp->d_trk_list, sp->d_trk_n, v_too_fast);
  LAUNCH_CHECK(c);
  if (TRACK) ISKB_TRY(launch_advance_tracked_list(sp, dt, MX, MY));   // the rows this kernel left to track! / check!
  ISKB_TRY(prof_end(c));
  if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD || TRACK) sp->counts_stale = true;
  return ISKB_OK;
}

template <int WE, int WR, int WARPS, int MINB, bool CLAIM>
static int32_t launch_variant(iskb_species *sp, double dt, int mode_x, int mode_y) {
  switch (mode_x * 3 + mode_y) {
    case 0: return launch_modes<WE, WR, WARPS, MINB, 0, 0, CLAIM>(sp, dt);
    case 1: return launch_modes<WE, WR, WARPS, MINB, 0, 1, CLAIM>(sp, dt);
    case 2: return launch_modes<WE, WR, WARPS, MINB, 0, 2, CLAIM>(sp, dt);
    case 3: return launch_modes<WE, WR, WARPS, MINB, 1, 0, CLAIM>(sp, dt);
    case 4: return launch_modes<WE, WR, WARPS, MINB, 1, 1, CLAIM>(sp, dt);
    case 5: return launch_modes<WE, WR, WARPS, MINB, 1, 2, CLAIM>(sp, dt);
    case 6: return launch_modes<WE, WR, WARPS, MINB, 2, 0, CLAIM>(sp, dt);
    case 7: return launch_modes<WE, WR, WARPS, MINB, 2, 1, CLAIM>(sp, dt);
    default: return launch_modes<WE, WR, WARPS, MINB, 2, 2, CLAIM>(sp, dt);
  }
}

// advance! with config.tracker on the tiled path (dx == dy, grid large enough for the windows)
int32_t launch_advance_tiled_tracked(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(fields_join(c));
  if (c->g.nx < 20 || c->g.ny < 20 || c->g.dx != c->g.dy) return launch_advance_tracked(sp, dt, mode_x, mode_y, true);
  switch (mode_x * 3 + mode_y) {
    case 0: return launch_modes<16, 16, 8, 3, 0, 0, true, true>(sp, dt);
    case 1: return launch_modes<16, 16, 8, 3, 0, 1, true, true>(sp, dt);
    case 2: return launch_modes<16, 16, 8, 3, 0, 2, true, true>(sp, dt);
    case 3: return launch_modes<16, 16, 8, 3, 1, 0, true, true>(sp, dt);
    case 4: return launch_modes<16, 16, 8, 3, 1, 1, true, true>(sp, dt);
    case 5: return launch_modes<16, 16, 8, 3, 1, 2, true, true>(sp, dt);
    case 6: return launch_modes<16, 16, 8, 3, 2, 0, true, true>(sp, dt);
    case 7: return launch_modes<16, 16, 8, 3, 2, 1, true, true>(sp, dt);
    default: return launch_modes<16, 16, 8, 3, 2, 2, true, true>(sp, dt);
  }
}

int32_t launch_advance_tiled(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(fields_join(c));
  if (c->g.nx < 20 || c->g.ny < 20)   // windows do not fit small / quasi-1D grids: use the simple kernel
    return launch_advance_simple(sp, dt, mode_x, mode_y, true, false);
  // 8 warps x 3 CTAs per SM with claim rounds; the variants measured against it are listed in profiles/r1_ncu_advance_tiled.md
  return launch_variant<16, 16, 8, 3, true>(sp, dt, mode_x, mode_y);
}
