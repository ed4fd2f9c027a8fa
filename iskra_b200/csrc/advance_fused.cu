// Fused advance! + density for one species: the per-timestep particle hot path in ONE pass over
// the particle columns (88 B per particle: read x,y,vx,vy,vz,wg, write x,y,vx,vy,vz).
//   gather E (cloud_in_cell.jl:20-36) -> push (pushers.jl:37-50) -> after_push (wrap.jl:1-33)
//   -> CIC deposit of wg (cloud_in_cell.jl:1-18)
//
// Design (measurements behind it: profiles/r1b_ncu_advance_tiled.md, r1_microbench_warp_ops_b200.txt):
//   * rows are kept sorted by 8x8-cell tile and, inside a tile, round-robin over its cells
//     (sort.cu), so the 32 rows of a batch sit in 32 different cells of one small patch;
//   * every WARP owns a private window of 16x16 nodes in shared memory: the E field of the patch
//     (double2 per node, loaded when the window moves) and the rho accumulator;
//   * gather reads the four corner nodes from the shared E window (4 x LDS.128);
//   * deposit adds the four CIC products into the shared rho window with atomicAdd(double)
//     (a CAS loop on sm_100a: 6 SM-cycles per warp instruction x conflict degree, and the
//     interleaved row order keeps the conflict degree near 1) -- no lane sort, no scan;
//   * the window is flushed with one red.global.add.f64 per non-zero node when the warp moves to
//     another patch or finishes its rows; rows outside the window (fast particles between sorts)
//     use the global field / global REDs directly;
//   * every warp walks one contiguous range of rows, so its window moves tile by tile;
//   * cell indices use the exact three-instruction division (pic_device.cuh).
// The particle arithmetic (gather, push, boundary, cell index) is bit-identical to the oracle;
// the deposit differs from the reference's sequential sum only in summation order.
#include <cstdlib>

#include "pic_device.cuh"

namespace {

constexpr int MISS_LIMIT = 8;       // re-anchor when more rows of a batch miss the window

struct Window {
  int i0, j0;       // node coordinates of the window's lower-left corner
  bool anchored;
};

template <int WN>
__device__ __forceinline__ void flush_rho(double *rho, const Window &w, const GridDev &g, double *u, int lane) {
#pragma unroll
  for (int k = 0; k < (WN * WN + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WN * WN) break;
    const double v = rho[e];
    if (v != 0.0) {
      atomicAdd(&u[(int64_t)(w.i0 + (e % WN)) + (int64_t)(w.j0 + (e / WN)) * g.nx], v);
      rho[e] = 0.0;
    }
  }
}

template <int WN>
__device__ __forceinline__ void load_E(double2 *sE, const Window &w, const GridDev &g, const double2 *__restrict__ E2,
                                       int lane) {
#pragma unroll
  for (int k = 0; k < (WN * WN + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WN * WN) break;
    sE[e] = __ldg(&E2[(int64_t)(w.i0 + (e % WN)) + (int64_t)(w.j0 + (e / WN)) * g.nx]);
  }
}

// Four shared-memory FP64 adds with their CAS loops interleaved (the compiler's atomicAdd(double)
// on shared memory is one serial LDS -> DADD -> ATOMS.CAS chain per address; issuing the four
// chains together hides most of their latency).
template <int WN>
__device__ __forceinline__ void smem_add4(double *w, int o, double d00, double d10, double d01, double d11) {
  unsigned long long *a0 = (unsigned long long *)(w + o), *a1 = a0 + 1, *a2 = a0 + WN, *a3 = a0 + WN + 1;
  unsigned long long o0 = *a0, o1 = *a1, o2 = *a2, o3 = *a3;
  unsigned todo = 0xfu;
  while (todo) {
    if (todo & 1u) {
      const unsigned long long as = o0;
      o0 = atomicCAS(a0, as, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)as) + d00));
      if (o0 == as) todo &= ~1u;
    }
    if (todo & 2u) {
      const unsigned long long as = o1;
      o1 = atomicCAS(a1, as, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)as) + d10));
      if (o1 == as) todo &= ~2u;
    }
    if (todo & 4u) {
      const unsigned long long as = o2;
      o2 = atomicCAS(a2, as, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)as) + d01));
      if (o2 == as) todo &= ~4u;
    }
    if (todo & 8u) {
      const unsigned long long as = o3;
      o3 = atomicCAS(a3, as, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)as) + d11));
      if (o3 == as) todo &= ~8u;
    }
  }
}

template <int WN>
__device__ __forceinline__ bool in_window(const Window &w, int ci, int cj) {
  return ci >= w.i0 && ci < w.i0 + (WN - 1) && cj >= w.j0 && cj < w.j0 + (WN - 1);
}

template <int DEP, int WN, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_advance_tiled(double *__restrict__ X, double *__restrict__ Y, double *__restrict__ VX, double *__restrict__ VY,
                double *__restrict__ VZ, const double *__restrict__ WG, int64_t *cnt, GridDev g,
                const double2 *__restrict__ E2, double qm, double dt, int mode_x, int mode_y, double *u,
                int *status, unsigned long long *vmax2, int64_t n_sorted) {
  extern __shared__ double2 s_dyn[];   // per warp: E window (double2), rho window (double), claim bytes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 *sE = s_dyn + warp * (WN * WN);
  double *rho = (double *)(s_dyn + WARPS * (WN * WN)) + warp * (WN * WN);
  unsigned char *claim = (unsigned char *)((double *)(s_dyn + WARPS * (WN * WN)) + WARPS * (WN * WN)) + warp * (WN * WN);
  unsigned n_gmiss = 0, n_dmiss = 0, n_anchor = 0, n_rounds = 0;
  for (int e = lane; e < WN * WN; e += 32) rho[e] = 0.0;
  __syncwarp();
  constexpr int MARGIN = (WN - 9) / 2;   // cells of slack below the 8x8 tile (the rest above)
  const int64_t n = cnt[CNT_NSLOTS];
  const double c1 = __dmul_rn(__dmul_rn(0.5, dt), qm);
  // Every warp walks one contiguous range of the SORTED rows [0, ns) (its window follows the tiles)
  // plus one slice of the unsorted tail [ns, n) (rows appended by ionisation since the last sort;
  // they take the global path and never move the window).  Both ranges are multiples of 32.
  const int64_t nwarps = (int64_t)gridDim.x * WARPS, gwarp = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t ns = n_sorted < n ? n_sorted : n;
  const int64_t per = ((ns + nwarps - 1) / nwarps + 31) / 32 * 32;
  int64_t rbeg = gwarp * per;
  if (rbeg > ns) rbeg = ns;
  const int64_t rend = rbeg + per < ns ? rbeg + per : ns;
  const int64_t tper = ((n - ns + nwarps - 1) / nwarps + 31) / 32 * 32;
  int64_t tbeg = ns + gwarp * tper;
  if (tbeg > n) tbeg = n;
  const int64_t tend = tbeg + tper < n ? tbeg + tper : n;
  const int64_t nbm = (rend - rbeg + 31) / 32, nbt = (tend - tbeg + 31) / 32;
  Window w{0, 0, false};
  unsigned dead_total = 0;
  double vm2 = 0.0;   // max |v|^2 of the rows of this lane (bound used by the MCC pruning)

  int64_t p = (nbm ? rbeg : tbeg) + lane;
  int64_t lim = nbm ? rend : tend;
  double px = 0, py = 0, vx = 0, vy = 0, vz = 0, wq = 0;
  if (p < lim) { px = X[p]; py = Y[p]; vx = VX[p]; vy = VY[p]; vz = VZ[p]; wq = WG[p]; }
  for (int64_t kb = 0; kb < nbm + nbt; ++kb) {
    const bool windowed = kb < nbm;
    const bool in_range = p < lim;
    // next batch: next of this range, or the first batch of the tail slice
    const int64_t limn = kb + 1 < nbm ? rend : tend;
    const int64_t pn = kb + 1 == nbm ? tbeg + lane : p + 32;
    double nx_ = 0, ny_ = 0, nvx_ = 0, nvy_ = 0, nvz_ = 0, nwq_ = 0;
    if (kb + 1 < nbm + nbt && pn < limn) { nx_ = X[pn]; ny_ = Y[pn]; nvx_ = VX[pn]; nvy_ = VY[pn]; nvz_ = VZ[pn]; nwq_ = WG[pn]; }

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
    bool fit = ing && w.anchored && in_window<WN>(w, i - 1, j - 1);
    if (gm) {
      const unsigned fm = __ballot_sync(0xffffffffu, fit);
      const unsigned miss = gm & ~fm;
      if (windowed && (!w.anchored || __popc(miss) > MISS_LIMIT)) {
        ++n_anchor;
        if (w.anchored) flush_rho<WN>(rho, w, g, u, lane);
        const int src = __ffs(miss) - 1;
        const int ti = (__shfl_sync(0xffffffffu, i, src) - 1) >> 3, tj = (__shfl_sync(0xffffffffu, j, src) - 1) >> 3;
        w.i0 = ti * 8 - MARGIN;
        w.j0 = tj * 8 - MARGIN;
        if (w.i0 > g.nx - WN) w.i0 = g.nx - WN;
        if (w.j0 > g.ny - WN) w.j0 = g.ny - WN;
        if (w.i0 < 0) w.i0 = 0;
        if (w.j0 < 0) w.j0 = 0;
        w.anchored = true;
        __syncwarp();
        load_E<WN>(sE, w, g, E2, lane);
        __syncwarp();
        fit = ing && in_window<WN>(w, i - 1, j - 1);
      }
      n_gmiss += __popc(gm & ~__ballot_sync(0xffffffffu, fit));
    }
    bool dead_now = false;
    bool dep_win = false;
    int dep_o = 0;
    double d00 = 0, d10 = 0, d01 = 0, d11 = 0;
    if (live) {
      // ---- gather ----
      double ex = 0.0, ey = 0.0;
      if (ing) {
        const CicW gw = cic_weights(hx, hy);
        double2 e00, e10, e01, e11;
        if (fit) {
          const int o = (j - 1 - w.j0) * WN + (i - 1 - w.i0);
          e00 = sE[o]; e10 = sE[o + 1]; e01 = sE[o + WN]; e11 = sE[o + WN + 1];
        } else {
          const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
          e00 = __ldg(&E2[n00]); e10 = __ldg(&E2[n00 + 1]);
          e01 = __ldg(&E2[n00 + g.nx]); e11 = __ldg(&E2[n00 + g.nx + 1]);
        }
        ex = cic_gather(gw, e00.x, e10.x, e01.x, e11.x);
        ey = cic_gather(gw, e00.y, e10.y, e01.y, e11.y);
      }
      // ---- push ----
      vx = push_v(vx, ex, c1, qm, dt);
      vy = push_v(vy, ey, c1, qm, dt);
      vz = push_v(vz, 0.0, c1, qm, dt);
      px = push_x(px, vx, dt);
      py = push_x(py, vy, dt);
      vm2 = fmax(vm2, fma(vz, vz, fma(vy, vy, vx * vx)));
      // ---- after_push: discards first, then wraps ----
      bool dead = (mode_x == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, mode_x);
      if (!dead) dead = (mode_y == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, mode_y);
      if (!dead) {
        if (mode_x == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, mode_x);
        if (mode_y == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, mode_y);
      }
      VX[p] = vx; VY[p] = vy; VZ[p] = vz; Y[p] = py;
      if (dead) {
        X[p] = __longlong_as_double(0x7ff8000000000000LL);
        dead_now = true;
      } else {
        X[p] = px;
        // ---- deposit (new position) ----
        cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
        cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
        if (cell_in_grid(i, j, g.nx, g.ny)) {
          const CicW cw = cic_weights(hx, hy);
          d00 = __dmul_rn(cw.w00, wq); d10 = __dmul_rn(cw.w10, wq);
          d01 = __dmul_rn(cw.w01, wq); d11 = __dmul_rn(cw.w11, wq);
          if (w.anchored && in_window<WN>(w, i - 1, j - 1)) {
            dep_o = (j - 1 - w.j0) * WN + (i - 1 - w.i0);
            if (DEP == 2) {
              dep_win = true;
            } else if (DEP == 1) {
              smem_add4<WN>(rho, dep_o, d00, d10, d01, d11);
            } else {
              atomicAdd(&rho[dep_o], d00);
              atomicAdd(&rho[dep_o + 1], d10);
              atomicAdd(&rho[dep_o + WN], d01);
              atomicAdd(&rho[dep_o + WN + 1], d11);
            }
          } else {
            const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
            atomicAdd(&u[n00], d00);
            atomicAdd(&u[n00 + 1], d10);
            atomicAdd(&u[n00 + g.nx], d01);
            atomicAdd(&u[n00 + g.nx + 1], d11);
            ++n_dmiss;
          }
        } else {
          atomicOr(status, ISKB_ST_OOB);
        }
      }
    }
    dead_total += __popc(__ballot_sync(0xffffffffu, dead_now));
    if (DEP == 2) {
      // conflict-free shared-memory deposit: lanes claim their cell; the winners (distinct cells)
      // add their four corners with plain load/add/store in four steps (within a step all winners
      // address different words); lanes that lost a claim retry in the next round.
      unsigned pending = __ballot_sync(0xffffffffu, dep_win);
      while (pending) {
        ++n_rounds;
        if (dep_win) claim[dep_o] = (unsigned char)lane;
        __syncwarp();
        const bool win = dep_win && claim[dep_o] == (unsigned char)lane;
        __syncwarp();
        if (win) rho[dep_o] += d00;
        __syncwarp();
        if (win) rho[dep_o + 1] += d10;
        __syncwarp();
        if (win) rho[dep_o + WN] += d01;
        __syncwarp();
        if (win) rho[dep_o + WN + 1] += d11;
        __syncwarp();
        dep_win = dep_win && !win;
        pending = __ballot_sync(0xffffffffu, dep_win);
      }
    }
    __syncwarp();   // rho window updates of this batch are ordered before a possible flush
    p = pn;
    lim = limn;
    px = nx_; py = ny_; vx = nvx_; vy = nvy_; vz = nvz_; wq = nwq_;
  }
  __syncwarp();
  if (w.anchored) flush_rho<WN>(rho, w, g, u, lane);
  if (lane == 0 && dead_total) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)dead_total);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmax(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if (lane == 0 && vm2 > 0.0) atomicMax(vmax2, (unsigned long long)__double_as_longlong(vm2));
  // window statistics (diagnostics): gather misses, deposit misses, anchors, deposit rounds
  n_dmiss = __reduce_add_sync(0xffffffffu, n_dmiss);
  if (lane == 0) {
    atomicAdd((unsigned long long *)&cnt[3], (unsigned long long)n_gmiss);
    atomicAdd((unsigned long long *)&cnt[4], (unsigned long long)n_dmiss);
    atomicAdd((unsigned long long *)&cnt[5], (unsigned long long)n_anchor);
    atomicAdd((unsigned long long *)&cnt[6], (unsigned long long)n_rounds);
  }
}

}  // namespace

int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit, bool from_begin);

template <int DEP, int WN, int WARPS, int MINB>
static int32_t launch_variant(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  constexpr int SMEM = WARPS * WN * WN * (16 + 8 + 1);
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_advance_tiled<DEP, WN, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const double qm = sp->q / sp->m;
  const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
  int64_t blocks = (bound + (int64_t)1024 * WARPS - 1) / ((int64_t)1024 * WARPS);   // >= 1024 rows per warp
  const int64_t maxb = (int64_t)c->n_sm * MINB;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  ISKB_TRY(sp_vmax_reset(sp));
  ISKB_TRY(prof_begin(c));
  k_advance_tiled<DEP, WN, WARPS, MINB><<<(int)blocks, WARPS * 32, SMEM, c->stream>>>(
      sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5], sp->d_cnt, c->g, c->d_E2, qm, dt, mode_x,
      mode_y, sp->d_u, c->d_status, sp->d_vmax2, sp->h_nsorted);
  LAUNCH_CHECK(c);
  ISKB_TRY(prof_end(c));
  if (mode_x == ISKB_BND_DISCARD || mode_y == ISKB_BND_DISCARD) sp->counts_stale = true;
  return ISKB_OK;
}

int32_t launch_advance_tiled(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  static const int variant = getenv("ISKB_ADV_VARIANT") ? atoi(getenv("ISKB_ADV_VARIANT")) : 0;
  const int wn = variant >= 10 ? 20 : 16;
  if (c->g.nx < wn || c->g.ny < wn)   // window does not fit small / quasi-1D grids: use the simple kernel
    return launch_advance_simple(sp, dt, mode_x, mode_y, true, false);
  switch (variant) {
    case 2: return launch_variant<2, 16, 8, 3>(sp, dt, mode_x, mode_y);
    case 10: return launch_variant<0, 20, 4, 5>(sp, dt, mode_x, mode_y);
    case 12: return launch_variant<2, 20, 4, 5>(sp, dt, mode_x, mode_y);
    case 22: return launch_variant<2, 20, 8, 2>(sp, dt, mode_x, mode_y);
    default: return launch_variant<0, 16, 8, 3>(sp, dt, mode_x, mode_y);
  }
}
