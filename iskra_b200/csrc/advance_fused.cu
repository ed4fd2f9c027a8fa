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
//     another patch or finishes its rows; rows outside the window (fast particles between sorts)
//     use the global field / global REDs directly;
//   * every warp walks one contiguous range of rows, so its window moves tile by tile;
//   * cell indices use the exact three-instruction division (pic_device.cuh).
// The particle arithmetic (gather, push, boundary, cell index) is bit-identical to the oracle;
// the deposit differs from the reference's sequential sum only in summation order.
#include <cstdlib>

#include "pic_device.cuh"

namespace {

constexpr int MISS_LIMIT = 8;   // re-anchor when more rows of a batch miss the window

// Per-warp shared windows anchored on the 8x8-cell tile the warp is working on:
//   E window   WE x WE nodes (double2), origin tile*8 - (WE-9)/2   -> gather
//   rho window WR x WR nodes (double),  origin tile*8 - (WR-9)/2   -> deposit
// The E window is the larger one: a gather miss stalls the whole warp on an L2 round trip,
// a deposit miss only issues fire-and-forget global REDs.
// The rho window is the central WR x WR part of the E window (origin + (WE-WR)/2).
struct Window {
  int ei0, ej0;     // node coordinates of the E window's lower-left corner
  bool anchored;
};

template <int WR>
__device__ __forceinline__ void flush_rho(double *rho, const Window &w, int off, const GridDev &g, double *u, int lane) {
#pragma unroll
  for (int k = 0; k < (WR * WR + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WR * WR) break;
    const double v = rho[e];
    if (v != 0.0) {
      atomicAdd(&u[(int64_t)(w.ei0 + off + (e % WR)) + (int64_t)(w.ej0 + off + (e / WR)) * g.nx], v);
      rho[e] = 0.0;
    }
  }
}

template <int WE>
__device__ __forceinline__ void load_E(double2 *sE, const Window &w, const GridDev &g, const double2 *__restrict__ E2,
                                       int lane) {
#pragma unroll
  for (int k = 0; k < (WE * WE + 31) / 32; ++k) {
    const int e = k * 32 + lane;
    if (e >= WE * WE) break;
    sE[e] = __ldg(&E2[(int64_t)(w.ei0 + (e % WE)) + (int64_t)(w.ej0 + (e / WE)) * g.nx]);
  }
}

__device__ __forceinline__ int clamp_origin(int o, int n, int wn) {
  if (o > n - wn) o = n - wn;
  return o < 0 ? 0 : o;
}

template <int WE, int WR, int WARPS, int MINB, int MX, int MY>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_advance_tiled(double *__restrict__ X, double *__restrict__ Y, double *__restrict__ VX, double *__restrict__ VY,
                double *__restrict__ VZ, const double *__restrict__ WG, int64_t *cnt, GridDev g,
                const double2 *__restrict__ E2, double qm, double dt, double *u, int *status,
                unsigned long long *vmax2, int64_t n_sorted) {
  constexpr int mode_x = MX, mode_y = MY;   // after_push modes are compile-time: no mode branches per row
  extern __shared__ double2 s_dyn[];        // per warp: E window (double2), then the rho windows (double)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 *sE = s_dyn + warp * (WE * WE);
  double *rho = (double *)(s_dyn + WARPS * (WE * WE)) + warp * (WR * WR);
  unsigned n_gmiss = 0, n_dmiss = 0, n_anchor = 0;
  for (int e = lane; e < WR * WR; e += 32) rho[e] = 0.0;
  __syncwarp();
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
  Window w{0, 0, false};
  unsigned dead_total = 0;
  double vm2 = 0.0;   // max |v|^2 of the rows of this lane (bound used by the MCC pruning)

  // two passes over the same loop body: rg = 0 the sorted range (windowed), rg = 1 the tail slice.
  // Inside a range rows are addressed with 32-bit offsets from per-range base pointers.
  for (int rg = 0; rg < 2; ++rg) {
    const bool windowed = rg == 0;
    const int64_t gbeg = windowed ? rbeg : tbeg;
    const int cntr = (int)((windowed ? rend : tend) - gbeg);   // rows of this range (< 2^31 by construction)
    if (cntr <= 0) continue;
    double *__restrict__ Xr = X + gbeg, *__restrict__ Yr = Y + gbeg, *__restrict__ VXr = VX + gbeg;
    double *__restrict__ VYr = VY + gbeg, *__restrict__ VZr = VZ + gbeg;
    const double *__restrict__ WGr = WG + gbeg;
    int p = lane;
    double px = 0, py = 0, vx = 0, vy = 0, vz = 0, wq = 0;
    if (p < cntr) { px = Xr[p]; py = Yr[p]; vx = VXr[p]; vy = VYr[p]; vz = VZr[p]; wq = WGr[p]; }
    for (int b0 = 0; b0 < cntr; b0 += 32) {
      const bool in_range = p < cntr;
      const int pn = p + 32;
      double nx_ = 0, ny_ = 0, nvx_ = 0, nvy_ = 0, nvz_ = 0, nwq_ = 0;
      if (pn < cntr) { nx_ = Xr[pn]; ny_ = Yr[pn]; nvx_ = VXr[pn]; nvy_ = VYr[pn]; nvz_ = VZr[pn]; nwq_ = WGr[pn]; }

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
      bool fit = ing && w.anchored && (unsigned)(i - 1 - w.ei0) < (unsigned)(WE - 1) &&
                 (unsigned)(j - 1 - w.ej0) < (unsigned)(WE - 1);
      if (gm) {
        const unsigned miss = gm & ~__ballot_sync(0xffffffffu, fit);
        if (windowed && (!w.anchored || __popc(miss) > MISS_LIMIT)) {
          // move the window to the tile of the first row that missed
          ++n_anchor;
          if (w.anchored) flush_rho<WR>(rho, w, (WE - WR) / 2, g, u, lane);
          const int src = __ffs(miss) - 1;
          const int ti = (__shfl_sync(0xffffffffu, i, src) - 1) >> 3, tj = (__shfl_sync(0xffffffffu, j, src) - 1) >> 3;
          w.ei0 = clamp_origin(ti * 8 - (WE - 9) / 2, g.nx, WE);
          w.ej0 = clamp_origin(tj * 8 - (WE - 9) / 2, g.ny, WE);
          w.anchored = true;
          __syncwarp();
          load_E<WE>(sE, w, g, E2, lane);
          __syncwarp();
          fit = ing && (unsigned)(i - 1 - w.ei0) < (unsigned)(WE - 1) && (unsigned)(j - 1 - w.ej0) < (unsigned)(WE - 1);
          n_gmiss += __popc(gm & ~__ballot_sync(0xffffffffu, fit));
        } else {
          n_gmiss += __popc(miss);
        }
      }
      bool dead_now = false;
      if (live) {
        // ---- gather ----
        double ex = 0.0, ey = 0.0;
        if (ing) {
          const CicW gw = cic_weights(hx, hy);
          double2 e00, e10, e01, e11;
          if (fit) {
            const int o = (j - 1 - w.ej0) * WE + (i - 1 - w.ei0);
            e00 = sE[o]; e10 = sE[o + 1]; e01 = sE[o + WE]; e11 = sE[o + WE + 1];
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
        VXr[p] = vx; VYr[p] = vy; VZr[p] = vz; Yr[p] = py;
        if (dead) {
          Xr[p] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
        } else {
          Xr[p] = px;
          // ---- deposit (new position) ----
          cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
          cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
          if (cell_in_grid(i, j, g.nx, g.ny)) {
            const CicW cw = cic_weights(hx, hy);
            const double d00 = __dmul_rn(cw.w00, wq), d10 = __dmul_rn(cw.w10, wq);
            const double d01 = __dmul_rn(cw.w01, wq), d11 = __dmul_rn(cw.w11, wq);
            const int ri = i - 1 - w.ei0 - (WE - WR) / 2, rj = j - 1 - w.ej0 - (WE - WR) / 2;
            if (w.anchored && (unsigned)ri < (unsigned)(WR - 1) && (unsigned)rj < (unsigned)(WR - 1)) {
              // shared CAS adds: 6 SM-cycles per warp instruction x conflict degree (about 1 here)
              double *r0 = rho + rj * WR + ri;
              // (hand-interleaved CAS sequences and claim/retry rounds were both measured slower than
              //  the compiler's ATOMS.CAST.SPIN loops: 1.62 / 1.58 ms vs 1.47 ms per 62.5 M rows)
              atomicAdd(r0, d00);
              atomicAdd(r0 + 1, d10);
              atomicAdd(r0 + WR, d01);
              atomicAdd(r0 + WR + 1, d11);
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
      if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD) dead_total += __popc(__ballot_sync(0xffffffffu, dead_now));
      __syncwarp();   // rho window updates of this batch are ordered before a possible flush
      p = pn;
      px = nx_; py = ny_; vx = nvx_; vy = nvy_; vz = nvz_; wq = nwq_;
    }
  }
  __syncwarp();
  if (w.anchored) flush_rho<WR>(rho, w, (WE - WR) / 2, g, u, lane);
  if (lane == 0 && dead_total) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)dead_total);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmax(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if (lane == 0 && vm2 > 0.0) atomicMax(vmax2, (unsigned long long)__double_as_longlong(vm2));
  // window statistics (diagnostics + adaptive re-sort): gather misses, deposit misses, window moves
  n_dmiss = __reduce_add_sync(0xffffffffu, n_dmiss);
  if (lane == 0) {
    atomicAdd((unsigned long long *)&cnt[3], (unsigned long long)n_gmiss);
    atomicAdd((unsigned long long *)&cnt[4], (unsigned long long)n_dmiss);
    atomicAdd((unsigned long long *)&cnt[5], (unsigned long long)n_anchor);
  }
}

}  // namespace

int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit, bool from_begin);

template <int WE, int WR, int WARPS, int MINB, int MX, int MY>
static int32_t launch_modes(iskb_species *sp, double dt) {
  iskb_ctx *c = sp->ctx;
  constexpr int SMEM = WARPS * (WE * WE * 16 + WR * WR * 8);
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_advance_tiled<WE, WR, WARPS, MINB, MX, MY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
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
  k_advance_tiled<WE, WR, WARPS, MINB, MX, MY><<<(int)blocks, WARPS * 32, SMEM, c->stream>>>(
      sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5], sp->d_cnt, c->g, c->d_E2, qm, dt,
      sp->d_u, c->d_status, sp->d_vmax2, sp->h_nsorted);
  LAUNCH_CHECK(c);
  ISKB_TRY(prof_end(c));
  if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD) sp->counts_stale = true;
  return ISKB_OK;
}

template <int WE, int WR, int WARPS, int MINB>
static int32_t launch_variant(iskb_species *sp, double dt, int mode_x, int mode_y) {
  switch (mode_x * 3 + mode_y) {
    case 0: return launch_modes<WE, WR, WARPS, MINB, 0, 0>(sp, dt);
    case 1: return launch_modes<WE, WR, WARPS, MINB, 0, 1>(sp, dt);
    case 2: return launch_modes<WE, WR, WARPS, MINB, 0, 2>(sp, dt);
    case 3: return launch_modes<WE, WR, WARPS, MINB, 1, 0>(sp, dt);
    case 4: return launch_modes<WE, WR, WARPS, MINB, 1, 1>(sp, dt);
    case 5: return launch_modes<WE, WR, WARPS, MINB, 1, 2>(sp, dt);
    case 6: return launch_modes<WE, WR, WARPS, MINB, 2, 0>(sp, dt);
    case 7: return launch_modes<WE, WR, WARPS, MINB, 2, 1>(sp, dt);
    default: return launch_modes<WE, WR, WARPS, MINB, 2, 2>(sp, dt);
  }
}

int32_t launch_advance_tiled(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  if (c->g.nx < 20 || c->g.ny < 20)   // windows do not fit small / quasi-1D grids: use the simple kernel
    return launch_advance_simple(sp, dt, mode_x, mode_y, true, false);
  return launch_variant<16, 16, 8, 3>(sp, dt, mode_x, mode_y);
}
