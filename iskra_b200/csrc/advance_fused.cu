// Fused advance! + density for one species: the per-timestep particle hot path in ONE pass over
// the particle columns (88 B per particle: read x,y,vx,vy,vz,wg, write x,y,vx,vy,vz).
//   gather E (cloud_in_cell.jl:20-36) -> push (pushers.jl:37-50) -> after_push (wrap.jl:1-33)
//   -> CIC deposit of wg (cloud_in_cell.jl:1-18)
//
// Deposition design (no FP64 shared-memory atomics exist; global REDs per particle would cost
// more LSU time than the whole HBM budget):
//   * rows are kept sorted by 8x8-cell tile (sort.cu), so a run of rows lives in a small patch;
//   * every WARP owns a private 16x16-node accumulation window in shared memory that follows its
//     rows (anchored 3 cells outside the current tile);
//   * per batch of 32 rows the warp sorts the lanes by window cell (bitonic network on packed
//     key|lane words), pulls the four CIC products from the source lanes, and runs a segmented
//     inclusive scan over equal cells -- a fixed-shape reduction tree, so the result is
//     deterministic for a given row order;
//   * the last lane of every segment adds its four sums into the window with plain
//     load/add/store (segments have distinct cells; the four corners are issued as four
//     separate warp steps, so no two lanes touch one word in the same step);
//   * the window is flushed with one red.global.add.f64 per non-zero node when the warp moves to
//     another patch or finishes its rows.  Rows that left the window (fast particles between
//     sorts) fall back to global REDs.
#include "pic_device.cuh"

namespace {

constexpr int TPB = 256;
constexpr int WARPS = TPB / 32;
constexpr int WN = 16;              // window nodes per side
constexpr int WCELLS = WN - 1;      // cells per side that fit
constexpr int MARGIN = 3;           // cells of slack around the 8x8 tile
constexpr int CHUNK = 1024;         // rows per warp task
constexpr uint32_t KEY_NONE = 0xff; // lanes that do not deposit through the window

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__device__ __forceinline__ void flush_window(double *win, int wi0, int wj0, const GridDev &g, double *u, int lane) {
#pragma unroll
  for (int k = 0; k < WN * WN / 32; ++k) {
    const int e = k * 32 + lane;
    const double v = win[e];
    if (v != 0.0) {
      const int a = wi0 + (e % WN), b = wj0 + (e / WN);
      atomicAdd(&u[(int64_t)a + (int64_t)b * g.nx], v);
      win[e] = 0.0;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(TPB)
k_advance_tiled(double *__restrict__ X, double *__restrict__ Y, double *__restrict__ VX, double *__restrict__ VY,
                double *__restrict__ VZ, const double *__restrict__ WG, int64_t *cnt, GridDev g,
                const double2 *__restrict__ E2, double qm, double dt, int mode_x, int mode_y, double *u,
                int *status) {
  __shared__ double s_win[WARPS][WN * WN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *win = s_win[warp];
#pragma unroll
  for (int k = 0; k < WN * WN / 32; ++k) win[k * 32 + lane] = 0.0;
  __syncwarp();
  const int64_t n = cnt[CNT_NSLOTS];
  const double c1 = __dmul_rn(__dmul_rn(0.5, dt), qm);
  const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp, nwarps = (int64_t)gridDim.x * WARPS;
  int wi0 = 0, wj0 = 0;
  bool anchored = false;
  unsigned long long dead_total = 0;

  for (int64_t chunk = gwarp; chunk * CHUNK < n; chunk += nwarps) {
    const int64_t cbeg = chunk * CHUNK;
    const int64_t cend = cbeg + CHUNK < n ? cbeg + CHUNK : n;
    // software prefetch of the first batch
    int64_t p = cbeg + lane;
    double px = 0, py = 0, vx = 0, vy = 0, vz = 0, wq = 0;
    if (p < cend) { px = X[p]; py = Y[p]; vx = VX[p]; vy = VY[p]; vz = VZ[p]; wq = WG[p]; }
    for (int64_t b0 = cbeg; b0 < cend; b0 += 32) {
      const bool in_range = p < cend;
      // prefetch next batch
      const int64_t pn = p + 32;
      double nx_ = 0, ny_ = 0, nvx_ = 0, nvy_ = 0, nvz_ = 0, nwq_ = 0;
      if (pn < cend) { nx_ = X[pn]; ny_ = Y[pn]; nvx_ = VX[pn]; nvy_ = VY[pn]; nvz_ = VZ[pn]; nwq_ = WG[pn]; }

      bool live = in_range && !is_dead(px);
      bool dead_now = false;
      int ci = 0, cj = 0;
      CicW w{0, 0, 0, 0};
      bool dep = false;
      if (live) {
        // ---- gather (old position) ----
        int i, j;
        double hx, hy, ex = 0.0, ey = 0.0;
        cell1(px, g.dx, i, hx);
        cell1(py, g.dy, j, hy);
        if (cell_in_grid(i, j, g.nx, g.ny)) {
          const CicW gw = cic_weights(hx, hy);
          const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
          const double2 e00 = __ldg(&E2[n00]), e10 = __ldg(&E2[n00 + 1]);
          const double2 e01 = __ldg(&E2[n00 + g.nx]), e11 = __ldg(&E2[n00 + g.nx + 1]);
          ex = cic_gather(gw, e00.x, e10.x, e01.x, e11.x);
          ey = cic_gather(gw, e00.y, e10.y, e01.y, e11.y);
        } else {
          atomicOr(status, ISKB_ST_OOB);
        }
        // ---- push ----
        vx = push_v(vx, ex, c1, qm, dt);
        vy = push_v(vy, ey, c1, qm, dt);
        vz = push_v(vz, 0.0, c1, qm, dt);
        px = push_x(px, vx, dt);
        py = push_x(py, vy, dt);
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
          cell1(px, g.dx, i, hx);
          cell1(py, g.dy, j, hy);
          if (cell_in_grid(i, j, g.nx, g.ny)) {
            const CicW cw = cic_weights(hx, hy);
            w.w00 = __dmul_rn(cw.w00, wq); w.w10 = __dmul_rn(cw.w10, wq);
            w.w01 = __dmul_rn(cw.w01, wq); w.w11 = __dmul_rn(cw.w11, wq);
            ci = i - 1; cj = j - 1;
            dep = true;
          } else {
            atomicOr(status, ISKB_ST_OOB);
          }
        }
      }
      const unsigned dm = __ballot_sync(0xffffffffu, dead_now);
      if (lane == 0) dead_total += __popc(dm);

      // ---- window management ----
      const unsigned depm = __ballot_sync(0xffffffffu, dep);
      if (depm) {
        bool fits = dep && anchored && ci >= wi0 && ci < wi0 + WCELLS && cj >= wj0 && cj < wj0 + WCELLS;
        unsigned fitm = __ballot_sync(0xffffffffu, fits);
        const unsigned missm = depm & ~fitm;
        if (__popc(missm) > 8 || (!anchored && missm)) {
          // move the window to the patch of the first row that missed
          if (anchored) flush_window(win, wi0, wj0, g, u, lane);
          const int src = __ffs(missm) - 1;
          const int ti = __shfl_sync(0xffffffffu, ci, src) >> 3, tj = __shfl_sync(0xffffffffu, cj, src) >> 3;
          wi0 = ti * 8 - MARGIN; wj0 = tj * 8 - MARGIN;
          if (wi0 > g.nx - WN) wi0 = g.nx - WN;
          if (wj0 > g.ny - WN) wj0 = g.ny - WN;
          if (wi0 < 0) wi0 = 0;
          if (wj0 < 0) wj0 = 0;
          anchored = true;
          fits = dep && ci >= wi0 && ci < wi0 + WCELLS && cj >= wj0 && cj < wj0 + WCELLS;
          fitm = __ballot_sync(0xffffffffu, fits);
        }
        // rows outside the window: global REDs
        if (dep && !fits) {
          const int64_t n00 = (int64_t)ci + (int64_t)cj * g.nx;
          atomicAdd(&u[n00], w.w00);
          atomicAdd(&u[n00 + 1], w.w10);
          atomicAdd(&u[n00 + g.nx], w.w01);
          atomicAdd(&u[n00 + g.nx + 1], w.w11);
        }
        if (fitm) {
          // ---- warp sort of (cell key | lane) ----
          uint32_t key = fits ? (uint32_t)((cj - wj0) * WN + (ci - wi0)) : KEY_NONE;
          uint32_t kv = (key << 5) | (uint32_t)lane;
#pragma unroll
          for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
              const uint32_t other = __shfl_xor_sync(0xffffffffu, kv, jj);
              const bool up = ((lane & k) == 0);            // ascending block
              const bool lower = ((lane & jj) == 0);
              const uint32_t mn = kv < other ? kv : other, mx = kv < other ? other : kv;
              kv = (up == lower) ? mn : mx;
            }
          }
          const int src = (int)(kv & 31u);
          key = kv >> 5;
          double s00 = shfl_d(w.w00, src), s10 = shfl_d(w.w10, src);
          double s01 = shfl_d(w.w01, src), s11 = shfl_d(w.w11, src);
          // ---- segmented inclusive scan over equal keys ----
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ko = __shfl_up_sync(0xffffffffu, key, d);
            const double t00 = shfl_up_d(s00, d), t10 = shfl_up_d(s10, d);
            const double t01 = shfl_up_d(s01, d), t11 = shfl_up_d(s11, d);
            if (lane >= d && ko == key) {
              s00 = __dadd_rn(s00, t00); s10 = __dadd_rn(s10, t10);
              s01 = __dadd_rn(s01, t01); s11 = __dadd_rn(s11, t11);
            }
          }
          const uint32_t knext = __shfl_down_sync(0xffffffffu, key, 1);
          const bool tail = key != KEY_NONE && (lane == 31 || knext != key);
          // four corner steps; within a step all tails address distinct words
          if (tail) win[key] += s00;
          __syncwarp();
          if (tail) win[key + 1] += s10;
          __syncwarp();
          if (tail) win[key + WN] += s01;
          __syncwarp();
          if (tail) win[key + WN + 1] += s11;
          __syncwarp();
        }
      }
      // rotate the prefetched batch in
      p = pn;
      px = nx_; py = ny_; vx = nvx_; vy = nvy_; vz = nvz_; wq = nwq_;
    }
    // keep the window across chunks handled by this warp only if the next chunk is adjacent;
    // chunks of one warp are nwarps*CHUNK rows apart, so flush now.
    if (anchored) {
      flush_window(win, wi0, wj0, g, u, lane);
      anchored = false;
    }
  }
  if (lane == 0 && dead_total) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], dead_total);
}

}  // namespace

int32_t launch_advance_tiled(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  if (c->g.nx < WN || c->g.ny < WN) {
    // window does not fit small / quasi-1D grids: use the simple kernel
    extern int32_t launch_advance_simple(iskb_species *, double, int, int, bool, bool);
    return launch_advance_simple(sp, dt, mode_x, mode_y, true, false);
  }
  const double qm = sp->q / sp->m;
  const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
  int64_t blocks = (bound + (int64_t)CHUNK * WARPS - 1) / ((int64_t)CHUNK * WARPS);
  const int64_t maxb = (int64_t)c->n_sm * 6;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  ISKB_TRY(prof_begin(c));
  k_advance_tiled<<<(int)blocks, TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4],
                                                     sp->col[5], sp->d_cnt, c->g, c->d_E2, qm, dt, mode_x, mode_y,
                                                     sp->d_u, c->d_status);
  LAUNCH_CHECK(c);
  ISKB_TRY(prof_end(c));
  if (mode_x == ISKB_BND_DISCARD || mode_y == ISKB_BND_DISCARD) sp->counts_stale = true;
  return ISKB_OK;
}
