// Surface tracker, walls and electrodes:
//   ParticleInCell/src/pic/surfaces/build.jl   Dict{(cell,cell) -> Surface}      -> per-cell face table in HBM
//   ParticleInCell/src/pic/surfaces/track.jl   track!  (before the push)         -> k_track / inline in k_advance_tracked
//   ParticleInCell/src/pic/surfaces/check.jl   check / check! (after the push)   -> walk_tracked
//   ParticleInCell/src/pic/surfaces/hit.jl, circuit_coupling.jl:44-61  hit!      -> walk_tracked
// The reference keeps a FIFO of tracked particles and pops/pushes tuples; every tuple belongs to one
// particle and its successor depends only on that particle, so the queue is equivalent to an independent
// walk per particle -- one thread each.  The only cross-particle state is the electrodes' collected
// charge (a sum, accumulated with atomics: order differs from the reference's FIFO order, <= 1e-13 rel).
#include <climits>
#include <cstring>

#include "surfaces_device.cuh"

namespace {

constexpr int TPB = 256;

// ---- track!  track.jl:42-52 (operator-level API) ---------------------------------------------------
__global__ void k_track(const double *__restrict__ x, const double *__restrict__ y, const int64_t *__restrict__ cnt,
                        TrackerDev t, int32_t *ti, int32_t *tj, double *thx, double *thy, unsigned long long *counts) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x; p0 < n; p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    bool trk = false;
    if (p < n) {
      const double px = x[p];
      int i = INT_MIN, j = 0;
      double hx = 0, hy = 0;
      if (!is_dead(px)) trk = tracked_cell(t, px, y[p], i, j, hx, hy);
      ti[p] = trk ? i : INT_MIN;
      if (trk) { tj[p] = j; thx[p] = hx; thy[p] = hy; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, trk);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&counts[0], (unsigned long long)__popc(m));
  }
}

// ---- check!  check.jl:39-68 (operator-level API) ---------------------------------------------------
__global__ void k_check(double *x, double *y, double *vx, double *vy, const double *__restrict__ vz,
                        const double *__restrict__ wg, int64_t *cnt, TrackerDev t, double dt, double q,
                        const int32_t *__restrict__ ti, const int32_t *__restrict__ tj, const double *__restrict__ thx,
                        const double *__restrict__ thy, unsigned long long *counts, int *status) {
  const int64_t n = cnt[CNT_NSLOTS];
  const double vmax = __ddiv_rn(t.dh, dt);                                                 // :42
  bool too_fast = false;
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x; p0 < n; p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    bool dead_now = false;
    if (p < n) {
      double px = x[p];
      if (!is_dead(px)) {
        double pvx = vx[p], pvy = vy[p];
        too_fast |= fabs(pvx) > vmax || fabs(pvy) > vmax || fabs(vz[p]) > vmax;            // :43-46
        const int i = ti[p];
        if (i != INT_MIN) {
          double py = y[p];
          const double ox = px, oy = py, ovx = pvx, ovy = pvy;
          dead_now = walk_tracked(t, dt, i, tj[p], thx[p], thy[p], px, py, pvx, pvy, __dmul_rn(q, wg[p]), status);
          if (dead_now) {
            x[p] = nan_dead();                                                             // remove!  :64-66
          } else if (px != ox || py != oy || pvx != ovx || pvy != ovy) {
            x[p] = px; y[p] = py; vx[p] = pvx; vy[p] = pvy;
          }
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, dead_now);
    if (m && (threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)__popc(m));
      atomicAdd(&counts[1], (unsigned long long)__popc(m));
    }
  }
  if (__any_sync(0xffffffffu, too_fast) && (threadIdx.x & 31) == 0) atomicOr(status, ISKB_ST_TOO_FAST);
}

// ---- advance!(part, E, B, dt, config) with a tracker, one pass  ParticleInCell.jl:51-61 -------------
//   track! -> grid_to_particle -> push_particles! -> check! -> after_push -> (density's particle_to_grid)
__global__ void k_advance_tracked(double *x, double *y, double *vx, double *vy, double *vz, const double *__restrict__ wg,
                                  int64_t *cnt, GridDev g, TrackerDev t, const double2 *__restrict__ E2, double q, double qm,
                                  double dt, int mode_x, int mode_y, double *u, int *status, unsigned long long *vmax2,
                                  unsigned long long *counts, const uint32_t *__restrict__ list,
                                  const unsigned *__restrict__ n_list) {
  // list == nullptr: every row; otherwise exactly the rows the tiled advance left to the tracker
  const int64_t n = list ? (int64_t)*n_list : cnt[CNT_NSLOTS];
  double vm2 = 0.0;
  const double c1 = __dmul_rn(__dmul_rn(0.5, dt), qm);
  const double vmax = __ddiv_rn(t.dh, dt);
  bool too_fast = false;
  unsigned n_abs = 0;
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x; p0 < n; p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = p0 + threadIdx.x;
    bool dead_now = false;
    if (k < n) {
      const int64_t p = list ? (int64_t)list[k] : k;
      double px = x[p], py = y[p];
      if (!is_dead(px)) {
        int ti, tj;
        double thx, thy;
        const bool trk = tracked_cell(t, px, py, ti, tj, thx, thy);                        // :56 track!
        double ex, ey;
        if (!gather_E(E2, g, px, py, ex, ey)) atomicOr(status, ISKB_ST_OOB);               // :57
        double nvx = push_v(vx[p], ex, c1, qm, dt);                                        // :59
        double nvy = push_v(vy[p], ey, c1, qm, dt);
        const double nvz = push_v(vz[p], 0.0, c1, qm, dt);
        px = push_x(px, nvx, dt);
        py = push_x(py, nvy, dt);
        too_fast |= fabs(nvx) > vmax || fabs(nvy) > vmax || fabs(nvz) > vmax;
        bool dead = false;
        if (trk) dead = walk_tracked(t, dt, ti, tj, thx, thy, px, py, nvx, nvy, __dmul_rn(q, wg[p]), status);   // :60 check!
        if (dead) ++n_abs;
        vm2 = fmax(vm2, (nvx * nvx + nvy * nvy) + nvz * nvz);
        if (!dead) {                                                                       // :61 after_push
          dead = (mode_x == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, mode_x);
          if (!dead) dead = (mode_y == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, mode_y);
          if (!dead) {
            if (mode_x == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, mode_x);
            if (mode_y == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, mode_y);
          }
        }
        vx[p] = nvx;
        vy[p] = nvy;
        vz[p] = nvz;
        y[p] = py;
        if (dead) {
          x[p] = nan_dead();
          dead_now = true;
        } else {
          x[p] = px;
          if (u) {                                                                         // kinetic.jl:53
            int i, j;
            double hx, hy;
            cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
            cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
            if (!cell_in_grid(i, j, g.nx, g.ny)) {
              atomicOr(status, ISKB_ST_OOB);
            } else {
              const CicW w = cic_weights(hx, hy);
              const double wq = wg[p];
              const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
              atomicAdd(&u[n00], __dmul_rn(w.w00, wq));
              atomicAdd(&u[n00 + 1], __dmul_rn(w.w10, wq));
              atomicAdd(&u[n00 + g.nx], __dmul_rn(w.w01, wq));
              atomicAdd(&u[n00 + g.nx + 1], __dmul_rn(w.w11, wq));
            }
          }
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, dead_now);
    if (m && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)__popc(m));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmax(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if ((threadIdx.x & 31) == 0 && vm2 > 0.0) atomicMax(vmax2, (unsigned long long)__double_as_longlong(vm2));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) n_abs += __shfl_xor_sync(0xffffffffu, n_abs, d);
  if ((threadIdx.x & 31) == 0 && n_abs) atomicAdd(&counts[1], (unsigned long long)n_abs);
  if (__any_sync(0xffffffffu, too_fast) && (threadIdx.x & 31) == 0) atomicOr(status, ISKB_ST_TOO_FAST);
}

inline int grid_for(const iskb_species *sp) {
  const iskb_ctx *c = sp->ctx;
  const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
  int64_t b = (bound + TPB - 1) / TPB;
  const int64_t maxb = (int64_t)c->n_sm * 8;
  if (b > maxb) b = maxb;
  if (b < 1) b = 1;
  return (int)b;
}

// directions of the face table: 0:(i,j-1) 1:(i+1,j) 2:(i,j+1) 3:(i-1,j)
inline int dir_of(int i, int j, int k, int l) {
  if (k == i && l == j - 1) return 0;
  if (k == i + 1 && l == j) return 1;
  if (k == i && l == j + 1) return 2;
  if (k == i - 1 && l == j) return 3;
  return -1;
}

}  // namespace

// ================================ host side ======================================================
static void set_face(iskb_tracker *st, int i, int j, int dir, int sid, bool only_if_absent) {
  const int nx = st->ctx->g.nx, ny = st->ctx->g.ny;
  if (i < 0 || i > nx || j < 0 || j > ny) return;
  uint8_t &f = st->h_face[(size_t)(4 * (i + j * (nx + 1)) + dir)];
  if (only_if_absent && f) return;       // get!  build.jl:28-29,36-42
  f = (uint8_t)sid;
  st->dirty = true;
}

static int32_t add_surface(iskb_tracker *st, int kind, int dof, double area, int *sid_out) {
  if ((int)st->h_kind.size() >= ISKB_MAX_SURFACES) return iskb_fail(ISKB_E_UNSUPPORTED, "more than %d surfaces", ISKB_MAX_SURFACES);
  st->h_kind.push_back(kind);
  st->h_dof.push_back(dof);
  st->h_area.push_back(area);
  *sid_out = (int)st->h_kind.size() - 1;   // id 0 is reserved for "no entry"
  st->dirty = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_create(iskb_ctx *c, int32_t default_kind, iskb_tracker **out) {
  if (!c || !c->has_grid || !out) return iskb_fail(ISKB_E_INVALID, "iskb_grid_set must be called first");
  if (c->tracker) return iskb_fail(ISKB_E_INVALID, "this context already has a surface tracker");
  if (default_kind < ISKB_SURF_PERIODIC || default_kind > ISKB_SURF_REFLECTIVE)
    return iskb_fail(ISKB_E_INVALID, "default surface must be periodic, absorbing or reflective");
  iskb_tracker *st = new iskb_tracker();
  st->ctx = c;
  st->default_kind = default_kind;
  const int nx = c->g.nx, ny = c->g.ny;
  st->h_face.assign((size_t)4 * (nx + 1) * (ny + 1), 0);
  st->h_kind.assign(1, -1);
  st->h_dof.assign(1, -1);
  st->h_area.assign(1, 0.0);
  int sid = 0;
  ISKB_TRY(add_surface(st, default_kind, -1, 0.0, &sid));
  for (int i = 1; i <= nx - 1; ++i) {                    // build_default_surface!  build.jl:33-44
    set_face(st, i, 1, 0, sid, true);                    // ((i,1),(i,0))
    set_face(st, i, ny - 1, 2, sid, true);               // ((i,ny-1),(i,ny))
  }
  for (int j = 1; j <= ny - 1; ++j) {
    set_face(st, 1, j, 3, sid, true);                    // ((1,j),(0,j))
    set_face(st, nx - 1, j, 1, sid, true);               // ((nx-1,j),(nx,j))
  }
  CU_TRY(cudaMalloc(&st->d_counts, 2 * sizeof(unsigned long long)));
  CU_TRY(cudaMemsetAsync(st->d_counts, 0, 2 * sizeof(unsigned long long), c->stream));
  CU_TRY(cudaMalloc(&st->d_dq, (ISKB_MAX_SURFACES + 2) * sizeof(double)));
  CU_TRY(cudaMemsetAsync(st->d_dq, 0, (ISKB_MAX_SURFACES + 2) * sizeof(double), c->stream));
  c->tracker = st;
  *out = st;
  return ISKB_OK;
}

int32_t tracker_free(iskb_tracker *st) {
  if (!st) return ISKB_OK;
  cudaFree(st->d_face); cudaFree(st->d_tracked); cudaFree(st->d_kind); cudaFree(st->d_dof); cudaFree(st->d_area);
  cudaFree(st->d_dq); cudaFree(st->d_ti); cudaFree(st->d_tj); cudaFree(st->d_thx); cudaFree(st->d_thy); cudaFree(st->d_counts);
  delete st;
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_track_surface(iskb_tracker *st, const uint8_t *mask, int32_t kind, int32_t sigma_dof,
                                              double area, int32_t *sid_out) {
  if (!st || !mask) return iskb_fail(ISKB_E_INVALID, "null tracker / mask");
  if (kind < ISKB_SURF_PERIODIC || kind > ISKB_SURF_ELECTRODE_FLOATING) return iskb_fail(ISKB_E_INVALID, "bad surface kind");
  iskb_ctx *c = st->ctx;
  if (kind == ISKB_SURF_ELECTRODE_FLOATING && (sigma_dof < 1 || sigma_dof > c->ps.n_sigma || !(area > 0)))
    return iskb_fail(ISKB_E_INVALID, "a floating electrode needs an existing sigma dof and a positive area");
  const int nx = c->g.nx, ny = c->g.ny;
  int sid = 0;
  ISKB_TRY(add_surface(st, kind, kind == ISKB_SURF_ELECTRODE_FLOATING ? sigma_dof - 1 : -1, area, &sid));
  auto nd = [&](int i, int j) { return mask[(size_t)((i - 1) + (int64_t)(j - 1) * nx)] != 0; };   // 1-based
  for (int i = 1; i <= nx - 1; ++i)                      // build_surface_lookup!  build.jl:46-60
    for (int j = 1; j <= ny - 1; ++j) {
      const bool a = nd(i, j), b = nd(i + 1, j), cc = nd(i + 1, j + 1), d = nd(i, j + 1);
      if (a && b) set_face(st, i, j, 0, sid, false);
      if (b && cc) set_face(st, i, j, 1, sid, false);
      if (cc && d) set_face(st, i, j, 2, sid, false);
      if (d && a) set_face(st, i, j, 3, sid, false);
    }
  if (sid_out) *sid_out = sid;
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_lookup(iskb_tracker *st, int32_t i, int32_t j, int32_t k, int32_t l, int32_t *kind_out) {
  if (!st || !kind_out) return iskb_fail(ISKB_E_INVALID, "null");
  const int nx = st->ctx->g.nx, ny = st->ctx->g.ny;
  *kind_out = -1;
  const int dir = dir_of(i, j, k, l);
  if (dir < 0 || i < 0 || i > nx || j < 0 || j > ny) return ISKB_OK;
  const int sid = st->h_face[(size_t)(4 * (i + j * (nx + 1)) + dir)];
  if (sid) *kind_out = st->h_kind[(size_t)sid];
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_route_hits_to_sigma(iskb_tracker *st, int32_t on) {
  if (!st) return iskb_fail(ISKB_E_INVALID, "null tracker");
  st->route_hits = on != 0;
  return ISKB_OK;
}

template <typename T>
static int32_t up(T **d, const std::vector<T> &h, cudaStream_t s) {
  if (*d) { cudaFree(*d); *d = nullptr; }
  CU_TRY(cudaMalloc(d, h.size() * sizeof(T)));
  CU_TRY(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return ISKB_OK;
}

int32_t tracker_prepare(iskb_tracker *st, TrackerDev *out) {
  iskb_ctx *c = st->ctx;
  const int nx = c->g.nx, ny = c->g.ny;
  if (st->dirty) {
    // cells on either side of a key  (Base.in(::BoundaryCell, st), build.jl:86-93)
    st->h_tracked.assign((size_t)(nx + 1) * (ny + 1), 0);
    static const int di[4] = {0, 1, 0, -1}, dj[4] = {-1, 0, 1, 0};
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i)
        for (int d = 0; d < 4; ++d)
          if (st->h_face[(size_t)(4 * (i + j * (nx + 1)) + d)]) {
            st->h_tracked[(size_t)(i + j * (nx + 1))] = 1;
            const int k = i + di[d], l = j + dj[d];
            if (k >= 0 && k <= nx && l >= 0 && l <= ny) st->h_tracked[(size_t)(k + l * (nx + 1))] = 1;
          }
    CU_TRY(cudaStreamSynchronize(c->stream));
    ISKB_TRY(up(&st->d_face, st->h_face, c->stream));
    ISKB_TRY(up(&st->d_tracked, st->h_tracked, c->stream));
    ISKB_TRY(up(&st->d_kind, st->h_kind, c->stream));
    ISKB_TRY(up(&st->d_dof, st->h_dof, c->stream));
    ISKB_TRY(up(&st->d_area, st->h_area, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    st->dirty = false;
  }
  double *d_sigma = nullptr;
  if (st->route_hits && c->n_ranks > 1)   // every rank would add only its own hits to its replica of sigma
    return iskb_fail(ISKB_E_UNSUPPORTED, "route_hits_to_sigma with more than one rank: sum iskb_surface_charge over the ranks instead");
  if (st->route_hits && c->ps.created && c->ps.n_sigma > 0) ISKB_TRY(poisson_sigma_device(c, &d_sigma));
  out->nx = nx; out->ny = ny;
  out->dh = c->g.dx;                      // create_surface_tracker(grid): dx, ~ = grid.dh  build.jl:96-97
  out->face = st->d_face; out->tracked = st->d_tracked;
  out->s_kind = st->d_kind; out->s_dof = st->d_dof; out->s_area = st->d_area;
  out->s_dq = st->d_dq; out->sigma = d_sigma;
  out->route_hits = st->route_hits && d_sigma;
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_track(iskb_tracker *st, iskb_species *sp, double dt, int64_t *n_tracked) {
  if (!st || !sp || sp->ctx != st->ctx) return iskb_fail(ISKB_E_INVALID, "bad tracker / species");
  (void)dt;   // the reference stores dt in every tuple (track.jl:49); check! gets the same value again
  iskb_ctx *c = st->ctx;
  ISKB_TRY(sp_compact(sp));
  TrackerDev t;
  memset(&t, 0, sizeof(t));
  ISKB_TRY(tracker_prepare(st, &t));
  if (st->trk_cap < sp->cap) {
    cudaFree(st->d_ti); cudaFree(st->d_tj); cudaFree(st->d_thx); cudaFree(st->d_thy);
    CU_TRY(cudaMalloc(&st->d_ti, sp->cap * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&st->d_tj, sp->cap * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&st->d_thx, sp->cap * sizeof(double)));
    CU_TRY(cudaMalloc(&st->d_thy, sp->cap * sizeof(double)));
    st->trk_cap = sp->cap;
  }
  CU_TRY(cudaMemsetAsync(st->d_counts, 0, 2 * sizeof(unsigned long long), c->stream));
  k_track<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, t, st->d_ti, st->d_tj, st->d_thx, st->d_thy,
                                               st->d_counts);
  LAUNCH_CHECK(c);
  st->trk_sp = sp;
  st->trk_rows = sp->h_nslots;
  if (n_tracked) {
    CU_TRY(cudaMemcpyAsync(c->h_scratch, st->d_counts, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *n_tracked = c->h_scratch[0];
  }
  return ISKB_OK;
}

extern "C" int32_t iskb_tracker_check(iskb_tracker *st, iskb_species *sp, double dt, int64_t *n_absorbed, int32_t *too_fast) {
  if (!st || !sp || sp->ctx != st->ctx) return iskb_fail(ISKB_E_INVALID, "bad tracker / species");
  iskb_ctx *c = st->ctx;
  if (st->trk_sp != sp || sp->counts_stale || sp->h_nslots != st->trk_rows)
    return iskb_fail(ISKB_E_INVALID, "iskb_tracker_check must follow iskb_tracker_track + iskb_push on the same species");
  TrackerDev t;
  memset(&t, 0, sizeof(t));
  ISKB_TRY(tracker_prepare(st, &t));
  sp_touch(sp);
  k_check<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5], sp->d_cnt,
                                               t, dt, sp->q, st->d_ti, st->d_tj, st->d_thx, st->d_thy, st->d_counts, c->d_status);
  LAUNCH_CHECK(c);
  sp->counts_stale = true;
  st->trk_sp = nullptr;
  CU_TRY(cudaMemcpyAsync(c->h_scratch, st->d_counts, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  if (n_absorbed) *n_absorbed = c->h_scratch[1];
  c->warn_too_fast = false;
  const int32_t rc = ctx_check_status(c);
  if (too_fast) *too_fast = c->warn_too_fast ? 1 : 0;
  return rc;
}

extern "C" int32_t iskb_surface_charge(iskb_tracker *st, int32_t sid, double *dq_out, int32_t reset) {
  if (!st || sid < 1 || sid >= (int)st->h_kind.size()) return iskb_fail(ISKB_E_INVALID, "bad surface id");
  iskb_ctx *c = st->ctx;
  if (dq_out) {
    CU_TRY(cudaMemcpyAsync(c->h_scratch, st->d_dq + sid, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    memcpy(dq_out, c->h_scratch, sizeof(double));
  }
  if (reset) CU_TRY(cudaMemsetAsync(st->d_dq + sid, 0, sizeof(double), c->stream));
  return ISKB_OK;
}

int32_t launch_advance_tracked(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(fields_join(c));
  TrackerDev t;
  memset(&t, 0, sizeof(t));
  ISKB_TRY(tracker_prepare(c->tracker, &t));
  const double qm = sp->q / sp->m;
  ISKB_TRY(sp_vmax_reset(sp));
  ISKB_TRY(prof_begin(c));
  sp_touch(sp);
  k_advance_tracked<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5],
                                                         sp->d_cnt, c->g, t, c->d_E2, sp->q, qm, dt, mode_x, mode_y,
                                                         deposit ? sp->d_u : nullptr, c->d_status, sp->d_vmax2,
                                                         c->tracker->d_counts, nullptr, nullptr);
  LAUNCH_CHECK(c);
  ISKB_TRY(prof_end(c));
  sp->counts_stale = true;
  return ISKB_OK;
}

// Second half of the tiled advance with a tracker: the rows it appended to the species' tracked list.
int32_t launch_advance_tracked_list(iskb_species *sp, double dt, int mode_x, int mode_y) {
  iskb_ctx *c = sp->ctx;
  TrackerDev t;
  memset(&t, 0, sizeof(t));
  ISKB_TRY(tracker_prepare(c->tracker, &t));
  const double qm = sp->q / sp->m;
  // the list length is only known on the device; a few blocks per SM cover the usual <1 % of the rows
  sp_touch(sp);
  k_advance_tracked<<<c->n_sm * 2, TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5],
                                                      sp->d_cnt, c->g, t, c->d_E2, sp->q, qm, dt, mode_x, mode_y, sp->d_u,
                                                      c->d_status, sp->d_vmax2, c->tracker->d_counts, sp->d_trk_list,
                                                      sp->d_trk_n);
  LAUNCH_CHECK(c);
  sp->counts_stale = true;
  return ISKB_OK;
}
