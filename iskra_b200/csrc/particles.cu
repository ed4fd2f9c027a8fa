// Particle operators, one kernel per reference function (the "separate operator" path used by
// the parity tests and by the API-level drop-in), plus rho finalisation.  The fused fast path
// lives in advance_fused.cu.  All kernels are grid-stride with the particle count read from
// device memory, so appends/discards never need a host round trip.
#include "pic_device.cuh"

namespace {

constexpr int TPB = 256;

inline int grid_for(const iskb_species *sp) {
  const iskb_ctx *c = sp->ctx;
  const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
  int64_t b = (bound + TPB - 1) / TPB;
  const int64_t maxb = (int64_t)c->n_sm * 8;
  if (b > maxb) b = maxb;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- particle_cell  ParticleInCell.jl:28-35 ----------------------------------------------------
__global__ void k_cell_index(const double *__restrict__ x, const double *__restrict__ y,
                             const int64_t *__restrict__ cnt, GridDev g, int32_t *ci, int32_t *cj,
                             double *hx, double *hy) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    int i, j;
    double fx, fy;
    cell1(x[p], g.dx, g.rdx, g.fast_div, i, fx);
    cell1(y[p], g.dy, g.rdy, g.fast_div, j, fy);
    ci[p] = i;
    cj[p] = j;
    if (hx) hx[p] = fx;
    if (hy) hy[p] = fy;
  }
}

// ---- grid_to_particle  cloud_in_cell.jl:20-36 : gather_E lives in pic_device.cuh -----------------
__global__ void k_gather(const double *__restrict__ x, const double *__restrict__ y,
                         const int64_t *__restrict__ cnt, GridDev g, const double2 *__restrict__ E2,
                         double *pE, int64_t ld, int *status) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    double ex, ey;
    const double px = x[p];
    if (is_dead(px)) continue;
    if (!gather_E(E2, g, px, y[p], ex, ey)) atomicOr(status, ISKB_ST_OOB);
    pE[p] = ex;
    pE[p + ld] = ey;
    pE[p + 2 * ld] = 0.0;   // Ez == 0 (generalized_poisson.jl:398-410 never writes E[:,:,3])
  }
}

// ---- push_particles!  pushers.jl:8-11,37-50 ----------------------------------------------------
// partE == nullptr: gather on the fly (advance!, ParticleInCell.jl:57-59)
__global__ void k_push(double *x, double *y, double *vx, double *vy, double *vz,
                       const int64_t *__restrict__ cnt, GridDev g, const double2 *__restrict__ E2,
                       const double *__restrict__ pE, int64_t ld, double qm, double dt, int *status, int rz) {
  const int64_t n = cnt[CNT_NSLOTS];
  const double c1 = __dmul_rn(__dmul_rn(0.5, dt), qm);   // (0.5dt)*qm  pushers.jl:41
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double px = x[p], py = y[p];
    if (is_dead(px)) continue;
    double ex, ey, ez = 0.0;
    if (pE) {
      ex = pE[p];
      ey = pE[p + ld];
      ez = pE[p + 2 * ld];
    } else if (!gather_E(E2, g, px, py, ex, ey)) {
      atomicOr(status, ISKB_ST_OOB);
    }
    double nvx = push_v(vx[p], ex, c1, qm, dt);
    const double nvy = push_v(vy[p], ey, c1, qm, dt);
    double nvz = push_v(vz[p], ez, c1, qm, dt);
    double nx_ = push_x(px, nvx, dt);
    if (rz) to_cylindrical(nx_, nvx, nvz, dt);   // push_particles!(::BorisPusher{:rz}, ...)  pushers.jl:13-17
    vx[p] = nvx;
    vy[p] = nvy;
    vz[p] = nvz;
    x[p] = nx_;
    y[p] = push_x(py, nvy, dt);
  }
}

__global__ void k_to_cylindrical(double *x, double *vx, double *vz, const int64_t *__restrict__ cnt, double dt) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    double px = x[p], pvx = vx[p], pvz = vz[p];
    if (is_dead(px)) continue;
    to_cylindrical(px, pvx, pvz, dt);
    x[p] = px; vx[p] = pvx; vz[p] = pvz;
  }
}

// ---- discard! / wrap!  wrap.jl:1-33 ------------------------------------------------------------
__global__ void k_boundary(double *x, double *y, int64_t *cnt, GridDev g, int mode_x, int mode_y) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x; p0 < n; p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    bool dead_now = false;
    if (p < n) {
      double px = x[p], py = y[p];
      if (!is_dead(px)) {
        // discards first (dims in order), then wraps -- the scripts' after_push order
        bool dead = (mode_x == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, mode_x);
        if (!dead) dead = (mode_y == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, mode_y);
        if (!dead) {
          if (mode_x == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, mode_x);
          if (mode_y == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, mode_y);
          x[p] = px;
          y[p] = py;
        } else {
          x[p] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, dead_now);
    if (m && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)__popc(m));
  }
}

// ---- particle_to_grid  cloud_in_cell.jl:1-18 with pu(p) = wg[p] ---------------------------------
__global__ void k_deposit_atomic(const double *__restrict__ x, const double *__restrict__ y,
                                 const double *__restrict__ wg, const int64_t *__restrict__ cnt,
                                 GridDev g, double *u, int *status) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double px = x[p];
    if (is_dead(px)) continue;
    int i, j;
    double hx, hy;
    cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
    cell1(y[p], g.dy, g.rdy, g.fast_div, j, hy);
    if (!cell_in_grid(i, j, g.nx, g.ny)) {
      atomicOr(status, ISKB_ST_OOB);
      continue;
    }
    const CicW w = cic_weights(hx, hy);
    const double q = wg[p];
    const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
    atomicAdd(&u[n00], __dmul_rn(w.w00, q));
    atomicAdd(&u[n00 + 1], __dmul_rn(w.w10, q));
    atomicAdd(&u[n00 + g.nx], __dmul_rn(w.w01, q));
    atomicAdd(&u[n00 + g.nx + 1], __dmul_rn(w.w11, q));
  }
}

// u += sum of the private copies (and clear them for the next launch)
__global__ void k_reduce_copies(double *upriv, int64_t nn, double *u) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn; k += (int64_t)gridDim.x * blockDim.x) {
    double s = u[k];
    for (int c = 0; c < PRIV_COPIES; ++c) {
      s += upriv[(int64_t)c * nn + k];
      upriv[(int64_t)c * nn + k] = 0.0;
    }
    u[k] = s;
  }
}

// n = u ./ cell_volume(grid)   kinetic.jl:53
__global__ void k_density(const double *__restrict__ u, const double *__restrict__ V, double *n,
                          int64_t nn) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn;
       k += (int64_t)gridDim.x * blockDim.x)
    n[k] = __ddiv_rn(u[k], V[k]);
}

// rho .+= part.n .* part.q   ParticleInCell.jl:121
__global__ void k_rho_acc(double *rho, const double *__restrict__ n, double q, int64_t nn) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn;
       k += (int64_t)gridDim.x * blockDim.x)
    rho[k] = __dadd_rn(rho[k], __dmul_rn(n[k], q));
}

struct RhoFin {
  const double *u[8];
  double *n[8];
  double q[8];
  int ns;
};
// fused: for every species n_s = u_s ./ V ; rho = sum_s n_s * q_s (species order)  :118-124
__global__ void k_rho_finalize(RhoFin f, const double *__restrict__ V, double *rho, int64_t nn) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn;
       k += (int64_t)gridDim.x * blockDim.x) {
    const double v = V[k];
    double r = 0.0;
    for (int s = 0; s < f.ns; ++s) {
      const double d = __ddiv_rn(f.u[s][k], v);
      f.n[s][k] = d;
      r = __dadd_rn(r, __dmul_rn(d, f.q[s]));
    }
    rho[k] = r;
  }
}

// The same from the fixed-point deposits of the tile path (advance_tile.cu, add_fixed): n_s = (ufix_s / fscale) ./ V and
// rho from the INTEGER combination sum_s Z_s * ufix_s (Z_s = q_s / q0): that sum is what ranks all-reduce, so rho is
// bit-identical on every rank and from run to run.
struct RhoFix {
  const long long *u[8];
  double *n[8];
  long long z[8];
  int ns;
};
// The per-species densities (species.n) are formed from the fixed-point sums only when somebody asks for them
// (iskb_species_density_download): two node arrays less to write per step.  One rank: rho directly; several ranks:
// the integer charge sum first (all-reduced as integers); rho itself is formed by its consumer (rho_materialize / the FFT solve).
__global__ void k_rho_fixed_local(RhoFix f, const double *__restrict__ V, double q0_over_scale, long long *rho_int, double *rho,
                                  int64_t nn) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn; k += (int64_t)gridDim.x * blockDim.x) {
    long long r = 0;
    for (int s = 0; s < f.ns; ++s) r += f.z[s] * f.u[s][k];
    if (rho) rho[k] = __ddiv_rn(__dmul_rn((double)r, q0_over_scale), V[k]);
    else rho_int[k] = r;
  }
}
__global__ void k_density_fixed(const long long *__restrict__ u, const double *__restrict__ V, double inv_scale, double *n, int64_t nn) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn; k += (int64_t)gridDim.x * blockDim.x)
    n[k] = __ddiv_rn(__dmul_rn((double)u[k], inv_scale), V[k]);
}

// simple one-thread-per-particle advance! (gather + push + after_push [+ atomic deposit]);
// the reference-order building block behind iskb_step when the tiled kernel is not applicable.
__global__ void k_advance_simple(double *x, double *y, double *vx, double *vy, double *vz,
                                 const double *__restrict__ wg, int64_t *cnt, int first_from_begin,
                                 GridDev g, const double2 *__restrict__ E2, double qm, double dt,
                                 int mode_x, int mode_y, double *u, int *status, unsigned long long *vmax2, int rz,
                                 double *upriv) {
  if (upriv) u = upriv + (int64_t)(blockIdx.x % PRIV_COPIES) * ((int64_t)g.nx * g.ny);   // private copy of this block
  const int64_t n = cnt[CNT_NSLOTS];
  const int64_t first = first_from_begin ? cnt[CNT_BEGIN] : 0;
  double vm2 = 0.0;
  const double c1 = __dmul_rn(__dmul_rn(0.5, dt), qm);
  for (int64_t p0 = first + blockIdx.x * (int64_t)blockDim.x; p0 < n;
       p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    bool dead_now = false;
    if (p < n) {
      double px = x[p], py = y[p];
      if (!is_dead(px)) {
        double ex, ey;
        if (!gather_E(E2, g, px, py, ex, ey)) atomicOr(status, ISKB_ST_OOB);
        double nvx = push_v(vx[p], ex, c1, qm, dt);
        const double nvy = push_v(vy[p], ey, c1, qm, dt);
        double nvz = push_v(vz[p], 0.0, c1, qm, dt);
        px = push_x(px, nvx, dt);
        py = push_x(py, nvy, dt);
        if (rz) to_cylindrical(px, nvx, nvz, dt);
        vm2 = fmax(vm2, (nvx * nvx + nvy * nvy) + nvz * nvz);
        bool dead = (mode_x == ISKB_BND_DISCARD) && boundary_axis(px, g.ox, g.Lx, mode_x);
        if (!dead) dead = (mode_y == ISKB_BND_DISCARD) && boundary_axis(py, g.oy, g.Ly, mode_y);
        if (!dead) {
          if (mode_x == ISKB_BND_WRAP) boundary_axis(px, g.ox, g.Lx, mode_x);
          if (mode_y == ISKB_BND_WRAP) boundary_axis(py, g.oy, g.Ly, mode_y);
        }
        vx[p] = nvx;
        vy[p] = nvy;
        vz[p] = nvz;
        y[p] = py;
        if (dead) {
          x[p] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
        } else {
          x[p] = px;
          if (u) {
            int i, j;
            double hx, hy;
            cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
            cell1(py, g.dy, g.rdy, g.fast_div, j, hy);
            if (!cell_in_grid(i, j, g.nx, g.ny)) {
              atomicOr(status, ISKB_ST_OOB);
            } else {
              const CicW w = cic_weights(hx, hy);
              const double q = wg[p];
              const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
              atomicAdd(&u[n00], __dmul_rn(w.w00, q));
              atomicAdd(&u[n00 + 1], __dmul_rn(w.w10, q));
              atomicAdd(&u[n00 + g.nx], __dmul_rn(w.w01, q));
              atomicAdd(&u[n00 + g.nx + 1], __dmul_rn(w.w11, q));
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
}

// remove!(sp, i)  kinetic.jl:20-27 (one thread: an API-level operation)
__global__ void k_remove_one(double *x, double *y, double *vx, double *vy, double *vz, double *wg, uint32_t *id,
                             int64_t *cnt, int64_t i, double w0) {
  const int64_t l = cnt[CNT_NSLOTS] - 1;
  x[i] = x[l]; y[i] = y[l];                 // :22
  vx[i] = vx[l]; vy[i] = vy[l]; vz[i] = vz[l];   // :23
  wg[i] = wg[l]; wg[l] = w0;                // :24
  const uint32_t t = id[i]; id[i] = id[l]; id[l] = t;   // :25
  cnt[CNT_NSLOTS] = l;                      // :26
  cnt[CNT_BEGIN] = l;
}

__global__ void k_set_np(int64_t *cnt, int64_t np) {
  cnt[CNT_NSLOTS] = np;
  cnt[CNT_NDEAD] = 0;
  cnt[CNT_BEGIN] = np;
}

// remove_particles!(part, dh, matches)  kinetic.jl:39-50: rows whose cell matches are marked dead
__global__ void k_remove_in_cells(double *x, const double *__restrict__ y, int64_t *cnt, GridDev g,
                                  const uint8_t *__restrict__ mask) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x; p0 < n; p0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + threadIdx.x;
    bool dead_now = false;
    if (p < n) {
      const double px = x[p];
      if (!is_dead(px)) {
        int i, j;
        double hx, hy;
        cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
        cell1(y[p], g.dy, g.rdy, g.fast_div, j, hy);
        if ((unsigned)(i - 1) < (unsigned)g.nx && (unsigned)(j - 1) < (unsigned)g.ny &&
            mask[(i - 1) + (int64_t)(j - 1) * g.nx]) {
          x[p] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, dead_now);
    if (m && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&cnt[CNT_NDEAD], (unsigned long long)__popc(m));
  }
}

}  // namespace

// ================================ host side ====================================================
static int32_t need_grid(iskb_species *sp) {
  if (!sp) return iskb_fail(ISKB_E_INVALID, "null species");
  if (!sp->ctx->has_grid) return iskb_fail(ISKB_E_INVALID, "iskb_grid_set must be called first");
  return ISKB_OK;
}

extern "C" int32_t iskb_cell_index(iskb_species *sp, int32_t *i_out, int32_t *j_out, double *hx_out,
                                   double *hy_out) {
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(sp_compact(sp));
  const int64_t n = sp->h_nslots;
  if (n == 0) return ISKB_OK;
  int32_t *di = nullptr, *dj = nullptr;
  double *dhx = nullptr, *dhy = nullptr;
  CU_TRY(cudaMalloc(&di, n * sizeof(int32_t)));
  CU_TRY(cudaMalloc(&dj, n * sizeof(int32_t)));
  if (hx_out) CU_TRY(cudaMalloc(&dhx, n * sizeof(double)));
  if (hy_out) CU_TRY(cudaMalloc(&dhy, n * sizeof(double)));
  k_cell_index<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, c->g, di, dj, dhx, dhy);
  LAUNCH_CHECK(c);
  if (i_out) CU_TRY(cudaMemcpyAsync(i_out, di, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (j_out) CU_TRY(cudaMemcpyAsync(j_out, dj, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (hx_out) CU_TRY(cudaMemcpyAsync(hx_out, dhx, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (hy_out) CU_TRY(cudaMemcpyAsync(hy_out, dhy, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(di); cudaFree(dj); cudaFree(dhx); cudaFree(dhy);
  return ISKB_OK;
}

extern "C" int32_t iskb_gather(iskb_species *sp, double *partE_out) {
  if (sp) ISKB_TRY(fields_join(sp->ctx));
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  if (!partE_out) return iskb_fail(ISKB_E_INVALID, "partE_out is NULL");
  ISKB_TRY(sp_compact(sp));
  const int64_t n = sp->h_nslots;
  if (n == 0) return ISKB_OK;
  double *d = nullptr;
  CU_TRY(cudaMalloc(&d, 3 * n * sizeof(double)));
  k_gather<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, c->g, c->d_E2, d, n, c->d_status);
  LAUNCH_CHECK(c);
  CU_TRY(cudaMemcpyAsync(partE_out, d, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return ctx_check_status(c);
}

extern "C" int32_t iskb_push(iskb_species *sp, const double *partE, double dt) {
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  double *d = nullptr;
  int64_t n = 0;
  if (partE) {
    ISKB_TRY(sp_compact(sp));   // rows of partE follow the compacted device order
    n = sp->h_nslots;
    if (n == 0) return ISKB_OK;
    CU_TRY(cudaMalloc(&d, 3 * n * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(d, partE, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  const double qm = sp->q / sp->m;   // pushers.jl:39
  ISKB_TRY(sp_vmax_unknown(sp));
  sp_touch(sp);
  k_push<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4],
                                              sp->d_cnt, c->g, c->d_E2, d, n, qm, dt, c->d_status, c->pusher_rz);
  LAUNCH_CHECK(c);
  if (d) {
    CU_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(d);
  }
  return ISKB_OK;
}

extern "C" int32_t iskb_boundary(iskb_species *sp, int32_t mode_x, int32_t mode_y, int64_t *n_removed) {
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  int64_t before = 0;
  if (n_removed) {
    ISKB_TRY(sp_sync_counts(sp));
    before = sp->h_nslots - sp->h_ndead;
  }
  sp_touch(sp);
  k_boundary<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, c->g, mode_x, mode_y);
  LAUNCH_CHECK(c);
  if (mode_x == ISKB_BND_DISCARD || mode_y == ISKB_BND_DISCARD) sp->counts_stale = true;
  if (n_removed) {
    ISKB_TRY(sp_sync_counts(sp));
    *n_removed = before - (sp->h_nslots - sp->h_ndead);
  }
  return ISKB_OK;
}

static int32_t deposit_u(iskb_species *sp) {
  iskb_ctx *c = sp->ctx;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  CU_TRY(cudaMemsetAsync(sp->d_u, 0, nn * sizeof(double), c->stream));
  k_deposit_atomic<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[5], sp->d_cnt, c->g,
                                                        sp->d_u, c->d_status);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

extern "C" int32_t iskb_density(iskb_species *sp, double *n_out) {
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  ISKB_TRY(deposit_u(sp));
  int blocks = (int)((nn + TPB - 1) / TPB);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  k_density<<<blocks, TPB, 0, c->stream>>>(sp->d_u, c->d_V, sp->d_n, nn);
  LAUNCH_CHECK(c);
  if (n_out) {
    CU_TRY(cudaMemcpyAsync(n_out, sp->d_n, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return ctx_check_status(c);
  }
  return ISKB_OK;
}

extern "C" int32_t iskb_species_density_download(iskb_species *sp, double *n_out) {
  ISKB_TRY(need_grid(sp));
  iskb_ctx *c = sp->ctx;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  if (sp->n_in_ufix) {   // fused tiled step: the density of the last step still sits in the fixed-point sums
    int blocks = (int)((nn + TPB - 1) / TPB);
    if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
    k_density_fixed<<<blocks, TPB, 0, c->stream>>>(sp->d_ufix, c->d_V, 1.0 / c->fscale, sp->d_n, nn);
    LAUNCH_CHECK(c);
    sp->n_in_ufix = false;
  }
  CU_TRY(cudaMemcpyAsync(n_out, sp->d_n, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}

extern "C" int32_t iskb_rho_zero(iskb_ctx *c) {
  if (c) ISKB_TRY(fields_join(c));
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "no grid");
  c->rho_lazy = false;
  CU_TRY(cudaMemsetAsync(c->d_rho, 0, (int64_t)c->g.nx * c->g.ny * sizeof(double), c->stream));
  return ISKB_OK;
}

extern "C" int32_t iskb_rho_accumulate(iskb_ctx *c, iskb_species *sp) {
  if (c) ISKB_TRY(fields_join(c));
  if (c) ISKB_TRY(rho_materialize(c));   // (rho of a fused step may still sit in the fixed-point sums)
  ISKB_TRY(need_grid(sp));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  int blocks = (int)((nn + TPB - 1) / TPB);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  k_rho_acc<<<blocks, TPB, 0, c->stream>>>(c->d_rho, sp->d_n, sp->q, nn);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

int32_t launch_rho_finalize(iskb_ctx *c, const std::vector<iskb_species *> *list) {
  if (c) ISKB_TRY(fields_join(c));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  const std::vector<iskb_species *> &sp = list ? *list : c->species;
  if (sp.size() > 8) return iskb_fail(ISKB_E_UNSUPPORTED, "more than 8 kinetic species");
  RhoFin f;
  f.ns = (int)sp.size();
  for (int s = 0; s < f.ns; ++s) {
    f.u[s] = sp[s]->d_u;
    f.n[s] = sp[s]->d_n;
    f.q[s] = sp[s]->q;
  }
  int blocks = (int)((nn + TPB - 1) / TPB);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  k_rho_finalize<<<blocks, TPB, 0, c->stream>>>(f, c->d_V, c->d_rho, nn);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

int32_t launch_rho_finalize_fixed(iskb_ctx *c, const std::vector<iskb_species *> &sp) {
  ISKB_TRY(fields_join(c));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  if (sp.size() > 8) return iskb_fail(ISKB_E_UNSUPPORTED, "more than 8 kinetic species");
  if (!c->d_rho_int) CU_TRY(cudaMalloc(&c->d_rho_int, nn * sizeof(long long)));
  RhoFix f;
  f.ns = (int)sp.size();
  for (int s = 0; s < f.ns; ++s) {
    f.u[s] = sp[s]->d_ufix;
    f.n[s] = sp[s]->d_n;
    f.z[s] = (long long)llround(sp[s]->q / c->q0);
    sp[s]->n_in_ufix = true;   // species.n of this step: formed on demand from the sums (until the next step zeroes them)
  }
  int blocks = (int)((nn + TPB - 1) / TPB);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  // rho itself is formed where it is consumed: by the first transform of the FFT solve on its way in, or by
  // rho_materialize for everybody else (download, dense / GEMM solve)
  if (c->n_ranks > 1) {
    k_rho_fixed_local<<<blocks, TPB, 0, c->stream>>>(f, c->d_V, c->q0 / c->fscale, c->d_rho_int, nullptr, nn);
    LAUNCH_CHECK(c);
    ISKB_TRY(comm_allreduce_sum_i64(c, c->d_rho_int, nn));
    c->rho_ns = 1;
    c->rho_u[0] = c->d_rho_int;
    c->rho_z[0] = 1;
  } else {
    c->rho_ns = f.ns;
    for (int s = 0; s < f.ns; ++s) { c->rho_u[s] = f.u[s]; c->rho_z[s] = f.z[s]; }
  }
  c->rho_lazy = true;
  return ISKB_OK;
}

int32_t rho_materialize(iskb_ctx *c) {
  if (!c->rho_lazy) return ISKB_OK;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  RhoFix f;
  f.ns = c->rho_ns;
  for (int s = 0; s < f.ns; ++s) { f.u[s] = c->rho_u[s]; f.n[s] = nullptr; f.z[s] = c->rho_z[s]; }
  int blocks = (int)((nn + TPB - 1) / TPB);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  k_rho_fixed_local<<<blocks, TPB, 0, c->stream>>>(f, c->d_V, c->q0 / c->fscale, nullptr, c->d_rho, nn);
  LAUNCH_CHECK(c);
  c->rho_lazy = false;
  return ISKB_OK;
}

// advance! for one species, simple kernel (see advance_fused.cu for the tiled one)
int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit,
                              bool from_begin) {
  ISKB_TRY(fields_join(sp->ctx));
  iskb_ctx *c = sp->ctx;
  const double qm = sp->q / sp->m;
  int blocks = from_begin ? c->n_sm : grid_for(sp);
  if (!from_begin) ISKB_TRY(sp_vmax_reset(sp));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  const bool priv = deposit && nn <= PRIV_MAX_NODES && blocks >= 2 * PRIV_COPIES;
  if (priv && !c->d_upriv) {
    CU_TRY(cudaMalloc(&c->d_upriv, PRIV_COPIES * nn * sizeof(double)));
    CU_TRY(cudaMemsetAsync(c->d_upriv, 0, PRIV_COPIES * nn * sizeof(double), c->stream));
  }
  if (!from_begin) ISKB_TRY(prof_begin(c));
  sp_touch(sp);
  k_advance_simple<<<blocks, TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4],
                                                  sp->col[5], sp->d_cnt, from_begin ? 1 : 0, c->g, c->d_E2, qm,
                                                  dt, mode_x, mode_y, deposit ? sp->d_u : nullptr, c->d_status, sp->d_vmax2,
                                                  c->pusher_rz, priv ? c->d_upriv : nullptr);
  LAUNCH_CHECK(c);
  if (priv) {
    k_reduce_copies<<<(int)((nn + TPB - 1) / TPB), TPB, 0, c->stream>>>(c->d_upriv, nn, sp->d_u);
    LAUNCH_CHECK(c);
  }
  if (!from_begin) ISKB_TRY(prof_end(c));
  if (mode_x == ISKB_BND_DISCARD || mode_y == ISKB_BND_DISCARD) sp->counts_stale = true;
  return ISKB_OK;
}

// ---- kinetic.jl:20-50 remove! / add! / remove_particles! ------------------------------------------
extern "C" int32_t iskb_species_remove(iskb_species *sp, int64_t i) {
  if (!sp) return iskb_fail(ISKB_E_INVALID, "null species");
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(sp_compact(sp));
  if (i < 1 || i > sp->h_nslots) return iskb_fail(ISKB_E_INVALID, "remove!: row %lld outside 1..np = %lld", (long long)i, (long long)sp->h_nslots);
  sp_touch(sp);
  k_remove_one<<<1, 1, 0, c->stream>>>(sp->col[0], sp->col[1], sp->col[2], sp->col[3], sp->col[4], sp->col[5], sp->id, sp->d_cnt,
                                       i - 1, sp->w0);
  LAUNCH_CHECK(c);
  sp->h_nslots -= 1;
  sp->h_nsorted = 0;
  return ISKB_OK;
}

extern "C" int32_t iskb_species_add(iskb_species *src, iskb_species *dst) {
  if (!src || !dst || src->ctx != dst->ctx || src == dst) return iskb_fail(ISKB_E_INVALID, "bad species");
  iskb_ctx *c = dst->ctx;
  ISKB_TRY(sp_compact(src));
  ISKB_TRY(sp_compact(dst));
  const int64_t n = src->h_nslots, at = dst->h_nslots;
  if (n == 0) return ISKB_OK;                       // :30
  if (at + n > dst->cap) return iskb_fail(ISKB_E_CAPACITY, "add!: %lld + %lld rows exceed the capacity %lld", (long long)at, (long long)n, (long long)dst->cap);
  for (int k = 0; k < 5; ++k)                       // :32-33 (wg and id of dst stay)
    CU_TRY(cudaMemcpyAsync(dst->col[k] + at, src->col[k], n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  const int64_t newn = at + n;                      // :35
  k_set_np<<<1, 1, 0, c->stream>>>(dst->d_cnt, newn);
  LAUNCH_CHECK(c);
  dst->h_nslots = newn; dst->h_ndead = 0; dst->counts_stale = false;
  dst->h_nsorted = dst->h_nsorted < at ? dst->h_nsorted : at;
  ISKB_TRY(sp_vmax_unknown(dst));
  return ISKB_OK;
}

extern "C" int32_t iskb_species_remove_in_cells(iskb_species *sp, const uint8_t *cell_mask, int64_t *n_removed) {
  ISKB_TRY(need_grid(sp));
  if (!cell_mask) return iskb_fail(ISKB_E_INVALID, "cell_mask is NULL");
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(sp_sync_counts(sp));
  const int64_t before = sp->h_nslots - sp->h_ndead;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  uint8_t *d = nullptr;
  CU_TRY(cudaMalloc(&d, nn));
  CU_TRY(cudaMemcpyAsync(d, cell_mask, nn, cudaMemcpyHostToDevice, c->stream));
  sp_touch(sp);
  k_remove_in_cells<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, c->g, d);
  LAUNCH_CHECK(c);
  sp->counts_stale = true;
  ISKB_TRY(sp_sync_counts(sp));
  cudaFree(d);
  if (n_removed) *n_removed = before - (sp->h_nslots - sp->h_ndead);
  return ISKB_OK;
}

// ---- axisymmetric variant -----------------------------------------------------
extern "C" int32_t iskb_set_pusher(iskb_ctx *c, int32_t kind) {
  if (!c || (kind != ISKB_PUSHER_XY && kind != ISKB_PUSHER_RZ)) return iskb_fail(ISKB_E_INVALID, "bad pusher kind");
  c->pusher_rz = kind == ISKB_PUSHER_RZ;
  return ISKB_OK;
}

extern "C" int32_t iskb_transform_cylindrical(iskb_species *sp, double dt) {
  if (!sp) return iskb_fail(ISKB_E_INVALID, "null species");
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(sp_vmax_unknown(sp));
  sp_touch(sp);
  k_to_cylindrical<<<grid_for(sp), TPB, 0, c->stream>>>(sp->col[0], sp->col[2], sp->col[4], sp->d_cnt, dt);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

extern "C" int32_t iskb_cell_volume_set(iskb_ctx *c, const double *V) {
  if (!c || !c->has_grid || !V) return iskb_fail(ISKB_E_INVALID, "no grid / V");
  CU_TRY(cudaMemcpyAsync(c->d_V, V, (int64_t)c->g.nx * c->g.ny * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}
