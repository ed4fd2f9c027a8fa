// DSMC between two kinetic species: Chemistry/src/dsmc.jl:25-142.
//   cache!          :25-30    per-cell lists of rows          -> k_dsmc_count / k_dsmc_scan / k_dsmc_fill (counting sort
//                                                                of row indices by cell; order inside a cell is free, the
//                                                                reference only ever draws from the list at random)
//   PIC.perform!    :87-142   candidate pairs per cell        -> k_dsmc_collide, ONE THREAD PER CELL: the pairs of a cell are
//   perform!(ElasticCollision) :32-79                            processed in sequence like the reference (a later pair sees the
//                                                                velocities an earlier one left behind); cells are independent
// RNG: Philox4x32-10, counter = (cell, call, draw); the reference's MersenneTwister stream cannot be reproduced, so parity is
// statistical, plus the stream-independent invariants (candidate-pair count and its carry exact; momentum and energy conserved).
// Reference quirks kept (oracle/dsmc_oracle.py D1-D5): one collision per DSMC object; sigma_g_max = max(sigma) * argmax(sigma);
// the unequal-weight branch writes target.v[s] with mr2; `sR == tR` compares rows of two different species.
#include <cmath>
#include <cstring>

#include "pic_device.cuh"

struct iskb_dsmc {
  iskb_ctx *ctx = nullptr;
  iskb_species *source = nullptr, *target = nullptr;
  int n_nodes = 0;
  double *d_gn = nullptr, *d_sg = nullptr;   // sigma(g) table
  double sgmax = 0.0;
  uint64_t seed = 0, calls = 0;
  double *d_rem = nullptr;                   // collisions_remaining  dsmc.jl:18,118
  uint32_t *d_count[2] = {nullptr, nullptr}, *d_start[2] = {nullptr, nullptr}, *d_cursor = nullptr;
  uint32_t *d_list[2] = {nullptr, nullptr};
  float *d_nu = nullptr;
  unsigned long long *d_stats = nullptr;     // [0] candidate pairs, [1] collisions
  int64_t totals[2] = {0, 0};
};

struct DsmcCell {
  // everything one cell needs; host pointers in the debug hook, device pointers in the kernel
  double *svx, *svy, *svz, *tvx, *tvy, *tvz;
  const double *swg, *twg;
  const uint32_t *slist, *tlist;   // rows of this cell
  uint32_t Na, Nb;
  int64_t t_rows;                  // rows of the target species (bound for the D2 write)
  double ms, mt, Wa, Wb, dx, dy, dt, sgmax;
  const double *gn, *sg;
  int n_nodes;
  int same;                        // source === target
  uint32_t k0, k1, call;
  uint64_t cell;
};

namespace {

constexpr int TPB = 256;

// CrossSection(g): piecewise linear, Flat() outside  (cross_section.jl:8-14)
__host__ __device__ inline double xsec(const double *xs, const double *ys, int n, double x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xs[mid] <= x) lo = mid; else hi = mid;
  }
  const double f = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
  return (1.0 - f) * ys[lo] + f * ys[lo + 1];
}

struct CellRng {
  uint32_t c0, c1, call, k0, k1, draw;
  uint32_t buf[4];
  int have;
  __host__ __device__ CellRng(uint64_t cell, uint32_t call_, uint32_t k0_, uint32_t k1_)
      : c0((uint32_t)cell), c1((uint32_t)(cell >> 32)), call(call_), k0(k0_), k1(k1_), draw(0), have(0) {}
  __host__ __device__ double u01() {
    if (have < 2) {
      const Philox4 o = philox4x32_10(c0, c1, call, draw++, k0, k1);
      buf[0] = o.c[0]; buf[1] = o.c[1]; buf[2] = o.c[2]; buf[3] = o.c[3];
      have = 4;
    }
    have -= 2;
    return u01_53(buf[have], buf[have + 1]);
  }
};

// The per-cell body of PIC.perform!(dsmc, ...)  dsmc.jl:109-139.  Returns the number of candidate pairs; *n_coll the collisions.
__host__ __device__ inline uint32_t dsmc_cell(const DsmcCell &c, double *remaining, uint32_t *n_coll, int *oob_write) {
  *n_coll = 0;
  if (c.Na < 2 || c.Nb < 2) return 0;                                 // :111-113
  double Pab, Pba;
  if (c.Wa > c.Wb) { Pab = c.Wb / c.Wa; Pba = 1.0; } else { Pab = 1.0; Pba = c.Wa / c.Wb; }   // :101-105
  const double na = (double)c.Na * c.Wa / (c.dx * c.dy);             // :114
  double Nc = na * (double)c.Nb * c.dt * c.sgmax;                     // :115
  Nc /= Pab + (c.Wb / c.Wa) * Pba;                                    // :116
  if (!c.same) Nc *= 2;                                               // :117-119
  Nc += *remaining;                                                   // :121
  const double fl = floor(Nc);
  *remaining = Nc - fl;                                               // :122
  const uint32_t npairs = fl < 4.0e9 ? (uint32_t)fl : 4000000000u;
  const double mr1 = c.ms / (c.ms + c.mt), mr2 = c.mt / (c.ms + c.mt);   // :35-36
  CellRng rng(c.cell, c.call, c.k0, c.k1);
  for (uint32_t it = 0; it < npairs; ++it) {                          // :124
    uint32_t a = (uint32_t)(rng.u01() * c.Na), b = (uint32_t)(rng.u01() * c.Nb);
    if (a >= c.Na) a = c.Na - 1;
    if (b >= c.Nb) b = c.Nb - 1;
    const uint32_t s = c.slist[a];                                    // :125 rand(candidates)
    uint32_t t = c.tlist[b];
    while (!c.same && s == t) {                                       // :127-129 (D3)
      b = (uint32_t)(rng.u01() * c.Nb);
      if (b >= c.Nb) b = c.Nb - 1;
      t = c.tlist[b];
    }
    const double gx = c.svx[s] - c.tvx[t], gy = c.svy[s] - c.tvy[t], gz = c.svz[s] - c.tvz[t];
    const double g = sqrt(gx * gx + gy * gy + gz * gz);               // :131
    const double sgg = xsec(c.gn, c.sg, c.n_nodes, g) * g;            // :132
    const double P = sgg / c.sgmax;
    const double R = rng.u01();
    if (P < R) continue;                                              // :134-138
    // perform!(collision::ElasticCollision, s, t)  :32-79 with vss_inv == 1
    const double cmx = mr1 * c.svx[s] + mr2 * c.tvx[t], cmy = mr1 * c.svy[s] + mr2 * c.tvy[t], cmz = mr1 * c.svz[s] + mr2 * c.tvz[t];
    const double B = 2 * rng.u01() - 1.0;
    const double A = sqrt(1 - B * B);
    const double C = 2 * 3.141592653589793 * rng.u01();
    const double rx = g * B, ry = g * (A * cos(C)), rz = g * (A * sin(C));
    if (c.swg[s] == c.twg[t]) {                                       // :61-63
      c.svx[s] = cmx + mr2 * rx; c.svy[s] = cmy + mr2 * ry; c.svz[s] = cmz + mr2 * rz;
      c.tvx[t] = cmx - mr1 * rx; c.tvy[t] = cmy - mr1 * ry; c.tvz[t] = cmz - mr1 * rz;
    } else {                                                          // :64-78
      const double Pab2 = c.twg[t] / c.swg[s], Pba2 = c.swg[s] / c.twg[t];
      const double R2 = rng.u01();
      if (Pab2 > R2) { c.svx[s] = cmx + mr2 * rx; c.svy[s] = cmy + mr2 * ry; c.svz[s] = cmz + mr2 * rz; }
      if (Pba2 > R2) {                                                // D2: target.v[s,:] = vc_cm .- mr2*vr_cp, as written
        if ((int64_t)s < c.t_rows) { c.tvx[s] = cmx - mr2 * rx; c.tvy[s] = cmy - mr2 * ry; c.tvz[s] = cmz - mr2 * rz; }
        else *oob_write = 1;                                          // the reference would raise BoundsError
      }
    }
    ++*n_coll;
  }
  return npairs;
}

// cache!: cell of every live row (particle_cell, ParticleInCell.jl:28-35); candidates is nx x ny (:94-95)
__global__ void k_dsmc_count(const double *__restrict__ x, const double *__restrict__ y, const int64_t *__restrict__ cnt, GridDev g,
                             uint32_t *count, uint32_t *rowcell, uint32_t *rowrank, int *status) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const double px = x[p];
    uint32_t cell = 0xffffffffu;
    if (!is_dead(px)) {
      int i, j;
      double hx, hy;
      cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
      cell1(y[p], g.dy, g.rdy, g.fast_div, j, hy);
      if ((unsigned)(i - 1) < (unsigned)g.nx && (unsigned)(j - 1) < (unsigned)g.ny) {
        cell = (uint32_t)((i - 1) + (j - 1) * g.nx);
        rowrank[p] = atomicAdd(&count[cell], 1u);   // the slot of the row in its cell's list: k_dsmc_fill needs no second atomic
      } else {
        atomicOr(status, ISKB_ST_OOB);
      }
    }
    rowcell[p] = cell;
  }
}

// (the exclusive scan of the counts is sort.cu's multi-block exclusive_scan_u32: the single-block scan that stood here took
// 0.89 ms for 1025^2 cells)
__global__ void k_dsmc_fill(const uint32_t *__restrict__ rowcell, const uint32_t *__restrict__ rowrank, const int64_t *__restrict__ cnt,
                            const uint32_t *__restrict__ start, uint32_t *list) {
  const int64_t n = cnt[CNT_NSLOTS];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = rowcell[p];
    if (c != 0xffffffffu) list[start[c] + rowrank[p]] = (uint32_t)p;
  }
}

struct DsmcDev {
  DsmcCell proto;                 // species-wide fields filled in; lists / counts / cell set per thread
  const uint32_t *count_s, *count_t, *start_s, *start_t, *list_s, *list_t;
  double *rem;
  float *nu;
  unsigned long long *stats;
  int *status;
  int64_t nn;
};

__global__ void k_dsmc_collide(DsmcDev d) {
  unsigned long long cand = 0, coll = 0;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < d.nn; cell += (int64_t)gridDim.x * blockDim.x) {
    DsmcCell c = d.proto;
    c.cell = (uint64_t)cell;
    c.Na = d.count_s[cell];
    c.Nb = d.count_t[cell];
    c.slist = d.list_s + d.start_s[cell];
    c.tlist = d.list_t + d.start_t[cell];
    uint32_t ncoll = 0;
    int oob = 0;
    double rem = d.rem[cell];
    const uint32_t np = dsmc_cell(c, &rem, &ncoll, &oob);
    d.rem[cell] = rem;
    if (d.nu) d.nu[cell] = (float)ncoll;
    if (oob) atomicOr(d.status, ISKB_ST_OOB);
    cand += np;
    coll += ncoll;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    cand += __shfl_xor_sync(0xffffffffu, cand, k);
    coll += __shfl_xor_sync(0xffffffffu, coll, k);
  }
  if ((threadIdx.x & 31) == 0 && cand) {
    atomicAdd(&d.stats[0], cand);
    atomicAdd(&d.stats[1], coll);
  }
}

__global__ void k_nu_to_double(const float *__restrict__ nu, int64_t nn, double *out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nn; k += (int64_t)gridDim.x * blockDim.x) out[k] = nu[k];
}

int blocks_for(const iskb_ctx *c, int64_t n) {
  int64_t b = (n + TPB - 1) / TPB;
  if (b > (int64_t)c->n_sm * 8) b = (int64_t)c->n_sm * 8;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

// Test hook: the per-cell body on HOST arrays (the same __host__ __device__ code the kernel runs), so that the
// stream-independent invariants can be checked without a GPU.  list_s / list_t: rows of the cell.
extern "C" int32_t iskb_debug_dsmc_cell(double *sv /* 3 x ns col-major */, int64_t ns, double *tv, int64_t nt, const double *swg,
                                        const double *twg, const uint32_t *list_s, uint32_t Na, const uint32_t *list_t, uint32_t Nb,
                                        double ms, double mt, double Wa, double Wb, double dx, double dy, double dt,
                                        const double *gn, const double *sg, int32_t n_nodes, int32_t same, uint64_t seed,
                                        uint32_t call, uint64_t cell, double *remaining, uint32_t *n_pairs, uint32_t *n_coll) {
  DsmcCell c;
  memset(&c, 0, sizeof(c));
  c.svx = sv; c.svy = sv + ns; c.svz = sv + 2 * ns;
  c.tvx = tv; c.tvy = tv + nt; c.tvz = tv + 2 * nt;
  c.swg = swg; c.twg = twg; c.slist = list_s; c.tlist = list_t; c.Na = Na; c.Nb = Nb; c.t_rows = nt;
  c.ms = ms; c.mt = mt; c.Wa = Wa; c.Wb = Wb; c.dx = dx; c.dy = dy; c.dt = dt;
  c.gn = gn; c.sg = sg; c.n_nodes = n_nodes; c.same = same;
  int kmax = 0;
  for (int k = 1; k < n_nodes; ++k) if (sg[k] > sg[kmax]) kmax = k;
  c.sgmax = sg[kmax] * gn[kmax];                     // maximum(rate) * argmax(rate)  dsmc.jl:107
  c.k0 = (uint32_t)seed; c.k1 = (uint32_t)(seed >> 32); c.call = call; c.cell = cell;
  int oob = 0;
  *n_pairs = dsmc_cell(c, remaining, n_coll, &oob);
  return oob ? ISKB_E_OOB : ISKB_OK;
}

extern "C" int32_t iskb_dsmc_create(iskb_ctx *c, iskb_species *source, iskb_species *target, const double *g_nodes, const double *sigma,
                                    int32_t n_nodes, uint64_t seed, iskb_dsmc **out) {
  if (!c || !c->has_grid || !source || !target || !g_nodes || !sigma || n_nodes < 2 || !out)
    return iskb_fail(ISKB_E_INVALID, "iskb_dsmc_create: bad arguments (grid set first, table of >= 2 rows)");
  if (source->ctx != c || target->ctx != c) return iskb_fail(ISKB_E_INVALID, "species of another context");
  if ((int64_t)c->g.nx * c->g.ny >= 0xffffffffll) return iskb_fail(ISKB_E_UNSUPPORTED, "grid too large for 32-bit cell ids");
  iskb_dsmc *d = new iskb_dsmc();
  d->ctx = c; d->source = source; d->target = target; d->n_nodes = n_nodes; d->seed = seed;
  int kmax = 0;
  for (int k = 1; k < n_nodes; ++k) if (sigma[k] > sigma[kmax]) kmax = k;
  d->sgmax = sigma[kmax] * g_nodes[kmax];            // dsmc.jl:107
  if (!(d->sgmax > 0.0)) { delete d; return iskb_fail(ISKB_E_INVALID, "sigma_g_max must be positive"); }
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  CU_TRY(cudaMalloc(&d->d_gn, n_nodes * sizeof(double)));
  CU_TRY(cudaMalloc(&d->d_sg, n_nodes * sizeof(double)));
  CU_TRY(cudaMemcpyAsync(d->d_gn, g_nodes, n_nodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(d->d_sg, sigma, n_nodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMalloc(&d->d_rem, nn * sizeof(double)));
  CU_TRY(cudaMemsetAsync(d->d_rem, 0, nn * sizeof(double), c->stream));       // :90-92
  const int nsp = source == target ? 1 : 2;
  for (int k = 0; k < nsp; ++k) {
    iskb_species *sp = k == 0 ? source : target;
    CU_TRY(cudaMalloc(&d->d_count[k], nn * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&d->d_start[k], (nn + 1) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&d->d_list[k], 3 * sp->cap * sizeof(uint32_t)));         // list + per-row cell and rank scratch
  }
  CU_TRY(cudaMalloc(&d->d_cursor, nn * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&d->d_nu, nn * sizeof(float)));
  CU_TRY(cudaMalloc(&d->d_stats, 2 * sizeof(unsigned long long)));
  CU_TRY(cudaMemsetAsync(d->d_stats, 0, 2 * sizeof(unsigned long long), c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->dsmcs.push_back(d);
  *out = d;
  return ISKB_OK;
}

int32_t dsmc_free(iskb_dsmc *d) {
  if (!d) return ISKB_OK;
  cudaFree(d->d_gn); cudaFree(d->d_sg); cudaFree(d->d_rem); cudaFree(d->d_cursor); cudaFree(d->d_nu); cudaFree(d->d_stats);
  for (int k = 0; k < 2; ++k) { cudaFree(d->d_count[k]); cudaFree(d->d_start[k]); cudaFree(d->d_list[k]); }
  delete d;
  return ISKB_OK;
}

int32_t dsmc_launch(iskb_dsmc *d, double dt, bool want_nu) {
  iskb_ctx *c = d->ctx;
  // Candidate pairs per cell go with N_a * N_b of that cell (dsmc.jl:109-122).  Ranks own index slices of every
  // species, so each would see only N/R of both lists and the collision frequency would drop by ~R.
  if (c->n_ranks > 1)
    return iskb_fail(ISKB_E_UNSUPPORTED, "DSMC with particles sharded over %d ranks: per-cell pair counts need all rows of a cell on one rank", c->n_ranks);
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  const bool same = d->source == d->target;
  for (int k = 0; k < (same ? 1 : 2); ++k) {          // cache!  :98-99
    iskb_species *sp = k == 0 ? d->source : d->target;
    uint32_t *rowcell = d->d_list[k] + sp->cap, *rowrank = d->d_list[k] + 2 * sp->cap;
    const int64_t bound = sp->counts_stale ? sp->cap : sp->h_nslots;
    CU_TRY(cudaMemsetAsync(d->d_count[k], 0, nn * sizeof(uint32_t), c->stream));
    k_dsmc_count<<<blocks_for(c, bound), TPB, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, c->g, d->d_count[k], rowcell, rowrank,
                                                              c->d_status);
    LAUNCH_CHECK(c);
    CU_TRY(cudaMemcpyAsync(d->d_start[k], d->d_count[k], nn * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    CU_TRY(cudaMemsetAsync(d->d_start[k] + nn, 0, sizeof(uint32_t), c->stream));
    ISKB_TRY(exclusive_scan_u32(c, d->d_start[k], nn + 1, d->d_cursor));   // d_cursor: scratch of the scan's partial sums
    k_dsmc_fill<<<blocks_for(c, bound), TPB, 0, c->stream>>>(rowcell, rowrank, sp->d_cnt, d->d_start[k], d->d_list[k]);
    LAUNCH_CHECK(c);
  }
  DsmcDev v;
  memset(&v, 0, sizeof(v));
  iskb_species *s = d->source, *t = d->target;
  v.proto.svx = s->col[2]; v.proto.svy = s->col[3]; v.proto.svz = s->col[4];
  v.proto.tvx = t->col[2]; v.proto.tvy = t->col[3]; v.proto.tvz = t->col[4];
  v.proto.swg = s->col[5]; v.proto.twg = t->col[5];
  v.proto.t_rows = t->counts_stale ? t->cap : t->h_nslots;
  v.proto.ms = s->m; v.proto.mt = t->m; v.proto.Wa = s->w0; v.proto.Wb = t->w0;
  v.proto.dx = c->g.dx; v.proto.dy = c->g.dy; v.proto.dt = dt; v.proto.sgmax = d->sgmax;
  v.proto.gn = d->d_gn; v.proto.sg = d->d_sg; v.proto.n_nodes = d->n_nodes; v.proto.same = same ? 1 : 0;
  v.proto.k0 = (uint32_t)d->seed;
  v.proto.k1 = (uint32_t)(d->seed >> 32) ^ (0x9E3779B9u * (uint32_t)(c->rank + 1));
  v.proto.call = (uint32_t)(d->calls++);
  const int ti = same ? 0 : 1;
  v.count_s = d->d_count[0]; v.count_t = d->d_count[ti];
  v.start_s = d->d_start[0]; v.start_t = d->d_start[ti];
  v.list_s = d->d_list[0]; v.list_t = d->d_list[ti];
  v.rem = d->d_rem; v.nu = want_nu ? d->d_nu : nullptr; v.stats = d->d_stats; v.status = c->d_status; v.nn = nn;
  ISKB_TRY(sp_vmax_unknown(s));
  if (!same) ISKB_TRY(sp_vmax_unknown(t));
  k_dsmc_collide<<<blocks_for(c, nn), TPB, 0, c->stream>>>(v);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

extern "C" int32_t iskb_dsmc_perform(iskb_dsmc *d, double dt, double *nu_out, int64_t *n_candidates, int64_t *n_collisions) {
  if (!d) return iskb_fail(ISKB_E_INVALID, "null dsmc");
  iskb_ctx *c = d->ctx;
  CU_TRY(cudaMemsetAsync(d->d_stats, 0, 2 * sizeof(unsigned long long), c->stream));
  ISKB_TRY(dsmc_launch(d, dt, nu_out != nullptr));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  double *tmp = nullptr;
  if (nu_out) {
    CU_TRY(cudaMalloc(&tmp, nn * sizeof(double)));
    k_nu_to_double<<<blocks_for(c, nn), TPB, 0, c->stream>>>(d->d_nu, nn, tmp);
    LAUNCH_CHECK(c);
    CU_TRY(cudaMemcpyAsync(nu_out, tmp, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(cudaMemcpyAsync(c->h_scratch, d->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(tmp);
  if (n_candidates) *n_candidates = c->h_scratch[0];
  if (n_collisions) *n_collisions = c->h_scratch[1];
  d->totals[0] += c->h_scratch[0];
  d->totals[1] += c->h_scratch[1];
  return ctx_check_status(c);
}
