// C-ABI entry points: context, grid, fields, species storage, and the fused step.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tiles.cuh"

int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit, bool from_begin);
int32_t launch_advance_tiled(iskb_species *sp, double dt, int mode_x, int mode_y);

// ---- errors -------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void iskb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int32_t iskb_fail(int32_t code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char *iskb_last_error(void) { return g_err; }
extern "C" int32_t iskb_version(void) { return 100; }
// test hook: the Philox4x32-10 block function on the host (known-answer tests run without a GPU)
extern "C" void iskb_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Philox4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int k = 0; k < 4; ++k) out[k] = o.c[k];
}

// ---- small kernels ------------------------------------------------------------------------------
namespace {
constexpr int TPB = 256;

__global__ void k_cell_volume(int nx, int ny, double dx, double dy, int b0, int b1, int b2, int b3, double *V) {
  const int64_t nn = (int64_t)nx * ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(n % nx), j = (int)(n / nx);
    double v = 0.0 + dx * dy;                         // RegularGrids.jl:29-30
    if (b0 != ISKB_BC_PERIODIC && i == 0) v *= 0.5;   // :33
    if (b1 != ISKB_BC_PERIODIC && i == nx - 1) v *= 0.5;
    if (b2 != ISKB_BC_PERIODIC && j == 0) v *= 0.5;
    if (b3 != ISKB_BC_PERIODIC && j == ny - 1) v *= 0.5;
    V[n] = v;
  }
}

__global__ void k_init_species(double *wg, uint32_t *id, int64_t cap, double w0) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < cap;
       p += (int64_t)gridDim.x * blockDim.x) {
    wg[p] = w0;                 // ones(N) * weight, configuration.jl:99
    id[p] = (uint32_t)(p + 1);  // particle_uuids, kinetic.jl:15
  }
}

__global__ void k_E_interleave(const double *__restrict__ E, int64_t nn, double2 *E2) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x)
    E2[n] = make_double2(E[n], E[n + nn]);
}
__global__ void k_E_split(const double2 *__restrict__ E2, int64_t nn, double *E) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x) {
    const double2 e = E2[n];
    E[n] = e.x;
    E[n + nn] = e.y;
    E[n + 2 * nn] = 0.0;
  }
}

// sample!(MaxwellianSource)  sources.jl:31-32 : x = rand*wx + dx ; v = randn*wv + dv
__global__ void k_sample(double *x, double *y, double *vx, double *vy, double *vz, int64_t first, int64_t n,
                         double wx0, double wx1, double dx0, double dx1, double wv0, double wv1, double wv2,
                         double dv0, double dv1, double dv2, uint32_t k0, uint32_t k1, uint32_t call) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    const Philox4 a = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), call, 0u, k0, k1);
    const Philox4 b = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), call, 1u, k0, k1);
    const Philox4 c = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), call, 2u, k0, k1);
    const int64_t p = first + t;
    x[p] = __dadd_rn(__dmul_rn(u01_53(a.c[0], a.c[1]), wx0), dx0);
    y[p] = __dadd_rn(__dmul_rn(u01_53(a.c[2], a.c[3]), wx1), dx1);
    const double r0 = sqrt(-2.0 * log(u01_open(b.c[0], b.c[1])));
    const double r1 = sqrt(-2.0 * log(u01_open(c.c[0], c.c[1])));
    double s0, c0, s1, c1;
    sincospi(2.0 * u01_53(b.c[2], b.c[3]), &s0, &c0);
    sincospi(2.0 * u01_53(c.c[2], c.c[3]), &s1, &c1);
    vx[p] = __dadd_rn(__dmul_rn(r0 * c0, wv0), dv0);
    vy[p] = __dadd_rn(__dmul_rn(r0 * s0, wv1), dv1);
    vz[p] = __dadd_rn(__dmul_rn(r1 * c1, wv2), dv2);
  }
}

__global__ void k_fill3(double *vx, double *vy, double *vz, int64_t n, double a, double b, double c) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    vx[p] = a; vy[p] = b; vz[p] = c;
  }
}

__global__ void k_set_counts(int64_t *cnt, int64_t nslots, int64_t ndead) {
  cnt[CNT_NSLOTS] = nslots;
  cnt[CNT_NDEAD] = ndead;
  cnt[CNT_BEGIN] = nslots;
}

int blocks_for(const iskb_ctx *c, int64_t n) {
  int64_t b = (n + TPB - 1) / TPB;
  if (b > (int64_t)c->n_sm * 8) b = (int64_t)c->n_sm * 8;
  return b < 1 ? 1 : (int)b;
}
}  // namespace

// ---- lifecycle ----------------------------------------------------------------------------------
extern "C" int32_t iskb_create(int32_t device, iskb_ctx **out) {
  if (!out) return iskb_fail(ISKB_E_INVALID, "out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return iskb_fail(ISKB_E_CUDA, "no CUDA device available (%s); iskra_b200 has no CPU fallback",
                     cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return iskb_fail(ISKB_E_INVALID, "device %d out of range (%d devices)", device, ndev);
  CU_TRY(cudaSetDevice(device));
  iskb_ctx *c = new iskb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  c->n_sm = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  {
    // high priority: the solve is a chain of small latency-bound kernels; its blocks should get SM slots
    // ahead of the bandwidth-bound particle kernels it overlaps with
    int lo = 0, hi = 0;
    CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU_TRY(cudaStreamCreateWithPriority(&c->fstream, cudaStreamNonBlocking, hi));
  }
  CU_TRY(cudaStreamCreateWithFlags(&c->mstream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->pstream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_p0, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_m0, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_m1, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_rho, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_E, cudaEventDisableTiming));
  CU_TRY(cudaMalloc(&c->d_status, sizeof(int)));
  CU_TRY(cudaMemset(c->d_status, 0, sizeof(int)));
  CU_TRY(cudaMallocHost(&c->h_status, sizeof(int)));
  CU_TRY(cudaMallocHost(&c->h_scratch, 16 * sizeof(int64_t)));
  *out = c;
  return ISKB_OK;
}

static void free_species(iskb_species *s) {
  for (int q = 0; q < 6; ++q) { cudaFree(s->col[q]); cudaFree(s->alt[q]); }
  cudaFree(s->id); cudaFree(s->alt_id); cudaFree(s->d_cnt); cudaFree(s->d_vmax2); cudaFree(s->d_vz2max); cudaFree(s->d_u); cudaFree(s->d_n);
  for (int k = 0; k < 2; ++k) { cudaFree(s->d_key[k]); cudaFree(s->d_idx[k]); }
  if (s->h_wstats) cudaFreeHost(s->h_wstats);
  for (int k = 0; k < 2; ++k) if (s->ev_wstats[k]) cudaEventDestroy(s->ev_wstats[k]);
  cudaFree(s->d_hist); cudaFree(s->d_trk_list); cudaFree(s->d_trk_n);
  tdir_free(s);
  delete s;
}

extern "C" int32_t iskb_destroy(iskb_ctx *c) {
  if (!c) return ISKB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->fstream);
  cudaStreamSynchronize(c->pstream);
  cudaStreamSynchronize(c->mstream);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  for (iskb_mcc *m : c->mccs) {
    if (m->ev_pre) cudaEventDestroy(m->ev_pre);
    cudaFree(m->d_tn); cudaFree(m->d_eps); cudaFree(m->d_sig); cudaFree(m->d_stats); cudaFree(m->d_nu); cudaFree(m->d_cand); cudaFree(m->d_coll); cudaFree(m->d_lists_cnt); cudaFree(m->d_pk);
    delete m;
  }
  for (iskb_species *s : c->species) free_species(s);
  for (iskb_dsmc *d : c->dsmcs) dsmc_free(d);
  tracker_free(c->tracker);
  poisson_free(c);
  cudaFree(c->d_rho_int); cudaFree(c->d_see_counts);
  cudaFree(c->d_upriv); cudaFree(c->d_V); cudaFree(c->d_rho); cudaFree(c->d_phi); cudaFree(c->d_E2); cudaFree(c->d_status);
  cudaFreeHost(c->h_status); cudaFreeHost(c->h_scratch);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_m0); cudaEventDestroy(c->ev_m1); cudaStreamDestroy(c->mstream);
  cudaEventDestroy(c->ev_p0); cudaStreamDestroy(c->pstream);
  cudaEventDestroy(c->ev_rho);
  cudaEventDestroy(c->ev_E);
  cudaStreamDestroy(c->fstream);
  cudaStreamDestroy(c->own_stream);
  delete c;
  return ISKB_OK;
}

extern "C" int32_t iskb_set_stream(iskb_ctx *c, void *s) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  CU_TRY(cudaStreamSynchronize(c->fstream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->fields_pending = false;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return ISKB_OK;
}

int32_t fields_join(iskb_ctx *c) {
  if (c->fields_pending) {
    CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_E, 0));
    c->fields_pending = false;
  }
  return ISKB_OK;
}

extern "C" int32_t iskb_stream_join(iskb_ctx *c) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  return fields_join(c);
}

extern "C" int32_t iskb_synchronize(iskb_ctx *c) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  ISKB_TRY(fields_join(c));
  CU_TRY(cudaStreamSynchronize(c->pstream));   // test phases of the next step's MCC that ran ahead
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ctx_check_status(c);
}

extern "C" int32_t iskb_warning_too_fast(iskb_ctx *c, int32_t *out) {
  if (!c || !out) return iskb_fail(ISKB_E_INVALID, "null");
  *out = c->warn_too_fast ? 1 : 0;
  c->warn_too_fast = false;
  return ISKB_OK;
}

extern "C" int32_t iskb_launch_count(iskb_ctx *c, int64_t *out) {
  if (!c || !out) return iskb_fail(ISKB_E_INVALID, "null");
  *out = c->launches;
  return ISKB_OK;
}

int32_t prof_begin(iskb_ctx *c) {
  if (!c->profile) return ISKB_OK;
  if (c->prof_used + 2 > c->prof_ev.size()) {
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreate(&a));
    CU_TRY(cudaEventCreate(&b));
    c->prof_ev.push_back(a);
    c->prof_ev.push_back(b);
  }
  CU_TRY(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  return ISKB_OK;
}
int32_t prof_end(iskb_ctx *c) {
  if (!c->profile) return ISKB_OK;
  CU_TRY(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream));
  c->prof_used += 2;
  c->prof_launches++;
  return ISKB_OK;
}
extern "C" int32_t iskb_profile_enable(iskb_ctx *c, int32_t on) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  c->profile = on != 0;
  return ISKB_OK;
}
extern "C" int32_t iskb_profile_read(iskb_ctx *c, double *ms, int64_t *launches) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  CU_TRY(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k + 1 < c->prof_used; k += 2) {
    float t = 0.f;
    CU_TRY(cudaEventElapsedTime(&t, c->prof_ev[k], c->prof_ev[k + 1]));
    c->prof_ms += t;
  }
  c->prof_used = 0;
  if (ms) *ms = c->prof_ms;
  if (launches) *launches = c->prof_launches;
  c->prof_ms = 0.0;
  c->prof_launches = 0;
  return ISKB_OK;
}

int32_t ctx_check_status(iskb_ctx *c) {
  CU_TRY(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  int st = *c->h_status;
  if (!st) return ISKB_OK;
  CU_TRY(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
  if (st & ISKB_ST_TOO_FAST) c->warn_too_fast = true;   // the reference only prints a message (check.jl:44-46)
  st &= ~ISKB_ST_TOO_FAST;
  if (!st) return ISKB_OK;
  if (st & ISKB_ST_WALK) return iskb_fail(ISKB_E_UNSUPPORTED, "a tracked particle crossed more than 65536 cell faces in one step");
  if (st & ISKB_ST_CAPACITY) return iskb_fail(ISKB_E_CAPACITY, "species capacity exceeded while appending particles");
  if (st & ISKB_ST_PK) return iskb_fail(ISKB_E_PK, "collision probability P_k > 1 (energy outside of the range)");
  return iskb_fail(ISKB_E_OOB, "a live particle lies outside the grid (reference would raise BoundsError)");
}

// ---- grid ---------------------------------------------------------------------------------------
extern "C" int32_t iskb_grid_set(iskb_ctx *c, int32_t nx, int32_t ny, double dx, double dy, double ox, double oy,
                                 const int32_t bcs[4]) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  if (nx < 2 || ny < 2 || !(dx > 0) || !(dy > 0)) return iskb_fail(ISKB_E_INVALID, "grid needs nx, ny >= 2 and dh > 0");
  if (c->has_grid) return iskb_fail(ISKB_E_INVALID, "grid already set for this context");
  CU_TRY(cudaSetDevice(c->device));
  c->g.nx = nx; c->g.ny = ny; c->g.dx = dx; c->g.dy = dy; c->g.ox = ox; c->g.oy = oy;
  c->g.Lx = (double)(nx - 1) * dx;   // wrap.jl:3-4  Lx = nx*dx with nx = grid.n[i]-1
  c->g.Ly = (double)(ny - 1) * dy;
  c->g.rdx = 1.0 / dx;   // IEEE division: correctly rounded reciprocals
  c->g.rdy = 1.0 / dy;
  auto all_ones = [](double v) {
    uint64_t b;
    memcpy(&b, &v, sizeof(b));
    return (b & 0xfffffffffffffull) == 0xfffffffffffffull;
  };
  c->g.fast_div = (all_ones(dx) || all_ones(dy)) ? 0 : 1;
  for (int k = 0; k < 4; ++k) c->bcs[k] = bcs ? bcs[k] : ISKB_BC_OPEN;
  const int64_t nn = (int64_t)nx * ny;
  CU_TRY(cudaMalloc(&c->d_V, nn * sizeof(double)));
  CU_TRY(cudaMalloc(&c->d_rho, nn * sizeof(double)));
  CU_TRY(cudaMalloc(&c->d_phi, nn * sizeof(double)));
  CU_TRY(cudaMalloc(&c->d_E2, nn * sizeof(double2)));
  CU_TRY(cudaMemsetAsync(c->d_rho, 0, nn * sizeof(double), c->stream));   // zeros, ParticleInCell.jl:97-99
  CU_TRY(cudaMemsetAsync(c->d_phi, 0, nn * sizeof(double), c->stream));
  CU_TRY(cudaMemsetAsync(c->d_E2, 0, nn * sizeof(double2), c->stream));
  k_cell_volume<<<blocks_for(c, nn), TPB, 0, c->stream>>>(nx, ny, dx, dy, c->bcs[0], c->bcs[1], c->bcs[2], c->bcs[3], c->d_V);
  LAUNCH_CHECK(c);
  c->has_grid = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_cell_volume(iskb_ctx *c, double *V_out) {
  if (!c || !c->has_grid || !V_out) return iskb_fail(ISKB_E_INVALID, "no grid");
  CU_TRY(cudaMemcpyAsync(V_out, c->d_V, (int64_t)c->g.nx * c->g.ny * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}

extern "C" int32_t iskb_fields_download(iskb_ctx *c, double *rho, double *phi, double *E) {
  if (c) ISKB_TRY(fields_join(c));
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "no grid");
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  double *tmp = nullptr;
  if (rho) {
    ISKB_TRY(rho_materialize(c));   // fused tiled step: rho is still in the fixed-point sums
    CU_TRY(cudaMemcpyAsync(rho, c->d_rho, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (phi) CU_TRY(cudaMemcpyAsync(phi, c->d_phi, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (E) {
    CU_TRY(cudaMalloc(&tmp, 3 * nn * sizeof(double)));
    k_E_split<<<blocks_for(c, nn), TPB, 0, c->stream>>>(c->d_E2, nn, tmp);
    LAUNCH_CHECK(c);
    CU_TRY(cudaMemcpyAsync(E, tmp, 3 * nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(tmp);
  return ISKB_OK;
}

extern "C" int32_t iskb_fields_upload(iskb_ctx *c, const double *rho, const double *phi, const double *E) {
  if (c) ISKB_TRY(fields_join(c));
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "no grid");
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  double *tmp = nullptr;
  if (rho) {
    c->rho_lazy = false;
    CU_TRY(cudaMemcpyAsync(c->d_rho, rho, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  if (phi) CU_TRY(cudaMemcpyAsync(c->d_phi, phi, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (E) {
    CU_TRY(cudaMalloc(&tmp, 2 * nn * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(tmp, E, 2 * nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_E_interleave<<<blocks_for(c, nn), TPB, 0, c->stream>>>(tmp, nn, c->d_E2);
    LAUNCH_CHECK(c);
  }
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(tmp);
  return ISKB_OK;
}

// ---- species ------------------------------------------------------------------------------------
extern "C" int32_t iskb_species_create(iskb_ctx *c, int64_t capacity, double q, double m, double w0,
                                       iskb_species **out) {
  if (!c || !c->has_grid || !out) return iskb_fail(ISKB_E_INVALID, "iskb_grid_set must be called first");
  if (capacity < 1 || capacity > 0xfffffff0ll) return iskb_fail(ISKB_E_INVALID, "capacity out of range");
  CU_TRY(cudaSetDevice(c->device));
  iskb_species *s = new iskb_species();
  s->ctx = c; s->cap = capacity; s->q = q; s->m = m; s->w0 = w0;
  for (int k = 0; k < 6; ++k) {
    CU_TRY(cudaMalloc(&s->col[k], capacity * sizeof(double)));
    if (k < 5) CU_TRY(cudaMemsetAsync(s->col[k], 0, capacity * sizeof(double), c->stream));   // zeros(N,D), zeros(N,V)
  }
  CU_TRY(cudaMalloc(&s->id, capacity * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&s->d_vmax2, sizeof(unsigned long long)));
  CU_TRY(cudaMalloc(&s->d_vz2max, sizeof(unsigned long long)));
  CU_TRY(cudaMalloc(&s->d_cnt, CNT_N * sizeof(int64_t)));
  CU_TRY(cudaMemsetAsync(s->d_cnt, 0, CNT_N * sizeof(int64_t), c->stream));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  CU_TRY(cudaMalloc(&s->d_u, nn * sizeof(double)));
  CU_TRY(cudaMalloc(&s->d_n, nn * sizeof(double)));
  CU_TRY(cudaMemsetAsync(s->d_u, 0, nn * sizeof(double), c->stream));
  CU_TRY(cudaMemsetAsync(s->d_n, 0, nn * sizeof(double), c->stream));
  k_init_species<<<blocks_for(c, capacity), TPB, 0, c->stream>>>(s->col[5], s->id, capacity, w0);
  LAUNCH_CHECK(c);
  ISKB_TRY(sp_vmax_unknown(s));
  c->species.push_back(s);
  *out = s;
  return ISKB_OK;
}

// |v|^2 bound bookkeeping (used by the MCC pruning): +inf = unknown, 0 = about to be recomputed
int32_t sp_vmax_unknown(iskb_species *s) {
  static const unsigned long long inf_bits = 0x7ff0000000000000ull;
  CU_TRY(cudaMemcpyAsync(s->d_vmax2, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, s->ctx->stream));
  CU_TRY(cudaMemcpyAsync(s->d_vz2max, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, s->ctx->stream));
  s->vz2_known = false;
  return ISKB_OK;
}
int32_t sp_vmax_reset(iskb_species *s) {
  CU_TRY(cudaMemsetAsync(s->d_vmax2, 0, sizeof(unsigned long long), s->ctx->stream));
  return ISKB_OK;
}

int32_t sp_ensure_alt(iskb_species *s) {
  if (s->alt[0]) return ISKB_OK;
  for (int k = 0; k < 6; ++k) CU_TRY(cudaMalloc(&s->alt[k], s->cap * sizeof(double)));
  CU_TRY(cudaMalloc(&s->alt_id, s->cap * sizeof(uint32_t)));
  // the lean re-grouping launches never write wg (all weights are w0 then): the twin column must hold them already
  CU_TRY(cudaMemcpyAsync(s->alt[5], s->col[5], s->cap * sizeof(double), cudaMemcpyDeviceToDevice, s->ctx->stream));
  return ISKB_OK;
}

int32_t sp_sync_counts(iskb_species *s) {
  iskb_ctx *c = s->ctx;
  if (!s->counts_stale) return ISKB_OK;
  CU_TRY(cudaMemcpyAsync(c->h_scratch, s->d_cnt, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  s->h_nslots = c->h_scratch[CNT_NSLOTS];
  s->h_ndead = c->h_scratch[CNT_NDEAD];
  s->counts_stale = false;
  return ISKB_OK;
}

static int32_t set_counts(iskb_species *s, int64_t nslots, int64_t ndead) {
  iskb_ctx *c = s->ctx;
  k_set_counts<<<1, 1, 0, c->stream>>>(s->d_cnt, nslots, ndead);
  LAUNCH_CHECK(c);
  s->h_nslots = nslots; s->h_ndead = ndead; s->counts_stale = false;
  s->h_nsorted = 0;   // set_counts is only used by upload / sample / copy: the layout is unknown
  sp_touch(s);
  return ISKB_OK;
}

extern "C" int32_t iskb_species_upload(iskb_species *s, const double *x, const double *v, const double *wg,
                                       const uint32_t *id, int64_t np, int64_t ld) {
  if (!s) return iskb_fail(ISKB_E_INVALID, "null species");
  iskb_ctx *c = s->ctx;
  if (np < 0 || np > s->cap) return iskb_fail(ISKB_E_CAPACITY, "np = %lld exceeds capacity %lld", (long long)np, (long long)s->cap);
  if (np > 0 && (!x || !v || ld < np)) return iskb_fail(ISKB_E_INVALID, "x, v required with ld >= np");
  sp_touch(s);   // (before the copies: kernels that ran ahead on the old rows are waited for)
  const size_t b = (size_t)np * sizeof(double);
  if (np > 0) {
    CU_TRY(cudaMemcpyAsync(s->col[0], x, b, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(s->col[1], x + ld, b, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(s->col[2], v, b, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(s->col[3], v + ld, b, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(s->col[4], v + 2 * ld, b, cudaMemcpyHostToDevice, c->stream));
  }
  if (wg) {
    CU_TRY(cudaMemcpyAsync(s->col[5], wg, s->cap * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    bool uni = true;   // every weight equal to w0 (the reference's `ones(N) * weight`): the lean kernels skip the column
    double wmax = s->w0;
    for (int64_t k = 0; k < s->cap; ++k) {
      uni = uni && wg[k] == s->w0;
      if (wg[k] > wmax) wmax = wg[k];
    }
    s->wg_uniform = uni;
    s->wmax = wmax;   // bound of the fixed-point deposit (fixed_point_setup)
    if (s->alt[5]) CU_TRY(cudaMemcpyAsync(s->alt[5], wg, s->cap * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  if (id) CU_TRY(cudaMemcpyAsync(s->id, id, s->cap * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  ISKB_TRY(set_counts(s, np, 0));
  ISKB_TRY(sp_vmax_unknown(s));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}

extern "C" int32_t iskb_species_download(iskb_species *s, double *x, double *v, double *wg, uint32_t *id, int64_t ld) {
  if (!s) return iskb_fail(ISKB_E_INVALID, "null species");
  iskb_ctx *c = s->ctx;
  ISKB_TRY(sp_compact(s));
  const int64_t np = s->h_nslots;
  if ((x || v) && ld < np) return iskb_fail(ISKB_E_INVALID, "ld < np");
  const size_t b = (size_t)np * sizeof(double);
  if (x && np) {
    CU_TRY(cudaMemcpyAsync(x, s->col[0], b, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(x + ld, s->col[1], b, cudaMemcpyDeviceToHost, c->stream));
  }
  if (v && np) {
    CU_TRY(cudaMemcpyAsync(v, s->col[2], b, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(v + ld, s->col[3], b, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(v + 2 * ld, s->col[4], b, cudaMemcpyDeviceToHost, c->stream));
  }
  if (wg) CU_TRY(cudaMemcpyAsync(wg, s->col[5], s->cap * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (id) CU_TRY(cudaMemcpyAsync(id, s->id, s->cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ctx_check_status(c);
}

extern "C" int32_t iskb_species_np(iskb_species *s, int64_t *np_out) {
  if (!s || !np_out) return iskb_fail(ISKB_E_INVALID, "null");
  ISKB_TRY(sp_sync_counts(s));
  *np_out = s->h_nslots - s->h_ndead;
  return ISKB_OK;
}

extern "C" int32_t iskb_species_window_stats(iskb_species *s, int64_t out[4]) {
  if (!s || !out) return iskb_fail(ISKB_E_INVALID, "null");
  iskb_ctx *c = s->ctx;
  CU_TRY(cudaMemcpyAsync(c->h_scratch, s->d_cnt + 3, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemsetAsync(s->d_cnt + 3, 0, 4 * sizeof(int64_t), c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 4; ++k) out[k] = c->h_scratch[k];
  return ISKB_OK;
}

extern "C" int32_t iskb_species_sample_maxwellian(iskb_species *s, int64_t n, const double wx[2], const double dx[2],
                                                  const double wv[3], const double dv[3], uint64_t seed) {
  if (!s || !wx || !wv) return iskb_fail(ISKB_E_INVALID, "null");
  iskb_ctx *c = s->ctx;
  ISKB_TRY(sp_compact(s));
  const int64_t free_rows = s->cap - s->h_nslots;
  if (n > free_rows) n = free_rows;                      // sources.jl:30  minimum([size(px,1), ...])
  if (n <= 0) return ISKB_OK;
  const double z2[2] = {0, 0}, z3[3] = {0, 0, 0};
  if (!dx) dx = z2;
  if (!dv) dv = z3;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (0x85EBCA6Bu * (uint32_t)(c->rank + 1));
  k_sample<<<blocks_for(c, n), TPB, 0, c->stream>>>(s->col[0], s->col[1], s->col[2], s->col[3], s->col[4], s->h_nslots, n,
                                                   wx[0], wx[1], dx[0], dx[1], wv[0], wv[1], wv[2], dv[0], dv[1], dv[2],
                                                   k0, k1, (uint32_t)(s->sample_calls++));
  LAUNCH_CHECK(c);
  ISKB_TRY(sp_vmax_unknown(s));
  return set_counts(s, s->h_nslots + n, 0);
}

extern "C" int32_t iskb_species_copy_positions(iskb_species *dst, iskb_species *src, const double v_fill[3]) {
  if (!dst || !src || dst->ctx != src->ctx) return iskb_fail(ISKB_E_INVALID, "bad species");
  iskb_ctx *c = dst->ctx;
  ISKB_TRY(sp_compact(src));
  const int64_t n = src->h_nslots;
  if (n > dst->cap) return iskb_fail(ISKB_E_CAPACITY, "destination capacity too small");
  CU_TRY(cudaMemcpyAsync(dst->col[0], src->col[0], n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(dst->col[1], src->col[1], n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  const double z[3] = {0, 0, 0};
  if (!v_fill) v_fill = z;
  if (n) {
    k_fill3<<<blocks_for(c, n), TPB, 0, c->stream>>>(dst->col[2], dst->col[3], dst->col[4], n, v_fill[0], v_fill[1], v_fill[2]);
    LAUNCH_CHECK(c);
  }
  ISKB_TRY(sp_vmax_unknown(dst));
  return set_counts(dst, n, 0);
}

// ---- fused step: ParticleInCell.jl:102-135 ------------------------------------------------------
extern "C" int32_t iskb_set_after_push(iskb_ctx *c, int32_t mx, int32_t my) {
  if (!c || mx < 0 || mx > 2 || my < 0 || my > 2) return iskb_fail(ISKB_E_INVALID, "bad boundary mode");
  c->after_push[0] = mx; c->after_push[1] = my;
  return ISKB_OK;
}
extern "C" int32_t iskb_set_sort_interval(iskb_ctx *c, int32_t interval) {
  if (!c || interval < 0) return iskb_fail(ISKB_E_INVALID, "bad interval");
  c->sort_interval = interval;
  return ISKB_OK;
}

extern "C" int32_t iskb_set_sort_policy(iskb_ctx *c, double miss_threshold, int32_t max_interval) {
  if (!c || miss_threshold < 0 || max_interval < 0) return iskb_fail(ISKB_E_INVALID, "bad sort policy");
  c->sort_miss_threshold = miss_threshold;
  c->sort_max_interval = max_interval;
  return ISKB_OK;
}

extern "C" int32_t iskb_set_sort_full_interval(iskb_ctx *c, int32_t full_interval) {
  if (!c || full_interval < 0) return iskb_fail(ISKB_E_INVALID, "bad interval");
  c->sort_full_interval = full_interval;
  return ISKB_OK;
}

// Decide whether species s is re-sorted before this step.  The window statistics travel through
// an async copy + event; the host waits for the snapshot taken TWO steps ago, which keeps one whole
// step of work queued on the GPU (no bubble) while bounding how far the host runs ahead.
static int32_t maybe_sort(iskb_ctx *c, iskb_species *s) {
  if (c->sort_miss_threshold > 0.0 && s->h_wstats) {
    const int slot = (int)(s->wstats_step & 1);   // the older of the two snapshots
    if (s->wstats_pending[slot]) {
      CU_TRY(cudaEventSynchronize(s->ev_wstats[slot]));
      const int64_t g = s->h_wstats[4 * slot];
      const int64_t rows = s->h_nslots > 0 ? s->h_nslots : 1;
      const int64_t d = g >= s->last_gmiss ? g - s->last_gmiss : g;   // counters may have been reset by the user
      s->last_gmiss = g;
      // this snapshot has index wstats_step-2; ignore it if it predates the last sort
      if (s->wstats_step - 2 >= s->wstats_sort_mark) s->miss_rate = (double)d / (double)rows;
      s->wstats_pending[slot] = false;
    }
  }
  bool do_sort = s->steps_since_sort >= c->sort_interval;
  if (do_sort && c->sort_miss_threshold > 0.0 && s->steps_since_sort < (1 << 29)) {
    const bool too_old = c->sort_max_interval > 0 && s->steps_since_sort >= c->sort_max_interval;
    do_sort = too_old || s->miss_rate > c->sort_miss_threshold;
  }
  if (do_sort) {
    // full sort (cells + checkerboard interleave) the first time and every sort_full_interval steps;
    // in between a cheaper stable re-group by tile
    const bool full = c->sort_full_interval <= 0 || s->steps_since_full >= c->sort_full_interval;
    if (full) {
      ISKB_TRY(sp_sort(s, nullptr, true));
      s->steps_since_full = 0;
    } else {
      ISKB_TRY(sp_regroup(s));
    }
    s->steps_since_sort = 0;
    s->miss_rate = 0.0;
    s->wstats_sort_mark = s->wstats_step;
  }
  return ISKB_OK;
}

static int32_t post_advance_stats(iskb_ctx *c, iskb_species *s) {
  if (c->sort_miss_threshold <= 0.0) return ISKB_OK;
  if (!s->h_wstats) {
    CU_TRY(cudaMallocHost(&s->h_wstats, 8 * sizeof(int64_t)));
    CU_TRY(cudaEventCreateWithFlags(&s->ev_wstats[0], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&s->ev_wstats[1], cudaEventDisableTiming));
  }
  const int slot = (int)(s->wstats_step & 1);
  CU_TRY(cudaMemcpyAsync(s->h_wstats + 4 * slot, s->d_cnt + 3, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaEventRecord(s->ev_wstats[slot], c->stream));
  s->wstats_pending[slot] = true;
  s->wstats_step++;
  return ISKB_OK;
}

extern "C" int32_t iskb_rho_allreduce(iskb_ctx *c) {
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "no grid");
  if (c->n_ranks == 1) return ISKB_OK;
  return comm_allreduce_sum(c, c->d_rho, (int64_t)c->g.nx * c->g.ny);
}

// ---- re-group policy of the tile-aware path (advance_tile.cu) -----------------------------------------
// Per species and step: a FULL sort (radix sort by cell + interleave, builds the tile directory) when there is no
// valid directory, when the unsorted tail has grown past 1 % of the rows, or every sort_full_interval steps if
// that is set; otherwise the advance itself re-groups the rows on its way out (`move`) when the window-miss
// rate of the launch two steps back exceeds the threshold, when discarded rows make up more than 2 % of the
// slots, or after sort_max_interval steps (fixed mode: every sort_interval steps).  The statistics travel
// through an async copy + event, two steps late, so the host never waits for the step it has just queued.
static int32_t tile_policy(iskb_ctx *c, iskb_species *s, bool *move, bool *mark) {
  *move = false;
  *mark = false;
  if (s->h_tstats) {
    const int slot = (int)(s->tstats_step & 1);   // the older of the two snapshots
    if (s->tstats_pending[slot]) {
      CU_TRY(cudaEventSynchronize(s->ev_tstats[slot]));
      const int64_t *h = s->h_tstats + 10 * slot;
      const int64_t n = h[CNT_NSLOTS] > 0 ? h[CNT_NSLOTS] : 1;
      const int64_t g = h[3];
      const int64_t d = g >= s->last_gmiss ? g - s->last_gmiss : g;   // the counters may have been reset by the caller
      s->last_gmiss = g;
      if (s->tstats_step - 2 >= s->tstats_sort_mark) {   // ignore snapshots that predate the last re-ordering
        s->miss_rate = (double)d / (double)n;
        if (s->miss_rate > 1e-5) s->drifting = true;
        s->tail_frac = (double)(h[CNT_NSLOTS] - h[9]) / (double)n;
        s->tail_rows = h[CNT_NSLOTS] - h[9];
        s->dead_frac = (double)h[CNT_NDEAD] / (double)n;
      }
      s->tstats_pending[slot] = false;
    }
  }
  // the unsorted tail (rows born since the last re-group) joins the tiles on a MOVE; a full sort is for a layout that
  // has no directory yet (upload, host-side edits) or a tail beyond the merge scratch (cap / 16 rows)
  const bool full = !s->tdir_valid || s->tail_frac > 0.05 ||
                    (c->sort_full_interval > 0 && s->steps_since_full >= c->sort_full_interval);
  if (full) {
    ISKB_TRY(sp_sort(s, nullptr, true));
    s->steps_since_full = 0;
    s->full_sorts++;
    s->miss_rate = s->tail_frac = s->dead_frac = 0.0;
    s->tstats_sort_mark = s->tstats_step;
    // (the sort left valid marks behind; the launch of this step keeps them valid only if a MOVE is due right after it)
    *mark = c->sort_miss_threshold > 0.0 ? (c->sort_max_interval > 0 && c->sort_max_interval <= 2 && s->drifting)
                                         : c->sort_interval <= 2;
    return ISKB_OK;
  }
  // A MOVE needs the codes and counts of the launch before it (MARK).  Launches mark only when a MOVE is due at the
  // next step: the periodic one is known a step ahead, a MOVE asked for by the statistics waits one marked launch.
  const int64_t since = s->steps_since_move + 1;
  bool want, soon = false;
  if (c->sort_miss_threshold > 0.0) {
    // a species whose rows hardly ever leave their windows (ions) is not re-grouped just because time has passed
    want = s->miss_rate > c->sort_miss_threshold || s->dead_frac > 0.02 || s->tail_frac > 0.01 ||
           (c->sort_max_interval > 0 && since >= c->sort_max_interval && s->drifting);
    soon = c->sort_max_interval > 0 && since + 1 >= c->sort_max_interval && s->drifting;
  } else {
    want = since >= c->sort_interval || s->tail_frac > 0.01;
    soon = since + 1 >= c->sort_interval;
  }
  *move = want && s->marks_valid;
  *mark = (want && !s->marks_valid) || soon;
  if (*move)   // a MOVE marks only when the very next launch is a MOVE again (intervals of one or two steps)
    *mark = c->sort_miss_threshold > 0.0 ? (c->sort_max_interval > 0 && c->sort_max_interval <= 2 && s->drifting) : c->sort_interval <= 2;
  if (*move) {
    s->miss_rate = s->dead_frac = s->tail_frac = 0.0;
    // "drifting" is earned again after every re-group (by the first snapshots taken after it): electrons are back over the
    // bar within two steps, ions -- whose strays pile up over ~100 steps and are then cured by ONE re-group -- are not, so
    // they do not go on re-grouping every max_interval steps for the rest of the run (a 500-step run showed exactly that)
    s->drifting = false;
    s->tstats_sort_mark = s->tstats_step;
    s->moves++;
  }
  return ISKB_OK;
}

static int32_t tile_stats_snapshot(iskb_ctx *c, iskb_species *s) {
  if (!s->h_tstats) {
    CU_TRY(cudaMallocHost(&s->h_tstats, 20 * sizeof(int64_t)));
    memset(s->h_tstats, 0, 20 * sizeof(int64_t));
    CU_TRY(cudaEventCreateWithFlags(&s->ev_tstats[0], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&s->ev_tstats[1], cudaEventDisableTiming));
  }
  const int slot = (int)(s->tstats_step & 1);
  const TileGeom tg = tile_geom(c->g);
  s->h_tstats[10 * slot + 9] = 0;
  CU_TRY(cudaMemcpyAsync(s->h_tstats + 10 * slot, s->d_cnt, CNT_N * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(s->h_tstats + 10 * slot + 9, s->d_ts[0] + tg.ntiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaEventRecord(s->ev_tstats[slot], c->stream));
  s->tstats_pending[slot] = true;
  s->tstats_step++;
  return ISKB_OK;
}

extern "C" int32_t iskb_set_lean(iskb_ctx *c, int32_t on) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  c->lean_ok = on != 0;
  return ISKB_OK;
}

extern "C" int32_t iskb_set_advance_path(iskb_ctx *c, int32_t path) {
  if (!c || path < 0 || path > 1) return iskb_fail(ISKB_E_INVALID, "advance path is 0 (tile directory) or 1 (per-warp windows)");
  c->adv_path = path;
  return ISKB_OK;
}

extern "C" int32_t iskb_species_sort_stats(iskb_species *s, int64_t out[8]) {
  if (!s || !out) return iskb_fail(ISKB_E_INVALID, "null");
  out[0] = s->full_sorts;
  out[1] = s->moves;
  out[2] = s->steps_since_full;
  out[3] = s->steps_since_move;
  // the newest snapshot the policy has looked at (two steps old): slots, discarded rows inside them, rows in the tile directory
  const int slot = (int)((s->tstats_step + 1) & 1);
  out[4] = s->h_tstats ? s->h_tstats[10 * slot + CNT_NSLOTS] : 0;
  out[5] = s->h_tstats ? s->h_tstats[10 * slot + CNT_NDEAD] : 0;
  out[6] = s->h_tstats ? s->h_tstats[10 * slot + 9] : 0;
  out[7] = 0;
  return ISKB_OK;
}

// Fixed-point unit of the tile path's deposit (advance_tile.cu, add_fixed).  q0 = the smallest |charge|; every active
// species must carry an integer multiple Z_s of it (e, -e, 2e ...), so that rho = q0 * sum_s Z_s u_s / V can be formed
// from integers.  fscale = the largest power of two for which sum_s |Z_s| * capacity_s * wmax_s * fscale < 2^62:
// even with every slot of every species in one node nothing overflows.
static int32_t fixed_point_setup(iskb_ctx *c, const std::vector<iskb_species *> &species) {
  double q0 = 0.0;
  for (iskb_species *s : species) {
    const double a = fabs(s->q);
    if (a > 0.0 && (q0 == 0.0 || a < q0)) q0 = a;
  }
  if (q0 == 0.0) q0 = 1.0;   // neutral species only: rho == 0 whatever the unit
  double total = 0.0;
  for (iskb_species *s : species) {
    const double z = s->q / q0, zr = (double)llround(z);
    if (fabs(z - zr) > 1e-9 * (fabs(z) + 1.0))
      return iskb_fail(ISKB_E_UNSUPPORTED, "species charges are not integer multiples of the smallest one (%g vs %g)", s->q, q0);
    if (s->wmax <= 0.0) s->wmax = s->w0 > 0.0 ? s->w0 : 1.0;
    total += (fabs(zr) > 1.0 ? fabs(zr) : 1.0) * (double)s->cap * s->wmax;
  }
  total *= (double)(c->n_ranks > 1 ? c->n_ranks : 1);   // the all-reduce adds the sums of every rank (slices of equal size)
  int e = 0;
  frexp(4.0e18 / total, &e);          // 2^(e-1) <= 4e18 / total < 2^e  (4e18 < 2^62)
  c->fscale = ldexp(1.0, e - 1);
  c->q0 = q0;
  return ISKB_OK;
}

extern "C" int32_t iskb_ctx_counts(iskb_ctx *c, int64_t *n_species, int64_t *n_mcc, int64_t *n_dsmc) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  if (n_species) *n_species = (int64_t)c->species.size();
  if (n_mcc) *n_mcc = (int64_t)c->mccs.size();
  if (n_dsmc) *n_dsmc = (int64_t)c->dsmcs.size();
  return ISKB_OK;
}

extern "C" int32_t iskb_step_set_active(iskb_ctx *c, iskb_species *const *species, int32_t n_species,
                                        void *const *interactions, int32_t n_interactions) {
  if (!c) return iskb_fail(ISKB_E_INVALID, "null ctx");
  if (n_species < 0) {   // back to "everything on the context"
    c->active_set = false;
    c->active_species.clear();
    c->active_inter.clear();
    return ISKB_OK;
  }
  if ((n_species > 0 && !species) || (n_interactions > 0 && !interactions) || n_interactions < 0)
    return iskb_fail(ISKB_E_INVALID, "iskb_step_set_active: null list");
  std::vector<iskb_species *> sp;
  std::vector<std::pair<int, void *>> in;
  for (int k = 0; k < n_species; ++k) {
    bool ok = false;
    for (iskb_species *s : c->species) ok = ok || s == species[k];
    if (!ok) return iskb_fail(ISKB_E_INVALID, "iskb_step_set_active: species %d does not belong to this context", k);
    sp.push_back(species[k]);
  }
  for (int k = 0; k < n_interactions; ++k) {
    int kind = -1;
    for (iskb_mcc *m : c->mccs) if ((void *)m == interactions[k]) kind = 0;
    for (iskb_dsmc *d : c->dsmcs) if ((void *)d == interactions[k]) kind = 1;
    if (kind < 0) return iskb_fail(ISKB_E_INVALID, "iskb_step_set_active: interaction %d does not belong to this context", k);
    in.push_back(std::make_pair(kind, interactions[k]));
  }
  c->active_species.swap(sp);
  c->active_inter.swap(in);
  c->active_set = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_step(iskb_ctx *c, double dt, int32_t n_steps) {
  if (!c || !c->has_grid || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "grid and Poisson solver must be set");
  CU_TRY(cudaSetDevice(c->device));
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  ISKB_TRY(poisson_prepare(c));
  const std::vector<iskb_species *> &species = c->active_set ? c->active_species : c->species;
  for (int it = 0; it < n_steps; ++it) {
    const bool windows_fit = c->g.nx >= 20 && c->g.ny >= 20;   // small / quasi-1D grids use the simple kernels
    const bool tiled = c->sort_interval > 0 && windows_fit;
    // tile directory path: everything but the surface tracker (which still runs on the per-warp windows of advance_fused.cu)
    // A warp owns whole tiles: a grid of a few tiles cannot feed 148 SMs (33 x 65 nodes = 32 tiles: 4e7 rows took 0.7 s per
    // step).  Many rows on fewer than 2048 tiles go to the kernels that split by rows (per-warp windows / simple kernels).
    int64_t rows_bound = 0;
    for (const iskb_species *s : species) rows_bound += s->counts_stale ? s->cap : s->h_nslots;
    const TileGeom tg0 = tile_geom(c->g);
    const bool tiles_feed_gpu = (int64_t)tg0.tiles_x * tg0.tiles_y >= 2048 || rows_bound <= 4000000;
    const bool tile_dir = tiled && !c->tracker && c->adv_path == 0 && c->g.fast_div && tiles_feed_gpu;   // (a dh with an all-ones significand: simple kernels)
    const bool legacy = tiled && !tile_dir && !c->pusher_rz && (c->tracker || c->adv_path == 1);
    if (c->pusher_rz && c->tracker) return iskb_fail(ISKB_E_UNSUPPORTED, "surface tracker with the axial pusher");
    c->rho_lazy = false;   // the sums of the previous step are about to be zeroed
    bool move[64], mark[64];
    if (species.size() > 64) return iskb_fail(ISKB_E_UNSUPPORTED, "more than 64 species");
    if (tile_dir) {
      for (size_t k = 0; k < species.size(); ++k) ISKB_TRY(tile_policy(c, species[k], &move[k], &mark[k]));
    } else if (legacy) {
      for (iskb_species *s : species) ISKB_TRY(maybe_sort(c, s));
    }
    // :109-111, config.interactions in order.  The last one, if it is an MCC with a neutral target, runs on its own
    // stream: it only touches its source and product species, so the species advanced before those overlap it.
    std::vector<std::pair<int, void *>> inter;
    if (c->active_set) inter = c->active_inter;
    else {
      for (iskb_mcc *m : c->mccs) inter.push_back(std::make_pair(0, (void *)m));
      for (iskb_dsmc *d : c->dsmcs) inter.push_back(std::make_pair(1, (void *)d));
    }
    iskb_mcc *deferred = nullptr;
    bool only_neutral_mcc = tile_dir;   // the test phases may run ahead only if nothing else edits velocities in between
    for (const std::pair<int, void *> &in : inter) only_neutral_mcc = only_neutral_mcc && in.first == 0 && ((iskb_mcc *)in.second)->tq == 0.0;
    bool waited_pre = false;
    for (size_t k = 0; k < inter.size(); ++k) {
      if (inter[k].first == 1) { ISKB_TRY(dsmc_launch((iskb_dsmc *)inter[k].second, dt, false)); continue; }
      iskb_mcc *m = (iskb_mcc *)inter[k].second;
      // phase 1 (selection + test) of this step may have run right after the previous advance of the source species
      int phase = 3;
      if (m->pre_valid) {
        if (only_neutral_mcc && m->pre_dt == dt && m->pre_epoch == m->source->epoch) {
          if (!waited_pre) {   // every test phase that ran ahead has taken its snapshot before anything appends rows
            CU_TRY(cudaEventRecord(c->ev_p0, c->pstream));
            CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_p0, 0));
            waited_pre = true;
          }
          phase = 2;
          m->pre_valid = false;
        } else {
          ISKB_TRY(mcc_discard_pre(m));
        }
      }
      if (k + 1 == inter.size() && inter.size() > 1 && m->tq == 0.0 && tile_dir) {
        CU_TRY(cudaEventRecord(c->ev_m0, c->stream));
        CU_TRY(cudaStreamWaitEvent(c->mstream, c->ev_m0, 0));
        ISKB_TRY(mcc_launch(m, dt, false, c->mstream, phase));
        CU_TRY(cudaEventRecord(c->ev_m1, c->mstream));
        deferred = m;
      } else {
        ISKB_TRY(mcc_launch(m, dt, false, nullptr, phase));
      }
    }
    ISKB_TRY(fields_join(c));   // E of the previous step (the re-sort and MCC above did not need it)
    if (tile_dir) ISKB_TRY(fixed_point_setup(c, species));
    for (size_t k = 0; k < species.size(); ++k) {                          // :113-115
      iskb_species *s = species[k];
      if (deferred) {   // the species the deferred MCC reads or appends to wait for it
        bool touches = deferred->source == s;
        for (const MccProc &p : deferred->procs) touches = touches || p.product == s;
        if (touches) {
          CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_m1, 0));
          deferred = nullptr;
        }
      }
      if (tile_dir) {
        if (s->n_in_ufix) {   // nobody asked for the density of the previous step: species.n keeps an older one
          s->n_in_ufix = false;
        }
        CU_TRY(cudaMemsetAsync(s->d_ufix, 0, nn * sizeof(long long), c->stream));
        ISKB_TRY(launch_advance_tile(s, dt, c->after_push[0], c->after_push[1], move[k], mark[k]));
        ISKB_TRY(tile_stats_snapshot(c, s));
        s->steps_since_move++;
        s->steps_since_full++;
        if (only_neutral_mcc) {
          // selection + acceptance test of the NEXT step for the MCC objects that draw from this species: they read
          // what this launch has just written and nothing else, so they run next to the other species' advance
          for (const std::pair<int, void *> &in : inter) {
            iskb_mcc *m = (iskb_mcc *)in.second;
            if (m->source != s) continue;
            CU_TRY(cudaEventRecord(c->ev_p0, c->stream));
            CU_TRY(cudaStreamWaitEvent(c->pstream, c->ev_p0, 0));
            ISKB_TRY(mcc_launch(m, dt, false, c->pstream, 1));
            CU_TRY(cudaEventRecord(m->ev_pre, c->pstream));
            m->pre_valid = true;
            m->pre_dt = dt;
            m->pre_epoch = s->epoch;
          }
        }
      } else if (c->tracker && !legacy) {
        // config.tracker != nothing: track! -> gather -> push -> check! -> after_push  (:56-61), one pass
        CU_TRY(cudaMemsetAsync(s->d_u, 0, nn * sizeof(double), c->stream));
        ISKB_TRY(launch_advance_tracked(s, dt, c->after_push[0], c->after_push[1], true));
      } else if (legacy) {
        CU_TRY(cudaMemsetAsync(s->d_u, 0, nn * sizeof(double), c->stream));
        if (c->tracker) ISKB_TRY(launch_advance_tiled_tracked(s, dt, c->after_push[0], c->after_push[1]));
        else ISKB_TRY(launch_advance_tiled(s, dt, c->after_push[0], c->after_push[1]));
        ISKB_TRY(post_advance_stats(c, s));
        s->steps_since_sort++;
        s->steps_since_full++;
      } else {
        CU_TRY(cudaMemsetAsync(s->d_u, 0, nn * sizeof(double), c->stream));
        ISKB_TRY(launch_advance_simple(s, dt, c->after_push[0], c->after_push[1], true, false));
      }
    }
    if (deferred) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_m1, 0));
    if (tile_dir) {
      ISKB_TRY(launch_rho_finalize_fixed(c, species));                     // :118-124, integer all-reduce inside
    } else {
      ISKB_TRY(launch_rho_finalize(c, &species));                          // :118-124
      if (c->n_ranks > 1) ISKB_TRY(comm_allreduce_sum(c, c->d_rho, nn));
    }
    ISKB_TRY(poisson_solve(c));                                            // :126-128
    c->step_count++;
  }
  return ISKB_OK;
}
