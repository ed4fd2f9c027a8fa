/* CPU ORACLE (plain C) -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Line-faithful single-thread restatement of bchaber/iskra's per-timestep particle hot
 * path (the reference is pure Julia and cannot run in this image).  Compiled with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp
 * so that no FMA contraction or re-association changes the reference's rounding.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The shipped package never does.
 *
 * File:line citations are relative to /root/reference.  Arrays are column-major and
 * indices 1-based in the comments, exactly like the Julia source; C indices are 0-based.
 *
 * PARITY PINNING: see oracle/pic_oracle.py header.  Deterministic particle arithmetic is
 * cross-checked bit-for-bit against the independent numpy restatement; sigma(eps)
 * interpolation (Interpolations.jl 0.13.1), the dense LU (LAPACK behind `A\b`) and all
 * RNG-dependent results are "parity unpinned".
 *
 * The *_mt entry points are OpenMP variants used only for the multi-core CPU baseline
 * timing; they change the deposit summation order and are never used as the checker.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  double *x, *y, *vx, *vy, *vz, *wg;   /* kinetic.jl:2-9  (x :: N x 2, v :: N x 3, column major) */
  uint32_t *id;                        /* kinetic.jl:10 */
  int64_t np, cap;                     /* kinetic.jl:7 ; capacity N */
  double q, m, w0;                     /* kinetic.jl:5,6,8 */
} orc_species;

typedef struct {
  int32_t nx, ny;                      /* RegularGrids.jl:9  n */
  double dx, dy;                       /* :10 dh */
  double ox, oy;                       /* :14 origin */
  int32_t bcs[4];                      /* :11 left,right,bottom,top ; 0 = open, 1 = periodic */
} orc_grid;

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- ParticleInCell.jl:28-35 particle_cell ------------------------------------------------ */
static inline void cell1(double x, double d, int64_t *i, double *h) {
  double f = 1.0 + x / d;            /* :29  divide, then add 1 */
  double fl = floor(f);              /* :30 */
  *i = (int64_t)fl;
  *h = f - fl;                       /* :31 */
}

void orc_particle_cell(const double *x, const double *y, int64_t np, double dx, double dy,
                       int64_t *ci, int64_t *cj, double *hx, double *hy) {
  for (int64_t p = 0; p < np; ++p) {
    cell1(x[p], dx, &ci[p], &hx[p]);
    cell1(y[p], dy, &cj[p], &hy[p]);
  }
}

/* ---- cloud_in_cell.jl:20-36 grid_to_particle ---------------------------------------------- */
/* E: nx*ny*3 column-major; pE: np x 3 column-major (component c at pE + c*np). */
static inline void gather1(const orc_grid *g, const double *E, double x, double y, double *e3) {
  int64_t i, j; double hx, hy;
  cell1(x, g->dx, &i, &hx); cell1(y, g->dy, &j, &hy);
  const int64_t nx = g->nx, nn = (int64_t)g->nx * g->ny;
  const int64_t n00 = (i - 1) + (j - 1) * nx;
  const double w00 = (1.0 - hx) * (1.0 - hy), w10 = hx * (1.0 - hy);
  const double w01 = (1.0 - hx) * hy, w11 = hx * hy;
  for (int c = 0; c < 3; ++c) {
    const double *u = E + c * nn;
    e3[c] = ((w00 * u[n00] + w10 * u[n00 + 1]) + w01 * u[n00 + nx]) + w11 * u[n00 + nx + 1];
  }
}

void orc_gather(const orc_grid *g, const orc_species *s, const double *E, double *pE) {
  const int64_t np = s->np;
  for (int64_t p = 0; p < np; ++p) {
    double e3[3];
    gather1(g, E, s->x[p], s->y[p], e3);
    pE[p] = e3[0]; pE[p + np] = e3[1]; pE[p + 2 * np] = e3[2];
  }
}

/* ---- pushers.jl:37-50 push_in_cartesian! with B == 0 (generalized_poisson.jl:412-419) ----- */
static inline void push1(double *x, double *y, double *vx, double *vy, double *vz,
                         const double *e3, double qm, double dt) {
  const double c1 = 0.5 * dt * qm;                 /* :41  (0.5dt*qm) */
  double vm[3] = { c1 * e3[0] + *vx, c1 * e3[1] + *vy, c1 * e3[2] + *vz };
  /* :42-46 with B = 0: t = 0, v' = v- + v- x 0, s = 0, v+ = v- + v' x 0  => v+ = v- (+0.0) */
  double vp[3] = { vm[0] + 0.0, vm[1] + 0.0, vm[2] + 0.0 };
  double v1[3] = { vm[0] + 0.0, vm[1] + 0.0, vm[2] + 0.0 };
  (void)vp;
  *vx = ((dt * e3[0]) * qm) * 0.5 + v1[0];         /* :48 */
  *vy = ((dt * e3[1]) * qm) * 0.5 + v1[1];
  *vz = ((dt * e3[2]) * qm) * 0.5 + v1[2];
  *x = dt * *vx + *x;                              /* :49 */
  *y = dt * *vy + *y;
}

void orc_push(orc_species *s, const double *pE, double dt) {
  const int64_t np = s->np;
  const double qm = s->q / s->m;                   /* :39 */
  for (int64_t p = 0; p < np; ++p) {
    double e3[3] = { pE[p], pE[p + np], pE[p + 2 * np] };
    push1(&s->x[p], &s->y[p], &s->vx[p], &s->vy[p], &s->vz[p], e3, qm, dt);
  }
}

/* ---- Julia Base: mod / fld for Float64 (float.jl, div.jl) --------------------------------- */
static inline double jl_mod(double x, double y) {
  double r = fmod(x, y);
  if (r == 0.0) return copysign(r, y);
  if ((r > 0.0) != (y > 0.0)) return r + y;
  return r;
}
static inline double jl_fld(double x, double y) { return rint((x - jl_mod(x, y)) / y); }

/* ---- kinetic.jl:20-27 remove! ------------------------------------------------------------- */
static void remove1(orc_species *s, int64_t i /*0-based*/) {
  const int64_t l = s->np - 1;
  s->x[i] = s->x[l]; s->y[i] = s->y[l];
  s->vx[i] = s->vx[l]; s->vy[i] = s->vy[l]; s->vz[i] = s->vz[l];
  s->wg[i] = s->wg[l]; s->wg[l] = s->w0;
  uint32_t t = s->id[i]; s->id[i] = s->id[l]; s->id[l] = t;
  s->np = l;
}

/* ---- surfaces/wrap.jl:20-33 wrap! , :1-18 discard!  (dim = 1 or 2) ------------------------ */
void orc_wrap(orc_species *s, const orc_grid *g, int dim) {
  double *px = dim == 1 ? s->x : s->y;
  const double L = (double)((dim == 1 ? g->nx : g->ny) - 1) * (dim == 1 ? g->dx : g->dy);
  const double o = dim == 1 ? g->ox : g->oy;
  for (int64_t p = 0; p < s->np; ++p) {
    double a = jl_fld(px[p] - o, L);
    if (a != 0) px[p] -= a * L;
  }
}

int64_t orc_discard(orc_species *s, const orc_grid *g, int dim) {
  const int64_t before = s->np;
  double *px = dim == 1 ? s->x : s->y;
  const double L = (double)((dim == 1 ? g->nx : g->ny) - 1) * (dim == 1 ? g->dx : g->dy);
  const double o = dim == 1 ? g->ox : g->oy;
  for (int64_t p = s->np - 1; p >= 0; --p) {       /* reverse(1:np) */
    double a = jl_fld(px[p] - o, L);
    if (a != 0) remove1(s, p);
  }
  return before - s->np;
}

/* ---- cloud_in_cell.jl:1-18 particle_to_grid with pu(p) = wg[p]  (kinetic.jl:53) ----------- */
void orc_deposit(const orc_grid *g, const orc_species *s, double *u /* nx*ny, zeroed here */) {
  const int64_t nx = g->nx;
  memset(u, 0, sizeof(double) * (size_t)g->nx * g->ny);
  for (int64_t p = 0; p < s->np; ++p) {
    int64_t i, j; double hx, hy;
    cell1(s->x[p], g->dx, &i, &hx); cell1(s->y[p], g->dy, &j, &hy);
    const int64_t n00 = (i - 1) + (j - 1) * nx;
    const double w = s->wg[p];
    u[n00]          += (1.0 - hx) * (1.0 - hy) * w;
    u[n00 + 1]      += (hx) * (1.0 - hy) * w;
    u[n00 + nx]     += (1.0 - hx) * (hy) * w;
    u[n00 + nx + 1] += (hx) * (hy) * w;
  }
}

/* ---- RegularGrids.jl:26-38 cell_volume ---------------------------------------------------- */
void orc_cell_volume(const orc_grid *g, double *V) {
  const int nx = g->nx, ny = g->ny;
  const double V0 = g->dx * g->dy;
  for (int64_t n = 0; n < (int64_t)nx * ny; ++n) V[n] = 0.0 + V0;
  if (g->bcs[0] != 1) for (int j = 0; j < ny; ++j) V[0 + (int64_t)j * nx] *= 0.5;
  if (g->bcs[1] != 1) for (int j = 0; j < ny; ++j) V[(nx - 1) + (int64_t)j * nx] *= 0.5;
  if (g->bcs[2] != 1) for (int i = 0; i < nx; ++i) V[i] *= 0.5;
  if (g->bcs[3] != 1) for (int i = 0; i < nx; ++i) V[i + (int64_t)(ny - 1) * nx] *= 0.5;
}

/* density (kinetic.jl:53) and rho accumulation (ParticleInCell.jl:118-124) */
void orc_density(const orc_grid *g, const orc_species *s, const double *V, double *n) {
  orc_deposit(g, s, n);
  for (int64_t k = 0; k < (int64_t)g->nx * g->ny; ++k) n[k] = n[k] / V[k];
}
void orc_rho_accumulate(const orc_grid *g, const double *n, double q, double *rho) {
  for (int64_t k = 0; k < (int64_t)g->nx * g->ny; ++k) rho[k] += n[k] * q;
}

/* ---- generalized_poisson.jl:34-68 dense assembly, :286-324 periodic, :205-215 dirichlet --- */
void orc_poisson_assemble(const orc_grid *g, double *A /* nn*nn col-major, zeroed here */) {
  const int nx = g->nx, ny = g->ny; const int64_t nn = (int64_t)nx * ny;
  memset(A, 0, sizeof(double) * (size_t)(nn * nn));
#define AT(r, c) A[(r) + (c) * nn]
  for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
    const int64_t r = i + (int64_t)j * nx;
    if (i < nx - 1) { AT(r, r) -= 1.0; AT(r, r + 1) += 1.0; }
    if (i > 0)      { AT(r, r) -= 1.0; AT(r, r - 1) += 1.0; }
    if (j < ny - 1) { AT(r, r) -= 1.0; AT(r, r + nx) += 1.0; }
    if (j > 0)      { AT(r, r) -= 1.0; AT(r, r - nx) += 1.0; }
  }
  const double d2 = g->dx * g->dx;                 /* :65 A ./= dh[1]^2 */
  for (int64_t k = 0; k < nn * nn; ++k) A[k] /= d2;
}

void orc_poisson_apply_periodic(const orc_grid *g, double *A, int axis) {
  const int nx = g->nx, ny = g->ny; const int64_t nn = (int64_t)nx * ny;
  if (axis == 1) {                                 /* couples j = 1 <-> j = ny, uses dx^2 */
    const double c = (0.5 * 1.0 + 0.5 * 1.0) / (g->dx * g->dx);
    for (int jj = 0; jj < 2; ++jj) {               /* `for j in (1, ny)` */
      const int j = jj == 0 ? 0 : ny - 1;
      for (int i = 0; i < nx; ++i) {
        const int64_t r = i + (int64_t)j * nx;
        if (j == ny - 1) { AT(r, r) -= c; AT(r, (int64_t)i) += c; }
        if (j == 0)      { AT(r, r) -= c; AT(r, i + (int64_t)(ny - 1) * nx) += c; }
      }
    }
  }
  if (axis == 2) {                                 /* couples i = 1 <-> i = nx, uses dy^2 */
    const double c = (0.5 * 1.0 + 0.5 * 1.0) / (g->dy * g->dy);
    for (int j = 0; j < ny; ++j) {
      for (int ii = 0; ii < 2; ++ii) {
        const int i = ii == 0 ? 0 : nx - 1;
        const int64_t r = i + (int64_t)j * nx;
        if (i == nx - 1) { AT(r, r) -= c; AT(r, 0 + (int64_t)j * nx) += c; }
        if (i == 0)      { AT(r, r) -= c; AT(r, (nx - 1) + (int64_t)j * nx) += c; }
      }
    }
  }
}

/* mask: nx*ny bytes; rho_dof: nx*ny bytes (1 = node still receives f/eps0), b: rhs */
void orc_poisson_apply_dirichlet(const orc_grid *g, double *A, double *b, uint8_t *rho_dof,
                                 const uint8_t *mask, double phi0) {
  const int64_t nn = (int64_t)g->nx * g->ny;
  for (int64_t r = 0; r < nn; ++r) if (mask[r]) {
    for (int64_t c = 0; c < nn; ++c) AT(r, c) = 0.0;
    AT(r, r) = 1.0;
    b[r] = phi0;
    rho_dof[r] = 0;
  }
}
#undef AT

/* ---- :201-203 solve(A,b) = A\b : dense LU, partial pivoting (dgetrf/dgetrs class) --------- */
/* A is copied (the reference refactors every step).  Returns 0, or k+1 if pivot k is exactly 0. */
int orc_dense_solve(const double *A, const double *b, int64_t n, double *x) {
  double *M = (double *)malloc(sizeof(double) * (size_t)(n * n));
  if (!M) return -1;
  memcpy(M, A, sizeof(double) * (size_t)(n * n));
  memcpy(x, b, sizeof(double) * (size_t)n);
  int rc = 0;
  for (int64_t k = 0; k < n; ++k) {
    int64_t piv = k; double best = fabs(M[k + k * n]);
    for (int64_t r = k + 1; r < n; ++r) { double a = fabs(M[r + k * n]); if (a > best) { best = a; piv = r; } }
    if (best == 0.0) { rc = (int)(k + 1); break; }
    if (piv != k) {
      for (int64_t c = 0; c < n; ++c) { double t = M[k + c * n]; M[k + c * n] = M[piv + c * n]; M[piv + c * n] = t; }
      double t = x[k]; x[k] = x[piv]; x[piv] = t;
    }
    const double inv = 1.0 / M[k + k * n];
    for (int64_t r = k + 1; r < n; ++r) M[r + k * n] *= inv;
    for (int64_t c = k + 1; c < n; ++c) {
      const double ukc = M[k + c * n];
      if (ukc != 0.0) for (int64_t r = k + 1; r < n; ++r) M[r + c * n] -= M[r + k * n] * ukc;
    }
    for (int64_t r = k + 1; r < n; ++r) x[r] -= M[r + k * n] * x[k];
  }
  if (!rc) for (int64_t k = n - 1; k >= 0; --k) {
    x[k] /= M[k + k * n];
    for (int64_t r = 0; r < k; ++r) x[r] -= M[r + k * n] * x[k];
  }
  free(M);
  return rc;
}

/* :372-378 calculate_electric_potential: b[rho] = f[rho]/eps0 with f = -rho (ParticleInCell.jl:126) */
int orc_electric_potential(const double *A, double *b, const uint8_t *rho_dof, const double *rho,
                           double eps0, int64_t nn, double *phi) {
  for (int64_t r = 0; r < nn; ++r) if (rho_dof[r]) b[r] = (-rho[r]) / eps0;
  return orc_dense_solve(A, b, nn, phi);
}

/* ---- :398-410 calculate_electric_field! ---------------------------------------------------- */
void orc_electric_field(const orc_grid *g, const double *phi, double *E /* nx*ny*3 */) {
  const int nx = g->nx, ny = g->ny; const int64_t nn = (int64_t)nx * ny;
  const double dx = g->dx, dy = g->dy;
  memset(E, 0, sizeof(double) * (size_t)(3 * nn));
#define P(i, j) phi[(i) + (int64_t)(j) * nx]
  for (int j = 0; j < ny; ++j) {
    for (int i = 1; i < nx - 1; ++i) E[i + (int64_t)j * nx] = (P(i - 1, j) - P(i + 1, j)) / (2 * dx);
    E[0 + (int64_t)j * nx] = (P(0, j) - P(1, j)) / dx;
    E[(nx - 1) + (int64_t)j * nx] = (P(nx - 2, j) - P(nx - 1, j)) / dx;
  }
  double *Ey = E + nn;
  for (int i = 0; i < nx; ++i) {
    for (int j = 1; j < ny - 1; ++j) Ey[i + (int64_t)j * nx] = (P(i, j - 1) - P(i, j + 1)) / (2 * dy);
    Ey[i] = (P(i, 0) - P(i, 1)) / dy;
    Ey[i + (int64_t)(ny - 1) * nx] = (P(i, ny - 2) - P(i, ny - 1)) / dy;
  }
#undef P
}

/* ---- cross_section.jl:8-14: LinearInterpolation(xs, ys; extrapolation_bc = Flat()) -------- */
double orc_xsec_eval(const double *xs, const double *ys, int32_t n, double x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int lo = 0, hi = n - 1;                          /* xs[lo] <= x < xs[hi] */
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (xs[mid] <= x) lo = mid; else hi = mid; }
  const double f = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
  return (1.0 - f) * ys[lo] + f * ys[lo + 1];
}

/* ---- RNG for the RNG-dependent parts (stream is NOT the reference's; statistical parity) -- */
typedef struct { uint64_t s[4]; int has_spare; double spare; } orc_rng;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix(uint64_t *z) { uint64_t r = (*z += 0x9e3779b97f4a7c15ULL);
  r = (r ^ (r >> 30)) * 0xbf58476d1ce4e5b9ULL; r = (r ^ (r >> 27)) * 0x94d049bb133111ebULL; return r ^ (r >> 31); }
void orc_rng_seed(orc_rng *r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&seed); r->has_spare = 0; }
static inline uint64_t rng_next(orc_rng *r) {      /* xoshiro256++ */
  uint64_t *s = r->s; const uint64_t res = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return res; }
static inline double rng_u01(orc_rng *r) { return (double)(rng_next(r) >> 11) * 0x1.0p-53; }
static double rng_randn(orc_rng *r) {
  if (r->has_spare) { r->has_spare = 0; return r->spare; }
  double u, v, s;
  do { u = 2.0 * rng_u01(r) - 1.0; v = 2.0 * rng_u01(r) - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
  const double f = sqrt(-2.0 * log(s) / s);
  r->spare = v * f; r->has_spare = 1; return u * f;
}

/* ---- sources.jl:24-34 sample! (Maxwellian load) -------------------------------------------- */
int64_t orc_sample(orc_species *s, double rate, double dt, const double *wx, const double *dx0,
                   const double *wv, const double *dv, orc_rng *rng) {
  int64_t n = (int64_t)floor(rate * dt);
  if (n > s->cap - s->np) n = s->cap - s->np;
  const int64_t b = s->np;
  /* Julia fills rand(n,D) column by column */
  for (int64_t p = 0; p < n; ++p) s->x[b + p] = rng_u01(rng) * wx[0] + dx0[0];
  for (int64_t p = 0; p < n; ++p) s->y[b + p] = rng_u01(rng) * wx[1] + dx0[1];
  for (int64_t p = 0; p < n; ++p) s->vx[b + p] = rng_randn(rng) * wv[0] + dv[0];
  for (int64_t p = 0; p < n; ++p) s->vy[b + p] = rng_randn(rng) * wv[1] + dv[1];
  for (int64_t p = 0; p < n; ++p) s->vz[b + p] = rng_randn(rng) * wv[2] + dv[2];
  s->np += n;
  return n;
}

/* ---- Chemistry/src/mcc.jl -------------------------------------------------------------------- */
enum { ORC_ELASTIC_ISOTROPIC = 0, ORC_ELASTIC_BACKWARD = 1, ORC_INELASTIC_BACKWARD = 2,
       ORC_EXCITATION = 3, ORC_IONIZATION = 4 };

typedef struct {
  int32_t kind;               /* mcc.jl:4-8 */
  double energy;              /* threshold [eV] for excitation / ionization */
  const double *eps, *sigma;  /* CrossSection nodes, cross_section.jl:3-6 */
  int32_t n_nodes;
  orc_species *product;       /* ion product of an ionization (products != source), or NULL */
} orc_collision;

typedef struct {
  orc_collision *coll; int32_t N;
  orc_species *source;
  double tq, tm, tT;          /* target FluidSpecies q, m, T (fluid.jl:1-8) */
  const double *tn;           /* target density n (nx*ny) */
  double max_sigma_g, m_eV, remainder;   /* mcc.jl:18-24 */
} orc_mcc;

#define ORC_QE_MCC 1.60217646e-19      /* mcc.jl:26 */
#define ORC_KB_MCC 1.3806503e-23       /* mcc.jl:75 */

static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

/* mcc.jl:27-51 MonteCarloCollisions ctor: max over union grid of sum sigma_i(eps) * sqrt(2 eps/m) */
void orc_mcc_setup(orc_mcc *m) {
  int64_t tot = 1;
  for (int k = 0; k < m->N; ++k) tot += m->coll[k].n_nodes;
  double *e = (double *)malloc(sizeof(double) * (size_t)tot);
  int64_t c = 0; e[c++] = 0.0;
  for (int k = 0; k < m->N; ++k) for (int q = 0; q < m->coll[k].n_nodes; ++q) e[c++] = m->coll[k].eps[q];
  qsort(e, (size_t)tot, sizeof(double), cmp_d);
  m->m_eV = m->source->m / ORC_QE_MCC;
  const double alpha = sqrt(2.0 / m->m_eV);
  double best = -INFINITY;
  for (int64_t q = 0; q < tot; ++q) {
    if (q > 0 && e[q] == e[q - 1]) continue;
    const double v = alpha * sqrt(e[q]);
    double sg = 0.0;
    for (int k = 0; k < m->N; ++k) sg += orc_xsec_eval(m->coll[k].eps, m->coll[k].sigma, m->coll[k].n_nodes, e[q]) * v;
    if (sg > best) best = sg;
  }
  free(e);
  m->max_sigma_g = best; m->remainder = 0.0;
}

static inline double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* mcc.jl:89-106 euler_angles, :83-87 unrotated, row-vector * T (:112,:126) */
static void scatter3(const double *v, double sc, double cc, double se, double ce, double *out) {
  const double nv = norm3(v);
  const double ct = v[2] / nv;
  const double st = sqrt(1.0 - ct * ct);
  double cp, sp;
  if (st == 0.0) { cp = 1.0; sp = 0.0; } else { cp = v[0] / nv / st; sp = v[1] / nv / st; }
  const double T[3][3] = { { cp * ct, -sp * ct, -st }, { -sp, cp, 0.0 }, { cp * st, sp * st, ct } };
  const double r[3] = { sc * ce, sc * se, cc };
  for (int c = 0; c < 3; ++c) out[c] = (r[0] * T[0][c] + r[1] * T[1][c]) + r[2] * T[2][c];
}
static void isotropic_scattering(const double *v, orc_rng *g, double *out) {      /* :53-62,:108-113 */
  const double chi = 2 * M_PI * rng_u01(g); const double sc = sin(chi), cc = cos(chi);
  const double eta = 2 * M_PI * rng_u01(g);
  scatter3(v, sc, cc, sin(eta), cos(eta), out);
}
static void diffuse_reflection(const double *v, orc_rng *g, double *out) {        /* :64-72,:122-127 */
  const double sc = sqrt(rng_u01(g)); const double cc = -sqrt(1.0 - sc * sc);
  const double eta = 2 * M_PI * rng_u01(g);
  scatter3(v, sc, cc, sin(eta), cos(eta), out);
}

/* mcc.jl:129-229.  Returns 0 ok, 1 capacity overflow (reference: BoundsError). */
static int collide1(orc_mcc *m, orc_collision *c, int64_t p, orc_rng *g) {
  orc_species *s = m->source;
  double sv[3] = { s->vx[p], s->vy[p], s->vz[p] };
  if (c->kind <= ORC_INELASTIC_BACKWARD) {
    const double mr1 = s->m / (s->m + m->tm), mr2 = m->tm / (s->m + m->tm);
    const double vth = sqrt(2 * ORC_KB_MCC * m->tT / m->tm);
    double tv[3]; for (int k = 0; k < 3; ++k) tv[k] = rng_randn(g) * vth;
    double vr[3], w[3], dir[3];
    for (int k = 0; k < 3; ++k) { vr[k] = sv[k] - tv[k]; w[k] = mr1 * sv[k] + mr2 * tv[k]; }
    double mag = norm3(vr);
    if (c->kind == ORC_ELASTIC_ISOTROPIC) isotropic_scattering(vr, g, dir);
    else if (c->kind == ORC_ELASTIC_BACKWARD) diffuse_reflection(vr, g, dir);
    else { mag = rng_u01(g) * mag; diffuse_reflection(vr, g, dir); }
    s->vx[p] = w[0] + mr2 * (mag * dir[0]);
    s->vy[p] = w[1] + mr2 * (mag * dir[1]);
    s->vz[p] = w[2] + mr2 * (mag * dir[2]);
    return 0;
  }
  const double ms = m->m_eV;
  const double sE = 0.5 * ms * ((sv[0] * sv[0] + sv[1] * sv[1]) + sv[2] * sv[2]) - c->energy;
  if (sE < 0) return 0;
  if (c->kind == ORC_EXCITATION) {
    const double ev = sqrt(2 / ms) * sqrt(sE); double dir[3];
    isotropic_scattering(sv, g, dir);
    s->vx[p] = ev * dir[0]; s->vy[p] = ev * dir[1]; s->vz[p] = ev * dir[2];
    return 0;
  }
  const double e1E = sE * rng_u01(g), e2E = sE - e1E;
  const double alpha = sqrt(2.0 / ms), e1v = alpha * sqrt(e1E), e2v = alpha * sqrt(e2E);
  double d1[3], d2[3];
  diffuse_reflection(sv, g, d1);
  for (int k = 0; k < 3; ++k) sv[k] = e1v * d1[k];
  s->vx[p] = sv[0]; s->vy[p] = sv[1]; s->vz[p] = sv[2];
  if (s->np >= s->cap) return 1;
  diffuse_reflection(sv, g, d2);
  const int64_t q = s->np;
  s->x[q] = s->x[p]; s->y[q] = s->y[p];
  s->vx[q] = e2v * d2[0]; s->vy[q] = e2v * d2[1]; s->vz[q] = e2v * d2[2];
  s->np += 1;
  const double vth = sqrt(2 * ORC_KB_MCC * m->tT / m->tm);
  double tv[3]; for (int k = 0; k < 3; ++k) tv[k] = rng_randn(g) * vth;
  if (c->product && c->product != s) {
    orc_species *pr = c->product;
    const int64_t cnt = (int64_t)rint(s->w0 / pr->w0);
    for (int64_t k = 0; k < cnt; ++k) {
      if (pr->np >= pr->cap) return 1;
      const int64_t r = pr->np;
      pr->x[r] = s->x[p]; pr->y[r] = s->y[p];
      pr->vx[r] = tv[0]; pr->vy[r] = tv[1]; pr->vz[r] = tv[2];
      pr->np += 1;
    }
  }
  return 0;
}

/* PIC.perform!(mcc, E, dt, config)  mcc.jl:231-289.
 * nu: nx*ny*N counters (may be NULL).  out[0] = Nc, out[1] = collisions.
 * Returns 0 ok, 1 capacity, 2 max_Pt > 1/N (:244-246), 3 Pk > 1 (:273-279). */
int orc_mcc_perform(orc_mcc *m, const orc_grid *g, const double *E, double dt, double *nu,
                    int64_t *out, orc_rng *rng) {
  const int64_t nx = g->nx, nn = (int64_t)g->nx * g->ny;
  const int N = m->N;
  orc_species *s = m->source;
  if (nu) memset(nu, 0, sizeof(double) * (size_t)(nn * N));
  double max_n0 = -INFINITY;
  for (int64_t k = 0; k < nn; ++k) if (m->tn[k] > max_n0) max_n0 = m->tn[k];
  const double max_Pt = 1.0 - exp(-max_n0 * m->max_sigma_g * dt);      /* :243 */
  if (max_Pt > 1.0 / N) return 2;
  double ip; m->remainder = modf(N * max_Pt * (double)s->np + m->remainder, &ip);   /* :248 */
  const int64_t Nc = (int64_t)ip;
  int64_t ncoll = 0;
  const double tqm = m->tq / m->tm;
  for (int64_t it = 0; it < Nc; ++it) {
    const int64_t p = (int64_t)(rng_u01(rng) * (double)s->np);          /* :251 rand(1:np) */
    int64_t i, j; double hx, hy;
    cell1(s->x[p], g->dx, &i, &hx); cell1(s->y[p], g->dy, &j, &hy);
    const int64_t n00 = (i - 1) + (j - 1) * nx;
    const double n = m->tn[n00];
    if (n < 0) continue;
    const double U = rng_u01(rng);
    const int k = (int)floor(N * U + 1);                                /* :261 */
    orc_collision *c = &m->coll[k - 1];
    double d[3];
    d[0] = (tqm * E[n00]) * dt - s->vx[p];                              /* :266-267 */
    d[1] = (tqm * E[n00 + nn]) * dt - s->vy[p];
    d[2] = (tqm * E[n00 + 2 * nn]) * dt - s->vz[p];
    const double gg = norm3(d);
    const double eps = 0.5 * m->m_eV * (gg * gg);                       /* :268 */
    const double skg = orc_xsec_eval(c->eps, c->sigma, c->n_nodes, eps) * gg;
    double Pk = 1.0 - exp(-n * skg * dt);                               /* :271 */
    Pk /= N * max_Pt;                                                   /* :272 */
    if (Pk > 1.0) return 3;
    if (U > (double)k / N - Pk) {                                       /* :281 */
      if (collide1(m, c, p, rng)) return 1;
      if (nu) nu[n00 + (int64_t)(k - 1) * nn] += 1;
      ++ncoll;
    }
  }
  if (out) { out[0] = Nc; out[1] = ncoll; }
  return 0;
}

/* ---- ParticleInCell.jl:51-72 advance! for one species (gather, push, after_push) ---------- */
/* bmode[d]: 0 none, 1 wrap, 2 discard -- the after_push hooks of the BASELINE problem scripts
 * (10_two_streams.jl:56-58, 11_rf_discharge.jl:80-83, 12_avalanche.jl:55-59): discards first. */
void orc_advance(orc_species *s, const orc_grid *g, const double *E, double dt, const int32_t *bmode) {
  const double qm = s->q / s->m;
  for (int64_t p = 0; p < s->np; ++p) {
    double e3[3];
    gather1(g, E, s->x[p], s->y[p], e3);
    push1(&s->x[p], &s->y[p], &s->vx[p], &s->vy[p], &s->vz[p], e3, qm, dt);
  }
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 2) orc_discard(s, g, d);
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 1) orc_wrap(s, g, d);
}

/* ---- OpenMP variants: multi-core CPU baseline timing only ---------------------------------- */
void orc_advance_mt(orc_species *s, const orc_grid *g, const double *E, double dt, const int32_t *bmode) {
  const double qm = s->q / s->m;
  const int64_t np = s->np;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < np; ++p) {
    double e3[3];
    gather1(g, E, s->x[p], s->y[p], e3);
    push1(&s->x[p], &s->y[p], &s->vx[p], &s->vy[p], &s->vz[p], e3, qm, dt);
  }
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 2) orc_discard(s, g, d);
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 1) {
    double *px = d == 1 ? s->x : s->y;
    const double L = (double)((d == 1 ? g->nx : g->ny) - 1) * (d == 1 ? g->dx : g->dy);
    const double o = d == 1 ? g->ox : g->oy;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < s->np; ++p) { double a = jl_fld(px[p] - o, L); if (a != 0) px[p] -= a * L; }
  }
}

void orc_deposit_mt(const orc_grid *g, const orc_species *s, double *u) {
  const int64_t nx = g->nx, nn = (int64_t)g->nx * g->ny;
  memset(u, 0, sizeof(double) * (size_t)nn);
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < s->np; ++p) {
    int64_t i, j; double hx, hy;
    cell1(s->x[p], g->dx, &i, &hx); cell1(s->y[p], g->dy, &j, &hy);
    const int64_t n00 = (i - 1) + (j - 1) * nx;
    const double w = s->wg[p];
    const double c00 = (1.0 - hx) * (1.0 - hy) * w, c10 = hx * (1.0 - hy) * w;
    const double c01 = (1.0 - hx) * hy * w, c11 = hx * hy * w;
#pragma omp atomic
    u[n00] += c00;
#pragma omp atomic
    u[n00 + 1] += c10;
#pragma omp atomic
    u[n00 + nx] += c01;
#pragma omp atomic
    u[n00 + nx + 1] += c11;
  }
}

/* ==== SURVEY.md 8f row N1: surface tracker (ParticleInCell/src/pic/surfaces/{build,track,check,hit}.jl) ==================
 * The reference's Dict{(cell,cell) -> Surface} is flattened into a table over the cells
 * 0..nx x 0..ny with four directed faces per cell  0:(i,j-1) 1:(i+1,j) 2:(i,j+1) 3:(i-1,j)
 * holding a surface id (0 = no key).  Kinds: 0 periodic (no-op hit!), 1 absorbing, 2 reflective,
 * 3 fixed electrode, 4 floating electrode.  The FIFO of check! (check.jl:48-62) is kept literally:
 * a ring buffer of tuples, popfirst!/push!, so electrode charge sums run in the reference's order. */
typedef struct {
  int32_t nx, ny;          /* nodes */
  double dh;               /* st.dh  build.jl:17 */
  uint8_t *face;           /* 4*(nx+1)*(ny+1) */
  uint8_t *tracked;        /* (nx+1)*(ny+1): Base.in(::BoundaryCell, st)  build.jl:86-93 */
  int32_t *kind;           /* per surface id */
  double *area;
  double *dq;              /* s.dq  circuit_coupling.jl:49-50 */
} orc_tracker;

typedef struct { double dt; int64_t p; int32_t i, j; double hx, hy; } orc_tp;   /* TrackedParticle{2}  build.jl:5 */

static inline int trk_dir(int32_t i, int32_t j, int32_t k, int32_t l) {
  if (k == i && l == j - 1) return 0;
  if (k == i + 1 && l == j) return 1;
  if (k == i && l == j + 1) return 2;
  return 3;
}

/* check(pt::TrackedParticle{2}, pv, dh)  check.jl:17-36 */
static orc_tp trk_check1(orc_tp pt, const orc_species *s, double dh) {
  double vx = s->vx[pt.p], vy = s->vy[pt.p];
  double dx = (vx > 0) ? dh * (1 - pt.hx) : dh * pt.hx;
  double dy = (vy > 0) ? dh * (1 - pt.hy) : dh * pt.hy;
  double dtx = dx / fabs(vx), dty = dy / fabs(vy);
  if (pt.dt < dtx && pt.dt < dty) return pt;
  orc_tp o = pt;
  if (dtx < dty) {
    o.dt = pt.dt - dtx;
    o.hy = pt.hy + vy * dtx / dh;
    if (vx > 0) { o.i = pt.i + 1; o.hx = 0.; } else { o.i = pt.i - 1; o.hx = 1.; }
  } else {
    o.dt = pt.dt - dty;
    o.hx = pt.hx + vx * dty / dh;
    if (vy > 0) { o.j = pt.j + 1; o.hy = 0.; } else { o.j = pt.j - 1; o.hy = 1.; }
  }
  return o;
}

static int cmp_i64_desc(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x < y) - (x > y);
}

/* track! (track.jl:42-52) -> push (caller) -> check! (check.jl:39-68).  orc_track fills `queue`
 * (capacity >= np) and returns the number of tracked particles; orc_check consumes it and returns the
 * number absorbed; *too_fast is the condition of the printed message (:41-46). */
int64_t orc_track(const orc_tracker *t, const orc_species *s, double dt, orc_tp *queue) {
  int64_t n = 0;
  for (int64_t p = 0; p < s->np; ++p) {
    int64_t i, j;
    double hx, hy;
    cell1(s->x[p], t->dh, &i, &hx);      /* particle_cell(px, p, st.dh): scalar dh for both axes */
    cell1(s->y[p], t->dh, &j, &hy);
    if (i < 0 || i > t->nx || j < 0 || j > t->ny) continue;
    if (!t->tracked[i + j * (t->nx + 1)]) continue;
    orc_tp pt = {dt, p, (int32_t)i, (int32_t)j, hx, hy};
    queue[n++] = pt;
  }
  return n;
}

int64_t orc_check(orc_tracker *t, orc_species *s, double dt, orc_tp *queue, int64_t n_tracked, int64_t qcap,
                  int32_t *too_fast) {
  double vmax = t->dh / dt;
  int tf = 0;
  for (int64_t p = 0; p < s->np; ++p)
    if (fabs(s->vx[p]) > vmax || fabs(s->vy[p]) > vmax || fabs(s->vz[p]) > vmax) tf = 1;
  if (too_fast) *too_fast = tf;
  int64_t *absorbed = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_tracked > 0 ? n_tracked : 1));
  int64_t nabs = 0, head = 0, tail = n_tracked % qcap, count = n_tracked;
  while (count > 0) {
    orc_tp pt = queue[head];                         /* popfirst! */
    head = (head + 1) % qcap;
    --count;
    orc_tp p2 = trk_check1(pt, s, t->dh);
    if (p2.i == pt.i && p2.j == pt.j) continue;
    int sid = 0;
    if (pt.i >= 0 && pt.i <= t->nx && pt.j >= 0 && pt.j <= t->ny)
      sid = t->face[4 * (pt.i + pt.j * (t->nx + 1)) + trk_dir(pt.i, pt.j, p2.i, p2.j)];
    if (!sid) {                                      /* track!(st, pt') */
      queue[tail] = p2; tail = (tail + 1) % qcap; ++count;
      continue;
    }
    int kind = t->kind[sid];
    if (kind == 1 || kind == 3) {                    /* absorbed! */
      absorbed[nabs++] = pt.p;
    } else if (kind == 4) {                          /* circuit_coupling.jl:44-53 (sigma update: quirk S1, no effect) */
      t->dq[sid] += s->q * s->wg[pt.p];
      absorbed[nabs++] = pt.p;
    } else if (kind == 2) {                          /* hit.jl:39-56 */
      int64_t p = pt.p;
      s->x[p] -= s->vx[p] * p2.dt;
      s->y[p] -= s->vy[p] * p2.dt;
      if (p2.i != pt.i) s->vx[p] *= -1;
      if (p2.j != pt.j) s->vy[p] *= -1;
      s->x[p] += s->vx[p] * p2.dt;
      s->y[p] += s->vy[p] * p2.dt;
      orc_tp p3 = p2;                                /* scattered!  hit.jl:12-20 */
      if (p2.hx == 0.) { p3.i = p2.i - 1; p3.hx = 1.; }
      if (p2.hx == 1.) { p3.i = p2.i + 1; p3.hx = 0.; }
      if (p2.hy == 0.) { p3.j = p2.j - 1; p3.hy = 1.; }
      if (p2.hy == 1.) { p3.j = p2.j + 1; p3.hy = 0.; }
      queue[tail] = p3; tail = (tail + 1) % qcap; ++count;
    }                                                /* kind 0: generic no-op hit! */
  }
  qsort(absorbed, (size_t)nabs, sizeof(int64_t), cmp_i64_desc);   /* SortedSet(Reverse) */
  for (int64_t k = 0; k < nabs; ++k)
    if (k == 0 || absorbed[k] != absorbed[k - 1]) remove1(s, absorbed[k]);
  free(absorbed);
  return nabs;
}

/* advance!(part, E, B, dt, config) with config.tracker  ParticleInCell.jl:51-61 */
int64_t orc_advance_tracked(orc_species *s, const orc_grid *g, orc_tracker *t, const double *E, double dt,
                            const int32_t *bmode, orc_tp *queue, int64_t qcap, int32_t *too_fast) {
  int64_t n = orc_track(t, s, dt, queue);
  double *pE = (double *)malloc(sizeof(double) * 3 * (size_t)(s->np > 0 ? s->np : 1));
  orc_gather(g, s, E, pE);
  orc_push(s, pE, dt);
  free(pE);
  int64_t nabs = orc_check(t, s, dt, queue, n, qcap, too_fast);
  for (int d = 0; d < 2; ++d) if (bmode[d] == 2) orc_discard(s, g, d + 1);
  for (int d = 0; d < 2; ++d) if (bmode[d] == 1) orc_wrap(s, g, d + 1);
  return nabs;
}

/* ==== SURVEY.md 8f row N3: axisymmetric r-z pusher (ParticleInCell/src/pic/pushers.jl:13-17, 52-66) ========= */
static inline void to_cylindrical1(double *x, double *vx, double *vz, double dt) {
  double y = dt * *vz;                               /* :54 */
  double r = sqrt(*x * *x + y * y);                  /* :55 */
  double sn = y / r;                                 /* :57 */
  if (r == 0.0) sn = 0.0;                            /* :58  r .~ 0.0 is an exact zero test */
  double cs = sqrt(1.0 - sn * sn);                   /* :59 */
  double vr = cs * *vx + sn * *vz;                   /* :61 */
  double vy = -sn * *vx + cs * *vz;                  /* :62 */
  *x = r; *vx = vr; *vz = vy;                        /* :63-65 */
}

void orc_to_cylindrical(orc_species *s, double dt) {
  for (int64_t p = 0; p < s->np; ++p) to_cylindrical1(&s->x[p], &s->vx[p], &s->vz[p], dt);
}

/* advance! with BorisPusher{:rz}: gather, push_in_cartesian!, transform, after_push */
void orc_advance_rz(orc_species *s, const orc_grid *g, const double *E, double dt, const int32_t *bmode) {
  const double qm = s->q / s->m;
  for (int64_t p = 0; p < s->np; ++p) {
    double e3[3];
    gather1(g, E, s->x[p], s->y[p], e3);
    push1(&s->x[p], &s->y[p], &s->vx[p], &s->vy[p], &s->vz[p], e3, qm, dt);
    to_cylindrical1(&s->x[p], &s->vx[p], &s->vz[p], dt);
  }
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 2) orc_discard(s, g, d);
  for (int d = 1; d <= 2; ++d) if (bmode[d - 1] == 1) orc_wrap(s, g, d);
}

/* ==== SURVEY.md 8f row N4: DSMC (Chemistry/src/dsmc.jl:25-142), one DSMC.ElasticCollision ====================
 * Same restatement as oracle/dsmc_oracle.py (quirks D1-D5 there); RNG = this file's xoshiro256++, so only the
 * stream-independent parts (candidate pairs per cell, carry) agree exactly with the Python oracle.
 * work: int64[2*(nn+1) + np_s + np_t] scratch for the cell lists (cache!, :25-30). */
static void dsmc_lists(const orc_grid *g, const orc_species *s, int64_t *start /* nn+1 */, int64_t *list) {
  const int64_t nn = (int64_t)g->nx * g->ny;
  memset(start, 0, sizeof(int64_t) * (size_t)(nn + 1));
  for (int64_t p = 0; p < s->np; ++p) {
    int64_t i, j; double hx, hy;
    cell1(s->x[p], g->dx, &i, &hx);
    cell1(s->y[p], g->dy, &j, &hy);
    start[(i - 1) + (j - 1) * g->nx + 1]++;
  }
  for (int64_t c = 0; c < nn; ++c) start[c + 1] += start[c];
  int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)nn);
  memcpy(cur, start, sizeof(int64_t) * (size_t)nn);
  for (int64_t p = 0; p < s->np; ++p) {
    int64_t i, j; double hx, hy;
    cell1(s->x[p], g->dx, &i, &hx);
    cell1(s->y[p], g->dy, &j, &hy);
    list[cur[(i - 1) + (j - 1) * g->nx]++] = p;      /* push!(candidates[i, j], p): rows in increasing order */
  }
  free(cur);
}

int64_t orc_dsmc_perform(orc_species *src, orc_species *tgt, const orc_grid *g, const double *gn, const double *sg,
                         int32_t n_nodes, double dt, double *remaining /* nx*ny */, double *nu /* nx*ny or NULL */,
                         int64_t *n_candidates, orc_rng *rng) {
  const int64_t nn = (int64_t)g->nx * g->ny;
  const int same = src == tgt;
  int64_t *ss = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nn + 1)), *ls = (int64_t *)malloc(sizeof(int64_t) * (size_t)(src->np + 1));
  int64_t *st = ss, *lt = ls;
  dsmc_lists(g, src, ss, ls);
  if (!same) {
    st = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nn + 1));
    lt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(tgt->np + 1));
    dsmc_lists(g, tgt, st, lt);
  }
  int kmax = 0;
  for (int k = 1; k < n_nodes; ++k) if (sg[k] > sg[kmax]) kmax = k;
  const double sgmax = sg[kmax] * gn[kmax];                                   /* :107 */
  const double Wa = src->w0, Wb = tgt->w0;
  double Pab, Pba;
  if (Wa > Wb) { Pab = Wb / Wa; Pba = 1.0; } else { Pab = 1.0; Pba = Wa / Wb; }   /* :101-105 */
  const double mr1 = src->m / (src->m + tgt->m), mr2 = tgt->m / (src->m + tgt->m);
  int64_t ncoll = 0, ncand = 0;
  for (int64_t i = 0; i < g->nx; ++i)                                         /* :108-109 for i, for j */
    for (int64_t j = 0; j < g->ny; ++j) {
      const int64_t c = i + j * g->nx;
      const int64_t Na = ss[c + 1] - ss[c], Nb = st[c + 1] - st[c];
      if (nu) nu[c] = 0.0;
      if (Na < 2 || Nb < 2) continue;
      const double na = (double)Na * Wa / (g->dx * g->dy);
      double Nc = na * (double)Nb * dt * sgmax;
      Nc /= Pab + (Wb / Wa) * Pba;
      if (!same) Nc *= 2;
      Nc += remaining[c];
      const double fl = floor(Nc);
      remaining[c] = Nc - fl;
      ncand += (int64_t)fl;
      for (int64_t it = 0; it < (int64_t)fl; ++it) {
        const int64_t s = ls[ss[c] + (int64_t)(rng_u01(rng) * (double)Na)];
        int64_t t = lt[st[c] + (int64_t)(rng_u01(rng) * (double)Nb)];
        while (!same && s == t) t = lt[st[c] + (int64_t)(rng_u01(rng) * (double)Nb)];      /* D3 */
        const double gx = src->vx[s] - tgt->vx[t], gy = src->vy[s] - tgt->vy[t], gz = src->vz[s] - tgt->vz[t];
        const double gg = sqrt(gx * gx + gy * gy + gz * gz);
        const double sgg = orc_xsec_eval(gn, sg, n_nodes, gg) * gg;
        if (sgg / sgmax < rng_u01(rng)) continue;
        const double cmx = mr1 * src->vx[s] + mr2 * tgt->vx[t], cmy = mr1 * src->vy[s] + mr2 * tgt->vy[t],
                     cmz = mr1 * src->vz[s] + mr2 * tgt->vz[t];
        const double B = 2 * rng_u01(rng) - 1.0, A = sqrt(1 - B * B), C = 2 * M_PI * rng_u01(rng);
        const double rx = gg * B, ry = gg * (A * cos(C)), rz = gg * (A * sin(C));
        if (src->wg[s] == tgt->wg[t]) {
          src->vx[s] = cmx + mr2 * rx; src->vy[s] = cmy + mr2 * ry; src->vz[s] = cmz + mr2 * rz;
          tgt->vx[t] = cmx - mr1 * rx; tgt->vy[t] = cmy - mr1 * ry; tgt->vz[t] = cmz - mr1 * rz;
        } else {
          const double Pab2 = tgt->wg[t] / src->wg[s], Pba2 = src->wg[s] / tgt->wg[t], R2 = rng_u01(rng);
          if (Pab2 > R2) { src->vx[s] = cmx + mr2 * rx; src->vy[s] = cmy + mr2 * ry; src->vz[s] = cmz + mr2 * rz; }
          if (Pba2 > R2 && s < tgt->np) { tgt->vx[s] = cmx - mr2 * rx; tgt->vy[s] = cmy - mr2 * ry; tgt->vz[s] = cmz - mr2 * rz; }   /* D2 */
        }
        if (nu) nu[c] += 1.0;
        ++ncoll;
      }
    }
  if (n_candidates) *n_candidates = ncand;
  free(ss); free(ls);
  if (!same) { free(st); free(lt); }
  return ncoll;
}
