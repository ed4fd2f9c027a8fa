// Per-particle device arithmetic of the hot path.  Every function keeps the reference's
// operation order and uses the *_rn intrinsics (never contracted into FMAs) so that results
// are bit-identical to the Julia code / the oracle.
#pragma once
#include "common.cuh"

__device__ __forceinline__ bool is_dead(double x) { return x != x; }   // dead slots carry x = NaN

// Correctly rounded x/d in three FP64 operations (Markstein 1990; Muller et al., Handbook of
// Floating-Point Arithmetic, thm. "correction of a faithful quotient"): with r = RN(1/d),
//   q0 = RN(x*r) ;  rem = x - q0*d  (exact in one FMA) ;  q = RN(q0 + rem*r) = RN(x/d).
// Valid away from over/underflow (guarded: otherwise the IEEE divide is used) and not for a
// divisor whose significand is all ones (the host clears fast_div then).  Exhaustively compared
// with exact rational arithmetic on the host and with __ddiv_rn in the GPU tests.
// __ddiv_rn costs 4.6 SM-cycles per warp on B200 (profiles/r1_microbench_warp_ops_b200.txt).
__device__ __forceinline__ double div_exact(double x, double d, double r, int fast) {
  const double ax = fabs(x);
  if (fast && ((ax > 1e-250 && ax < 1e250) || x == 0.0)) {
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-q0, d, x);
    return __fma_rn(rem, r, q0);
  }
  return __ddiv_rn(x, d);
}

// particle_cell(px, p, dh)  ParticleInCell.jl:28-35:  f = 1 + x/dh ; i = floor(f) ; h = f - i
// In cell1 the magnitude guards of div_exact are unnecessary: a quotient below 2^-53 cannot change
// f = 1 + x/d (both ways f == 1 and h == 0), and for |x| so large that q0 overflows the cell is
// out of the grid either way ((int) saturates identically).  So only the uniform `fast` flag is tested.
__device__ __forceinline__ void cell1(double x, double d, double rd, int fast, int &i, double &h) {
  double q;
  if (fast) {
    const double q0 = __dmul_rn(x, rd);
    q = __fma_rn(__fma_rn(-q0, d, x), rd, q0);
  } else {
    q = __ddiv_rn(x, d);
  }
  const double f = __dadd_rn(1.0, q);
  const double fl = floor(f);
  i = (int)fl;
  h = __dsub_rn(f, fl);
}
// the same with the three-operation division unconditionally (callers checked GridDev::fast_div on the host)
__device__ __forceinline__ void cell1_fast(double x, double d, double rd, int &i, double &h) {
  const double q0 = __dmul_rn(x, rd);
  const double f = __dadd_rn(1.0, __fma_rn(__fma_rn(-q0, d, x), rd, q0));
  const double fl = floor(f);
  i = (int)fl;
  h = __dsub_rn(f, fl);
}
__device__ __forceinline__ void cell1(double x, double d, int &i, double &h) {
  const double f = __dadd_rn(1.0, __ddiv_rn(x, d));
  const double fl = floor(f);
  i = (int)fl;
  h = __dsub_rn(f, fl);
}

struct CicW {
  double w00, w10, w01, w11;
};
// weights formed first: (1-hx)*(1-hy), hx*(1-hy), (1-hx)*hy, hx*hy  (cloud_in_cell.jl:11-14,29-32)
__device__ __forceinline__ CicW cic_weights(double hx, double hy) {
  const double ax = __dsub_rn(1.0, hx), ay = __dsub_rn(1.0, hy);
  CicW w;
  w.w00 = __dmul_rn(ax, ay);
  w.w10 = __dmul_rn(hx, ay);
  w.w01 = __dmul_rn(ax, hy);
  w.w11 = __dmul_rn(hx, hy);
  return w;
}

// grid_to_particle, cloud_in_cell.jl:29-32: left-to-right sum of weight*value
__device__ __forceinline__ double cic_gather(const CicW &w, double u00, double u10, double u01,
                                             double u11) {
  double s = __dadd_rn(__dmul_rn(w.w00, u00), __dmul_rn(w.w10, u10));
  s = __dadd_rn(s, __dmul_rn(w.w01, u01));
  return __dadd_rn(s, __dmul_rn(w.w11, u11));
}

// push_in_cartesian!  pushers.jl:37-50 with B == 0 (generalized_poisson.jl:412-419):
//   v- = ((0.5dt)*qm)*E + v ; v+ = v- (+0.0 twice) ; v = ((dt*E)*qm)*0.5 + v+ ; x = dt*v + x
__device__ __forceinline__ double push_v(double v, double e, double c1, double qm, double dt) {
  double vm = __dadd_rn(__dmul_rn(c1, e), v);
  vm = __dadd_rn(vm, 0.0);   // :44 v' = v- + v- x B ; :46 v+ = v- + v' x s   (B = s = 0)
  return __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(dt, e), qm), 0.5), vm);
}
// Same value with one multiplication less: ((dt*e)*qm)*0.5 == (dt*e)*(0.5*qm) bit for bit, because scaling by a power
// of two commutes with rounding (hqm = 0.5*qm is exact) -- except where the product is subnormal (|.| < 2^-1022),
// which no velocity increment of a plasma reaches.
__device__ __forceinline__ double push_v_h(double v, double e, double c1, double hqm, double dt) {
  double vm = __dadd_rn(__dmul_rn(c1, e), v);
  vm = __dadd_rn(vm, 0.0);
  return __dadd_rn(__dmul_rn(__dmul_rn(dt, e), hqm), vm);
}
__device__ __forceinline__ double push_x(double x, double v, double dt) {
  return __dadd_rn(__dmul_rn(dt, v), x);
}

// transform_from_cartesian_to_cylindrical!(part, dt)  pushers.jl:52-66 for one particle (BorisPusher{:rz}):
//   y = dt*vz ; r = sqrt(x^2 + y^2) ; sin = y/r (0 where r == 0: `r .~ 0.0` with the default tolerances is an
//   exact zero test) ; cos = sqrt(1 - sin^2) ; vr = cos*vx + sin*vz ; vy = -sin*vx + cos*vz ; x = r
__device__ __forceinline__ void to_cylindrical(double &x, double &vx, double &vz, double dt) {
  const double y = __dmul_rn(dt, vz);
  const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
  double sn = __ddiv_rn(y, r);
  if (r == 0.0) sn = 0.0;
  const double cs = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(sn, sn)));
  const double vr = __dadd_rn(__dmul_rn(cs, vx), __dmul_rn(sn, vz));
  const double vy = __dadd_rn(__dmul_rn(-sn, vx), __dmul_rn(cs, vz));
  x = r;
  vx = vr;
  vz = vy;
}

// Julia Base mod(x, y) for Float64 (float.jl) and fld = round((x - mod(x,y))/y) (div.jl).
__device__ __forceinline__ double jl_fld(double x, double y) {
  const double r = fmod(x, y);
  double md;
  if (r == 0.0) md = copysign(r, y);
  else if ((r > 0.0) != (y > 0.0)) md = __dadd_rn(r, y);
  else md = r;
  return rint(__ddiv_rn(__dsub_rn(x, md), y));
}

// One axis of discard!/wrap!  (wrap.jl:1-33).  Returns true when the particle is discarded.
// Fast path: 0 <= x-o < L  <=>  fld(x-o, L) == 0 (DESIGN.md "boundary fast path").
// Everything is passed and returned BY VALUE: a reference parameter of a non-inlined function would
// pin the caller's x to a local-memory slot, and a local load in the hot loop shares its scoreboard
// with the row prefetch.  Returns the wrapped coordinate, or NaN when the particle is discarded
// (a live particle never has a NaN coordinate).
static __device__ __noinline__ double boundary_axis_slow(double x, double xs, double L, int mode) {
  const double a = jl_fld(xs, L);
  if (a != 0.0) {
    if (mode == ISKB_BND_DISCARD) return __longlong_as_double(0x7ff8000000000000LL);
    x = __dsub_rn(x, __dmul_rn(a, L));
  }
  return x;
}
__device__ __forceinline__ bool boundary_axis(double &x, double o, double L, int mode) {
  if (mode == ISKB_BND_NONE) return false;
  const double xs = __dsub_rn(x, o);
  if (xs >= 0.0 && xs < L) return false;
  const double r = boundary_axis_slow(x, xs, L, mode);   // rare: only rows that actually crossed an edge
  if (r != r) return true;
  x = r;
  return false;
}

__device__ __forceinline__ bool cell_in_grid(int i, int j, int nx, int ny) {
  // 1 <= i <= nx-1 and 1 <= j <= ny-1 as two unsigned range checks
  return (unsigned)(i - 1) < (unsigned)(nx - 1) && (unsigned)(j - 1) < (unsigned)(ny - 1);
}

// grid_to_particle  cloud_in_cell.jl:20-36 for one particle; false when the cell is outside the grid
__device__ __forceinline__ bool gather_E(const double2 *__restrict__ E2, const GridDev &g, double x,
                                         double y, double &ex, double &ey) {
  int i, j;
  double hx, hy;
  cell1(x, g.dx, g.rdx, g.fast_div, i, hx);
  cell1(y, g.dy, g.rdy, g.fast_div, j, hy);
  if (!cell_in_grid(i, j, g.nx, g.ny)) {
    ex = ey = 0.0;
    return false;
  }
  const CicW w = cic_weights(hx, hy);
  const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * g.nx;
  const double2 e00 = __ldg(&E2[n00]), e10 = __ldg(&E2[n00 + 1]);
  const double2 e01 = __ldg(&E2[n00 + g.nx]), e11 = __ldg(&E2[n00 + g.nx + 1]);
  ex = cic_gather(w, e00.x, e10.x, e01.x, e11.x);
  ey = cic_gather(w, e00.y, e10.y, e01.y, e11.y);
  return true;
}


// ---- Philox4x32-10 counter-based RNG (Salmon et al. 2011), written from the published spec ---
struct Philox4 {
  uint32_t c[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  Philox4 o;
  o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
  return o;
}
// uniform in [0,1) with 53 bits (like Julia's rand(Float64) range)
__host__ __device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  const uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)b * (1.0 / 9007199254740992.0);
}
// uniform in (0,1] for log()
__host__ __device__ __forceinline__ double u01_open(uint32_t hi, uint32_t lo) {
  const uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)b + 1.0) * (1.0 / 9007199254740992.0);
}
