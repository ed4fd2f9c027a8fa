// Device helpers shared by the collision kernels (mcc.cu) and the wall emission kernel (see.cu): the per-row Philox
// stream, the reference's scattering kinematics (Chemistry/src/mcc.jl:53-127) and the row append.
#pragma once
#include <cmath>

#include "pic_device.cuh"

namespace {

struct SpDev {
  double *col[6];
  int64_t *cnt;
  int64_t cap;
  unsigned long long *vmax2, *vz2max;
};

struct Rng {   // per-row stream: counter = (row_lo, row_hi, call, draw)
  uint32_t r0, r1, call, k0, k1, draw;
  uint32_t buf[4];
  int have;
  __device__ Rng(int64_t row, uint32_t call_, uint32_t k0_, uint32_t k1_, uint32_t first_draw)
      : r0((uint32_t)row), r1((uint32_t)(row >> 32)), call(call_), k0(k0_), k1(k1_), draw(first_draw), have(0) {}
  __device__ uint32_t next32() {
    if (!have) {
      const Philox4 o = philox4x32_10(r0, r1, call, draw++, k0, k1);
      buf[0] = o.c[0]; buf[1] = o.c[1]; buf[2] = o.c[2]; buf[3] = o.c[3];
      have = 4;
    }
    return buf[--have];
  }
  __device__ double u01() { const uint32_t a = next32(), b = next32(); return u01_53(a, b); }
  __device__ void randn2(double &z0, double &z1) {   // Box-Muller
    const uint32_t a = next32(), b = next32(), c = next32(), d = next32();
    const double r = sqrt(-2.0 * log(u01_open(a, b)));
    double s, co;
    sincospi(2.0 * u01_53(c, d), &s, &co);
    z0 = r * co;
    z1 = r * s;
  }
};

__device__ __forceinline__ double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// euler_angles :89-106, unrotated :83-87, [sc*ce sc*se cc] * T  (:112, :126)
__device__ void scatter3(const double *v, double sc, double cc, double se, double ce, double *out) {
  const double nv = norm3(v);
  const double ct = v[2] / nv;
  const double st = sqrt(1.0 - ct * ct);
  double cp, sp;
  if (st == 0.0) { cp = 1.0; sp = 0.0; }
  else { cp = v[0] / nv / st; sp = v[1] / nv / st; }
  const double r0 = sc * ce, r1 = sc * se, r2 = cc;
  out[0] = (r0 * (cp * ct) + r1 * (-sp)) + r2 * (cp * st);
  out[1] = (r0 * (-sp * ct) + r1 * cp) + r2 * (sp * st);
  out[2] = (r0 * (-st) + r1 * 0.0) + r2 * ct;
}
__device__ void isotropic_scattering(const double *v, Rng &g, double *out) {   // :53-62, :108-113
  double sc, cc, se, ce;
  sincos(2.0 * M_PI * g.u01(), &sc, &cc);
  sincos(2.0 * M_PI * g.u01(), &se, &ce);
  scatter3(v, sc, cc, se, ce, out);
}
__device__ void diffuse_reflection(const double *v, Rng &g, double *out) {     // :64-72, :122-127
  const double sc = sqrt(g.u01());
  const double cc = -sqrt(1.0 - sc * sc);
  double se, ce;
  sincos(2.0 * M_PI * g.u01(), &se, &ce);
  scatter3(v, sc, cc, se, ce, out);
}

__device__ __forceinline__ void raise_vmax(const SpDev &s, const double *v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  // almost never taken: same-address atomics are slow.  Plain L2 loads (not volatile): a stale bound only costs a redundant
  // atomicMax, and the compiler may issue the two loads ahead of the kinematics instead of at the end of the dependent chain
  if (b > __ldcg(s.vmax2)) atomicMax(s.vmax2, b);
  const unsigned long long bz = (unsigned long long)__double_as_longlong(v[2] * v[2]);   // kept for the lean advance (advance_tile.cu)
  if (bz > __ldcg(s.vz2max)) atomicMax(s.vz2max, bz);
}

__device__ bool append_row(const SpDev &s, double x, double y, const double *v, int *status) {
  const int64_t slot = (int64_t)atomicAdd((unsigned long long *)&s.cnt[CNT_NSLOTS], 1ull);
  if (slot >= s.cap) {
    atomicAdd((unsigned long long *)&s.cnt[CNT_NSLOTS], (unsigned long long)(-1ll));
    atomicOr(status, ISKB_ST_CAPACITY);
    return false;
  }
  s.col[0][slot] = x; s.col[1][slot] = y;
  s.col[2][slot] = v[0]; s.col[3][slot] = v[1]; s.col[4][slot] = v[2];
  raise_vmax(s, v);
  // wg and id of the slot stay as parked there (kinetic.jl:29-37 "dst has already correct ID")
  return true;
}

inline SpDev spdev(const iskb_species *s) {
  SpDev d;
  for (int q = 0; q < 6; ++q) d.col[q] = s->col[q];
  d.cnt = s->d_cnt;
  d.cap = s->cap;
  d.vmax2 = s->d_vmax2;
  d.vz2max = s->d_vz2max;
  return d;
}


}  // namespace
