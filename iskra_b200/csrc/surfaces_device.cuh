// Device side of the surface tracker (ParticleInCell/src/pic/surfaces/{track,check,hit}.jl,
// circuit_coupling.jl:44-61), shared by surfaces.cu (operator-level kernels, simple fused advance) and
// advance_fused.cu (tiled fused advance).  See surfaces.cu for the mapping from the reference's FIFO of
// tracked tuples to one independent walk per particle.
#pragma once
#include "pic_device.cuh"

constexpr int WALK_CAP = 1 << 16;

__device__ __forceinline__ double nan_dead() { return __longlong_as_double(0x7ff8000000000000LL); }

// particle_cell(px, p, st.dh)  track.jl:47 -- BOTH coordinates are divided by the scalar st.dh (quirk S2)
__device__ __forceinline__ bool tracked_cell(const TrackerDev &t, double x, double y, int &i, int &j, double &hx,
                                             double &hy) {
  cell1(x, t.dh, i, hx);
  cell1(y, t.dh, j, hy);
  if ((unsigned)i > (unsigned)t.nx || (unsigned)j > (unsigned)t.ny) return false;   // not a key of the Dict
  return t.tracked[i + j * (t.nx + 1)] != 0;                                       // (i,j) in st, build.jl:86-93
}

// check! loop body for one particle, check.jl:48-62 with check :17-36 and hit! inlined.
// Returns true when the particle was absorbed.  x, y, vx, vy are updated by reflections.
__device__ __forceinline__ bool walk_tracked(const TrackerDev &t, double dt, int i, int j, double hx, double hy, double &x,
                                             double &y, double &vx, double &vy, double qw, int *status) {
  const double dh = t.dh;
  for (int it = 0; it < WALK_CAP; ++it) {
    const double dx = vx > 0 ? __dmul_rn(dh, __dsub_rn(1.0, hx)) : __dmul_rn(dh, hx);   // :20
    const double dy = vy > 0 ? __dmul_rn(dh, __dsub_rn(1.0, hy)) : __dmul_rn(dh, hy);   // :21
    const double dtx = __ddiv_rn(dx, fabs(vx)), dty = __ddiv_rn(dy, fabs(vy));          // :23
    if (dt < dtx && dt < dty) return false;                                              // :24-26 stayed in the cell
    int i2 = i, j2 = j, dir;
    double hx2, hy2, dt2;
    if (dtx < dty) {                                                                      // :28-31
      dt2 = __dsub_rn(dt, dtx);
      hy2 = __dadd_rn(hy, __ddiv_rn(__dmul_rn(vy, dtx), dh));
      if (vx > 0) { i2 = i + 1; hx2 = 0.0; dir = 1; } else { i2 = i - 1; hx2 = 1.0; dir = 3; }
    } else {                                                                              // :32-35
      dt2 = __dsub_rn(dt, dty);
      hx2 = __dadd_rn(hx, __ddiv_rn(__dmul_rn(vx, dty), dh));
      if (vy > 0) { j2 = j + 1; hy2 = 0.0; dir = 2; } else { j2 = j - 1; hy2 = 1.0; dir = 0; }
    }
    int sid = 0;                                                                          // get(st, (ij, ij'), nothing) :55
    if ((unsigned)i <= (unsigned)t.nx && (unsigned)j <= (unsigned)t.ny) sid = t.face[4 * (i + j * (t.nx + 1)) + dir];
    if (sid == 0) {                                                                       // :59 track!(st, pt')
      i = i2; j = j2; hx = hx2; hy = hy2; dt = dt2;
      continue;
    }
    const int kind = t.s_kind[sid];
    if (kind == ISKB_SURF_ABSORBING || kind == ISKB_SURF_ELECTRODE_FIXED) return true;    // hit.jl:32-38, circuit_coupling.jl:55-61
    if (kind == ISKB_SURF_ELECTRODE_FLOATING) {                                           // circuit_coupling.jl:44-53
      atomicAdd(&t.s_dq[sid], qw);                                                        // s.dq += q*wg[p]
      // s.sigma .+= dq/s.area lands in the solution vector in the reference (quirk S1) and is overwritten by
      // the next solve; it reaches the sigma right-hand side only when the caller asks for it
      if (t.route_hits && t.s_dof[sid] >= 0) atomicAdd(&t.sigma[t.s_dof[sid]], __ddiv_rn(qw, t.s_area[sid]));
      return true;
    }
    if (kind == ISKB_SURF_REFLECTIVE) {                                                   // hit.jl:39-56
      x = __dsub_rn(x, __dmul_rn(vx, dt2));                                               // :47 px .-= pv*dt'
      y = __dsub_rn(y, __dmul_rn(vy, dt2));
      if (i2 != i) vx = -vx;                                                              // :48-52 n = [i'-i, j'-j, 0]
      if (j2 != j) vy = -vy;
      x = __dadd_rn(x, __dmul_rn(vx, dt2));                                               // :53 px .+= pv*dt'
      y = __dadd_rn(y, __dmul_rn(vy, dt2));
      int i3 = i2, j3 = j2;                                                               // scattered!  hit.jl:12-20
      double hx3 = hx2, hy3 = hy2;
      if (hx2 == 0.0) { i3 = i2 - 1; hx3 = 1.0; }
      if (hx2 == 1.0) { i3 = i2 + 1; hx3 = 0.0; }
      if (hy2 == 0.0) { j3 = j2 - 1; hy3 = 1.0; }
      if (hy2 == 1.0) { j3 = j2 + 1; hy3 = 0.0; }
      i = i3; j = j3; hx = hx3; hy = hy3; dt = dt2;
      continue;
    }
    return false;                                                                          // PeriodicSurface: no-op hit!, hit.jl:26-31
  }
  atomicOr(status, ISKB_ST_WALK);
  return false;
}


