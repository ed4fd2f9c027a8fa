// Secondary-electron emission at a wall: emit!(primary, secondary, grid, material; boundary, gamma_t, gamma_e, gamma_i)
// Chemistry/src/see.jl:114-181, one thread per row of the primary species.
//
// The reference walks the rows in reverse with one sequential MersenneTwister; here every row has its own Philox
// stream (counter = row, call, draw), so parity is statistical like for the MCC step.  Everything
// else follows the reference line by line, quirks included (oracle/see_oracle.py lists them): "theta" is the sine of
// the incidence angle, dt = mod(x_i, L_i)/|v_i|, an emitted secondary does not remove its primary, and only
// `R2 >= gamma` absorbs it.  Absorbed rows are marked dead (x = NaN) and parked by the next compaction together
// with their ids -- the same permutation invariant as remove! (kinetic.jl:20-27).
// The emission coefficients are the reference's own closures (see.jl:17-58) with their parameters passed by value.
#include "mcc_device.cuh"

namespace {

constexpr double QE_SEE = 1.60217646e-19;   // see.jl:65

struct SeeDev {
  SpDev prim, sec;
  GridDev g;
  int axis;          // 0: left / right, 1: bottom / top
  int lower;         // predicate: alpha < 0 (left, bottom) or alpha > 0 (right, top)   see.jl:67-79
  double n[3];       // wall normal  see.jl:81-89
  double m_prim, m_sec;   // mass(species) = m / qe  [eV s^2/m^2]
  // vaughan(w0, w0max, g0max, ks); elastic(we, wemax, gemax, De, re) (gemax < 0: gamma_e == 0); inelastic(ri) (< 0: == 0);
  // secondary(re_t, ri_t)
  double w0, w0max, g0max, ks, we, wemax, gemax, De, re, ri, re_t, ri_t;
  uint32_t k0, k1, call;
  unsigned long long *counts;   // elastic, inelastic, secondaries, absorbed
  int *status;
};

__device__ double vaughan(const SeeDev &s, double w, double th) {   // see.jl:17-26
  const double wmax = s.w0max * (1.0 + s.ks / M_PI * (th * th));
  const double gmax = s.g0max * (1.0 + s.ks / (2.0 * M_PI) * (th * th));
  const double v = w > s.w0 ? (w - s.w0) / (wmax - s.w0) : 0.0;
  const double k = w > wmax ? 0.25 : 0.62;
  return gmax * pow(v * exp(1.0 - v), k);
}
__device__ double gamma_e(const SeeDev &s, double w, double th) {   // see.jl:29-42
  if (s.gemax < 0.0) return 0.0;
  if (s.we < w && w <= s.wemax) {
    const double v1 = (w - s.we) / (s.wemax - s.we);
    return s.re * vaughan(s, w, th) + s.gemax * v1 * exp(1.0 - v1);
  }
  if (w > s.wemax) {
    const double v2 = (w - s.wemax) / s.De;
    return s.re * vaughan(s, w, th) + s.gemax * (1.0 + v2) * exp(-v2);
  }
  return 0.0;
}

__device__ double jl_mod_dev(double x, double y) {   // Julia Base mod(x, y), float.jl
  const double r = fmod(x, y);
  if (r == 0.0) return copysign(r, y);
  return ((r > 0.0) != (y > 0.0)) ? r + y : r;
}

// inject_secondary!(secondary, x, n, dt)  see.jl:107-117
__device__ void inject_secondary(const SeeDev &s, Rng &g, const double *x0, double dt, unsigned &n_sec) {
  double z0, z1;
  g.randn2(z0, z1);
  const double eps = exp(1.65 + 1.1 * z0);               // rand(LogNormal(1.65, 1.1))  :11-14
  double dir[3];
  diffuse_reflection(s.n, g, dir);
  const double sp = sqrt(2.0 * eps / s.m_sec);
  const double v[3] = {dir[0] * sp, dir[1] * sp, dir[2] * sp};
  if (append_row(s.sec, dt * v[0] + x0[0], dt * v[1] + x0[1], v, s.status)) ++n_sec;
}

__global__ void __launch_bounds__(128) k_see_emit(SeeDev s) {
  __shared__ unsigned sc[4];
  if (threadIdx.x < 4) sc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n = s.prim.cnt[CNT_BEGIN];   // rows that existed when the call started (appended secondaries are not revisited)
  const double L = s.axis == 0 ? s.g.Lx : s.g.Ly, o = s.axis == 0 ? s.g.ox : s.g.oy;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    double x[2] = {s.prim.col[0][p], s.prim.col[1][p]};
    if (is_dead(x[0])) continue;
    const double a = jl_fld(x[s.axis] - o, L);
    if (!(s.lower ? a < 0.0 : a > 0.0)) continue;         // hit(alpha)
    double pv[3] = {s.prim.col[2][p], s.prim.col[3][p], s.prim.col[4][p]};
    Rng g(p, s.call, s.k0, s.k1, 0u);
    const double cx = pv[1] * s.n[2] - pv[2] * s.n[1], cy = pv[2] * s.n[0] - pv[0] * s.n[2], cz = pv[0] * s.n[1] - pv[1] * s.n[0];
    const double th = sqrt(cx * cx + cy * cy + cz * cz) / norm3(pv);                 // :133 (the sine)
    const double w = 0.5 * s.m_prim * ((pv[0] * pv[0] + pv[1] * pv[1]) + pv[2] * pv[2]);
    const double R1 = g.u01();
    const double gv = vaughan(s, w, th);
    const double ge = gamma_e(s, w, th), gi = s.ri < 0.0 ? 0.0 : s.ri * gv, gt = (1.0 - s.re_t - s.ri_t) * gv;
    const double dt = jl_mod_dev(x[s.axis], L) / fabs(pv[s.axis]);                  // :141
    const double x0[2] = {x[0] - pv[0] * dt, x[1] - pv[1] * dt};
    const bool inel = ge + gi > R1 && R1 > ge;
    if (inel || R1 < ge) {
      // x -= v*dt ; v = [rand()*] snells_law(v, n) ; x += v*dt   :146-162
      const double f = inel ? g.u01() : 1.0;
      const double nd = 2.0 * (s.n[0] * pv[0] + s.n[1] * pv[1] + s.n[2] * pv[2]);
      for (int k = 0; k < 3; ++k) pv[k] = f * (pv[k] - nd * s.n[k]);
      s.prim.col[0][p] = x0[0] + pv[0] * dt;
      s.prim.col[1][p] = x0[1] + pv[1] * dt;
      s.prim.col[2][p] = pv[0]; s.prim.col[3][p] = pv[1]; s.prim.col[4][p] = pv[2];
      atomicAdd(&sc[inel ? 1 : 0], 1u);
      continue;
    }
    double gam = ge + gi + gt;
    unsigned n_sec = 0;
    while (gam > 1.0) {                                    // :165-169
      inject_secondary(s, g, x0, dt, n_sec);
      gam -= 1.0;
    }
    if (g.u01() < gam) {
      inject_secondary(s, g, x0, dt, n_sec);
    } else {
      s.prim.col[0][p] = __longlong_as_double(0x7ff8000000000000LL);   // remove!(primary, p): absorbed
      atomicAdd((unsigned long long *)&s.prim.cnt[CNT_NDEAD], 1ull);
      atomicAdd(&sc[3], 1u);
    }
    if (n_sec) atomicAdd(&sc[2], n_sec);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sc[threadIdx.x]) atomicAdd(&s.counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

__global__ void k_see_begin(int64_t *cnt, unsigned long long *counts) {
  cnt[CNT_BEGIN] = cnt[CNT_NSLOTS];
  for (int k = 0; k < 4; ++k) counts[k] = 0;
}

}  // namespace

// boundary: ISKB_EDGE_LEFT / RIGHT / BOTTOM / TOP.  coef[12] = {w0, w0max, g0max, ks, we, wemax, gemax, De, re, ri, re_t, ri_t}
// (gemax < 0: gamma_e == gamma_0; ri < 0: gamma_i == gamma_0).  counts_out[4] = elastic, inelastic, secondaries, absorbed.
extern "C" int32_t iskb_see_emit(iskb_species *primary, iskb_species *secondary, int32_t boundary, const double *coef,
                                 uint64_t seed, int64_t *counts_out) {
  if (!primary || !secondary || !coef || primary->ctx != secondary->ctx) return iskb_fail(ISKB_E_INVALID, "iskb_see_emit: bad arguments");
  if (boundary < ISKB_EDGE_LEFT || boundary > ISKB_EDGE_TOP)
    return iskb_fail(ISKB_E_UNSUPPORTED, "emit! with boundary = :all has a zero wall normal (see.jl:94): NaN secondaries in the reference");
  iskb_ctx *c = primary->ctx;
  CU_TRY(cudaSetDevice(c->device));
  sp_touch(primary);
  if (secondary != primary) sp_touch(secondary);
  SeeDev s;
  s.prim = spdev(primary);
  s.sec = spdev(secondary);
  s.g = c->g;
  s.axis = (boundary == ISKB_EDGE_LEFT || boundary == ISKB_EDGE_RIGHT) ? 0 : 1;
  s.lower = (boundary == ISKB_EDGE_LEFT || boundary == ISKB_EDGE_BOTTOM) ? 1 : 0;
  s.n[0] = boundary == ISKB_EDGE_LEFT ? -1.0 : boundary == ISKB_EDGE_RIGHT ? 1.0 : 0.0;
  s.n[1] = boundary == ISKB_EDGE_BOTTOM ? -1.0 : boundary == ISKB_EDGE_TOP ? 1.0 : 0.0;
  s.n[2] = 0.0;
  s.m_prim = primary->m / QE_SEE;
  s.m_sec = secondary->m / QE_SEE;
  s.w0 = coef[0]; s.w0max = coef[1]; s.g0max = coef[2]; s.ks = coef[3];
  s.we = coef[4]; s.wemax = coef[5]; s.gemax = coef[6]; s.De = coef[7]; s.re = coef[8];
  s.ri = coef[9]; s.re_t = coef[10]; s.ri_t = coef[11];
  s.k0 = (uint32_t)seed;
  s.k1 = (uint32_t)(seed >> 32) ^ (0xC2B2AE35u * (uint32_t)(c->rank + 1));
  s.call = (uint32_t)(primary->see_calls++);
  if (!c->d_see_counts) CU_TRY(cudaMalloc(&c->d_see_counts, 4 * sizeof(unsigned long long)));
  s.counts = c->d_see_counts;
  s.status = c->d_status;
  k_see_begin<<<1, 1, 0, c->stream>>>(primary->d_cnt, c->d_see_counts);
  LAUNCH_CHECK(c);
  const int64_t bound = primary->counts_stale ? primary->cap : primary->h_nslots;
  int64_t blocks = (bound + 127) / 128;
  if (blocks > (int64_t)c->n_sm * 16) blocks = (int64_t)c->n_sm * 16;
  if (blocks < 1) blocks = 1;
  k_see_emit<<<(int)blocks, 128, 0, c->stream>>>(s);
  LAUNCH_CHECK(c);
  primary->counts_stale = true;
  secondary->counts_stale = true;
  ISKB_TRY(sp_vmax_unknown(primary));
  ISKB_TRY(sp_vmax_unknown(secondary));
  unsigned long long h[4];
  CU_TRY(cudaMemcpyAsync(h, c->d_see_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  if (counts_out) for (int k = 0; k < 4; ++k) counts_out[k] = (int64_t)h[k];
  return ctx_check_status(c);
}
