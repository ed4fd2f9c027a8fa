// Monte-Carlo null-collision step, Chemistry/src/mcc.jl:231-289, on the device.
//
// The reference draws Nc = N*max_Pt*np candidates *with replacement* from one sequential
// MersenneTwister stream (mcc.jl:248-251).  Here every particle row gets its own counter-based
// Philox4x32-10 stream keyed (seed, rank) with counter (row, call#, draw#).  The reference's null-collision rate is
// N times larger than it has to be: a candidate picks one of the N processes uniformly and accepts it with
// P_k / max_Pt, where max_Pt already bounds the SUM over the processes (mcc.jl:27-51, 243).  Here a row is a candidate
// with probability p_sel = n_max * sup(sum_k sigma_k g) * dt >= sum_k P_k, and a candidate collides with process k
// when U * p_sel falls into [P_1 + .. + P_{k-1}, P_1 + .. + P_k): per particle and process the collision probability
// is 1 - exp(-n sigma_k g dt) in both schemes (they differ at second order in P, like the reference's sampling with
// replacement), but only 1/N of the rows are fetched from memory -- the test phase is bound by exactly
// that random DRAM traffic (profiles/r2_ncu_mcc.md).  Kinematics follow the reference line by line.
//
// Three dense phases with warp-aggregated compaction between them (no divergent fat paths):
//   k_mcc_select  : the candidate decision touches every row but needs no particle data -- one
//                   Philox call decides four rows; candidate rows are appended to a list
//                   (one atomic per warp via ballot/popc);
//   k_mcc_test    : one thread per candidate: cell, density, process choice, sigma(eps), acceptance
//                   (mcc.jl:252-281); accepted rows go to the collider list the same way;
//   k_mcc_collide : one thread per collider: kinematics (mcc.jl:129-229), ionisation appends.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mcc_device.cuh"

namespace {

constexpr int TPB = 256;
constexpr double QE_MCC = 1.60217646e-19;   // mcc.jl:26
constexpr double KB_MCC = 1.3806503e-23;    // mcc.jl:75

struct ProcDev {
  int kind;
  double threshold;
  int offset, len;
  int prod;        // index into MccDev::prod or -1
  int n_prod;      // round(source.w0 / product.w0)   mcc.jl:205-207
};

struct MccDev {
  SpDev src;
  SpDev prod[4];
  ProcDev proc[8];
  int N;
  const double *eps, *sig, *tn;
  const double2 *E2;
  GridDev g;
  double tqm;        // target.q / target.m            mcc.jl:266
  double vth_t;      // thermal_speed(target.T, target.m)  mcc.jl:74-77
  double mr1, mr2;   // mass ratios                    mcc.jl:131-132
  double m_eV;       // mass(source)                   mcc.jl:26
  double dt;
  double p_cand;     // host estimate of the candidate probability (grid sizes only)
  double sup_total;      // sup over the tables of sum_k sigma_k(eps)*g(eps), interior maxima included
  double sig_last_total; // sum_k sigma_k(last knot): Flat() beyond the tables, so sum sigma*g <= sig_last_total*|v|max there
  double n0max;
  int need_cell;     // 0: the target density is the same on every node and no nu map is asked for
  const unsigned long long *vmax2;   // bits of the species-wide bound of |v|^2 (+inf = unknown)
  double *pk_dev;       // [0] p_sel, [1] 1/log(1 - p_sel) (geometric gaps), [2] p_sel * 2^32: computed on the device per call
  uint32_t k0, k1;   // Philox key
  uint32_t call;
  unsigned long long *stats;
  float *nu;         // nullable
  int *status;
};

// cross_section.jl:8-14: LinearInterpolation(xs, ys; extrapolation_bc = Flat())
__device__ double xsec_eval(const double *__restrict__ xs, const double *__restrict__ ys, int n, double x) {
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

// perform!(collision, p, ...)  mcc.jl:129-229
__device__ void collide(const MccDev &m, const ProcDev &pc, int64_t p, Rng &g) {
  double sv[3] = {m.src.col[2][p], m.src.col[3][p], m.src.col[4][p]};
  if (pc.kind <= ISKB_MCC_INELASTIC_BACKWARD) {
    double tv[3], z;
    g.randn2(tv[0], tv[1]);
    g.randn2(tv[2], z);
    double vr[3], w[3], dir[3];
    for (int k = 0; k < 3; ++k) {
      tv[k] *= m.vth_t;
      vr[k] = sv[k] - tv[k];
      w[k] = m.mr1 * sv[k] + m.mr2 * tv[k];
    }
    double mag = norm3(vr);
    if (pc.kind == ISKB_MCC_ELASTIC_ISOTROPIC) isotropic_scattering(vr, g, dir);
    else if (pc.kind == ISKB_MCC_ELASTIC_BACKWARD) diffuse_reflection(vr, g, dir);
    else { mag = g.u01() * mag; diffuse_reflection(vr, g, dir); }
    double nv[3];
    for (int k = 0; k < 3; ++k) { nv[k] = w[k] + m.mr2 * (mag * dir[k]); m.src.col[2 + k][p] = nv[k]; }
    raise_vmax(m.src, nv);
    return;
  }
  const double sE = 0.5 * m.m_eV * ((sv[0] * sv[0] + sv[1] * sv[1]) + sv[2] * sv[2]) - pc.threshold;
  if (sE < 0) return;                                   // :179-182, :220-223
  if (pc.kind == ISKB_MCC_EXCITATION) {
    const double ev = sqrt(2.0 / m.m_eV) * sqrt(sE);
    double dir[3];
    isotropic_scattering(sv, g, dir);
    double nv[3];
    for (int k = 0; k < 3; ++k) { nv[k] = ev * dir[k]; m.src.col[2 + k][p] = nv[k]; }
    raise_vmax(m.src, nv);
    return;
  }
  // ionization :184-212
  const double e1E = sE * g.u01(), e2E = sE - e1E;
  const double alpha = sqrt(2.0 / m.m_eV), e1v = alpha * sqrt(e1E), e2v = alpha * sqrt(e2E);
  double d1[3], d2[3];
  diffuse_reflection(sv, g, d1);
  for (int k = 0; k < 3; ++k) { sv[k] = e1v * d1[k]; m.src.col[2 + k][p] = sv[k]; }
  raise_vmax(m.src, sv);
  diffuse_reflection(sv, g, d2);
  const double x = m.src.col[0][p], y = m.src.col[1][p];
  double nv[3] = {e2v * d2[0], e2v * d2[1], e2v * d2[2]};
  if (!append_row(m.src, x, y, nv, m.status)) return;
  double tv[3], z;
  g.randn2(tv[0], tv[1]);
  g.randn2(tv[2], z);
  for (int k = 0; k < 3; ++k) tv[k] *= m.vth_t;
  if (pc.prod >= 0)
    for (int c = 0; c < pc.n_prod; ++c)
      if (!append_row(m.prod[pc.prod], x, y, tv, m.status)) return;
}

__global__ void k_snapshot_begin(int64_t *cnt, unsigned int *lists_cnt, MccDev m) {
  cnt[CNT_BEGIN] = cnt[CNT_NSLOTS];
  lists_cnt[0] = 0;   // candidates
  lists_cnt[1] = 0;   // colliders
  // candidate probability: p_sel >= sum_k P_k for EVERY live row.  1 - exp(-a) <= a, so n sup(sum sigma g) dt bounds
  // the sum; beyond the last knot every sigma_k is flat, so there sum sigma g <= sig_last_total*|v|max with the
  // species-wide bound of |v|^2 kept by the advance kernels (neutral target: g = |v|).  A row that still exceeds the
  // bound raises ISKB_ST_PK in the test phase like mcc.jl:273-279; without a speed bound (unknown after an upload, charged
  // target) the rate is the reference's N * max_Pt, so that the error needs what it needs there: P > N times the table bound.
  const double v2 = __longlong_as_double((long long)*m.vmax2);
  double sg = m.sup_total;
  if (m.tqm == 0.0 && v2 < 1e300) sg = fmax(sg, m.sig_last_total * sqrt(v2));
  else sg *= m.N;   // no speed bound (first call after an upload, charged target): the reference's own margin, N times the table bound
  double psel = m.n0max * sg * m.dt * (1.0 + 1e-9);
  if (!(psel < 1.0)) psel = 1.0;
  if (!(psel > 0.0)) psel = 0.0;
  m.pk_dev[0] = psel;
  m.pk_dev[1] = psel < 1.0 ? 1.0 / log1p(-psel) : 0.0;   // p = 1: every gap is 0
  const double p32 = psel * 4294967296.0;
  m.pk_dev[2] = p32 >= 4294967295.0 ? 4294967295.0 : floor(p32);
}

constexpr int SEL_CALLS = 4;   // Philox calls (groups of 4 rows) per thread and block iteration: 4096 rows per block iteration
__global__ void __launch_bounds__(256) k_mcc_select(MccDev m, unsigned int *lists_cnt, uint32_t *cand,
                                                    unsigned int cand_cap) {
  __shared__ unsigned int s_warp[8];
  __shared__ unsigned int s_base;
  const int64_t n = m.src.cnt[CNT_BEGIN];   // rows that existed when the step started
  const int64_t nq = (n + 3) / 4;
  const uint32_t p_u32 = (uint32_t)m.pk_dev[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long my_cand = 0;
  for (int64_t q0 = (int64_t)blockIdx.x * (256 * SEL_CALLS); q0 < nq; q0 += (int64_t)gridDim.x * (256 * SEL_CALLS)) {
    // thread t owns the groups q0 + SEL_CALLS*t .. +SEL_CALLS-1, i.e. 4*SEL_CALLS consecutive rows: the
    // candidate list stays in increasing row order within a block iteration
    const int64_t qb = q0 + (int64_t)threadIdx.x * SEL_CALLS;
    unsigned keep = 0;
#pragma unroll
    for (int u = 0; u < SEL_CALLS; ++u) {
      const int64_t q = qb + u;
      if (q < nq) {
        // one Philox call decides candidacy of rows 4q..4q+3 (draw index 0 of row 4q)
        const Philox4 o = philox4x32_10((uint32_t)(q * 4), (uint32_t)((q * 4) >> 32), m.call, 0u, m.k0, m.k1);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (q * 4 + s < n && o.c[s] < p_u32) keep |= 1u << (4 * u + s);
      }
    }
    const unsigned cnt = __popc(keep);
    my_cand += cnt;
    // block-aggregated append: ONE global atomic per block iteration
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0;
      for (int w = 0; w < 8; ++w) {
        const unsigned c = s_warp[w];
        s_warp[w] = tot;
        tot += c;
      }
      s_base = tot ? atomicAdd(&lists_cnt[0], tot) : 0u;
    }
    __syncthreads();
    unsigned slot = s_base + s_warp[warp] + incl - cnt;
    while (keep) {
      const int b = __ffs(keep) - 1;
      keep &= keep - 1;
      if (slot < cand_cap) cand[slot] = (uint32_t)(qb * 4 + b);
      else atomicOr(m.status, ISKB_ST_CAPACITY);
      ++slot;
    }
    __syncthreads();
  }
  // candidates statistic (rows drawn; includes the few discarded rows still parked in the columns)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_cand += __shfl_down_sync(0xffffffffu, my_cand, d);
  if (lane == 0 && my_cand) atomicAdd(&m.stats[0], my_cand);
}

// Same decision, drawn the other way round: with independent Bernoulli(p) trials per row the gap between
// two consecutive candidates is geometric, gap = floor(log U / log(1-p)), and because the geometric law is
// memoryless the sequence may restart at every chunk boundary.  Each thread owns 64 consecutive rows and
// draws gaps until it leaves them: the work is proportional to the candidates (6 % / 1 % of the rows at
// C5) instead of the rows (one Philox call per 4 rows before: 2 x 120 us per step under ncu).  Measured on
// the C5 step: 3.82 -> 3.77 ms -- most of the selection was already hidden behind the field solve on the
// side stream.  k_mcc_select (one draw per row) remains for p = 0 / p = 1.  One Philox call yields four gaps;
// counter = (first row of the chunk, call, 0x40000000 + k), disjoint from the per-row streams (draws 0, 1, ...).
constexpr int SKIP_ROWS = 256;   // rows per thread: one Philox call (four gaps) is the fixed cost of a chunk, and at p = 0.5 % (ions)
                                 // a 64-row chunk held 0.3 candidates -- the kernel was all fixed cost (30 us for 3e5 candidates)
constexpr int SKIP_WORDS = SKIP_ROWS / 64;
__global__ void __launch_bounds__(256) k_mcc_select_skip(MccDev m, unsigned int *lists_cnt, uint32_t *cand,
                                                         unsigned int cand_cap) {
  __shared__ unsigned int s_warp[8];
  __shared__ unsigned int s_base;
  const int64_t n = m.src.cnt[CNT_BEGIN];   // rows that existed when the step started
  const int64_t nchunk = (n + SKIP_ROWS - 1) / SKIP_ROWS;
  const double inv_log1mp = m.pk_dev[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long my_cand = 0;
  for (int64_t c0 = (int64_t)blockIdx.x * 256; c0 < nchunk; c0 += (int64_t)gridDim.x * 256) {
    const int64_t ch = c0 + threadIdx.x;
    const int64_t row0 = ch * SKIP_ROWS;
    unsigned long long keep[SKIP_WORDS];
#pragma unroll
    for (int w = 0; w < SKIP_WORDS; ++w) keep[w] = 0;
    if (ch < nchunk) {
      const int lim = n - row0 < SKIP_ROWS ? (int)(n - row0) : SKIP_ROWS;
      int pos = -1;
      for (uint32_t k = 0; pos < lim; ++k) {
        const Philox4 o = philox4x32_10((uint32_t)row0, (uint32_t)(row0 >> 32), m.call, 0x40000000u + k, m.k0, m.k1);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (pos >= lim) break;   // the FP64 log is most of this kernel: do not take it for gaps nobody needs
          // U uniform on (0,1): 32 bits are ample, the law is cut off at (1-p)^gap = 2^-33
          const double u = ((double)o.c[s] + 0.5) * (1.0 / 4294967296.0);
          const double gq = log(u) * inv_log1mp;
          const int gap = gq < 1048576.0 ? (int)gq : 1048576;
          pos += gap + 1;
          if (pos < lim) {
#pragma unroll
            for (int w = 0; w < SKIP_WORDS; ++w)
              if ((pos >> 6) == w) keep[w] |= 1ull << (pos & 63);
          }
        }
      }
    }
    unsigned cnt = 0;
#pragma unroll
    for (int w = 0; w < SKIP_WORDS; ++w) cnt += __popcll(keep[w]);
    my_cand += cnt;
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0;
      for (int w = 0; w < 8; ++w) {
        const unsigned c = s_warp[w];
        s_warp[w] = tot;
        tot += c;
      }
      s_base = tot ? atomicAdd(&lists_cnt[0], tot) : 0u;
    }
    __syncthreads();
    unsigned slot = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
    for (int w = 0; w < SKIP_WORDS; ++w) {
      unsigned long long kw = keep[w];
      while (kw) {
        const int b = __ffsll((long long)kw) - 1;
        kw &= kw - 1;
        if (slot < cand_cap) cand[slot] = (uint32_t)(row0 + 64 * w + b);
        else atomicOr(m.status, ISKB_ST_CAPACITY);
        ++slot;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_cand += __shfl_down_sync(0xffffffffu, my_cand, d);
  if (lane == 0 && my_cand) atomicAdd(&m.stats[0], my_cand);
}

constexpr int TEST_TPB = 128;
constexpr int TEST_BUF = 512;
constexpr int TEST_SMEM_TABLE_MAX = 6144;   // table entries (eps + sigma = 16 B each) staged in shared memory

// One thread per candidate.  The phase is latency bound (a candidate costs one DRAM round trip for its row and a
// binary search in sigma(eps)), so: the grid covers every candidate at once (a capped grid walked the list in 26
// dependent rounds: 110 us per launch at C5), the row's five columns are requested together, and the tables are
// searched in shared memory.
__global__ void __launch_bounds__(TEST_TPB) k_mcc_test(MccDev m, unsigned int *lists_cnt, const uint32_t *__restrict__ cand,
                                                       unsigned int cand_cap, uint2 *coll, int ntab) {
  extern __shared__ double s_tab[];   // eps[ntab] | sigma[ntab]  (ntab == 0: search the global tables)
  __shared__ uint2 s_hit[TEST_BUF];
  __shared__ unsigned int s_nhit, s_base;
  __shared__ unsigned int s_proc[8];
  if (threadIdx.x == 0) s_nhit = 0;
  if (threadIdx.x < 8) s_proc[threadIdx.x] = 0;
  for (int e = threadIdx.x; e < ntab; e += TEST_TPB) {
    s_tab[e] = m.eps[e];
    s_tab[ntab + e] = m.sig[e];
  }
  __syncthreads();
  const double *teps = ntab ? s_tab : m.eps, *tsig = ntab ? s_tab + ntab : m.sig;
  unsigned int nc = lists_cnt[0];
  if (nc > cand_cap) nc = cand_cap;
  const double psel = m.pk_dev[0];
  const unsigned int nc_pad = (nc + TEST_TPB - 1) / TEST_TPB * TEST_TPB;   // block-uniform trip count
  for (unsigned int t = blockIdx.x * TEST_TPB + threadIdx.x; t < nc_pad; t += gridDim.x * TEST_TPB) {
    if (t < nc) {
      const int64_t p = cand[t];
      Rng g(p, m.call, m.k0, m.k1, 1u);
      const double tsel = g.u01() * psel;                                   // process k iff tsel in [cum_{k-1}, cum_k)
      const double vx = m.src.col[2][p], vy = m.src.col[3][p], vz = m.src.col[4][p];
      const double px = m.src.col[0][p];                          // (dead rows carry x = NaN)
      const double py = m.need_cell ? m.src.col[1][p] : px;       // uniform target density, no nu map: the cell is not needed
      if (!is_dead(px)) {
        int i, j;
        double hx, hy;
        cell1(px, m.g.dx, m.g.rdx, m.g.fast_div, i, hx);
        cell1(py, m.g.dy, m.g.rdy, m.g.fast_div, j, hy);
        if (m.need_cell && !cell_in_grid(i, j, m.g.nx, m.g.ny)) {
          atomicOr(m.status, ISKB_ST_OOB);
        } else {
          const int64_t node = m.need_cell ? (int64_t)(i - 1) + (int64_t)(j - 1) * m.g.nx : 0;   // lower-left node :252-253
          const double dens = m.need_cell ? m.tn[node] : m.n0max;
          if (dens >= 0) {                                                      // :254-257
            // neutral target (tqm == 0): (0*E)*dt contributes exactly +0 for any finite E, so E is not
            // read -- the field solve of the previous step may still be writing it on the field stream
            const double2 e = m.tqm == 0.0 ? make_double2(0.0, 0.0) : m.E2[node];
            double d[3];
            d[0] = (m.tqm * e.x) * m.dt - vx;                                   // :266-267
            d[1] = (m.tqm * e.y) * m.dt - vy;
            d[2] = (m.tqm * 0.0) * m.dt - vz;
            const double gg = norm3(d);
            const double eps = 0.5 * m.m_eV * (gg * gg);                        // :268
            double cum = 0.0;
            int hit = 0;
            for (int k = 0; k < m.N; ++k) {
              const ProcDev pc = m.proc[k];
              const double skg = xsec_eval(teps + pc.offset, tsig + pc.offset, pc.len, eps) * gg;
              cum += 1.0 - exp(-dens * skg * m.dt);                             // :271  P_k
              if (!hit && tsel < cum) hit = k + 1;
            }
            if (cum > psel) atomicOr(m.status, ISKB_ST_PK);                     // :273-279: the bound does not hold for this row
            else if (hit) {
              const unsigned slot = atomicAdd(&s_nhit, 1u);                     // block-local staging
              s_hit[slot] = make_uint2((uint32_t)p, (uint32_t)hit);
              atomicAdd(&s_proc[hit - 1], 1u);
              if (m.nu) atomicAdd(&m.nu[node + (int64_t)(hit - 1) * m.g.nx * m.g.ny], 1.0f);   // :283
            }
          }
        }
      }
    }
    __syncthreads();
    if (s_nhit > TEST_BUF - TEST_TPB) {   // flush the staged colliders with ONE global atomic
      if (threadIdx.x == 0) s_base = atomicAdd(&lists_cnt[1], s_nhit);
      __syncthreads();
      for (unsigned int q = threadIdx.x; q < s_nhit; q += TEST_TPB) coll[s_base + q] = s_hit[q];
      __syncthreads();
      if (threadIdx.x == 0) s_nhit = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  if (s_nhit) {
    if (threadIdx.x == 0) s_base = atomicAdd(&lists_cnt[1], s_nhit);
    __syncthreads();
    for (unsigned int q = threadIdx.x; q < s_nhit; q += TEST_TPB) coll[s_base + q] = s_hit[q];
  }
  if (threadIdx.x < 8 && s_proc[threadIdx.x]) atomicAdd(&m.stats[2 + threadIdx.x], (unsigned long long)s_proc[threadIdx.x]);
}

__global__ void k_mcc_collide(MccDev m, const unsigned int *__restrict__ lists_cnt, const uint2 *__restrict__ coll) {
  const unsigned int nc = lists_cnt[1];
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < nc; t += gridDim.x * blockDim.x) {
    const uint2 c = coll[t];
    Rng g((int64_t)c.x, m.call, m.k0, m.k1, 2u);   // draw block 1 belonged to the acceptance test
    collide(m, m.proc[c.y - 1], (int64_t)c.x, g);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && nc) atomicAdd(&m.stats[1], (unsigned long long)nc);
}

__global__ void k_commit_stats(unsigned long long *stats) {
  if (threadIdx.x < 10) {
    stats[threadIdx.x] += stats[10 + threadIdx.x];
    stats[10 + threadIdx.x] = 0;
  }
}

double xsec_eval_host(const double *xs, const double *ys, int n, double x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xs[mid] <= x) lo = mid; else hi = mid; }
  const double f = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
  return (1.0 - f) * ys[lo] + f * ys[lo + 1];
}

}  // namespace

// phase 1: candidate selection + acceptance test (reads the source species only, leaves the collider list on the
// device); phase 2: kinematics of the listed colliders (writes velocities, appends ionisation products);
// 3: both.  iskb_step runs phase 1 of the NEXT step right after the source species has been advanced, on a side
// stream next to the other species' advance, so only the kinematics remain between two steps.
// a phase 1 that ran ahead (iskb_step) is dropped: the context stream waits for it, its counts are forgotten
int32_t mcc_discard_pre(iskb_mcc *mc) {
  if (!mc->pre_valid) return ISKB_OK;
  iskb_ctx *c = mc->ctx;
  CU_TRY(cudaStreamWaitEvent(c->stream, mc->ev_pre, 0));
  CU_TRY(cudaMemsetAsync(mc->d_stats + 10, 0, 10 * sizeof(unsigned long long), c->stream));
  mc->pre_valid = false;
  return ISKB_OK;
}

int32_t mcc_launch(iskb_mcc *mc, double dt, bool count_nu, cudaStream_t st, int phase) {
  iskb_ctx *c = mc->ctx;
  if (!st) st = c->stream;
  iskb_species *src = mc->source;
  const int N = mc->N;
  const double max_Pt = 1.0 - exp(-mc->max_n0 * mc->max_sigma_g * dt);     // mcc.jl:243
  if (max_Pt > 1.0 / N)
    return iskb_fail(ISKB_E_PMAX, "Maximum probability (%g) is greater than 1/%d", max_Pt, N);   // :244-246
  MccDev m;
  m.src = spdev(src);
  if (mc->tq != 0.0) ISKB_TRY(fields_join(c));   // charged target: the test kernel reads E (mcc.jl:266-267)
  int nprod = 0;
  for (int k = 0; k < N; ++k) {
    const MccProc &p = mc->procs[(size_t)k];
    ProcDev &d = m.proc[k];
    d.kind = p.kind; d.threshold = p.threshold; d.offset = p.offset; d.len = p.len;
    d.prod = -1; d.n_prod = 0;
    if (p.kind == ISKB_MCC_IONIZATION && p.product && p.product != src) {
      if (nprod >= 4) return iskb_fail(ISKB_E_UNSUPPORTED, "too many ionisation products");
      m.prod[nprod] = spdev(p.product);
      d.prod = nprod++;
      d.n_prod = (int)std::rint(src->w0 / p.product->w0);                  // :205-207
      if (phase & 2) p.product->counts_stale = true;
    }
    if (p.kind == ISKB_MCC_IONIZATION && (phase & 2)) src->counts_stale = true;
  }
  m.N = N;
  m.eps = mc->d_eps; m.sig = mc->d_sig; m.tn = mc->d_tn; m.E2 = c->d_E2; m.g = c->g;
  m.tqm = mc->tq / mc->tm;
  m.vth_t = sqrt(2 * KB_MCC * mc->tT / mc->tm);
  m.mr1 = src->m / (src->m + mc->tm);
  m.mr2 = mc->tm / (src->m + mc->tm);
  m.m_eV = mc->m_eV;
  m.dt = dt;
  m.sup_total = mc->sup_total;
  m.sig_last_total = 0.0;
  for (int k = 0; k < N; ++k) m.sig_last_total += mc->sig_last[(size_t)k];
  m.n0max = mc->max_n0;
  m.p_cand = std::fmin(1.0, std::fmax(0.0, mc->max_n0 * mc->sup_total * dt));   // the device adds the speed bound (k_snapshot_begin)
  m.need_cell = (mc->uniform_n && !count_nu && mc->tq == 0.0) ? 0 : 1;
  m.vmax2 = src->d_vmax2;
  if (!mc->d_pk) CU_TRY(cudaMalloc(&mc->d_pk, 8 * sizeof(double)));
  if (mc->max_n0 < 0.0) m.p_cand = 0.0;
  m.pk_dev = mc->d_pk;
  m.k0 = (uint32_t)mc->seed;
  m.k1 = (uint32_t)(mc->seed >> 32) ^ (0x9E3779B9u * (uint32_t)(c->rank + 1));
  if (phase & 1) mc->cur_call = mc->calls++;
  m.call = (uint32_t)mc->cur_call;
  // phase 1 alone counts into the second half of d_stats; the numbers join the totals when phase 2 runs
  m.stats = phase == 1 ? mc->d_stats + 10 : mc->d_stats;
  m.status = c->d_status;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  m.nu = nullptr;
  if (count_nu) {
    if (!mc->d_nu) CU_TRY(cudaMalloc(&mc->d_nu, nn * N * sizeof(float)));
    if (phase & 1) CU_TRY(cudaMemsetAsync(mc->d_nu, 0, nn * N * sizeof(float), st));
    m.nu = mc->d_nu;
  }
  // candidate / collider lists (lazy; sized for the worst case of every row being a candidate)
  if (!mc->d_cand) {
    CU_TRY(cudaMalloc(&mc->d_cand, src->cap * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&mc->d_coll, src->cap * sizeof(uint2)));
    CU_TRY(cudaMalloc(&mc->d_lists_cnt, 2 * sizeof(unsigned int)));
  }
  const unsigned int cand_cap = (unsigned int)src->cap;
  const int64_t bound = src->counts_stale ? src->cap : src->h_nslots;
  int64_t exp_cand = (int64_t)(m.p_cand * (double)bound * 1.25) + 1024;
  if (phase & 1) {
  k_snapshot_begin<<<1, 1, 0, st>>>(src->d_cnt, mc->d_lists_cnt, m);
  LAUNCH_CHECK(c);
  int64_t blocks = (bound / 4 + TPB) / TPB;
  if (blocks > (int64_t)c->n_sm * 8) blocks = (int64_t)c->n_sm * 8;
  if (blocks < 1) blocks = 1;
  if (!(m.p_cand > 0.0) || m.p_cand >= 1.0) {   // degenerate probabilities: one draw per row
    k_mcc_select<<<(int)blocks, TPB, 0, st>>>(m, mc->d_lists_cnt, mc->d_cand, cand_cap);
  } else {
    int64_t bs = (bound / SKIP_ROWS + TPB) / TPB;
    if (bs > (int64_t)c->n_sm * 8) bs = (int64_t)c->n_sm * 8;
    if (bs < 1) bs = 1;
    k_mcc_select_skip<<<(int)bs, TPB, 0, st>>>(m, mc->d_lists_cnt, mc->d_cand, cand_cap);
  }
  LAUNCH_CHECK(c);
  // the list lengths live on the device: the dense phases are sized from the expected candidate count
  int64_t b2 = (exp_cand + TEST_TPB - 1) / TEST_TPB;
  if (b2 > (int64_t)c->n_sm * 256) b2 = (int64_t)c->n_sm * 256;
  int ntab = 0;
  for (const MccProc &p : mc->procs) ntab += p.len;
  if (ntab > TEST_SMEM_TABLE_MAX) ntab = 0;
  const size_t tab_bytes = (size_t)ntab * 2 * sizeof(double);
  if (tab_bytes > 40 * 1024)
    CU_TRY(cudaFuncSetAttribute(k_mcc_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
  k_mcc_test<<<(int)b2, TEST_TPB, tab_bytes, st>>>(m, mc->d_lists_cnt, mc->d_cand, cand_cap, mc->d_coll, ntab);
  LAUNCH_CHECK(c);
  }
  if (!(phase & 2)) return ISKB_OK;
  if (phase == 2) {
    k_commit_stats<<<1, 32, 0, st>>>(mc->d_stats);
    LAUNCH_CHECK(c);
  }
  // one thread per collider: the count lives on the device, so the grid covers the expected candidates (most of them
  // collide now that the candidate rate is the tight bound); the kinematics are a long dependent chain, a smaller grid
  // walking the list in rounds is latency bound.  Carrying the velocities from the test phase in the list instead of
  // re-reading three random lines per collider cut the DRAM reads by two thirds and changed nothing in the time.
  int64_t b3 = (exp_cand + 127) / 128;
  if (b3 > (int64_t)c->n_sm * 256) b3 = (int64_t)c->n_sm * 256;
  if (b3 < 1) b3 = 1;
  k_mcc_collide<<<(int)b3, 128, 0, st>>>(m, mc->d_lists_cnt, mc->d_coll);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

extern "C" int32_t iskb_mcc_create(iskb_ctx *c, iskb_species *source, double target_q, double target_m,
                                   double target_T, const double *target_n, int32_t n_proc, const int32_t *kind,
                                   const double *threshold, const int32_t *table_len, const double *eps,
                                   const double *sigma, iskb_species *const *ion_product, uint64_t seed,
                                   iskb_mcc **out) {
  if (!c || !c->has_grid || !source || !out || n_proc < 1 || n_proc > 8 || !target_n || !kind || !table_len ||
      !eps || !sigma)
    return iskb_fail(ISKB_E_INVALID, "iskb_mcc_create: bad arguments (1..8 processes, grid set first)");
  iskb_mcc *mc = new iskb_mcc();
  mc->ctx = c;
  mc->source = source;
  mc->tq = target_q; mc->tm = target_m; mc->tT = target_T;
  mc->N = n_proc;
  mc->seed = seed;
  int off = 0;
  for (int k = 0; k < n_proc; ++k) {
    if (table_len[k] < 2) { delete mc; return iskb_fail(ISKB_E_INVALID, "cross-section table needs >= 2 rows"); }
    MccProc p;
    p.kind = kind[k];
    p.threshold = threshold ? threshold[k] : 0.0;
    p.offset = off;
    p.len = table_len[k];
    p.product = ion_product ? ion_product[k] : nullptr;
    mc->procs.push_back(p);
    off += table_len[k];
  }
  // MonteCarloCollisions(collisions)  mcc.jl:27-51
  mc->m_eV = source->m / QE_MCC;
  std::vector<double> e(1, 0.0);
  e.insert(e.end(), eps, eps + off);
  std::sort(e.begin(), e.end());
  e.erase(std::unique(e.begin(), e.end()), e.end());
  const double alpha = sqrt(2.0 / mc->m_eV);
  double best = -INFINITY;
  for (double ee : e) {
    const double v = alpha * sqrt(ee);
    double sg = 0.0;
    for (const MccProc &p : mc->procs) sg += xsec_eval_host(eps + p.offset, sigma + p.offset, p.len, ee) * v;
    best = std::fmax(best, sg);
  }
  mc->max_sigma_g = best;
  // per-process supremum of sigma_k(eps)*alpha*sqrt(eps) on [0, eps_hi] (for the pruning bound)
  mc->eps_hi = e.back();
  mc->sup_sigma_g.assign((size_t)n_proc, 0.0);
  for (int k = 0; k < n_proc; ++k) {
    const MccProc &p = mc->procs[(size_t)k];
    double sup = 0.0;
    auto f = [&](double ee) { return xsec_eval_host(eps + p.offset, sigma + p.offset, p.len, ee) * alpha * sqrt(ee); };
    for (size_t q = 0; q + 1 < e.size(); ++q) {
      const double e0 = e[q], e1 = e[q + 1];
      sup = std::fmax(sup, std::fmax(f(e0), f(e1)));
      // sigma is linear on [e0,e1] (union grid contains every knot): s = a + b*eps
      const double s0 = xsec_eval_host(eps + p.offset, sigma + p.offset, p.len, e0);
      const double s1 = xsec_eval_host(eps + p.offset, sigma + p.offset, p.len, e1);
      const double b = (s1 - s0) / (e1 - e0), a = s0 - b * e0;
      if (b < 0.0 && a > 0.0) {
        const double es = -a / (3.0 * b);
        if (es > e0 && es < e1) sup = std::fmax(sup, f(es));
      }
    }
    mc->sup_sigma_g[(size_t)k] = sup;
    mc->sig_last.push_back(sigma[p.offset + p.len - 1]);   // sigma_k(last knot): Flat() beyond the table, so sigma*g <= sig_last*|v|max there
  }
  {   // supremum of the SUM over the processes (the candidate probability of the null-collision selection)
    auto tot = [&](double ee) {
      double sg = 0.0;
      for (const MccProc &p : mc->procs) sg += xsec_eval_host(eps + p.offset, sigma + p.offset, p.len, ee);
      return sg;
    };
    double sup = 0.0;
    for (size_t q = 0; q + 1 < e.size(); ++q) {
      const double e0 = e[q], e1 = e[q + 1], s0 = tot(e0), s1 = tot(e1);
      sup = std::fmax(sup, std::fmax(s0 * alpha * sqrt(e0), s1 * alpha * sqrt(e1)));
      const double b = (s1 - s0) / (e1 - e0), a = s0 - b * e0;   // every sigma_k is linear on [e0, e1]
      if (b < 0.0 && a > 0.0) {
        const double es = -a / (3.0 * b);
        if (es > e0 && es < e1) sup = std::fmax(sup, (a + b * es) * alpha * sqrt(es));
      }
    }
    mc->sup_total = sup;
  }
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  mc->max_n0 = -INFINITY;
  for (int64_t k = 0; k < nn; ++k) mc->max_n0 = std::fmax(mc->max_n0, target_n[k]);   // :242
  mc->uniform_n = true;
  for (int64_t k = 0; k < nn && mc->uniform_n; ++k) mc->uniform_n = target_n[k] == mc->max_n0;
  CU_TRY(cudaMalloc(&mc->d_tn, nn * sizeof(double)));
  CU_TRY(cudaMalloc(&mc->d_eps, off * sizeof(double)));
  CU_TRY(cudaMalloc(&mc->d_sig, off * sizeof(double)));
  CU_TRY(cudaMalloc(&mc->d_stats, 2 * (2 + 8) * sizeof(unsigned long long)));
  CU_TRY(cudaEventCreateWithFlags(&mc->ev_pre, cudaEventDisableTiming));
  CU_TRY(cudaMemcpyAsync(mc->d_tn, target_n, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(mc->d_eps, eps, off * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(mc->d_sig, sigma, off * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemsetAsync(mc->d_stats, 0, 2 * (2 + 8) * sizeof(unsigned long long), c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->mccs.push_back(mc);
  *out = mc;
  return ISKB_OK;
}

extern "C" int32_t iskb_mcc_constants(iskb_mcc *mc, double *max_sigma_g, double *m_eV) {
  if (!mc) return iskb_fail(ISKB_E_INVALID, "null mcc");
  if (max_sigma_g) *max_sigma_g = mc->max_sigma_g;
  if (m_eV) *m_eV = mc->m_eV;
  return ISKB_OK;
}

static int32_t read_stats(iskb_mcc *mc, unsigned long long *h) {
  iskb_ctx *c = mc->ctx;
  CU_TRY(cudaMemcpyAsync(h, mc->d_stats, (2 + 8) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}

extern "C" int32_t iskb_mcc_perform(iskb_mcc *mc, double dt, double *nu_out, int64_t *n_candidates,
                                    int64_t *n_collisions) {
  if (!mc) return iskb_fail(ISKB_E_INVALID, "null mcc");
  iskb_ctx *c = mc->ctx;
  unsigned long long before[10], after[10];
  ISKB_TRY(mcc_discard_pre(mc));
  ISKB_TRY(read_stats(mc, before));
  ISKB_TRY(mcc_launch(mc, dt, nu_out != nullptr));
  ISKB_TRY(read_stats(mc, after));
  if (n_candidates) *n_candidates = (int64_t)(after[0] - before[0]);
  if (n_collisions) *n_collisions = (int64_t)(after[1] - before[1]);
  if (nu_out) {
    const int64_t tot = (int64_t)c->g.nx * c->g.ny * mc->N;
    std::vector<float> h((size_t)tot);
    CU_TRY(cudaMemcpyAsync(h.data(), mc->d_nu, tot * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (int64_t k = 0; k < tot; ++k) nu_out[k] = h[(size_t)k];
  }
  return ctx_check_status(c);
}

extern "C" int32_t iskb_mcc_totals(iskb_mcc *mc, int64_t *out) {
  if (!mc || !out) return iskb_fail(ISKB_E_INVALID, "null");
  unsigned long long h[10];
  ISKB_TRY(read_stats(mc, h));
  for (int k = 0; k < 2 + mc->N; ++k) out[k] = (int64_t)h[k];
  return ISKB_OK;
}
