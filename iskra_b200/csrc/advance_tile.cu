// Tile-aware fused advance! + density with an INCREMENTAL re-group folded into the same pass.
//   gather E (cloud_in_cell.jl:20-36) -> push (pushers.jl:8-17,37-50) -> after_push (wrap.jl:1-33)
//   -> CIC deposit of wg (cloud_in_cell.jl:1-18), 88 B per particle-step.
//
// Row layout (DESIGN.md "tile directory"): rows [0, ns) are grouped by 8x8-cell tile in tile-ordinal order,
// ts[t] is the first row of tile t (ts[ntiles] = ns); rows [ns, n) are the unsorted TAIL (rows appended by
// ionisation, rows that left their tile's neighbourhood).  A row may lag behind: it is STORED in tile S but its
// position is in a neighbour of S -- the windows below have a 3.5-cell margin for exactly that.
//
// Every warp owns whole tiles (wr[w] .. wr[w+1]).  Per tile it anchors its two private shared-memory windows
// (16x16 nodes: E as double2, rho accumulator) on the tile and walks the tile's rows in 32-row batches:
// exact cell index -> gather from the shared E window -> push -> boundary -> store -> cell of the new position
// -> deposit into the shared rho window in claim rounds (advance_fused.cu describes why not atomics).
// Rows whose cell lies outside the window go to a list that k_advance_list advances from global memory.
//
// MARK (every launch): for each row the kernel records where its NEW position lies relative to its storage
// tile (1 byte, tiles.cuh) and counts rows per (storage tile, code).  MOVE (a launch that re-groups): the counts
// of the previous launch give, after one small scan (k_regroup_*), the destination of every row in a layout that is
// grouped by the tile of its CURRENT position; the kernel then writes its results (and wg, id) there instead of in
// place.  The rows of a destination tile are ordered by (source tile in a fixed neighbour order, previous row
// order): a pure function of the previous layout, no atomics involved in the placement.  This replaces the
// radix sort + 104 B/row permutation pass of a re-group by +16 B/row on the launch that moves.
#include <cstdlib>
#include <cstring>

#include "tiles.cuh"

namespace {

constexpr int WE = 16;    // window nodes per side (tile 9 nodes + 3.5 cells of margin each side)
constexpr int WRS = 20;   // row stride of the rho window in doubles.  64-bit shared accesses are served per half warp (16 lanes,
                          // 16 bank pairs); in the interleaved row order of a full sort (sort.cu: cells of one parity, row by row)
                          // 16 consecutive rows sit in 4 cell rows x 4 cells of alternating column parity, and with a row
                          // stride of 4 bank pairs (20 mod 16) those 16 nodes fall into 16 different bank pairs.  (Stride 16:
                          // 4-way conflicts; 18: 2-way.  The E window, double2 with stride 16, is conflict free per quarter warp.)

struct TileArgs {
  double *col[6];           // x y vx vy vz wg  (read; written in place unless MOVE)
  uint32_t *id;
  uint8_t *code;
  double *ocol[6];          // MOVE: destination buffers
  uint32_t *oid;
  uint8_t *ocode;
  int64_t *cnt;
  GridDev g;
  const double2 *E2;
  double qm, hqm, dt, c1, w0;   // q/m, 0.5*(q/m), dt, (0.5dt)*(q/m), default weight
  long long *ufix;          // deposited weights, fixed point (see add_fixed)
  double fscale;
  int *status;
  unsigned long long *vmax2;
  const uint32_t *ts;       // [ntiles + 1]
  const uint32_t *wr;       // [n_warps + 1] first tile of every warp
  uint32_t ntiles, mtx, tiles_x, tiles_y;
  uint32_t *tcnt;           // [ntiles * NCODE] rows per (storage tile, code) after this launch
  const uint32_t *tbase;    // MOVE: [ntiles * NCODE] first destination row per (storage tile, code)
  const uint32_t *seg;      // MOVE: scan results (see k_regroup_seg): [2*ntiles] = tail destination
  const unsigned long long *vz2max;   // LEAN: bits of the kept bound of v_z^2
  unsigned *ticket;         // chunk dispenser
  unsigned nchunks;
  // MOVE with tail merge: destination and key (tile, ntiles = stays in the tail, ntiles + 1 = dead) of tail row ns + k
  const uint32_t *taildst, *tailkey, *tstart, *tailbase, *tn;
  int mark;                 // write the codes and counts a MOVE needs: only on the launch before one
  uint2 *mlist;             // rows left to k_advance_list: (source row, destination row)
  unsigned *mlist_n;
  unsigned mlist_cap;
};

// staged miss-list entries: flushed MQ_FLUSH or more at a time.  In place the destination of a row is the row itself: one
// word per entry.
template <bool MOVE> struct Mq { static constexpr int FLUSH = 16, CAP = FLUSH + 31 + 1; typedef uint2 T; };
template <> struct Mq<false> { static constexpr int FLUSH = 8, CAP = FLUSH + 31 + 1; typedef unsigned T; };
__device__ __forceinline__ void mq_put(uint2 &e, unsigned row, unsigned dest) { e = make_uint2(row, dest); }
__device__ __forceinline__ void mq_put(unsigned &e, unsigned row, unsigned) { e = row; }
__device__ __forceinline__ uint2 mq_get(const uint2 &e) { return e; }
__device__ __forceinline__ uint2 mq_get(const unsigned &e) { return make_uint2(e, e); }
constexpr int NTC_INPLACE = 12;   // codes 0 .. CODE_DEAD
template <bool MOVE>
struct WarpSm {
  double2 E[WE * WE];
  double rho[WE * WRS];
  unsigned char claim[(WE - 1) * WE];   // cells: rows 0 .. WE-2
  // [code of the tile the row is stored in after the launch][code of its new position]; in-place launches: row CODE_STAY only
  unsigned tc[MOVE ? 9 * NCODE : NTC_INPLACE];
  unsigned mv[MOVE ? NCODE : 1];   // MOVE: next destination row per code of the current tile
  unsigned stats[4];      // window misses, deposits outside the window, tiles, discards
  unsigned misc[8];       // first tile of the next chunk, last readable row, tile coordinates of the current tile, staged misses
  typename Mq<MOVE>::T mq[Mq<MOVE>::CAP];
};   // 7.0 KB in place / 8.0 KB MOVE: three CTAs of 8 warps per SM either way
static_assert(sizeof(WarpSm<true>) * 8 + 1024 <= 233472 / 3, "three CTAs per SM");
static_assert(CODE_DEAD < NTC_INPLACE, "in-place counters");

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ int clamp_origin(int o, int n, int wn) {
  if (o > n - wn) o = n - wn;
  return o < 0 ? 0 : o;
}

// Everything that leaves a warp goes into a 64-bit FIXED-POINT accumulator (value * fscale, rounded to nearest): integer
// addition is associative, so the deposited weights do not depend on the order in which warps, tiles, list rows --
// or ranks, the all-reduce is an integer sum too -- arrive (rho reproducible from run to run, identical on
// every rank).  The sums INSIDE a warp are FP64, so a different split of the rows over warps or ranks still differs by their
// rounding (2e-13 of the largest node value between one and two GPUs, tests/multi_gpu).  fscale is a power of two chosen on the host such that the sum over ALL slots of all species fits
// (api.cu): at the C5 shard one unit is 1.5e-11 of a particle weight, 1e-12 of a typical node value.
__device__ __forceinline__ void add_fixed(long long *ufix, int64_t node, double v, double fscale) {
  atomicAdd((unsigned long long *)&ufix[node], (unsigned long long)__double2ll_rn(v * fscale));
}

__device__ __forceinline__ void flush_rho(double *rho, int i0, int j0, int nx, long long *ufix, double fscale, int lane) {
#pragma unroll
  for (int k = 0; k < WE * WE / 32; ++k) {
    const int e = k * 32 + lane;
    const int o = (e >> 4) * WRS + (e & 15);
    const double v = rho[o];
    if (v != 0.0) {
      add_fixed(ufix, (int64_t)(i0 + (e & 15)) + (int64_t)(j0 + (e >> 4)) * nx, v, fscale);
      rho[o] = 0.0;
    }
  }
}
__device__ __forceinline__ void load_E(double2 *sE, int i0, int j0, int nx, const double2 *__restrict__ E2, int lane) {
#pragma unroll
  for (int k = 0; k < WE * WE / 32; ++k) {
    const int e = k * 32 + lane;
    sE[e] = __ldg(&E2[(int64_t)(i0 + (e & 15)) + (int64_t)(j0 + (e >> 4)) * nx]);
  }
}

// code of tile (ntx, nty) seen from tile (stx, sty)
__device__ __forceinline__ unsigned rel_code(int ntx, int nty, int stx, int sty) {
  const unsigned ddx = (unsigned)(ntx - stx + 1), ddy = (unsigned)(nty - sty + 1);
  return (ddx < 3u && ddy < 3u) ? ddy * 3u + ddx : (unsigned)CODE_FAR;
}

// One particle, from registers to registers: gather is the caller's business (shared window or global memory).
// Returns true when the row was discarded.
template <int MX, int MY, bool RZ, bool LEAN>
__device__ __forceinline__ bool push_and_bound(double &px, double &py, double &vx, double &vy, double &vz, double ex,
                                               double ey, const TileArgs &a) {
  vx = push_v_h(vx, ex, a.c1, a.hqm, a.dt);
  vy = push_v_h(vy, ey, a.c1, a.hqm, a.dt);
  if (!LEAN) vz = push_v_h(vz, 0.0, a.c1, a.hqm, a.dt);   // v_z + 0: LEAN leaves the column alone
  px = push_x(px, vx, a.dt);
  if (RZ) to_cylindrical(px, vx, vz, a.dt);   // push_particles!(::BorisPusher{:rz}, ...)  pushers.jl:13-17
  py = push_x(py, vy, a.dt);
  bool dead = (MX == ISKB_BND_DISCARD) && boundary_axis(px, a.g.ox, a.g.Lx, MX);
  if (!dead) dead = (MY == ISKB_BND_DISCARD) && boundary_axis(py, a.g.oy, a.g.Ly, MY);
  if (!dead) {
    if (MX == ISKB_BND_WRAP) boundary_axis(px, a.g.ox, a.g.Lx, MX);
    if (MY == ISKB_BND_WRAP) boundary_axis(py, a.g.oy, a.g.Ly, MY);
  }
  return dead;
}

#define OUTC(q) (MOVE ? a.ocol[q] : a.col[q])   // column the launch writes to

// LEAN: the launch neither reads nor writes what cannot change.  With B == 0 (generalized_poisson.jl:412-419) and
// E_z == 0 (:398-410) push_in_cartesian! leaves v_z as it is (v_z + 0, pushers.jl:41-48), so the column is not
// touched (a re-grouping launch still has to carry it along); and a species whose weights are all w0 (every
// BASELINE config: configuration.jl:99 `ones(N) * weight`) needs no wg column traffic either.  72 -> 64 B per row
// instead of 88; positions, velocities and rho are bit-identical to the full path (tests/test_gpu_tile.py).
template <int MX, int MY, bool MOVE, bool RZ, bool LEAN>
__global__ void __launch_bounds__(256, 3) k_advance_tile(const TileArgs a) {
  extern __shared__ double2 s_dyn[];
  const int lane = threadIdx.x & 31;
  WarpSm<MOVE> &sm = ((WarpSm<MOVE> *)s_dyn)[threadIdx.x >> 5];
  constexpr int TCS = MOVE ? CODE_STAY * NCODE : 0;   // counters of the rows that are stored in this tile after the launch
  for (int e = lane; e < WE * WRS; e += 32) sm.rho[e] = 0.0;
  for (int e = lane; e < (MOVE ? 9 * NCODE : NTC_INPLACE); e += 32) sm.tc[e] = 0;
  if (MOVE && lane < NCODE) sm.mv[lane] = 0;
  if (lane < 4) sm.stats[lane] = 0;
  // The loop-carried state is kept small (tile, its row range, the batch, the window origin, a float velocity
  // bound); everything else that is constant per warp or per tile sits in shared memory: the body must fit
  // 80 registers without spills (a spill reload shares its scoreboard with the row prefetch).
  if (lane == 0) {
    sm.misc[1] = (unsigned)a.cnt[CNT_NSLOTS] - 1u;   // last row that may be read
    sm.misc[4] = 0;
  }
  __syncwarp();
  float vm2 = 0.0f;
  constexpr int NOT_ANCHORED = -(1 << 20);
  int ei0 = NOT_ANCHORED, ej0 = 0;
  // Chunks of whole tiles (~2000 rows, wr[]) are handed out through a ticket: the grid is persistent (3 CTAs per
  // SM) and no warp waits for the slowest sibling of its CTA.
  for (;;) {
    unsigned chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.ticket, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= a.nchunks) break;
    unsigned t = a.wr[chunk];
    if (lane == 0) sm.misc[0] = a.wr[chunk + 1];   // first tile of the next chunk
    __syncwarp();
    unsigned r0 = 0, r1 = 0;
    for (; t < sm.misc[0]; ++t) {   // first non-empty tile
      r0 = a.ts[t];
      r1 = a.ts[t + 1];
      if (r1 > r0) break;
    }
    if (t >= sm.misc[0]) continue;
    // Register prefetch, one batch ahead, issued ONLY inside the loop: the loop starts one (empty) batch before the
    // first real one, with dead rows in the registers.  If the first batch were loaded ahead of the loop, the
    // body's first use of px would have to wait on those loads' scoreboard -- and ptxas attaches the in-loop
    // prefetch to the same scoreboard, so every iteration would wait for the loads it has just issued
    // (measured: stall_long_sb 5.2 per issue, 1.49 instead of 1.09 ms per launch).
    unsigned rb = (r0 & ~31u) - 32u;   // may wrap: the rows of this batch are never valid
    const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
    double nx_ = nan_, ny_ = 0.0, nvx_ = 0.0, nvy_ = 0.0, nvz_ = 0.0, nwq_ = 0.0;
    uint32_t nid_ = 0;
    unsigned ncode_ = 0;
    for (;;) {
      double px = nx_, py = ny_, vx = nvx_, vy = nvy_, vz = nvz_;
      const double wq = LEAN ? a.w0 : nwq_;
      const uint32_t pid = nid_;
      const unsigned pcode = ncode_;
      // ---- prefetch of the next batch (next tile: the batch that holds its first row) ----
      unsigned rn = rb + 32;
      if (rn >= r1) {
        for (unsigned nt = t + 1; nt < sm.misc[0]; ++nt) {
          const unsigned q0 = a.ts[nt];
          if (a.ts[nt + 1] > q0) { rn = q0 & ~31u; break; }
        }
      }
      rn = min(rn + lane, sm.misc[1]);   // index clamped instead of branching
      nx_ = a.col[0][rn]; ny_ = a.col[1][rn]; nvx_ = a.col[2][rn]; nvy_ = a.col[3][rn];
      if (!LEAN || MOVE) nvz_ = a.col[4][rn];
      if (!LEAN) nwq_ = a.col[5][rn];
      if (MOVE) { nid_ = a.id[rn]; ncode_ = a.code[rn]; }
      {
        // pull the rows three batches further into L2: one batch of register prefetch does not always cover DRAM under
        // the mixed read / write traffic (-0.03 ms per step at the C5 shard).  Tiles are stored back to back, so "the
        // rows after these" is right except at the end of the warp's chunk.
        const unsigned rp = min(rn + 96u, sm.misc[1]);
        prefetch_l2(&a.col[0][rp]); prefetch_l2(&a.col[1][rp]); prefetch_l2(&a.col[2][rp]); prefetch_l2(&a.col[3][rp]);
        if (!LEAN || MOVE) prefetch_l2(&a.col[4][rp]);
        if (!LEAN) prefetch_l2(&a.col[5][rp]);
        if (MOVE && (lane & 1) == 0) prefetch_l2(&a.id[rp]);
        if (MOVE && (lane & 7) == 0) prefetch_l2(&a.code[rp]);
      }

      if (ei0 == NOT_ANCHORED) {   // first batch of a tile: anchor the windows on it
        int stx, sty;
        tile_coords(t, a.mtx, stx, sty);
        ei0 = clamp_origin(stx * 8 - (WE - 9) / 2, a.g.nx, WE);
        ej0 = clamp_origin(sty * 8 - (WE - 9) / 2, a.g.ny, WE);
        load_E(sm.E, ei0, ej0, a.g.nx, a.E2, lane);
        if (MOVE && lane < NCODE) sm.mv[lane] = a.tbase[t * NCODE + lane];
        if (lane == 0) {
          sm.stats[2] += 1;
          sm.misc[2] = (unsigned)stx;
          sm.misc[3] = (unsigned)sty;
        }
        __syncwarp();
      }

      const unsigned row = rb + lane;
      const bool valid = row >= r0 && row < r1;
      const bool live = valid && !is_dead(px);
      // cell of the old position (dead rows: NaN -> cell 0, fits nowhere); live rows outside the window -- outside
      // the grid included -- are left to k_advance_list
      int i, j;
      double hx, hy;
      cell1_fast(px, a.g.dx, a.g.rdx, i, hx);
      cell1_fast(py, a.g.dy, a.g.rdy, j, hy);
      const bool fit = live && (unsigned)(i - 1 - ei0) < (unsigned)(WE - 1) && (unsigned)(j - 1 - ej0) < (unsigned)(WE - 1);
      const bool miss = live && !fit;

      // ---- MOVE: destination row = base(storage tile, code) + rank among the tile's rows with that code ----
      unsigned dest = row;
      unsigned scode = CODE_STAY;   // code of the row in the layout it is written to
      if (MOVE) {
        // the stored code decides (the counts were taken with it): a row that was discarded by k_advance_list
        // still carries CODE_FAR and goes to the tail as a dead row
        scode = valid ? pcode : 31u;
        {   // most rows stay: one ballot
          const unsigned m = __ballot_sync(0xffffffffu, scode == (unsigned)CODE_STAY);
          const unsigned b = sm.mv[CODE_STAY];
          if (scode == (unsigned)CODE_STAY) dest = b + __popc(m & lanemask_lt());
          __syncwarp();
          if (lane == 0) sm.mv[CODE_STAY] = b + __popc(m);
        }
        unsigned rem = __ballot_sync(0xffffffffu, valid && scode != (unsigned)CODE_STAY);
        while (rem) {
          const int ld = __ffs(rem) - 1;
          const unsigned cl = __shfl_sync(0xffffffffu, scode, ld);
          const unsigned m = __ballot_sync(0xffffffffu, scode == cl);
          const unsigned b = sm.mv[cl];
          if (scode == cl) dest = b + __popc(m & lanemask_lt());
          __syncwarp();
          if (lane == ld) sm.mv[cl] = b + __popc(m);
          rem &= ~m;
        }
        __syncwarp();
      }

      bool dead_now = false, dep_win = false;
      double d00 = 0, d10 = 0, d01 = 0, d11 = 0;
      int ci = 0;
      unsigned ncode = CODE_FAR;
      if (fit) {
        double ex, ey;
        {
          const CicW gw = cic_weights(hx, hy);
          const int o = (j - 1 - ej0) * WE + (i - 1 - ei0);
          const double2 e00 = sm.E[o], e10 = sm.E[o + 1], e01 = sm.E[o + WE], e11 = sm.E[o + WE + 1];
          ex = cic_gather(gw, e00.x, e10.x, e01.x, e11.x);
          ey = cic_gather(gw, e00.y, e10.y, e01.y, e11.y);
        }
        const bool dead = push_and_bound<MX, MY, RZ, LEAN>(px, py, vx, vy, vz, ex, ey, a);
        if (LEAN) vm2 = fmaxf(vm2, __double2float_ru(fma(vy, vy, vx * vx)));
        else vm2 = fmaxf(vm2, __double2float_ru(fma(vz, vz, fma(vy, vy, vx * vx))));
        OUTC(2)[dest] = vx; OUTC(3)[dest] = vy; OUTC(1)[dest] = py;
        if (!LEAN || MOVE) OUTC(4)[dest] = vz;
        if (MOVE) {
          if (!LEAN) a.ocol[5][dest] = wq;
          a.oid[dest] = pid;
        }
        if (dead) {
          OUTC(0)[dest] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
          ncode = CODE_DEAD;
        } else {
          OUTC(0)[dest] = px;
          cell1_fast(px, a.g.dx, a.g.rdx, i, hx);
          cell1_fast(py, a.g.dy, a.g.rdy, j, hy);
          if (cell_in_grid(i, j, a.g.nx, a.g.ny)) {
            const CicW cw = cic_weights(hx, hy);
            d00 = __dmul_rn(cw.w00, wq); d10 = __dmul_rn(cw.w10, wq);
            d01 = __dmul_rn(cw.w01, wq); d11 = __dmul_rn(cw.w11, wq);
            // code of the new position seen from the tile the row is stored in after this launch
            int dtx = (int)sm.misc[2], dty = (int)sm.misc[3];
            if (MOVE && scode < 9u) { dtx += (int)(scode % 3u) - 1; dty += (int)(scode / 3u) - 1; }
            if (a.mark) ncode = rel_code((i - 1) >> 3, (j - 1) >> 3, dtx, dty);
            const int ri = i - 1 - ei0, rj = j - 1 - ej0;
            if ((unsigned)ri < (unsigned)(WE - 1) && (unsigned)rj < (unsigned)(WE - 1)) {
              ci = rj * WE + ri;
              dep_win = true;
            } else {
              const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * a.g.nx;
              add_fixed(a.ufix, n00, d00, a.fscale);
              add_fixed(a.ufix, n00 + 1, d10, a.fscale);
              add_fixed(a.ufix, n00 + a.g.nx, d01, a.fscale);
              add_fixed(a.ufix, n00 + a.g.nx + 1, d11, a.fscale);
              atomicAdd(&sm.stats[1], 1u);
            }
          } else {
            atomicOr(a.status, ISKB_ST_OOB);
          }
        }
      } else if (valid && !live) {
        ncode = CODE_DEAD;
        if (MOVE) {   // parked behind the live rows with its id; the slot's weight is reset like remove! does (kinetic.jl:24)
          a.ocol[0][dest] = __longlong_as_double(0x7ff8000000000000LL);
          if (!LEAN) a.ocol[5][dest] = a.w0;
          a.oid[dest] = pid;
        }
      }
      // rows outside the window: k_advance_list advances them; they join the tail at the next re-group.  The entries
      // are staged per warp and appended 16 or more at a time: one same-address atomic per batch cost 2-4 ms per launch
      // once most batches held a miss (same-address atomics serialise at ~2 ns each on B200).
      {
        const unsigned mm = __ballot_sync(0xffffffffu, miss);
        if (mm) {
          unsigned q = sm.misc[4];
          if (miss) mq_put(sm.mq[q + __popc(mm & lanemask_lt())], row, dest);
          q += __popc(mm);
          __syncwarp();
          if (q >= (unsigned)Mq<MOVE>::FLUSH) {
            unsigned b = 0;
            if (lane == 0) b = atomicAdd(a.mlist_n, q);
            b = __shfl_sync(0xffffffffu, b, 0);
            for (unsigned e = lane; e < q; e += 32) {
              if (b + e < a.mlist_cap) a.mlist[b + e] = mq_get(sm.mq[e]);
              else atomicOr(a.status, ISKB_ST_CAPACITY);
            }
            q = 0;
            __syncwarp();
          }
          if (lane == 0) { sm.misc[4] = q; sm.stats[0] += __popc(mm); }
          __syncwarp();
        }
      }
      // ---- MARK: code of the new position + counts per (tile the row is stored in, code); only the launch before a
      // MOVE (and the MOVE itself) needs them ----
      if (a.mark) {
        if (valid) (MOVE ? a.ocode : a.code)[dest] = (uint8_t)ncode;
        // counters of this tile and, on a MOVE, of its eight neighbours (rows that move there); rows that go to
        // the tail (FAR) or are parked (DEAD) belong to no tile any more
        const bool counted = valid && scode < 9u;
        const bool common = counted && scode == (unsigned)CODE_STAY && ncode == (unsigned)CODE_STAY;
        const unsigned ms = __ballot_sync(0xffffffffu, common);
        if (lane == 0 && ms) sm.tc[TCS + CODE_STAY] += __popc(ms);
        if (counted && !common) atomicAdd(&sm.tc[(MOVE ? scode * NCODE : 0u) + ncode], 1u);
      }
      // ---- deposit rounds without atomics (advance_fused.cu).  The claim is per cell, so the winners' four corner
      // updates hit distinct nodes in every phase.  Measured alternatives (B200, C5 shard, ms per step): one claim round
      // followed by atomicAdd(double) on shared memory -- a compare-and-swap loop on sm_100a -- for the losers +0.11;
      // match.any to merge equal cells in registers first costs ~2 cycles per distinct value (profiles/r1_microbench);
      // one rho window and claim array per HALF warp (2 instead of 2.9 rounds for rows in random order) +0.09. ----
      {
        unsigned pend = __ballot_sync(0xffffffffu, dep_win);
        double *r0p = sm.rho + ci + (WRS - WE) * (ci >> 4);
        while (pend) {
          if (dep_win) sm.claim[ci] = (unsigned char)lane;
          __syncwarp();
          const bool win = dep_win && sm.claim[ci] == (unsigned char)lane;
          if (win) r0p[0] = __dadd_rn(r0p[0], d00);
          __syncwarp();
          if (win) r0p[1] = __dadd_rn(r0p[1], d10);
          __syncwarp();
          if (win) r0p[WRS] = __dadd_rn(r0p[WRS], d01);
          __syncwarp();
          if (win) r0p[WRS + 1] = __dadd_rn(r0p[WRS + 1], d11);
          __syncwarp();
          if (win) dep_win = false;
          pend = __ballot_sync(0xffffffffu, dep_win);
        }
      }
      if (MX == ISKB_BND_DISCARD || MY == ISKB_BND_DISCARD) {
        const unsigned dmask = __ballot_sync(0xffffffffu, dead_now);
        if (dmask && lane == 0) sm.stats[3] += __popc(dmask);
      }
      __syncwarp();
      if (rb + 32 >= r1) {   // tile finished: flush its window, publish its counts, go to the next non-empty tile
        flush_rho(sm.rho, ei0, ej0, a.g.nx, a.ufix, a.fscale, lane);
        ei0 = NOT_ANCHORED;
        if (a.mark)
        for (int e = lane; e < (MOVE ? 9 * NCODE : NTC_INPLACE); e += 32) {
          const int sc = MOVE ? e / NCODE : CODE_STAY, nc = MOVE ? e % NCODE : e;
          const unsigned c = sm.tc[e];
          if (c) {
            const int dtx = (int)sm.misc[2] + sc % 3 - 1, dty = (int)sm.misc[3] + sc / 3 - 1;
            atomicAdd(&a.tcnt[(sc == CODE_STAY ? t : tile_ordinal((uint32_t)dtx, (uint32_t)dty, a.mtx)) * NCODE + nc], c);
            sm.tc[e] = 0;
          }
        }
        __syncwarp();
        const unsigned t_end = sm.misc[0];
        for (++t; t < t_end; ++t) {
          r0 = a.ts[t];
          r1 = a.ts[t + 1];
          if (r1 > r0) break;
        }
        if (t >= t_end) break;
        rb = r0 & ~31u;
      } else {
        rb += 32;
      }
    }
  }
  __syncwarp();
  {   // staged misses left over
    const unsigned q = sm.misc[4];
    if (q) {
      unsigned b = 0;
      if (lane == 0) b = atomicAdd(a.mlist_n, q);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (lane < (int)q) {
        if (b + lane < a.mlist_cap) a.mlist[b + lane] = mq_get(sm.mq[lane]);
        else atomicOr(a.status, ISKB_ST_CAPACITY);
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmaxf(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if (lane == 0) {
    if (vm2 > 0.0f) {
      // float -> double is exact and the float was rounded up: still an upper bound; LEAN adds the kept bound of v_z^2
      double b = (double)vm2;
      if (LEAN) b = __dadd_ru(b, __longlong_as_double((long long)*a.vz2max));
      atomicMax(a.vmax2, (unsigned long long)__double_as_longlong(b));
    }
    if (sm.stats[3]) atomicAdd((unsigned long long *)&a.cnt[CNT_NDEAD], (unsigned long long)sm.stats[3]);
    atomicAdd((unsigned long long *)&a.cnt[3], (unsigned long long)sm.stats[0]);
    atomicAdd((unsigned long long *)&a.cnt[4], (unsigned long long)sm.stats[1]);
    atomicAdd((unsigned long long *)&a.cnt[5], (unsigned long long)sm.stats[2]);
  }
}

// Rows the tiled kernel left out -- its miss list and the unsorted tail [ns, n) -- advanced one per thread straight
// from / to global memory: same arithmetic, E gathered from the global field, deposit with global REDs.
// MOVE: miss rows go where the tiled kernel said; tail row ns + k goes where the tail merge put it (k_tail_dest): into
// the segment of the tile it is in, behind the tile's other arrivals -- the tail never needs a full sort to empty.
template <int MX, int MY, bool MOVE, bool RZ, bool LEAN>
__global__ void __launch_bounds__(256) k_advance_list(const TileArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned nm = min(*a.mlist_n, a.mlist_cap);
  const unsigned n = (unsigned)a.cnt[CNT_NSLOTS], ns = a.ts[a.ntiles];
  const unsigned ntail = n > ns ? n - ns : 0u;
  const unsigned tmerged = MOVE ? *a.tn : 0u;   // tail rows that took part in the merge (the first tmerged ones)
  const unsigned tkeep = MOVE ? a.tailbase[a.ntiles] + (a.tstart[a.ntiles + 1] - a.tstart[a.ntiles]) : 0u;
  const unsigned total = nm + ntail;
  const unsigned total_pad = (total + 31u) & ~31u;
  double vm2 = 0.0;
  unsigned ndead = 0;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total_pad; k += gridDim.x * blockDim.x) {
    bool dead_now = false;
    if (k < total) {
      unsigned src, dst, key = 0xffffffffu;   // key < ntiles: the row joins that tile (MOVE)
      if (k < nm) { const uint2 e = a.mlist[k]; src = e.x; dst = e.y; }
      else {
        const unsigned kt = k - nm;
        src = ns + kt;
        if (!MOVE) dst = src;
        else if (MOVE && kt < tmerged) { dst = a.taildst[kt]; key = a.tailkey[kt]; }
        else dst = tkeep + (kt - tmerged);
      }
      unsigned ncode = CODE_FAR;
      double px = a.col[0][src];
      if (is_dead(px)) {
        if (MOVE) {
          a.ocol[0][dst] = px;
          if (!LEAN) a.ocol[5][dst] = a.col[5][src];
          a.oid[dst] = a.id[src];
        }
      } else {
        double py = a.col[1][src], vx = a.col[2][src], vy = a.col[3][src], vz = (!LEAN || MOVE) ? a.col[4][src] : 0.0;
        const double wq = LEAN ? a.w0 : a.col[5][src];
        double ex, ey;
        if (!gather_E(a.E2, a.g, px, py, ex, ey)) atomicOr(a.status, ISKB_ST_OOB);
        const bool dead = push_and_bound<MX, MY, RZ, LEAN>(px, py, vx, vy, vz, ex, ey, a);
        vm2 = fmax(vm2, LEAN ? fma(vy, vy, vx * vx) : fma(vz, vz, fma(vy, vy, vx * vx)));
        OUTC(2)[dst] = vx; OUTC(3)[dst] = vy; OUTC(1)[dst] = py;
        if (!LEAN || MOVE) OUTC(4)[dst] = vz;
        if (MOVE) {
          if (!LEAN) a.ocol[5][dst] = wq;
          a.oid[dst] = a.id[src];
        }
        if (dead) {
          OUTC(0)[dst] = __longlong_as_double(0x7ff8000000000000LL);
          dead_now = true;
          ncode = CODE_DEAD;
        } else {
          OUTC(0)[dst] = px;
          int i, j;
          double hx, hy;
          cell1(px, a.g.dx, a.g.rdx, a.g.fast_div, i, hx);
          cell1(py, a.g.dy, a.g.rdy, a.g.fast_div, j, hy);
          if (cell_in_grid(i, j, a.g.nx, a.g.ny)) {
            if (MOVE && key < a.ntiles) {
              int stx, sty;
              tile_coords(key, a.mtx, stx, sty);
              ncode = rel_code((i - 1) >> 3, (j - 1) >> 3, stx, sty);
            }
            const CicW cw = cic_weights(hx, hy);
            const int64_t n00 = (int64_t)(i - 1) + (int64_t)(j - 1) * a.g.nx;
            add_fixed(a.ufix, n00, __dmul_rn(cw.w00, wq), a.fscale);
            add_fixed(a.ufix, n00 + 1, __dmul_rn(cw.w10, wq), a.fscale);
            add_fixed(a.ufix, n00 + a.g.nx, __dmul_rn(cw.w01, wq), a.fscale);
            add_fixed(a.ufix, n00 + a.g.nx + 1, __dmul_rn(cw.w11, wq), a.fscale);
          } else {
            atomicOr(a.status, ISKB_ST_OOB);
          }
        }
        if (MOVE && a.mark && key < a.ntiles) {   // MARK of a row that has just joined a tile
          a.ocode[dst] = (uint8_t)ncode;
          atomicAdd(&a.tcnt[key * NCODE + ncode], 1u);
        }
      }
    }
    ndead += __popc(__ballot_sync(0xffffffffu, dead_now));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) vm2 = fmax(vm2, __shfl_xor_sync(0xffffffffu, vm2, d));
  if (lane == 0) {
    if (vm2 > 0.0) {
      // the bound is kept as a float rounded up (the tiled kernel does the same)
      double up = (double)__double2float_ru(vm2);
      if (LEAN) up = __dadd_ru(up, __longlong_as_double((long long)*a.vz2max));
      atomicMax(a.vmax2, (unsigned long long)__double_as_longlong(up));
    }
    if (ndead) atomicAdd((unsigned long long *)&a.cnt[CNT_NDEAD], (unsigned long long)ndead);
  }
}

// ---- tile directory -----------------------------------------------------------------------------
// ts[t] = first row whose sorted cell key is >= t << 6 (keys of a FULL sort, sort.cu); ts[ntiles] = ns.
__global__ void k_tile_starts(const uint32_t *__restrict__ keys, int64_t n, uint32_t ntiles, uint32_t *ts) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
    const uint32_t v = t << 6;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    ts[t] = (uint32_t)lo;
  }
}

// bits of max v_z^2 over the rows (dead rows carry finite velocities; they only loosen the bound)
__global__ void k_vz2max(const double *__restrict__ vz, const int64_t *__restrict__ cnt, unsigned long long *out) {
  const int64_t n = cnt[CNT_NSLOTS];
  double m = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const double v = vz[p];
    m = fmax(m, v * v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// wr[w] = first tile that starts at or after row w * per (per = rows per chunk): every tile belongs to one chunk
__global__ void k_warp_ranges(const uint32_t *__restrict__ ts, uint32_t ntiles, uint32_t nwarps, uint32_t *wr) {
  const uint32_t ns = ts[ntiles];
  const uint32_t per = (ns + nwarps - 1) / nwarps;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w <= nwarps; w += gridDim.x * blockDim.x) {
    uint32_t r = ntiles;
    if (w < nwarps) {
      const uint64_t v = (uint64_t)w * per;
      uint32_t lo = 0, hi = ntiles;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ts[mid] < v) lo = mid + 1; else hi = mid;
      }
      r = lo;
    }
    wr[w] = r;
  }
}

// ---- incremental re-group: destinations from the counts of the previous launch --------------------
// seg = [ rows per destination tile (ntiles) | FAR rows per source tile (ntiles) | rows that stay in the tail (1) |
//         DEAD rows per source tile (ntiles) | dead rows of the tail (1) | 0 ] ; its exclusive scan gives every base of the
// new layout.  tstart[key] = first sorted tail row with that key (k_tail_starts): the tail rows join their tiles.
__global__ void k_regroup_seg(const uint32_t *__restrict__ tcnt, const uint32_t *__restrict__ ts, const int64_t *__restrict__ cnt,
                              uint32_t ntiles, uint32_t mtx, uint32_t tiles_x, uint32_t tiles_y,
                              const uint32_t *__restrict__ tstart, const uint32_t *__restrict__ tn, uint32_t *seg) {
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < ntiles; d += gridDim.x * blockDim.x) {
    int tx, ty;
    tile_coords(d, mtx, tx, ty);
    uint32_t tot = 0;
    if ((uint32_t)tx < tiles_x && (uint32_t)ty < tiles_y) {
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        // source tile = d - (dx, dy) of code c
        const int sx = tx - (c % 3 - 1), sy = ty - (c / 3 - 1);
        if ((uint32_t)sx < tiles_x && (uint32_t)sy < tiles_y) tot += tcnt[tile_ordinal((uint32_t)sx, (uint32_t)sy, mtx) * NCODE + c];
      }
    }
    seg[d] = tot + (tstart[d + 1] - tstart[d]);
    seg[ntiles + d] = tcnt[d * NCODE + CODE_FAR];
    seg[2 * ntiles + 1 + d] = tcnt[d * NCODE + CODE_DEAD];
    if (d == 0) {
      const uint32_t n = (uint32_t)cnt[CNT_NSLOTS], ns = ts[ntiles];
      const uint32_t ntail = n > ns ? n - ns : 0u;
      seg[2 * ntiles] = (ntail - *tn) + (tstart[ntiles + 1] - tstart[ntiles]);   // not merged (beyond the scratch) + outside the grid
      seg[3 * ntiles + 1] = tstart[ntiles + 2] - tstart[ntiles + 1];
      seg[3 * ntiles + 2] = 0;
    }
  }
}

// after the scan: tbase[s][c] for every (source tile, code) and the new tile starts
__global__ void k_regroup_bases(const uint32_t *__restrict__ tcnt, const uint32_t *__restrict__ seg, uint32_t ntiles, uint32_t mtx,
                                uint32_t tiles_x, uint32_t tiles_y, uint32_t *tbase, uint32_t *ts_new, uint32_t *tailbase) {
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d <= ntiles; d += gridDim.x * blockDim.x) {
    ts_new[d] = seg[d];
    if (d == ntiles) {
      tailbase[ntiles] = seg[2 * ntiles];
      tailbase[ntiles + 1] = seg[3 * ntiles + 1];
      break;
    }
    int tx, ty;
    tile_coords(d, mtx, tx, ty);
    uint32_t run = seg[d];
    if ((uint32_t)tx < tiles_x && (uint32_t)ty < tiles_y) {
      // rows that stay come first, then the arrivals in the fixed order of the codes
      tbase[d * NCODE + CODE_STAY] = run;
      run += tcnt[d * NCODE + CODE_STAY];
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        if (c == CODE_STAY) continue;
        const int sx = tx - (c % 3 - 1), sy = ty - (c / 3 - 1);
        if ((uint32_t)sx < tiles_x && (uint32_t)sy < tiles_y) {
          const uint32_t s = tile_ordinal((uint32_t)sx, (uint32_t)sy, mtx);
          tbase[s * NCODE + c] = run;
          run += tcnt[s * NCODE + c];
        }
      }
    }
    tailbase[d] = run;   // the tail rows that are in this tile come last
    tbase[d * NCODE + CODE_FAR] = seg[ntiles + d];
    tbase[d * NCODE + CODE_DEAD] = seg[2 * ntiles + 1 + d];
  }
}

// ---- tail merge: the unsorted tail [ns, n) -- rows born since the last re-group (sources, ionisation) and rows that jumped
// further than a neighbouring tile -- joins the tile segments on every MOVE.  Keys are taken from the CURRENT position (the
// MOVE places every row by where it is before the push, like the codes of the tile rows), sorted stably (sort.cu), so the
// new order is a pure function of the old one: no full sort is needed in steady state.
__global__ void k_tail_keys(const double *__restrict__ x, const double *__restrict__ y, const int64_t *__restrict__ cnt,
                            const uint32_t *__restrict__ ts, GridDev g, uint32_t ntiles, uint32_t mtx, uint32_t tcap, uint32_t *keys,
                            uint32_t *tn) {
  const uint32_t n = (uint32_t)cnt[CNT_NSLOTS], ns = ts[ntiles];
  const uint32_t ntail = n > ns ? n - ns : 0u, T = min(ntail, tcap);
  if (blockIdx.x == 0 && threadIdx.x == 0) *tn = T;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < T; k += gridDim.x * blockDim.x) {
    const double px = x[ns + k];
    uint32_t key = ntiles + 1;   // dead: parked
    if (!is_dead(px)) {
      int i, j;
      double hx, hy;
      cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
      cell1(y[ns + k], g.dy, g.rdy, g.fast_div, j, hy);
      key = cell_in_grid(i, j, g.nx, g.ny) ? tile_ordinal((uint32_t)(i - 1) >> 3, (uint32_t)(j - 1) >> 3, mtx) : ntiles;
    }
    keys[k] = key;
  }
}

// tstart[v] = first sorted tail row whose key is >= v, v = 0 .. ntiles + 2
__global__ void k_tail_starts(const uint32_t *__restrict__ sk, const uint32_t *__restrict__ tn, uint32_t ntiles, uint32_t *tstart) {
  const uint32_t T = *tn;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v <= ntiles + 2; v += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = T;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sk[mid] < v) lo = mid + 1; else hi = mid;
    }
    tstart[v] = lo;
  }
}

__global__ void k_tail_dest(const uint32_t *__restrict__ sk, const uint32_t *__restrict__ si, const uint32_t *__restrict__ tn,
                            const uint32_t *__restrict__ tstart, const uint32_t *__restrict__ tailbase, uint32_t *taildst,
                            uint32_t *tailkey) {
  const uint32_t T = *tn;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < T; j += gridDim.x * blockDim.x) {
    const uint32_t key = sk[j], k = si[j];
    taildst[k] = tailbase[key] + (j - tstart[key]);
    tailkey[k] = key;
  }
}

// counters after a MOVE: live slots end where the parked (discarded) rows begin; those rows leave the dead count
__global__ void k_after_move(int64_t *cnt, const uint32_t *__restrict__ seg, uint32_t ntiles) {
  const int64_t n_new = seg[2 * ntiles + 1], parked = (int64_t)seg[3 * ntiles + 2] - n_new;
  cnt[CNT_NSLOTS] = n_new;
  cnt[CNT_NDEAD] -= parked;
}

// rows beyond the slot count (parked ids of removed rows, default weights) must survive the buffer swap of a MOVE
__global__ void k_copy_parked(const uint32_t *__restrict__ id, uint32_t *oid, const double *__restrict__ wg, double *owg,
                              const int64_t *__restrict__ cnt, int64_t cap) {
  for (int64_t p = cnt[CNT_NSLOTS] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < cap; p += (int64_t)gridDim.x * blockDim.x) {
    oid[p] = id[p];
    owg[p] = wg[p];
  }
}

// MARK state right after a full sort: every sorted row sits in its own tile
__global__ void k_marks_after_sort(const uint32_t *__restrict__ ts, uint32_t ntiles, uint32_t *tcnt, uint8_t *code, int64_t n) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = gid; t < (int64_t)ntiles * NCODE; t += gsz) {
    const uint32_t tile = (uint32_t)(t / NCODE), c = (uint32_t)(t % NCODE);
    tcnt[t] = c == (uint32_t)CODE_STAY ? ts[tile + 1] - ts[tile] : 0u;
  }
  for (int64_t p = gid; p < n; p += gsz) code[p] = (uint8_t)CODE_STAY;
}

}  // namespace

int32_t exclusive_scan_u32(iskb_ctx *c, uint32_t *d, int64_t n, uint32_t *partial);
int32_t sort_pairs_device_count(iskb_species *sp, int64_t ncap, const uint32_t *n_dev, int bits, uint32_t **keys_out,
                                uint32_t **idx_out, uint32_t **spare_key, uint32_t **spare_idx);
int32_t sort_scratch_ensure(iskb_species *sp);
int32_t launch_advance_simple(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit, bool from_begin);

// persistent CTAs of 8 warps, three per SM.  Four (64 registers, 7 KB per warp) measured +0.16 ms per step at the C5 shard:
// the kernel is bound by L1/shared wavefronts, not by latency, and the fourth CTA takes the L1 that is left.
static int advance_grid(const iskb_ctx *c) { return c->n_sm * 3; }
static int advance_chunks(const iskb_ctx *c) { return c->n_sm * 3 * 8 * 8; }   // ~8 chunks per resident warp

int32_t tdir_ensure(iskb_species *sp) {
  iskb_ctx *c = sp->ctx;
  if (sp->d_ts[0]) return ISKB_OK;
  const TileGeom tg = tile_geom(c->g);
  const size_t nt = tg.ntiles;
  for (int k = 0; k < 2; ++k) CU_TRY(cudaMalloc(&sp->d_ts[k], (nt + 1) * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_tcnt, nt * NCODE * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_tbase, nt * NCODE * sizeof(uint32_t)));
  const size_t nseg = 3 * nt + 3;
  CU_TRY(cudaMalloc(&sp->d_seg, (nseg + (nseg + 2047) / 2048 + 16) * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_tstart, (nt + 3) * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_tailbase, (nt + 2) * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_tn, sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_wr, ((size_t)advance_chunks(c) + 1) * sizeof(uint32_t)));
  CU_TRY(cudaMalloc(&sp->d_ticket, sizeof(unsigned)));
  CU_TRY(cudaMalloc(&sp->d_ufix, (size_t)c->g.nx * c->g.ny * sizeof(long long)));
  CU_TRY(cudaMalloc(&sp->d_code, sp->cap));
  CU_TRY(cudaMalloc(&sp->alt_code, sp->cap));
  CU_TRY(cudaMalloc(&sp->d_mlist, sp->cap * sizeof(uint2)));
  CU_TRY(cudaMalloc(&sp->d_mlist_n, sizeof(unsigned)));
  return sp_ensure_alt(sp);
}

void sp_touch(iskb_species *sp) {
  // an MCC test phase that ran ahead on this species (mcc_launch phase 1) is void now
  for (iskb_mcc *m : sp->ctx->mccs)
    if (m->source == sp) mcc_discard_pre(m);
  sp->epoch++;
  sp->tdir_valid = false;
  sp->marks_valid = false;
  sp->vz2_known = false;
}

void tdir_free(iskb_species *sp) {
  if (sp->h_tstats) cudaFreeHost(sp->h_tstats);
  for (int k = 0; k < 2; ++k) if (sp->ev_tstats[k]) cudaEventDestroy(sp->ev_tstats[k]);
  for (int k = 0; k < 2; ++k) cudaFree(sp->d_ts[k]);
  cudaFree(sp->d_tcnt); cudaFree(sp->d_tbase); cudaFree(sp->d_seg); cudaFree(sp->d_wr);
  cudaFree(sp->d_code); cudaFree(sp->alt_code); cudaFree(sp->d_mlist); cudaFree(sp->d_mlist_n); cudaFree(sp->d_ticket); cudaFree(sp->d_ufix);
  cudaFree(sp->d_tstart); cudaFree(sp->d_tailbase); cudaFree(sp->d_tn);
}

static int32_t warp_ranges(iskb_species *sp) {
  iskb_ctx *c = sp->ctx;
  const TileGeom tg = tile_geom(c->g);
  const uint32_t nw = (uint32_t)advance_chunks(c);
  k_warp_ranges<<<(nw + 256) / 256, 256, 0, c->stream>>>(sp->d_ts[0], tg.ntiles, nw, sp->d_wr);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

// called by the full sort (sort.cu) with the sorted cell keys still in place
int32_t tdir_build(iskb_species *sp, const uint32_t *sorted_keys, int64_t n) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(tdir_ensure(sp));
  const TileGeom tg = tile_geom(c->g);
  k_tile_starts<<<(tg.ntiles + 256) / 256, 256, 0, c->stream>>>(sorted_keys, n, tg.ntiles, sp->d_ts[0]);
  LAUNCH_CHECK(c);
  int blocks = (int)((n + 255) / 256);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  if (blocks < 1) blocks = 1;
  k_marks_after_sort<<<blocks, 256, 0, c->stream>>>(sp->d_ts[0], tg.ntiles, sp->d_tcnt, sp->d_code, n);
  LAUNCH_CHECK(c);
  ISKB_TRY(warp_ranges(sp));
  sp->tdir_valid = true;
  sp->marks_valid = true;
  sp->steps_since_move = 0;
  return ISKB_OK;
}

template <int MX, int MY, bool RZ, bool LEAN>
static int32_t launch_tile_modes(iskb_species *sp, double dt, bool move, bool mark) {
  iskb_ctx *c = sp->ctx;
  const TileGeom tg = tile_geom(c->g);
  TileArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < 6; ++q) { a.col[q] = sp->col[q]; a.ocol[q] = sp->alt[q]; }
  a.id = sp->id; a.oid = sp->alt_id;
  a.code = sp->d_code; a.ocode = sp->alt_code;
  a.cnt = sp->d_cnt;
  a.g = c->g;
  a.E2 = c->d_E2;
  a.qm = sp->q / sp->m;
  a.hqm = 0.5 * a.qm;
  a.dt = dt;
  a.c1 = 0.5 * dt * a.qm;
  a.w0 = sp->w0;
  a.ufix = sp->d_ufix;
  a.fscale = c->fscale;
  a.status = c->d_status;
  a.vmax2 = sp->d_vmax2;
  a.ts = sp->d_ts[0];
  a.wr = sp->d_wr;
  a.ntiles = tg.ntiles; a.mtx = tg.mtx; a.tiles_x = tg.tiles_x; a.tiles_y = tg.tiles_y;
  a.tcnt = sp->d_tcnt;
  a.tbase = sp->d_tbase;
  a.seg = sp->d_seg;
  a.mlist = sp->d_mlist;
  a.mlist_n = sp->d_mlist_n;
  a.mlist_cap = (unsigned)sp->cap;
  a.vz2max = sp->d_vz2max;
  a.ticket = sp->d_ticket;
  a.nchunks = (unsigned)advance_chunks(c);
  a.mark = mark ? 1 : 0;
  if (LEAN && !sp->vz2_known) {   // bound of v_z^2 (MCC pruning): the lean kernels do not see the column
    CU_TRY(cudaMemsetAsync(sp->d_vz2max, 0, sizeof(unsigned long long), c->stream));
    k_vz2max<<<c->n_sm * 4, 256, 0, c->stream>>>(sp->col[4], sp->d_cnt, sp->d_vz2max);
    LAUNCH_CHECK(c);
    sp->vz2_known = true;
  }
  const int grid = advance_grid(c);
  const int SMEM = 8 * (int)(move ? sizeof(WarpSm<true>) : sizeof(WarpSm<false>));
  if (move) {
    // destinations of this launch from the counts of the previous one
    const int nb = (int)((tg.ntiles + 255) / 256);
    const int64_t nseg = 3 * (int64_t)tg.ntiles + 3;
    // tail merge: keys of the tail rows, stable sort, rows per tile
    ISKB_TRY(sort_scratch_ensure(sp));
    // grid of the merge: the tail of the latest snapshot (two steps old) with room to grow; rows beyond it stay in the
    // tail for the next MOVE (k_tail_keys clamps), a tail beyond 5 % of the rows gets a full sort instead (api.cu)
    const int64_t tcap = std::min<int64_t>(sp->cap, 4 * std::max<int64_t>(sp->tail_rows, 0) + 262144);
    int bits = 1;
    while ((1ull << bits) <= (uint64_t)tg.ntiles + 1u) ++bits;
    const int tb = (int)std::min<int64_t>((tcap + 255) / 256, (int64_t)c->n_sm * 8);
    k_tail_keys<<<tb, 256, 0, c->stream>>>(sp->col[0], sp->col[1], sp->d_cnt, sp->d_ts[0], c->g, tg.ntiles, tg.mtx, (uint32_t)tcap,
                                           sp->d_key[0], sp->d_tn);
    LAUNCH_CHECK(c);
    uint32_t *sk, *si, *spare_key, *spare_idx;
    ISKB_TRY(sort_pairs_device_count(sp, tcap, sp->d_tn, bits, &sk, &si, &spare_key, &spare_idx));
    k_tail_starts<<<nb + 1, 256, 0, c->stream>>>(sk, sp->d_tn, tg.ntiles, sp->d_tstart);
    LAUNCH_CHECK(c);
    k_regroup_seg<<<nb, 256, 0, c->stream>>>(sp->d_tcnt, sp->d_ts[0], sp->d_cnt, tg.ntiles, tg.mtx, tg.tiles_x, tg.tiles_y,
                                             sp->d_tstart, sp->d_tn, sp->d_seg);
    LAUNCH_CHECK(c);
    ISKB_TRY(exclusive_scan_u32(c, sp->d_seg, nseg, sp->d_seg + nseg));
    k_regroup_bases<<<nb + 1, 256, 0, c->stream>>>(sp->d_tcnt, sp->d_seg, tg.ntiles, tg.mtx, tg.tiles_x, tg.tiles_y, sp->d_tbase,
                                                  sp->d_ts[1], sp->d_tailbase);
    LAUNCH_CHECK(c);
    k_tail_dest<<<tb, 256, 0, c->stream>>>(sk, si, sp->d_tn, sp->d_tstart, sp->d_tailbase, spare_idx, spare_key);
    LAUNCH_CHECK(c);
    a.taildst = spare_idx; a.tailkey = spare_key; a.tstart = sp->d_tstart; a.tailbase = sp->d_tailbase; a.tn = sp->d_tn;
  }
  if (a.mark) CU_TRY(cudaMemsetAsync(sp->d_tcnt, 0, (size_t)tg.ntiles * NCODE * sizeof(uint32_t), c->stream));
  CU_TRY(cudaMemsetAsync(sp->d_mlist_n, 0, sizeof(unsigned), c->stream));
  CU_TRY(cudaMemsetAsync(sp->d_ticket, 0, sizeof(unsigned), c->stream));
  ISKB_TRY(sp_vmax_reset(sp));
  ISKB_TRY(prof_begin(c));
  if (move) {
    CU_TRY(cudaFuncSetAttribute(k_advance_tile<MX, MY, true, RZ, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    k_advance_tile<MX, MY, true, RZ, LEAN><<<grid, 256, SMEM, c->stream>>>(a);
    LAUNCH_CHECK(c);
    k_advance_list<MX, MY, true, RZ, LEAN><<<c->n_sm * 4, 256, 0, c->stream>>>(a);
    LAUNCH_CHECK(c);
  } else {
    CU_TRY(cudaFuncSetAttribute(k_advance_tile<MX, MY, false, RZ, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    k_advance_tile<MX, MY, false, RZ, LEAN><<<grid, 256, SMEM, c->stream>>>(a);
    LAUNCH_CHECK(c);
    k_advance_list<MX, MY, false, RZ, LEAN><<<c->n_sm * 4, 256, 0, c->stream>>>(a);
    LAUNCH_CHECK(c);
  }
  ISKB_TRY(prof_end(c));
  if (move) {
    // rows beyond the old slot count (parked ids, default weights) must survive the buffer swap
    k_copy_parked<<<c->n_sm * 2, 256, 0, c->stream>>>(sp->id, sp->alt_id, sp->col[5], sp->alt[5], sp->d_cnt, sp->cap);
    LAUNCH_CHECK(c);
    k_after_move<<<1, 1, 0, c->stream>>>(sp->d_cnt, sp->d_seg, tg.ntiles);
    LAUNCH_CHECK(c);
    for (int q = 0; q < 6; ++q) std::swap(sp->col[q], sp->alt[q]);
    std::swap(sp->id, sp->alt_id);
    std::swap(sp->d_code, sp->alt_code);
    std::swap(sp->d_ts[0], sp->d_ts[1]);
    ISKB_TRY(warp_ranges(sp));
    sp->steps_since_move = 0;
  }
  sp->counts_stale = true;
  sp->marks_valid = a.mark != 0;   // an unmarked launch leaves codes and counts of older positions behind
  return ISKB_OK;
}

template <bool RZ, bool LEAN>
static int32_t launch_tile_rz(iskb_species *sp, double dt, int mx, int my, bool move, bool mark) {
  switch (mx * 3 + my) {
    case 0: return launch_tile_modes<0, 0, RZ, LEAN>(sp, dt, move, mark);
    case 1: return launch_tile_modes<0, 1, RZ, LEAN>(sp, dt, move, mark);
    case 2: return launch_tile_modes<0, 2, RZ, LEAN>(sp, dt, move, mark);
    case 3: return launch_tile_modes<1, 0, RZ, LEAN>(sp, dt, move, mark);
    case 4: return launch_tile_modes<1, 1, RZ, LEAN>(sp, dt, move, mark);
    case 5: return launch_tile_modes<1, 2, RZ, LEAN>(sp, dt, move, mark);
    case 6: return launch_tile_modes<2, 0, RZ, LEAN>(sp, dt, move, mark);
    case 7: return launch_tile_modes<2, 1, RZ, LEAN>(sp, dt, move, mark);
    default: return launch_tile_modes<2, 2, RZ, LEAN>(sp, dt, move, mark);
  }
}

// advance! + density of one species on the tile directory; `move` re-groups the rows on the way out
int32_t launch_advance_tile(iskb_species *sp, double dt, int mode_x, int mode_y, bool move, bool mark) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(fields_join(c));
  if (!sp->tdir_valid) return iskb_fail(ISKB_E_INVALID, "tile directory not built (internal)");
  if (move && !sp->marks_valid) return iskb_fail(ISKB_E_INVALID, "re-group without valid marks (internal)");
  if (c->pusher_rz) return launch_tile_rz<true, false>(sp, dt, mode_x, mode_y, move, mark);   // the r-z transform rotates (v_x, v_z)
  if (sp->wg_uniform && c->lean_ok) return launch_tile_rz<false, true>(sp, dt, mode_x, mode_y, move, mark);
  return launch_tile_rz<false, false>(sp, dt, mode_x, mode_y, move, mark);
}
