// Stable LSD radix sort of the particle rows by cell key, and stable compaction of discarded
// rows (the reference's remove!, kinetic.jl:20-27, keeps `id` a permutation of 1..N; here the
// discarded rows are parked behind the live ones with their ids, which preserves the same
// invariant -- SURVEY.md H5).  New relative to the reference: sorting exists only to make the
// tiled shared-memory deposition possible.
//
// Sort key (DESIGN.md "cell key"): cells are grouped in 8x8 tiles, tiles in 16x16 meta-tiles:
//   tile = (((ty>>4)*mtx + (tx>>4)) << 8) | ((ty&15) << 4) | (tx&15),   mtx = ceil(tiles_x/16)
//   key  = (tile << 6) | ((cy & 7) << 3) | (cx & 7)
// with cx = i-1, cy = j-1 from particle_cell (ParticleInCell.jl:28-35), tx = cx>>3, ty = cy>>3.
// The meta-tile level keeps the rows of vertically adjacent tiles within ~16 tiles of each other in
// memory, so the gather of a re-sort finds the rows that crossed a tile edge in L2.  Out-of-grid rows get key_max and
// dead rows key_max+1, so they land at the end.  The sort is stable in the previous row order, so the resulting
// permutation is a pure function of the cell indices (bit-exact contract of north_star).
#include "tiles.cuh"

int32_t tdir_build(iskb_species *sp, const uint32_t *sorted_keys, int64_t n);

namespace {

constexpr int TPB = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_WARPS = TPB / 32;
constexpr int RS_TILE = TPB * RS_ITEMS;   // 4096 keys per block

__global__ void k_cell_keys(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                            GridDev g, uint32_t mtx, uint32_t key_max, uint32_t *keys) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double px = x[p];
    uint32_t key = key_max + 1u;   // dead rows go last, behind live out-of-grid rows (key_max)
    if (!is_dead(px)) {
      key = key_max;
      int i, j;
      double hx, hy;
      cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
      cell1(y[p], g.dy, g.rdy, g.fast_div, j, hy);
      if (cell_in_grid(i, j, g.nx, g.ny)) {
        const uint32_t cx = (uint32_t)(i - 1), cy = (uint32_t)(j - 1);
        key = (tile_ordinal(cx >> 3, cy >> 3, mtx) << 6) | ((cy & 7u) << 3) | (cx & 7u);
      }
    }
    keys[p] = key;
  }
}

// Tile-only key (re-group between full sorts): rows keep their previous relative order inside the
// tile (stable sort), which was the cell-interleaved order of the last full sort plus drift.
__global__ void k_tile_keys(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                            GridDev g, uint32_t mtx, uint32_t ntiles, uint32_t *keys) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double px = x[p];
    uint32_t key = ntiles + 1u;
    if (!is_dead(px)) {
      key = ntiles;
      int i, j;
      double hx, hy;
      cell1(px, g.dx, g.rdx, g.fast_div, i, hx);
      cell1(y[p], g.dy, g.rdy, g.fast_div, j, hy);
      if (cell_in_grid(i, j, g.nx, g.ny)) key = tile_ordinal((uint32_t)(i - 1) >> 3, (uint32_t)(j - 1) >> 3, mtx);
    }
    keys[p] = key;
  }
}

__global__ void k_dead_keys(const double *__restrict__ x, int64_t n, uint32_t *keys) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x)
    keys[p] = is_dead(x[p]) ? 1u : 0u;
}

template <int BITS>
__global__ void k_radix_hist(const uint32_t *__restrict__ keys, int64_t n, int shift, uint32_t *hist,
                             int nblocks, const uint32_t *__restrict__ n_dev = nullptr) {
  constexpr int BINS = 1 << BITS;
  constexpr uint32_t MASK = BINS - 1;
  __shared__ uint32_t h[BINS];
  if (n_dev) n = min(n, (int64_t)*n_dev);   // device-side count (tail merge of a MOVE, advance_tile.cu)
  for (int d = threadIdx.x; d < BINS; d += TPB) h[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int t = 0; t < RS_ITEMS; ++t) {
    const int64_t p = base + t * TPB + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & MASK], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < BINS; d += TPB) hist[(int64_t)d * nblocks + blockIdx.x] = h[d];
}

// ---- exclusive scan over uint32 (3 kernels) ----------------------------------------------------
constexpr int SC_PER_BLOCK = 2048;   // 256 threads x 8

__global__ void k_scan_reduce(const uint32_t *__restrict__ in, int64_t n, uint32_t *partial) {
  __shared__ uint32_t s[TPB];
  const int64_t base = (int64_t)blockIdx.x * SC_PER_BLOCK + threadIdx.x * 8;
  uint32_t a = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) if (base + k < n) a += in[base + k];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = TPB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void k_scan_partials(uint32_t *partial, int np) {   // single block, exclusive in place
  __shared__ uint32_t s[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < np; base += 1024) {
    const int k = base + threadIdx.x;
    const uint32_t v = k < np ? partial[k] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (k < np) partial[k] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
}

__global__ void k_scan_down(uint32_t *data, int64_t n, const uint32_t *__restrict__ partial) {
  __shared__ uint32_t s[TPB];
  const int64_t base = (int64_t)blockIdx.x * SC_PER_BLOCK + threadIdx.x * 8;
  uint32_t v[8];
  uint32_t a = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = base + k < n ? data[base + k] : 0;
    a += v[k];
  }
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 1; o < TPB; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = partial[blockIdx.x] + s[threadIdx.x] - a;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}

// ---- stable scatter of one BITS-wide digit -----------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(TPB, 4) k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ idx_in,
                                uint32_t *keys_out, uint32_t *idx_out, int64_t n, int shift,
                                const uint32_t *__restrict__ offs, int nblocks, const uint32_t *__restrict__ n_dev = nullptr) {
  constexpr int BINS = 1 << BITS;
  constexpr uint32_t MASK = BINS - 1;
  __shared__ uint32_t wcnt[RS_WARPS][BINS];
  if (n_dev) n = min(n, (int64_t)*n_dev);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  for (int k = threadIdx.x; k < RS_WARPS * BINS; k += TPB) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ITEMS * 32);
  uint32_t key[RS_ITEMS], src[RS_ITEMS];
  // phase 1: per-warp digit counts (warp-private rows; one writer per digit and iteration)
#pragma unroll
  for (int t = 0; t < RS_ITEMS; ++t) {
    const int64_t p = wbase + t * 32 + lane;
    const bool valid = p < n;
    key[t] = valid ? keys_in[p] : 0xffffffffu;
    src[t] = (valid && idx_in) ? idx_in[p] : (uint32_t)p;
    const uint32_t d = valid ? ((key[t] >> shift) & MASK) : (uint32_t)BINS;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt) == 0) wcnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix across warps + global base of (digit, block)
  for (int d = threadIdx.x; d < BINS; d += TPB) {
    uint32_t run = offs[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = wcnt[w][d];
      wcnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase 2: stable ranks and scatter
#pragma unroll
  for (int t = 0; t < RS_ITEMS; ++t) {
    const int64_t p = wbase + t * 32 + lane;
    const bool valid = p < n;
    const uint32_t d = valid ? ((key[t] >> shift) & MASK) : (uint32_t)BINS;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = wcnt[warp][d];
      wcnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    if (valid) {
      const uint32_t dest = old + __popc(peers & lt);
      keys_out[dest] = key[t];
      idx_out[dest] = src[t];
    }
    __syncwarp();
  }
}

// ---- cell-interleaved order inside every 8x8 tile ----------------------------------------------
// After the stable sort the rows of a tile are grouped by cell.  The fused advance kernel deposits
// with shared-memory CAS adds whose cost is proportional to how many lanes of a 32-row batch hit
// the same node, so the rows of a tile are re-ordered round-robin over its cells: first the
// rank-0 row of every non-empty cell (cell order), then the rank-1 rows, ...  A batch of 32
// consecutive rows then touches 32 different cells.  The round-robin visits the cells in CHECKERBOARD
// order pi(c) = (((cx+cy)&1) << 5) | (c >> 1): first the 32 "even" cells, then the 32 "odd" ones, so the
// cells of one batch are never edge neighbours and stay distinct longer while the rows drift.
// dest(c, r) = start + F[r] + #{c' : pi(c') < pi(c), cnt[c'] > r}, F[r] = sum_c' min(cnt[c'], r);
// ranks >= IL_RMAX keep the sorted order behind the interleaved part.
// Deterministic: a pure function of the sorted (cell, previous-row) order.
constexpr int IL_RMAX = 512;
__device__ __forceinline__ uint32_t il_pi(uint32_t c) { return ((((c & 7u) + (c >> 3)) & 1u) << 5) | (c >> 1); }
__device__ __forceinline__ uint32_t il_pi_inv(uint32_t t) {   // cell whose checkerboard position is t
  const uint32_t par = t >> 5, r = t & 31u, cy = r >> 2, cxh = r & 3u;
  return (cy << 3) | (2u * cxh + ((cy & 1u) ^ par));
}

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t *__restrict__ a, int64_t lo, int64_t hi, uint32_t v) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(64) k_tile_interleave(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ idx_in,
                                                        uint32_t *__restrict__ idx_out, int64_t n, uint32_t ntiles) {
  __shared__ int64_t s_cs[65];            // start of every cell's run
  __shared__ uint32_t s_cnt[64], s_over[65];
  __shared__ unsigned long long s_mask[IL_RMAX];
  __shared__ uint32_t s_F[IL_RMAX + 1];
  __shared__ uint32_t s_maxcnt;
  const uint32_t tile = blockIdx.x;
  const int c = threadIdx.x;
  if (tile == ntiles) {   // out-of-grid and dead rows: keep the sorted order
    const int64_t s0 = lower_bound_u32(keys, 0, n, ntiles << 6);
    for (int64_t p = s0 + c; p < n; p += 64) idx_out[p] = idx_in[p];
    return;
  }
  s_cs[c] = lower_bound_u32(keys, 0, n, (tile << 6) | (uint32_t)c);
  if (c == 0) s_cs[64] = lower_bound_u32(keys, 0, n, (tile + 1) << 6);
  __syncthreads();
  const int64_t start = s_cs[0], end = s_cs[64];
  if (end == start) return;
  const uint32_t cnt = (uint32_t)(s_cs[c + 1] - s_cs[c]);
  s_cnt[c] = cnt;
  if (c == 0) s_maxcnt = 0;
  __syncthreads();
  atomicMax(&s_maxcnt, cnt);
  __syncthreads();
  const uint32_t R = s_maxcnt < IL_RMAX ? s_maxcnt : IL_RMAX;
  const uint32_t cnt_pi = s_cnt[il_pi_inv((uint32_t)c)];   // thread c owns checkerboard position c
  for (uint32_t r = 0; r < R; ++r) {      // mask_r (bit = checkerboard position): cells that still have a row of rank r
    const unsigned b = __ballot_sync(0xffffffffu, cnt_pi > r);
    if ((c & 31) == 0) ((unsigned *)&s_mask[r])[c >> 5] = b;
  }
  __syncthreads();
  if (c == 0) {
    uint32_t f = 0, o = 0;
    for (uint32_t r = 0; r < R; ++r) { s_F[r] = f; f += __popcll(s_mask[r]); }
    s_F[R] = f;
    for (int q = 0; q < 64; ++q) { s_over[q] = o; o += s_cnt[q] > R ? s_cnt[q] - R : 0; }
  }
  __syncthreads();
  for (int64_t p = start + c; p < end; p += 64) {
    const uint32_t cell = keys[p] & 63u;
    const uint32_t r = (uint32_t)(p - s_cs[cell]);
    int64_t dest;
    if (r < R) dest = start + s_F[r] + __popcll(s_mask[r] & ((1ull << il_pi(cell)) - 1ull));
    else dest = start + s_F[R] + s_over[cell] + (r - R);
    idx_out[dest] = idx_in[p];
  }
}

struct Cols {
  const double *in[6];
  double *out[6];
  const uint32_t *id_in;
  uint32_t *id_out;
  double w0;   // weight a vacated slot is reset to (remove!, kinetic.jl:24)
};

// Each block moves a CONTIGUOUS range of destination rows: after a re-sort the sources of
// neighbouring destination rows sit in the same few tiles, so the 32-byte sectors fetched for one
// row are reused from L1 by the other rows of the block instead of being re-read through L2.
__global__ void __launch_bounds__(256) k_permute(Cols c, const uint32_t *__restrict__ idx, int64_t n, int chunk,
                                                 uint32_t *ticket) {
  // chunks are handed out in increasing order (ticket), so the rows in flight always form one
  // compact range and the 64-byte lines shared between neighbouring chunks are found in L2
  __shared__ uint32_t s_chunk;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t base = (int64_t)s_chunk * chunk;
    if (base >= n) break;
    const int64_t end = base + chunk < n ? base + chunk : n;
    for (int64_t k = base + threadIdx.x; k < end; k += blockDim.x) {
      const uint32_t s = idx[k];
      double v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = c.in[q][s];
      const uint32_t id = c.id_in[s];
      if (is_dead(v[0])) v[5] = c.w0;   // a parked row is a free slot: whatever is created there later starts with the default weight
#pragma unroll
      for (int q = 0; q < 6; ++q) c.out[q][k] = v[q];
      c.id_out[k] = id;
    }
  }
}

__global__ void k_counts_after_sort(int64_t *cnt) {
  cnt[CNT_NSLOTS] -= cnt[CNT_NDEAD];
  cnt[CNT_NDEAD] = 0;
}

}  // namespace

int32_t exclusive_scan_u32(iskb_ctx *c, uint32_t *d, int64_t n, uint32_t *partial) {
  const int nb = (int)((n + SC_PER_BLOCK - 1) / SC_PER_BLOCK);
  k_scan_reduce<<<nb, TPB, 0, c->stream>>>(d, n, partial);
  LAUNCH_CHECK(c);
  k_scan_partials<<<1, 1024, 0, c->stream>>>(partial, nb);
  LAUNCH_CHECK(c);
  k_scan_down<<<nb, TPB, 0, c->stream>>>(d, n, partial);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}

namespace {

int32_t ensure_sort_scratch(iskb_species *sp) {
  if (!sp->d_key[0]) {
    for (int k = 0; k < 2; ++k) {
      CU_TRY(cudaMalloc(&sp->d_key[k], sp->cap * sizeof(uint32_t)));
      CU_TRY(cudaMalloc(&sp->d_idx[k], sp->cap * sizeof(uint32_t)));
    }
    const int64_t nblocks = (sp->cap + RS_TILE - 1) / RS_TILE;
    const int64_t hn = 512 * nblocks;
    sp->hist_cap = hn + (hn + SC_PER_BLOCK - 1) / SC_PER_BLOCK + 16;
    CU_TRY(cudaMalloc(&sp->d_hist, sp->hist_cap * sizeof(uint32_t)));
  }
  return sp_ensure_alt(sp);
}

// keys (of `bits` significant bits) already in d_key[0][0..n); runs the digit passes, permutes all
// columns, fixes counters.  Digits are 8 bits wide, or 9 when that saves a whole pass (17-bit tile keys).
int32_t sort_by_keys(iskb_species *sp, int64_t n, int bits, uint32_t *perm_out_host, uint32_t interleave_tiles = 0) {
  iskb_ctx *c = sp->ctx;
  sp_touch(sp);   // rows are about to move: whatever ran ahead on the old layout is void (and is waited for)
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  const int width = (bits + 8) / 9 < (bits + 7) / 8 ? 9 : 8;
  const int passes = (bits + width - 1) / width;
  const int64_t hn = ((int64_t)1 << width) * nblocks;
  int cur = 0;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = width * pass;
    const uint32_t *idx_in = pass == 0 ? nullptr : sp->d_idx[cur];
    if (width == 9) k_radix_hist<9><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], n, shift, sp->d_hist, nblocks);
    else k_radix_hist<8><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], n, shift, sp->d_hist, nblocks);
    LAUNCH_CHECK(c);
    ISKB_TRY(exclusive_scan_u32(c, sp->d_hist, hn, sp->d_hist + hn));
    if (width == 9)
      k_radix_scatter<9><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], idx_in, sp->d_key[cur ^ 1], sp->d_idx[cur ^ 1],
                                                         n, shift, sp->d_hist, nblocks);
    else
      k_radix_scatter<8><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], idx_in, sp->d_key[cur ^ 1], sp->d_idx[cur ^ 1],
                                                         n, shift, sp->d_hist, nblocks);
    LAUNCH_CHECK(c);
    cur ^= 1;
  }
  if (interleave_tiles) {
    k_tile_interleave<<<interleave_tiles + 1, 64, 0, c->stream>>>(sp->d_key[cur], sp->d_idx[cur], sp->d_idx[cur ^ 1], n,
                                                                interleave_tiles);
    LAUNCH_CHECK(c);
    std::swap(sp->d_idx[cur], sp->d_idx[cur ^ 1]);   // keys stay in d_key[cur]; idx now interleaved
  }
  Cols cols;
  for (int q = 0; q < 6; ++q) {
    cols.in[q] = sp->col[q];
    cols.out[q] = sp->alt[q];
  }
  cols.id_in = sp->id;
  cols.id_out = sp->alt_id;
  cols.w0 = sp->w0;
  constexpr int PERM_CHUNK = 512;
  int blocks = (int)((n + PERM_CHUNK - 1) / PERM_CHUNK);
  if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
  uint32_t *ticket = sp->d_hist + sp->hist_cap - 1;   // last scratch word (hist_cap has 16 words of slack)
  CU_TRY(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), c->stream));
  k_permute<<<blocks, TPB, 0, c->stream>>>(cols, sp->d_idx[cur], n, PERM_CHUNK, ticket);
  LAUNCH_CHECK(c);
  // rows >= n (parked ids / default weights) must survive the buffer swap
  if (sp->cap > n) {
    CU_TRY(cudaMemcpyAsync(sp->alt_id + n, sp->id + n, (sp->cap - n) * sizeof(uint32_t),
                           cudaMemcpyDeviceToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(sp->alt[5] + n, sp->col[5] + n, (sp->cap - n) * sizeof(double),
                           cudaMemcpyDeviceToDevice, c->stream));
  }
  for (int q = 0; q < 6; ++q) std::swap(sp->col[q], sp->alt[q]);
  std::swap(sp->id, sp->alt_id);
  k_counts_after_sort<<<1, 1, 0, c->stream>>>(sp->d_cnt);
  LAUNCH_CHECK(c);
  // the tile directory of the tile-aware advance (advance_tile.cu) describes the layout of a FULL sort only
  sp->tdir_valid = false;
  sp->marks_valid = false;
  if (interleave_tiles) ISKB_TRY(tdir_build(sp, sp->d_key[cur], n));
  const int64_t nlive = n - sp->h_ndead;
  if (perm_out_host && nlive > 0)
    CU_TRY(cudaMemcpyAsync(perm_out_host, sp->d_idx[cur], nlive * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                           c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  sp->h_nslots = nlive;
  sp->h_nsorted = nlive;
  sp->h_ndead = 0;
  sp->counts_stale = false;
  return ISKB_OK;
}

}  // namespace

// Stable sort of (key, position) pairs whose count lives on the device: keys in d_key[0][0 .. min(*n_dev, ncap)).  The grid
// covers ncap keys.  On return *keys_out / *idx_out are the sorted keys and their source positions, *spare_key / *spare_idx
// the other half of the scratch (free for the caller).  Used by the tail merge of a MOVE (advance_tile.cu).
int32_t sort_pairs_device_count(iskb_species *sp, int64_t ncap, const uint32_t *n_dev, int bits, uint32_t **keys_out,
                                uint32_t **idx_out, uint32_t **spare_key, uint32_t **spare_idx) {
  iskb_ctx *c = sp->ctx;
  const int nblocks = (int)((ncap + RS_TILE - 1) / RS_TILE);
  const int width = (bits + 8) / 9 < (bits + 7) / 8 ? 9 : 8;
  const int passes = (bits + width - 1) / width;
  const int64_t hn = ((int64_t)1 << width) * nblocks;
  int cur = 0;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = width * pass;
    const uint32_t *idx_in = pass == 0 ? nullptr : sp->d_idx[cur];
    if (width == 9) k_radix_hist<9><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], ncap, shift, sp->d_hist, nblocks, n_dev);
    else k_radix_hist<8><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], ncap, shift, sp->d_hist, nblocks, n_dev);
    LAUNCH_CHECK(c);
    ISKB_TRY(exclusive_scan_u32(c, sp->d_hist, hn, sp->d_hist + hn));
    if (width == 9)
      k_radix_scatter<9><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], idx_in, sp->d_key[cur ^ 1], sp->d_idx[cur ^ 1],
                                                         ncap, shift, sp->d_hist, nblocks, n_dev);
    else
      k_radix_scatter<8><<<nblocks, TPB, 0, c->stream>>>(sp->d_key[cur], idx_in, sp->d_key[cur ^ 1], sp->d_idx[cur ^ 1],
                                                         ncap, shift, sp->d_hist, nblocks, n_dev);
    LAUNCH_CHECK(c);
    cur ^= 1;
  }
  *keys_out = sp->d_key[cur];
  *idx_out = sp->d_idx[cur];
  *spare_key = sp->d_key[cur ^ 1];
  *spare_idx = sp->d_idx[cur ^ 1];
  return ISKB_OK;
}

int32_t sort_scratch_ensure(iskb_species *sp) { return ensure_sort_scratch(sp); }

int32_t sp_compact(iskb_species *sp) {
  ISKB_TRY(sp_sync_counts(sp));
  if (sp->h_ndead == 0) return ISKB_OK;
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(ensure_sort_scratch(sp));
  const int64_t n = sp->h_nslots;
  int blocks = (int)((n + TPB - 1) / TPB);
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  k_dead_keys<<<blocks, TPB, 0, c->stream>>>(sp->col[0], n, sp->d_key[0]);
  LAUNCH_CHECK(c);
  return sort_by_keys(sp, n, 1, nullptr);   // 1-bit key
}

// Re-group by tile only (2 radix passes for <= 65535 tiles, no interleave pass): about half the
// cost of the full sort, and the gather of the permutation stays nearly sequential.
int32_t sp_regroup(iskb_species *sp) {
  iskb_ctx *c = sp->ctx;
  ISKB_TRY(sp_sync_counts(sp));
  const int64_t n = sp->h_nslots;
  if (n == 0) return ISKB_OK;
  ISKB_TRY(ensure_sort_scratch(sp));
  const TileGeom tg = tile_geom(c->g);
  int bits = 1;
  while ((1ull << bits) <= (uint64_t)tg.ntiles + 1u) ++bits;
  int blocks = (int)((n + TPB - 1) / TPB);
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  k_tile_keys<<<blocks, TPB, 0, c->stream>>>(sp->col[0], sp->col[1], n, c->g, tg.mtx, tg.ntiles, sp->d_key[0]);
  LAUNCH_CHECK(c);
  return sort_by_keys(sp, n, bits, nullptr, 0);
}

int32_t sp_sort(iskb_species *sp, uint32_t *perm_out_host, bool interleave) {
  iskb_ctx *c = sp->ctx;
  if (!c->has_grid) return iskb_fail(ISKB_E_INVALID, "iskb_grid_set must be called first");
  ISKB_TRY(sp_sync_counts(sp));
  const int64_t n = sp->h_nslots;
  if (n == 0) return interleave ? tdir_build(sp, nullptr, 0) : ISKB_OK;   // an empty species still gets its (empty) tile directory
  ISKB_TRY(ensure_sort_scratch(sp));
  const TileGeom tg = tile_geom(c->g);
  const uint64_t kmax64 = (uint64_t)tg.ntiles * 64u;
  if (kmax64 >= 0xffffffffull)
    return iskb_fail(ISKB_E_UNSUPPORTED, "grid too large for 32-bit cell keys");
  const uint32_t key_max = (uint32_t)kmax64;
  int bits = 1;
  while ((1ull << bits) <= (uint64_t)key_max + 1u) ++bits;
  int blocks = (int)((n + TPB - 1) / TPB);
  if (blocks > c->n_sm * 16) blocks = c->n_sm * 16;
  k_cell_keys<<<blocks, TPB, 0, c->stream>>>(sp->col[0], sp->col[1], n, c->g, tg.mtx, key_max, sp->d_key[0]);
  LAUNCH_CHECK(c);
  return sort_by_keys(sp, n, bits, perm_out_host, interleave ? tg.ntiles : 0u);
}

extern "C" int32_t iskb_sort_by_cell(iskb_species *sp, uint32_t *perm_out) {
  if (!sp) return iskb_fail(ISKB_E_INVALID, "null species");
  ISKB_TRY(sp_compact(sp));   // perm_out refers to the compacted (download) row order
  return sp_sort(sp, perm_out, false);
}

extern "C" int32_t iskb_sort_for_deposit(iskb_species *sp, uint32_t *perm_out) {
  if (!sp) return iskb_fail(ISKB_E_INVALID, "null species");
  ISKB_TRY(sp_compact(sp));
  return sp_sort(sp, perm_out, true);
}
