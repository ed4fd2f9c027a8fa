// iskra_b200 internal declarations shared by the .cu translation units.
// Device state of the hot path: particle SoA columns, rho/phi/E, the
// Poisson operator in separable form, sigma tables and RNG counters -- all resident in HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/iskra_b200.h"

// status bits OR-ed by kernels into ctx->d_status
#define ISKB_ST_OOB 1
#define ISKB_ST_CAPACITY 2
#define ISKB_ST_PK 4
#define ISKB_ST_TOO_FAST 8   // check!'s "particle is too fast" condition (check.jl:41-46): a warning, not an error
#define ISKB_ST_WALK 16      // a tracked particle did not finish its cell walk within the iteration cap

void iskb_set_error(const char *fmt, ...);
int32_t iskb_fail(int32_t code, const char *fmt, ...);

#define CU_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return iskb_fail(ISKB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                 \
  } while (0)

#define ISKB_TRY(expr)            \
  do {                            \
    int32_t rc__ = (expr);        \
    if (rc__ != ISKB_OK) return rc__; \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return iskb_fail(ISKB_E_CUDA, "kernel launch failed: %s (%s:%d)",                     \
                       cudaGetErrorString(e__), __FILE__, __LINE__);                        \
  } while (0)

// Grid description passed by value to kernels (RegularGrids.jl:7-15).
struct GridDev {
  int nx, ny;          // nodes
  double dx, dy;       // dh
  double ox, oy;       // origin (used by wrap!/discard! only, wrap.jl:5,24)
  double Lx, Ly;       // (n-1)*dh  (wrap.jl:3-4)
  double rdx, rdy;     // RN(1/dx), RN(1/dy) for the exact fast division (pic_device.cuh)
  int fast_div;        // 0: always use the IEEE divide instruction sequence
};

// Axis kinds of the separable operator (generalized_poisson.jl:34-68, 205-215, 286-324)
struct PoissonState {
  bool created = false;
  double eps0 = 0.0;
  bool periodic_i = false;   // apply_periodic(ps, 2): couples i=1 <-> i=nx
  bool periodic_j = false;   // apply_periodic(ps, 1): couples j=1 <-> j=ny
  std::vector<uint8_t> isdir;   // host copy: node is Dirichlet
  std::vector<double> dval;     // host copy: Dirichlet value
  bool structure_dirty = true;  // operator structure changed -> rebuild solver
  bool values_dirty = true;     // only Dirichlet values changed -> re-upload dval
  int mode = 0;                 // 1 separable, 2 dense
  // device
  uint8_t *d_isdir = nullptr;
  double *d_dval = nullptr;
  // --- separable solver (transform along the contiguous axis of the working layout) ---
  bool transposed = false;      // working layout is (ny x nx)
  int na = 0, nb = 0;           // working sizes: transform axis a (contiguous), solve axis b
  int a0 = 0, ma = 0;           // unknown range along a: [a0, a0+ma)
  int b0 = 0, mb = 0;           // unknown range along b
  bool b_cyclic = false;
  bool singular = false;
  bool use_fft = false;         // DST-I by FFT (ma+1 power of two)
  int fft_log2M = 0;
  double *d_V = nullptr;        // ma x ma orthonormal eigenvectors (column-major), V[p,k]
  double *d_Vt = nullptr;       // transpose of V
  double *d_lam = nullptr;      // ma eigenvalues of the a-axis operator
  double *d_gam = nullptr;      // Sherman-Morrison gamma per mode
  double *d_msing = nullptr;    // Thomas factors of the pinned singular mode
  int singular_mode = -1;
  double *d_cp = nullptr;       // Thomas c' table  (ma x mb)
  double *d_cpT = nullptr, *d_qT = nullptr;   // the same tables as [k + q*ma] (FFT path, k_thomas_seg)
  double *d_fvec = nullptr;     // Sherman-Morrison factor per mode of the current solve
  double *d_q = nullptr;        // Sherman-Morrison q table (ma x mb), cyclic only
  double *d_qden = nullptr;     // 1/(1 + v.q) per mode
  double *d_w1 = nullptr, *d_w2 = nullptr, *d_w3 = nullptr;   // work arrays (nx*ny)
  double2 *d_tw = nullptr;      // FFT twiddles
  // --- dense solver ---
  double *d_Ainv = nullptr;
  int64_t nn_dense = 0;
  double *d_rowscale = nullptr;   // row equilibration of the dense system (dh^2 / 1 / Neumann row scale)
  // --- operator assembled by the caller (axial grids, generalized_poisson.jl:70-199) ---
  bool has_custom = false;
  std::vector<double> custom_A;   // nn x nn column-major, before Dirichlet / Neumann rows
  // --- sigma dofs and Neumann rows (add_new_dof / apply_neumann, generalized_poisson.jl:217-269) ---
  int n_sigma = 0;
  std::vector<double> sigma;        // host copy of b[sigma dofs], authoritative while sigma_host_newer
  bool sigma_host_newer = false;
  double *d_sigma = nullptr;        // device copy (electrode hits may add to it), capacity MAX_SIGMA
  std::vector<uint8_t> neu_kind;    // per node: 0 none, 1 strip row (:248-255), 2 strip-end row (:256-267)
  std::vector<int32_t> neu_i2, neu_j2, neu_dof;   // neighbour nodes i', j' (0-based) and sigma dof (0-based)
  double *d_neu_coef = nullptr;     // per node: A[row, sigma dof] (0 where no Neumann row)
  int32_t *d_neu_dof = nullptr;
};
constexpr int ISKB_MAX_SIGMA = 64;

struct iskb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  // The field solve (rho -> phi -> E) runs on its own stream so that the next step's re-sort and MCC,
  // which do not read E, overlap it.  ev_rho: rho is final (main stream); ev_E: E is final (field stream).
  cudaStream_t fstream = nullptr;
  // the LAST interaction of a step runs on its own stream next to the advance of the species it does not touch
  cudaStream_t mstream = nullptr;
  cudaEvent_t ev_m0 = nullptr, ev_m1 = nullptr;
  cudaStream_t pstream = nullptr;   // phase 1 of the next step's MCC objects (see mcc_launch)
  cudaEvent_t ev_p0 = nullptr;
  cudaEvent_t ev_rho = nullptr, ev_E = nullptr;
  bool fields_pending = false;   // a solve is in flight on fstream; main-stream users of rho/phi/E must join first
  int64_t launches = 0;
  int n_sm = 148;
  // grid + fields
  bool has_grid = false;
  GridDev g{};
  int bcs[4] = {0, 0, 0, 0};
  double *d_V = nullptr;      // cell_volume, RegularGrids.jl:26-38
  double *d_rho = nullptr, *d_phi = nullptr;
  double2 *d_E2 = nullptr;    // (Ex,Ey) interleaved per node; Ez == 0 is not stored
  int *d_status = nullptr;
  int *h_status = nullptr;    // pinned
  int64_t *h_scratch = nullptr;  // pinned, 16 int64
  PoissonState ps;
  std::vector<iskb_species *> species;
  std::vector<iskb_mcc *> mccs;
  std::vector<iskb_dsmc *> dsmcs;
  // what iskb_step runs, in this order (config.species / config.interactions, ParticleInCell.jl:109-115);
  // unset: everything created on the context, MCC objects before DSMC objects
  bool active_set = false;
  std::vector<iskb_species *> active_species;
  std::vector<std::pair<int, void *>> active_inter;   // (0 = MCC, 1 = DSMC, handle)
  iskb_tracker *tracker = nullptr;   // config.tracker (create_surface_tracker); nullptr == `nothing`
  int after_push[2] = {ISKB_BND_WRAP, ISKB_BND_WRAP};   // default hook wrap!, ParticleInCell.jl:41
  int sort_interval = 0;
  double sort_miss_threshold = 0.0;   // adaptive: sort a species only when its window-miss rate exceeds this
  int sort_max_interval = 0;          //           ... or this many steps have passed
  int sort_full_interval = 0;         // > 0: between full sorts re-group by tile only (cheaper)
  int64_t step_count = 0;
  // small grids (<= PRIV_MAX_NODES): the simple advance deposits into PRIV_COPIES private copies of u (block b uses
  // copy b % PRIV_COPIES) that are summed afterwards -- divides the same-address atomic pressure by PRIV_COPIES
  double *d_upriv = nullptr;
  double fscale = 0.0;               // fixed-point unit of the tile path's deposit: value * fscale is accumulated as int64
  double q0 = 0.0;                   // base charge: every active species carries an integer multiple of it (else 0)
  unsigned long long *d_see_counts = nullptr;   // emit! bookkeeping (see.cu)
  long long *d_rho_int = nullptr;    // sum_s Z_s * ufix_s, the quantity that is all-reduced
  // rho of the last fused tiled step still sits in the fixed-point sums: the FFT solve reads it from there, everybody
  // else calls rho_materialize first (particles.cu); void once the next step zeroes the sums
  bool rho_lazy = false;
  int rho_ns = 0;
  const long long *rho_u[8] = {nullptr};
  long long rho_z[8] = {0};
  bool lean_ok = true;               // iskb_set_lean(ctx, 0) forces the full 88 B/row kernels (A/B measurements)
  int adv_path = 0;                  // 0: tile directory (advance_tile.cu), 1: per-warp windows (advance_fused.cu)
  int pusher_rz = 0;                 // BorisPusher{:rz}: transform_from_cartesian_to_cylindrical! after the push
  bool warn_too_fast = false;        // check!'s "particle is too fast" message condition was seen (sticky until read)
  // optional per-kernel timing of the dominant (advance) kernel, CUDA events on the launch stream
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;     // pairs (start, stop)
  size_t prof_used = 0;
  double prof_ms = 0.0;
  int64_t prof_launches = 0;
  // comm (NCCL through dlopen, see comm.cu)
  int n_ranks = 1, rank = 0;
  void *nccl_comm = nullptr;
};

// counters living in device memory (kernels read loop bounds from here)
enum { CNT_NSLOTS = 0, CNT_NDEAD = 1, CNT_BEGIN = 2, CNT_N = 8 };

struct iskb_species {
  iskb_ctx *ctx = nullptr;
  int64_t cap = 0;
  double q = 0, m = 0, w0 = 0;
  double *col[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x y vx vy vz wg
  uint32_t *id = nullptr;
  double *alt[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // sort ping-pong
  uint32_t *alt_id = nullptr;
  int64_t *d_cnt = nullptr;     // CNT_*
  uint64_t epoch = 0;           // bumped whenever rows change outside the fused step (sp_touch)
  int64_t h_nslots = 0, h_ndead = 0;   // host mirror, valid when !counts_stale
  int64_t h_nsorted = 0;               // rows [0, h_nsorted) are in the sorted layout of the last re-sort
  bool counts_stale = false;
  unsigned long long *d_vmax2 = nullptr;   // bits of an upper bound of |v|^2 over the live rows (+inf = unknown)
  double *d_u = nullptr;        // deposited weights  (particle_to_grid, cloud_in_cell.jl:1-18)
  double *d_n = nullptr;        // density n = u ./ V (kinetic.jl:53)
  // sort scratch
  uint32_t *d_key[2] = {nullptr, nullptr}, *d_idx[2] = {nullptr, nullptr};
  uint32_t *d_hist = nullptr;
  int64_t hist_cap = 0;
  uint64_t sample_calls = 0, see_calls = 0;
  // adaptive re-sort bookkeeping (iskb_step)
  int64_t steps_since_sort = 1 << 30;   // never sorted yet
  int64_t steps_since_full = 1 << 30;   // steps since the last FULL (cell + interleave) sort
  int64_t *h_wstats = nullptr;          // pinned ring (2 x 4) of cnt[3..6] snapshots taken after the advance
  cudaEvent_t ev_wstats[2] = {nullptr, nullptr};
  bool wstats_pending[2] = {false, false};
  int64_t wstats_step = 0;              // number of snapshots issued
  int64_t wstats_sort_mark = 0;         // snapshots with index < this were taken before the last sort
  int64_t last_gmiss = 0;
  double miss_rate = 1.0;               // gather-miss fraction of the last measured step
  // rows the tiled advance leaves to the surface tracker (cells next to a surface), filled per launch
  uint32_t *d_trk_list = nullptr;
  unsigned *d_trk_n = nullptr;
  // ---- tile directory and incremental re-group (advance_tile.cu) ----
  uint32_t *d_ts[2] = {nullptr, nullptr};   // [ntiles+1] first row of every tile ([0] current, [1] next layout)
  uint32_t *d_wr = nullptr;                 // [warps+1] first tile of every warp of the advance grid
  uint32_t *d_tcnt = nullptr;               // [ntiles*NCODE] rows per (storage tile, code), rebuilt by every advance
  uint32_t *d_tbase = nullptr;              // [ntiles*NCODE] destination bases of a re-grouping launch
  uint32_t *d_seg = nullptr;                // scan scratch of the re-group (3*ntiles+3 words + partials)
  uint32_t *d_tstart = nullptr, *d_tailbase = nullptr, *d_tn = nullptr;   // tail merge of a MOVE (advance_tile.cu)
  bool n_in_ufix = false;                   // species.n of the last fused step not formed yet (particles.cu)
  int64_t tail_rows = 0;                    // rows of the unsorted tail at the latest snapshot (sizes the merge grid)
  uint8_t *d_code = nullptr, *alt_code = nullptr;   // per row: where its position lies relative to its storage tile
  uint2 *d_mlist = nullptr;                 // rows outside their tile's window: (source row, destination row)
  unsigned *d_mlist_n = nullptr;
  bool tdir_valid = false;                  // d_ts describes the current row layout
  bool marks_valid = false;                 // d_code / d_tcnt describe the current positions
  int64_t steps_since_move = 0;
  int64_t moves = 0, full_sorts = 0;        // statistics (iskb_species_sort_stats)
  double tail_frac = 0.0, dead_frac = 0.0;  // from the last snapshot
  long long *d_ufix = nullptr;              // deposited weights of the tile path in fixed point (nx*ny)
  double wmax = 0.0;                        // largest weight any slot may carry (w0 unless an upload says otherwise)
  unsigned *d_ticket = nullptr;             // chunk dispenser of the advance grid
  unsigned long long *d_vz2max = nullptr;   // bits of a bound of v_z^2 (the lean advance does not touch the column)
  bool vz2_known = false;
  bool wg_uniform = true;                   // every wg[p] == w0 (configuration.jl:99) until an upload says otherwise
  bool drifting = false;                    // some launch saw more than 1e-5 of the rows outside their window
  int64_t *h_tstats = nullptr;              // pinned ring (2 x 10): cnt[0..8), -, rows in the tile directory
  cudaEvent_t ev_tstats[2] = {nullptr, nullptr};
  bool tstats_pending[2] = {false, false};
  int64_t tstats_step = 0, tstats_sort_mark = 0;
};

struct MccProc {
  int kind;
  double threshold;
  int offset, len;      // into the concatenated tables
  iskb_species *product;
};

struct iskb_mcc {
  iskb_ctx *ctx = nullptr;
  iskb_species *source = nullptr;
  double tq = 0, tm = 0, tT = 0;
  int N = 0;
  std::vector<MccProc> procs;
  double max_sigma_g = 0, m_eV = 0;
  double max_n0 = 0;
  bool uniform_n = false;            // the target density is the same on every node
  double eps_hi = 0;                 // largest tabulated energy
  double sup_total = 0;              // sup of sum_k sigma_k*g on [0, eps_hi]
  std::vector<double> sup_sigma_g;   // per process: sup of sigma_k*g on [0, eps_hi]
  std::vector<double> sig_last;      // per process: sigma_k at its last knot
  double *d_pk = nullptr;            // device pruning bounds of the current call
  uint64_t seed = 0;
  uint64_t calls = 0, cur_call = 0;
  // phase 1 of the next step already ran (iskb_step): valid while dt and the source species are what they were
  bool pre_valid = false;
  double pre_dt = 0.0;
  uint64_t pre_epoch = 0;
  cudaEvent_t ev_pre = nullptr;
  double *d_tn = nullptr;       // target density on nodes
  double *d_eps = nullptr, *d_sig = nullptr;
  void *d_procs = nullptr;      // MccProcDev[N]
  unsigned long long *d_stats = nullptr;   // [0] candidates, [1] collisions, [2+k] per process
  float *d_nu = nullptr;        // nx*ny*N counters of the last perform (lazy)
  uint32_t *d_cand = nullptr;   // candidate rows of the current call
  uint2 *d_coll = nullptr;      // (row, process) of accepted collisions
  unsigned int *d_lists_cnt = nullptr;
  int64_t totals[2 + 16] = {0};
};

// SurfaceTracker{2}  ParticleInCell/src/pic/surfaces/build.jl:13-18, flattened for the device:
// cells (i,j) with 0 <= i <= nx, 0 <= j <= ny (ghost ring included, build.jl:33-44); per cell four
// directed faces  0:(i,j-1)  1:(i+1,j)  2:(i,j+1)  3:(i-1,j)  holding a surface id (0 = no entry).
struct TrackerDev {
  int nx, ny;                 // nodes
  double dh;                  // st.dh = grid.dh[1]  (build.jl:97)
  const uint8_t *face;        // 4 * (nx+1) * (ny+1)
  const uint8_t *tracked;     // (nx+1) * (ny+1): cell is on either side of some key (build.jl:86-93)
  const int32_t *s_kind;      // per surface id
  const int32_t *s_dof;       // sigma dof (0-based) of a floating electrode, -1 otherwise
  const double *s_area;
  double *s_dq;               // collected charge per surface
  double *sigma;              // device sigma right-hand side (only touched when route_hits != 0)
  int route_hits;
};

struct iskb_tracker {
  iskb_ctx *ctx = nullptr;
  int default_kind = ISKB_SURF_ABSORBING;
  std::vector<uint8_t> h_face, h_tracked;
  std::vector<int32_t> h_kind, h_dof;
  std::vector<double> h_area;
  bool dirty = true;
  uint8_t *d_face = nullptr, *d_tracked = nullptr;
  int32_t *d_kind = nullptr, *d_dof = nullptr;
  double *d_area = nullptr, *d_dq = nullptr;
  int route_hits = 0;
  // track! -> check! hand-over of the operator-level API (one species at a time)
  iskb_species *trk_sp = nullptr;
  int64_t trk_rows = 0, trk_cap = 0;
  int32_t *d_ti = nullptr, *d_tj = nullptr;
  double *d_thx = nullptr, *d_thy = nullptr;
  unsigned long long *d_counts = nullptr;   // [0] tracked, [1] absorbed
};
constexpr int ISKB_MAX_SURFACES = 250;
constexpr int PRIV_COPIES = 64;
constexpr int64_t PRIV_MAX_NODES = 16384;

// ---- internal entry points across translation units ------------------------------------------
int32_t sp_sync_counts(iskb_species *sp);
int32_t sp_compact(iskb_species *sp);
int32_t sp_sort(iskb_species *sp, uint32_t *perm_out_host, bool interleave);
int32_t sp_regroup(iskb_species *sp);
int32_t fields_join(iskb_ctx *c);   // main stream waits for a field solve in flight (stream-ordered, no host sync)
int32_t sp_ensure_alt(iskb_species *sp);
int32_t ctx_check_status(iskb_ctx *ctx);
int32_t poisson_prepare(iskb_ctx *ctx);
int32_t poisson_solve(iskb_ctx *ctx);
int32_t poisson_free(iskb_ctx *ctx);
int32_t mcc_launch(iskb_mcc *mcc, double dt, bool count_nu, cudaStream_t st = nullptr, int phase = 3);
int32_t mcc_discard_pre(iskb_mcc *mcc);
int32_t comm_allreduce_sum(iskb_ctx *ctx, double *d_buf, int64_t n);
int32_t comm_destroy(iskb_ctx *ctx);
int32_t launch_advance(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit,
                       int64_t first_slot_from_cnt_begin);
int32_t launch_rho_finalize(iskb_ctx *ctx, const std::vector<iskb_species *> *list = nullptr);
int32_t launch_rho_finalize_fixed(iskb_ctx *ctx, const std::vector<iskb_species *> &list);
int32_t rho_materialize(iskb_ctx *ctx);
int32_t comm_allreduce_sum_i64(iskb_ctx *ctx, long long *d_buf, int64_t n);
int32_t sp_vmax_unknown(iskb_species *sp);
int32_t sp_vmax_reset(iskb_species *sp);
int32_t dsmc_launch(iskb_dsmc *d, double dt, bool want_nu);
int32_t dsmc_free(iskb_dsmc *d);
int32_t tracker_prepare(iskb_tracker *st, TrackerDev *out);
int32_t tracker_free(iskb_tracker *st);
int32_t launch_advance_tracked(iskb_species *sp, double dt, int mode_x, int mode_y, bool deposit);
int32_t launch_advance_tiled_tracked(iskb_species *sp, double dt, int mode_x, int mode_y);
int32_t launch_advance_tracked_list(iskb_species *sp, double dt, int mode_x, int mode_y);
int32_t poisson_sigma_device(iskb_ctx *ctx, double **d_sigma_out);
int32_t launch_advance_tile(iskb_species *sp, double dt, int mode_x, int mode_y, bool move, bool mark);
int32_t tdir_build(iskb_species *sp, const uint32_t *sorted_keys, int64_t n);
void tdir_free(iskb_species *sp);
void sp_touch(iskb_species *sp);   // rows were changed outside the tile-aware advance: its marks no longer hold
int32_t exclusive_scan_u32(iskb_ctx *c, uint32_t *d, int64_t n, uint32_t *partial);
int32_t prof_begin(iskb_ctx *ctx);
int32_t prof_end(iskb_ctx *ctx);
