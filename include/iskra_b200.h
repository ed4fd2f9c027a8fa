/* iskra_b200 -- C ABI of the B200-native particle hot path of bchaber/iskra.
 *
 * This header is the drop-in boundary.  The reference (pure Julia) has no FFI today: the
 * path sits behind multiple dispatch on generic functions of its ParticleInCell,
 * FiniteDifferenceMethod, RegularGrids and Chemistry modules.  Every entry point below
 * names the reference function (file:line under /root/reference) whose work it replaces;
 * the Julia-side `ccall` methods a maintainer would add are in INTEGRATION.md and
 * julia/ParticleInCellB200.jl, the Python/ctypes mirror used by the tests is iskra_b200/.
 *
 * Conventions
 *   - plain C, opaque handles, no exceptions cross the boundary; every function returns an
 *     int32 status (ISKB_OK = 0, errors < 0); iskb_last_error() gives the message.
 *   - all host arrays are COLUMN-MAJOR exactly as Julia lays them out: x :: N x 2, v :: N x 3
 *     (kinetic.jl:2-3) with leading dimension `ld` (= the species capacity N for the
 *     reference's arrays), grid fields nx x ny (x 3) with i fastest.
 *   - host pointers are only read/written during the call (the library copies); device state
 *     (particles, rho, phi, E, sigma tables, RNG counters) lives in HBM inside the handles.
 *   - indices returned to the host are 1-based like the reference's.
 *   - calls on one context are stream-ordered on the stream given to iskb_set_stream()
 *     (default: a private non-blocking stream); calls that return host data synchronise.
 *   - single caller thread per context (the reference loop is single-threaded,
 *     ParticleInCell.jl:84-139).
 */
#ifndef ISKRA_B200_H
#define ISKRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iskb_ctx iskb_ctx;
typedef struct iskb_species iskb_species;
typedef struct iskb_mcc iskb_mcc;
typedef struct iskb_tracker iskb_tracker;
typedef struct iskb_dsmc iskb_dsmc;

/* status codes */
#define ISKB_OK 0
#define ISKB_E_INVALID (-1)      /* bad argument / call order */
#define ISKB_E_CUDA (-2)         /* CUDA runtime error (message has the cudaError string) */
#define ISKB_E_CAPACITY (-3)     /* species capacity exceeded (reference: BoundsError, mcc.jl:195-197, TODO:2) */
#define ISKB_E_PMAX (-4)         /* max_Pt > 1/N            (reference: @assert, mcc.jl:244-246) */
#define ISKB_E_PK (-5)           /* P_k > 1                 (reference: @assert, mcc.jl:273-279) */
#define ISKB_E_OOB (-6)          /* live particle outside the grid at gather/deposit (reference: BoundsError) */
#define ISKB_E_NCCL (-7)         /* NCCL missing or failed */
#define ISKB_E_UNSUPPORTED (-8)  /* configuration outside the implemented scope */
#define ISKB_E_SINGULAR (-9)     /* dense Poisson operator is singular */

/* boundary handling per axis after the push (ParticleInCell/src/pic/surfaces/wrap.jl) */
#define ISKB_BND_NONE 0
#define ISKB_BND_WRAP 1          /* wrap!    wrap.jl:20-33 */
#define ISKB_BND_DISCARD 2       /* discard! wrap.jl:1-18  */

/* grid edge kinds, order left,right,bottom,top (RegularGrids.jl:11,55-57) */
#define ISKB_BC_OPEN 0
#define ISKB_BC_PERIODIC 1

/* edges for iskb_poisson_apply_dirichlet_edge */
#define ISKB_EDGE_LEFT 0         /* i = 1  */
#define ISKB_EDGE_RIGHT 1        /* i = nx */
#define ISKB_EDGE_BOTTOM 2       /* j = 1  */
#define ISKB_EDGE_TOP 3          /* j = ny */

/* collision kinds (Chemistry/src/mcc.jl:4-8) */
#define ISKB_MCC_ELASTIC_ISOTROPIC 0
#define ISKB_MCC_ELASTIC_BACKWARD 1
#define ISKB_MCC_INELASTIC_BACKWARD 2
#define ISKB_MCC_EXCITATION 3
#define ISKB_MCC_IONIZATION 4

/* surface kinds (ParticleInCell/src/pic/surfaces/build.jl:8-11, circuit_coupling.jl:5-16) */
#define ISKB_SURF_PERIODIC 0            /* PeriodicSurface: generic no-op hit!, hit.jl:26-31 */
#define ISKB_SURF_ABSORBING 1           /* AbsorbingSurface            hit.jl:32-38 */
#define ISKB_SURF_REFLECTIVE 2          /* ReflectiveSurface (specular) hit.jl:39-56 */
#define ISKB_SURF_ELECTRODE_FIXED 3     /* FixedPotentialElectrode     circuit_coupling.jl:55-61 */
#define ISKB_SURF_ELECTRODE_FLOATING 4  /* FloatingPotentialElectrode  circuit_coupling.jl:44-53 */

/* ---- lifecycle -------------------------------------------------------------------------- */
int32_t iskb_version(void);
const char *iskb_last_error(void);
/* One context per process/GPU; replaces the implicit single address space of
 * ParticleInCell.solve (ParticleInCell.jl:84-100: phi, rho, E, B = zeros(size(grid)...)). */
int32_t iskb_create(int32_t device, iskb_ctx **out);
int32_t iskb_destroy(iskb_ctx *ctx);
/* Run all work of this context on the caller's CUDA stream (a cudaStream_t); NULL restores
 * the private stream. */
int32_t iskb_set_stream(iskb_ctx *ctx, void *cuda_stream);
int32_t iskb_synchronize(iskb_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int32_t iskb_launch_count(iskb_ctx *ctx, int64_t *out);

/* Time the dominant kernel (fused advance) with CUDA events on the launching stream.
 * iskb_profile_read synchronises, returns the accumulated milliseconds and launch count since the
 * last read, and resets both. */
int32_t iskb_profile_enable(iskb_ctx *ctx, int32_t on);
int32_t iskb_profile_read(iskb_ctx *ctx, double *ms_advance, int64_t *launches_advance);

/* ---- multi-GPU: particles sharded by index slice, rho all-reduced --------- */
/* The reference has no communication layer; this is new.  id128 is an ncclUniqueId. */
int32_t iskb_comm_unique_id(void *id128);
int32_t iskb_comm_init(iskb_ctx *ctx, int32_t n_ranks, int32_t rank, const void *id128);

/* ---- grid: create_uniform_grid(xx, yy; left, right, bottom, top)  RegularGrids.jl:55-69 -- */
/* Also fixes cell_volume(grid) (RegularGrids.jl:26-38) and allocates rho, phi, E in HBM.     */
int32_t iskb_grid_set(iskb_ctx *ctx, int32_t nx, int32_t ny, double dx, double dy,
                      double ox, double oy, const int32_t bcs[4]);
int32_t iskb_cell_volume(iskb_ctx *ctx, double *V_out /* nx*ny */);

/* ---- field solve: FiniteDifferenceMethod/src/generalized_poisson.jl ---------------------- */
/* create_poisson_solver(grid, eps0)  :27-30 -> :34-68 (5-point operator, all edges open) */
int32_t iskb_poisson_create(iskb_ctx *ctx, double eps0);
/* apply_periodic(ps, axis)  :286-324  (axis 1 couples j=1<->ny, axis 2 couples i=1<->nx) */
int32_t iskb_poisson_apply_periodic(iskb_ctx *ctx, int32_t axis);
/* apply_dirichlet(ps, nodes::BitArray, phi0)  :205-215 ; mask is nx*ny bytes, column-major */
int32_t iskb_poisson_apply_dirichlet(iskb_ctx *ctx, const uint8_t *mask, double phi0);
/* Same for a whole edge without shipping a mask (the per-step RF drive of
 * problem/11_rf_discharge.jl:95). */
int32_t iskb_poisson_apply_dirichlet_edge(iskb_ctx *ctx, int32_t edge, double phi0);
/* Dense matrix exactly as the reference assembles it (debug / parity): A is nn*nn column-major */
int32_t iskb_poisson_get_dense(iskb_ctx *ctx, double *A_out, double *b_out);
/* 1 = separable fast solver, 2 = dense inverse; decided from the boundary structure */
int32_t iskb_poisson_mode(iskb_ctx *ctx, int32_t *mode_out);
/* phi = calculate_electric_potential(solver, -rho) :372-378 ; E = calculate_electric_field(solver, phi)
 * :398-410 ; B = calculate_magnetic_field == 0 :412-419 is never materialised.
 * (ParticleInCell.jl:126-128).  rho -> phi -> E stays in HBM. */
int32_t iskb_field_solve(iskb_ctx *ctx);
/* any pointer may be NULL; E is nx*ny*3 column-major (Ez == 0) */
int32_t iskb_fields_download(iskb_ctx *ctx, double *rho, double *phi, double *E);
int32_t iskb_fields_upload(iskb_ctx *ctx, const double *rho, const double *phi, const double *E);

/* ---- species: KineticSpecies{2,3}  ParticleInCell/src/pic/kinetic.jl:1-18 ----------------- */
/* create_kinetic_species(name, N, q, m, weight)  problem/configuration.jl:95-102 */
int32_t iskb_species_create(iskb_ctx *ctx, int64_t capacity, double q, double m, double w0,
                            iskb_species **out);
/* x: np x 2, v: np x 3 with leading dimension ld; wg, id: `capacity` entries or NULL (defaults
 * ones*w0 and 1..N, kinetic.jl:15-18).  Sets species.np = np. */
int32_t iskb_species_upload(iskb_species *sp, const double *x, const double *v, const double *wg,
                            const uint32_t *id, int64_t np, int64_t ld);
/* Live particles come back in rows 1..np (device order; compare keyed by id); any pointer may
 * be NULL. */
int32_t iskb_species_download(iskb_species *sp, double *x, double *v, double *wg, uint32_t *id,
                              int64_t ld);
int32_t iskb_species_np(iskb_species *sp, int64_t *np_out);
/* Diagnostics of the fused advance kernel since the last call (then reset): rows whose gather
 * missed the shared window, rows whose deposit missed it, window moves, deposit rounds. */
int32_t iskb_species_window_stats(iskb_species *sp, int64_t out[4]);
/* sample!(src::MaxwellianSource, species, dt)  sources.jl:24-34 with n given: appends n
 * particles x = rand*wx + dx, v = randn*wv + dv  (device Philox stream, key = seed). */
int32_t iskb_species_sample_maxwellian(iskb_species *sp, int64_t n, const double wx[2],
                                       const double dx[2], const double wv[3], const double dv[3],
                                       uint64_t seed);
/* iHe.x .= e.x ; iHe.np = e.np   (problem/10_two_streams.jl:66-68); v filled with v_fill */
int32_t iskb_species_copy_positions(iskb_species *dst, iskb_species *src, const double v_fill[3]);
/* remove!(sp, i)  kinetic.jl:20-27 ; i is the 1-based row of the download order.  Exactly the reference's
 * swap with the last row: x, v of row np move to row i, wg[i] = wg[np], wg[np] = w0, id[i] <-> id[np], np -= 1. */
int32_t iskb_species_remove(iskb_species *sp, int64_t i);
/* add!(src, dst)  kinetic.jl:29-37: appends x, v of src's live rows to dst (wg and id of dst stay, like the
 * reference); ISKB_E_CAPACITY where the reference would raise BoundsError. */
int32_t iskb_species_add(iskb_species *src, iskb_species *dst);
/* remove_particles!(part, dh, matches)  kinetic.jl:39-50 with matches(i,j) given as a byte mask over the
 * cells (nx*ny, column-major, entry (i-1) + (j-1)*nx for the 1-based lower-left node (i,j)).  Survivors keep
 * their ids; row order differs from the reference's swap sequence (compare keyed by id). */
int32_t iskb_species_remove_in_cells(iskb_species *sp, const uint8_t *cell_mask, int64_t *n_removed);
/* density field n of the last iskb_density()/iskb_step() (part.n, kinetic.jl:4) */
int32_t iskb_species_density_download(iskb_species *sp, double *n_out);

/* ---- per-step operators (each usable on its own for parity tests) ------------------------ */
/* particle_cell(px, p, dh)  ParticleInCell.jl:28-35 : 1-based lower-left node and fractions */
int32_t iskb_cell_index(iskb_species *sp, int32_t *i_out, int32_t *j_out, double *hx_out,
                        double *hy_out);
/* Stable counting (radix) sort of the live particles by cell key; new, enables tiled deposition.
 * perm_out[k] (0-based) = previous row of the particle now in row k.  Key layout: DESIGN.md. */
int32_t iskb_sort_by_cell(iskb_species *sp, uint32_t *perm_out);
/* The layout the fused step keeps between re-sorts: the cell sort above followed by a
 * deterministic round-robin over the cells inside every 8x8 tile (DESIGN.md "row order"), so
 * that the 32 rows of a warp batch deposit into 32 different cells. */
int32_t iskb_sort_for_deposit(iskb_species *sp, uint32_t *perm_out);
/* grid_to_particle(grid, part, (i,j)->E[i,j,:])  cloud_in_cell.jl:20-36 ; out is np x 3, ld = np */
int32_t iskb_gather(iskb_species *sp, double *partE_out);
/* push_particles!(::BorisPusher{:xy}, part, E, B, dt)  pushers.jl:8-11,37-50 with B == 0.
 * partE (np x 3, ld = np) may be NULL: then E is gathered on the fly from the context's field. */
int32_t iskb_push(iskb_species *sp, const double *partE, double dt);
/* discard!(part, grid; dims) then wrap!(part, grid; dims)  wrap.jl:1-33 ; mode per axis */
int32_t iskb_boundary(iskb_species *sp, int32_t mode_x, int32_t mode_y, int64_t *n_removed);
/* density(species, grid) = particle_to_grid(species, grid, p->wg[p]) ./ cell_volume(grid)
 * kinetic.jl:53, cloud_in_cell.jl:1-18.  n_out (nx*ny) may be NULL. */
int32_t iskb_density(iskb_species *sp, double *n_out);
/* fill!(rho, 0) ; rho .+= part.n .* part.q  ParticleInCell.jl:118-121 */
int32_t iskb_rho_zero(iskb_ctx *ctx);
int32_t iskb_rho_accumulate(iskb_ctx *ctx, iskb_species *sp);
/* sum of rho over ranks (no-op for one rank) */
int32_t iskb_rho_allreduce(iskb_ctx *ctx);

/* ---- secondary-electron emission at a wall: emit!(primary, secondary, grid, material; boundary, gamma_t, gamma_e, gamma_i)
 * Chemistry/src/see.jl:114-181.  boundary = ISKB_EDGE_LEFT / RIGHT / BOTTOM / TOP (:all has a zero wall normal in the
 * reference, see.jl:94, and yields NaN secondaries there: rejected).  The emission coefficients are the reference's own
 * closures with their parameters by value: coef[12] = { vaughan(w0, w0max, gamma0max, ks) (:17-26), elastic(we, wemax,
 * gamma_e_max, Delta_e, r_e) (:29-42; gamma_e_max < 0: gamma_e = gamma_0 == 0), inelastic(r_i) (:45-49; r_i < 0: gamma_0),
 * secondary(r_e, r_i) (:52-56) }.  Rows beyond the wall are reflected (elastic / inelastic), emit true secondaries into
 * `secondary` (energy ~ LogNormal(1.65, 1.1) eV, cosine law about the inward normal) or are absorbed (removed).
 * counts_out[4] (nullable) = elastic, inelastic, injected secondaries, absorbed. */
int32_t iskb_see_emit(iskb_species *primary, iskb_species *secondary, int32_t boundary, const double *coef,
                      uint64_t seed, int64_t *counts_out);

/* ---- fused fast path: the loop body of ParticleInCell.solve  ParticleInCell.jl:102-135 ---- */
/* after_push hook (ParticleInCell.jl:41; problem scripts override it), applied to every species */
int32_t iskb_set_after_push(iskb_ctx *ctx, int32_t mode_x, int32_t mode_y);
/* interval > 0 keeps the rows of every species grouped by 8x8-cell tile (new relative to the reference: the row
 * order of a KineticSpecies carries no meaning there) and runs advance! + density as one fused kernel;
 * 0 = one simple kernel per operator.  Fixed mode: the rows are re-grouped every `interval` steps. */
int32_t iskb_set_sort_interval(iskb_ctx *ctx, int32_t interval);
/* Adaptive mode (miss_threshold > 0): a species is re-grouped when the fraction of its rows that lay outside
 * their tile's shared-memory window in the last measured step exceeds miss_threshold, when discarded rows make up
 * more than 2 % of its slots, or after max_interval steps (0 = no limit).  Slow species (ions) are then
 * re-grouped rarely, fast ones (electrons) every few steps. */
int32_t iskb_set_sort_policy(iskb_ctx *ctx, double miss_threshold, int32_t max_interval);
/* A re-group is folded into the advance kernel (rows are written to their new place instead of in place); a FULL
 * sort (radix sort by cell) happens when a species has no tile directory yet, when its unsorted tail (appended
 * rows) exceeds 1 % of its rows, and every full_interval steps if full_interval > 0. */
int32_t iskb_set_sort_full_interval(iskb_ctx *ctx, int32_t full_interval);
/* 0 (default): tile directory + incremental re-group (advance_tile.cu); 1: per-warp windows that follow the
 * rows (advance_fused.cu, re-grouped by radix sort) -- the path a context with a surface tracker always takes. */
int32_t iskb_set_advance_path(iskb_ctx *ctx, int32_t path);
/* 1 (default): the fused advance does not touch what cannot change -- v_z under B == 0, E_z == 0 (pushers.jl:41-48 add
 * zero to it) and the wg column of a species whose weights all equal w0 -- 64 instead of 88 B per particle-step,
 * bit-identical results.  0: every launch reads and writes all columns (A/B measurements). */
int32_t iskb_set_lean(iskb_ctx *ctx, int32_t on);
/* out[0] full sorts, out[1] re-grouping launches so far, out[2] steps since the last full sort, out[3] since the last
 * re-group; from the newest statistics snapshot the policy has read: out[4] slots, out[5] discarded rows among them,
 * out[6] rows covered by the tile directory (slots beyond it form the unsorted tail); out[7] reserved */
int32_t iskb_species_sort_stats(iskb_species *sp, int64_t out[8]);
/* n_steps iterations of: MCC, then DSMC (registered interactions, each kind in creation order) -> advance! every species
 * (gather, push, after_push) -> density / rho -> all-reduce -> phi -> E.  With a surface tracker on
 * the context advance! is track! -> gather -> push -> check! -> after_push (ParticleInCell.jl:56-61);
 * the circuit (advance!(circuit, ...), :116) stays on the host between steps: call iskb_step(ctx, dt, 1)
 * and iskb_poisson_sigma_add from the after_loop hook's side. */
int32_t iskb_step(iskb_ctx *ctx, double dt, int32_t n_steps);
/* What iskb_step runs: config.species (kinetic ones) and config.interactions in the reference's order
 * (ParticleInCell.jl:109-115).  interactions[] holds iskb_mcc* / iskb_dsmc* handles.  Species or interaction
 * objects that merely exist on the context (a source buffer for add!, a scratch species for density) are then
 * neither advanced nor deposited.  n_species < 0 restores the default: everything created on the context, MCC
 * objects before DSMC objects, each in creation order. */
int32_t iskb_step_set_active(iskb_ctx *ctx, iskb_species *const *species, int32_t n_species,
                             void *const *interactions, int32_t n_interactions);
int32_t iskb_ctx_counts(iskb_ctx *ctx, int64_t *n_species, int64_t *n_mcc, int64_t *n_dsmc);
/* The field solve of a step runs on a private stream so that the next step's re-sort and MCC overlap
 * it.  Every entry point that touches rho / phi / E joins it automatically; call this to make the
 * ctx stream wait for it explicitly (stream-ordered, no host sync), e.g. before recording a timing
 * event on the ctx stream. */
int32_t iskb_stream_join(iskb_ctx *ctx);

/* ---- surfaces, electrodes, circuit coupling --------------------------- */
/* add_new_dof(ps, :sigma)  generalized_poisson.jl:217-230 -> 1-based index of the new sigma dof.
 * Systems with sigma dofs are solved by the dense path (grids up to 8192 nodes). */
int32_t iskb_poisson_add_dof(iskb_ctx *ctx, int32_t *dof_out);
/* apply_neumann(ps, nodes, dof)  :235-269 (eps_r == 1); mask is nx*ny bytes, column-major.  Rows of
 * nodes on a vertical strip and at its ends are replaced, others are left alone like the reference. */
int32_t iskb_poisson_apply_neumann(iskb_ctx *ctx, const uint8_t *mask, int32_t dof);
/* get_rhs(ps, :sigma, dof) .= v / .+= dv / read   :367-370 (circuit_coupling.jl:41-42) */
int32_t iskb_poisson_sigma_set(iskb_ctx *ctx, int32_t dof, double value);
int32_t iskb_poisson_sigma_add(iskb_ctx *ctx, int32_t dof, double delta);
int32_t iskb_poisson_sigma_get(iskb_ctx *ctx, int32_t dof, double *out);
/* number of unknowns of the dense system (nx*ny + sigma dofs); iskb_poisson_get_dense returns that size */
int32_t iskb_poisson_dense_size(iskb_ctx *ctx, int64_t *n_out);
/* get_solution(ps, :phi, i, j)  :363-364 ; 1-based node */
int32_t iskb_phi_at(iskb_ctx *ctx, int32_t i, int32_t j, double *out);

/* create_surface_tracker(grid, ds)  build.jl:95-100 -> :76-87: default surface `default_kind` on the
 * four domain faces, from the inside out.  One tracker per context (config.tracker). */
int32_t iskb_tracker_create(iskb_ctx *ctx, int32_t default_kind, iskb_tracker **out);
/* track_surface!(st, nodes, surface)  build.jl:109-111 -> :46-60.  For electrodes: sigma_dof (1-based,
 * 0 for a fixed one) and area (create_electrode, problem/configuration.jl:45-53).  Returns a surface id. */
int32_t iskb_tracker_track_surface(iskb_tracker *st, const uint8_t *node_mask, int32_t kind,
                                   int32_t sigma_dof, double area, int32_t *surface_id_out);
/* face lookup get(st, ((i,j),(k,l)), nothing)  build.jl:113-117: kind of the surface or -1 */
int32_t iskb_tracker_lookup(iskb_tracker *st, int32_t i, int32_t j, int32_t k, int32_t l, int32_t *kind_out);
/* track!(st, part, dt)  track.jl:42-52: remembers cell and fractions of every particle in a cell next
 * to a surface.  Must be followed by iskb_push and iskb_tracker_check on the same species with no
 * other call in between.  n_tracked may be NULL. */
int32_t iskb_tracker_track(iskb_tracker *st, iskb_species *sp, double dt, int64_t *n_tracked);
/* check!(st, part, dt)  check.jl:39-68: walks every tracked particle through the cell faces it
 * crossed (check, :17-36), applies hit! (hit.jl:26-56, circuit_coupling.jl:44-61), removes the
 * absorbed ones.  too_fast: the reference's "particle is too fast" message condition (:41-46). */
int32_t iskb_tracker_check(iskb_tracker *st, iskb_species *sp, double dt, int64_t *n_absorbed,
                           int32_t *too_fast);
/* Sticky flag: some particle exceeded dh/dt per step since the last call (the reference prints
 * "ERROR: ... particle is too fast", check.jl:41-46, and carries on); reading clears it. */
int32_t iskb_warning_too_fast(iskb_ctx *ctx, int32_t *out);
/* electrode.dq: charge collected by the surface so far (circuit_coupling.jl:49-50); reset != 0 zeroes it
 * (foo!, :30).  With particles sharded over several ranks this is the rank's own share. */
int32_t iskb_surface_charge(iskb_tracker *st, int32_t surface_id, double *dq_out, int32_t reset);
/* The reference builds its floating electrodes with sigma and phi swapped (problem/configuration.jl:69-70
 * vs circuit_coupling.jl:11-16), so collected charge never reaches the sigma right-hand side.  on != 0
 * routes dq/area into it (the evident intent); default 0 = the reference's behaviour. */
int32_t iskb_tracker_route_hits_to_sigma(iskb_tracker *st, int32_t on);

/* ---- axisymmetric r-z variant ------------------------------------------------- */
/* create_axial_grid(rr, zz)  RegularGrids.jl:84-97 is iskb_grid_set with (dr, dz); gather and deposit of an
 * AxialGrid{2} (cloud_in_cell.jl:38-73) are the Cartesian ones.  What differs: */
/* cell_volume(g::AxialGrid{2})  RegularGrids.jl:40-53 (ring volumes) -- or any other node volume array */
int32_t iskb_cell_volume_set(iskb_ctx *ctx, const double *V /* nx*ny */);
/* create_poisson_solver(grid::AxialGrid{2}, eps0)  generalized_poisson.jl:70-199: the caller assembles the
 * operator with the reference's own assembler and hands it over (nn*nn, column-major, nn = nx*ny); Dirichlet /
 * Neumann rows set through this API are applied on top as in the reference.  Dense path (nn <= 8192). */
int32_t iskb_poisson_set_dense(iskb_ctx *ctx, const double *A, int64_t nn);
/* create_boris_pusher() / create_axial_boris_pusher()  pushers.jl:5-6 for iskb_step and iskb_push */
#define ISKB_PUSHER_XY 0
#define ISKB_PUSHER_RZ 1   /* push_in_cartesian! then transform_from_cartesian_to_cylindrical!  pushers.jl:13-17 */
int32_t iskb_set_pusher(iskb_ctx *ctx, int32_t kind);
/* transform_from_cartesian_to_cylindrical!(part, dt)  pushers.jl:52-66 on its own */
int32_t iskb_transform_cylindrical(iskb_species *sp, double dt);

/* ---- MCC: Chemistry/src/mcc.jl ----------------------------------------------------------- */
/* mcc(reactions) -> MonteCarloCollisions(collisions)  mcc.jl:313-320, :27-51.
 * One source species colliding with one fluid target (accept, :291-311).  Process k has
 * kind[k], threshold[k] (eV) and a CrossSection table (cross_section.jl:3-14) of table_len[k]
 * rows stored back to back in eps[] / sigma[].  ion_product[k] is the product species != source
 * of an ionisation (or NULL).  target_n is the fluid density on the nodes (nx*ny). */
int32_t iskb_mcc_create(iskb_ctx *ctx, iskb_species *source, double target_q, double target_m,
                        double target_T, const double *target_n, int32_t n_proc,
                        const int32_t *kind, const double *threshold, const int32_t *table_len,
                        const double *eps, const double *sigma, iskb_species *const *ion_product,
                        uint64_t seed, iskb_mcc **out);
/* mcc.max_sigma_g and mcc.m  (mcc.jl:20-21) */
int32_t iskb_mcc_constants(iskb_mcc *mcc, double *max_sigma_g, double *m_eV);
/* PIC.perform!(mcc, E, dt, config)  mcc.jl:231-289.  nu_out: nx*ny*N collision counters of
 * this call (may be NULL).  Per-particle Philox null-collision test, see DESIGN.md (H7). */
int32_t iskb_mcc_perform(iskb_mcc *mcc, double dt, double *nu_out, int64_t *n_candidates,
                         int64_t *n_collisions);
/* totals accumulated by iskb_step since creation: [candidates, collisions, per-process...] */
int32_t iskb_mcc_totals(iskb_mcc *mcc, int64_t *out /* 2 + N */);

/* ---- DSMC: Chemistry/src/dsmc.jl ---------------------------------------------- */
/* dsmc(reactions) -> DirectSimulationMonteCarlo  dsmc.jl:143-166 with ONE DSMC.ElasticCollision (source, target kinetic
 * species, possibly the same one; rate = CrossSection over the relative speed g, cross_section.jl:3-14).  More than one
 * collision per object is ill defined in the reference (its cell lists accumulate across collisions, :94-99). */
int32_t iskb_dsmc_create(iskb_ctx *ctx, iskb_species *source, iskb_species *target, const double *g_nodes,
                         const double *sigma, int32_t n_nodes, uint64_t seed, iskb_dsmc **out);
/* PIC.perform!(dsmc, E, dt, config)  dsmc.jl:87-142.  nu_out: collisions per cell of this call (nx*ny, nullable);
 * n_candidates: sum over cells of floor(Nc) (:109-122), independent of the random stream. */
int32_t iskb_dsmc_perform(iskb_dsmc *dsmc, double dt, double *nu_out, int64_t *n_candidates, int64_t *n_collisions);

#ifdef __cplusplus
}
#endif
#endif /* ISKRA_B200_H */
