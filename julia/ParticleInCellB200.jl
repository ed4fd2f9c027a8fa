# ParticleInCellB200.jl -- the reference-side binding of libiskra_b200.so.
#
# What a maintainer of bchaber/iskra adds to run the per-timestep particle hot path on a B200:
# new types + methods on the reference's OWN generic functions (multiple dispatch is the
# reference's plugin mechanism, ParticleInCell.jl:44-45), no edits to reference sources.
# Every `ccall` below binds one entry point of include/iskra_b200.h.
#
# NOTE: Julia is not installed in the build image (no `julia` binary, no network), so this file
# has been syntax-reviewed only.  The executed mirror of exactly these calls is the Python/ctypes
# host code in iskra_b200/ (same entry points, same argument order), which the test-suite drives.
module ParticleInCellB200

import ParticleInCell
import ParticleInCell: KineticSpecies, FluidSpecies
import FiniteDifferenceMethod
import RegularGrids: CartesianGrid
import Chemistry
import Circuit

const LIB = get(ENV, "ISKRA_B200_LIB", "libiskra_b200.so")

struct IskraB200Error <: Exception
  code :: Int32
  msg :: String
end

function check(rc :: Int32)
  rc == 0 && return
  throw(IskraB200Error(rc, unsafe_string(ccall((:iskb_last_error, LIB), Cstring, ()))))
end

# ---- context = one GPU + one grid (replaces phi, rho, E, B = zeros(size(grid)...), ParticleInCell.jl:97-100)
mutable struct Context
  h :: Ptr{Cvoid}
  grid :: CartesianGrid{2}
end

function Context(grid :: CartesianGrid{2}; device = parse(Int, get(ENV, "LOCAL_RANK", "0")))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:iskb_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
  nx, ny = grid.n
  bc(s) = s == :periodic ? Int32(1) : Int32(0)
  (l, r), (b, t) = grid.bcs
  bcs = Int32[bc(l), bc(r), bc(b), bc(t)]
  check(ccall((:iskb_grid_set, LIB), Int32,
              (Ptr{Cvoid}, Int32, Int32, Float64, Float64, Float64, Float64, Ptr{Int32}),
              h[], nx, ny, grid.Δh[1], grid.Δh[2], grid.origin[1], grid.origin[2], bcs))
  ctx = Context(h[], grid)
  finalizer(c -> ccall((:iskb_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
  ctx
end

# ---- species: device-backed KineticSpecies{2,3} with lazily synced host mirrors --------------------
mutable struct B200Species
  host :: KineticSpecies{2,3}      # e.np, e.x, e.v ... keep working through this mirror
  h :: Ptr{Cvoid}
  ctx :: Context
  device_newer :: Bool
end

function B200Species(ctx :: Context, part :: KineticSpecies{2,3})
  N = size(part.x, 1)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:iskb_species_create, LIB), Int32,
              (Ptr{Cvoid}, Int64, Float64, Float64, Float64, Ref{Ptr{Cvoid}}),
              ctx.h, N, part.q, part.m, part.w0, h))
  sp = B200Species(part, h[], ctx, false)
  upload!(sp)
  sp
end

function upload!(sp :: B200Species)
  p = sp.host
  GC.@preserve p begin   # the library copies during the call and keeps no host pointer
    check(ccall((:iskb_species_upload, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}, Int64, Int64),
                sp.h, p.x, p.v, p.wg, p.id, p.np, size(p.x, 1)))
  end
  sp.device_newer = false
end

function download!(sp :: B200Species)
  p = sp.host
  np = Ref{Int64}(0)
  GC.@preserve p begin
    check(ccall((:iskb_species_download, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}, Int64),
                sp.h, p.x, p.v, p.wg, p.id, size(p.x, 1)))
  end
  check(ccall((:iskb_species_np, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), sp.h, np))
  p.np = np[]
  sp.device_newer = false
  p
end

Base.getproperty(sp :: B200Species, s :: Symbol) =
  s in (:host, :h, :ctx, :device_newer) ? getfield(sp, s) :
  (getfield(sp, :device_newer) && download!(sp); getproperty(getfield(sp, :host), s))

# ---- operators: methods on the reference's generic functions --------------------------------------
# grid_to_particle(grid, part, (i,j)->E[i,j,:])           cloud_in_cell.jl:20-36
function ParticleInCell.grid_to_particle(grid :: CartesianGrid{2}, sp :: B200Species, u)
  np = sp.np
  pu = zeros(np, 3)
  check(ccall((:iskb_gather, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), sp.h, pu))
  pu
end

# push_particles!(pusher, part, E, B, dt)                  pushers.jl:8-11
function ParticleInCell.push_particles!(:: ParticleInCell.BorisPusher{:xy}, sp :: B200Species, E, B, Δt)
  check(ccall((:iskb_push, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64), sp.h, E === nothing ? C_NULL : E, Δt))
  sp.device_newer = true
end

# wrap!(part, grid; dims) / discard!(part, grid; dims)     surfaces/wrap.jl:1-33
modes(dims, m) = (1 in dims ? Int32(m) : Int32(0), 2 in dims ? Int32(m) : Int32(0))
function ParticleInCell.wrap!(sp :: B200Species, grid; dims = 1:2)
  mx, my = modes(dims, 1)
  check(ccall((:iskb_boundary, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}), sp.h, mx, my, C_NULL))
  sp.device_newer = true
end
function ParticleInCell.discard!(sp :: B200Species, grid; dims = 1:2)
  mx, my = modes(dims, 2)
  removed = Ref{Int64}(0)
  check(ccall((:iskb_boundary, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ref{Int64}), sp.h, mx, my, removed))
  sp.device_newer = true
  removed[]
end

# density(species, grid)                                    kinetic.jl:53
function ParticleInCell.density(sp :: B200Species, grid :: CartesianGrid{2})
  n = zeros(grid.n...)
  check(ccall((:iskb_density, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), sp.h, n))
  n
end

# ---- field solve ------------------------------------------------------------------------------------
struct B200Poisson
  ctx :: Context
end
function B200Poisson(ctx :: Context, ε0 :: Float64)
  check(ccall((:iskb_poisson_create, LIB), Int32, (Ptr{Cvoid}, Float64), ctx.h, ε0))
  B200Poisson(ctx)
end
FiniteDifferenceMethod.apply_periodic(ps :: B200Poisson, axis) =
  check(ccall((:iskb_poisson_apply_periodic, LIB), Int32, (Ptr{Cvoid}, Int32), ps.ctx.h, axis))
function FiniteDifferenceMethod.apply_dirichlet(ps :: B200Poisson, nodes :: BitArray{2}, ϕ0)
  mask = UInt8.(nodes)                       # nx*ny bytes, column-major like the BitArray
  check(ccall((:iskb_poisson_apply_dirichlet, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Float64), ps.ctx.h, mask, ϕ0))
end
# phi = calculate_electric_potential(solver, -rho) ; E = calculate_electric_field(solver, phi)
function FiniteDifferenceMethod.calculate_electric_potential(ps :: B200Poisson, f)
  ρ = -f
  check(ccall((:iskb_fields_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ps.ctx.h, ρ, C_NULL, C_NULL))
  check(ccall((:iskb_field_solve, LIB), Int32, (Ptr{Cvoid},), ps.ctx.h))
  ϕ = zeros(size(f))
  check(ccall((:iskb_fields_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ps.ctx.h, C_NULL, ϕ, C_NULL))
  ϕ
end
function FiniteDifferenceMethod.calculate_electric_field(ps :: B200Poisson, ϕ)
  E = zeros(size(ϕ)..., 3)
  check(ccall((:iskb_fields_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ps.ctx.h, C_NULL, C_NULL, E))
  E
end

# ---- MCC: perform!(mcc, E, dt, config)                    Chemistry/src/mcc.jl:231-289 --------------
kind(::Chemistry.MCC.ElasticIsotropic) = (Int32(0), 0.0)
kind(::Chemistry.MCC.ElasticBackward) = (Int32(1), 0.0)
kind(::Chemistry.MCC.InelasticBackward) = (Int32(2), 0.0)
kind(t::Chemistry.MCC.Excitation) = (Int32(3), t.energy)
kind(t::Chemistry.MCC.Ionization) = (Int32(4), t.energy)

mutable struct B200MCC
  h :: Ptr{Cvoid}
  source :: B200Species
  products :: Vector{B200Species}
end

function B200MCC(ctx :: Context, mcc :: Chemistry.MonteCarloCollisions, lookup; seed = UInt64(0))
  cs = mcc.collisions
  src, tgt = lookup(first(cs).source), first(cs).target
  kinds = Int32[kind(c.type)[1] for c in cs]
  thr = Float64[kind(c.type)[2] for c in cs]
  lens = Int32[size(c.rate.nodes, 1) for c in cs]
  eps = vcat([c.rate.nodes[:, 1] for c in cs]...)
  sig = vcat([c.rate.nodes[:, 2] for c in cs]...)
  prods = Ptr{Cvoid}[C_NULL for _ in cs]
  used = B200Species[]
  for (k, c) in enumerate(cs), p in c.products
    if kinds[k] == 4 && p !== first(cs).source
      bp = lookup(p); prods[k] = bp.h; push!(used, bp)
    end
  end
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:iskb_mcc_create, LIB), Int32,
              (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Float64, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Float64},
               Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}, UInt64, Ref{Ptr{Cvoid}}),
              ctx.h, src.h, tgt.q, tgt.m, tgt.T, tgt.n, length(cs), kinds, thr, lens, eps, sig, prods, seed, h))
  B200MCC(h[], src, used)
end

function ParticleInCell.perform!(m :: B200MCC, E, Δt, config)
  nc, ncoll = Ref{Int64}(0), Ref{Int64}(0)
  check(ccall((:iskb_mcc_perform, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
              m.h, Δt, C_NULL, nc, ncoll))
  m.source.device_newer = true
  foreach(p -> p.device_newer = true, m.products)
end

# ---- surfaces, electrodes, circuit (SURVEY.md 8f row N1) ------------------------------------------------
# ParticleInCell/src/pic/surfaces/{build,track,check,hit}.jl and pic/circuit_coupling.jl.  The tracker's
# Dict{(cell,cell) => Surface} lives on the device as a per-cell face table; the reference's own Surface
# values stay the host-side handles (their kind decides hit!, electrodes also carry dq / area).
const SURF_PERIODIC, SURF_ABSORBING, SURF_REFLECTIVE, SURF_FIXED, SURF_FLOATING = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)

surface_kind(:: ParticleInCell.PeriodicSurface)            = SURF_PERIODIC
surface_kind(:: ParticleInCell.AbsorbingSurface)           = SURF_ABSORBING
surface_kind(:: ParticleInCell.ReflectiveSurface)          = SURF_REFLECTIVE
surface_kind(:: ParticleInCell.FixedPotentialElectrode)    = SURF_FIXED
surface_kind(:: ParticleInCell.FloatingPotentialElectrode) = SURF_FLOATING

mutable struct B200SurfaceTracker
  h :: Ptr{Cvoid}
  ctx :: Context
  ids :: IdDict{Any,Int32}         # Surface => surface id on the device
end

# create_surface_tracker(grid, ds)  build.jl:95-100
function B200SurfaceTracker(ctx :: Context, ds :: ParticleInCell.Surface = ParticleInCell.AbsorbingSurface())
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:iskb_tracker_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx.h, surface_kind(ds), h))
  B200SurfaceTracker(h[], ctx, IdDict{Any,Int32}())
end

# track_surface!(st, bcs::BitArray{2}, ss)  build.jl:109-111 ; sigma_dof: the electrode's dof from add_new_dof
function ParticleInCell.track_surface!(st :: B200SurfaceTracker, bcs :: BitArray{2}, ss :: ParticleInCell.Surface;
                                       sigma_dof :: Integer = 0)
  mask = UInt8.(bcs)
  sid = Ref{Int32}(0)
  area = hasproperty(ss, :area) ? Float64(ss.area) : 0.0
  GC.@preserve mask check(ccall((:iskb_tracker_track_surface, LIB), Int32,
                                (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Float64, Ref{Int32}),
                                st.h, mask, surface_kind(ss), Int32(sigma_dof), area, sid))
  st.ids[ss] = sid[]
end

# track!(st, part, dt)  track.jl:42-52  /  check!(st, part, dt)  check.jl:39-68
function ParticleInCell.track!(st :: B200SurfaceTracker, sp :: B200Species, Δt)
  upload!(sp)
  check(ccall((:iskb_tracker_track, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Int64}), st.h, sp.h, Δt, C_NULL))
end
function ParticleInCell.check!(st :: B200SurfaceTracker, sp :: B200Species, Δt)
  nabs, fast = Ref{Int64}(0), Ref{Int32}(0)
  check(ccall((:iskb_tracker_check, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ref{Int64}, Ref{Int32}),
              st.h, sp.h, Δt, nabs, fast))
  sp.device_newer = true
  fast[] != 0 && println("ERROR: $(sp.host) particle is too fast")     # check.jl:44-46
  nabs[]
end

# electrode.dq (circuit_coupling.jl:49-50) accumulates on the device
function collected_charge(st :: B200SurfaceTracker, ss; reset = false)
  dq = Ref{Float64}(0.0)
  check(ccall((:iskb_surface_charge, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}, Int32), st.h, st.ids[ss], dq, reset ? 1 : 0))
  dq[]
end

# add_new_dof / apply_neumann / get_rhs(ps, :σ, dof)  generalized_poisson.jl:217-269, 367-370
function FiniteDifferenceMethod.add_new_dof(ps :: B200Poisson, symbol :: Symbol)
  symbol == :σ || error("only σ dofs exist")
  dof = Ref{Int32}(0)
  check(ccall((:iskb_poisson_add_dof, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), ps.ctx.h, dof))
  Int(dof[])
end
function FiniteDifferenceMethod.apply_neumann(ps :: B200Poisson, nodes :: BitArray{2}, dof)
  mask = UInt8.(nodes)
  GC.@preserve mask check(ccall((:iskb_poisson_apply_neumann, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32), ps.ctx.h, mask, Int32(dof)))
end
sigma_rhs(ps :: B200Poisson, dof) = (v = Ref{Float64}(0.0);
  check(ccall((:iskb_poisson_sigma_get, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}), ps.ctx.h, Int32(dof), v)); v[])
sigma_rhs!(ps :: B200Poisson, dof, value) =
  check(ccall((:iskb_poisson_sigma_set, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), ps.ctx.h, Int32(dof), Float64(value)))
sigma_rhs_add!(ps :: B200Poisson, dof, delta) =
  check(ccall((:iskb_poisson_sigma_add, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), ps.ctx.h, Int32(dof), Float64(delta)))
phi_at(ps :: B200Poisson, i, j) = (v = Ref{Float64}(0.0);
  check(ccall((:iskb_phi_at, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ref{Float64}), ps.ctx.h, Int32(i), Int32(j), v)); v[])

# advance!(circuit::CircuitRLC, ϕ, Δt, config)  circuit_coupling.jl:33-43: the RLC recurrence (Circuit.jl:117-136)
# stays Julia; only `σ .+= dσ` crosses the boundary (8 bytes).
function advance_circuit!(circuit, ps :: B200Poisson, Δt)
  circuit === nothing && return 0.0
  Circuit.advance_circuit!(circuit, 0, Δt)
  dσ = ParticleInCell.foo!(circuit.ext, circuit.i, Δt)
  sigma_rhs_add!(ps, 1, dσ)
  dσ
end

# ---- axisymmetric r-z variant (SURVEY.md 8f row N3) -----------------------------------------------------
# AxialGrid{2}: ring volumes and the operator come from the reference's own code (RegularGrids.cell_volume,
# FiniteDifferenceMethod.create_poisson_solver(::AxialGrid{2}, eps0)) and are handed to the device as arrays.
function upload_cell_volume!(ctx :: Context, V :: Matrix{Float64})
  GC.@preserve V check(ccall((:iskb_cell_volume_set, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, V))
end
function B200Poisson(ctx :: Context, reference_solver :: FiniteDifferenceMethod.PoissonSolver{:rz, 2})
  ps = B200Poisson(ctx, reference_solver.ε0)
  A = Matrix{Float64}(reference_solver.A)
  GC.@preserve A check(ccall((:iskb_poisson_set_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx.h, A, size(A, 1)))
  ps
end
function ParticleInCell.push_particles!(:: ParticleInCell.BorisPusher{:rz}, sp :: B200Species, E, B, Δt)
  check(ccall((:iskb_set_pusher, LIB), Int32, (Ptr{Cvoid}, Int32), sp.ctx.h, Int32(1)))
  upload!(sp)
  check(ccall((:iskb_push, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64), sp.h, C_NULL, Δt))
  sp.device_newer = true
end

# ---- DSMC (SURVEY.md 8f row N4): Chemistry/src/dsmc.jl ----------------------------------------------------
mutable struct B200DSMC
  h :: Ptr{Cvoid}
  source :: B200Species
  target :: B200Species
end

# dsmc(@reactions ...) with ONE DSMC.ElasticCollision; `lookup` maps the reference species to their device twins
function B200DSMC(ctx :: Context, d :: Chemistry.DirectSimulationMonteCarlo, lookup; seed = UInt64(0))
  length(d.collisions) == 1 || error("one collision per DSMC object (the reference's cell lists accumulate, dsmc.jl:94-99)")
  c = d.collisions[1]
  s, t = lookup(c.source), lookup(c.target)
  upload!(s); upload!(t)
  gn, sg = Vector{Float64}(c.rate.nodes[:, 1]), Vector{Float64}(c.rate.nodes[:, 2])
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve gn sg check(ccall((:iskb_dsmc_create, LIB), Int32,
                                 (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, UInt64, Ref{Ptr{Cvoid}}),
                                 ctx.h, s.h, t.h, gn, sg, Int32(length(gn)), seed, h))
  B200DSMC(h[], s, t)
end

function ParticleInCell.perform!(d :: B200DSMC, E, Δt, config)
  nc, ncoll = Ref{Int64}(0), Ref{Int64}(0)
  check(ccall((:iskb_dsmc_perform, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{Int64}, Ref{Int64}), d.h, Δt, C_NULL, nc, ncoll))
  d.source.device_newer = true
  d.target.device_newer = true
end

# ---- fused loop: drop-in for ParticleInCell.solve that still fires the hooks (ParticleInCell.jl:84-139)
# `interactions` :: B200MCC / B200DSMC objects in config.interactions order; species and interactions that exist on the
# context but are not listed (a source buffer for add!, a scratch species) are neither advanced nor deposited.
function solve(ctx :: Context, species :: Vector{B200Species}, Δt, timesteps; interactions = [], after_push = (1, 1),
               sort_interval = 4, miss_threshold = 5e-4, max_interval = 4, full_interval = 0, lean = true)
  check(ccall((:iskb_set_after_push, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), ctx.h, after_push...))
  check(ccall((:iskb_set_sort_interval, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, sort_interval))
  check(ccall((:iskb_set_sort_policy, LIB), Int32, (Ptr{Cvoid}, Float64, Int32), ctx.h, miss_threshold, max_interval))
  check(ccall((:iskb_set_sort_full_interval, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, full_interval))
  check(ccall((:iskb_set_lean, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, lean ? 1 : 0))
  check(ccall((:iskb_set_advance_path, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, 0))
  foreach(upload!, species)
  sh = Ptr{Cvoid}[sp.h for sp in species]
  ih = Ptr{Cvoid}[i.h for i in interactions]
  GC.@preserve sh ih check(ccall((:iskb_step_set_active, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Ptr{Cvoid}}, Int32),
                                 ctx.h, sh, Int32(length(sh)), ih, Int32(length(ih))))
  ParticleInCell.enter_loop()
  for it in 1:timesteps
    foreach(upload!, species)                          # host edits made in after_loop reach the device
    check(ccall((:iskb_step, LIB), Int32, (Ptr{Cvoid}, Float64, Int32), ctx.h, float(Δt), 1))
    foreach(sp -> sp.device_newer = true, species)
    ParticleInCell.after_loop(it, it*Δt - Δt, Δt)     # scripts' iteration(): diagnostics, RF apply_dirichlet
  end
  check(ccall((:iskb_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx.h))
  ParticleInCell.exit_loop()
end

# ---- secondary-electron emission at a wall: Chemistry.emit! (see.jl:114-181) for device-backed species.  The coefficient
# closures of see.jl cannot cross the C ABI; their parameters do (the module defaults of see.jl:60-63 shown here).
const SEE_EDGE = Dict(:left => 0, :right => 1, :bottom => 2, :top => 3)
function Chemistry.emit!(primary :: B200Species, secondary :: B200Species, grid, material :: Symbol; boundary = :all,
                         vaughan = (13., 500., 3., 1.), elastic = (2., 10., 0.55, 13., 0.03), inelastic = 0.07,
                         secondary_r = (0.03, 0.07), seed = UInt64(1))
  haskey(SEE_EDGE, boundary) || error("emit! with boundary = :all: the wall normal is the zero vector (see.jl:94)")
  upload!(primary); upload!(secondary)
  coef = Float64[vaughan..., elastic..., inelastic, secondary_r...]
  counts = zeros(Int64, 4)
  GC.@preserve coef counts check(ccall((:iskb_see_emit, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, UInt64, Ptr{Int64}),
                                       primary.h, secondary.h, Int32(SEE_EDGE[boundary]), coef, seed, counts))
  primary.device_newer = true
  secondary.device_newer = true
  (elastic = counts[1], inelastic = counts[2], secondaries = counts[3], absorbed = counts[4])
end

# bookkeeping of the row order (new relative to the reference; diagnostics only)
function sort_stats(sp :: B200Species)
  out = zeros(Int64, 8)
  check(ccall((:iskb_species_sort_stats, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), sp.h, out))
  (full_sorts = out[1], regroups = out[2], since_full = out[3], since_regroup = out[4], slots = out[5], dead = out[6], in_directory = out[7])
end
function context_counts(ctx :: Context)
  a, b, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
  check(ccall((:iskb_ctx_counts, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), ctx.h, a, b, c))
  (species = a[], mcc = b[], dsmc = c[])
end

end # module
