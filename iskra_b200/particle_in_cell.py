"""Host mirror of the ParticleInCell module API that problem/*.jl scripts call
(ParticleInCell/src/ParticleInCell.jl and pic/*.jl), backed by the device library.

Julia's `f!` functions are spelled `f_` here.  KineticSpecies keeps lazily-synced host mirrors so
that script idioms like `e.np = 0`, `iHe.x .= e.x` (problem/10_two_streams.jl:62-68) keep working:
reading `.x/.v/.wg/.id` pulls from the device if it is newer and marks the host copy as the
authority (the caller may write through the returned array) until the next device operation
pushes it back; `.x_ro/.v_ro/.wg_ro/.id_ro` are read-only views that leave the device copy the
authority (diagnostics use those).  Every device operation -- the fused loop included -- pushes
first, so a host copy that was only read is never stale when it goes back.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from .units_and_constants import kB


# ---- species ------------------------------------------------------------------------------------
class KineticSpecies:
    """KineticSpecies{2,3}  pic/kinetic.jl:1-18."""

    def __init__(self, name, N, D=2, V=3):
        if (D, V) != (2, 3):
            raise NotImplementedError("only KineticSpecies{2,3} is on the hot path (SURVEY.md 8a)")
        self.name = name
        self.N = int(N)
        self._x = np.zeros((N, 2), order="F")
        self._v = np.zeros((N, 3), order="F")
        self._wg = np.ones(N)
        self._id = np.arange(1, N + 1, dtype=np.uint32)      # particle_uuids, kinetic.jl:15
        self._np = 0
        self.n = np.zeros((0, 0))
        self.m = 0.0
        self.q = 0.0
        self.w0 = 1.0
        self._rt = None
        self._h = None
        self._host_newer = True
        self._dev_newer = False

    def __repr__(self):
        return self.name

    # -- lazy sync -----------------------------------------------------------------------------
    def _bind(self, grid):
        rt = grid._rt
        if self._rt is rt:
            return
        if self._rt is not None:
            raise RuntimeError("species %s is already bound to another grid" % self.name)
        h = L.vp()
        L.check(rt.lib.iskb_species_create(rt.h, self.N, self.q, self.m, self.w0, C.byref(h)))
        self._rt, self._h = rt, h
        self._host_newer = True

    def _push(self, grid=None):
        """make the device copy current"""
        if grid is not None:
            self._bind(grid)
        if self._rt is None:
            raise RuntimeError("species %s is not bound to a grid yet" % self.name)
        if self._host_newer:
            L.check(self._rt.lib.iskb_species_upload(self._h, L.ptr(self._x), L.ptr(self._v), L.ptr(self._wg),
                                                     L.ptr(self._id), self._np, self.N))
            self._host_newer = False
            self._dev_newer = False

    def _pull(self):
        """make the host copy current"""
        if self._dev_newer:
            L.check(self._rt.lib.iskb_species_download(self._h, L.ptr(self._x), L.ptr(self._v), L.ptr(self._wg),
                                                       L.ptr(self._id), self.N))
            self._np = self._query_np()
            self._dev_newer = False

    def _query_np(self):
        v = L.i64()
        L.check(self._rt.lib.iskb_species_np(self._h, C.byref(v)))
        return v.value

    def window_stats(self):
        """(gather misses, deposit misses, window moves, deposit rounds) of the fused kernel since the last call."""
        out = np.zeros(4, dtype=np.int64)
        L.check(self._rt.lib.iskb_species_window_stats(self._h, L.ptr(out)))
        return out

    def _touched_on_device(self):
        """a device operation changed the rows: the host copy (pushed just before) is no longer current"""
        self._dev_newer = True
        self._host_newer = False

    @property
    def np(self):
        if self._dev_newer:
            return self._query_np()
        return self._np

    @np.setter
    def np(self, value):
        self._pull()
        self._np = int(value)
        self._host_newer = True

    def _host(name):
        def get(self):
            self._pull()
            self._host_newer = True     # the caller may write through the returned array
            return getattr(self, name)

        def set_(self, value):
            self._pull()
            getattr(self, name)[...] = value
            self._host_newer = True
        return property(get, set_)

    def _host_ro(name):
        def get(self):
            self._pull()
            a = getattr(self, name).view()
            a.flags.writeable = False
            return a
        return property(get)

    x = _host("_x")
    v = _host("_v")
    wg = _host("_wg")
    id = _host("_id")
    x_ro = _host_ro("_x")
    v_ro = _host_ro("_v")
    wg_ro = _host_ro("_wg")
    id_ro = _host_ro("_id")


class FluidSpecies:
    """FluidSpecies  pic/fluid.jl:1-11 -- on the path only as the read-only MCC target."""

    def __init__(self, name, mu, q, m, n, T):
        self.name, self.mu, self.q, self.m, self.T = name, float(mu), float(q), float(m), float(T)
        self.n = np.asfortranarray(n, dtype=np.float64)

    def __repr__(self):
        return self.name


def is_fluid(species):
    return isinstance(species, FluidSpecies)


def create_kinetic_species(name, N, q, m, weight, D=2, V=3):
    """problem/configuration.jl:95-102"""
    sp = KineticSpecies(name, N, D, V)
    sp.q, sp.m = float(q), float(m)
    sp._wg *= weight
    sp.w0 = float(weight)
    return sp


def remove_(sp, i):
    """remove!(sp, i)  kinetic.jl:20-27 ; i is 1-based"""
    sp._push()
    L.check(sp._rt.lib.iskb_species_remove(sp._h, int(i)))
    sp._touched_on_device()


def add_(src, dst):
    """add!(src, dst)  kinetic.jl:29-37"""
    src._push()
    dst._push()
    L.check(dst._rt.lib.iskb_species_add(src._h, dst._h))
    dst._touched_on_device()


def remove_particles_(part, grid, matches):
    """remove_particles!(part, dh, matches)  kinetic.jl:39-50.  `matches(i, j)` (1-based cell) is evaluated on
    the host once per cell and shipped as a mask; returns the number of particles removed."""
    part._push(grid)
    nx, ny = grid.n
    mask = np.zeros((nx, ny), dtype=np.uint8, order="F")
    for j in range(1, ny + 1):
        for i in range(1, nx + 1):
            mask[i - 1, j - 1] = 1 if matches(i, j) else 0
    n = L.i64()
    L.check(part._rt.lib.iskb_species_remove_in_cells(part._h, L.ptr(mask), C.byref(n)))
    part._touched_on_device()
    return n.value


# ---- sources ------------------------------------------------------------------------------------
class MaxwellianSource:
    """MaxwellianSource{D,V}  pic/sources.jl:8-22"""

    def __init__(self, rate, wx, wv, dx=None, dv=None):
        self.rate = float(rate)
        self.wx = np.asarray(wx, dtype=np.float64).reshape(-1)
        self.wv = np.asarray(wv, dtype=np.float64).reshape(-1)
        self.dx = np.zeros_like(self.wx) if dx is None else np.asarray(dx, dtype=np.float64).reshape(-1)
        self.dv = np.zeros_like(self.wv) if dv is None else np.asarray(dv, dtype=np.float64).reshape(-1)
        self.seed = 0
        # solve() samples `src.species` for every src in config.sources (ParticleInCell.jl:104-106); the reference's
        # struct has no such field (sources.jl:8-22), so a script that fills config.sources must attach it


def thermal_speed(T, m):
    """problem/configuration.jl:79-81"""
    return math.sqrt(2 * kB * T / m)


class DensitySource:
    """DensitySource{D}  pic/sources.jl:3-6"""

    def __init__(self, delta, grid):
        self.delta, self.grid = np.asarray(delta, dtype=np.float64), grid


def create_thermalized_beam(species, x, vb, dx=None, T=300.0, rate=1.0):
    """problem/configuration.jl:89-93"""
    vth = thermal_speed(T, species.m) * np.ones(3)
    return MaxwellianSource(rate, x, vth, dx=dx, dv=vb)


_sample_calls = [0]


def sample_(src, species, dt, grid=None, seed=None):
    """sample!(src, species, dt)  pic/sources.jl:24-34 -- drawn on the device (Philox); a DensitySource adds its
    density to a fluid species on the host (:36-38)."""
    if isinstance(src, DensitySource):
        if not is_fluid(species):
            raise TypeError("DensitySource feeds a FluidSpecies")
        species.n += src.delta
        return
    n = int(math.floor(src.rate * dt))            # :30
    species._push(grid)
    if seed is None:
        _sample_calls[0] += 1
        seed = (src.seed << 20) + _sample_calls[0]
    L.check(species._rt.lib.iskb_species_sample_maxwellian(species._h, n, L.ptr(src.wx), L.ptr(src.dx),
                                                           L.ptr(src.wv), L.ptr(src.dv), int(seed)))
    species._touched_on_device()


def init(src, species, dt, grid=None):
    """init(src, species, dt)  ParticleInCell.jl:47-49"""
    sample_(src, species, dt, grid)


# ---- operators ----------------------------------------------------------------------------------
def particle_cell(part, grid):
    """particle_cell for every live particle  ParticleInCell.jl:28-35 -> (i, j, hx, hy), 1-based."""
    part._push(grid)
    n = part.np
    i = np.zeros(n, dtype=np.int32)
    j = np.zeros(n, dtype=np.int32)
    hx = np.zeros(n)
    hy = np.zeros(n)
    L.check(part._rt.lib.iskb_cell_index(part._h, L.ptr(i), L.ptr(j), L.ptr(hx), L.ptr(hy)))
    return i, j, hx, hy


def grid_to_particle(grid, part, u=None):
    """grid_to_particle(grid, part, (i,j)->E[i,j,:])  cloud_in_cell.jl:20-36.  `u`: (nx,ny,3) node
    array to gather, or None for the field currently on the device."""
    part._push(grid)
    if u is not None:
        grid._rt.set_fields(E=u)
    n = part.np
    out = np.zeros((n, 3), order="F")
    if n:
        L.check(part._rt.lib.iskb_gather(part._h, L.ptr(out)))
    return out


def particle_to_grid(part, grid):
    """particle_to_grid(part, grid, p->wg[p])  cloud_in_cell.jl:1-18"""
    return density(part, grid) * __import__("iskra_b200").regular_grids.cell_volume(grid)


def density(species, grid):
    """density(species, grid)  kinetic.jl:53 / fluid.jl:11"""
    if is_fluid(species):
        return species.n
    species._push(grid)
    n = np.zeros(grid.n, order="F")
    L.check(species._rt.lib.iskb_density(species._h, L.ptr(n)))
    species.n = n
    return n


class BorisPusher:
    """BorisPusher{T}  pushers.jl:4-6 ; kind "xy" or "rz" """

    def __init__(self, kind="xy"):
        self.kind = kind


def create_boris_pusher():
    return BorisPusher("xy")


def create_axial_boris_pusher():
    """create_axial_boris_pusher()  pushers.jl:6"""
    return BorisPusher("rz")


def _set_pusher(rt, pusher):
    kind = L.PUSHER_RZ if (pusher is not None and getattr(pusher, "kind", "xy") == "rz") else L.PUSHER_XY
    L.check(rt.lib.iskb_set_pusher(rt.h, kind))


def transform_from_cartesian_to_cylindrical_(part, dt):
    """transform_from_cartesian_to_cylindrical!(part, dt)  pushers.jl:52-66"""
    part._push()
    L.check(part._rt.lib.iskb_transform_cylindrical(part._h, float(dt)))
    part._touched_on_device()


def push_particles_(pusher, part, E, B, dt, grid=None):
    """push_particles!(pusher, part, E, B, dt)  pushers.jl:8-11.  E: (np,3) per-particle field, or
    None to gather from the device field on the fly.  B must be zero (generalized_poisson.jl:412-419)."""
    if B is not None and np.any(np.asarray(B) != 0):
        raise NotImplementedError("B != 0 is outside the hot path (calculate_magnetic_field == 0)")
    part._push(grid)
    _set_pusher(part._rt, pusher)
    pe = None if E is None else np.asfortranarray(E, dtype=np.float64)
    L.check(part._rt.lib.iskb_push(part._h, L.ptr(pe), float(dt)))
    part._touched_on_device()


def _modes(dims, mode):
    dims = (1, 2) if dims is None else tuple(dims)
    return (mode if 1 in dims else L.BND_NONE, mode if 2 in dims else L.BND_NONE)


def wrap_(part, grid, dims=None):
    """wrap!(part, grid; dims)  surfaces/wrap.jl:20-33"""
    part._push(grid)
    mx, my = _modes(dims, L.BND_WRAP)
    L.check(part._rt.lib.iskb_boundary(part._h, mx, my, None))
    part._touched_on_device()


def discard_(part, grid, dims=None):
    """discard!(part, grid; dims)  surfaces/wrap.jl:1-18 -> number removed"""
    part._push(grid)
    mx, my = _modes(dims, L.BND_DISCARD)
    removed = L.i64()
    L.check(part._rt.lib.iskb_boundary(part._h, mx, my, C.byref(removed)))
    part._touched_on_device()
    return removed.value


def sort_by_cell_(part, grid, for_deposit=False):
    """New (no reference counterpart): stable sort of the rows by cell; returns the permutation.
    for_deposit=True additionally interleaves the rows of every 8x8 tile round-robin over its cells
    (the layout the fused step keeps)."""
    part._push(grid)
    perm = np.zeros(part.np, dtype=np.uint32)
    fn = part._rt.lib.iskb_sort_for_deposit if for_deposit else part._rt.lib.iskb_sort_by_cell
    L.check(fn(part._h, L.ptr(perm)))
    part._touched_on_device()
    return perm


# ---- surfaces: pic/surfaces/{build,track,check,hit}.jl, pic/circuit_coupling.jl ----
class Surface:
    """abstract type Surface  build.jl:8"""
    kind = None
    _sid = None


class PeriodicSurface(Surface):
    kind = L.SURF_PERIODIC


class AbsorbingSurface(Surface):
    kind = L.SURF_ABSORBING


class ReflectiveSurface(Surface):
    kind = L.SURF_REFLECTIVE


def create_periodic_surface():
    return PeriodicSurface()


def create_absorbing_surface():
    return AbsorbingSurface()


def create_reflective_surface():
    return ReflectiveSurface()


class FixedPotentialElectrode(Surface):
    """FixedPotentialElectrode  circuit_coupling.jl:5-9 (phi, dq, area)"""
    kind = L.SURF_ELECTRODE_FIXED

    def __init__(self, phi, dq, area):
        self.phi, self._dq, self.area = phi, float(dq), float(area)
        self._st = None

    @property
    def dq(self):
        return self._dq


class FloatingPotentialElectrode(Surface):
    """FloatingPotentialElectrode  circuit_coupling.jl:11-16: fields (sigma, phi, dq, area) IN THIS ORDER.
    problem/configuration.jl:70 passes (phi0, sigma0, ...), so `.sigma` is the potential of the reference
    node and `.phi` the sigma right-hand side (quirk S1); kept as is."""
    kind = L.SURF_ELECTRODE_FLOATING

    def __init__(self, sigma, phi, dq, area):
        self.sigma, self.phi, self._dq0, self.area = sigma, phi, float(dq), float(area)
        self._st = None
        self._dof = 0

    @property
    def dq(self):
        """s.dq: charge collected on the device since the last reset"""
        if self._st is None or self._sid is None:
            return self._dq0
        v = L.f64()
        L.check(self._st._rt.lib.iskb_surface_charge(self._st._h, self._sid, C.byref(v), 0))
        return self._dq0 + v.value

    @dq.setter
    def dq(self, value):
        self._dq0 = float(value)
        if self._st is not None and self._sid is not None:
            L.check(self._st._rt.lib.iskb_surface_charge(self._st._h, self._sid, None, 1))


class SurfaceTracker:
    """SurfaceTracker{2}  build.jl:13-18 -- the Dict lives on the device as a per-cell face table."""

    def __init__(self, grid, ds=None):
        ds = AbsorbingSurface() if ds is None else ds
        self._rt = grid._rt
        self.dh = grid.dh[0]                                   # build.jl:96-97
        h = L.vp()
        L.check(self._rt.lib.iskb_tracker_create(self._rt.h, ds.kind, C.byref(h)))
        self._h = h
        self.surfaces = [ds]

    def get(self, bc, default=None):
        """get(st, ((i,j),(k,l)), nothing)  build.jl:113-117 -> surface kind or default"""
        (i, j), (k, l) = bc
        v = L.i32()
        L.check(self._rt.lib.iskb_tracker_lookup(self._h, i, j, k, l, C.byref(v)))
        return default if v.value < 0 else v.value

    def route_hits_to_sigma(self, on=True):
        L.check(self._rt.lib.iskb_tracker_route_hits_to_sigma(self._h, 1 if on else 0))


def create_surface_tracker(grid, ds=None):
    """create_surface_tracker(grid::CartesianGrid{2}, ds=AbsorbingSurface())  build.jl:95-100"""
    return SurfaceTracker(grid, ds)


def track_surface_(st, bcs, ss):
    """track_surface!(st, bcs::BitArray{2}, ss)  build.jl:109-111"""
    mask = np.asfortranarray(np.asarray(bcs, dtype=bool).astype(np.uint8))
    sid = L.i32()
    dof = getattr(ss, "_dof", 0) if ss.kind == L.SURF_ELECTRODE_FLOATING else 0
    area = getattr(ss, "area", 0.0)
    L.check(st._rt.lib.iskb_tracker_track_surface(st._h, L.ptr(mask), ss.kind, dof, area, C.byref(sid)))
    ss._sid, ss._st = sid.value, st
    st.surfaces.append(ss)


def track_(st, part, dt, grid=None):
    """track!(st, part, dt)  track.jl:41-52 -> number of tracked particles"""
    if st is None:
        return 0
    part._push(grid)
    n = L.i64()
    L.check(st._rt.lib.iskb_tracker_track(st._h, part._h, float(dt), C.byref(n)))
    return n.value


def check_(st, part, dt):
    """check!(st, part, dt)  check.jl:38-68 -> (too_fast, n_absorbed)"""
    if st is None:
        return False, 0
    n, tf = L.i64(), L.i32()
    L.check(st._rt.lib.iskb_tracker_check(st._h, part._h, float(dt), C.byref(n), C.byref(tf)))
    part._touched_on_device()
    if tf.value:
        print("ERROR: %s particle is too fast" % part)          # check.jl:44-46
    return bool(tf.value), n.value


class PlasmaDevice:
    """PlasmaDevice <: Circuit.CircuitDevice  circuit_coupling.jl:18-25"""

    def __init__(self, positive, negative):
        self.positive, self.negative = positive, negative

    def voltage(self):
        return self.positive.phi.value - self.negative.phi.value   # :23-25


def advance_circuit_coupling_(circuit, phi, dt, config):
    """advance!(circuit, phi, dt, config)  circuit_coupling.jl:33-43 (foo! :26-32 inlined)"""
    if circuit is None:
        return 0.0
    from . import circuit as CIR
    from . import finite_difference_method as FDM
    CIR.advance_circuit_(circuit, 0, dt)
    if isinstance(circuit.ext, PlasmaDevice):
        circuit.ext.positive.dq = 0.0
        dsig = -dt * circuit.i / circuit.ext.positive.area
    else:
        dsig = 0.0
    FDM.get_rhs(config.solver, "sigma", 1).add(dsig)            # :41-42
    return dsig


# ---- hooks and loop (ParticleInCell.jl:37-45, 51-72, 84-139) --------------------------------------
class Hooks:
    def __init__(self):
        self.enter_loop = lambda: None
        self.after_loop = lambda it, t, dt: None
        self.exit_loop = lambda: None
        self.after_push = lambda part, grid: wrap_(part, grid)   # default, ParticleInCell.jl:41


hooks = Hooks()


def advance_(part, E, B, dt, config):
    """advance!(part::KineticSpecies, E, B, dt, config)  ParticleInCell.jl:51-72 (operator by operator)."""
    grid = config.grid
    if E is not None:
        grid._rt.set_fields(E=E)
    track_(config.tracker, part, dt, grid)                          # :56
    push_particles_(config.pusher, part, None, None, dt, grid)     # gather :57 + push :59 fused
    check_(config.tracker, part, dt)                                # :60
    hooks.after_push(part, grid)                                    # :61


def perform_(interaction, E, dt, config):
    """extension point perform!(interaction, E, dt, config)  ParticleInCell.jl:44"""
    return interaction.perform_(E, dt, config)


def _check_active_lists(rt, config, kinetic):
    """iskb_step advances everything registered on the context; that must be exactly config's lists, in order."""
    handles = (L.vp * max(1, len(kinetic)))(*[s._h for s in kinetic])
    inter = [i._h for i in config.interactions]
    ih = (L.vp * max(1, len(inter)))(*inter)
    L.check(rt.lib.iskb_step_set_active(rt.h, handles, len(kinetic), ih, len(inter)))


def solve(config, dt=1e-5, timesteps=200, after_push=None, sort_interval=0, fused=True):
    """solve(config, dt, timesteps)  ParticleInCell.jl:84-139.

    fused=True runs the loop body as one device call per step (iskb_step) and still fires
    after_loop every step, so scripts' iteration() (diagnostics, RF Dirichlet update) keep working.
    `after_push` = (mode_x, mode_y) boundary modes replacing the after_push hook on that path."""
    grid, rt = config.grid, config.grid._rt
    dt = float(dt)
    hooks.enter_loop()                                                        # :95
    kinetic = [s for s in config.species if not is_fluid(s)]
    for s in kinetic:
        s._push(grid)
    for inter in config.interactions:
        inter._bind(config)
    from . import diagnostics
    diagnostics.register_solve_records(config)      # @field / @particle of :63-71,:122-133 -- fetched on demand only
    _set_pusher(rt, config.pusher)
    if fused:
        mx, my = after_push if after_push is not None else (L.BND_WRAP, L.BND_WRAP)
        rt.set_after_push(mx, my)
        rt.set_sort_interval(sort_interval)
    for s in config.species:
        # advance!(::FluidSpecies) (:74-82) and the fluid term of rho (:119-124) are not on the path (DESIGN.md
        # "out of scope"): a fluid that would need either must not be dropped silently
        if is_fluid(s) and (s.q != 0.0 or s.mu != 0.0 and getattr(s, "mobile", False)):
            raise NotImplementedError("FluidSpecies %s carries charge: fluid advection / fluid charge density are out of scope" % s)
    saved_hook = None
    if not fused and after_push is not None:
        # operator-by-operator loop: the boundary modes become the after_push hook (discards first, then wraps,
        # like the scripts' overrides: 11_rf_discharge.jl:80-83) for the duration of this call
        mxy = tuple(after_push)
        saved_hook = hooks.after_push

        def _modes_hook(part, g_):
            dd = [d + 1 for d in (0, 1) if mxy[d] == L.BND_DISCARD]
            ww = [d + 1 for d in (0, 1) if mxy[d] == L.BND_WRAP]
            if dd:
                discard_(part, g_, dims=dd)
            if ww:
                wrap_(part, g_, dims=ww)
        hooks.after_push = _modes_hook
    if fused:
        # the device step runs every species / interaction bound to the context, in creation order
        # (iskb_step); the reference runs config.species / config.interactions (:109-115).  Refuse to diverge.
        _check_active_lists(rt, config, kinetic)
    for it in range(1, timesteps + 1):
        if fused and config.circuit is not None:
            raise NotImplementedError("the circuit advances between advance! and density (:116); use fused=False")
        for src in config.sources:                                            # :104-106  sample!(src, src.species, dt)
            sample_(src, src.species, dt, grid)      # AttributeError without .species, like the reference's field access
        if fused:
            for s in kinetic:
                s._push(grid)      # host edits made in after_loop (e.np = 0, e.x[...] = ...) reach the device
            rt.step(dt, 1)
            for s in kinetic:
                s._touched_on_device()
        else:
            for inter in config.interactions:                                 # :109-111
                inter.perform_(None, dt, config)
            for part in kinetic:                                              # :113-115
                advance_(part, None, None, dt, config)
            advance_circuit_coupling_(config.circuit, None, dt, config)       # :116
            L.check(rt.lib.iskb_rho_zero(rt.h))                               # :118
            for part in kinetic:                                              # :119-124
                part._push(grid)
                L.check(rt.lib.iskb_density(part._h, None))
                L.check(rt.lib.iskb_rho_accumulate(rt.h, part._h))
            L.check(rt.lib.iskb_rho_allreduce(rt.h))
            L.check(rt.lib.iskb_field_solve(rt.h))                            # :126-128
        hooks.after_loop(it, it * dt - dt, dt)                                # :134
    rt.synchronize()
    if saved_hook is not None:
        hooks.after_push = saved_hook
    hooks.exit_loop()                                                         # :137
