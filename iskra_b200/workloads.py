"""Synthetic BASELINE workloads built through the host mirror.

  c5 : 2-D XY RF discharge + MCC  (BASELINE.json configs[4]) -- 2049x2049 nodes, dh and dt of
       problem/11_rf_discharge.jl, Dirichlet electrodes at i=1 (450 sin(2 pi 13.56 MHz t)) and
       i=nx (0 V), "periodic" in j, discard! dim 1 / wrap! dim 2, e- at 30 000 K + He+ at 300 K
       loaded uniformly, He background 9.64e20 m^-3 at 300 K, 4 + 2 MCC processes (synthetic
       tables, datasets.py).  1e9 particles over 8 GPUs = 1.25e8 per GPU (index-slice sharding).
  walls : SURVEY.md 8f row N1 at scale -- a bounded RF cell: 1025x1025 nodes, the RF-driven electrode
       (i = 1) and the grounded one (i = nx) registered as FixedPotentialElectrode surfaces, default
       AbsorbingSurface on the other two faces, a ReflectiveSurface block inside, e- + He+ loaded
       uniformly outside the block, no MCC.  advance! runs with config.tracker (track! / check!).
  seed  : SURVEY.md 8f row N3 at scale -- problem/13_seed.jl geometry (axial r-z grid, plates at z = 0 / Lz, axial
       Boris pusher, discard! dim 2) with a plasma column instead of a single seed electron: 33 x 65 nodes (or 65 x 125; the
       operator of an AxialGrid is dense: <= 8192 nodes), e- + Ar+ loaded in r < R/2, no MCC.
  c4 : 2-D XY two-stream (configs[3]) -- 1025x1025 nodes, dh and CFL of problem/10_two_streams.jl,
       fully "periodic", two +-1e7 m/s electron beams at 300 K + co-located ions, wrap! both axes.
"""
import math

import numpy as np

from . import _lib as L
from . import chemistry as CH
from . import datasets
from . import finite_difference_method as FDM
from . import particle_in_cell as PIC
from . import regular_grids as RG
from .configuration import Config
from .units_and_constants import c0, eps0, me, mp, qe


class Workload:
    def __init__(self, name, config, dt, after_push, rf=None, meta=None):
        self.name, self.config, self.dt, self.after_push, self.rf = name, config, dt, after_push, rf
        self.meta = meta or {}
        self.step_index = 0

    @property
    def rt(self):
        return self.config.grid._rt

    def kinetic(self):
        return [s for s in self.config.species if not PIC.is_fluid(s)]

    def n_particles(self):
        return sum(s.np for s in self.kinetic())

    def prepare(self, sort_interval, miss_threshold=0.0, max_interval=0, full_interval=0):
        rt = self.rt
        rt.set_sort_policy(miss_threshold, max_interval, full_interval)
        for s in self.kinetic():
            s._push(self.config.grid)
        for m in self.config.interactions:
            m._bind(self.config)
        rt.set_after_push(*self.after_push)
        rt.set_sort_interval(sort_interval)

    def step(self, n=1):
        """n iterations of the loop body; the RF electrode value is refreshed before every step the
        way problem/11_rf_discharge.jl:95 does it from after_loop (host -> device, 8 bytes)."""
        rt = self.rt
        for _ in range(n):
            if self.rf is not None:
                edge, amp, freq = self.rf
                t = (self.step_index - 1) * self.dt
                v = amp * math.sin(2 * math.pi * freq * t) if self.step_index > 0 else 0.0
                L.check(rt.lib.iskb_poisson_apply_dirichlet_edge(rt.h, edge, v))
            rt.step(self.dt, 1)
            self.step_index += 1
        rt.join()   # the last field solve runs on the field stream; order it before whatever follows
        for s in self.kinetic():
            s._touched_on_device()


def _load_uniform(species, grid, n, T, drift, seed):
    nx, ny = grid.n
    dx, dy = grid.dh
    src = PIC.create_thermalized_beam(species, [(nx - 1) * dx, (ny - 1) * dy], drift, T=T, rate=1.0)
    species._push(grid)
    L.check(species._rt.lib.iskb_species_sample_maxwellian(species._h, int(n), L.ptr(src.wx), L.ptr(src.dx),
                                                           L.ptr(src.wv), L.ptr(src.dv), int(seed)))
    species._touched_on_device()


def build_c5(particles_per_gpu=125_000_000, cells=2048, n_gpus_total=8, seed=2, capacity_factor=1.1, device=None):
    """C5 shard for one rank: particles_per_gpu rows (half e-, half He+)."""
    dh = 6.7 * 0.01 / 128                         # 11_rf_discharge.jl:8,12,14
    f = 13.56e6
    dt = 1 / (400 * f)                            # :29
    nHe, ne = 9.64e20, 2.56e14                    # :9-10
    n_each = particles_per_gpu // 2
    grid = RG.create_uniform_grid(np.arange(cells + 1) * dh, np.arange(cells + 1) * dh, device=device)
    grid._rt.comm_init_torch()                    # rank enters the Philox keys; rho all-reduce
    nx, ny = grid.n
    # weight: ne over the whole box shared by all electrons of the full job (weak scaling keeps it)
    wgt = ne * (cells * dh) ** 2 / (n_each * n_gpus_total)
    cap = int(n_each * capacity_factor) + 1024
    e = PIC.create_kinetic_species("e-", cap, -qe, me, wgt)
    iHe = PIC.create_kinetic_species("He+", cap, +qe, 3.99 * mp, wgt)          # :42 (electron weight, H8)
    He = PIC.FluidSpecies("He", 1.0, 0.0, 3.99 * mp, nHe * np.ones((nx, ny)), 300.0)
    solver = FDM.create_poisson_solver(grid, eps0)
    FDM.apply_periodic(solver, 1)                                                # :76
    L.check(grid._rt.lib.iskb_poisson_apply_dirichlet_edge(grid._rt.h, L.EDGE_LEFT, 0.0))    # :77
    L.check(grid._rt.lib.iskb_poisson_apply_dirichlet_edge(grid._rt.h, L.EDGE_RIGHT, 0.0))   # :78
    rank = grid._rt.rank
    _load_uniform(e, grid, n_each, 30000.0, [0.0, 0.0, 0.0], seed * 1000 + 1)
    _load_uniform(iHe, grid, n_each, 300.0, [0.0, 0.0, 0.0], seed * 1000 + 2)
    s1, s2, s3, s4 = [CH.CrossSection(t) for t in datasets.helium_electron()]
    sb, si = [CH.CrossSection(t) for t in datasets.helium_ion()]
    names = {"e": e, "He": He, "iHe": iHe}
    electron = CH.mcc(CH.reactions([(s1, "e + He --> e + He"),
                                    (s2, "e + He --> e + He", CH.MCC.Excitation(19.82)),
                                    (s3, "e + He --> e + He", CH.MCC.Excitation(20.61)),
                                    (s4, "e + He --> e + e + iHe", CH.MCC.Ionization(24.587))], names),
                      seed=seed * 7919 + 11)                                     # :52-57
    ion = CH.mcc(CH.reactions([(sb, "iHe + He --> iHe + He", CH.MCC.ElasticBackward()),
                               (si, "iHe + He --> iHe + He", CH.MCC.ElasticIsotropic())], names),
                 seed=seed * 7919 + 13)                                          # :59-62
    cfg = Config()
    cfg.grid, cfg.solver, cfg.pusher = grid, solver, PIC.create_boris_pusher()
    cfg.species, cfg.interactions = [e, iHe, He], [electron, ion]
    meta = {"grid_nodes": [nx, ny], "dh": dh, "dt": dt, "particles_per_gpu": 2 * n_each, "mcc_processes": [4, 2],
            "rank": rank}
    return Workload("c5", cfg, dt, (L.BND_DISCARD, L.BND_WRAP), rf=(L.EDGE_LEFT, 450.0, f), meta=meta)


def build_c4(particles_per_gpu=100_000_000, cells=1024, seed=1, capacity_factor=1.02, device=None):
    """C4 (or a shard of it): two-stream, fully periodic."""
    nHe = 1e24
    w = 2 * math.pi * 9e3 * math.sqrt(2e-6 * nHe)        # 10_two_streams.jl:13-14
    dh = 5e-3 * c0 / w                                    # :15
    vd = 1e7
    dt = 0.4 * dh / vd / math.sqrt(2.0)                   # :26
    n_each = particles_per_gpu // 2
    grid = RG.create_uniform_grid(np.arange(cells + 1) * dh, np.arange(cells + 1) * dh, device=device)
    grid._rt.comm_init_torch()
    nx, ny = grid.n
    wgt = nHe * (cells * dh) ** 2 / n_each
    cap = int(n_each * capacity_factor) + 1024
    mHe = 4.002602 * me / 5.48579903e-04                  # :10
    e = PIC.create_kinetic_species("e-", cap, -qe, me, wgt)
    iHe = PIC.create_kinetic_species("He+", cap, +qe, mHe, wgt)
    solver = FDM.create_poisson_solver(grid, eps0)
    FDM.apply_periodic(solver, 1)                         # :53
    FDM.apply_periodic(solver, 2)                         # :54
    half = n_each // 2
    _load_uniform(e, grid, half, 300.0, [+vd, 0.0, 0.0], seed * 1000 + 1)     # fwd beam :41
    _load_uniform(e, grid, n_each - half, 300.0, [-vd, 0.0, 0.0], seed * 1000 + 2)   # rev beam :42
    iHe._push(grid)
    vfill = np.array([1280.0, 1280.0, 1280.0])            # `iHe.v .= iHe.np = e.np` quirk (:67-68, H8) at C1 size
    L.check(grid._rt.lib.iskb_species_copy_positions(iHe._h, e._h, L.ptr(vfill)))   # :66
    iHe._touched_on_device()
    cfg = Config()
    cfg.grid, cfg.solver, cfg.pusher = grid, solver, PIC.create_boris_pusher()
    cfg.species, cfg.interactions = [e, iHe], []
    meta = {"grid_nodes": [nx, ny], "dh": dh, "dt": dt, "particles_per_gpu": 2 * n_each, "mcc_processes": []}
    return Workload("c4", cfg, dt, (L.BND_WRAP, L.BND_WRAP), meta=meta)


def build_walls(particles_per_gpu=100_000_000, cells=1024, seed=3, device=None):
    """N1 at scale: electrodes, absorbing walls and a reflecting block (see the module docstring)."""
    from . import configuration as CFG
    dh = 6.7 * 0.01 / 128
    f = 13.56e6
    dt = 1 / (400 * f)
    ne = 2.56e14
    n_each = particles_per_gpu // 2
    grid = RG.create_uniform_grid(np.arange(cells + 1) * dh, np.arange(cells + 1) * dh, device=device)
    grid._rt.comm_init_torch()
    nx, ny = grid.n
    wgt = ne * (cells * dh) ** 2 / n_each
    cap = n_each + 1024
    e = PIC.create_kinetic_species("e-", cap, -qe, me, wgt)
    iHe = PIC.create_kinetic_species("He+", cap, +qe, 3.99 * mp, wgt)
    cfg = Config()
    cfg.grid, cfg.solver, cfg.pusher = grid, FDM.create_poisson_solver(grid, eps0), PIC.create_boris_pusher()
    left = np.zeros((nx, ny), dtype=bool)
    left[0, :] = True
    right = np.zeros((nx, ny), dtype=bool)
    right[nx - 1, :] = True
    block = np.zeros((nx, ny), dtype=bool)
    b0, b1 = (3 * cells) // 8, (5 * cells) // 8
    block[b0:b1 + 1, b0:b1 + 1] = True
    CFG.create_electrode(left, cfg, fixed=True, phi=0.0)
    CFG.create_electrode(right, cfg, fixed=True, phi=0.0)
    PIC.track_surface_(cfg.tracker, block, PIC.create_reflective_surface())
    _load_uniform(e, grid, n_each, 30000.0, [0.0, 0.0, 0.0], seed * 1000 + 1)
    _load_uniform(iHe, grid, n_each, 300.0, [0.0, 0.0, 0.0], seed * 1000 + 2)
    # empty the block: particles loaded inside it are mirrored to the strip left of it
    for s in (e, iHe):
        x = s.x
        n = s.np
        inside = (x[:n, 0] >= b0 * dh) & (x[:n, 0] <= b1 * dh) & (x[:n, 1] >= b0 * dh) & (x[:n, 1] <= b1 * dh)
        x[:n, 0][inside] -= (b1 - b0 + 1) * dh
    cfg.species, cfg.interactions = [e, iHe], []
    meta = {"grid_nodes": [nx, ny], "dh": dh, "dt": dt, "particles_per_gpu": 2 * n_each, "mcc_processes": [],
            "surfaces": "2 fixed electrodes (whole edges), default absorbing, reflective block %d..%d" % (b0, b1)}
    return Workload("walls", cfg, dt, (L.BND_DISCARD, L.BND_DISCARD), rf=(L.EDGE_LEFT, 450.0, f), meta=meta)


def build_seed(particles_per_gpu=40_000_000, cells=32, seed=5, device=None):
    """N3 at scale (see the module docstring).  cells = radial cells: 32 -> 33 x 65 nodes (13_seed.jl's own grid),
    64 -> 65 x 125 nodes (the largest the dense path takes; its one-off host inversion needs minutes)."""
    nr = cells + 1
    nz = min(2 * cells + 1, 8192 // (cells + 1) - 1)
    dh = 0.08 / 32                                    # 13_seed.jl:9-13
    dt = 0.075e-9                                     # :18
    n_each = particles_per_gpu // 2
    grid = RG.create_axial_grid(np.arange(nr) * dh, np.arange(nz) * dh, device=device)
    grid._rt.comm_init_torch()
    cap = n_each + 1024
    wgt = 5e5 * 200_000 / n_each
    e = PIC.create_kinetic_species("e-", cap, -qe, me, wgt)
    iAr = PIC.create_kinetic_species("Ar+", cap, +qe, 3.99 * mp, wgt)
    solver = FDM.create_poisson_solver(grid, eps0)
    bot = np.zeros((nr, nz), dtype=bool)
    bot[:, 0] = True
    top = np.zeros((nr, nz), dtype=bool)
    top[:, nz - 1] = True
    FDM.apply_dirichlet(solver, bot, 0.0)             # :49-52
    FDM.apply_dirichlet(solver, top, 20.0)
    R, Lz = (nr - 1) * dh, (nz - 1) * dh
    for sp, T, sd in ((e, 11600.0, 1), (iAr, 300.0, 2)):
        src = PIC.create_thermalized_beam(sp, [0.5 * R, 0.5 * Lz], [0.0, 0.0, 0.0], dx=[0.0, 0.25 * Lz], T=T, rate=1.0)
        sp._push(grid)
        L.check(sp._rt.lib.iskb_species_sample_maxwellian(sp._h, int(n_each), L.ptr(src.wx), L.ptr(src.dx), L.ptr(src.wv),
                                                          L.ptr(src.dv), int(seed * 1000 + sd)))
        sp._touched_on_device()
    cfg = Config()
    cfg.grid, cfg.solver, cfg.pusher = grid, solver, PIC.create_axial_boris_pusher()
    cfg.species, cfg.interactions = [e, iAr], []
    PIC._set_pusher(grid._rt, cfg.pusher)
    meta = {"grid_nodes": [nr, nz], "dh": dh, "dt": dt, "particles_per_gpu": 2 * n_each, "mcc_processes": []}
    return Workload("seed", cfg, dt, (L.BND_NONE, L.BND_DISCARD), meta=meta)
