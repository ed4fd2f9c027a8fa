#!/usr/bin/env python
"""problem/07_boundaries.jl -- electrodes (sigma-driven, floating, grounded) and a reflecting block; default absorbing walls."""
import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K, cm, eps0, me, ns, qe

PIC, FDM, RG, CFG = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.configuration
ts, every = _common.steps(200)

nx, ny = 10, 10                      # :7-13
dh, dt = 10 * cm, 10 * ns
Lx, Ly = nx * dh, ny * dh
config = CFG.Config()
config.grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
e = PIC.create_kinetic_species("e-", 50_000, -1 * qe, 1 * me, 50e3)                                   # :16
gamma = PIC.create_thermalized_beam(e, [Lx / 4, Ly / 2], [0.0, 0.0, 0.0], dx=[Lx / 4, Lx / 4], T=300 * K, rate=20_000 / dt)
config.solver = FDM.create_poisson_solver(config.grid, eps0)
config.pusher = PIC.create_boris_pusher()
config.species = [e]
gnx, gny = config.grid.n
bcs = np.zeros((gnx, gny), dtype=np.int8)                                                             # :27-32
bcs[0, 1:gny - 1] = 1
bcs[gnx - 1, 4:7] = 2
bcs[gnx - 2, 0] = 3
bcs[gnx - 2, gny - 1] = 3
bcs[5:8, 4:7] = 4
reflecting = PIC.create_reflective_surface()
driven = CFG.create_electrode(bcs == 1, config, sigma=1 * eps0)                                       # :34-37
floating = CFG.create_electrode(bcs == 2, config)
grounded = CFG.create_electrode(bcs == 3, config, fixed=True)
PIC.track_surface_(config.tracker, bcs == 4, reflecting)


def after_loop(i, t, dt_):
    if i % every == 0 or i == ts:
        print([("iteration", i), ("e", e.np), ("dq floating", floating.dq), ("dq driven", driven.dq)])


PIC.hooks.after_loop = after_loop
PIC.init(gamma, e, dt, config.grid)                                                                   # :66
PIC.solve(config, dt, ts)                          # default after_push: wrap!(part, grid)  ParticleInCell.jl:41
print("Complete!")
