#!/usr/bin/env python
"""problem/05_dsmc.jl -- DSMC collisions between kinetic electrons and kinetic oxygen atoms (field part left out: the
script's point electrodes need create_electrode on single nodes; here the particles move in E = 0 with wrap!)."""
import math

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import cm, eps0, me, mp, ns, qe

PIC, FDM, RG, CH = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.chemistry
ts, every = _common.steps(50)

nx, ny = 20, 20                      # :7-13
dh, dt = 5 * cm, 1 * ns
Lx, Ly = nx * dh, ny * dh
grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
O_ = PIC.create_kinetic_species("O", 20_000, 0 * qe, 8 * mp, 1)                                        # :17-19
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, 1)
sigma = CH.CrossSection(np.arange(3e6, 6.1e6, 1e6), [0.01, 0.1, 2.0, 0.01])                          # :27
collisions = CH.dsmc(CH.reactions([(sigma, "e + O --> O + e")], {"e": e, "O": O_}), seed=5)         # :28-30
solver = FDM.create_poisson_solver(grid, eps0)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_boris_pusher()
config.species, config.interactions = [e, O_], [collisions]
vth = math.sqrt(2 * 1.3806503e-23 * 300 / O_.m)                                                     # :46-49
e.np = 0
O_.np = 0
PIC.init(PIC.MaxwellianSource(5e3 / dt, [1.0 * Lx, 1.0 * Ly], [0.5e6, 0.5e6, 0.0]), e, dt, grid)      # :54-55
PIC.init(PIC.MaxwellianSource(5e3 / dt, [1.0 * Lx, 1.0 * Ly], [vth, vth, vth]), O_, dt, grid)
k0 = 0.5 * me * float((e.v[: e.np] ** 2).sum()) + 0.5 * O_.m * float((O_.v[: O_.np] ** 2).sum())


def iteration(i, t, dt_):
    if i % every == 0 or i == ts:
        k1 = 0.5 * me * float((e.v[: e.np] ** 2).sum()) + 0.5 * O_.m * float((O_.v[: O_.np] ** 2).sum())
        print([("iteration", i), ("e", e.np), ("O", O_.np), ("kinetic energy / initial", k1 / k0)])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP))                            # :41-43
print("Complete!")
