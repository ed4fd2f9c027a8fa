#!/usr/bin/env python
"""problem/12_avalanche.jl -- electron avalanche with ionisation MCC (the particle count grows)."""
import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K, eps0, me, mp, ns, qe

PIC, FDM, RG, CH = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.chemistry
ts, every = _common.steps(256)

# + parameters (:3-19)
nAr, T = 1e22, 300.0 * K
Efield, d = 5_000, 0.04
nx, ny = 32, 64
dh = d / nx
dt = 0.075 * ns
Lx, Ly = nx * dh, ny * dh

# + species (:22-27)
grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
gnx, gny = grid.n
e = PIC.create_kinetic_species("e-", 100_000, -1 * qe, 1 * me, 1)
iAr = PIC.create_kinetic_species("Ar+", 100_000, +1 * qe, 3.99 * mp, 1)
Ar = PIC.FluidSpecies("Ar", 1.0, 0 * qe, 3.99 * mp, nAr * np.ones((gnx, gny)), T)
se = PIC.create_thermalized_beam(e, [dh, dh], [0.0, 0.0, 0.0], dx=[0.0, Ly / 2], T=T, rate=1 / dt)

# + reactions (:30-37)
t1, t2, t3, t4 = [CH.CrossSection(t) for t in ib.datasets.argon_electron()]
names = {"e": e, "Ar": Ar, "iAr": iAr}
electron = CH.mcc(CH.reactions([(t1, "e + Ar --> e + Ar"),
                                (t2, "e + Ar --> e + Ar", CH.MCC.Excitation(11.55)),
                                (t3, "e + Ar --> e + Ar", CH.MCC.Excitation(13.00)),
                                (t4, "e + Ar --> e + e + iAr", CH.MCC.Ionization(15.7))], names), seed=3)

# + solver and boundary conditions (:40-59)
solver = FDM.create_poisson_solver(grid, eps0)
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[0, :] = 1
bcs[gnx - 1, :] = 2
FDM.apply_periodic(solver, 1)
FDM.apply_dirichlet(solver, bcs == 1, 0.0)
FDM.apply_dirichlet(solver, bcs == 2, Efield * d)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_boris_pusher()
config.species, config.interactions = [e, iAr, Ar], [electron]

# + hooks (:62-64, :67-86)
PIC.init(se, e, dt, grid)


def iteration(i, t, dt_):
    if i % every == 0 or i == ts:
        print([("iteration", i), ("electrons", e.np), ("ions", iAr.np)])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_DISCARD, ib._lib.BND_WRAP))      # :55-59
print("Complete!", "MCC totals:", electron.totals().tolist())
