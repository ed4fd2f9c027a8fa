#!/usr/bin/env python
"""problem/13_seed.jl -- axisymmetric (r-z) avalanche seeded by one electron per step near the axis."""
import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K, eps0, me, mp, ns, qe

PIC, FDM, RG, CH = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.chemistry
ts, every = _common.steps(128)

nAr, T = 1e22, 300.0 * K             # :3-18
Efield, d = 20_000, 0.08
nx, ny = 32, 64
dh = d / nx
dt = 0.075 * ns
Lx, Ly = nx * dh, ny * dh
grid = RG.create_axial_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)                            # :40
gnx, gny = grid.n
e = PIC.create_kinetic_species("e-", 200_000, -1 * qe, 1.00 * me, 5e5)                                # :24-27
iAr = PIC.create_kinetic_species("Ar+", 200_000, +1 * qe, 3.99 * mp, 5e5)
Ar = PIC.FluidSpecies("Ar", 1.0, 0 * qe, 3.99 * mp, nAr * np.ones((gnx, gny)), T)
se = PIC.create_thermalized_beam(e, [dh, dh], [0.0, 0.0, 0.0], T=T, rate=1 / dt)
t1, t2, t3, t4 = [CH.CrossSection(t) for t in ib.datasets.argon_electron()]
names = {"e": e, "Ar": Ar, "iAr": iAr}
electron = CH.mcc(CH.reactions([(t1, "e + Ar --> e + Ar"),                                            # :31-36
                                (t2, "e + Ar --> e + Ar", CH.MCC.Excitation(11.55)),
                                (t3, "e + Ar --> e + Ar", CH.MCC.Excitation(13.00)),
                                (t4, "e + Ar --> e + e + iAr", CH.MCC.Ionization(15.7))], names), seed=4)
solver = FDM.create_poisson_solver(grid, eps0)                                                        # :41
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[:, 0] = 1
bcs[:, gny - 1] = 2
FDM.apply_dirichlet(solver, bcs == 1, 0.0)                                                            # :49-52
FDM.apply_dirichlet(solver, bcs == 2, Efield * d)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_axial_boris_pusher()             # :42
config.species, config.interactions = [e, iAr, Ar], [electron]
PIC.init(se, e, dt, grid)                                                                             # :59


def iteration(i, t, dt_):
    if i % every == 0 or i == ts:
        print([("iteration", i), ("electrons", e.np), ("ions", iAr.np)])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_NONE, ib._lib.BND_DISCARD))                         # discard!(dims=2)  :54-56
print("Complete!", "MCC totals:", electron.totals().tolist())
