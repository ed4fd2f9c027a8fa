#!/usr/bin/env python
"""problem/11_rf_discharge.jl -- 1-D capacitive RF discharge with e- / He+ MCC collisions."""
import math

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K, MHz, cm, eps0, me, mp, qe

PIC, FDM, RG, CH = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.chemistry
ts, every = _common.steps(400)

# + spatial and temporal parameters (:8-36)
nHe, ne = 9.64e20, 2.56e14
f = 13.56 * MHz
nx, ny = 128, 1
Lx = 6.7 * cm
dh = Lx / nx
Ly = ny * dh
dt = 1 / (400 * f)
numParticles = 512 * nx
weight = ne * (Lx * Ly) / numParticles
Te, Ti = 30_000 * K, 300 * K

# + species (:39-46)
grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
gnx, gny = grid.n
e = PIC.create_kinetic_species("e-", 200_000, -1 * qe, 1 * me, weight)
iHe = PIC.create_kinetic_species("He+", 200_000, +1 * qe, 3.99 * mp, weight)
He = PIC.FluidSpecies("He", 1.0, 0 * qe, 3.99 * mp, nHe * np.ones((gnx, gny)), Ti)
se = PIC.create_thermalized_beam(e, [Lx, Ly], [0.0, 0.0, 0.0], T=Te, rate=numParticles / dt)
si = PIC.create_thermalized_beam(iHe, [Lx, Ly], [0.0, 0.0, 0.0], T=Ti, rate=numParticles / dt)

# + reactions (:49-62); synthetic tables on the reference's thresholds (datasets.py)
s1, s2, s3, s4 = [CH.CrossSection(t) for t in ib.datasets.helium_electron()]
sb, si_ = [CH.CrossSection(t) for t in ib.datasets.helium_ion()]
names = {"e": e, "He": He, "iHe": iHe}
electron = CH.mcc(CH.reactions([(s1, "e + He --> e + He"),
                                (s2, "e + He --> e + He", CH.MCC.Excitation(19.82)),
                                (s3, "e + He --> e + He", CH.MCC.Excitation(20.61)),
                                (s4, "e + He --> e + e + iHe", CH.MCC.Ionization(24.587))], names), seed=1)
ion = CH.mcc(CH.reactions([(sb, "iHe + He --> iHe + He", CH.MCC.ElasticBackward()),
                           (si_, "iHe + He --> iHe + He", CH.MCC.ElasticIsotropic())], names), seed=2)

# + grid, solver and pusher, boundary conditions (:65-83)
solver = FDM.create_poisson_solver(grid, eps0)
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[0, :] = 1
bcs[gnx - 1, :] = 2
FDM.apply_periodic(solver, 1)
FDM.apply_dirichlet(solver, bcs == 1, 0.0)
FDM.apply_dirichlet(solver, bcs == 2, 0.0)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_boris_pusher()
config.species, config.interactions = [e, iHe, He], [electron, ion]

# + hooks: start (:86-89), iteration (:92-115)
PIC.init(se, e, dt, grid)
PIC.init(si, iHe, dt, grid)


def iteration(i, t, dt_):
    FDM.apply_dirichlet(solver, bcs == 1, 450 * math.sin(2 * math.pi * f * t))      # :95
    if i % every == 0 or i == ts:
        print([("iteration", i), ("e-", e.np), ("He+", iHe.np)])


PIC.hooks.after_loop = iteration
# after_push: discard!(dims=1), wrap!(dims=2) (:80-83)
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_DISCARD, ib._lib.BND_WRAP))
print("Complete!", "MCC totals e-:", electron.totals().tolist(), "He+:", ion.totals().tolist())
