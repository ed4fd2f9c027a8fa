#!/usr/bin/env python
"""problem/ts_unstable.jl and problem/ts_stable.jl -- the normalised two-stream cases (me = qe = eps0 = 1, 32 x 1 cells of
2 pi / 32, two cold beams of +-0.01): they differ only in the density, nHe = 0.1^2 (k v0 / omega_p = 0.1, unstable) or
0.001^2 (k v0 / omega_p = 10, stable).   python p_ts.py [--stable] [--steps N]"""
import math
import sys

stable = "--stable" in sys.argv
if stable:
    sys.argv.remove("--stable")
import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K

PIC, FDM, RG = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids
ts, every = _common.steps(300)

# + spatial and temporal parameters (:8-31)
me = qe = 1.0
vdrift = 0.01
mHe = 4.002602 * me / 5.48579903e-04
nHe = (0.001) ** 2 if stable else (0.1) ** 2
ne = 1 * nHe
nx, ny = 32, 1
dh = 2 * math.pi / nx
eps0 = 1.0
wp = math.sqrt(ne * qe ** 2 / (me * eps0))
electronParticles = 10_000
electronNumRatio = ne * (nx * dh * ny * dh) / electronParticles
dt = 0.19634954084936207
Lx, Ly = nx * dh, ny * dh

# + species and sources (:40-46)
grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, electronNumRatio)
iHe = PIC.create_kinetic_species("He+", 20_000, +1 * qe, 1 * mHe, electronNumRatio)
fwd = PIC.create_thermalized_beam(e, [Lx, Ly], [+vdrift, 0, 0], T=0.0 * K, rate=electronParticles / 2 / dt)
rev = PIC.create_thermalized_beam(e, [Lx, Ly], [-vdrift, 0, 0], T=0.0 * K, rate=electronParticles / 2 / dt)

# + grid, solver and pusher, boundary conditions (:49-61)
solver = FDM.create_poisson_solver(grid, eps0)
FDM.apply_periodic(solver, 1)
FDM.apply_periodic(solver, 2)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_boris_pusher()
config.species, config.interactions = [e, iHe], []

# + hooks: start (:64-73)
e.np = 0
PIC.init(fwd, e, dt, grid)
PIC.init(rev, e, dt, grid)
x = e.x
x[:, 0] += 0.001 * np.cos(2 * math.pi * x[:, 0] / Lx)
# The perturbation pushes the electrons within 0.001 of x = Lx past the last node; the reference's first gather would then
# index E[nx+2, ...] (BoundsError) unless its seed happens to leave that strip empty.  They are wrapped here before the loop.
PIC.wrap_(e, grid)
iHe.x[...] = e.x
iHe.v[...] = 0.0
iHe.np = e.np
print("omega_p:", wp, " k v0 / omega_p:", 1.0 * vdrift / wp, " electrons:", e.np, " wg:", electronNumRatio)


def iteration(i, t, dt_):
    if i % every == 0 or i == ts:
        _, _, E = grid._rt.fields(rho=False, phi=False)
        print([("iteration", i), ("e", e.np), ("U_E", float(np.sum(E[..., 0] ** 2)))])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP))
print("Complete!")
