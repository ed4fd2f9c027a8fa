#!/usr/bin/env python
"""problem/10_two_streams.jl -- 1-D electrostatic two-stream instability (quasi-2D, one cell in y)."""
import math

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200.units_and_constants import K, c0, eps0, me, qe

PIC, FDM, RG = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids
ts, every = _common.steps(200)

# + spatial and temporal parameters (:4-31)
nHe = 1e24
f = 9e3 * math.sqrt(2e-6 * nHe)
w = 2 * math.pi * f
mHe = 4.002602 * me / 5.48579903e-04
nx, ny = 128, 1
dh = 5e-3 * c0 / w
vdrift = 1e7
dt = 0.4 * dh / vdrift / math.sqrt(2.0)
Lx, Ly = nx * dh, ny * dh
electronParticles = nx * 10
electronNumRatio = nHe * (nx * dh * ny * dh) / electronParticles

# + species and sources (:34-42)
xs, ys = np.arange(nx + 1) * dh, np.arange(ny + 1) * dh
grid = RG.create_uniform_grid(xs, ys)
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, electronNumRatio)
iHe = PIC.create_kinetic_species("He+", 20_000, +1 * qe, mHe, electronNumRatio)
fwd = PIC.create_thermalized_beam(e, [Lx, Ly], [+vdrift, 0, 0], T=300 * K, rate=electronParticles / 2 / dt)
rev = PIC.create_thermalized_beam(e, [Lx, Ly], [-vdrift, 0, 0], T=300 * K, rate=electronParticles / 2 / dt)

# + grid, solver and pusher (:45-50), boundary conditions (:53-58)
solver = FDM.create_poisson_solver(grid, eps0)
FDM.apply_periodic(solver, 1)
FDM.apply_periodic(solver, 2)
config = ib.configuration.Config()
config.grid, config.solver, config.pusher = grid, solver, PIC.create_boris_pusher()
config.species, config.interactions = [e, iHe], []

# + hooks: start (:61-69)
e.np = 0
PIC.init(fwd, e, dt, grid)
PIC.init(rev, e, dt, grid)
iHe.x[...] = e.x          # iHe.x .= e.x
iHe.v[...] = e.np         # `iHe.v .= iHe.np = e.np` (:67-68) assigns the particle count to the velocities, kept as is
iHe.np = e.np


def iteration(i, t, dt_):
    if i % every == 0 or i == ts:
        _, _, E = grid._rt.fields(rho=False, phi=False)
        print([("iteration", i), ("e", e.np), ("U_E", float(np.sum(E[..., 0] ** 2)))])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP))     # after_push: wrap!(part, grid) (:56-58)
print("Complete!")
