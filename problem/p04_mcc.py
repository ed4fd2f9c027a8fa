#!/usr/bin/env python
"""problem/04_mcc.jl -- electrons scattering elastically off a uniform oxygen background (one MCC process), two point
electrodes, periodic wrap; fields, densities, the collision-frequency map and both kinetic species go to openPMD-HDF5 every
iteration, XDMF export at the end."""
import os
import tempfile

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200 import diagnostics as DG
from iskra_b200 import xdmf as X
from iskra_b200.units_and_constants import cm, eps0, me, mp, ns, qe

PIC, FDM, RG, CFG, CH = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.configuration, ib.chemistry
ts, every = _common.steps(1000)

nx, ny = 20, 20                       # :1-8
dh, dt = 5 * cm, 1 * ns
Lx, Ly = nx * dh, ny * dh
O = CFG.create_fluid_species("O", 1.0, 0 * qe, 8 * mp, nx + 1, ny + 1)                  # :13-15
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, 1)
iO = PIC.create_kinetic_species("O+", 20_000, +1 * qe, 8 * mp, 1)
config = CFG.Config()
config.grid = grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
config.solver = solver = FDM.create_poisson_solver(grid, eps0)
config.pusher = PIC.create_boris_pusher()
config.species = [e, O, iO]
sigma = CH.CrossSection(np.array([[3e6, 0.01], [4e6, 0.1], [5e6, 2.0], [6e6, 0.01]]))    # :25
collisions = CH.mcc(CH.reactions([(sigma, "e + O --> O + e")], {"e": e, "O": O}), seed=4)
config.interactions = [collisions]
gnx, gny = grid.n                                                                       # :31-40
delta = np.ones((gnx, gny))
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[gnx - 1, 0] = 1
bcs[gnx - 1, gny - 1] = 2
CFG.create_electrode(bcs == 1, solver, grid, sigma=1e3 * eps0)
CFG.create_electrode(bcs == 2, solver, grid, fixed=True)

# + hooks: start (:47-52)
e.np = 0
O.n[...] = 0.0
PIC.init(PIC.MaxwellianSource(5e3 / dt, [1.0 * Lx, 1.0 * Ly], [0.5e6, 0.5e6, 0.0]), e, dt, grid)
PIC.init(PIC.DensitySource(5e3 * delta, grid), O, dt)
prefix = os.path.join(tempfile.gettempdir(), "04_mcc")


def iteration(i, t, dt_):                                                               # :55-69
    def save(it):
        for k in ("phi", "nuMCC-e--1", "nO", "ne-", "nO+", "E"):
            DG.save_record(it, k)
        DG.save_records(it, "e-/")
        DG.save_records(it, "O+/")
    DG.new_iteration(prefix, i, t, dt_, save)
    if i % every == 0 or i == ts:
        print([("iteration", i), ("e", e.np), ("iO", iO.np)])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP))              # after_push: wrap!(part, grid) (:42-44)

print("Exporting to XDMF...")                                                           # :71-87
electrons, fields, ions = X.new_document(), X.new_document(), X.new_document()
X.xdmf(lambda it: (X.write_species(it, electrons, "e-"), X.write_species(it, ions, "O+"), X.write_fields(it, fields)),
       range(1, ts + 1), prefix=prefix)
for doc, name in ((electrons, "electrons"), (fields, "fields"), (ions, "ions")):
    print(X.save_document(doc, name, prefix=prefix))
print("Complete!")
