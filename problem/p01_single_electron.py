#!/usr/bin/env python
"""problem/01_single_electron.jl -- one electron between two point electrodes (one sigma-driven, one grounded), absorbing walls;
every iteration goes to openPMD-HDF5 (hdf5.jl) and the run ends with the XDMF export (XDMF.jl), like the reference's stop()."""
import os
import tempfile

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200 import diagnostics as DG
from iskra_b200 import xdmf as X
from iskra_b200.units_and_constants import K, cm, eps0, me, ns, qe

PIC, FDM, RG, CFG = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.configuration
ts, every = _common.steps(1000)

# + spatial and temporal parameters (:1-8)
nx, ny = 20, 20
dh, dt = 5 * cm, 20 * ns
Lx, Ly = nx * dh, ny * dh

# + species and sources (:10-14)
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, 1)
gamma = PIC.create_thermalized_beam(e, [Lx, Ly], [+0.05 * dh / dt, 0.0, 0.0], T=300 * K, rate=1.0 / dt)

# + grid, solver and pusher (:16-22)
config = CFG.Config()
config.grid = grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
config.solver = solver = FDM.create_poisson_solver(grid, eps0)
config.pusher = PIC.create_boris_pusher()
config.species, config.interactions = [e], []

# + boundary conditions (:24-35): two single nodes
gnx, gny = grid.n
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[gnx - 1, gny - 1] = 1
bcs[gnx - 1, 0] = 2
CFG.create_electrode(bcs == 1, solver, grid, sigma=-1 * eps0)
CFG.create_electrode(bcs == 2, solver, grid, fixed=True)

# + hooks (:37-52)
prefix = os.path.join(tempfile.gettempdir(), "01_single_particle")
e.np = 0
PIC.init(gamma, e, dt, grid)


def iteration(i, t, dt_):
    def save(it):
        DG.save_records(it, "e-/")
        DG.save_record(it, "rho")
        DG.save_record(it, "phi")
    DG.new_iteration(prefix, i, t, dt_, save)
    if i % every == 0 or i == ts:
        print([("iteration", i), ("e", e.np)])


PIC.hooks.after_loop = iteration
PIC.solve(config, dt, ts, after_push=(ib._lib.BND_DISCARD, ib._lib.BND_DISCARD))      # after_push: discard!(part, grid) (:33-35)

# + stop (:54-66): XDMF export of what was written
print("Exporting to XDMF...")
electrons, fields = X.new_document(), X.new_document()
X.xdmf(lambda it: (X.write_species(it, electrons, "e-"), X.write_fields(it, fields)), range(1, ts + 1), prefix=prefix)
print(X.save_document(electrons, "electrons", prefix=prefix))
print(X.save_document(fields, "fields", prefix=prefix))
print("Complete!")
