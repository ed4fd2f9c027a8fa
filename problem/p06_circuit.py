#!/usr/bin/env python
"""problem/06_circuit.jl -- a driven electrode fed through an RLC circuit (V1, L1, C1, R1), a grounded one opposite, 1000
electrons in between; the circuit's probes and the electrons go to openPMD-HDF5 every iteration, XDMF export at the end."""
import math
import os
import tempfile

import _common
import numpy as np

import iskra_b200 as ib
from iskra_b200 import circuit as CIR
from iskra_b200 import diagnostics as DG
from iskra_b200 import xdmf as X
from iskra_b200.units_and_constants import cm, eps0, me, mp, ns, qe

PIC, FDM, RG, CFG = ib.particle_in_cell, ib.finite_difference_method, ib.regular_grids, ib.configuration
ts, every = _common.steps(250)

nx, ny = 20, 20                       # :9-15
dh, dt = 5 * cm, 10 * ns
Lx, Ly = nx * dh, ny * dh
config = CFG.Config()
O = CFG.create_fluid_species("O", 1.0, 0 * qe, 8 * mp, nx + 1, ny + 1)                 # :18
e = PIC.create_kinetic_species("e-", 20_000, -1 * qe, 1 * me, 50e3)
config.grid = RG.create_uniform_grid(np.arange(nx + 1) * dh, np.arange(ny + 1) * dh)
config.cells = RG.create_staggered_grid(config.grid)
config.solver = FDM.create_poisson_solver(config.grid, eps0)
config.pusher = PIC.create_boris_pusher()
config.species = [e, O]
gnx, gny = config.grid.n                                                                # :27-35
delta = np.ones((gnx, gny))
bcs = np.zeros((gnx, gny), dtype=np.int8)
bcs[0, :] = 1
bcs[gnx - 1, :] = 2
driven = CFG.create_electrode(bcs == 1, config, sigma=1 * eps0)
grounded = CFG.create_electrode(bcs == 2, config, fixed=True)
NH, NF = 1e-9, 1e-9
config.circuit = CIR.rlc(CIR.netlist([                                                  # :37-42
    ("V1", 3, "GND", lambda t: math.sin(2 * math.pi * 5e6 * t)),
    ("L1", "NOD", "VCC", 1000 * NH),
    ("C1", "NOD", "VCC", 1000 * NF),
    ("R1", "GND", "NOD", 1),
]))
prefix = os.path.join(tempfile.gettempdir(), "06_circuit")


def after_loop(i, t, dt_):                                                              # :48-57
    def save(it):
        DG.save_records(it, "e-/")
        for k in ("Q1", "I1", "V1", "Vext"):
            DG.save_record(it, k)
    DG.new_iteration(prefix, i, t, dt_, save)
    if i % every == 0 or i == ts:
        print([("iteration", i), ("e", e.np), ("I1", config.circuit.i), ("Q1", config.circuit.q)])


PIC.hooks.after_loop = after_loop
PIC.init(PIC.MaxwellianSource(1e3 / dt, [1.0 * Lx, 1.0 * Ly], [0.0, 0.0, 0.0]), e, dt, config.grid)   # :74
PIC.init(PIC.DensitySource(0 * delta, config.grid), O, dt)                                               # :75
PIC.solve(config, dt, ts, fused=False)        # the circuit advances between advance! and density (ParticleInCell.jl:116)

print("Exporting to XDMF...")                                                           # :59-72
electrons, probes = X.new_document(), X.new_document()
X.xdmf(lambda it: (X.write_species(it, electrons, "e-"), X.write_probes(it, probes)), range(1, ts + 1), prefix=prefix)
print(X.save_document(electrons, "electrons", prefix=prefix))
print(X.save_document(probes, "probes", prefix=prefix))
print("Complete!")
