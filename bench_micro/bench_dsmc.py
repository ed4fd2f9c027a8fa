#!/usr/bin/env python
"""Times PIC.perform!(dsmc, ...) on the device: cell lists (count / scan / fill) + collisions, CUDA events
around the call, and the Python oracle on a small sample beside it.  Prints one JSON line.
  python bench_micro/bench_dsmc.py [--cells 1024] [--particles 100000000] [--calls 10]"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--particles", type=int, default=100_000_000)
    ap.add_argument("--calls", type=int, default=10)
    a = ap.parse_args()
    import torch
    import iskra_b200 as ib
    from iskra_b200 import _lib as L
    from oracle import pic_oracle as O
    PIC, CH = ib.particle_in_cell, ib.chemistry
    dh = 0.05
    n = a.particles // 2
    g = ib.regular_grids.create_uniform_grid(np.arange(a.cells + 1) * dh, np.arange(a.cells + 1) * dh)
    g._rt.use_torch_stream()
    sig = np.stack([np.arange(3e6, 6.1e6, 1e6), [0.01, 0.1, 2.0, 0.01]], axis=1)
    e = PIC.create_kinetic_species("e-", n + 1024, -O.qe, O.me, 1.0)
    ox = PIC.create_kinetic_species("O", n + 1024, 0.0, 8 * O.mp, 1.0)
    L_ = a.cells * dh
    for sp, vth, sd in ((e, 3e6, 1), (ox, 560.0, 2)):
        src = PIC.MaxwellianSource(1.0, [L_, L_], [vth, vth, vth])
        sp._push(g)
        L.check(sp._rt.lib.iskb_species_sample_maxwellian(sp._h, n, L.ptr(src.wx), L.ptr(src.dx), L.ptr(src.wv), L.ptr(src.dv), sd))
        sp._touched_on_device()
    ppc = n / float(a.cells * a.cells)
    # dt for ~2 candidate pairs per cell and call: Nc = Na*W/(dx dy) * Nb * dt * sgmax
    dt = 2.0 / (ppc / (dh * dh) * ppc * 1e7)
    d = CH.dsmc(CH.reactions([(CH.CrossSection(sig), "e + O --> O + e")], {"e": e, "O": ox}), seed=3)
    cfg = ib.configuration.Config()
    cfg.grid, cfg.species, cfg.interactions = g, [e, ox], [d]
    d.perform_(None, dt, cfg, want_nu=False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    cand = coll = 0
    for _ in range(a.calls):
        _, nc, ncoll = d.perform_(None, dt, cfg, want_nu=False)
        cand += nc
        coll += ncoll
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.calls
    # the C oracle (one thread, like the reference) on a bounded sample with the same particles per cell
    import ctypes as C
    from oracle import c_oracle as CO
    cs = 256
    no = int(ppc * cs * cs)
    cg = CO.make_grid(cs + 1, cs + 1, dh, dh)
    rng = np.random.default_rng(0)
    csp = []
    for m, vth in ((O.me, 3e6), (8 * O.mp, 560.0)):
        s = CO.CSpecies(no, 0.0, m, 1.0)
        vv = rng.standard_normal((3, no)) * vth
        s.set(rng.random(no) * cs * dh, rng.random(no) * cs * dh, vv[0], vv[1], vv[2])
        csp.append(s)
    gn, sgv = np.ascontiguousarray(sig[:, 0]), np.ascontiguousarray(sig[:, 1])
    rem = np.zeros((cs + 1) * (cs + 1))
    crng = CO.make_rng(1)
    fn = CO.lib().orc_dsmc_perform
    fn.restype = C.c_int64
    t0 = time.perf_counter()
    for _ in range(3):
        fn(csp[0].ref(), csp[1].ref(), C.byref(cg), CO.dp(gn), CO.dp(sgv), C.c_int32(len(gn)), C.c_double(dt), CO.dp(rem), None, None,
           C.byref(crng))
    cpu_s = (time.perf_counter() - t0) / 3
    print(json.dumps({"what": "PIC.perform!(dsmc) on the device (cell lists + one thread per cell)", "grid_cells": [a.cells, a.cells],
                      "particles": 2 * n, "particles_per_cell_per_species": ppc, "ms_per_call": ms,
                      "particle_visits_per_s": 2 * n / (ms * 1e-3), "candidate_pairs_per_call": cand / a.calls,
                      "collisions_per_call": coll / a.calls,
                      "list_build_bytes_per_particle": 16 + 4 + 4 + 4, "list_build_GBps_if_alone": 2 * n * 28 / (ms * 1e-3) / 1e9,
                      "cpu_oracle_c_1_thread": {"particles": 2 * no, "seconds_per_call": cpu_s, "particle_visits_per_s": 2 * no / cpu_s}}))


if __name__ == "__main__":
    main()
