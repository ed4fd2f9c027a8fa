#!/usr/bin/env python
"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (not collected by pytest: needs one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu/check_rank_invariance.py

DESIGN.md 14.5 / SURVEY.md 8e: the particles are sharded by index slice, every rank deposits its slice into 64-bit fixed-point
sums, the sums are all-reduced as integers and every rank solves the same field.  Checked here:
 (a) rho, phi and E are BIT-IDENTICAL on every rank (every rank must push its slice in the same field);
 (b) against a single-GPU run of all the particles: the same live rows, fields and rows equal up to the rounding of the
     FP64 sums inside a warp (a warp of the sharded run adds other rows than a warp of the single-GPU run; only what leaves
     a warp is an integer) -- 1e-11 of the largest value after 12 steps.
Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(ib, O, x, v, ids, device, shard):
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx = ny = 257
    dx = 5.234375e-4
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dx, device=device)
    ps = FDM.create_poisson_solver(g, O.eps0)
    FDM.apply_periodic(ps, 1)
    left = np.zeros((nx, ny), bool)
    left[0, :] = True
    right = np.zeros((nx, ny), bool)
    right[nx - 1, :] = True
    FDM.apply_dirichlet(ps, left, 30.0)
    FDM.apply_dirichlet(ps, right, 0.0)
    if shard:
        g._rt.comm_init_torch()
    sps = []
    for k, (name, q, m) in enumerate((("e-", -O.qe, O.me), ("He+", O.qe, 3.99 * O.mp))):
        n = x[k].shape[0]
        sp = PIC.create_kinetic_species(name, n + 64, q, m, 5.0e5)
        sp.x[:n], sp.v[:n] = x[k], v[k]
        sp.id[:n] = ids[k]
        sp.id[n:] = np.arange(10_000_000 + 1, 10_000_000 + 65, dtype=np.uint32)     # parked ids: distinct from the live ones
        sp.np = n
        sps.append(sp)
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), sps
    return g, cfg, sps


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(rank)
    import iskra_b200 as ib
    from oracle import pic_oracle as O
    from sharding_helpers import slice_for_rank
    PIC = ib.particle_in_cell
    n_tot, steps, dt = 1_200_000, 12, 1.8436578171091445e-10
    L = 256 * 5.234375e-4
    rng = np.random.default_rng(2024)
    X, V, I = [], [], []
    for T, m in ((30000.0, O.me), (300.0, 3.99 * O.mp)):
        X.append(rng.random((n_tot, 2)) * L)
        V.append(rng.standard_normal((n_tot, 3)) * O.thermal_speed(T, m))
        I.append(np.arange(1, n_tot + 1, dtype=np.uint32))
    lo, hi = slice_for_rank(n_tot, rank, world)
    g, cfg, sps = build(ib, O, [a[lo:hi] for a in X], [a[lo:hi] for a in V], [a[lo:hi] for a in I], rank, True)
    PIC.solve(cfg, dt, steps, after_push=(2, 1), sort_interval=3)
    rho, phi, E = g._rt.fields()
    mine = torch.from_numpy(np.concatenate([rho.ravel(), phi.ravel(), E.ravel()]).view(np.int64).copy()).cuda()
    allf = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allf, mine)
    same_on_ranks = all(bool(torch.equal(allf[0], t)) for t in allf)
    res = {"world": world, "fields_identical_on_all_ranks": same_on_ranks}
    # the rows of this rank against the single-GPU run (rank 0 runs it; the others receive the rows they need by id)
    if rank == 0:
        g1, cfg1, sps1 = build(ib, O, X, V, I, 0, False)
        PIC.solve(cfg1, dt, steps, after_push=(2, 1), sort_interval=3)
        r1, p1, e1 = g1._rt.fields()
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        res["rel_diff_to_single_gpu"] = {"rho": rel(rho, r1), "phi": rel(phi, p1), "E": rel(E, e1)}
        res["fields_agree_with_single_gpu"] = bool(max(res["rel_diff_to_single_gpu"].values()) < 1e-11)
        res["max_abs_rho"] = float(np.abs(rho).max())
        ok_rows = True
        for a, b in zip(sps, sps1):
            ids1 = b.id[: b.np]
            o1 = np.argsort(ids1)
            ida = a.id[: a.np]
            pos = np.searchsorted(ids1[o1], ida)
            ok_rows &= bool(np.all(ids1[o1][pos] == ida))
            dx = np.abs(b.x[: b.np][o1][pos] - a.x[: a.np]).max() / L
            dv = np.abs(b.v[: b.np][o1][pos] - a.v[: a.np]).max() / np.abs(a.v[: a.np]).max()
            ok_rows &= bool(dx < 1e-11 and dv < 1e-11)
        res["rank0_rows_agree_with_single_gpu"] = ok_rows
        res["live_rows_single_gpu"] = [int(s.np) for s in sps1]
    live = torch.tensor([float(s.np) for s in sps], dtype=torch.float64, device="cuda")
    dist.all_reduce(live)
    if rank == 0:
        res["live_rows_sum_over_ranks"] = [int(v) for v in live.tolist()]
        res["ok"] = bool(res["fields_identical_on_all_ranks"] and res["fields_agree_with_single_gpu"] and
                         res["rank0_rows_agree_with_single_gpu"] and res["live_rows_sum_over_ranks"] == res["live_rows_single_gpu"])
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
