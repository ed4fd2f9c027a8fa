"""world_size-2 gloo tests (CPU) of the N>1 host logic: particles sharded by index slice, every
rank deposits its slice, rho is summed with one all-reduce, and the result equals the unsharded
deposit (linearity of the CIC scatter) -- SURVEY.md 8e.  The per-rank compute here is the C
oracle (the GPU path is exercised by the -m gpu tests and the multi-GPU bench)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sharding_helpers import max_over_ranks, slice_for_rank, sum_over_ranks
    from oracle import c_oracle as CO
    Lc = CO.lib()
    nx, ny, dx, n = 33, 17, 1e-3, 20001
    rng = np.random.default_rng(0)                      # identical stream on both ranks
    x, y = rng.random(n) * (nx - 1) * dx, rng.random(n) * (ny - 1) * dx
    wg = 1e5 * (0.5 + rng.random(n))
    lo, hi = slice_for_rank(n, rank, world)
    g = CO.make_grid(nx, ny, dx, dx)
    s = CO.CSpecies(hi - lo, -1.6e-19, 9.1e-31, 1e5)
    z = np.zeros(hi - lo)
    s.set(x[lo:hi], y[lo:hi], z, z, z, wg[lo:hi])
    u = np.zeros(nx * ny)
    Lc.orc_deposit(C.byref(g), s.ref(), CO.dp(u))
    sum_over_ranks(u, dist)                             # the one exchange step of the path
    # unique-id style broadcast plumbing (Runtime.comm_init_torch uses the same call pattern)
    idbuf = torch.from_numpy(np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, dtype=np.uint8))
    dist.broadcast(idbuf, 0)
    tmax = max_over_ranks(1.0 + rank, dist)
    np.save(os.path.join(out_dir, "u%d.npy" % rank), u)
    np.save(os.path.join(out_dir, "meta%d.npy" % rank), np.array([idbuf.numpy().sum(), tmax, lo, hi], dtype=np.float64))
    dist.destroy_process_group()


def test_sharded_deposit_allreduce_equals_unsharded(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import c_oracle as CO
    Lc = CO.lib()
    nx, ny, dx, n = 33, 17, 1e-3, 20001
    rng = np.random.default_rng(0)
    x, y = rng.random(n) * (nx - 1) * dx, rng.random(n) * (ny - 1) * dx
    wg = 1e5 * (0.5 + rng.random(n))
    g = CO.make_grid(nx, ny, dx, dx)
    s = CO.CSpecies(n, -1.6e-19, 9.1e-31, 1e5)
    z = np.zeros(n)
    s.set(x, y, z, z, z, wg)
    full = np.zeros(nx * ny)
    Lc.orc_deposit(C.byref(g), s.ref(), CO.dp(full))
    u0, u1 = np.load(tmp_path / "u0.npy"), np.load(tmp_path / "u1.npy")
    assert np.array_equal(u0, u1)                       # replicated result, bitwise identical on all ranks
    assert np.allclose(u0, full, rtol=1e-13, atol=1e-13 * full.max())   # differs by summation order only
    m0, m1 = np.load(tmp_path / "meta0.npy"), np.load(tmp_path / "meta1.npy")
    assert m0[0] == m1[0] == sum(range(128))            # broadcast reached rank 1
    assert m0[1] == m1[1] == 2.0                        # max-over-ranks timing rule
    assert (m0[2], m0[3], m1[2], m1[3]) == (0, 10001, 10001, 20001)
