"""SURVEY.md 8f row N4 -- DSMC (Chemistry/src/dsmc.jl).  The reference's MersenneTwister stream cannot be reproduced, so:
  * CPU: the oracle restatement is pinned by its stream-independent parts, and the device's per-cell routine -- the same
    __host__ __device__ code the kernel runs, reached through a host test hook -- is checked against them without a GPU:
    candidate-pair count and fractional carry exact, momentum and energy conserved, acceptance ratio within 4 sigma;
  * GPU: the whole perform! (cell lists + one thread per cell) on the problem/05_dsmc.jl case against the oracle."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import dsmc_oracle as D
from oracle import pic_oracle as O

SIG = np.stack([np.arange(3e6, 6.1e6, 1e6), [0.01, 0.1, 2.0, 0.01]], axis=1)     # problem/05_dsmc.jl:27


def _load(n, seed, ve=3e6):
    rng = np.random.default_rng(seed)
    x = rng.random((2, n, 2)) * 1.0
    v = rng.standard_normal((2, n, 3))
    v[0] *= ve
    v[1] *= math.sqrt(2 * 1.3806503e-23 * 300 / (8 * O.mp))
    return x, v


def test_sigma_g_max_quirk_D4():
    assert D.sigma_g_max(O.CrossSection(SIG)) == 2.0 * 5e6


def test_candidate_pairs_formula_and_carry():
    # problem/05_dsmc.jl numbers: 12 + 13 particles in a 5 cm cell, W = 1, dt = 1 ns, sigma_g_max = 1e7
    k, rem = D.candidate_pairs(12, 13, 1.0, 1.0, 0.05, 0.05, 1.03e-9, 1e7, False, 0.0)
    nc = 12 * 1.0 / 0.0025 * 13 * 1.03e-9 * 1e7 / 2.0 * 2            # = 642.72
    assert k == 642 and rem == pytest.approx(nc - 642, abs=1e-9)
    k2, rem2 = D.candidate_pairs(12, 13, 1.0, 1.0, 0.05, 0.05, 1.03e-9, 1e7, False, rem)
    assert k + k2 == 1285 and rem2 == pytest.approx(2 * nc - 1285, abs=1e-9)
    # unequal weights: Wa > Wb  =>  Pab = Wb/Wa, Pba = 1  (:101-105)
    k3, _ = D.candidate_pairs(10, 10, 4.0, 1.0, 0.1, 0.1, 1.031e-9, 1e7, True, 0.0)
    assert k3 == 824                                                  # 10*4/0.01 * 10 * 1.031e-9 * 1e7 / (0.25 + 0.25) = 824.8


def test_oracle_conserves_momentum_and_energy():
    g = O.CartesianGrid2(np.arange(6) * 0.2, np.arange(6) * 0.2)
    n = 600
    x, v = _load(n, 1)
    e = O.KineticSpecies("e-", n, -O.qe, O.me, 1.0)
    ox = O.KineticSpecies("O", n, 0.0, 8 * O.mp, 1.0)
    e.x[:], e.v[:], e.np = x[0], v[0], n
    ox.x[:], ox.v[:], ox.np = x[1], v[1], n
    d = D.DirectSimulationMonteCarlo(D.ElasticCollision(O.CrossSection(SIG), e, ox))
    p0 = e.m * e.v.sum(0) + ox.m * ox.v.sum(0)
    k0 = 0.5 * e.m * (e.v ** 2).sum() + 0.5 * ox.m * (ox.v ** 2).sum()
    rng = np.random.default_rng(2)
    tot = cand = 0
    for _ in range(5):
        nu, nc = D.perform_(d, 2e-11, g, rng)
        tot += nu.sum()
        cand += nc
    assert cand > 300 and 0.05 * cand < tot < 0.6 * cand
    p1 = e.m * e.v.sum(0) + ox.m * ox.v.sum(0)
    k1 = 0.5 * e.m * (e.v ** 2).sum() + 0.5 * ox.m * (ox.v ** 2).sum()
    assert np.abs(p1 - p0).max() <= 1e-12 * np.abs(e.m * e.v).sum() and abs(k1 - k0) <= 1e-12 * k0


def test_c_and_numpy_oracles_agree_on_the_stream_independent_parts():
    """Two independent restatements (Python FIFO-free loops, C) with different RNGs: candidate pairs and the per-cell carry are
    equal exactly on every call, the accepted fraction agrees statistically, both conserve momentum and energy."""
    from oracle import c_oracle as CO
    nxn = 6
    g = O.CartesianGrid2(np.arange(nxn) * 0.2, np.arange(nxn) * 0.2)
    cg = CO.make_grid(nxn, nxn, 0.2, 0.2)
    n = 800
    x, v = _load(n, 7)
    osp, csp = [], []
    for k, m in enumerate((O.me, 8 * O.mp)):
        o = O.KineticSpecies("s%d" % k, n, 0.0, m, 1.0)
        o.x[:], o.v[:], o.np = x[k], v[k], n
        c = CO.CSpecies(n, 0.0, m, 1.0)
        c.set(x[k][:, 0], x[k][:, 1], v[k][:, 0], v[k][:, 1], v[k][:, 2])
        osp.append(o)
        csp.append(c)
    d = D.DirectSimulationMonteCarlo(D.ElasticCollision(O.CrossSection(SIG), osp[0], osp[1]))
    gn, sg = np.ascontiguousarray(SIG[:, 0]), np.ascontiguousarray(SIG[:, 1])
    rem = np.zeros(nxn * nxn)
    fn = CO.lib().orc_dsmc_perform
    fn.restype = C.c_int64
    rng, crng = np.random.default_rng(3), CO.make_rng(4)
    co = cc = cand = 0
    k0 = 0.5 * O.me * (v[0] ** 2).sum() + 0.5 * 8 * O.mp * (v[1] ** 2).sum()
    for _ in range(6):
        nu_o, nc_o = D.perform_(d, 1e-10, g, rng)
        ncand = C.c_int64(0)
        ncoll = fn(csp[0].ref(), csp[1].ref(), C.byref(cg), CO.dp(gn), CO.dp(sg), C.c_int32(len(gn)), C.c_double(1e-10), CO.dp(rem),
                   None, C.byref(ncand), C.byref(crng))
        assert ncand.value == nc_o
        assert np.array_equal(rem.reshape((nxn, nxn), order="F"), d.collisions_remaining)
        co += nu_o.sum()
        cc += ncoll
        cand += nc_o
    pa = co / cand
    assert abs(cc / cand - pa) <= 5 * math.sqrt(2 * pa * (1 - pa) / cand), (cc / cand, pa)
    k1 = 0.5 * O.me * (csp[0].v ** 2).sum() + 0.5 * 8 * O.mp * (csp[1].v ** 2).sum()
    assert abs(k1 - k0) <= 1e-12 * k0


def _hook():
    from iskra_b200 import _lib
    fn = _lib.lib().iskb_debug_dsmc_cell
    dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    fn.argtypes = [dp, C.c_int64, dp, C.c_int64, dp, dp, u32p, C.c_uint32, u32p, C.c_uint32] + [C.c_double] * 7 + \
                  [dp, dp, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, dp, u32p, u32p]
    fn.restype = C.c_int32
    return fn


def test_device_cell_routine_on_host_pairs_carry_conservation_acceptance():
    fn = _hook()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    Na, Nb, ns, nt = 40, 35, 64, 50
    rng = np.random.default_rng(5)
    sv = np.asfortranarray(rng.standard_normal((ns, 3)) * 3e6)
    tv = np.asfortranarray(rng.standard_normal((nt, 3)) * 400.0)
    swg, twg = np.ones(ns), np.ones(nt)
    ls = rng.permutation(ns)[:Na].astype(np.uint32)
    lt = rng.permutation(nt)[:Nb].astype(np.uint32)
    ms, mt = O.me, 8 * O.mp
    gn, sg = np.ascontiguousarray(SIG[:, 0]), np.ascontiguousarray(SIG[:, 1])
    dx = dy = 0.05
    dt = 3.1e-11
    rem = C.c_double(0.0)
    p0 = ms * sv.sum(0) + mt * tv.sum(0)
    k0 = 0.5 * ms * (sv ** 2).sum() + 0.5 * mt * (tv ** 2).sum()
    tot_pairs = tot_coll = 0
    orem = 0.0
    # acceptance probability of the oracle on the same (frozen) velocities, for the statistical comparison
    gg = np.linalg.norm(sv[ls][:, None, :] - tv[lt][None, :, :], axis=2)
    p_acc = float(np.mean(np.interp(gg, gn, sg) * gg / (2.0 * 5e6)))
    for call in range(40):
        npairs, ncoll = C.c_uint32(0), C.c_uint32(0)
        rc = fn(dp(sv), ns, dp(tv), nt, dp(swg), dp(twg), up(ls), Na, up(lt), Nb, ms, mt, 1.0, 1.0, dx, dy, dt,
                dp(gn), dp(sg), len(gn), 0, 1234, call, 7, C.byref(rem), C.byref(npairs), C.byref(ncoll))
        assert rc == 0
        k, orem = D.candidate_pairs(Na, Nb, 1.0, 1.0, dx, dy, dt, 2.0 * 5e6, False, orem)
        assert npairs.value == k and rem.value == orem               # stream-independent: exact
        tot_pairs += npairs.value
        tot_coll += ncoll.value
    assert tot_pairs > 3000
    p1 = ms * sv.sum(0) + mt * tv.sum(0)
    k1 = 0.5 * ms * (sv ** 2).sum() + 0.5 * mt * (tv ** 2).sum()
    assert np.abs(p1 - p0).max() <= 1e-12 * np.abs(ms * sv).sum() and abs(k1 - k0) <= 1e-12 * k0
    # electrons are so much lighter that |g| barely changes: the acceptance probability stays that of the start
    sigma = math.sqrt(p_acc * (1 - p_acc) / tot_pairs)
    assert abs(tot_coll / tot_pairs - p_acc) <= 5 * sigma + 0.02 * p_acc, (tot_coll / tot_pairs, p_acc)
    # rows outside the cell lists were never touched
    untouched = np.setdiff1d(np.arange(ns), ls)
    assert np.array_equal(sv[untouched], np.asfortranarray(np.random.default_rng(5).standard_normal((ns, 3)) * 3e6)[untouched])


def test_device_cell_routine_skips_sparse_cells_and_same_species():
    fn = _hook()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    sv = np.asfortranarray(np.random.default_rng(0).standard_normal((8, 3)) * 4e6)
    wg = np.ones(8)
    l1 = np.array([3], dtype=np.uint32)
    l8 = np.arange(8, dtype=np.uint32)
    gn, sg = np.ascontiguousarray(SIG[:, 0]), np.ascontiguousarray(SIG[:, 1])
    rem, a, b = C.c_double(0.25), C.c_uint32(9), C.c_uint32(9)
    fn(dp(sv), 8, dp(sv), 8, dp(wg), dp(wg), up(l1), 1, up(l8), 8, 1.0, 1.0, 1.0, 1.0, 0.1, 0.1, 1e-9, dp(gn), dp(sg), 4, 0, 1, 0, 0,
       C.byref(rem), C.byref(a), C.byref(b))
    assert a.value == 0 and b.value == 0 and rem.value == 0.25         # Na < 2: continue (:111-113), carry untouched
    # same species: no factor 2 (:117-119), a pair may be (p, p) -- the reference only re-draws when source != target
    k, r = D.candidate_pairs(8, 8, 1.0, 1.0, 0.1, 0.1, 1e-10, 1e7, True, 0.0)
    rem = C.c_double(0.0)
    p0 = sv.sum(0).copy()
    fn(dp(sv), 8, dp(sv), 8, dp(wg), dp(wg), up(l8), 8, up(l8), 8, 1.0, 1.0, 1.0, 1.0, 0.1, 0.1, 1e-10, dp(gn), dp(sg), 4, 1, 1, 0, 0,
       C.byref(rem), C.byref(a), C.byref(b))
    assert a.value == k and rem.value == r
    assert np.abs(sv.sum(0) - p0).max() <= 1e-9 * np.abs(sv).sum()


# ------------------------------------------------------------------ GPU ---------------------------------
@pytest.mark.gpu
def test_dsmc_perform_on_device_vs_oracle():
    """problem/05_dsmc.jl: 21 x 21 nodes of 5 cm, e- + O, sigma(g) of :27; dt shortened so that the Python oracle finishes."""
    import iskra_b200 as ib
    PIC, CH = ib.particle_in_cell, ib.chemistry
    nx = ny = 21
    dh, dt, n = 0.05, 2e-11, 5000
    xs = np.arange(nx) * dh
    og = O.CartesianGrid2(xs, xs)
    g = ib.regular_grids.create_uniform_grid(xs, xs)
    x, v = _load(n, 3)
    osp, gsp = [], []
    for k, (name, q, m) in enumerate((("e-", -O.qe, O.me), ("O", 0.0, 8 * O.mp))):
        o = O.KineticSpecies(name, n + 8, q, m, 1.0)
        o.x[:n], o.v[:n], o.np = x[k], v[k], n
        s = PIC.create_kinetic_species(name, n + 8, q, m, 1.0)
        s.x[:n] = x[k]
        s.v[:n] = v[k]
        s.np = n
        osp.append(o)
        gsp.append(s)
    od = D.DirectSimulationMonteCarlo(D.ElasticCollision(O.CrossSection(SIG), osp[0], osp[1]))
    sig = CH.CrossSection(SIG)
    gd = CH.dsmc(CH.reactions([(sig, "e + O --> O + e")], {"e": gsp[0], "O": gsp[1]}), seed=11)
    cfg = ib.configuration.Config()
    cfg.grid, cfg.species, cfg.interactions = g, gsp, [gd]
    p0 = O.me * v[0].sum(0) + 8 * O.mp * v[1].sum(0)
    k0 = 0.5 * O.me * (v[0] ** 2).sum() + 0.5 * 8 * O.mp * (v[1] ** 2).sum()
    rng = np.random.default_rng(4)
    cand_o = cand_g = coll_o = coll_g = 0
    nu_sum = np.zeros((nx, ny))
    for it in range(6):
        nu_o, nc_o = D.perform_(od, dt, og, rng)
        nu_g, nc_g, ncoll_g = gd.perform_(None, dt, cfg)
        assert nc_g == nc_o                                       # candidate pairs: stream independent, exact every call
        assert nu_g.sum() == ncoll_g
        cand_o += nc_o
        cand_g += nc_g
        coll_o += nu_o.sum()
        coll_g += ncoll_g
        nu_sum += nu_g
    assert cand_g > 10000
    # collisions: two binomial samples of the same acceptance law
    pa = coll_o / cand_o
    sigma = math.sqrt(2 * pa * (1 - pa) / cand_o)
    assert abs(coll_g / cand_g - pa) <= 5 * sigma, (coll_g / cand_g, pa)
    assert nu_sum[nx - 1, :].sum() == 0 and nu_sum[:, ny - 1].sum() == 0       # no particle has i = nx or j = ny
    # exact invariants on the device state
    ve, vo = gsp[0].v[:n], gsp[1].v[:n]
    p1 = O.me * ve.sum(0) + 8 * O.mp * vo.sum(0)
    k1 = 0.5 * O.me * (ve ** 2).sum() + 0.5 * 8 * O.mp * (vo ** 2).sum()
    assert np.abs(p1 - p0).max() <= 1e-11 * np.abs(O.me * v[0]).sum() and abs(k1 - k0) <= 1e-11 * k0
    assert not np.array_equal(ve, v[0])
    # and through the fused step (the DSMC object is registered with the context)
    ps = ib.finite_difference_method.create_poisson_solver(g, O.eps0)
    cfg.solver, cfg.pusher = ps, PIC.create_boris_pusher()
    PIC.solve(cfg, dt, 2, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP), fused=True)
    assert gsp[0].np == n
