"""CPU tests that pin the oracle.

1. every known answer the reference holds for the path: the restated capacitor test
   (FiniteDifferenceMethod/test/runtests.jl:8-22) and the constants printed in the
   reference's notebooks (docs/capacitively_induced_discharge.ipynb, docs/nanbu-scratchbook.ipynb);
2. the two independent restatements (numpy, C) must agree bit-for-bit on the deterministic
   particle arithmetic and to rounding on the dense solve.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import pic_oracle as O


# ------------------------------------------------------------------ known answers ----------
def test_capacitor_known_answer_numpy():
    # runtests.jl:5  grid 0:0.5:1 x 0:0.5:2  (3 x 5 nodes); phi=0 on i=1, phi=1 on i=3
    grid = O.CartesianGrid2(np.arange(0, 1.01, 0.5), np.arange(0, 2.01, 0.5))
    ps = O.PoissonSolver(grid, O.eps0)
    bcs = np.zeros(grid.n, dtype=np.int8)
    bcs[0, :] = 1
    bcs[2, :] = 2
    O.apply_dirichlet(ps, bcs == 1, 0.0)
    O.apply_dirichlet(ps, bcs == 2, 1.0)
    phi = O.calculate_electric_potential(ps, np.zeros(grid.n))
    E = O.calculate_electric_field(ps, phi)
    assert np.allclose(E[:, :, 0], -np.ones((3, 5)))             # runtests.jl:20
    assert np.allclose(E[:, :, 1], np.zeros((3, 5)), atol=1e-15)  # runtests.jl:21


def _c_poisson(nx, ny, dx, dy, periodic=(), dirichlet=()):
    L = CO.lib()
    g = CO.make_grid(nx, ny, dx, dy)
    nn = nx * ny
    A = np.zeros(nn * nn)
    b = np.zeros(nn)
    dof = np.ones(nn, dtype=np.uint8)
    L.orc_poisson_assemble(C.byref(g), CO.dp(A))
    for ax in periodic:
        L.orc_poisson_apply_periodic(C.byref(g), CO.dp(A), C.c_int(ax))
    for mask, val in dirichlet:
        m = np.ascontiguousarray(mask.ravel(order="F").astype(np.uint8))
        L.orc_poisson_apply_dirichlet(C.byref(g), CO.dp(A), CO.dp(b), dof.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      m.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_double(val))
    return g, A, b, dof


def test_capacitor_known_answer_c():
    L = CO.lib()
    bcs = np.zeros((3, 5), dtype=np.int8)
    bcs[0, :] = 1
    bcs[2, :] = 2
    g, A, b, dof = _c_poisson(3, 5, 0.5, 0.5, dirichlet=[(bcs == 1, 0.0), (bcs == 2, 1.0)])
    phi = np.zeros(15)
    rc = L.orc_electric_potential(CO.dp(A), CO.dp(b), dof.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  CO.dp(np.zeros(15)), C.c_double(O.eps0), C.c_int64(15), CO.dp(phi))
    assert rc == 0
    E = np.zeros(45)
    L.orc_electric_field(C.byref(g), CO.dp(phi), CO.dp(E))
    assert np.allclose(E[:15], -1.0)
    assert np.allclose(E[15:30], 0.0, atol=1e-15)
    assert np.all(E[30:] == 0.0)


def test_notebook_constants():
    # docs/capacitively_induced_discharge.ipynb:145-163 (cell 2 output) and cell 0 output
    ms = O.me / O.QE_MCC                       # mcc.jl:26
    assert ms == 5.685630721038056e-12
    assert math.sqrt(2.0 / ms) * math.sqrt(989.6379) == 1.865794604401031e7   # mcc.jl:39-40,48
    mi = 3.99 * O.mp / O.QE_MCC
    assert mi == 4.165434669922687e-8
    assert math.sqrt(2.0 / mi) * math.sqrt(10000.0) == 692922.772032998
    # docs/nanbu-scratchbook.ipynb cell 1 output (He+ with m = 6.64647645970479e-27)
    assert math.sqrt(2.0 / (6.64647645970479e-27 / O.QE_MCC)) * math.sqrt(10000.0) == 694343.5981451317
    # "kB Te / me: 476807.16512475203" = .5thermal_speed(Te, me)   (configuration.jl:79-81)
    assert 0.5 * O.thermal_speed(30000.0, O.me) == 476807.16512475203
    # printed run parameters of the RF case
    f = 13.56 * 1e6
    dt = 1 / (400 * f)
    assert dt == 1.8436578171091445e-10
    dh = 6.7 * 0.01 / 128
    assert dh == 0.0005234375
    vol = 128 * dh * 1 * dh
    assert 2.56e14 * vol / (128 * 128) == 547973.6328125001
    assert 9.64e20 * vol / (128 * 128) == 2.0634632110595708e12


def test_notebook_candidate_counts():
    # "Collisions: 5.0 out of 1037.0" / "4.0 out of 159.0" at step 1 (np = 16384; N = 4 / 2)
    # with the printed max_sigma_g -- pins the formula of mcc.jl:242-248.
    dt = 1.8436578171091445e-10
    for sg, N, expect in ((8.976965143603543e-14, 4, 1037), (2.7462885393092625e-14, 2, 159)):
        max_Pt = 1.0 - math.exp(-9.64e20 * sg * dt)
        frac, Nc = math.modf(N * max_Pt * 16384 + 0.0)
        assert int(Nc) == expect


def test_mcc_candidate_count_and_remainder_both_oracles():
    # flat synthetic cross-section chosen so that max_sigma_g equals the notebook constant:
    # sigma * v_max = 8.976965143603543e-14 at eps_max
    ms = O.me / O.QE_MCC
    epsmax = 100.0
    vmax = math.sqrt(2.0 / ms) * math.sqrt(epsmax)
    sig = 8.976965143603543e-14 / vmax / 4
    nodes = np.array([[0.0, sig], [epsmax, sig]])
    e = O.KineticSpecies("e-", 20000, -O.qe, O.me, 547973.6328125001)
    rng = np.random.default_rng(0)
    e.np = 16384
    e.x[:e.np] = rng.random((e.np, 2)) * [0.067, 0.0005234375]
    e.v[:e.np] = rng.standard_normal((e.np, 3)) * 1e5
    grid = O.CartesianGrid2(np.arange(129) * 0.0005234375, np.arange(2) * 0.0005234375)
    He = O.FluidSpecies("He", 1.0, 0.0, 3.99 * O.mp, 9.64e20 * np.ones((129, 2)), 300.0)
    colls = [O.Collision(O.ELASTIC_ISOTROPIC, O.CrossSection(nodes), e, He) for _ in range(4)]
    mcc = O.MonteCarloCollisions(colls)
    assert mcc.max_sigma_g == pytest.approx(8.976965143603543e-14, rel=1e-14)
    E = np.zeros((129, 2, 3))
    nu, Nc, ncoll = O.mcc_perform_(mcc, E, 1.8436578171091445e-10, grid, rng)
    assert Nc == 1037 and 0 <= ncoll <= Nc
    assert mcc.remainder == pytest.approx(0.3060804980123, abs=1e-6)
    assert nu.sum() == ncoll
    # C oracle, same set-up
    cs = CO.CSpecies(20000, -O.qe, O.me, 547973.6328125001)
    cs.set(e.x[:16384, 0], e.x[:16384, 1], e.v[:16384, 0], e.v[:16384, 1], e.v[:16384, 2])
    cm = CO.CMcc(cs, [(CO_KIND["iso"], 0.0, nodes[:, 0], nodes[:, 1], None)] * 4, 0.0, 3.99 * O.mp, 300.0,
                 9.64e20 * np.ones(129 * 2))
    assert cm.c.max_sigma_g == mcc.max_sigma_g
    assert cm.c.m_eV == mcc.m
    g = CO.make_grid(129, 2, 0.0005234375, 0.0005234375)
    rc, nu_c, Nc_c, ncoll_c = cm.perform(g, np.zeros(129 * 2 * 3), 1.8436578171091445e-10, CO.make_rng(1))
    assert rc == 0 and Nc_c == 1037 and nu_c.sum() == ncoll_c
    assert cm.c.remainder == pytest.approx(mcc.remainder, abs=1e-12)


CO_KIND = {"iso": 0, "back": 1, "inel": 2, "exc": 3, "ion": 4}


# ------------------------------------------------------ numpy vs C, bit-for-bit -------------
def _random_species(n, cap, nx, ny, dx, dy, seed, q=-O.qe, m=O.me, w=3.5e7, spill=0.0):
    rng = np.random.default_rng(seed)
    Lx, Ly = (nx - 1) * dx, (ny - 1) * dy
    x = rng.random(n) * Lx * (1 + 2 * spill) - spill * Lx
    y = rng.random(n) * Ly * (1 + 2 * spill) - spill * Ly
    v = rng.standard_normal((n, 3)) * 1e6
    wg = w * (0.5 + rng.random(n))
    po = O.KineticSpecies("s", cap, q, m, w)
    po.np = n
    po.x[:n, 0], po.x[:n, 1], po.v[:n], po.wg[:n] = x, y, v, wg
    pc = CO.CSpecies(cap, q, m, w)
    pc.set(x, y, v[:, 0], v[:, 1], v[:, 2], wg)
    return po, pc


def _colmajor3(E):
    return np.ascontiguousarray(E.transpose(2, 1, 0)).ravel()


@pytest.mark.parametrize("nx,ny", [(129, 2), (33, 65), (17, 9)])
def test_cell_gather_push_deposit_bitwise(nx, ny):
    L = CO.lib()
    dx = dy = 1.8743613985989574e-08
    n = 5000
    po, pc = _random_species(n, n + 10, nx, ny, dx, dy, seed=nx * 7 + ny)
    # adversarial positions: exact nodes, one ulp below / above nodes
    k = np.arange(50)
    po.x[k, 0] = (k % (nx - 1)) * dx
    po.x[50:100, 0] = np.nextafter((k % (nx - 1) + 1) * dx, 0.0)
    po.x[100:150, 0] = np.nextafter((k % (nx - 2)) * dx, 1.0)
    pc.xy[0, :n] = po.x[:n, 0]
    grid = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dy)
    g = CO.make_grid(nx, ny, dx, dy)
    i, j, hx, hy = O.particle_cell(po.x[:n], grid.dh)
    ci, cj = np.zeros(n, np.int64), np.zeros(n, np.int64)
    chx, chy = np.zeros(n), np.zeros(n)
    L.orc_particle_cell(CO.dp(pc.xy[0]), CO.dp(pc.xy[1]), C.c_int64(n), C.c_double(dx), C.c_double(dy),
                        ci.ctypes.data_as(CO.c_i64p), cj.ctypes.data_as(CO.c_i64p), CO.dp(chx), CO.dp(chy))
    assert np.array_equal(i, ci) and np.array_equal(j, cj)
    assert np.array_equal(hx, chx) and np.array_equal(hy, chy)
    assert i.min() >= 1 and i.max() <= nx - 1 and j.min() >= 1 and j.max() <= ny - 1
    # gather
    rng = np.random.default_rng(5)
    E = rng.standard_normal((nx, ny, 3)) * 1e5
    E[:, :, 2] = 0.0
    pE = O.grid_to_particle(grid, po, E)
    Ec = _colmajor3(E)
    pEc = np.zeros(3 * n)
    L.orc_gather(C.byref(g), pc.ref(), CO.dp(Ec), CO.dp(pEc))
    assert np.array_equal(pE.T.ravel(), pEc)
    # push
    dt = 5.3e-16
    O.push_in_cartesian_(po, pE, dt)
    L.orc_push(pc.ref(), CO.dp(pEc), C.c_double(dt))
    assert np.array_equal(po.x[:n, 0], pc.xy[0, :n]) and np.array_equal(po.x[:n, 1], pc.xy[1, :n])
    assert np.array_equal(po.v[:n].T, pc.v[:, :n])
    # wrap both dims, then deposit
    O.wrap_(po, grid)
    L.orc_wrap(pc.ref(), C.byref(g), C.c_int(1))
    L.orc_wrap(pc.ref(), C.byref(g), C.c_int(2))
    assert np.array_equal(po.x[:n, 0], pc.xy[0, :n]) and np.array_equal(po.x[:n, 1], pc.xy[1, :n])
    assert po.x[:n, 0].min() >= 0 and po.x[:n, 0].max() <= (nx - 1) * dx
    u = O.particle_to_grid(po, grid, po.wg[:n])
    uc = np.zeros(nx * ny)
    L.orc_deposit(C.byref(g), pc.ref(), CO.dp(uc))
    assert np.array_equal(u.ravel(order="F"), uc)
    assert u.sum() == pytest.approx(po.wg[:n].sum(), rel=1e-12)   # CIC conserves weight
    V = O.cell_volume(grid)
    Vc = np.zeros(nx * ny)
    L.orc_cell_volume(C.byref(g), CO.dp(Vc))
    assert np.array_equal(V.ravel(order="F"), Vc)


def test_discard_matches_and_keeps_id_permutation():
    L = CO.lib()
    nx, ny, dx = 33, 9, 1.25e-3
    n, cap = 4000, 4100
    po, pc = _random_species(n, cap, nx, ny, dx, dx, seed=3, spill=0.2)
    grid = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    g = CO.make_grid(nx, ny, dx, dx)
    r1 = O.discard_(po, grid, dims=(1,))
    r2 = L.orc_discard(pc.ref(), C.byref(g), C.c_int(1))
    assert r1 == r2 and po.np == pc.np and 0 < r1 < n
    m = po.np
    assert np.array_equal(po.x[:m, 0], pc.xy[0, :m]) and np.array_equal(po.v[:m].T, pc.v[:, :m])
    assert np.array_equal(po.id, pc.id) and np.array_equal(po.wg, pc.wg)
    assert sorted(po.id.tolist()) == list(range(1, cap + 1))       # id stays a permutation
    Lx = (nx - 1) * dx
    assert po.x[:m, 0].min() >= 0 and po.x[:m, 0].max() < Lx
    O.wrap_(po, grid, dims=(2,))
    L.orc_wrap(pc.ref(), C.byref(g), C.c_int(2))
    assert np.array_equal(po.x[:m, 1], pc.xy[1, :m])


def test_fld_edge_cases():
    L = 2.4e-6
    x = np.array([0.0, -0.0, L, np.nextafter(L, 0), -1e-30, 2.5 * L, -0.25 * L, -L, 7.5 * L])
    a = O.jl_fld(x, L)
    assert a.tolist() == [0, 0, 1, 0, -1, 2, -1, -1, 7]
    # tiny negative wraps to exactly L (reference quirk; next particle_cell would go out of bounds)
    assert (-1e-30) - (-1.0) * L == L


@pytest.mark.parametrize("periodic,dirichlet_edges", [((1, 2), ()), ((1,), ("l", "r")), ((), ("l",)), ((2,), ("b",))])
def test_dense_operator_and_solve_numpy_vs_c(periodic, dirichlet_edges):
    L = CO.lib()
    nx, ny, dx = 9, 6, 0.37
    grid = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = O.PoissonSolver(grid, O.eps0)
    masks = []
    for e in dirichlet_edges:
        m = np.zeros((nx, ny), bool)
        if e == "l":
            m[0, :] = True
        if e == "r":
            m[nx - 1, :] = True
        if e == "b":
            m[:, 0] = True
        masks.append((m, 1.5 if e == "l" else -0.5))
    for ax in periodic:
        O.apply_periodic(ps, ax)
    for m, v in masks:
        O.apply_dirichlet(ps, m, v)
    g, A, b, dof = _c_poisson(nx, ny, dx, dx, periodic=periodic, dirichlet=masks)
    assert np.array_equal(ps.A.ravel(order="F"), A)
    rng = np.random.default_rng(1)
    rho = rng.standard_normal((nx, ny)) * 1e-9
    if not dirichlet_edges:
        return  # singular (fully periodic / pure Neumann): solution ill-posed, see DESIGN.md (H3)
    phi = O.calculate_electric_potential(ps, -rho)
    phic = np.zeros(nx * ny)
    rc = L.orc_electric_potential(CO.dp(A), CO.dp(b), dof.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  CO.dp(rho.ravel(order="F").copy()), C.c_double(O.eps0), C.c_int64(nx * ny),
                                  CO.dp(phic))
    assert rc == 0
    assert np.allclose(phi.ravel(order="F"), phic, rtol=1e-11, atol=1e-11 * np.abs(phi).max())
    E = O.calculate_electric_field(ps, phi)
    Ec = np.zeros(3 * nx * ny)
    L.orc_electric_field(C.byref(g), CO.dp(np.ascontiguousarray(phi.ravel(order="F"))), CO.dp(Ec))
    assert np.array_equal(_colmajor3(E), Ec)


def test_cross_section_interpolation():
    xs = np.array([0.0, 1.0, 2.5, 10.0])
    ys = np.array([1.0, 3.0, 2.0, 8.0])
    s = O.CrossSection(np.stack([xs, ys], 1))
    L = CO.lib()
    q = np.array([-1.0, 0.0, 0.5, 1.0, 1.75, 2.5, 9.9, 10.0, 1e6])
    expect = np.array([1.0, 1.0, 2.0, 3.0, 2.5, 2.0, None, 8.0, 8.0], dtype=object)
    got = s(q)
    for k, e in enumerate(expect):
        c = L.orc_xsec_eval(CO.dp(xs), CO.dp(ys), 4, float(q[k]))
        assert c == got[k]
        if e is not None:
            assert got[k] == e


def test_add_and_remove_particles_restatement():
    """kinetic.jl:29-50: add! appends x, v only; remove_particles! re-tests the row swapped in, keeps id a permutation."""
    g = O.CartesianGrid2(np.arange(5) * 1.0, np.arange(4) * 1.0)
    a = O.KineticSpecies("a", 10, -1.0, 1.0, 3.0)
    b = O.KineticSpecies("b", 10, -1.0, 1.0, 5.0)
    a.x[:3] = [[0.5, 0.5], [1.5, 0.5], [3.5, 2.5]]
    a.v[:3] = [[1, 0, 0], [2, 0, 0], [3, 0, 0]]
    a.np = 3
    b.x[:2] = [[2.5, 1.5], [3.5, 2.6]]
    b.np = 2
    O.add_(a, b)
    assert b.np == 5 and np.array_equal(b.x[2:5], a.x[:3]) and np.array_equal(b.v[2:5, 0], [1, 2, 3])
    assert np.all(b.wg == 5.0) and np.array_equal(b.id, np.arange(1, 11))
    # remove everything in cell (4,3): rows 2 and 5 (1-based) -- row 5 is swapped into row 2 first and re-tested
    O.remove_particles_(b, g.dh, lambda i, j: (i, j) == (4, 3))
    assert b.np == 3
    assert sorted(map(tuple, b.x[:3].tolist())) == [(0.5, 0.5), (1.5, 0.5), (2.5, 1.5)]
    assert sorted(b.id.tolist()) == list(range(1, 11))
