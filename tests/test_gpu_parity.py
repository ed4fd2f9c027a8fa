"""GPU parity tests: the CUDA path (through the C ABI / the host mirror) against the oracle.

Bars (north_star): cell indices and sort permutations bit-exact; gather / push / wrap are
bit-exact as well because the kernels keep the reference's operation order without FMA
contraction; rho / phi / E and particle state after many steps within 1e-10 relative;
RNG-dependent results statistically.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import pic_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-10   # tolerance stated by north_star for deterministic rho/phi/E and particle state


@pytest.fixture(scope="module")
def ib():
    import iskra_b200
    return iskra_b200


def _species_pair(ib, grid, n, cap, seed, q=-O.qe, m=O.me, w=3.5e7, spill=0.0, vscale=1e6):
    nx, ny = grid.n
    dx, dy = grid.dh
    rng = np.random.default_rng(seed)
    Lx, Ly = (nx - 1) * dx, (ny - 1) * dy
    x = rng.random(n) * Lx * (1 + 2 * spill) - spill * Lx
    y = rng.random(n) * Ly * (1 + 2 * spill) - spill * Ly
    v = rng.standard_normal((n, 3)) * vscale
    wg = w * (0.5 + rng.random(n))
    pc = CO.CSpecies(cap, q, m, w)
    pc.set(x, y, v[:, 0], v[:, 1], v[:, 2], wg)
    pg = ib.particle_in_cell.create_kinetic_species("s", cap, q, m, w)
    pg.x[:n, 0], pg.x[:n, 1] = x, y
    pg.v[:n] = v
    pg.wg[:n] = wg
    pg.np = n
    return pc, pg


def _grid_pair(ib, nx, ny, dx, dy=None):
    dy = dx if dy is None else dy
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dy)
    cg = CO.make_grid(nx, ny, g.dh[0], g.dh[1])
    return g, cg


def _colmajor3(E):
    return np.ascontiguousarray(np.asarray(E).transpose(2, 1, 0)).ravel()


def _by_id(ids, *cols):
    o = np.argsort(ids, kind="stable")
    return [np.asarray(c)[o] for c in cols]


# -------------------------------------------------------------------------- single operators ---
@pytest.mark.parametrize("nx,ny", [(129, 2), (33, 65), (257, 129)])
def test_cell_index_gather_push_wrap_bitexact(ib, nx, ny):
    PIC = ib.particle_in_cell
    Lc = CO.lib()
    dx = 1.8743613985989574e-08
    g, cg = _grid_pair(ib, nx, ny, dx)
    n = 20000
    pc, pg = _species_pair(ib, g, n, n + 7, seed=nx + ny)
    k = np.arange(64)
    xs = pg.x
    xs[k, 0] = (k % (nx - 1)) * dx                                  # exactly on nodes
    xs[64 + k, 0] = np.nextafter((k % (nx - 1) + 1) * dx, 0.0)      # one ulp below a node
    xs[128 + k, 0] = np.nextafter((k % (nx - 2)) * dx, 1.0)         # one ulp above a node
    pc.xy[0, :n] = xs[:n, 0]
    ci, cj = np.zeros(n, np.int64), np.zeros(n, np.int64)
    chx, chy = np.zeros(n), np.zeros(n)
    Lc.orc_particle_cell(CO.dp(pc.xy[0]), CO.dp(pc.xy[1]), C.c_int64(n), C.c_double(dx), C.c_double(dx),
                         ci.ctypes.data_as(CO.c_i64p), cj.ctypes.data_as(CO.c_i64p), CO.dp(chx), CO.dp(chy))
    i, j, hx, hy = PIC.particle_cell(pg, g)
    assert np.array_equal(i, ci) and np.array_equal(j, cj)          # bit-exact contract
    assert np.array_equal(hx, chx) and np.array_equal(hy, chy)
    # gather
    rng = np.random.default_rng(11)
    E = rng.standard_normal((nx, ny, 3)) * 1e5
    E[:, :, 2] = 0.0
    pEc = np.zeros(3 * n)
    Lc.orc_gather(C.byref(cg), pc.ref(), CO.dp(_colmajor3(E)), CO.dp(pEc))
    pE = PIC.grid_to_particle(g, pg, E)
    assert np.array_equal(pE.ravel(order="F"), pEc)
    # push with the host-provided per-particle field (push_particles! signature) ...
    dt = 5.3e-16
    Lc.orc_push(pc.ref(), CO.dp(pEc), C.c_double(dt))
    PIC.push_particles_(None, pg, pE, None, dt, g)
    assert np.array_equal(pg.x[:n, 0], pc.xy[0, :n]) and np.array_equal(pg.x[:n, 1], pc.xy[1, :n])
    assert np.array_equal(pg.v[:n].T, pc.v[:, :n])
    # ... and fused gather+push from the device field
    pEc2 = np.zeros(3 * n)
    Lc.orc_wrap(pc.ref(), C.byref(cg), C.c_int(1))
    Lc.orc_wrap(pc.ref(), C.byref(cg), C.c_int(2))
    PIC.wrap_(pg, g)
    assert np.array_equal(pg.x[:n, 0], pc.xy[0, :n]) and np.array_equal(pg.x[:n, 1], pc.xy[1, :n])
    Lc.orc_gather(C.byref(cg), pc.ref(), CO.dp(_colmajor3(E)), CO.dp(pEc2))
    Lc.orc_push(pc.ref(), CO.dp(pEc2), C.c_double(dt))
    PIC.push_particles_(None, pg, None, None, dt, g)
    assert np.array_equal(pg.x[:n, 0], pc.xy[0, :n]) and np.array_equal(pg.v[:n].T, pc.v[:, :n])


def test_discard_and_wrap_keyed_by_id(ib):
    PIC = ib.particle_in_cell
    Lc = CO.lib()
    nx, ny, dx = 33, 17, 1.25e-3
    g, cg = _grid_pair(ib, nx, ny, dx)
    n, cap = 30000, 30100
    pc, pg = _species_pair(ib, g, n, cap, seed=5, spill=0.15)
    r_ref = Lc.orc_discard(pc.ref(), C.byref(cg), C.c_int(1))
    Lc.orc_wrap(pc.ref(), C.byref(cg), C.c_int(2))
    r_gpu = PIC.discard_(pg, g, dims=[1])
    PIC.wrap_(pg, g, dims=[2])
    assert r_gpu == r_ref and pg.np == pc.np and 0 < r_ref < n
    m = pc.np
    ids_g, ids_c = pg.id[:m].copy(), pc.id[:m].copy()
    assert sorted(pg.id.tolist()) == list(range(1, cap + 1))        # id stays a permutation (H5)
    xg, yg, vg = _by_id(ids_g, pg.x[:m, 0], pg.x[:m, 1], pg.v[:m, 0])
    xc, yc, vc = _by_id(ids_c, pc.xy[0, :m], pc.xy[1, :m], pc.v[0, :m])
    assert np.array_equal(np.sort(ids_g), np.sort(ids_c))
    assert np.array_equal(xg, xc) and np.array_equal(yg, yc) and np.array_equal(vg, vc)
    wg_g, = _by_id(ids_g, pg.wg[:m])
    wg_c, = _by_id(ids_c, pc.wg[:m])
    assert np.array_equal(wg_g, wg_c)


@pytest.mark.parametrize("nx,ny,n", [(129, 2, 5000), (33, 65, 40000), (65, 65, 300000)])
def test_density_matches_oracle(ib, nx, ny, n):
    PIC = ib.particle_in_cell
    Lc = CO.lib()
    g, cg = _grid_pair(ib, nx, ny, 2.5e-4)
    pc, pg = _species_pair(ib, g, n, n, seed=n)
    V = np.zeros(nx * ny)
    Lc.orc_cell_volume(C.byref(cg), CO.dp(V))
    assert np.array_equal(ib.regular_grids.cell_volume(g).ravel(order="F"), V)
    dens = np.zeros(nx * ny)
    Lc.orc_density(C.byref(cg), pc.ref(), CO.dp(V), CO.dp(dens))
    got = PIC.density(pg, g).ravel(order="F")
    assert np.allclose(got, dens, rtol=1e-12, atol=1e-12 * dens.max())


def _cell_key(i, j, nx, ny):
    """The sort key documented in DESIGN.md / sort.cu (8x8-cell tiles inside 16x16-tile meta-tiles)."""
    cx, cy = i - 1, j - 1
    tiles_x = (nx - 1 + 7) // 8
    mtx = (tiles_x + 15) // 16
    tx, ty = cx >> 3, cy >> 3
    tile = (((ty >> 4) * mtx + (tx >> 4)) << 8) | ((ty & 15) << 4) | (tx & 15)
    return (tile << 6) | ((cy & 7) << 3) | (cx & 7)


@pytest.mark.parametrize("nx,ny,n", [(129, 2, 3000), (65, 33, 100000), (257, 257, 1000000)])
def test_sort_permutation_bitexact(ib, nx, ny, n):
    PIC = ib.particle_in_cell
    g, cg = _grid_pair(ib, nx, ny, 1e-3)
    pc, pg = _species_pair(ib, g, n, n + 5, seed=n + 1)
    i, j, _, _ = O.particle_cell(np.stack([pc.xy[0, :n], pc.xy[1, :n]], 1), g.dh)
    expect = np.argsort(_cell_key(i, j, nx, ny), kind="stable")
    x_before = pg.x[:n, 0].copy()
    id_before = pg.id.copy()
    perm = PIC.sort_by_cell_(pg, g)
    assert np.array_equal(perm.astype(np.int64), expect)            # bit-exact contract
    assert np.array_equal(pg.x[:n, 0], x_before[expect])
    assert np.array_equal(pg.id[:n], id_before[:n][expect])
    assert np.array_equal(pg.id[n:], id_before[n:])
    i2, j2, _, _ = PIC.particle_cell(pg, g)
    k2 = _cell_key(i2.astype(np.int64), j2.astype(np.int64), nx, ny)
    assert np.all(np.diff(k2) >= 0)                                 # sortedness


@pytest.mark.parametrize("nx,ny,n", [(65, 33, 100000), (257, 257, 1500000), (17, 17, 200000)])
def test_deposit_layout_permutation_bitexact(ib, nx, ny, n):
    """Row order kept by the fused step: cell sort, then round-robin over the cells of each 8x8
    tile in checkerboard order pi(c): order by (tile, rank-in-cell, pi(cell)) for ranks < 512, the
    rest behind in (cell, rank) order."""
    PIC = ib.particle_in_cell
    g, cg = _grid_pair(ib, nx, ny, 1e-3)
    pc, pg = _species_pair(ib, g, n, n + 5, seed=n + 3)
    if nx == 17:                                      # pile-up in one cell: exercises ranks >= 512
        pg.x[:5000, 0] = 2.5e-3
        pg.x[:5000, 1] = 3.5e-3
        pc.xy[0, :5000], pc.xy[1, :5000] = 2.5e-3, 3.5e-3
    i, j, _, _ = O.particle_cell(np.stack([pc.xy[0, :n], pc.xy[1, :n]], 1), g.dh)
    key = _cell_key(i, j, nx, ny)
    srt = np.argsort(key, kind="stable")
    ks = key[srt]
    first = np.r_[True, ks[1:] != ks[:-1]]
    run_start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    rank = np.arange(n) - run_start
    tile, cell = ks >> 6, ks & 63
    capped = np.minimum(rank, 512)
    order = np.lexsort((np.where(capped < 512, 0, rank), cell, capped, tile)) if False else None
    # rows with rank < 512: (tile, rank, cell); rows with rank >= 512: (tile, 512, cell, rank)
    sec = np.where(rank < 512, rank, 512)
    pi = ((((cell & 7) + (cell >> 3)) & 1) << 5) | (cell >> 1)
    order = np.lexsort((np.where(rank < 512, 0, rank), np.where(rank < 512, pi, cell), sec, tile))
    expect = srt[order]
    perm = PIC.sort_by_cell_(pg, g, for_deposit=True)
    assert np.array_equal(perm.astype(np.int64), expect)
    # property: the first rows of every populated tile hit distinct cells
    i2, j2, _, _ = PIC.particle_cell(pg, g)
    k2 = _cell_key(i2.astype(np.int64), j2.astype(np.int64), nx, ny)
    assert np.all(np.diff(k2 >> 6) >= 0)
    t0 = np.flatnonzero(np.r_[True, (k2[1:] >> 6) != (k2[:-1] >> 6)])
    for s0 in t0[:200]:
        t = k2[s0] >> 6
        seg = k2[s0:s0 + 64]
        seg = seg[(seg >> 6) == t]
        ncells = len(np.unique(k2[(k2 >> 6) == t] & 63)) if len(t0) < 300 else None
        head = seg[:min(len(seg), ncells)] if ncells else seg[:8]
        assert len(np.unique(head & 63)) == len(head)


def test_fast_division_matches_ieee_on_adversarial_positions(ib):
    """cell1 uses q = RN(x*r); rem = fma(-q,d,x); q' = fma(rem,r,q): must equal IEEE x/d for every
    input (bit-exact cell index contract).  Positions one/two ulps around 4096 cell edges."""
    PIC = ib.particle_in_cell
    for dx in (1.8743613985989574e-08, 5.234375e-4, 1.25e-3, 1.0 / 3.0):
        nx = 4097
        g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(17) * dx)
        k = np.arange(1, nx - 1, dtype=np.float64)
        edges = k * dx
        xs = [edges]
        lo, hi = edges.copy(), edges.copy()
        for _ in range(3):
            lo, hi = np.nextafter(lo, 0.0), np.nextafter(hi, 1e9)
            xs += [lo.copy(), hi.copy()]
        rng = np.random.default_rng(0)
        xs.append(rng.random(200000) * (nx - 1) * dx)
        x = np.concatenate(xs)
        x = x[(x >= 0) & (x < (nx - 1) * dx)]
        n = len(x)
        pg = PIC.create_kinetic_species("s", n, -O.qe, O.me, 1.0)
        pg.x[:n, 0] = x
        pg.x[:n, 1] = rng.random(n) * 16 * dx
        pg.np = n
        i, j, hx, hy = PIC.particle_cell(pg, g)
        f = 1.0 + x / dx
        assert np.array_equal(i, np.floor(f).astype(np.int32))
        assert np.array_equal(hx, f - np.floor(f))


# ------------------------------------------------------------------------------ field solve ---
def _oracle_poisson(nx, ny, dx, periodic, edges):
    grid = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = O.PoissonSolver(grid, O.eps0)
    for ax in periodic:
        O.apply_periodic(ps, ax)
    for name, val in edges:
        m = np.zeros((nx, ny), bool)
        if name == "l":
            m[0, :] = True
        elif name == "r":
            m[nx - 1, :] = True
        elif name == "b":
            m[:, 0] = True
        else:
            m[:, ny - 1] = True
        O.apply_dirichlet(ps, m, val)
    return grid, ps


def _gpu_poisson(ib, nx, ny, dx, periodic, edges):
    FDM = ib.finite_difference_method
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = FDM.create_poisson_solver(g, O.eps0)
    for ax in periodic:
        FDM.apply_periodic(ps, ax)
    for name, val in edges:
        m = np.zeros((nx, ny), bool)
        if name == "l":
            m[0, :] = True
        elif name == "r":
            m[nx - 1, :] = True
        elif name == "b":
            m[:, 0] = True
        else:
            m[:, ny - 1] = True
        FDM.apply_dirichlet(ps, m, val)
    return g, ps


POISSON_CASES = [
    # nx, ny, periodic axes, dirichlet edges          (reference configs first)
    (129, 2, (1,), (("l", 0.0), ("r", 0.0))),         # C2  11_rf_discharge.jl:76-78
    (129, 2, (1,), (("l", 37.5), ("r", 0.0))),        # C2 with the RF electrode driven
    (33, 65, (1,), (("l", 0.0), ("r", 200.0))),       # C3  12_avalanche.jl:51-53
    (3, 5, (), (("l", 0.0), ("r", 1.0))),             # capacitor test, runtests.jl:8-22
    (17, 12, (), (("l", 1.0),)),
    (17, 12, (), (("r", -2.0),)),
    (12, 17, (2,), (("b", 0.5),)),
    (12, 17, (2,), (("b", 0.5), ("t", -1.0))),
    (20, 9, (), (("l", 1.0), ("b", 2.0))),
    (20, 9, (), (("l", 1.0), ("r", 3.0), ("b", 2.0), ("t", -1.0))),
    (16, 16, (1,), (("r", 3.0),)),
]


@pytest.mark.parametrize("nx,ny,periodic,edges", POISSON_CASES)
def test_poisson_separable_vs_dense_oracle(ib, nx, ny, periodic, edges):
    FDM = ib.finite_difference_method
    dx = 5.234375e-4
    grid, ops = _oracle_poisson(nx, ny, dx, periodic, edges)
    g, ps = _gpu_poisson(ib, nx, ny, dx, periodic, edges)
    assert ps.mode == "separable"
    A, b = ps.dense()
    assert np.array_equal(A, ops.A)                                 # same operator, bit for bit
    rng = np.random.default_rng(nx * ny)
    rho = rng.standard_normal((nx, ny)) * 1e-7
    phi_lu = O.calculate_electric_potential(ops, -rho)              # the reference's one-shot A\\b
    phi_ref = _refined_solve(ops.A, ops.b, dx).reshape((nx, ny), order="F")
    E_ref = O.calculate_electric_field(ops, phi_ref)
    phi = FDM.calculate_electric_potential(ps, -rho)
    E = FDM.calculate_electric_field(ps, phi)
    sc = np.abs(phi_ref).max()
    assert np.abs(phi - phi_ref).max() <= REL * sc
    assert np.abs(E - E_ref).max() <= REL * np.abs(E_ref).max()
    assert np.all(E[:, :, 2] == 0)
    # the reference's own LU sits within its (measured) rounding noise of both
    lu_noise = np.abs(phi_lu - phi_ref).max() / sc
    assert lu_noise <= 1e-6
    assert np.abs(phi - phi_lu).max() <= (lu_noise + REL) * sc


@pytest.mark.parametrize("nx,ny,periodic", [(129, 2, (1, 2)), (24, 17, (1, 2)), (17, 24, (1,)), (20, 20, ())])
def test_poisson_singular_cases_gauge_free(ib, nx, ny, periodic):
    """Fully periodic / all-open operators are singular: the reference's LU result
    is polluted by the null vector.  Parity is defined on E with the constant mode of the rhs
    projected out; phi is compared after removing its mean."""
    FDM = ib.finite_difference_method
    dx = 1.8743613985989574e-08
    grid, ops = _oracle_poisson(nx, ny, dx, periodic, ())
    g, ps = _gpu_poisson(ib, nx, ny, dx, periodic, ())
    assert ps.mode == "separable"
    assert np.array_equal(ps.dense()[0], ops.A)
    rng = np.random.default_rng(3)
    rho = rng.standard_normal((nx, ny)) * 1e-3
    b = (-rho).ravel(order="F") / O.eps0                            # b = f/eps0 with f = -rho (:375)
    bp = b - b.mean()
    x = np.linalg.lstsq(ops.A, bp, rcond=None)[0]
    x -= x.mean()
    phi_ref = x.reshape((nx, ny), order="F")
    E_ref = O.calculate_electric_field(ops, phi_ref)
    phi = FDM.calculate_electric_potential(ps, -rho)
    E = FDM.calculate_electric_field(ps, phi)
    assert abs(phi.mean()) <= 1e-9 * np.abs(phi).max()
    assert np.abs(phi - phi_ref).max() <= 1e-8 * np.abs(phi_ref).max()
    assert np.abs(E - E_ref).max() <= 1e-8 * np.abs(E_ref).max()
    # the residual of the projected system is at rounding level
    r = ops.A @ phi.ravel(order="F") - bp
    assert np.abs(r).max() <= 1e-9 * np.abs(bp).max()


def test_poisson_dense_fallback_irregular_mask(ib):
    FDM = ib.finite_difference_method
    nx, ny, dx = 14, 11, 1e-3
    grid = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    ops = O.PoissonSolver(grid, O.eps0)
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = FDM.create_poisson_solver(g, O.eps0)
    m1 = np.zeros((nx, ny), bool)
    m1[0, :] = True
    m2 = np.zeros((nx, ny), bool)
    m2[5:8, 4:6] = True                                             # an internal electrode
    for mask, val in ((m1, 0.0), (m2, 12.0)):
        O.apply_dirichlet(ops, mask, val)
        FDM.apply_dirichlet(ps, mask, val)
    assert ps.mode == "dense"
    rho = np.random.default_rng(0).standard_normal((nx, ny)) * 1e-8
    O.calculate_electric_potential(ops, -rho)
    phi_ref = _refined_solve(ops.A, ops.b, dx).reshape((nx, ny), order="F")
    phi = FDM.calculate_electric_potential(ps, -rho)
    assert np.abs(phi - phi_ref).max() <= REL * np.abs(phi_ref).max()
    assert np.all(phi[m2] == 12.0)


# ------------------------------------------------------------------------- multi-step parity ---
def _refined_solve(A, b, dx):
    """Exact solution of the reference's linear system A x = b: LU of the row-equilibrated matrix
    plus iterative refinement with extended-precision residuals.

    Why the pin is not the one-shot LU: the reference's A mixes identity rows (Dirichlet, scale 1)
    with stencil rows of scale 1/dh^2 (generalized_poisson.jl:65, :210-211), so `A\\b` (dgetrf without
    equilibration) returns phi with a relative error of 2e-10 .. 2e-8 on these grids
    (test_reference_lu_noise_level measures it).  That noise depends on LAPACK's pivot order and
    cannot be reproduced by any second implementation, so phi/E parity at the 1e-10 bar of
    north_star is checked against the exact solution of the same system, and the one-shot LU is
    checked at its own noise level."""
    import scipy.linalg as sla
    d = np.where((np.abs(np.diag(A)) == 1.0) & (np.count_nonzero(A, axis=1) == 1), 1.0, dx * dx)
    As, bs = A * d[:, None], b * d
    lu = sla.lu_factor(As)
    x = sla.lu_solve(lu, bs)
    Al = As.astype(np.longdouble)
    for _ in range(3):
        r = (bs.astype(np.longdouble) - Al @ x.astype(np.longdouble)).astype(np.float64)
        x = x + sla.lu_solve(lu, r)
    return x


class _OracleSolver:
    """Field solve of the oracle operator for the multi-step tests.  The reference refactors A every
    step (A\\b); here one LU of the row-equilibrated matrix is reused and refined (see
    _refined_solve for why).  Singular operators use the pseudo-inverse of the mean-projected
    system."""

    def __init__(self, A, nn, singular, dx):
        import scipy.linalg as sla
        self.singular, self.sla = singular, sla
        Am = A.reshape((nn, nn), order="F")
        if singular:
            self.pinv = np.linalg.pinv(Am * dx * dx, rcond=1e-12) * dx * dx
        else:
            self.d = np.where((np.abs(np.diag(Am)) == 1.0) & (np.count_nonzero(Am, axis=1) == 1), 1.0, dx * dx)
            self.As = Am * self.d[:, None]
            self.Al = self.As.astype(np.longdouble)
            self.lu = sla.lu_factor(self.As)

    def __call__(self, b):
        if self.singular:
            bb = b - b.mean()
            phi = self.pinv @ bb
            return phi - phi.mean()
        bs = b * self.d
        x = self.sla.lu_solve(self.lu, bs)
        for _ in range(2):
            r = (bs.astype(np.longdouble) - self.Al @ x.astype(np.longdouble)).astype(np.float64)
            x = x + self.sla.lu_solve(self.lu, r)
        return x


def _oracle_step(species, cg, solver, b, dof, E, dt, bmode, V, nn):
    Lc = CO.lib()
    rho = np.zeros(nn)
    dens = np.zeros(nn)
    for s in species:
        Lc.orc_advance(s.ref(), C.byref(cg), CO.dp(E), C.c_double(dt), (C.c_int32 * 2)(*bmode))
    for s in species:
        Lc.orc_density(C.byref(cg), s.ref(), CO.dp(V), CO.dp(dens))
        Lc.orc_rho_accumulate(C.byref(cg), CO.dp(dens), C.c_double(s.c.q), CO.dp(rho))
    m = dof.astype(bool)
    b[m] = (-rho[m]) / O.eps0                                        # generalized_poisson.jl:375
    phi = np.ascontiguousarray(solver(b))
    Enew = np.zeros(3 * nn)
    Lc.orc_electric_field(C.byref(cg), CO.dp(phi), CO.dp(Enew))
    return rho, phi, Enew


def _c_operator(cg, nn, periodic, edges, nx, ny):
    Lc = CO.lib()
    A = np.zeros(nn * nn)
    b = np.zeros(nn)
    dof = np.ones(nn, dtype=np.uint8)
    Lc.orc_poisson_assemble(C.byref(cg), CO.dp(A))
    for ax in periodic:
        Lc.orc_poisson_apply_periodic(C.byref(cg), CO.dp(A), C.c_int(ax))
    for name, val in edges:
        m = np.zeros((nx, ny), bool)
        if name == "l":
            m[0, :] = True
        else:
            m[nx - 1, :] = True
        mm = np.ascontiguousarray(m.ravel(order="F").astype(np.uint8))
        Lc.orc_poisson_apply_dirichlet(C.byref(cg), CO.dp(A), CO.dp(b), dof.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       mm.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_double(val))
    return A, b, dof


def _two_species(ib, g, cg, n, cap, seed, Te=300.0, drift=1e7, wgt=3.5e7, ion_m=4.002602 * O.me / 5.48579903e-04):
    """Two-stream style load (10_two_streams.jl:41-42,61-69): two cold beams + co-located ions."""
    PIC = ib.particle_in_cell
    nx, ny = g.n
    dx, dy = g.dh
    rng = np.random.default_rng(seed)
    x = rng.random(n) * (nx - 1) * dx
    y = rng.random(n) * (ny - 1) * dy
    vth = O.thermal_speed(Te, O.me)
    v = rng.standard_normal((n, 3)) * vth
    v[: n // 2, 0] += drift
    v[n // 2:, 0] -= drift
    out_c, out_g = [], []
    for name, q, m, vv in (("e-", -O.qe, O.me, v), ("He+", +O.qe, ion_m, np.full((n, 3), 1280.0))):
        pc = CO.CSpecies(cap, q, m, wgt)
        pc.set(x, y, vv[:, 0], vv[:, 1], vv[:, 2])
        pg = PIC.create_kinetic_species(name, cap, q, m, wgt)
        pg.x[:n, 0], pg.x[:n, 1] = x, y
        pg.v[:n] = vv
        pg.np = n
        out_c.append(pc)
        out_g.append(pg)
    return out_c, out_g


def _compare_species(pcs, pgs, tol, lengths):
    """State keyed by id; positions relative to the box size, velocities to max |v|."""
    for pc, pg in zip(pcs, pgs):
        m = pc.np
        assert pg.np == m and m > 100
        xg, yg, v0, v1, v2 = _by_id(pg.id[:m], pg.x[:m, 0], pg.x[:m, 1], pg.v[:m, 0], pg.v[:m, 1], pg.v[:m, 2])
        xc, yc, c0, c1, c2 = _by_id(pc.id[:m], pc.xy[0, :m], pc.xy[1, :m], pc.v[0, :m], pc.v[1, :m], pc.v[2, :m])
        assert np.array_equal(np.sort(pg.id[:m]), np.sort(pc.id[:m]))
        vmax = max(np.abs(c0).max(), np.abs(c1).max(), np.abs(c2).max())
        for a, r, sc in ((xg, xc, lengths[0]), (yg, yc, lengths[1]), (v0, c0, vmax), (v1, c1, vmax), (v2, c2, vmax)):
            assert np.abs(a - r).max() <= tol * sc


@pytest.mark.parametrize("mode", ["operators", "fused", "fused-tiled"])
def test_two_stream_100_steps(ib, mode):
    """C1-like (10_two_streams.jl): periodic x periodic, wrap! on both axes, 100 steps."""
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx, ny = (129, 2) if mode != "fused-tiled" else (65, 33)
    dx = 1.8743613985989574e-08
    dt = 5.301494621374497e-16
    g, cg = _grid_pair(ib, nx, ny, dx)
    nn = nx * ny
    ps = FDM.create_poisson_solver(g, O.eps0)
    FDM.apply_periodic(ps, 1)
    FDM.apply_periodic(ps, 2)
    A, b, dof = _c_operator(cg, nn, (1, 2), (), nx, ny)
    n = 1280 if mode != "fused-tiled" else 60000
    pcs, pgs = _two_species(ib, g, cg, n, n + 100, seed=7, wgt=3.513e7 if ny == 2 else 3.513e7 * 40)
    V = np.zeros(nn)
    CO.lib().orc_cell_volume(C.byref(cg), CO.dp(V))
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), pgs
    E = np.zeros(3 * nn)
    steps = 100
    osolve = _OracleSolver(A, nn, True, dx)
    for _ in range(steps):
        rho, phi, E = _oracle_step(pcs, cg, osolve, b, dof, E, dt, (1, 1), V, nn)
    if mode == "operators":
        PIC.hooks.after_push = lambda part, grid: PIC.wrap_(part, grid)
        PIC.solve(cfg, dt, steps, fused=False)
    else:
        PIC.solve(cfg, dt, steps, after_push=(1, 1), sort_interval=10 if mode == "fused-tiled" else 0)
    rho_g, phi_g, E_g = g._rt.fields()
    _compare_species(pcs, pgs, REL, ((nx - 1) * dx, (ny - 1) * dx))
    assert np.abs(rho_g.ravel(order="F") - rho).max() <= REL * np.abs(rho).max()
    Eg = _colmajor3(E_g)
    assert np.abs(Eg - E).max() <= 1e-8 * np.abs(E).max()           # singular operator: gauge-free E
    assert np.abs((phi_g - phi_g.mean()).ravel(order="F") - phi).max() <= 1e-8 * np.abs(phi).max()


@pytest.mark.parametrize("mode", ["operators", "fused", "fused-tiled"])
def test_rf_like_100_steps_discard(ib, mode):
    """C2-like (11_rf_discharge.jl) without MCC: Dirichlet electrodes in x with a driven voltage,
    'periodic' in y, discard!(dims=1) + wrap!(dims=2), 100 steps."""
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx, ny = (129, 2) if mode != "fused-tiled" else (129, 33)
    dx = 5.234375e-4
    dt = 1.8436578171091445e-10
    g, cg = _grid_pair(ib, nx, ny, dx)
    nn = nx * ny
    ps = FDM.create_poisson_solver(g, O.eps0)
    FDM.apply_periodic(ps, 1)
    left = np.zeros((nx, ny), bool)
    left[0, :] = True
    right = np.zeros((nx, ny), bool)
    right[nx - 1, :] = True
    FDM.apply_dirichlet(ps, left, 0.0)
    FDM.apply_dirichlet(ps, right, 0.0)
    A, b, dof = _c_operator(cg, nn, (1,), (("l", 0.0), ("r", 0.0)), nx, ny)
    n = 4000 if mode != "fused-tiled" else 50000
    rng = np.random.default_rng(21)
    pcs, pgs = [], []
    for name, q, m, T in (("e-", -O.qe, O.me, 10000.0), ("He+", O.qe, 3.99 * O.mp, 300.0)):
        x = rng.random(n) * (nx - 1) * dx
        y = rng.random(n) * (ny - 1) * dx
        v = rng.standard_normal((n, 3)) * O.thermal_speed(T, m)
        pc = CO.CSpecies(n + 10, q, m, 1.37e5)
        pc.set(x, y, v[:, 0], v[:, 1], v[:, 2])
        pg = PIC.create_kinetic_species(name, n + 10, q, m, 1.37e5)
        pg.x[:n, 0], pg.x[:n, 1] = x, y
        pg.v[:n] = v
        pg.np = n
        pcs.append(pc)
        pgs.append(pg)
    V = np.zeros(nn)
    CO.lib().orc_cell_volume(C.byref(cg), CO.dp(V))
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), pgs
    f = 13.56e6
    VRF = 20.0   # reduced from the script's 450 V: the thin test plasma cannot shield the full drive
    E = np.zeros(3 * nn)
    steps = 100
    lmask = np.ascontiguousarray(left.ravel(order="F").astype(np.uint8))
    osolve = _OracleSolver(A, nn, False, dx)
    for it in range(1, steps + 1):
        rho, phi, E = _oracle_step(pcs, cg, osolve, b, dof, E, dt, (2, 1), V, nn)
        t = it * dt - dt
        CO.lib().orc_poisson_apply_dirichlet(C.byref(cg), CO.dp(A), CO.dp(b), dof.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             lmask.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             C.c_double(VRF * math.sin(2 * math.pi * f * t)))   # 11_rf_discharge.jl:95
    PIC.hooks.after_loop = lambda i, t, dt_: FDM.apply_dirichlet(ps, left, VRF * math.sin(2 * math.pi * f * t))

    def after_push(part, grid):
        PIC.discard_(part, grid, dims=[1])
        PIC.wrap_(part, grid, dims=[2])
    try:
        if mode == "operators":
            PIC.hooks.after_push = after_push
            PIC.solve(cfg, dt, steps, fused=False)
        else:
            PIC.solve(cfg, dt, steps, after_push=(2, 1), sort_interval=7 if mode == "fused-tiled" else 0)
    finally:
        PIC.hooks.after_loop = lambda i, t, dt_: None
        PIC.hooks.after_push = lambda part, grid: PIC.wrap_(part, grid)
    assert pcs[0].np < n                                            # electrons did reach the walls
    rho_g, phi_g, E_g = g._rt.fields()
    _compare_species(pcs, pgs, REL, ((nx - 1) * dx, (ny - 1) * dx))
    assert np.abs(rho_g.ravel(order="F") - rho).max() <= REL * np.abs(rho).max()
    assert np.abs(phi_g.ravel(order="F") - phi).max() <= REL * np.abs(phi).max()
    assert np.abs(_colmajor3(E_g) - E).max() <= REL * np.abs(E).max()


def test_tiled_deposit_matches_simple_when_particles_leave_window(ib):
    """Fast particles between sorts fall back to global adds; the total must not change."""
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx, ny, dx = 129, 129, 1e-3
    res = []
    for sort_interval in (0, 1000):
        g, cg = _grid_pair(ib, nx, ny, dx)
        ps = FDM.create_poisson_solver(g, O.eps0)
        FDM.apply_periodic(ps, 1)
        FDM.apply_periodic(ps, 2)
        pc, pg = _species_pair(ib, g, 200000, 200000, seed=9, w=1.0, vscale=3.0 * dx / 1e-9)   # ~3 cells per step
        cfg = ib.configuration.Config()
        cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), [pg]
        PIC.solve(cfg, 1e-9, 5, after_push=(1, 1), sort_interval=sort_interval)
        res.append((PIC.density(pg, g).copy(), g._rt.fields()[0].copy(), np.sort(pg.x[:pg.np, 0])))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-12, atol=1e-12 * res[0][0].max())
    assert np.allclose(res[0][1], res[1][1], rtol=1e-11, atol=1e-11 * np.abs(res[0][1]).max())
    assert np.allclose(res[0][2], res[1][2], rtol=1e-12, atol=1e-12 * nx * dx)


# ------------------------------------------------------- kinetic.jl:20-50 remove! / add! / remove_particles! ---
def test_remove_add_remove_particles_vs_oracle(ib):
    PIC = ib.particle_in_cell
    nx, ny, dx = 17, 9, 0.25
    g, cg = _grid_pair(ib, nx, ny, dx)
    og = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    n, cap = 500, 1200
    rng = np.random.default_rng(21)
    x = rng.random((n, 2)) * np.array([(nx - 1) * dx, (ny - 1) * dx])
    v = rng.standard_normal((n, 3))
    wg = 2.0 + rng.random(n)

    def make():
        o = O.KineticSpecies("a", cap, -O.qe, O.me, 7.0)
        o.x[:n], o.v[:n], o.wg[:n], o.np = x, v, wg, n
        s = PIC.create_kinetic_species("a", cap, -O.qe, O.me, 7.0)
        s.x[:n] = x
        s.v[:n] = v
        s.wg[:n] = wg
        s.np = n
        s._push(g)
        return o, s
    o1, s1 = make()
    # remove!: exactly the reference's swap with the last row, repeated; whole arrays bit-identical
    for i in (1, 250, 498, 3, 3):
        O.remove_(o1, i)
        PIC.remove_(s1, i)
    assert s1.np == o1.np == n - 5
    assert np.array_equal(s1.x[: o1.np], o1.x[: o1.np]) and np.array_equal(s1.v[: o1.np], o1.v[: o1.np])
    assert np.array_equal(s1.wg, o1.wg) and np.array_equal(s1.id, o1.id)
    # add!: rows appended, wg and id of the destination untouched
    o2, s2 = make()
    O.add_(o1, o2)
    PIC.add_(s1, s2)
    assert s2.np == o2.np == 2 * n - 5
    assert np.array_equal(s2.x[: o2.np], o2.x[: o2.np]) and np.array_equal(s2.v[: o2.np], o2.v[: o2.np])
    assert np.array_equal(s2.wg, o2.wg) and np.array_equal(s2.id, o2.id)
    with pytest.raises(ib.IskraError):
        PIC.add_(s2, s1)                                            # 995 + 495 rows > capacity 1200 (reference: BoundsError)
    # remove_particles!: same survivors (keyed by id), ids stay a permutation
    matches = lambda i, j: (i + j) % 3 == 0 or i == 5
    O.remove_particles_(o2, og.dh, matches)
    removed = PIC.remove_particles_(s2, g, matches)
    assert removed == 2 * n - 5 - o2.np and s2.np == o2.np and removed > 100
    m = o2.np
    gx, gy, gw = _by_id(s2.id[:m], s2.x[:m, 0], s2.x[:m, 1], s2.wg[:m])
    ox, oy, ow = _by_id(o2.id[:m], o2.x[:m, 0], o2.x[:m, 1], o2.wg[:m])
    assert np.array_equal(np.sort(s2.id[:m]), np.sort(o2.id[:m]))
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gw, ow)
    assert np.array_equal(np.sort(s2.id), np.arange(1, cap + 1))


def test_vacated_slots_get_the_default_weight(ib):
    """remove! resets the weight of the slot it vacates (kinetic.jl:24), so rows created later in those slots (add!,
    sample!, ionisation) start with w0 -- also when the rows were discarded in bulk and parked by a compaction."""
    PIC = ib.particle_in_cell
    nx, ny, dx = 33, 33, 1e-3
    g, cg = _grid_pair(ib, nx, ny, dx)
    rng = np.random.default_rng(4)
    n, w0 = 3000, 2.5
    sp = PIC.create_kinetic_species("s", n + 500, -O.qe, O.me, w0)
    sp.x[:n, 0] = rng.random(n) * (nx - 1) * dx * 1.3          # a quarter lies beyond the right edge
    sp.x[:n, 1] = rng.random(n) * (ny - 1) * dx
    sp.v[:n] = rng.standard_normal((n, 3))
    sp.wg[:n] = 0.5 + rng.random(n)
    sp.np = n
    PIC.discard_(sp, g)
    kept = sp.np
    assert 0 < kept < n
    src = PIC.create_kinetic_species("src", 400, -O.qe, O.me, w0)
    src.x[:400] = rng.random((400, 2)) * (nx - 1) * dx
    src.np = 400
    src._push(g)
    PIC.add_(src, sp)
    assert sp.np == kept + 400
    assert np.all(sp.wg[kept:kept + 400] == w0)
    assert np.all(sp.wg[kept + 400:] == w0)
    assert sorted(sp.id.tolist()) == list(range(1, n + 501))
