"""SURVEY.md 8f row N3 -- the axisymmetric r-z variant (problem/13_seed.jl): oracle pinning on the CPU, parity of the
CUDA path on the GPU.  Bars: transform / push bit-exact, ring volumes and the assembled operator bit-exact, a
13_seed-like loop (Dirichlet plates in z, discard! dim 2) within 1e-10."""
import math

import numpy as np
import pytest

from oracle import axial_oracle as AX
from oracle import pic_oracle as O

REL = 1e-10


# ------------------------------------------------------------------ CPU: oracle pinning ------
def test_ring_volumes_known_answer():
    g = AX.AxialGrid2(np.arange(9) * 0.01, np.arange(7) * 0.02)
    V = AX.cell_volume(g)
    # the rings of one full-height column tile the disc of radius (nr - 1) dr  (RegularGrids.jl:45-49)
    assert V[:, 3].sum() == pytest.approx(math.pi * 0.02 * (8 * 0.01) ** 2, rel=1e-13)
    assert V[0, 3] == pytest.approx(math.pi * 0.02 * 0.005 ** 2, rel=1e-15)
    assert np.array_equal(V[:, 0], 0.5 * V[:, 3]) and np.array_equal(V[:, 6], 0.5 * V[:, 3])


def test_axial_operator_known_answers():
    g = AX.AxialGrid2(np.arange(9) * 0.01, np.arange(7) * 0.02)
    ps = AX.PoissonSolver(g, 1.0)
    nr, nz = g.n
    assert np.abs(ps.A.sum(axis=1)).max() < 1e-9                   # constants are in the null space (all-Neumann)
    r = (np.arange(nr) * g.dh[0])[:, None] * np.ones((1, nz))
    res = (ps.A @ (5.0 - 3.0 * r ** 2).reshape(-1, order="F")).reshape((nr, nz), order="F")
    assert np.allclose(res[1:-1, :], -12.0, rtol=1e-9)             # (1/r)(r phi')' = -4b exactly on a - b r^2
    assert res[0, 3] == pytest.approx(-3.0, rel=1e-9)              # the reference's axis row is (phi2 - phi1)/dr^2 (:137-139)


def test_host_assembler_matches_oracle_bit_for_bit():
    from iskra_b200 import finite_difference_method as FDM
    from iskra_b200 import regular_grids as RG

    class G:
        pass
    for nr, nz, dr, dz in ((9, 7, 0.01, 0.02), (33, 65, 0.0025, 0.0025), (3, 3, 1.0, 0.5)):
        g = G()
        g.n, g.dh, g.bcs = (nr, nz), (dr, dz), (("other", "other"), ("open", "open"))
        og = AX.AxialGrid2(np.arange(nr) * dr, np.arange(nz) * dz)
        og.dh = (dr, dz)
        assert np.array_equal(FDM.assemble_axial_operator(g), AX.PoissonSolver(og, 1.0).A)
        assert np.array_equal(RG.axial_cell_volume(g), AX.cell_volume(og))


def test_transform_known_answer():
    sp = O.KineticSpecies("e", 4, -1.0, 1.0, 1.0)
    sp.x[:3, 0] = [3.0, 0.0, 2.0]
    sp.v[:3] = [[1.0, 7.0, 4.0], [5.0, 0.0, 0.0], [0.5, 0.0, 0.0]]
    sp.np = 3
    AX.transform_from_cartesian_to_cylindrical_(sp, 1.0)
    assert sp.x[0, 0] == 5.0 and sp.v[0, 1] == 7.0                  # r = sqrt(3^2 + 4^2)
    assert sp.v[0, 0] == pytest.approx(0.6 * 1.0 + 0.8 * 4.0) and sp.v[0, 2] == pytest.approx(-0.8 * 1.0 + 0.6 * 4.0)
    assert sp.x[1, 0] == 0.0 and sp.v[1, 0] == 5.0                  # r == 0: sin = 0, cos = 1
    assert sp.x[2, 0] == 2.0 and sp.v[2, 0] == 0.5                  # no azimuthal motion: unchanged


def test_numpy_and_c_axial_advance_agree_bit_for_bit():
    import ctypes as C
    from oracle import c_oracle as CO
    nr, nz, dh, dt, n = 17, 33, 0.0025, 7.5e-11, 5000
    og = AX.AxialGrid2(np.arange(nr) * dh, np.arange(nz) * dh)
    cg = CO.make_grid(nr, nz, dh, dh)
    rng = np.random.default_rng(2)
    x = rng.random((n, 2)) * np.array([(nr - 1) * dh * 0.8, (nz - 1) * dh])
    v = rng.standard_normal((n, 3)) * 0.1 * dh / dt
    x[:3, 0] = 0.0
    v[:3, 2] = 0.0
    E = np.zeros((nr, nz, 3))
    E[:, :, :2] = rng.standard_normal((nr, nz, 2)) * 1e4
    osp = O.KineticSpecies("e-", n + 4, -O.qe, O.me, 1.0)
    osp.x[:n], osp.v[:n], osp.np = x, v, n
    cs = CO.CSpecies(n + 4, -O.qe, O.me, 1.0)
    cs.set(x[:, 0], x[:, 1], v[:, 0], v[:, 1], v[:, 2])
    Ec = np.ascontiguousarray(E.reshape(-1, order="F"))
    for _ in range(5):
        AX.advance_(osp, E, dt, og, lambda p, gg: O.discard_(p, gg, dims=(2,)))
        CO.lib().orc_advance_rz(cs.ref(), C.byref(cg), CO.dp(Ec), C.c_double(dt), (C.c_int32 * 2)(0, 2))
        m = osp.np
        assert cs.np == m
        assert np.array_equal(osp.x[:m, 0], cs.xy[0, :m]) and np.array_equal(osp.x[:m, 1], cs.xy[1, :m])
        assert np.array_equal(osp.v[:m].T, cs.v[:, :m]) and np.array_equal(osp.id, cs.id)
    assert m < n


# ------------------------------------------------------------------ GPU parity ----------------
@pytest.fixture(scope="module")
def ib():
    import iskra_b200
    return iskra_b200


def _by_id(ids, *cols):
    o = np.argsort(ids, kind="stable")
    return [np.asarray(c)[o] for c in cols]


def _pair(ib, nr, nz, dr, dz, n, seed, q=-O.qe, m=O.me, w=5e5, vscale=None):
    PIC = ib.particle_in_cell
    og = AX.AxialGrid2(np.arange(nr) * dr, np.arange(nz) * dz)
    g = ib.regular_grids.create_axial_grid(np.arange(nr) * dr, np.arange(nz) * dz)
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2)) * np.array([(nr - 1) * dr * 0.9, (nz - 1) * dz])
    v = rng.standard_normal((n, 3)) * (vscale if vscale else 0.05 * dr / 7.5e-11)
    osp = O.KineticSpecies("e-", n + 8, q, m, w)
    osp.x[:n], osp.v[:n], osp.np = x, v, n
    gsp = PIC.create_kinetic_species("e-", n + 8, q, m, w)
    gsp.x[:n] = x
    gsp.v[:n] = v
    gsp.np = n
    return og, g, osp, gsp


@pytest.mark.gpu
def test_axial_volume_transform_push_density_bitexact(ib):
    PIC, RG = ib.particle_in_cell, ib.regular_grids
    nr, nz, dh, dt, n = 33, 65, 0.0025, 7.5e-11, 20000
    og, g, osp, gsp = _pair(ib, nr, nz, dh, dh, n, seed=4)
    assert np.array_equal(RG.cell_volume(g), AX.cell_volume(og))
    osp.x[:5, 0] = 0.0
    osp.v[:5, 2] = 0.0                                                # r == 0 rows
    gsp.x[:5, 0] = 0.0
    gsp.v[:5, 2] = 0.0
    # transform on its own
    AX.transform_from_cartesian_to_cylindrical_(osp, dt)
    gsp._push(g)
    PIC.transform_from_cartesian_to_cylindrical_(gsp, dt)
    assert np.array_equal(gsp.x[:n], osp.x[:n]) and np.array_equal(gsp.v[:n], osp.v[:n])
    # push_particles!(::BorisPusher{:rz}) from a device field
    rng = np.random.default_rng(5)
    E = np.zeros((nr, nz, 3))
    E[:, :, :2] = rng.standard_normal((nr, nz, 2)) * 2e4
    AX.push_particles_rz_(osp, O.grid_to_particle(og, osp, E), dt)
    g._rt.set_fields(E=E)
    PIC.push_particles_(PIC.create_axial_boris_pusher(), gsp, None, None, dt, g)
    assert np.array_equal(gsp.x[:n], osp.x[:n]) and np.array_equal(gsp.v[:n], osp.v[:n])
    # density with the ring volumes (deposit order differs: 1e-13); rows that left the box are discarded first
    assert O.discard_(osp, og) == PIC.discard_(gsp, g)
    nd = PIC.density(gsp, g)
    nref = AX.density(osp, og)
    assert np.max(np.abs(nd - nref)) <= 1e-12 * np.max(np.abs(nref))


@pytest.mark.gpu
def test_axial_operator_on_device_bitexact(ib):
    FDM = ib.finite_difference_method
    nr, nz, dh = 17, 33, 0.0025
    og = AX.AxialGrid2(np.arange(nr) * dh, np.arange(nz) * dh)
    g = ib.regular_grids.create_axial_grid(np.arange(nr) * dh, np.arange(nz) * dh)
    ops = AX.PoissonSolver(og, O.eps0)
    ps = FDM.create_poisson_solver(g, O.eps0)
    bot = np.zeros((nr, nz), bool)
    bot[:, 0] = True
    top = np.zeros((nr, nz), bool)
    top[:, nz - 1] = True
    for s, o in ((FDM, ps), (O, ops)):
        s.apply_dirichlet(o, bot, 0.0)
        s.apply_dirichlet(o, top, 1600.0)
    A, b = ps.dense()
    assert np.array_equal(A, ops.A) and np.array_equal(b, ops.b)
    assert ps.mode == "dense"


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False, "tiled"])
def test_seed_like_loop_vs_oracle(ib, fused):
    """problem/13_seed.jl without MCC: axial grid 33 x 65, 0 V / E*d plates at z = 0 / Lz, axial pusher,
    discard!(dims = 2); phi against the exact solution of the reference's system."""
    import scipy.linalg as sla
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nr, nz, dh, dt, n = 33, 65, 0.08 / 32, 7.5e-11, 30000
    og, g, osp, gsp = _pair(ib, nr, nz, dh, dh, n, seed=8)
    ops = AX.PoissonSolver(og, O.eps0)
    ps = FDM.create_poisson_solver(g, O.eps0)
    bot = np.zeros((nr, nz), bool)
    bot[:, 0] = True
    top = np.zeros((nr, nz), bool)
    top[:, nz - 1] = True
    for s, o in ((FDM, ps), (O, ops)):
        s.apply_dirichlet(o, bot, 0.0)
        s.apply_dirichlet(o, top, 20000 * 0.08)
    sc = 1.0 / np.max(np.abs(ops.A), axis=1)
    As = ops.A * sc[:, None]
    lu = sla.lu_factor(As)
    Al = As.astype(np.longdouble)

    def exact(b):
        bs = b * sc
        x = sla.lu_solve(lu, bs)
        for _ in range(3):
            x = x + sla.lu_solve(lu, (bs.astype(np.longdouble) - Al @ x.astype(np.longdouble)).astype(np.float64))
        return x
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_axial_boris_pusher(), [gsp]
    E = np.zeros((nr, nz, 3))
    rd = np.asarray(ops.rho_dof, dtype=np.int64)
    PIC.hooks.after_push = lambda part, grid: PIC.discard_(part, grid, dims=[2])
    try:
        for it in range(12):
            AX.advance_(osp, E, dt, og, lambda p, gg: O.discard_(p, gg, dims=(2,)))
            rho = AX.density(osp, og) * osp.q
            ops.b[rd] = (-rho).reshape(-1, order="F")[rd] / ops.eps0
            phi = exact(ops.b).reshape((nr, nz), order="F")
            E = O.calculate_electric_field(ops, phi)
            # "tiled": the axial pusher inside the tile-directory kernels (re-group every 3 steps)
            PIC.solve(cfg, dt, 1, after_push=(ib._lib.BND_NONE, ib._lib.BND_DISCARD), fused=bool(fused),
                      sort_interval=3 if fused == "tiled" else 0)
            grho, gphi, gE = g._rt.fields()
            assert np.max(np.abs(gphi - phi)) <= REL * np.max(np.abs(phi)), it
            assert np.max(np.abs(grho - rho)) <= REL * np.max(np.abs(rho)), it
            assert np.max(np.abs(gE - E)) <= REL * np.max(np.abs(E)), it
            assert gsp.np == osp.np
    finally:
        PIC.hooks.after_push = lambda part, grid: PIC.wrap_(part, grid)
        ib.particle_in_cell._set_pusher(g._rt, None)
    m = osp.np
    gx, gy, gvx, gvz = _by_id(gsp.id[:m], gsp.x[:m, 0], gsp.x[:m, 1], gsp.v[:m, 0], gsp.v[:m, 2])
    ox, oy, ovx, ovz = _by_id(osp.id[:m], osp.x[:m, 0], osp.x[:m, 1], osp.v[:m, 0], osp.v[:m, 2])
    assert np.max(np.abs(gx - ox)) <= REL * 0.08 and np.max(np.abs(gy - oy)) <= REL * 0.16
    assert np.max(np.abs(gvx - ovx)) <= REL * np.max(np.abs(ovx)) and np.max(np.abs(gvz - ovz)) <= REL * np.max(np.abs(ovz))
