"""GPU parity tests of SURVEY.md 8f row N1 (surface tracker, electrodes, sigma dofs, circuit) against
the oracle (oracle/surfaces_oracle.py, the tracker part of oracle/iskra_oracle.c), through the C ABI.

Bars: face lookups, tracked / absorbed counts and which particles survive (keyed by id) exact;
positions and velocities of survivors bit-exact for one advance! (the walk keeps the reference's
operation order without FMA contraction); the assembled operator with sigma dofs bit-exact; phi / E /
particle state over several steps within 1e-10; collected electrode charge within 1e-12 (sum order).
"""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import c_oracle as CO
from oracle import pic_oracle as O
from oracle import surfaces_oracle as S

pytestmark = pytest.mark.gpu
REL = 1e-10


@pytest.fixture(scope="module")
def ib():
    import iskra_b200
    return iskra_b200


def _by_id(ids, *cols):
    o = np.argsort(ids, kind="stable")
    return [np.asarray(ids)[o]] + [np.asarray(c)[o] for c in cols]


def _boundaries(ib, n=20000, seed=3, nx=11, ny=11, dh=0.1, vfrac=0.3, dt=1e-8, electrodes=True):
    """problem/07_boundaries.jl geometry on both sides (oracle objects and device objects)."""
    PIC, FDM, CFG = ib.particle_in_cell, ib.finite_difference_method, ib.configuration
    xs, ys = np.arange(nx) * dh, np.arange(ny) * dh
    og = O.CartesianGrid2(xs, ys)
    ops = O.PoissonSolver(og, O.eps0)
    ost = S.create_surface_tracker(og)
    g = ib.regular_grids.create_uniform_grid(xs, ys)
    cfg = CFG.Config()
    cfg.grid = g
    cfg.solver = FDM.create_poisson_solver(g, O.eps0)
    cfg.pusher = PIC.create_boris_pusher()
    bcs = np.zeros((nx, ny), dtype=np.int8)
    bcs[0, 1:ny - 1] = 1
    bcs[nx - 1, 4:7] = 2
    bcs[nx - 2, 0] = 3
    bcs[nx - 2, ny - 1] = 3
    bcs[5:8, 4:7] = 4
    oel, gel = [], []
    if electrodes:
        oel.append(S.create_electrode(bcs == 1, ops, og, ost, sigma=1 * O.eps0))
        oel.append(S.create_electrode(bcs == 2, ops, og, ost))
        oel.append(S.create_electrode(bcs == 3, ops, og, ost, fixed=True))
        gel.append(CFG.create_electrode(bcs == 1, cfg, sigma=1 * O.eps0))
        gel.append(CFG.create_electrode(bcs == 2, cfg))
        gel.append(CFG.create_electrode(bcs == 3, cfg, fixed=True))
    else:
        cfg.tracker = PIC.create_surface_tracker(g)
    S.track_surface_(ost, bcs == 4, S.create_reflective_surface())
    PIC.track_surface_(cfg.tracker, bcs == 4, PIC.create_reflective_surface())
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2)) * np.array([(nx - 1) * dh, (ny - 1) * dh])
    inside = (x[:, 0] > 0.5) & (x[:, 0] < 0.7) & (x[:, 1] > 0.4) & (x[:, 1] < 0.6)
    x = x[~inside]
    n = len(x)
    v = rng.standard_normal((n, 3)) * vfrac * dh / dt
    wg = 50e3 * (0.5 + rng.random(n))
    cap = n + 16
    osp = O.KineticSpecies("e-", cap, -O.qe, O.me, 50e3)
    osp.x[:n], osp.v[:n], osp.wg[:n], osp.np = x, v, wg, n
    gsp = PIC.create_kinetic_species("e-", cap, -O.qe, O.me, 50e3)
    gsp.x[:n] = x
    gsp.v[:n] = v
    gsp.wg[:n] = wg
    gsp.np = n
    cfg.species = [gsp]
    return dict(og=og, ops=ops, ost=ost, oel=oel, osp=osp, g=g, cfg=cfg, gel=gel, gsp=gsp, dt=dt, bcs=bcs, n=n)


def test_face_table_matches_the_dict(ib):
    c = _boundaries(ib, n=10)
    ost, st = c["ost"], c["cfg"].tracker
    nx, ny = c["og"].n
    for i in range(0, nx + 1):
        for j in range(0, ny + 1):
            for (k, l) in ((i, j - 1), (i + 1, j), (i, j + 1), (i - 1, j)):
                ref = ost.surface.get(((i, j), (k, l)))
                got = st.get(((i, j), (k, l)), None)
                assert (ref.kind if ref is not None else None) == got, ((i, j), (k, l))
    assert st.get(((3, 3), (5, 5)), None) is None


def test_track_push_check_bitexact_vs_numpy_oracle(ib):
    PIC = ib.particle_in_cell
    c = _boundaries(ib, n=20000)
    og, ost, osp, g, cfg, gsp, dt = c["og"], c["ost"], c["osp"], c["g"], c["cfg"], c["gsp"], c["dt"]
    nx, ny = og.n
    rng = np.random.default_rng(1)
    E = np.zeros((nx, ny, 3))
    E[:, :, :2] = rng.standard_normal((nx, ny, 2)) * 200.0
    # oracle: track! -> gather -> push -> check!
    S.track_(ost, osp, dt)
    n_trk_ref = len(ost.tracked)
    O.push_in_cartesian_(osp, O.grid_to_particle(og, osp, E), dt)
    tf_ref, nabs_ref = S.check_(ost, osp, dt)
    # device, operator by operator
    g._rt.set_fields(E=E)
    n_trk = PIC.track_(cfg.tracker, gsp, dt, g)
    PIC.push_particles_(cfg.pusher, gsp, None, None, dt, g)
    tf, nabs = PIC.check_(cfg.tracker, gsp, dt)
    assert n_trk == n_trk_ref and n_trk > 1000
    assert nabs == nabs_ref and nabs > 100 and tf == tf_ref
    m = osp.np
    assert gsp.np == m
    ids, gx, gy, gvx, gvy, gvz = _by_id(gsp.id[:m], gsp.x[:m, 0], gsp.x[:m, 1], gsp.v[:m, 0], gsp.v[:m, 1], gsp.v[:m, 2])
    oid, ox, oy, ovx, ovy, ovz = _by_id(osp.id[:m], osp.x[:m, 0], osp.x[:m, 1], osp.v[:m, 0], osp.v[:m, 1], osp.v[:m, 2])
    assert np.array_equal(np.sort(gsp.id[:m]), np.sort(osp.id[:m]))          # the same particles survive
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)                  # bit-exact, reflections included
    assert np.array_equal(gvx, ovx) and np.array_equal(gvy, ovy) and np.array_equal(gvz, ovz)
    assert np.array_equal(np.sort(gsp.id), np.arange(1, gsp.N + 1))           # id stays a permutation (kinetic.jl:20-27)
    # electrodes: collected charge (sum order differs), sigma right-hand side untouched by hits (quirk S1)
    assert c["oel"][1].dq != 0.0
    assert c["gel"][1].dq == pytest.approx(c["oel"][1].dq, rel=1e-12)
    assert c["gel"][0].dq == pytest.approx(c["oel"][0].dq, rel=1e-12)
    assert c["gel"][0].phi.value == 1 * O.eps0 and c["gel"][1].phi.value == 0.0


def test_dense_operator_with_sigma_dofs_bitexact(ib):
    c = _boundaries(ib, n=10)
    A, b = c["cfg"].solver.dense()
    assert A.shape == c["ops"].A.shape == (11 * 11 + 2, 11 * 11 + 2)
    assert np.array_equal(A, c["ops"].A)
    assert np.array_equal(b, c["ops"].b)
    assert c["cfg"].solver.mode == "dense"


def _exact_solve(A, b):
    sc = 1.0 / np.max(np.abs(A), axis=1)
    As = A * sc[:, None]
    lu = sla.lu_factor(As)
    bs = b * sc
    x = sla.lu_solve(lu, bs)
    Al = As.astype(np.longdouble)
    for _ in range(3):
        x = x + sla.lu_solve(lu, (bs.astype(np.longdouble) - Al @ x.astype(np.longdouble)).astype(np.float64))
    return x


def test_sigma_driven_plate_known_answer_on_device(ib):
    """06_circuit.jl geometry: sigma on the left plate, grounded right plate, no charge: E_x = sigma."""
    FDM, CFG = ib.finite_difference_method, ib.configuration
    nx, ny, dh = 21, 21, 0.05
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dh, np.arange(ny) * dh)
    cfg = CFG.Config()
    cfg.grid, cfg.solver = g, FDM.create_poisson_solver(g, O.eps0)
    bcs = np.zeros((nx, ny), dtype=np.int8)
    bcs[0, :] = 1
    bcs[nx - 1, :] = 2
    driven = CFG.create_electrode(bcs == 1, cfg, sigma=7.5)
    CFG.create_electrode(bcs == 2, cfg, fixed=True)
    phi = FDM.calculate_electric_potential(cfg.solver, np.zeros((nx, ny)))
    E = FDM.calculate_electric_field(cfg.solver)
    assert np.allclose(E[:, :, 0], 7.5, rtol=1e-10)
    assert np.allclose(E[:, :, 1], 0.0, atol=1e-9)
    assert np.allclose(phi[nx - 1, :], 0.0, atol=1e-12)
    assert driven.sigma.value == pytest.approx(7.5 * (nx - 1) * dh, rel=1e-10)   # S1: .sigma reads phi at node (1,1)
    driven.phi.add(2.5)                                                           # ... and .phi is the sigma rhs
    FDM.calculate_electric_potential(cfg.solver, np.zeros((nx, ny)))
    assert np.allclose(FDM.calculate_electric_field(cfg.solver)[:, :, 0], 10.0, rtol=1e-10)


@pytest.mark.parametrize("fused", [True, False])
def test_loop_with_tracker_and_electrodes_vs_oracle(ib, fused):
    """Several iterations of the loop body (advance! with tracker -> density -> rho -> phi -> E) on
    the 07_boundaries geometry; phi against the exact solution of the reference's system."""
    PIC = ib.particle_in_cell
    c = _boundaries(ib, n=30000, vfrac=0.2)
    og, ops, ost, osp, g, cfg, gsp, dt = c["og"], c["ops"], c["ost"], c["osp"], c["g"], c["cfg"], c["gsp"], c["dt"]
    nx, ny = og.n
    E = np.zeros((nx, ny, 3))
    steps = 8
    PIC.hooks.after_push = lambda part, grid: PIC.discard_(part, grid)
    try:
        for it in range(steps):
            S.advance_(osp, E, dt, og, ost, lambda p, gg: O.discard_(p, gg))
            rho = O.density(osp, og) * osp.q
            ff = (-rho).reshape(-1, order="F")
            rd = np.asarray(ops.rho_dof, dtype=np.int64)
            ops.b[rd] = ff[rd] / ops.eps0
            ops.x[:] = _exact_solve(ops.A, ops.b)
            phi = ops.x[ops.phi_dof]
            E = O.calculate_electric_field(ops, phi)
            PIC.solve(cfg, dt, 1, after_push=(ib._lib.BND_DISCARD, ib._lib.BND_DISCARD), fused=fused)
            grho, gphi, gE = g._rt.fields()
            scale = np.max(np.abs(phi))
            assert np.max(np.abs(gphi - phi)) <= REL * scale, it
            assert np.max(np.abs(grho - rho)) <= REL * np.max(np.abs(rho)), it
            assert np.max(np.abs(gE - E)) <= REL * np.max(np.abs(E)), it
            assert gsp.np == osp.np
    finally:
        PIC.hooks.after_push = lambda part, grid: PIC.wrap_(part, grid)
    m = osp.np
    assert m < c["n"] - 200                                                       # walls and electrodes absorbed some
    ids, gx, gy, gvx = _by_id(gsp.id[:m], gsp.x[:m, 0], gsp.x[:m, 1], gsp.v[:m, 0])
    oid, ox, oy, ovx = _by_id(osp.id[:m], osp.x[:m, 0], osp.x[:m, 1], osp.v[:m, 0])
    assert np.array_equal(ids, oid)
    assert np.max(np.abs(gx - ox)) <= REL * 1.0 and np.max(np.abs(gy - oy)) <= REL * 1.0
    assert np.max(np.abs(gvx - ovx)) <= REL * np.max(np.abs(ovx))
    assert c["gel"][1].dq == pytest.approx(c["oel"][1].dq, rel=1e-9)


def test_circuit_coupled_loop_vs_oracle(ib):
    """problem/06_circuit.jl shape with the PlasmaDevice connected: RLC -> d sigma -> field, 20 steps."""
    PIC, FDM, CFG, CIR = ib.particle_in_cell, ib.finite_difference_method, ib.configuration, ib.circuit
    import math
    nx = ny = 21
    dh, dt = 0.05, 1e-8
    xs = np.arange(nx) * dh
    og = O.CartesianGrid2(xs, xs)
    ops = O.PoissonSolver(og, O.eps0)
    ost = S.create_surface_tracker(og)
    g = ib.regular_grids.create_uniform_grid(xs, xs)
    cfg = CFG.Config()
    cfg.grid, cfg.solver, cfg.pusher = g, FDM.create_poisson_solver(g, O.eps0), PIC.create_boris_pusher()
    bcs = np.zeros((nx, ny), dtype=np.int8)
    bcs[0, :] = 1
    bcs[nx - 1, :] = 2
    od = S.create_electrode(bcs == 1, ops, og, ost, sigma=1 * O.eps0)
    on = S.create_electrode(bcs == 2, ops, og, ost, fixed=True)
    gd = CFG.create_electrode(bcs == 1, cfg, sigma=1 * O.eps0)
    gn = CFG.create_electrode(bcs == 2, cfg, fixed=True)
    V = lambda t: math.sin(2 * math.pi * 5e6 * t)
    ocir = S.CircuitRLC(R=1.0, L=1e-6, C=1e-6, V=V, ext=S.PlasmaDevice(od, on))
    cfg.circuit = CIR.rlc(CIR.netlist([("V1", 3, "GND", V), ("L1", "NOD", "VCC", 1e-6), ("C1", "NOD", "VCC", 1e-6),
                                       ("R1", "GND", "NOD", 1.0), ("EXT1", "NOD", "GND", PIC.PlasmaDevice(gd, gn))]))
    n = 5000
    rng = np.random.default_rng(9)
    x = rng.random((n, 2)) * (nx - 1) * dh
    v = rng.standard_normal((n, 3)) * 0.2 * dh / dt
    osp = O.KineticSpecies("e-", n + 8, -O.qe, O.me, 1e3)
    osp.x[:n], osp.v[:n], osp.np = x, v, n
    gsp = PIC.create_kinetic_species("e-", n + 8, -O.qe, O.me, 1e3)
    gsp.x[:n] = x
    gsp.v[:n] = v
    gsp.np = n
    cfg.species = [gsp]
    E = np.zeros((nx, ny, 3))
    PIC.hooks.after_push = lambda part, grid: PIC.discard_(part, grid)
    try:
        for it in range(20):
            S.advance_(osp, E, dt, og, ost, lambda p, gg: O.discard_(p, gg))
            S.advance_circuit_coupling_(ocir, ops, dt)
            rho = O.density(osp, og) * osp.q
            ff = (-rho).reshape(-1, order="F")
            rd = np.asarray(ops.rho_dof, dtype=np.int64)
            ops.b[rd] = ff[rd] / ops.eps0
            ops.x[:] = _exact_solve(ops.A, ops.b)
            phi = ops.x[ops.phi_dof]
            E = O.calculate_electric_field(ops, phi)
            PIC.solve(cfg, dt, 1, fused=False)
            _, gphi, gE = g._rt.fields(rho=False)
            assert cfg.circuit.i == pytest.approx(ocir.i, rel=1e-12, abs=1e-300)
            assert cfg.circuit.q == pytest.approx(ocir.q, rel=1e-12, abs=1e-300)
            assert gd.phi.value == pytest.approx(ops.b[ops.sigma_dof[0]], rel=1e-12)
            assert np.max(np.abs(gphi - phi)) <= REL * np.max(np.abs(phi)), it
            assert np.max(np.abs(gE - E)) <= REL * np.max(np.abs(E)), it
            assert gsp.np == osp.np
    finally:
        PIC.hooks.after_push = lambda part, grid: PIC.wrap_(part, grid)
    assert abs(ocir.i) > 0 and ops.b[ops.sigma_dof[0]] != 1 * O.eps0


def test_route_hits_to_sigma_option(ib):
    """Not the reference's behaviour (quirk S1) but its evident intent: collected charge / area is added
    to the electrode's sigma right-hand side when asked for."""
    PIC = ib.particle_in_cell
    c = _boundaries(ib, n=20000)
    g, cfg, gsp, dt = c["g"], c["cfg"], c["gsp"], c["dt"]
    cfg.tracker.route_hits_to_sigma(True)
    PIC.track_(cfg.tracker, gsp, dt, g)
    PIC.push_particles_(cfg.pusher, gsp, None, None, dt, g)
    PIC.check_(cfg.tracker, gsp, dt)
    fl = c["gel"][1]
    assert fl.dq != 0.0
    assert fl.phi.value == pytest.approx(fl.dq / fl.area, rel=1e-12)


def test_too_fast_flag_and_reflective_box(ib):
    PIC = ib.particle_in_cell
    nx = ny = 11
    dh, dt = 0.1, 1e-3
    xs = np.arange(nx) * dh
    g = ib.regular_grids.create_uniform_grid(xs, xs)
    st = PIC.create_surface_tracker(g, PIC.create_reflective_surface())
    sp = PIC.create_kinetic_species("e-", 8, -O.qe, O.me, 1.0)
    sp.x[:3] = np.array([[0.03, 0.52], [0.02, 0.03], [0.5, 0.5]])
    sp.v[:3] = np.array([[-50.0, 20.0, 0.0], [-60.0, -70.0, 0.0], [0.0, 0.0, 150.0]])
    sp.np = 3
    assert PIC.track_(st, sp, dt, g) == 2
    PIC.push_particles_(None, sp, np.zeros((3, 3)), None, dt, g)
    tf, nabs = PIC.check_(st, sp, dt)
    assert tf and nabs == 0 and sp.np == 3
    x, v = sp.x[:3], sp.v[:3]
    assert x[0, 0] == pytest.approx(0.02, rel=1e-12) and v[0, 0] == 50.0 and v[0, 1] == 20.0
    assert v[1, 0] == 60.0 and v[1, 1] == 70.0
    assert x[1, 0] == pytest.approx(0.04, rel=1e-12) and x[1, 1] == pytest.approx(0.04, rel=1e-12)


@pytest.mark.parametrize("sort_interval", [0, 1])
def test_large_grid_tracked_advance_vs_c_oracle(ib, sort_interval):
    """513^2 nodes, 2e6 particles: fixed electrodes on two whole edges (separable solver), a reflecting
    block, default absorbing walls; fused steps against the C oracle, keyed by id.  sort_interval = 0 runs
    the simple tracked kernel, 1 the tiled one (rows sorted by cell, tracked rows through the drain path)."""
    PIC, FDM, CFG = ib.particle_in_cell, ib.finite_difference_method, ib.configuration
    Lc = CO.lib()
    nx = ny = 513
    dh, dt = 1e-3, 1e-9
    xs = np.arange(nx) * dh
    og = O.CartesianGrid2(xs, xs)
    ost = S.create_surface_tracker(og)
    g = ib.regular_grids.create_uniform_grid(xs, xs)
    cfg = CFG.Config()
    cfg.grid, cfg.solver, cfg.pusher = g, FDM.create_poisson_solver(g, O.eps0), PIC.create_boris_pusher()
    left = np.zeros((nx, ny), bool)
    left[0, :] = True
    right = np.zeros((nx, ny), bool)
    right[nx - 1, :] = True
    block = np.zeros((nx, ny), bool)
    block[200:300, 150:350] = True
    CFG.create_electrode(left, cfg, fixed=True, phi=25.0)
    CFG.create_electrode(right, cfg, fixed=True, phi=0.0)
    PIC.track_surface_(cfg.tracker, block, PIC.create_reflective_surface())
    S.track_surface_(ost, left, S.FixedPotentialElectrode(None, 0.0))
    S.track_surface_(ost, right, S.FixedPotentialElectrode(None, 0.0))
    S.track_surface_(ost, block, S.create_reflective_surface())
    assert cfg.solver.mode == "separable"
    n = 2_000_000
    rng = np.random.default_rng(17)
    x = rng.random((n, 2)) * (nx - 1) * dh
    inside = (x[:, 0] > 0.2) & (x[:, 0] < 0.299) & (x[:, 1] > 0.15) & (x[:, 1] < 0.349)
    x = x[~inside]
    n = len(x)
    v = rng.standard_normal((n, 3)) * 0.25 * dh / dt
    cs = CO.CSpecies(n + 8, -O.qe, O.me, 1e4)
    cs.set(x[:, 0], x[:, 1], v[:, 0], v[:, 1], v[:, 2])
    gsp = PIC.create_kinetic_species("e-", n + 8, -O.qe, O.me, 1e4)
    gsp.x[:n] = x
    gsp.v[:n] = v
    gsp.np = n
    cfg.species = [gsp]
    ct = CO.CTracker(ost, nx, ny)
    cg = CO.make_grid(nx, ny, dh, dh)
    E = np.zeros(3 * nx * ny)
    V = np.zeros(nx * ny)
    Lc.orc_cell_volume(C.byref(cg), CO.dp(V))
    tot = 0
    for it in range(4):
        nabs, _ = ct.advance(cs, cg, E, dt, bmode=(2, 2))
        tot += nabs
        PIC.solve(cfg, dt, 1, after_push=(ib._lib.BND_DISCARD, ib._lib.BND_DISCARD), fused=True,
                  sort_interval=sort_interval)
        _, _, gE = g._rt.fields(rho=False, phi=False)
        E = np.ascontiguousarray(gE.reshape(-1, order="F"))       # the device field drives both sides
        assert gsp.np == cs.np, it
    assert tot > 1000
    m = cs.np
    ids, gx, gy, gvx, gvy = _by_id(gsp.id[:m], gsp.x[:m, 0], gsp.x[:m, 1], gsp.v[:m, 0], gsp.v[:m, 1])
    oid, ox, oy, ovx, ovy = _by_id(cs.id[:m], cs.xy[0, :m], cs.xy[1, :m], cs.v[0, :m], cs.v[1, :m])
    assert np.array_equal(ids, oid)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)      # same E on both sides => bit-exact
    assert np.array_equal(gvx, ovx) and np.array_equal(gvy, ovy)
