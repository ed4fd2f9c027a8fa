"""Committed golden fixtures (tests/golden/pic_golden.npz, made by tests/golden/make_golden.py from the
numpy oracle) against (a) the numpy oracle as it is now, (b) the C oracle, (c) the CUDA path.
The reference itself is Julia and cannot run here or on the GPU box; the oracle is
pinned on the reference's stored known answers in test_oracle.py and frozen by these vectors."""
import ctypes as C
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import pic_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "pic_golden.npz"))


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _colmajor3(E):
    return np.ascontiguousarray(np.asarray(E).transpose(2, 1, 0)).ravel()


# ------------------------------------------------------------------------------------ CPU ----
def test_golden_file_is_what_the_numpy_oracle_produces():
    mk = _maker()
    out = {}
    mk.operators_case(out)
    mk.rf_steps_case(out)
    mk.xsec_case(out)
    assert sorted(out) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(np.asarray(out[k]), GOLD[k]), k


def test_notebook_values_follow_from_the_oracle_constants():
    nb = json.load(open(os.path.join(HERE, "golden", "reference_notebook_values.json")))
    assert O.me / O.QE_MCC == nb["m_eV_electron"][0]
    assert 0.5 * O.thermal_speed(30000.0, O.me) == nb["half_thermal_speed_Te"][0]
    dt, nHe, npart = 1.8436578171091445e-10, 9.64e20, 16384
    for key, msg, N in (("electron", "max_sigma_g_electron", 4), ("ion", "max_sigma_g_ion", 2)):
        cand = N * (1.0 - np.exp(-nHe * nb[msg][0] * dt)) * npart          # mcc.jl:242-248
        assert int(cand) == int(nb["candidates_%s_step1" % key][0])


def test_c_oracle_reproduces_golden_operators():
    Lc = CO.lib()
    nx, ny, dx = int(GOLD["ops_grid"][0]), int(GOLD["ops_grid"][1]), float(GOLD["ops_grid"][2])
    dt = float(GOLD["ops_dt"][0])
    cg = CO.make_grid(nx, ny, dx, dx)
    x0, v0, wg = GOLD["ops_x0"], GOLD["ops_v0"], GOLD["ops_wg"]
    n = len(wg)
    pc = CO.CSpecies(n + 8, -O.qe, O.me, 1.37e5)
    pc.set(x0[:, 0], x0[:, 1], v0[:, 0], v0[:, 1], v0[:, 2], wg)
    ci, cj, chx, chy = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n), np.zeros(n)
    Lc.orc_particle_cell(CO.dp(pc.xy[0]), CO.dp(pc.xy[1]), C.c_int64(n), C.c_double(dx), C.c_double(dx),
                         ci.ctypes.data_as(CO.c_i64p), cj.ctypes.data_as(CO.c_i64p), CO.dp(chx), CO.dp(chy))
    assert np.array_equal(ci, GOLD["ops_i"]) and np.array_equal(cj, GOLD["ops_j"])
    assert np.array_equal(chx, GOLD["ops_hx"]) and np.array_equal(chy, GOLD["ops_hy"])
    pE = np.zeros(3 * n)
    Lc.orc_gather(C.byref(cg), pc.ref(), CO.dp(_colmajor3(GOLD["ops_E"])), CO.dp(pE))
    assert np.array_equal(pE.reshape(3, n).T, GOLD["ops_partE"])
    Lc.orc_push(pc.ref(), CO.dp(pE), C.c_double(dt))
    assert np.array_equal(pc.xy[:, :n].T, GOLD["ops_x_pushed"]) and np.array_equal(pc.v[:, :n].T, GOLD["ops_v_pushed"])
    xb = GOLD["ops_x_before_wrap"]
    pc.xy[0, :n], pc.xy[1, :n] = xb[:, 0], xb[:, 1]
    Lc.orc_wrap(pc.ref(), C.byref(cg), C.c_int(2))
    assert np.array_equal(pc.xy[:, :n].T, GOLD["ops_x_wrapped"])
    removed = Lc.orc_discard(pc.ref(), C.byref(cg), C.c_int(1))
    m = pc.np
    assert [m, removed] == GOLD["ops_np_after_discard"].tolist()
    assert np.array_equal(pc.id, GOLD["ops_id_after_discard"])
    assert np.array_equal(pc.xy[:, :m].T, GOLD["ops_x_after_discard"])
    V, u, dens = np.zeros(nx * ny), np.zeros(nx * ny), np.zeros(nx * ny)
    Lc.orc_cell_volume(C.byref(cg), CO.dp(V))
    assert np.array_equal(V, GOLD["ops_cell_volume"].ravel(order="F"))
    Lc.orc_deposit(C.byref(cg), pc.ref(), CO.dp(u))
    assert np.array_equal(u, GOLD["ops_deposit"].ravel(order="F"))        # same sequential order, no FMA
    Lc.orc_density(C.byref(cg), pc.ref(), CO.dp(V), CO.dp(dens))
    assert np.array_equal(dens, GOLD["ops_density"].ravel(order="F"))
    sig = np.array([Lc.orc_xsec_eval(CO.dp(np.ascontiguousarray(GOLD["xsec_nodes"][:, 0])),
                                     CO.dp(np.ascontiguousarray(GOLD["xsec_nodes"][:, 1])),
                                     C.c_int(len(GOLD["xsec_nodes"])), C.c_double(e)) for e in GOLD["xsec_eps"]])
    assert np.array_equal(sig, GOLD["xsec_sigma"])


# ------------------------------------------------------------------------------------ GPU ----
@pytest.fixture(scope="module")
def ib():
    import iskra_b200
    return iskra_b200


@pytest.mark.gpu
def test_device_reproduces_golden_operators(ib):
    PIC = ib.particle_in_cell
    nx, ny, dx = int(GOLD["ops_grid"][0]), int(GOLD["ops_grid"][1]), float(GOLD["ops_grid"][2])
    dt = float(GOLD["ops_dt"][0])
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dx)
    x0, v0, wg = GOLD["ops_x0"], GOLD["ops_v0"], GOLD["ops_wg"]
    n = len(wg)
    sp = PIC.create_kinetic_species("e-", n + 8, -O.qe, O.me, 1.37e5)
    sp.x[:n], sp.v[:n], sp.wg[:n], sp.np = x0, v0, wg, n
    i, j, hx, hy = PIC.particle_cell(sp, g)
    assert np.array_equal(i, GOLD["ops_i"]) and np.array_equal(j, GOLD["ops_j"])          # bit-exact contract
    assert np.array_equal(hx, GOLD["ops_hx"]) and np.array_equal(hy, GOLD["ops_hy"])
    pE = PIC.grid_to_particle(g, sp, GOLD["ops_E"])
    assert np.array_equal(pE, GOLD["ops_partE"])
    PIC.push_particles_(None, sp, pE, None, dt, g)
    assert np.array_equal(sp.x[:n], GOLD["ops_x_pushed"]) and np.array_equal(sp.v[:n], GOLD["ops_v_pushed"])
    sp.x[:n] = GOLD["ops_x_before_wrap"]
    PIC.wrap_(sp, g, dims=[2])
    assert np.array_equal(sp.x[:n], GOLD["ops_x_wrapped"])
    removed = PIC.discard_(sp, g, dims=[1])
    m = sp.np
    assert [m, removed] == GOLD["ops_np_after_discard"].tolist()
    # the device keeps the survivors in place instead of swapping from the end: compare keyed by id
    gid = GOLD["ops_id_after_discard"]
    assert sorted(sp.id.tolist()) == sorted(gid.tolist())
    og, od = np.argsort(gid[:m]), np.argsort(sp.id[:m])
    assert np.array_equal(sp.id[:m][od], gid[:m][og])
    assert np.array_equal(sp.x[:m][od], GOLD["ops_x_after_discard"][og])
    assert np.array_equal(ib.regular_grids.cell_volume(g), GOLD["ops_cell_volume"])
    dens = PIC.density(sp, g)
    ref = GOLD["ops_density"]
    assert np.abs(dens - ref).max() <= 1e-13 * np.abs(ref).max()       # summation order only


@pytest.mark.gpu
@pytest.mark.parametrize("sort_interval", [0, 2])
def test_device_reproduces_golden_rf_steps(ib, sort_interval):
    """Five iterations of the loop body with a driven electrode (apply_dirichlet inside the loop,
    11_rf_discharge.jl:95), discard x / wrap y, through the fused iskb_step path."""
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx, ny, dx = int(GOLD["rf_grid"][0]), int(GOLD["rf_grid"][1]), float(GOLD["rf_grid"][2])
    dt = float(GOLD["rf_dt"][0])
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = FDM.create_poisson_solver(g, O.eps0)
    FDM.apply_periodic(ps, 1)
    left, right = np.zeros((nx, ny), bool), np.zeros((nx, ny), bool)
    left[0, :], right[nx - 1, :] = True, True
    FDM.apply_dirichlet(ps, right, 0.0)
    FDM.apply_dirichlet(ps, left, float(GOLD["rf_volts"][0]))
    species = []
    for k, name in enumerate(("e-", "He+")):
        w0, m, q = GOLD["rf_weight_mass"][k]
        x0, v0 = GOLD["rf_x0_" + name], GOLD["rf_v0_" + name]
        sp = PIC.create_kinetic_species(name, len(x0) + 16, q, m, w0)
        sp.x[:len(x0)], sp.v[:len(x0)], sp.np = x0, v0, len(x0)
        species.append(sp)
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), species
    volts = GOLD["rf_volts"]

    def iteration(it, t, dt_):                       # fires after step `it`; sets the value of step it+1
        if it < len(volts):
            FDM.apply_dirichlet(ps, left, float(volts[it]))
    PIC.hooks.after_loop = iteration
    try:
        PIC.solve(cfg, dt, len(volts), after_push=(2, 1), sort_interval=sort_interval)
    finally:
        PIC.hooks.after_loop = lambda *a: None
    rho, phi, E = g._rt.fields()
    tol = 1e-10
    for sp in species:
        m = int(GOLD["rf_np_" + sp.name][0])
        assert sp.np == m
        og, od = np.argsort(GOLD["rf_id_" + sp.name]), np.argsort(sp.id[:m])
        assert np.array_equal(sp.id[:m][od], GOLD["rf_id_" + sp.name][og])
        xr, vr = GOLD["rf_x_" + sp.name][og], GOLD["rf_v_" + sp.name][og]
        assert np.abs(sp.x[:m][od] - xr).max() <= tol * (nx - 1) * dx
        assert np.abs(sp.v[:m][od] - vr).max() <= tol * np.abs(vr).max()
    assert np.abs(rho - GOLD["rf_rho"]).max() <= tol * np.abs(GOLD["rf_rho"]).max()
    assert np.abs(phi - GOLD["rf_phi"]).max() <= tol * np.abs(GOLD["rf_phi"]).max()
    assert np.abs(E - GOLD["rf_E"]).max() <= tol * np.abs(GOLD["rf_E"]).max()
