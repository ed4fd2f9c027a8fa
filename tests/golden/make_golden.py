"""Generates tests/golden/pic_golden.npz and reference_notebook_values.json.

The reference is Julia and cannot be run here (no julia binary), so these vectors come
from the numpy restatement oracle/pic_oracle.py -- which is itself pinned on the reference's
known answers (tests/test_oracle.py).  They freeze the oracle: the C oracle, later edits of the numpy
oracle and the CUDA path are all checked against the same committed numbers.

    python tests/golden/make_golden.py        # rewrites both files; they are deterministic
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pic_oracle as O   # noqa: E402


def refined_solve(A, b, dx):
    """Exact solution of the reference's system A x = b (row-equilibrated LU + extended-precision
    refinement; DESIGN.md section 2 explains why the one-shot LU is not the comparison target)."""
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


def operators_case(out):
    """Every per-particle operator once, on a 17x9 grid with 400 seeded particles."""
    nx, ny, dx = 17, 9, 5.234375e-4
    g = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    rng = np.random.default_rng(20260101)
    n = 400
    sp = O.KineticSpecies("e-", n + 8, -O.qe, O.me, 1.37e5)
    sp.x[:n, 0] = rng.random(n) * (nx - 1) * dx
    sp.x[:n, 1] = rng.random(n) * (ny - 1) * dx
    sp.x[:8, 0] = np.arange(8) * dx                       # particles exactly on nodes
    sp.x[8:12, 1] = np.nextafter((ny - 1) * dx, 0.0)      # just inside the upper edge
    sp.v[:n] = rng.standard_normal((n, 3)) * O.thermal_speed(30000.0, O.me)
    sp.wg[:n] = 1.37e5 * (1.0 + 0.25 * rng.random(n))
    sp.np = n
    E = rng.standard_normal((nx, ny, 3)) * 2.0e3
    E[:, :, 2] = 0.0
    dt = 1.8436578171091445e-10
    out["ops_grid"] = np.array([nx, ny, dx])
    out["ops_dt"] = np.array([dt])
    out["ops_x0"], out["ops_v0"], out["ops_wg"] = sp.x[:n].copy(), sp.v[:n].copy(), sp.wg[:n].copy()
    out["ops_E"] = E
    i, j, hx, hy = O.particle_cell(sp.x[:n], g.dh)
    out["ops_i"], out["ops_j"], out["ops_hx"], out["ops_hy"] = i, j, hx, hy
    pE = O.grid_to_particle(g, sp, E)
    out["ops_partE"] = pE
    O.push_in_cartesian_(sp, pE, dt)
    out["ops_x_pushed"], out["ops_v_pushed"] = sp.x[:n].copy(), sp.v[:n].copy()
    far = sp.x[:n].copy()                                   # exercise several periods and both signs
    far[:40, 1] += np.linspace(-7.5, 7.5, 40) * (ny - 1) * dx
    sp.x[:n] = far
    out["ops_x_before_wrap"] = far.copy()
    O.wrap_(sp, g, dims=(2,))
    out["ops_x_wrapped"] = sp.x[:n].copy()
    removed = O.discard_(sp, g, dims=(1,))
    m = sp.np
    out["ops_np_after_discard"] = np.array([m, removed])
    out["ops_id_after_discard"] = sp.id.copy()
    out["ops_x_after_discard"] = sp.x[:m].copy()
    out["ops_cell_volume"] = O.cell_volume(g)
    out["ops_deposit"] = O.particle_to_grid(sp, g, sp.wg[:m])
    out["ops_density"] = O.density(sp, g)


def rf_steps_case(out):
    """Five iterations of the loop body (ParticleInCell.jl:102-135) without MCC on a C2-like case:
    0 V / driven electrode in x, 'periodic' in y, discard!(dims=1) + wrap!(dims=2)."""
    nx, ny, dx = 33, 9, 5.234375e-4
    dt = 1.8436578171091445e-10
    g = O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)
    ps = O.PoissonSolver(g, O.eps0)
    O.apply_periodic(ps, 1)
    left, right = np.zeros((nx, ny), bool), np.zeros((nx, ny), bool)
    left[0, :], right[nx - 1, :] = True, True
    O.apply_dirichlet(ps, right, 0.0)
    rng = np.random.default_rng(7)
    n = 3000
    species = []
    for name, q, m, T in (("e-", -O.qe, O.me, 30000.0), ("He+", O.qe, 4.002602 * 1.66053906660e-27, 300.0)):
        sp = O.KineticSpecies(name, n + 16, q, m, 2.0e7)
        sp.x[:n, 0] = rng.random(n) * (nx - 1) * dx
        sp.x[:n, 1] = rng.random(n) * (ny - 1) * dx
        sp.v[:n] = rng.standard_normal((n, 3)) * O.thermal_speed(T, m)
        sp.np = n
        species.append(sp)
        out["rf_x0_" + name], out["rf_v0_" + name] = sp.x[:n].copy(), sp.v[:n].copy()
    out["rf_grid"] = np.array([nx, ny, dx])
    out["rf_dt"] = np.array([dt])
    out["rf_weight_mass"] = np.array([[sp.w0, sp.m, sp.q] for sp in species])
    E = np.zeros((nx, ny, 3))
    volts = []
    after_push = lambda part, grid: (O.discard_(part, grid, dims=(1,)), O.wrap_(part, grid, dims=(2,)))
    for it in range(5):
        v = 20.0 * np.sin(2 * np.pi * 13.56e6 * it * dt) + 5.0 * it      # electrode value of this step
        volts.append(v)
        O.apply_dirichlet(ps, left, v)                                    # 11_rf_discharge.jl:95
        for sp in species:
            O.advance_(sp, E, dt, g, after_push)
        rho = np.zeros(g.n)
        for sp in species:
            sp.n = O.density(sp, g)
            rho += sp.n * sp.q
        ff = (-rho).reshape(-1, order="F")
        rd = np.asarray(ps.rho_dof, dtype=np.int64)
        ps.b[rd] = ff[rd] / ps.eps0
        phi = refined_solve(ps.A, ps.b, dx)[ps.phi_dof]
        E = O.calculate_electric_field(ps, phi)
    out["rf_volts"] = np.array(volts)
    out["rf_rho"], out["rf_phi"], out["rf_E"] = rho, phi, E
    for sp in species:
        m = sp.np
        out["rf_np_" + sp.name] = np.array([m])
        out["rf_id_" + sp.name] = sp.id[:m].copy()
        out["rf_x_" + sp.name], out["rf_v_" + sp.name] = sp.x[:m].copy(), sp.v[:m].copy()


def xsec_case(out):
    """sigma(eps): piecewise linear, flat outside (cross_section.jl:8-14) on an irregular table."""
    nodes = np.array([[0.0, 1.0e-20], [0.5, 3.0e-20], [0.75, 2.0e-20], [4.0, 8.0e-20], [19.8, 1.0e-21], [1000.0, 5.0e-22]])
    eps = np.array([-1.0, 0.0, 0.25, 0.5, 0.6, 0.75, 3.999, 4.0, 10.0, 19.8, 500.0, 1000.0, 2.0e4])
    out["xsec_nodes"], out["xsec_eps"], out["xsec_sigma"] = nodes, eps, O.CrossSection(nodes)(eps)


NOTEBOOK = {
    "_source": "numbers printed by the reference's notebooks (the only stored outputs of this path)",
    "m_eV_electron": [5.685630721038056e-12, "docs/capacitively_induced_discharge.ipynb:145-163"],
    "m_eV_He_ion": [4.165434669922687e-08, "docs/capacitively_induced_discharge.ipynb:145-163"],
    "half_thermal_speed_Te": [476807.16512475203, "docs/capacitively_induced_discharge.ipynb cell 2"],
    "candidates_electron_step1": [1037.0, "docs/capacitively_induced_discharge.ipynb:254 (np=16384, N=4)"],
    "candidates_ion_step1": [159.0, "docs/capacitively_induced_discharge.ipynb:255 (N=2)"],
    "max_sigma_g_electron": [8.976965143603543e-14, "docs/capacitively_induced_discharge.ipynb:145-163 (needs LXCat tables: synthetic tables are tuned to it)"],
    "max_sigma_g_ion": [2.7462885393092625e-14, "docs/capacitively_induced_discharge.ipynb:145-163"],
    "capacitor_test": ["3x5 nodes, dx=0.5, phi(i=1)=0, phi(i=3)=1 => Ex == -1, Ey == 0 (atol 1e-15)", "FiniteDifferenceMethod/test/runtests.jl:8-22"],
}


def main():
    out = {}
    operators_case(out)
    rf_steps_case(out)
    xsec_case(out)
    np.savez_compressed(os.path.join(HERE, "pic_golden.npz"), **out)
    with open(os.path.join(HERE, "reference_notebook_values.json"), "w") as f:
        json.dump(NOTEBOOK, f, indent=1)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
