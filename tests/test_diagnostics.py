"""SURVEY.md 8f row N2 -- the Diagnostics mirror: openPMD record names / paths / attributes as the reference writes them
(Diagnostics/src/hdf5.jl:47-95, openpmd/*.jl), records fetched on demand.  CPU: registry + sink with plain arrays;
GPU: the records of a running solve() equal the device state."""
import numpy as np
import pytest

from iskra_b200 import diagnostics as DG


class _G:
    n, dh = (4, 3), (0.5, 0.25)


class _S:
    name, np, m, q = "e-", 5, 2.0, -1.0


@pytest.mark.parametrize("sink", ["h5", "npz"])
def test_registry_paths_and_attributes(tmp_path, sink, monkeypatch):
    monkeypatch.setattr(DG, "SINK", sink)
    DG.records.clear()
    calls = []
    rho = np.arange(12.0).reshape(4, 3)
    E = np.arange(36.0).reshape(4, 3, 3)

    def fetch_rho():
        calls.append("rho")
        return rho
    DG.register_field("rho", "C/m^2", fetch_rho, _G())
    DG.register_field("E", "V/m", lambda: E, _G(), withcomponents=True)
    DG.register_particle("e-/position", "m", lambda: np.arange(10.0).reshape(5, 2), _S(), withcomponents=True)
    DG.register_particle("e-/mass", "kg", lambda: np.array([2.0]), _S())
    DG.register_particle("e-/weighting", "1", lambda: np.ones(5), _S(), weighted=True)
    assert calls == []                                            # registering copies nothing

    def save(it):
        DG.save_record(it, "rho")
        DG.save_record(it, "E")
        DG.save_records(it, "e-/")
        DG.save_record(it, "missing")
    path = DG.new_iteration(str(tmp_path / "run"), 7, 0.5, 0.1, save)
    assert calls == ["rho"]                                       # fetched exactly once, when saved
    arrays, attrs = DG.load(path)
    import os
    assert path.endswith("hdf5/data7") and os.path.exists(path + "." + sink)
    assert np.array_equal(arrays["data/7/fields/rho"], rho)
    assert np.array_equal(arrays["data/7/fields/E/y"], E[:, :, 1])
    assert np.array_equal(arrays["data/7/particles/e-/position/x"], np.arange(10.0).reshape(5, 2)[:, 0])
    assert np.array_equal(arrays["data/7/particles/e-/position/z"], np.zeros(5))       # D = 2: z component is zeros
    assert "data/7/particles/e-/mass" not in arrays                                    # constant record: attributes only
    assert attrs["data/7/particles/e-/mass"]["value"] == 2.0 and attrs["data/7/particles/e-/mass"]["shape"] == [5]
    assert attrs["data/7/particles/e-/weighting"]["macroWeighted"] == 1
    assert attrs[""]["openPMD"] == "1.1.0" and attrs[""]["meshesPath"] == "fields/" and attrs[""]["particlesPath"] == "particles/"
    assert attrs["data/7"] == {"dt": 0.1, "time": 0.5, "timeUnitSI": 1.0}
    assert attrs["data/7/fields/rho"]["unitDimension"] == [-2.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0]      # C/m^2 = m^-2 s A
    assert attrs["data/7/fields/rho"]["gridSpacing"] == [0.5, 0.25] and attrs["data/7/fields/rho"]["axisLabels"] == "xy"
    assert attrs["data/7/fields/E/x"]["unitSI"] == 1.0 and attrs["data/7/fields"]["fieldSolverParameters"] == "Nagel"
    assert attrs["data/7/fields"]["fieldBoundary"] == ["open"] * 4
    assert attrs["data/7/particles"]["particlePush"] == "Boris"


@pytest.mark.gpu
def test_records_of_a_running_solve_equal_the_device_state(tmp_path):
    import iskra_b200 as ib
    from oracle import pic_oracle as O
    PIC, FDM = ib.particle_in_cell, ib.finite_difference_method
    nx, ny, dh, dt, n = 33, 17, 1e-3, 1e-10, 4000
    g = ib.regular_grids.create_uniform_grid(np.arange(nx) * dh, np.arange(ny) * dh)
    ps = FDM.create_poisson_solver(g, O.eps0)
    FDM.apply_periodic(ps, 1)
    FDM.apply_periodic(ps, 2)
    rng = np.random.default_rng(1)
    e = PIC.create_kinetic_species("e-", n + 10, -O.qe, O.me, 1e6)
    e.x[:n] = rng.random((n, 2)) * np.array([(nx - 1) * dh, (ny - 1) * dh])
    e.v[:n] = rng.standard_normal((n, 3)) * 1e5
    e.np = n
    cfg = ib.configuration.Config()
    cfg.grid, cfg.solver, cfg.pusher, cfg.species = g, ps, PIC.create_boris_pusher(), [e]
    saved = []

    def after_loop(i, t, dt_):
        if i != 3:
            return

        def save(it):
            DG.save_records(it, "e-/")
            for k in ("rho", "phi", "E", "ne-"):
                DG.save_record(it, k)
        saved.append(DG.new_iteration(str(tmp_path / "run"), i, t, dt_, save))
    PIC.hooks.after_loop = after_loop
    try:
        PIC.solve(cfg, dt, 3, after_push=(ib._lib.BND_WRAP, ib._lib.BND_WRAP))
    finally:
        PIC.hooks.after_loop = lambda *a: None
    arrays, attrs = DG.load(saved[0])
    rho, phi, E = g._rt.fields()
    b = "data/3/"
    assert np.array_equal(arrays[b + "fields/rho"], rho) and np.array_equal(arrays[b + "fields/phi"], phi)
    assert np.array_equal(arrays[b + "fields/E/x"], E[:, :, 0]) and np.array_equal(arrays[b + "fields/E/y"], E[:, :, 1])
    assert np.array_equal(arrays[b + "particles/e-/position/x"], e.x[:n, 0])
    assert np.array_equal(arrays[b + "particles/e-/momentum/y"], O.me * e.v[:n, 1])
    assert np.array_equal(np.sort(arrays[b + "particles/e-/id"]), np.arange(1, n + 1))
    assert attrs[b + "particles/e-/charge"]["value"] == -O.qe and attrs[b + "particles/e-/mass"]["shape"] == [n]
    assert arrays[b + "fields/ne-"].shape == (nx, ny) and arrays[b + "fields/ne-"].sum() > 0


def test_probe_records_and_small_mirrors(tmp_path):
    """@probe (Diagnostics/src/circuit.jl), create_staggered_grid (RegularGrids.jl:99-108), DensitySource / create_fluid_species
    (sources.jl:3-6,36-38; configuration.jl:104-108): host-side pieces of problem/06_circuit.jl."""
    from iskra_b200 import circuit as CIR
    from iskra_b200 import configuration as CFG
    from iskra_b200 import particle_in_cell as PIC
    from iskra_b200 import regular_grids as RG
    DG.records.clear()
    cir = CIR.rlc(CIR.netlist([("V1", 3, "GND", lambda t: 2.0), ("L1", "N", "V", 1e-6), ("C1", "N", "V", 1e-6), ("R1", "G", "N", 1.0)]))
    CIR.advance_circuit_(cir, None, 1e-8)
    assert set(DG.records) >= {"Q1", "I1", "V1", "Vext"}
    path = DG.new_iteration(str(tmp_path / "run"), 1, 1e-8, 1e-8, lambda it: [DG.save_record(it, k) for k in ("Q1", "I1", "V1", "Vext")])
    arrays, attrs = DG.load(path)
    assert arrays["data/1/fields/V1"].shape == (1, 1) and arrays["data/1/fields/V1"][0, 0] == 2.0
    assert arrays["data/1/fields/I1"][0, 0] == cir.i
    assert attrs["data/1/fields/I1"]["unitDimension"] == [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0] and attrs["data/1/fields/I1"]["geometry"] == "cartesian"

    class G:
        origin, dh, n, bcs = (0.0, 1.0), (0.5, 0.25), (5, 3), None
    c = RG.create_staggered_grid(G())
    assert c.n == (6, 4) and c.origin == (-0.25, 0.875) and np.isclose(c.coords[0][-1, 0], 2.25) and np.isclose(c.coords[1][0, -1], 1.625)
    O_ = CFG.create_fluid_species("O", 1.0, 0.0, 8.0, 4, 3)
    PIC.init(PIC.DensitySource(2.0 * np.ones((4, 3)), None), O_, 1e-8)
    PIC.init(PIC.DensitySource(0.5 * np.ones((4, 3)), None), O_, 1e-8)
    assert O_.n.shape == (4, 3) and np.all(O_.n == 2.5) and O_.T == 300.0
