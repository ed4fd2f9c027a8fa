"""CPU tests: the C-ABI library loads and exports every symbol include/iskra_b200.h declares,
fails loudly without a GPU, and the host-side mirror logic (reactions parsing, lazy species
mirrors, workload parameters, synthetic tables) behaves like the reference's set-up code."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "iskra_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iskb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from iskra_b200 import _lib
    L = _lib.lib()
    declared = _header_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), "symbol %s declared in include/iskra_b200.h is not exported" % name
    # the ctypes table binds exactly the declared interface
    assert sorted(_lib.SIGNATURES) == declared
    assert L.iskb_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from iskra_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.iskb_create(0, C.byref(h))
    assert rc == _lib.E_CUDA and not h.value
    assert b"no CPU fallback" in L.iskb_last_error()
    with pytest.raises(_lib.IskraError):
        import iskra_b200
        iskra_b200.regular_grids.create_uniform_grid(np.arange(5.0), np.arange(5.0))


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (Salmon et al., SC'11 distribution kat_vectors)."""
    from iskra_b200 import _lib
    L = _lib.lib()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, expect in kats:
        out = (C.c_uint32 * 4)()
        L.iskb_debug_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == expect


def test_reactions_macro_semantics():
    from iskra_b200 import chemistry as CH, particle_in_cell as PIC
    e = PIC.create_kinetic_species("e-", 10, -1.0, 1.0, 2.0)
    iHe = PIC.create_kinetic_species("He+", 10, 1.0, 4.0, 2.0)
    He = PIC.FluidSpecies("He", 1.0, 0.0, 4.0, np.ones((3, 3)), 300.0)
    s = CH.CrossSection(np.array([[0.0, 1.0], [10.0, 2.0]]))
    rs = CH.reactions([(s, "e + He --> e + He"),
                       (s, "e + He --> e + He", CH.MCC.Excitation(19.82)),
                       (s, "e + He --> e + e + iHe", CH.MCC.Ionization(24.587))], {"e": e, "He": He, "iHe": iHe})
    assert [(r[0].name, r[1]) for r in rs[0].reactants] == [("e-", 1), ("He", 1)]
    assert rs[0].stoichiometry == []
    assert [(p.name, c) for p, c in rs[2].stoichiometry] == [("e-", 1), ("He", -1), ("He+", 1)]   # reactions.jl:39-51
    m = CH.mcc(rs)
    assert [c.type.kind for c in m.collisions] == [0, 3, 4]        # default ElasticIsotropic, mcc.jl:292
    assert m.collisions[2].products == [e, iHe] and m.collisions[2].source is e and m.collisions[2].target is He
    with pytest.raises(AssertionError):
        CH.mcc(CH.reactions([(s, "e + He + iHe --> e")], {"e": e, "He": He, "iHe": iHe}))
    with pytest.raises(ValueError):
        CH.mcc(CH.reactions([(s, "e + iHe --> e + iHe")], {"e": e, "iHe": iHe}))   # no fluid species


def test_species_host_mirror_unbound():
    from iskra_b200 import particle_in_cell as PIC
    sp = PIC.create_kinetic_species("e-", 100, -1.6e-19, 9.1e-31, 3.5e7)
    assert sp.np == 0 and sp.x.shape == (100, 2) and sp.v.shape == (100, 3)
    assert np.all(sp.wg == 3.5e7) and sp.w0 == 3.5e7                 # configuration.jl:99-100
    assert sp.id.tolist() == list(range(1, 101))                     # kinetic.jl:15
    sp.x[:5, 0] = 1.0
    sp.np = 5
    assert sp.np == 5 and sp.x[:5, 0].tolist() == [1.0] * 5
    with pytest.raises(RuntimeError):
        sp._push()                                                    # no grid -> no silent fallback


def test_synthetic_tables_reproduce_notebook_max_sigma_g():
    """max_sigma_g of the synthetic He tables through the oracle formula (mcc.jl:27-51) against
    the notebook constants (docs/capacitively_induced_discharge.ipynb:163)."""
    from iskra_b200 import datasets
    from oracle import pic_oracle as O
    e = O.KineticSpecies("e-", 1, -O.qe, O.me, 1.0)
    i = O.KineticSpecies("He+", 1, O.qe, 3.99 * O.mp, 1.0)
    He = O.FluidSpecies("He", 1.0, 0.0, 3.99 * O.mp, np.ones((2, 2)), 300.0)
    me_ = O.MonteCarloCollisions([O.Collision(0, O.CrossSection(t), e, He) for t in datasets.helium_electron()])
    mi_ = O.MonteCarloCollisions([O.Collision(0, O.CrossSection(t), i, He) for t in datasets.helium_ion()])
    assert me_.max_sigma_g == pytest.approx(8.976965143603543e-14, rel=5e-3, abs=0)
    assert mi_.max_sigma_g == pytest.approx(2.7462885393092625e-14, rel=5e-3, abs=0)
    for t in datasets.helium_electron() + datasets.helium_ion() + datasets.argon_electron():
        assert np.all(np.diff(t[:, 0]) > 0) and np.all(t[:, 1] >= 0)
    ext = [(t[0, 0], t[-1, 0]) for t in datasets.helium_electron()]
    assert ext == [(0.0, 965.0509), (19.82, 984.8709), (20.61, 985.6609), (24.59, 989.6379)]   # ipynb:145-156
    # the RF case keeps max_Pt <= 1/N (mcc.jl:244-246) and gives the notebook's candidate counts
    dt = 1.8436578171091445e-10
    assert 4 * (1 - math.exp(-9.64e20 * me_.max_sigma_g * dt)) * 16384 == pytest.approx(1037.3, rel=6e-3, abs=0)
    assert 2 * (1 - math.exp(-9.64e20 * mi_.max_sigma_g * dt)) * 16384 == pytest.approx(159.55, rel=6e-3, abs=0)


def test_sharding_slices():
    from sharding_helpers import slice_for_rank
    for n in (0, 1, 7, 1000, 125_000_001):
        for w in (1, 2, 3, 8):
            sl = [slice_for_rank(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


def test_circuit_mirror_matches_oracle_recurrence():
    """iskra_b200.circuit (host mirror of Circuit.jl) against the oracle restatement, 200 steps, bit for bit."""
    import math
    from iskra_b200 import circuit as CIR
    from oracle import surfaces_oracle as S
    V = lambda t: math.sin(2 * math.pi * 5e6 * t)
    cir = CIR.rlc(CIR.netlist([("V1", 3, "GND", V), ("L1", "NOD", "VCC", 1000e-9), ("C1", "NOD", "VCC", 1000e-9),
                               ("R1", "GND", "NOD", 1.0)]))                     # problem/06_circuit.jl:40-45
    ref = S.CircuitRLC(R=1.0, L=1000e-9, C=1000e-9, V=V)
    assert (cir.R, cir.L, cir.C) == (1.0, 1000e-9, 1000e-9) and isinstance(cir.ext, CIR.ShortedConnection)
    for _ in range(200):
        CIR.advance_circuit_(cir, 0, 10e-9)
        S.advance_circuit_(ref, 10e-9)
        assert cir.i == ref.i and cir.q == ref.q and cir.t == ref.t
    assert cir.probes["I1"] == cir.i and cir.probes["Vext"] == 0.0
    assert CIR.resonant_frequency(1e-6, 1e-6) == pytest.approx(1 / (2 * math.pi * 1e-6))
    with pytest.raises(ValueError):
        CIR.netlist([("X1", 1, 2, 3.0)])                                         # Circuit.jl:114 throw("unknown element")


def test_dsmc_constructor_semantics():
    """dsmc(reactions)  Chemistry/src/dsmc.jl:143-166 on the host mirror: two kinetic reactants, no products."""
    from iskra_b200 import chemistry as CH
    from iskra_b200 import particle_in_cell as PIC
    e = PIC.create_kinetic_species("e-", 10, -1.0, 1.0, 1.0)
    o = PIC.create_kinetic_species("O", 10, 0.0, 16.0, 1.0)
    io = PIC.create_kinetic_species("O+", 10, 1.0, 16.0, 1.0)
    sig = CH.CrossSection([3e6, 4e6, 5e6, 6e6], [0.01, 0.1, 2.0, 0.01])
    d = CH.dsmc(CH.reactions([(sig, "e + O --> O + e")], {"e": e, "O": o}))
    assert len(d.collisions) == 1 and d.collisions[0].source is e and d.collisions[0].target is o
    with pytest.raises(NotImplementedError):                                      # IonizationCollision has no perform! (:8-13)
        CH.dsmc(CH.reactions([(sig, "e + O --> e + e + iO")], {"e": e, "O": o, "iO": io}))
    with pytest.raises(NotImplementedError):                                      # quirk D1: one collision per object
        CH.dsmc(CH.reactions([(sig, "e + O --> O + e"), (sig, "e + O --> O + e")], {"e": e, "O": o}))
    with pytest.raises(AssertionError):                                           # :147 two reacting species
        CH.dsmc(CH.reactions([(sig, "O + O --> O + O")], {"O": o}))


def test_host_dense_inversion_hook():
    """The host-side Gauss-Jordan behind the dense Poisson path (csrc/poisson.cu invert_dense), without a GPU: a random
    well-conditioned matrix, the row-equilibrated 13_seed-like axial operator with Dirichlet plates, and a singular one."""
    import ctypes as C
    from iskra_b200 import _lib as L
    from oracle import axial_oracle as AX
    from oracle import pic_oracle as O
    fn = L.lib().iskb_debug_invert_dense
    fn.argtypes, fn.restype = [C.POINTER(C.c_double), C.c_int64], C.c_int32
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(0)
    M = np.asfortranarray(rng.standard_normal((200, 200)) + 200 * np.eye(200))
    A = M.copy(order="F")
    assert fn(dp(A), 200) == 0 and np.abs(A @ M - np.eye(200)).max() < 1e-13
    og = AX.AxialGrid2(np.arange(17) * 0.0025, np.arange(33) * 0.0025)
    ps = AX.PoissonSolver(og, 1.0)
    bot, top = np.zeros((17, 33), bool), np.zeros((17, 33), bool)
    bot[:, 0], top[:, 32] = True, True
    O.apply_dirichlet(ps, bot, 0.0)
    O.apply_dirichlet(ps, top, 1.0)
    As = np.asfortranarray(ps.A / np.abs(ps.A).max(axis=1)[:, None])
    B = As.copy(order="F")
    assert fn(dp(B), As.shape[0]) == 0 and np.abs(B @ As - np.eye(As.shape[0])).max() < 1e-12
    S = np.zeros((5, 5), order="F")
    assert fn(dp(S), 5) == L.E_SINGULAR


def _c_arity(decl_args):
    args = decl_args.strip()
    if args in ("", "void"):
        return 0
    depth, n = 0, 1
    for ch in args:
        depth += ch in "([" 
        depth -= ch in ")]"
        n += ch == "," and depth == 0
    return n


def test_header_ctypes_and_julia_glue_agree_on_names_and_arity():
    """Every entry point: the C declaration, the ctypes signature table and every `ccall` of the Julia glue must name
    an existing symbol and pass the same number of arguments (Julia is not installed here, so the glue is checked
    statically -- VERDICT r1 item 15)."""
    from iskra_b200 import _lib
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "iskra_b200.h")).read(), flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(iskb_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        decl[m.group(1)] = _c_arity(re.sub(r"\s+", " ", m.group(2)))
    assert sorted(decl) == sorted(_lib.SIGNATURES)
    for name, argtypes in _lib.SIGNATURES.items():
        assert len(argtypes) == decl[name], "%s: ctypes passes %d arguments, the header declares %d" % (name, len(argtypes), decl[name])
    jl = open(os.path.join(ROOT, "julia", "ParticleInCellB200.jl")).read()
    jl = re.sub(r"#[^\n]*", "", jl)
    seen = set()
    for m in re.finditer(r"ccall\(\(:(iskb_\w+),\s*LIB\),\s*(\w+),\s*\(", jl):
        name = m.group(1)
        assert name in decl, "Julia glue calls %s, which include/iskra_b200.h does not declare" % name
        # argument-type tuple: balanced parentheses from the match end
        i, depth = m.end(), 1
        while depth:
            depth += jl[i] == "("
            depth -= jl[i] == ")"
            i += 1
        tup = jl[m.end():i - 1].strip()
        n = 0 if tup == "" else _c_arity(tup.rstrip(","))
        assert n == decl[name], "Julia ccall of %s lists %d argument types, the header declares %d" % (name, n, decl[name])
        seen.add(name)
    assert len(seen) >= 40


def test_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the C oracle on the host cores) needs no GPU; stdout must carry ONE JSON line with the
    contract's keys, whatever libraries print while it runs (file descriptor 1 points at stderr until the line is out)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-particles", "100000"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
