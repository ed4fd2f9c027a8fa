"""CPU tests of the secondary-electron-emission restatement (oracle/see_oracle.py, Chemistry/src/see.jl).
The reference holds no test or stored output for see.jl; the oracle is pinned by the closed-form answers its formulas
have (see.jl:17-58) and by the invariants emit! keeps (see.jl:114-181)."""
import math

import numpy as np
import pytest

from oracle import pic_oracle as O
from oracle import see_oracle as S


def _grid(nx=33, ny=17, dx=1e-3):
    return O.CartesianGrid2(np.arange(nx) * dx, np.arange(ny) * dx)


def test_vaughan_known_answers():
    gv = S.vaughan(13.0, 500.0, 3.0, 1.0)
    assert gv(500.0, 0.0) == pytest.approx(3.0, rel=1e-15)          # v = 1: gamma_max at w_max
    assert gv(13.0, 0.0) == 0.0 and gv(1.0, 0.3) == 0.0             # below w0
    th = 1.0                                                         # grazing: sin(theta) = 1
    wmax, gmax = 500.0 * (1 + 1 / math.pi), 3.0 * (1 + 1 / (2 * math.pi))
    assert gv(wmax, th) == pytest.approx(gmax, rel=1e-15)
    # the exponent switches from 0.62 to 0.25 above w_max (continuous there, both branches give gmax)
    v = (2000.0 - 13.0) / (500.0 - 13.0)
    assert gv(2000.0, 0.0) == pytest.approx(3.0 * (v * math.exp(1 - v)) ** 0.25, rel=1e-15)
    v = (100.0 - 13.0) / (500.0 - 13.0)
    assert gv(100.0, 0.0) == pytest.approx(3.0 * (v * math.exp(1 - v)) ** 0.62, rel=1e-15)


def test_elastic_inelastic_secondary_known_answers():
    d = S.defaults()
    gv, ge, gi, gt = d["gv"], d["ge"], d["gi"], d["gt"]
    assert ge(2.0, 0.0) == 0.0 and ge(1.0, 0.0) == 0.0                 # w <= we
    assert ge(10.0, 0.0) == pytest.approx(0.55, rel=1e-15)             # v1 = 1 at wemax (gv = 0 below w0 = 13)
    lo, hi = ge(10.0, 0.0), ge(10.0 + 1e-9, 0.0)
    assert abs(lo - hi) < 1e-9                                         # continuous across wemax
    assert ge(23.0, 0.2) == pytest.approx(0.03 * gv(23.0, 0.2) + 0.55 * (1 + 1.0) * math.exp(-1.0), rel=1e-15)
    assert gi(300.0, 0.1) == pytest.approx(0.07 * gv(300.0, 0.1), rel=1e-15)
    assert gt(300.0, 0.1) == pytest.approx(0.90 * gv(300.0, 0.1), rel=1e-15)
    assert S.gamma0(5.0, 0.5) == 0.0


def test_true_secondary_energy_is_lognormal():
    rng = np.random.default_rng(0)
    le = np.log([S.true_secondary_energy(rng) for _ in range(20000)])
    assert abs(le.mean() - 1.65) < 0.03 and abs(le.std() - 1.1) < 0.03


def _beyond(sp, g, wall, n, rng, speed=4e6):
    Lx, Ly = (g.n[0] - 1) * g.dh[0], (g.n[1] - 1) * g.dh[1]
    d = rng.random(n) * 0.4 * g.dh[0] + 1e-6 * g.dh[0]
    sp.x[:n, 0] = rng.random(n) * Lx
    sp.x[:n, 1] = rng.random(n) * Ly
    sp.v[:n] = rng.standard_normal((n, 3)) * speed
    if wall == "right":
        sp.x[:n, 0], sp.v[:n, 0] = Lx + d, np.abs(sp.v[:n, 0]) + 1e3
    elif wall == "left":
        sp.x[:n, 0], sp.v[:n, 0] = -d, -np.abs(sp.v[:n, 0]) - 1e3
    elif wall == "top":
        sp.x[:n, 1], sp.v[:n, 1] = Ly + d, np.abs(sp.v[:n, 1]) + 1e3
    else:
        sp.x[:n, 1], sp.v[:n, 1] = -d, -np.abs(sp.v[:n, 1]) - 1e3
    sp.np = n
    return d


@pytest.mark.parametrize("wall", ["right", "top"])
def test_elastic_reflection_mirrors_the_row_at_an_upper_wall(wall):
    g = _grid()
    rng = np.random.default_rng(1)
    e, s = O.KineticSpecies("e-", 64, -O.qe, O.me, 1.0), O.KineticSpecies("s", 64, -O.qe, O.me, 1.0)
    d = _beyond(e, g, wall, 40, rng)
    e.np = 50                                                        # ten rows inside the domain: untouched
    e.x[40:50] = [[0.5 * 32e-3, 0.5 * 16e-3]] * 10
    x0, v0 = e.x.copy(), e.v.copy()
    c = S.emit_(e, s, g, wall, gt=S.gamma0, ge=lambda w, th: 2.0, rng=np.random.default_rng(2))
    assert c == {"elastic": 40, "inelastic": 0, "secondaries": 0, "absorbed": 0}
    i = 0 if wall == "right" else 1
    L = (g.n[i] - 1) * g.dh[i]
    assert np.allclose(e.x[:40, i], L - d, rtol=0, atol=1e-15)       # mirrored about the wall
    assert np.array_equal(e.v[:40, i], -v0[:40, i])
    o = 1 - i
    assert np.allclose(e.x[:40, o], x0[:40, o], rtol=0, atol=1e-15) and np.array_equal(e.v[:40, o], v0[:40, o])
    assert np.array_equal(e.x[40:], x0[40:]) and e.np == 50 and s.np == 0


def test_lower_wall_uses_mod_of_the_position_as_the_reference_does():
    """see.jl:141: dt = mod(x, L)/|v| -- at a lower wall that is (L - d)/|v|, not d/|v| (restated, not repaired)."""
    g = _grid()
    e, s = O.KineticSpecies("e-", 4, -O.qe, O.me, 1.0), O.KineticSpecies("s", 4, -O.qe, O.me, 1.0)
    e.x[0], e.v[0], e.np = [-1e-4, 5e-3], [-2e6, 0.0, 0.0], 1
    S.emit_(e, s, g, "left", gt=S.gamma0, ge=lambda w, th: 2.0, rng=np.random.default_rng(0))
    L = 32e-3
    dt = (L - 1e-4) / 2e6
    assert e.v[0, 0] == 2e6
    assert e.x[0, 0] == pytest.approx((-1e-4 + 2e6 * dt) + 2e6 * dt, rel=1e-14)


def test_yield_above_one_emits_floor_plus_bernoulli_and_keeps_the_primary_when_it_emits():
    g = _grid()
    rng = np.random.default_rng(3)
    n = 4000
    e, s = O.KineticSpecies("e-", n, -O.qe, O.me, 1.0), O.KineticSpecies("s", 4 * n, -O.qe, O.me, 1.0)
    _beyond(e, g, "right", n, rng)
    c = S.emit_(e, s, g, "right", gt=lambda w, th: 2.5, rng=np.random.default_rng(4))
    assert c["elastic"] == 0 and c["inelastic"] == 0
    third = c["secondaries"] - 2 * n
    assert third + c["absorbed"] == n                                 # R2 < 0.5: a third secondary, else absorbed
    assert abs(third - n / 2) < 5 * math.sqrt(n / 4)
    assert e.np == n - c["absorbed"] and s.np == c["secondaries"]
    assert sorted(e.id.tolist()) == list(range(1, n + 1))             # remove! keeps ids a permutation
    # diffuse_reflection (mcc.jl:64-72, 122-127) has cos(chi) = -sqrt(1 - R): back into the domain, cosine law about -n
    cosn = s.v[: s.np, 0] / np.linalg.norm(s.v[: s.np], axis=1)
    assert cosn.max() <= 0.0 and abs(cosn.mean() + 2.0 / 3.0) < 0.01
    eps = 0.5 * (O.me / S.QE_MCC) * np.sum(s.v[: s.np] ** 2, axis=1)
    assert abs(np.log(eps).mean() - 1.65) < 0.03


def test_inelastic_branch_scales_the_reflected_velocity():
    g = _grid()
    rng = np.random.default_rng(5)
    n = 2000
    e, s = O.KineticSpecies("e-", n, -O.qe, O.me, 1.0), O.KineticSpecies("s", n, -O.qe, O.me, 1.0)
    _beyond(e, g, "top", n, rng)
    v0 = e.v.copy()
    c = S.emit_(e, s, g, "top", gt=S.gamma0, ge=S.gamma0, gi=lambda w, th: 2.0, rng=np.random.default_rng(6))
    assert c["inelastic"] == n
    f = -e.v[:n, 1] / v0[:n, 1]
    assert f.min() >= 0.0 and f.max() < 1.0 and abs(f.mean() - 0.5) < 0.03
    assert np.allclose(e.v[:n, 0], f * v0[:n, 0], rtol=1e-14) and np.allclose(e.v[:n, 2], f * v0[:n, 2], rtol=1e-14)


def test_host_mirror_rejects_what_the_c_abi_cannot_carry():
    import iskra_b200.chemistry as CH
    with pytest.raises(NotImplementedError):
        CH.emit_(None, None, None, boundary="all", gamma_t=CH.gamma_t)
    with pytest.raises(TypeError):
        CH.emit_(None, None, None, boundary="left", gamma_t=lambda w, th: 0.0)
    other = CH.Vaughan(10.0, 400.0, 2.0)
    with pytest.raises(NotImplementedError):
        CH.emit_(None, None, None, boundary="left", gamma_t=CH.gamma_t, gamma_e=CH.Elastic(other, 2.0, 10.0, 0.5))
    assert (CH.gamma_v.w0, CH.gamma_v.w0max, CH.gamma_v.g0max, CH.gamma_v.ks) == (13.0, 500.0, 3.0, 1.0)   # see.jl:60
    assert (CH.gamma_e.we, CH.gamma_e.wemax, CH.gamma_e.gemax, CH.gamma_e.De, CH.gamma_e.re) == (2.0, 10.0, 0.55, 13.0, 0.03)
    assert CH.gamma_i.ri == 0.07 and (CH.gamma_t.re, CH.gamma_t.ri) == (0.03, 0.07)
