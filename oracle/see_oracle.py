"""TEST INFRASTRUCTURE: numpy restatement of the reference's secondary-electron emission at a wall,
Chemistry/src/see.jl.  Only tests/ may import this.

Line by line, quirks included:
  * the emission coefficients are closures over their parameters (see.jl:17-58; defaults :60-63);
  * "theta" is the SINE of the angle between the velocity and the wall normal (:133);
  * dt = mod(x_i, L_i) / |v_i| (:141): the time since the crossing at an upper wall, but (L + x)/|v| at a lower one;
  * an emitted secondary does not remove the primary (:165-168): only `R2 >= gamma` absorbs it (:169);
  * boundary = :all has the zero vector as normal (:94): specular reflection is the identity and
    diffuse_reflection([0,0,0]) is NaN -- not restated, the device path rejects it as well.
Parity unpinned: the reference holds no test, fixture or stored output for see.jl, and every branch draws from Julia's
MersenneTwister / Distributions.LogNormal; pinned by known answers instead (tests/test_see.py).
"""
import math

import numpy as np

from . import pic_oracle as O

QE_MCC = 1.60217646e-19     # see.jl:65


def true_secondary_energy(rng):
    """:11-14  rand(LogNormal(1.65, 1.1))  [eV]"""
    return float(rng.lognormal(1.65, 1.1))


def vaughan(w0, w0max, g0max, ks=0.0):
    """:17-26"""
    wmax = lambda th: w0max * (1.0 + ks / math.pi * th ** 2)
    gmax = lambda th: g0max * (1.0 + ks / (2 * math.pi) * th ** 2)
    v = lambda w, th: (w - w0) / (wmax(th) - w0) if w > w0 else 0.0

    def f(w, th):
        k = 0.25 if w > wmax(th) else 0.62
        return gmax(th) * (v(w, th) * math.exp(1.0 - v(w, th))) ** k
    return f


def elastic(gv, we, wemax, gemax, De=13.0, re=0.03):
    """:29-42"""
    v1 = lambda w: (w - we) / (wemax - we)
    v2 = lambda w: (w - wemax) / De

    def f(w, th):
        if we < w <= wemax:
            return re * gv(w, th) + gemax * v1(w) * math.exp(1.0 - v1(w))
        if w > wemax:
            return re * gv(w, th) + gemax * (1.0 + v2(w)) * math.exp(-v2(w))
        return 0.0
    return f


def inelastic(gv, ri=0.07):
    """:45-49"""
    return lambda w, th: ri * gv(w, th)


def secondary(gv, re, ri):
    """:52-56"""
    return lambda w, th: (1.0 - re - ri) * gv(w, th)


def defaults():
    """:60-63"""
    gv = vaughan(13.0, 500.0, 3.0, 1.0)
    return {"gv": gv, "ge": elastic(gv, 2.0, 10.0, 0.55, re=0.03), "gi": inelastic(gv, ri=0.07),
            "gt": secondary(gv, re=0.03, ri=0.07)}


gamma0 = lambda w, th: 0.0     # :103

_NORMAL = {"left": (-1.0, 0.0, 0.0), "right": (1.0, 0.0, 0.0), "top": (0.0, 1.0, 0.0), "bottom": (0.0, -1.0, 0.0)}   # :80-88
_LOWER = {"left", "bottom"}                                                                                          # :71-79
_DIM = {"left": 0, "right": 0, "top": 1, "bottom": 1}                                                                # :91-99


def snells_law(v, n):
    """:105"""
    return v - 2.0 * float(np.dot(n, v)) * n


def inject_secondary_(sec, x, nhat, dt, rng):
    """:107-117"""
    m = sec.m / QE_MCC
    eps = true_secondary_energy(rng)
    v = O.diffuse_reflection(nhat, rng) * math.sqrt(2.0 * eps / m)
    sec.v[sec.np, :] = v
    sec.x[sec.np, :] = dt * v[:2] + x
    sec.np += 1


def emit_(primary, sec, grid, boundary, gt, ge=gamma0, gi=gamma0, rng=None, counts=None):
    """emit!(primary, secondary, grid, material; boundary, gamma_t, gamma_e, gamma_i)  :114-181 for one wall.
    counts (optional dict) receives the number of elastic / inelastic reflections, injected secondaries, absorptions."""
    nhat = np.array(_NORMAL[boundary])
    i = _DIM[boundary]
    lower = boundary in _LOWER
    L = (grid.n[i] - 1) * grid.dh[i]
    ox = grid.origin[i]
    c = counts if counts is not None else {}
    for key in ("elastic", "inelastic", "secondaries", "absorbed"):
        c.setdefault(key, 0)
    mp = primary.m / QE_MCC
    for p in range(primary.np, 0, -1):                          # reverse(1:np)
        a = float(O.jl_fld(primary.x[p - 1, i] - ox, L))
        if not (a < 0.0 if lower else a > 0.0):
            continue
        pv = primary.v[p - 1]
        th = float(np.linalg.norm(np.cross(pv, nhat)) / np.linalg.norm(pv))     # :133
        w = 0.5 * mp * float(np.dot(pv, pv))
        R1 = rng.random()
        g_e, g_i, g_t = ge(w, th), gi(w, th), gt(w, th)
        xp = primary.x[p - 1]
        dt = float(O.jl_mod(xp[i], L)) / abs(pv[i])             # :141
        x0 = xp - pv[:2] * dt
        if g_e + g_i > R1 > g_e:                                # :146-153 inelastic
            xp -= pv[:2] * dt
            pv[:] = rng.random() * snells_law(pv, nhat)
            xp += pv[:2] * dt
            c["inelastic"] += 1
            continue
        if R1 < g_e:                                            # :155-162 elastic
            xp -= pv[:2] * dt
            pv[:] = snells_law(pv, nhat)
            xp += pv[:2] * dt
            c["elastic"] += 1
            continue
        g = g_e + g_i + g_t
        while g > 1.0:                                          # :165-169
            inject_secondary_(sec, x0, nhat, dt, rng)
            c["secondaries"] += 1
            g -= 1.0
        R2 = rng.random()
        if R2 < g:
            inject_secondary_(sec, x0, nhat, dt, rng)
            c["secondaries"] += 1
        else:
            O.remove_(primary, p)                               # :176 absorb
            c["absorbed"] += 1
    return c
