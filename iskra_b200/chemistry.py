"""Host mirror of the Chemistry module API for the MCC path (Chemistry/src/{mcc,cross_section,
reactions}.jl).  Set-up stays on the host (as it stays Julia in the reference); perform! runs on
the device."""
import ctypes as C
import re

import numpy as np

from . import _lib as L
from .particle_in_cell import is_fluid


class CrossSection:
    """CrossSection(nodes)  cross_section.jl:3-14: sigma(eps) table, piecewise linear, Flat() outside."""

    def __init__(self, nodes, ys=None):
        if ys is not None:
            nodes = np.stack([np.asarray(nodes, dtype=np.float64), np.asarray(ys, dtype=np.float64)], axis=1)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        if self.nodes.ndim != 2 or self.nodes.shape[1] != 2 or len(self.nodes) < 2:
            raise ValueError("CrossSection needs an (n>=2, 2) table")

    def __call__(self, x):
        xs, ys = self.nodes[:, 0], self.nodes[:, 1]
        return np.interp(x, xs, ys)          # host convenience only; the device evaluates its own copy

    def maximum(self):
        return float(self.nodes[:, 1].max())


class MCC:
    """collision type tags  mcc.jl:4-8"""

    class ElasticIsotropic:
        kind, energy = L.MCC_ELASTIC_ISOTROPIC, 0.0

    class ElasticBackward:
        kind, energy = L.MCC_ELASTIC_BACKWARD, 0.0

    class InelasticBackward:
        kind, energy = L.MCC_INELASTIC_BACKWARD, 0.0

    class Excitation:
        kind = L.MCC_EXCITATION

        def __init__(self, energy):
            self.energy = float(energy)

    class Ionization:
        kind = L.MCC_IONIZATION

        def __init__(self, energy):
            self.energy = float(energy)


class ChemicalReaction:
    """ChemicalReaction  reactions.jl:7-12"""

    def __init__(self, type_, rate, reactants, stoichiometry):
        self.type, self.rate, self.reactants, self.stoichiometry = type_, rate, reactants, stoichiometry


def reactions(entries, species):
    """Python stand-in for the @reactions macro (reactions.jl:3-101).

    entries: iterable of (rate, "a + b --> c + d", type-or-None); species: name -> object, in the
    order they first appear in the block (the macro's OrderedDict `mapping`, reactions.jl:16)."""
    out = []
    mapping = []
    for entry in entries:
        rate, eqn = entry[0], entry[1]
        type_ = entry[2] if len(entry) > 2 else None
        lhs, rhs = re.split(r"-->|->", eqn)
        reacs, prods = {}, {}
        for side, d in ((lhs, reacs), (rhs, prods)):
            for tok in side.split("+"):
                tok = tok.strip()
                if not tok:
                    continue
                mm = re.match(r"^(\d+)\s*\*?\s*(\w+)$", tok)
                coeff, name = (int(mm.group(1)), mm.group(2)) if mm else (1, tok)
                if name not in mapping:
                    mapping.append(name)
                d[name] = d.get(name, 0) + coeff
        reactants, stoich = [], []
        for name in mapping:                                     # reactions.jl:39-51
            P, R = prods.get(name, 0), reacs.get(name, 0)
            if name in reacs:
                reactants.append((species[name], R))
            if P - R != 0:
                stoich.append((species[name], P - R))
        out.append(ChemicalReaction(type_, rate, reactants, stoich))
    return out


class Collision:
    """MCC.Collision{T}  mcc.jl:9-15"""

    def __init__(self, type_, rate, source, target, products):
        self.type, self.rate, self.source, self.target, self.products = type_, rate, source, target, products


def accept(reaction):
    """accept(reaction)  mcc.jl:291-311"""
    source = target = None
    if len(reaction.reactants) != 2:
        raise AssertionError("Monte Carlo Collisions support only two reacting species: one fluid and one kinetic")
    for r, _ in reaction.reactants:
        if is_fluid(r):
            target = r
        else:
            source = r
    products = [p for p, c in reaction.stoichiometry if c > 0]
    if source is None:
        raise ValueError("Reaction without particle species")
    if target is None:
        raise ValueError("Reaction without fluid species")
    type_ = reaction.type if reaction.type is not None else MCC.ElasticIsotropic()
    return Collision(type_, reaction.rate, source, target, products)


class MonteCarloCollisions:
    """MonteCarloCollisions  mcc.jl:18-51 -- tables and constants are built by the device library."""

    def __init__(self, collisions, seed=0):
        self.collisions = collisions
        self.seed = int(seed)
        self._h = None
        self._rt = None
        self.max_sigma_g = None
        self.m = None
        self.last_nu = None          # nu of the last perform_ (the reference's "nuMCC-<source>-<k>" fields, mcc.jl:287)

    def _bind(self, config):
        if self._h is not None:
            return
        grid = config.grid
        rt = grid._rt
        first = self.collisions[0]
        source, target = first.source, first.target
        source._push(grid)
        N = len(self.collisions)
        kinds = np.array([c.type.kind for c in self.collisions], dtype=np.int32)
        thr = np.array([c.type.energy for c in self.collisions], dtype=np.float64)
        lens = np.array([len(c.rate.nodes) for c in self.collisions], dtype=np.int32)
        eps = np.ascontiguousarray(np.concatenate([c.rate.nodes[:, 0] for c in self.collisions]))
        sig = np.ascontiguousarray(np.concatenate([c.rate.nodes[:, 1] for c in self.collisions]))
        prods = (L.vp * N)()
        for k, c in enumerate(self.collisions):
            prods[k] = None
            if c.type.kind == L.MCC_IONIZATION:
                for p in c.products:                                 # mcc.jl:201-204
                    if p is not source:
                        p._push(grid)
                        prods[k] = p._h
        tn = np.asfortranarray(target.n, dtype=np.float64)
        if tn.shape != tuple(grid.n):
            raise ValueError("target density must live on the grid nodes %s" % (grid.n,))
        h = L.vp()
        L.check(rt.lib.iskb_mcc_create(rt.h, source._h, target.q, target.m, target.T, L.ptr(tn), N, L.ptr(kinds),
                                       L.ptr(thr), L.ptr(lens), L.ptr(eps), L.ptr(sig), prods, self.seed,
                                       C.byref(h)))
        self._h, self._rt = h, rt
        a, b = L.f64(), L.f64()
        L.check(rt.lib.iskb_mcc_constants(h, C.byref(a), C.byref(b)))
        self.max_sigma_g, self.m = a.value, b.value

    def perform_(self, E, dt, config, want_nu=True):
        """PIC.perform!(mcc, E, dt, config)  mcc.jl:231-289 -> (nu, n_candidates, n_collisions)."""
        self._bind(config)
        grid = config.grid
        if E is not None:
            grid._rt.set_fields(E=E)
        source = self.collisions[0].source
        source._push(grid)
        for c in self.collisions:
            for p in c.products:
                if not is_fluid(p):
                    p._push(grid)
        nx, ny = grid.n
        N = len(self.collisions)
        nu = np.zeros((nx, ny, N), order="F") if want_nu else None
        nc, ncoll = L.i64(), L.i64()
        L.check(self._rt.lib.iskb_mcc_perform(self._h, float(dt), L.ptr(nu), C.byref(nc), C.byref(ncoll)))
        source._touched_on_device()
        for c in self.collisions:
            for p in c.products:
                if not is_fluid(p):
                    p._touched_on_device()
        self.last_nu = nu
        return nu, nc.value, ncoll.value

    def totals(self):
        out = np.zeros(2 + len(self.collisions), dtype=np.int64)
        L.check(self._rt.lib.iskb_mcc_totals(self._h, L.ptr(out)))
        return out


def mcc(reaction_list, seed=0):
    """mcc(reactions)  mcc.jl:313-320"""
    return MonteCarloCollisions([accept(r) for r in reaction_list], seed=seed)


# ---- DSMC (Chemistry/src/dsmc.jl) ------------------------------------------------------
class DSMC:
    """module DSMC  dsmc.jl:1-14"""

    class ElasticCollision:
        def __init__(self, rate, source, target):
            self.rate, self.source, self.target = rate, source, target


class DirectSimulationMonteCarlo:
    """DirectSimulationMonteCarlo  dsmc.jl:16-23; perform! runs on the device, one thread per cell."""

    def __init__(self, collisions, seed=0):
        if len(collisions) != 1:
            raise NotImplementedError("one collision per DSMC object: the reference's cell lists accumulate across "
                                      "collisions (dsmc.jl:94-99), more than one is ill defined there")
        self.collisions, self.seed = collisions, int(seed)
        self._h = self._rt = None

    def _bind(self, config):
        if self._h is not None:
            return
        grid, c = config.grid, self.collisions[0]
        c.source._push(grid)
        c.target._push(grid)
        nodes = np.ascontiguousarray(c.rate.nodes)
        gn, sg = np.ascontiguousarray(nodes[:, 0]), np.ascontiguousarray(nodes[:, 1])
        h = L.vp()
        L.check(grid._rt.lib.iskb_dsmc_create(grid._rt.h, c.source._h, c.target._h, L.ptr(gn), L.ptr(sg), len(gn), self.seed,
                                              C.byref(h)))
        self._h, self._rt = h, grid._rt

    def perform_(self, E, dt, config, want_nu=True):
        """PIC.perform!(dsmc, E, dt, config)  dsmc.jl:87-142 -> (nu, n_candidate_pairs, n_collisions)"""
        self._bind(config)
        c = self.collisions[0]
        c.source._push(config.grid)
        c.target._push(config.grid)
        nu = np.zeros(config.grid.n, order="F") if want_nu else None
        nc, ncoll = L.i64(), L.i64()
        L.check(self._rt.lib.iskb_dsmc_perform(self._h, float(dt), L.ptr(nu), C.byref(nc), C.byref(ncoll)))
        c.source._touched_on_device()
        c.target._touched_on_device()
        return nu, nc.value, ncoll.value


def dsmc(reaction_list, seed=0):
    """dsmc(reactions)  dsmc.jl:143-166: two kinetic reactants, no products -> DSMC.ElasticCollision"""
    collisions = []
    for r in reaction_list:
        if len(r.reactants) != 2:
            raise AssertionError("Direct Simulation Monte Carlo support only two reacting, kinetic species")
        (source, _), (target, _) = r.reactants
        if any(cf > 0 for _, cf in r.stoichiometry):
            raise NotImplementedError("DSMC.IonizationCollision has no perform! method in the reference (dsmc.jl:8-13)")
        collisions.append(DSMC.ElasticCollision(r.rate, source, target))
    return DirectSimulationMonteCarlo(collisions, seed=seed)


# ---- secondary-electron emission at a wall (Chemistry/src/see.jl) ------------------------------------
class Vaughan:
    """vaughan(; w0, w0max, gamma0max, ks)  see.jl:17-26 -- the coefficient closure, as its parameters"""

    def __init__(self, w0, w0max, g0max, ks=0.0):
        self.w0, self.w0max, self.g0max, self.ks = float(w0), float(w0max), float(g0max), float(ks)


class Elastic:
    """elastic(gv, we, wemax, gamma_e_max; De = 13, re = 0.03)  see.jl:29-42"""

    def __init__(self, gv, we, wemax, gemax, De=13.0, re=0.03):
        self.gv, self.we, self.wemax, self.gemax, self.De, self.re = gv, float(we), float(wemax), float(gemax), float(De), float(re)


class Inelastic:
    """inelastic(gv; ri = 0.07)  see.jl:45-49"""

    def __init__(self, gv, ri=0.07):
        self.gv, self.ri = gv, float(ri)


class Secondary:
    """secondary(gv; re, ri)  see.jl:52-56"""

    def __init__(self, gv, re, ri):
        self.gv, self.re, self.ri = gv, float(re), float(ri)


# the reference's module-level defaults, see.jl:60-63
gamma_v = Vaughan(13.0, 500.0, 3.0, 1.0)
gamma_e = Elastic(gamma_v, 2.0, 10.0, 0.55, re=0.03)
gamma_i = Inelastic(gamma_v, ri=0.07)
gamma_t = Secondary(gamma_v, re=0.03, ri=0.07)
gamma_0 = None                                    # see.jl:103  (w, theta) -> 0

_SEE_EDGE = {"left": L.EDGE_LEFT, "right": L.EDGE_RIGHT, "bottom": L.EDGE_BOTTOM, "top": L.EDGE_TOP}
_see_calls = [0]


def emit_(primary, secondary, grid, material=None, boundary="all", gamma_t=None, gamma_e=None, gamma_i=None, seed=None):
    """emit!(primary, secondary, grid, material; boundary, gamma_t, gamma_e, gamma_i)  see.jl:114-181 on the device.
    The coefficient arguments are the parameter holders above (arbitrary closures cannot cross the C ABI); all of them must
    share one Vaughan curve, as the reference's defaults do.  Returns {"elastic", "inelastic", "secondaries", "absorbed"}."""
    if boundary not in _SEE_EDGE:
        raise NotImplementedError("emit! with boundary = %r: the reference's wall normal is the zero vector there (see.jl:94) "
                                  "and its secondaries are NaN" % (boundary,))
    if not isinstance(gamma_t, Secondary):
        raise TypeError("gamma_t must be a chemistry.Secondary (the reference's secondary(gv; re, ri) closure)")
    gv = gamma_t.gv
    for g in (gamma_e, gamma_i):
        if g is not None and g.gv is not gv:
            raise NotImplementedError("all emission coefficients must be built on the same Vaughan curve")
    coef = np.array([gv.w0, gv.w0max, gv.g0max, gv.ks,
                     gamma_e.we if gamma_e else 0.0, gamma_e.wemax if gamma_e else 1.0, gamma_e.gemax if gamma_e else -1.0,
                     gamma_e.De if gamma_e else 1.0, gamma_e.re if gamma_e else 0.0,
                     gamma_i.ri if gamma_i else -1.0, gamma_t.re, gamma_t.ri], dtype=np.float64)
    primary._push(grid)
    secondary._push(grid)
    if seed is None:
        _see_calls[0] += 1
        seed = 0x5EE0000 + _see_calls[0]
    out = np.zeros(4, dtype=np.int64)
    L.check(primary._rt.lib.iskb_see_emit(primary._h, secondary._h, _SEE_EDGE[boundary], L.ptr(coef), int(seed), L.ptr(out)))
    primary._touched_on_device()
    secondary._touched_on_device()
    return dict(zip(("elastic", "inelastic", "secondaries", "absorbed"), (int(v) for v in out)))
