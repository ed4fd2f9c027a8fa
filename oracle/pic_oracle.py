"""CPU ORACLE (numpy) -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A line-faithful restatement of the per-timestep particle hot path of bchaber/iskra
(pure Julia; cannot be executed in this image: no `julia` binary).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may
import this module; the shipped package `iskra_b200` never does.

Every function cites the reference file:line (relative to /root/reference) it
follows, and keeps the reference's *operation order* so that floating-point results
match a Julia run bit-for-bit wherever the arithmetic is deterministic (cell indices,
gather, push, wrap/discard, deposit).  Indices are 1-based like the reference.

PARITY PINNING:
  pinned   : restated capacitor known-answer test (FiniteDifferenceMethod/test/runtests.jl:8-22),
             notebook constants m_eV(e-), m_eV(He+), candidate counts 1037/159, Ptmax
             (docs/capacitively_induced_discharge.ipynb:163,254-255, docs/nanbu-scratchbook.ipynb:197).
  UNPINNED : everything that depends on third-party arithmetic that is absent from
             /root/reference -- Interpolations.jl 0.13.1 (sigma(eps) lookup, restated from its
             documented semantics: piecewise linear, Flat() extrapolation), LAPACK behind
             Julia's `A\\b` (here numpy.linalg.solve -> dgesv, same algorithm class),
             Julia's MersenneTwister stream (here numpy PCG64; RNG-dependent outputs are
             compared statistically only).  "parity unpinned" for those.
"""
import math
import numpy as np

# ---------------------------------------------------------------------------------
# constants -- problem/units_and_constants.jl:8-41 (Unitful 1.6.0 => CODATA 2018)
# ---------------------------------------------------------------------------------
kB = 1.380649e-23            # units_and_constants.jl:9  (u"k")
eps0 = 8.8541878128e-12      # units_and_constants.jl:37 (u"ɛ0")
mu0 = 1.25663706212e-6
c0 = math.sqrt((1.0 / eps0) * (1.0 / mu0))   # units_and_constants.jl:38
qe = 1.6021766208e-19        # units_and_constants.jl:39
me = 9.1093837015e-31        # units_and_constants.jl:40
mp = 1.6726218982e-27        # units_and_constants.jl:41
QE_MCC = 1.60217646e-19      # Chemistry/src/mcc.jl:26
KB_MCC = 1.3806503e-23       # Chemistry/src/mcc.jl:75


def thermal_speed(T, m):
    """problem/configuration.jl:79-81  sqrt(2kB*T/m)"""
    return math.sqrt(2 * kB * T / m)


# ---------------------------------------------------------------------------------
# RegularGrids/src/RegularGrids.jl
# ---------------------------------------------------------------------------------
class CartesianGrid2:
    """UniformGrid{:xy,2} -- RegularGrids.jl:7-15; create_uniform_grid :55-69."""

    def __init__(self, xx, yy, left="open", right="open", bottom="open", top="open"):
        xx = np.asarray(xx, dtype=np.float64)
        yy = np.asarray(yy, dtype=np.float64)
        self.n = (len(xx), len(yy))
        dx = xx[1] - xx[0] if len(xx) > 1 else 1.0      # :60
        dy = yy[1] - yy[0] if len(yy) > 1 else 1.0      # :61
        self.dh = (float(dx), float(dy))
        self.bcs = ((left, right), (bottom, top))       # :65
        self.origin = (float(xx[0]), float(yy[0]))      # :59,68


def jl_range(start, step, stop):
    """Julia `start:step:stop` for floats (length via floor((stop-start)/step)+1 with the
    usual range-length fix-up); good enough for the scripted 0:dh:L ranges."""
    n = int(math.floor((stop - start) / step + 1e-9)) + 1
    return start + step * np.arange(n)


def cell_volume(g):
    """RegularGrids.jl:26-38: dx*dy, halved on every non-periodic edge."""
    nx, ny = g.n
    dx, dy = g.dh
    V = np.zeros((nx, ny)) + dx * dy
    (left, right), (bottom, top) = g.bcs
    if left != "periodic":
        V[0, :] *= 0.5
    if right != "periodic":
        V[nx - 1, :] *= 0.5
    if bottom != "periodic":
        V[:, 0] *= 0.5
    if top != "periodic":
        V[:, ny - 1] *= 0.5
    return V


# ---------------------------------------------------------------------------------
# ParticleInCell/src/pic/kinetic.jl
# ---------------------------------------------------------------------------------
class KineticSpecies:
    """kinetic.jl:1-18 (+ create_kinetic_species, problem/configuration.jl:95-102)."""

    def __init__(self, name, N, q=0.0, m=0.0, weight=1.0, D=2, V=3):
        self.name = name
        self.x = np.zeros((N, D))
        self.v = np.zeros((N, V))
        self.n = np.zeros((0, 0))
        self.m = float(m)
        self.q = float(q)
        self.np = 0
        self.w0 = float(weight)                       # configuration.jl:100
        self.wg = np.ones(N) * weight                 # configuration.jl:99
        self.id = np.arange(1, N + 1, dtype=np.uint32)  # kinetic.jl:15


class FluidSpecies:
    """ParticleInCell/src/pic/fluid.jl:1-11."""

    def __init__(self, name, mu, q, m, n, T):
        self.name, self.mu, self.q, self.m, self.T = name, float(mu), float(q), float(m), float(T)
        self.n = np.array(n, dtype=np.float64)


def remove_(sp, i):
    """kinetic.jl:20-27 remove!(sp, i); i is 1-based."""
    np_ = sp.np
    sp.x[i - 1, :] = sp.x[np_ - 1, :]
    sp.v[i - 1, :] = sp.v[np_ - 1, :]
    sp.wg[i - 1], sp.wg[np_ - 1] = sp.wg[np_ - 1], sp.w0
    sp.id[i - 1], sp.id[np_ - 1] = sp.id[np_ - 1], sp.id[i - 1]
    sp.np = np_ - 1


def add_(src, dst):
    """kinetic.jl:29-37 add!(src, dst): x, v appended; wg and id of dst stay."""
    if src.np > 0:
        a, b = dst.np, dst.np + src.np
        dst.x[a:b, :] = src.x[: src.np]
        dst.v[a:b, :] = src.v[: src.np]
        dst.np += src.np


def remove_particles_(part, dh, matches):
    """kinetic.jl:39-50 remove_particles!(part, dh, matches): sequential scan, re-testing the row that was
    swapped in."""
    p = 1
    while p <= part.np:
        fx = 1.0 + part.x[p - 1, 0] / dh[0]
        fy = 1.0 + part.x[p - 1, 1] / dh[1]
        if matches(int(math.floor(fx)), int(math.floor(fy))):
            remove_(part, p)
            continue
        p += 1


# ---------------------------------------------------------------------------------
# ParticleInCell/src/ParticleInCell.jl:28-35
# ---------------------------------------------------------------------------------
def particle_cell(px, dh):
    """Vectorised particle_cell: f = 1 + x/dh (IEEE divide, then add), ij = floor(f), h = f - ij.
    Returns 1-based (i, j) as int64 and fractions hx, hy.  Origin is NOT subtracted."""
    fx = 1.0 + px[:, 0] / dh[0]
    fy = 1.0 + px[:, 1] / dh[1]
    i = np.floor(fx).astype(np.int64)
    j = np.floor(fy).astype(np.int64)
    return i, j, fx - i, fy - j


# ---------------------------------------------------------------------------------
# ParticleInCell/src/pic/cloud_in_cell.jl
# ---------------------------------------------------------------------------------
def grid_to_particle(grid, part, u):
    """cloud_in_cell.jl:20-36.  `u` is an (nx,ny,C) node array (the reference passes a
    closure (i,j)->E[i,j,:]).  Weights are formed first, then multiplied, then summed
    left to right."""
    np_ = part.np
    i, j, hx, hy = particle_cell(part.x[:np_], grid.dh)
    i0, j0 = i - 1, j - 1
    w00 = ((1.0 - hx) * (1.0 - hy))[:, None]
    w10 = ((hx) * (1.0 - hy))[:, None]
    w01 = ((1.0 - hx) * (hy))[:, None]
    w11 = ((hx) * (hy))[:, None]
    pu = w00 * u[i0, j0] + w10 * u[i0 + 1, j0]
    pu = pu + w01 * u[i0, j0 + 1]
    pu = pu + w11 * u[i0 + 1, j0 + 1]
    return pu


def particle_to_grid(part, grid, pu):
    """cloud_in_cell.jl:1-18.  Sequential `+=` over p (np.add.at applies updates in index
    order, so the summation order equals the reference's loop order)."""
    nx, ny = grid.n
    np_ = part.np
    u = np.zeros((nx, ny))
    i, j, hx, hy = particle_cell(part.x[:np_], grid.dh)
    i0, j0 = i - 1, j - 1
    c00 = (1.0 - hx) * (1.0 - hy) * pu
    c10 = (hx) * (1.0 - hy) * pu
    c01 = (1.0 - hx) * (hy) * pu
    c11 = (hx) * (hy) * pu
    # interleave so each particle's 4 updates happen together, in reference order
    ii = np.stack([i0, i0 + 1, i0, i0 + 1], axis=1).ravel()
    jj = np.stack([j0, j0, j0 + 1, j0 + 1], axis=1).ravel()
    cc = np.stack([c00, c10, c01, c11], axis=1).ravel()
    np.add.at(u, (ii, jj), cc)
    return u


def density(species, grid):
    """kinetic.jl:53 (kinetic) / fluid.jl:11 (fluid)."""
    if isinstance(species, FluidSpecies):
        return species.n
    return particle_to_grid(species, grid, species.wg[: species.np]) / cell_volume(grid)


# ---------------------------------------------------------------------------------
# ParticleInCell/src/pic/pushers.jl:37-50  (B == 0: generalized_poisson.jl:412-419)
# ---------------------------------------------------------------------------------
def push_in_cartesian_(part, E, dt):
    """E is (np,3).  Keeps the two different half-kick roundings (H2)."""
    np_ = part.np
    qm = part.q / part.m
    x, v = part.x[:np_], part.v[:np_]
    B = np.zeros_like(E)
    c1 = 0.5 * dt * qm
    vm = c1 * E + v                                 # :41
    t = c1 * B                                      # :42
    t2 = np.sum(t * t, axis=1)[:, None]             # :43
    vp = vm + np.cross(vm, B)                       # :44
    s = 2.0 / (1.0 + t2) * t                        # :45
    vplus = vm + np.cross(vp, s)                    # :46
    v[:, :] = dt * E * qm * 0.5 + vplus             # :48
    D = x.shape[1]
    x[:, :] = dt * v[:, :D] + x                     # :49


# ---------------------------------------------------------------------------------
# ParticleInCell/src/pic/surfaces/wrap.jl  (+ Julia Base fld/mod for Float64)
# ---------------------------------------------------------------------------------
def jl_mod(x, y):
    """Julia Base mod(x::Float64, y::Float64) (float.jl): r = rem(x,y); sign fix-up."""
    r = np.fmod(x, y)
    out = np.where(r == 0, np.copysign(r, y), np.where((r > 0) != (y > 0), r + y, r))
    return out


def jl_fld(x, y):
    """Julia Base fld for floats = div(x,y,RoundDown) = round((x - mod(x,y))/y) (div.jl)."""
    return np.rint((x - jl_mod(x, y)) / y)


def wrap_(part, grid, dims=(1, 2)):
    """wrap.jl:20-33."""
    for d in dims:
        a = d - 1
        L = (grid.n[a] - 1) * grid.dh[a]
        ox = grid.origin[a]
        px = part.x[: part.np, a]
        alpha = jl_fld(px - ox, L)
        m = alpha != 0
        px[m] -= alpha[m] * L


def discard_(part, grid, dims=(1, 2)):
    """wrap.jl:1-18: reverse scan with swap-from-last remove!."""
    before = part.np
    for d in dims:
        a = d - 1
        L = (grid.n[a] - 1) * grid.dh[a]
        ox = grid.origin[a]
        for p in range(part.np, 0, -1):
            alpha = jl_fld(np.float64(part.x[p - 1, a] - ox), L)
            if alpha != 0:
                remove_(part, p)
    return before - part.np


# ---------------------------------------------------------------------------------
# ParticleInCell/src/pic/sources.jl:24-34
# ---------------------------------------------------------------------------------
class MaxwellianSource:
    """sources.jl:8-22 / create_thermalized_beam, problem/configuration.jl:89-93."""

    def __init__(self, rate, wx, wv, dx=None, dv=None):
        self.rate = float(rate)
        self.wx = np.atleast_2d(np.asarray(wx, dtype=np.float64))
        self.wv = np.atleast_2d(np.asarray(wv, dtype=np.float64))
        self.dx = np.zeros_like(self.wx) if dx is None else np.atleast_2d(np.asarray(dx, dtype=np.float64))
        self.dv = np.zeros_like(self.wv) if dv is None else np.atleast_2d(np.asarray(dv, dtype=np.float64))


def create_thermalized_beam(species, x, vb, dx=None, T=300.0, rate=1.0):
    vth = thermal_speed(T, species.m) * np.ones((1, species.v.shape[1]))
    return MaxwellianSource(rate, x, vth, dx=dx, dv=vb)


def sample_(src, species, dt, rng):
    """sources.jl:24-34; RNG stream is numpy's (parity unpinned by construction)."""
    np_ = species.np
    free = species.x.shape[0] - np_
    n = min(free, int(math.floor(src.rate * dt)))
    D, V = species.x.shape[1], species.v.shape[1]
    species.x[np_:np_ + n, :] = rng.random((n, D)) * src.wx + src.dx
    species.v[np_:np_ + n, :] = rng.standard_normal((n, V)) * src.wv + src.dv
    species.np += n


# ---------------------------------------------------------------------------------
# FiniteDifferenceMethod/src/generalized_poisson.jl
# ---------------------------------------------------------------------------------
class PoissonSolver:
    """PoissonSolver{:xy,2} :11-20; create_poisson_solver :27-30 -> :34-68 (dense A)."""

    def __init__(self, grid, eps0_):
        nx, ny = grid.n
        nn = nx * ny
        A = np.zeros((nn, nn))
        phi = np.arange(nn).reshape((nx, ny), order="F")   # :40 (0-based dof ids)
        for j in range(ny):
            for i in range(nx):
                r = phi[i, j]
                if i < nx - 1:
                    A[r, r] -= 1.0
                    A[r, phi[i + 1, j]] += 1.0
                if i > 0:
                    A[r, r] -= 1.0
                    A[r, phi[i - 1, j]] += 1.0
                if j < ny - 1:
                    A[r, r] -= 1.0
                    A[r, phi[i, j + 1]] += 1.0
                if j > 0:
                    A[r, r] -= 1.0
                    A[r, phi[i, j - 1]] += 1.0
        A /= grid.dh[0] ** 2                                # :65
        self.A, self.b, self.x = A, np.zeros(nn), np.zeros(nn)
        self.eps0, self.dh = eps0_, grid.dh
        self.phi_dof = phi
        self.rho_dof = list(range(nn))                      # :41
        self.nx, self.ny = nx, ny


def apply_periodic(ps, axis):
    """:286-324 with eps_r == 1.  NOTE axis 1 couples along j, axis 2 along i."""
    A, phi = ps.A, ps.phi_dof
    dx, dy = ps.dh
    nx, ny = ps.nx, ps.ny
    if axis == 1:
        for j in (0, ny - 1):
            for i in range(nx):
                if j == ny - 1:
                    A[phi[i, j], phi[i, j]] -= (0.5 + 0.5) / dx ** 2
                    A[phi[i, j], phi[i, 0]] += (0.5 + 0.5) / dx ** 2
                if j == 0:
                    A[phi[i, j], phi[i, j]] -= (0.5 + 0.5) / dx ** 2
                    A[phi[i, j], phi[i, ny - 1]] += (0.5 + 0.5) / dx ** 2
    if axis == 2:
        for j in range(ny):
            for i in (0, nx - 1):
                if i == nx - 1:
                    A[phi[i, j], phi[i, j]] -= (0.5 + 0.5) / dy ** 2
                    A[phi[i, j], phi[0, j]] += (0.5 + 0.5) / dy ** 2
                if i == 0:
                    A[phi[i, j], phi[i, j]] -= (0.5 + 0.5) / dy ** 2
                    A[phi[i, j], phi[nx - 1, j]] += (0.5 + 0.5) / dy ** 2


def apply_dirichlet(ps, nodes, phi0):
    """:205-215; `nodes` boolean (nx,ny); iteration order = CartesianIndices (column-major)."""
    phi = ps.phi_dof
    nodes = np.asarray(nodes, dtype=bool)
    for j in range(ps.ny):
        for i in range(ps.nx):
            if nodes[i, j]:
                r = phi[i, j]
                ps.A[r, :] = 0.0
                ps.A[r, r] = 1.0
                ps.b[r] = phi0
                if r in ps.rho_dof:
                    ps.rho_dof.remove(r)


def calculate_electric_potential(ps, f):
    """:372-378; solve(A,b)=A\\b :201-203 -> dense LU with partial pivoting (dgesv)."""
    ff = np.asarray(f).reshape(-1, order="F")
    rd = np.asarray(ps.rho_dof, dtype=np.int64)
    ps.b[rd] = ff[rd] / ps.eps0
    ps.x[:] = np.linalg.solve(ps.A, ps.b)
    return ps.x[ps.phi_dof]


def calculate_electric_field(ps, phi):
    """:398-410: central inside, one-sided at all four edges; Ez = 0."""
    nx, ny = phi.shape
    dx, dy = ps.dh
    E = np.zeros((nx, ny, 3))
    E[1:nx - 1, :, 0] = (phi[0:nx - 2, :] - phi[2:nx, :]) / (2 * dx)
    E[:, 1:ny - 1, 1] = (phi[:, 0:ny - 2] - phi[:, 2:ny]) / (2 * dy)
    E[0, :, 0] = (phi[0, :] - phi[1, :]) / dx
    E[nx - 1, :, 0] = (phi[nx - 2, :] - phi[nx - 1, :]) / dx
    E[:, 0, 1] = (phi[:, 0] - phi[:, 1]) / dy
    E[:, ny - 1, 1] = (phi[:, ny - 2] - phi[:, ny - 1]) / dy
    return E


# ---------------------------------------------------------------------------------
# Chemistry/src/cross_section.jl:3-14  (Interpolations.jl 0.13.1 LinearInterpolation, Flat())
# ---------------------------------------------------------------------------------
class CrossSection:
    def __init__(self, nodes):
        self.nodes = np.asarray(nodes, dtype=np.float64)

    def __call__(self, x):
        """Piecewise linear on the knots, clamped outside.  Interpolations' Gridded(Linear())
        evaluates (1-f)*y[k] + f*y[k+1] with f=(x-x[k])/(x[k+1]-x[k])."""
        xs, ys = self.nodes[:, 0], self.nodes[:, 1]
        x = np.asarray(x, dtype=np.float64)
        xc = np.clip(x, xs[0], xs[-1])
        k = np.clip(np.searchsorted(xs, xc, side="right") - 1, 0, len(xs) - 2)
        f = (xc - xs[k]) / (xs[k + 1] - xs[k])
        return (1.0 - f) * ys[k] + f * ys[k + 1]


# ---------------------------------------------------------------------------------
# Chemistry/src/mcc.jl
# ---------------------------------------------------------------------------------
ELASTIC_ISOTROPIC, ELASTIC_BACKWARD, INELASTIC_BACKWARD, EXCITATION, IONIZATION = 0, 1, 2, 3, 4


class Collision:
    """MCC.Collision{T} mcc.jl:9-15."""

    def __init__(self, kind, rate, source, target, products=(), energy=0.0):
        self.kind, self.rate, self.source, self.target = kind, rate, source, target
        self.products, self.energy = list(products), float(energy)


def mcc_mass(species):
    """mcc.jl:26"""
    return species.m / QE_MCC


class MonteCarloCollisions:
    """mcc.jl:18-51: union energy grid, max over it of sum_i sigma_i(eps)*sqrt(2 eps/m)."""

    def __init__(self, collisions):
        eps = [0.0]
        for c in collisions:
            eps.extend(c.rate.nodes[:, 0].tolist())
        eps = np.array(sorted(set(eps)))                # :34 sort(unique(eps))
        ms = mcc_mass(collisions[0].source)
        alpha = math.sqrt(2.0 / ms)
        v = alpha * np.sqrt(eps)
        sg = np.zeros_like(v)
        for c in collisions:
            sg += c.rate(eps) * v                       # :45
        self.collisions = collisions
        self.max_sigma_g = float(np.max(sg))
        self.v_at_max = float(v[np.argmax(sg)])
        self.m = ms
        self.remainder = 0.0


def thermal_speed_mcc(T, m):
    """mcc.jl:74-77"""
    return math.sqrt(2 * KB_MCC * T / m)


def _unrotated(st, ct, sp, cp):
    """mcc.jl:83-87"""
    return np.array([[cp * ct, -sp * ct, -st],
                     [-sp, cp, 0.0],
                     [cp * st, sp * st, ct]])


def _euler_angles(v):
    """mcc.jl:89-106 (isapprox(sin,0) with default rtol => exact zero test)."""
    nv = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    ct = v[2] / nv
    st = math.sqrt(1.0 - ct ** 2)
    if st == 0.0:
        cp, sp = 1.0, 0.0
    else:
        cp = v[0] / nv / st
        sp = v[1] / nv / st
    return st, ct, sp, cp


def _isotropic_distribution(rng):
    """mcc.jl:53-62 (polar angle uniform in [0,2pi) -- reference quirk H8)."""
    chi = 2 * math.pi * rng.random()
    eta = 2 * math.pi * rng.random()
    return math.sin(chi), math.cos(chi), math.sin(eta), math.cos(eta)


def _cosine_distribution(rng):
    """mcc.jl:64-72"""
    sc = math.sqrt(rng.random())
    cc = -math.sqrt(1.0 - sc ** 2)
    eta = 2 * math.pi * rng.random()
    return sc, cc, math.sin(eta), math.cos(eta)


def _scatter(v, angles):
    st, ct, sp, cp = _euler_angles(v)
    sc, cc, se, ce = angles
    T = _unrotated(st, ct, sp, cp)
    return np.array([sc * ce, sc * se, cc]) @ T


def isotropic_scattering(v, rng):
    """mcc.jl:108-113"""
    return _scatter(v, _isotropic_distribution(rng))


def diffuse_reflection(v, rng):
    """mcc.jl:122-127"""
    return _scatter(v, _cosine_distribution(rng))


def _norm(v):
    return math.sqrt(float(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))


def perform_collision_(c, p, rng):
    """mcc.jl:129-229; p is 1-based.  Returns None."""
    source, target = c.source, c.target
    sv = source.v[p - 1]
    if c.kind in (ELASTIC_BACKWARD, INELASTIC_BACKWARD, ELASTIC_ISOTROPIC):
        mr1 = source.m / (source.m + target.m)
        mr2 = target.m / (source.m + target.m)
        tv = rng.standard_normal(3) * thermal_speed_mcc(target.T, target.m)
        vr = sv - tv
        w = mr1 * sv + mr2 * tv
        if c.kind == ELASTIC_BACKWARD:
            vr_cp = _norm(vr) * diffuse_reflection(vr, rng)            # :138
        elif c.kind == INELASTIC_BACKWARD:
            vr_cp = rng.random() * _norm(vr) * diffuse_reflection(vr, rng)  # :153
        else:
            vr_cp = _norm(vr) * isotropic_scattering(vr, rng)          # :169
        sv[:] = w + mr2 * vr_cp
        return
    ms = mcc_mass(source)
    sE = 0.5 * ms * float(sv @ sv) - c.energy
    if sE < 0:
        return                                                         # :179-182 / :220-223
    if c.kind == EXCITATION:
        ev = math.sqrt(2 / ms) * math.sqrt(sE)
        sv[:] = ev * isotropic_scattering(sv, rng)                     # :228
        return
    # IONIZATION :184-212
    e1E = sE * rng.random()
    e2E = sE - e1E
    alpha = math.sqrt(2.0 / ms)
    e1v, e2v = alpha * math.sqrt(e1E), alpha * math.sqrt(e2E)
    sv[:] = e1v * diffuse_reflection(sv, rng)
    source.x[source.np, :] = source.x[p - 1, :]
    source.v[source.np, :] = e2v * diffuse_reflection(sv, rng)
    source.np += 1
    tv = rng.standard_normal(3) * thermal_speed_mcc(target.T, target.m)
    for product in c.products:
        if product is source:
            continue
        mpn = source.w0 / product.w0
        for _ in range(int(np.rint(mpn))):
            product.x[product.np, :] = source.x[p - 1, :]
            product.v[product.np, :] = tv
            product.np += 1


def mcc_perform_(mcc, E, dt, grid, rng):
    """PIC.perform!(mcc, E, dt, config)  mcc.jl:231-289.  Returns (nu, Nc, n_collisions)."""
    nx, ny = grid.n
    N = len(mcc.collisions)
    nu = np.zeros((nx, ny, N))
    first = mcc.collisions[0]
    source, target = first.source, first.target
    dens = density(target, grid)
    max_n0 = float(np.max(dens))
    max_Pt = 1.0 - math.exp(-max_n0 * mcc.max_sigma_g * dt)
    if max_Pt > 1.0 / N:
        raise AssertionError("Maximum probability (%g) is greater than 1/%d" % (max_Pt, N))
    frac, Nc = math.modf(N * max_Pt * source.np + mcc.remainder)      # :248
    mcc.remainder = frac
    ncoll = 0
    tvE = (target.q / target.m) * E * dt                              # :266 (hoisted; same value)
    for _ in range(int(Nc)):
        p = int(rng.integers(1, source.np + 1))                       # :251
        i, j, _, _ = particle_cell(source.x[p - 1:p], grid.dh)
        i, j = int(i[0]), int(j[0])
        n = dens[i - 1, j - 1]
        if n < 0:
            continue
        U = rng.random()
        k = int(math.floor(N * U + 1))
        c = mcc.collisions[k - 1]
        d = tvE[i - 1, j - 1, :] - source.v[p - 1, :]
        g = _norm(d)
        eps = 0.5 * mcc.m * g ** 2
        skg = float(c.rate(eps)) * g
        Pk = 1.0 - math.exp(-n * skg * dt)
        Pk /= N * max_Pt
        if Pk > 1.0:
            raise AssertionError("Energy outside of the range")
        if U > k / N - Pk:
            perform_collision_(c, p, rng)
            nu[i - 1, j - 1, k - 1] += 1
            ncoll += 1
    return nu, int(Nc), ncoll


# ---------------------------------------------------------------------------------
# ParticleInCell/src/ParticleInCell.jl:51-72 advance!, :84-139 solve
# ---------------------------------------------------------------------------------
def advance_(part, E, dt, grid, after_push):
    partE = grid_to_particle(grid, part, E)         # :57  (B gather :58 is identically zero)
    push_in_cartesian_(part, partE, dt)             # :59
    after_push(part, grid)                          # :61


def step_(species, interactions, grid, solver, E, dt, after_push, rng=None):
    """One iteration of the loop body :102-135 (no sources, no circuit, no diagnostics).
    Returns (rho, phi, E_new)."""
    for mcc in interactions:                        # :109-111
        mcc_perform_(mcc, E, dt, grid, rng)
    for part in species:                            # :113-115
        if isinstance(part, KineticSpecies):
            advance_(part, E, dt, grid, after_push)
    rho = np.zeros(grid.n)                          # :118
    for part in species:                            # :119-124
        part.n = density(part, grid)
        rho += part.n * part.q
    phi = calculate_electric_potential(solver, -rho)   # :126
    Enew = calculate_electric_field(solver, phi)       # :127
    return rho, phi, Enew
