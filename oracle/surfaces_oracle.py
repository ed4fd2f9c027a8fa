"""CPU ORACLE (numpy / plain Python) for SURVEY.md section 8(f) row N1 -- TEST INFRASTRUCTURE ONLY.

Surface tracker (absorbing / reflective walls, electrodes), the sigma degrees of freedom of the
Poisson system, and the series RLC circuit of bchaber/iskra, restated line by line from

    ParticleInCell/src/pic/surfaces/build.jl, track.jl, check.jl, hit.jl
    ParticleInCell/src/pic/circuit_coupling.jl
    FiniteDifferenceMethod/src/generalized_poisson.jl:217-269  (add_new_dof, apply_neumann)
    Circuit/src/Circuit.jl:117-136                             (advance_circuit!)
    problem/configuration.jl:22-72                             (create_electrode)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may
import this module; the shipped package `iskra_b200` never does.  Indices are 1-based like the
reference's.  The reference has NO test, fixture or stored output for these files ("parity
unpinned" in the sense of SURVEY.md 8c); the restatement is pinned by analytic known answers
instead (tests/test_oracle_surfaces.py): specular reflection, absorption counts, the uniform field
E_x = sigma between a sigma-driven and a grounded plate, the closed-form RLC recurrence.

Reference quirks that are restated, not repaired (DESIGN.md "N1 hazards"):
  S1  create_electrode builds FloatingPotentialElectrode(phi0, sigma0, 0.0, area) but the struct's
      field order is (sigma, phi, dq, area) (circuit_coupling.jl:11-16, configuration.jl:69-70):
      the electrode's `.sigma` aliases the SOLUTION vector at its reference node and `.phi` aliases
      the sigma right-hand side.  hit! therefore adds dq/area to a solution entry that the next
      `x .= A\\b` overwrites -- collected charge never reaches the field; only `.dq` keeps it.
  S2  track! divides both coordinates by the scalar st.dh = dx (build.jl:88, track.jl:47).
  S3  directed lookups: a face is a key (from-cell, to-cell); default surfaces exist only from
      the inside out (build.jl:27-43), explicit ones in both directions where both cells exist.
  S4  apply_neumann shadows `b` and `c` (generalized_poisson.jl:241-246); rows of nodes that are
      neither on a vertical strip nor at its ends are left untouched.
"""
import math
from collections import deque

import numpy as np

from . import pic_oracle as O

PERIODIC, ABSORBING, REFLECTIVE, ELECTRODE_FIXED, ELECTRODE_FLOATING = 0, 1, 2, 3, 4


class Surface:
    """abstract type Surface + Periodic/Absorbing/Reflective  build.jl:8-11,24-26"""

    def __init__(self, kind):
        self.kind = kind


def create_periodic_surface():
    return Surface(PERIODIC)


def create_absorbing_surface():
    return Surface(ABSORBING)


def create_reflective_surface():
    return Surface(REFLECTIVE)


class SurfaceTracker:
    """SurfaceTracker{2}  build.jl:13-18"""

    def __init__(self, dh):
        self.surface = {}          # ((i,j),(k,l)) -> Surface
        self.tracked = deque()     # (dt, p, (i,j), (hx,hy))
        self.absorbed = set()      # SortedSet(Reverse): iterated from the largest p down
        self.dh = float(dh)
        self._cells = None

    def cells(self):
        """Base.in(x::BoundaryCell, st)  build.jl:86-93: every cell on either side of a key."""
        if self._cells is None:
            s = set()
            for (a, b) in self.surface:
                s.add(a)
                s.add(b)
            self._cells = s
        return self._cells


def build_default_surface_(st, nx, ny, ds):
    """build.jl:33-44 (get!: only where no entry exists yet)"""
    for i in range(1, nx):
        st.surface.setdefault(((i, 1), (i, 0)), ds)
        st.surface.setdefault(((i, ny - 1), (i, ny)), ds)
    for j in range(1, ny):
        st.surface.setdefault(((1, j), (0, j)), ds)
        st.surface.setdefault(((nx - 1, j), (nx, j)), ds)
    st._cells = None


def build_surface_lookup_(st, bcs, ss):
    """build.jl:46-60; bcs boolean (nx, ny) of nodes"""
    bcs = np.asarray(bcs, dtype=bool)
    nx, ny = bcs.shape
    for i in range(1, nx):
        for j in range(1, ny):
            a = bcs[i - 1, j - 1]
            b = bcs[i, j - 1]
            c = bcs[i, j]
            d = bcs[i - 1, j]
            if a and b:
                st.surface[((i, j), (i, j - 1))] = ss
            if b and c:
                st.surface[((i, j), (i + 1, j))] = ss
            if c and d:
                st.surface[((i, j), (i, j + 1))] = ss
            if d and a:
                st.surface[((i, j), (i - 1, j))] = ss
    st._cells = None


def create_surface_tracker(grid, ds=None):
    """create_surface_tracker(grid::CartesianGrid{2}, ds=AbsorbingSurface())  build.jl:95-100 -> :76-87"""
    if ds is None:
        ds = create_absorbing_surface()
    nx, ny = grid.n
    st = SurfaceTracker(grid.dh[0])
    build_default_surface_(st, nx, ny, ds)
    return st


def track_surface_(st, bcs, ss):
    """track_surface!  build.jl:109-111"""
    build_surface_lookup_(st, bcs, ss)


def track_(st, part, dt):
    """track!(st, part, dt)  track.jl:42-52"""
    if st is None:
        return
    st.tracked.clear()
    cells = st.cells()
    for p in range(1, part.np + 1):
        fx = 1.0 + part.x[p - 1, 0] / st.dh        # particle_cell with the scalar st.dh (S2)
        fy = 1.0 + part.x[p - 1, 1] / st.dh
        i, j = int(math.floor(fx)), int(math.floor(fy))
        if (i, j) in cells:
            st.tracked.append((dt, p, (i, j), (fx - i, fy - j)))


def _div(a, b):
    """IEEE division including x/0 (Julia never raises for Float64)"""
    return float(np.float64(a) / np.float64(b)) if b == 0.0 else a / b


def check(pt, pv, dh):
    """check(pt::TrackedParticle{2}, pv, dh)  check.jl:17-36"""
    dt, p, (i, j), (hx, hy) = pt
    vx, vy = float(pv[p - 1, 0]), float(pv[p - 1, 1])
    dx = dh * (1 - hx) if vx > 0 else dh * hx
    dy = dh * (1 - hy) if vy > 0 else dh * hy
    with np.errstate(divide="ignore", invalid="ignore"):
        dtx, dty = _div(dx, abs(vx)), _div(dy, abs(vy))
    if dt < dtx and dt < dty:
        return pt
    if dtx < dty:
        hy2 = hy + vy * dtx / dh
        return (dt - dtx, p, (i + 1, j), (0.0, hy2)) if vx > 0 else (dt - dtx, p, (i - 1, j), (1.0, hy2))
    hx2 = hx + vx * dty / dh
    return (dt - dty, p, (i, j + 1), (hx2, 0.0)) if vy > 0 else (dt - dty, p, (i, j - 1), (hx2, 1.0))


def scattered_(st, pt2):
    """scattered!  hit.jl:12-20"""
    dt, p, (i, j), (hx, hy) = pt2
    i2, j2, hx2, hy2 = i, j, hx, hy
    if hx == 0.0:
        i2, hx2 = i - 1, 1.0
    if hx == 1.0:
        i2, hx2 = i + 1, 0.0
    if hy == 0.0:
        j2, hy2 = j - 1, 1.0
    if hy == 1.0:
        j2, hy2 = j + 1, 0.0
    st.tracked.append((dt, p, (i2, j2), (hx2, hy2)))


class FixedPotentialElectrode(Surface):
    """circuit_coupling.jl:5-9"""

    def __init__(self, phi_ref, area):
        super().__init__(ELECTRODE_FIXED)
        self.phi = phi_ref      # (array, index) view into ps.b
        self.dq = 0.0
        self.area = area


class FloatingPotentialElectrode(Surface):
    """circuit_coupling.jl:11-16, constructed as configuration.jl:69-70 does (S1: swapped views)"""

    def __init__(self, first, second, area):
        super().__init__(ELECTRODE_FLOATING)
        self.sigma = first      # configuration.jl:70 passes phi0 = view(ps.x, phi[i,j]) here
        self.phi = second       # ... and sigma0 = view(ps.b, sigma dof) here
        self.dq = 0.0
        self.area = area


def hit_(surface, part, st, pt, pt2):
    """hit!  hit.jl:26-56, circuit_coupling.jl:40-60"""
    k = surface.kind
    if k == ABSORBING:
        st.absorbed.add(pt2[1])                       # absorbed!(st, pt')
    elif k == REFLECTIVE:
        dt2, p = pt2[0], pt2[1]
        (i, j), (i2, j2) = pt[2], pt2[2]
        nrm = (i2 - i, j2 - j)                         # normalvector hit.jl:1-5
        px, pv = part.x[p - 1], part.v[p - 1]
        px[0:2] -= pv[0:2] * dt2                       # :47
        for a in range(2):                             # :48-52 (n[3] == 0)
            if nrm[a] != 0:
                pv[a] *= -1
        px[0:2] += pv[0:2] * dt2                       # :53
        scattered_(st, pt2)
    elif k == ELECTRODE_FLOATING:
        p = pt[1]
        dq = part.q * part.wg[p - 1]
        surface.dq += dq
        arr, idx = surface.sigma
        arr[idx] += dq / surface.area                  # S1: lands in ps.x, not in the sigma rhs
        st.absorbed.add(p)
    elif k == ELECTRODE_FIXED:
        st.absorbed.add(pt[1])
    # PeriodicSurface: the generic no-op hit!  hit.jl:26-31


def check_(st, part, dt):
    """check!(st, part, dt)  check.jl:39-68.  Returns (too_fast, n_absorbed)."""
    if st is None:
        return False, 0
    pv = part.v[: part.np]
    vmax = st.dh / dt
    too_fast = bool(part.np > 0 and np.any(np.max(np.abs(pv), axis=0) > vmax))
    st.absorbed.clear()
    while st.tracked:
        pt = st.tracked.popleft()
        pt2 = check(pt, part.v, st.dh)
        if pt[2] == pt2[2]:
            continue
        surface = st.surface.get((pt[2], pt2[2]))
        if surface is not None:
            hit_(surface, part, st, pt, pt2)
        else:
            st.tracked.append(pt2)
    n = len(st.absorbed)
    for p in sorted(st.absorbed, reverse=True):        # SortedSet(Reverse)
        O.remove_(part, p)
    return too_fast, n


# ---------------------------------------------------------------------------------------------
# FiniteDifferenceMethod/src/generalized_poisson.jl:217-269
# ---------------------------------------------------------------------------------------------
def add_new_dof(ps):
    """add_new_dof(ps, :sigma)  :217-230 -> 1-based index of the new sigma dof"""
    if not hasattr(ps, "sigma_dof"):
        ps.sigma_dof = []
    N = len(ps.b)
    ps.sigma_dof.append(N)                              # 0-based row N == Julia's N+1
    A = np.zeros((N + 1, N + 1))
    A[:N, :N] = ps.A
    A[N, N] = 1.0
    ps.A = A
    ps.b = np.concatenate([ps.b, [0.0]])
    ps.x = np.concatenate([ps.x, [0.0]])
    return len(ps.sigma_dof)


def apply_neumann(ps, nodes, dof):
    """apply_neumann(ps::PoissonSolver{:xy,2}, nodes, dof)  :235-269 with eps_r == 1"""
    nodes = np.asarray(nodes, dtype=bool)
    dx, dy = ps.dh
    phi = ps.phi_dof
    nx, ny = ps.nx, ps.ny
    A = ps.A
    s = ps.sigma_dof[dof - 1]
    for jj in range(ny):                                # findall: column-major order
        for ii in range(nx):
            if not nodes[ii, jj]:
                continue
            i, j = ii + 1, jj + 1
            a = nodes[i - 2, j - 1] if i > 1 else False
            b = nodes[i, j - 1] if i < nx else False
            c = nodes[i - 1, j - 2] if j > 1 else True
            d = nodes[i - 1, j] if j < ny else True
            r = phi[i - 1, j - 1]
            if c and d and (not a) and (not b):
                i2 = i + 1 if i == 1 else i - 1
                A[r, :] = 0.0
                A[r, r] -= 2 * 1.0 / dx
                A[r, phi[i2 - 1, j - 1]] += 2 * 1.0 / dx
                A[r, s] += 2
            if c != d:
                i2 = i + 1 if i == 1 else i - 1
                j2 = j + 1 if c else j - 1
                A[r, :] = 0.0
                A[r, r] -= 4 * 1.0 / (3 * dx ** 2)
                A[r, r] -= 4 * 1.0 / (3 * dy ** 2)
                A[r, phi[i2 - 1, j - 1]] += 4 * 1.0 / (3 * dx ** 2)
                A[r, phi[i - 1, j2 - 1]] += 4 * 1.0 / (3 * dy ** 2)
                A[r, s] += (2 / 3) * (dx + dy) / (dx * dy)


def electrode_area(nodes, grid):
    """calculate_area  configuration.jl:45-53 (dz = 1)"""
    nodes = np.asarray(nodes, dtype=bool)
    nx, ny = grid.n
    dx, dy = grid.dh
    area = 0.0
    for jj in range(ny):
        for ii in range(nx):
            if nodes[ii, jj]:
                area += dx * 1.0 if (ii + 1 < nx and nodes[ii + 1, jj]) else 0
                area += dy * 1.0 if (jj + 1 < ny and nodes[ii, jj + 1]) else 0
    return area


def reference_node(nodes):
    """find_reference_node  configuration.jl:54-57: first true in column-major order (0-based here)"""
    nodes = np.asarray(nodes, dtype=bool)
    k = int(np.flatnonzero(nodes.reshape(-1, order="F"))[0])
    return k % nodes.shape[0], k // nodes.shape[0]


def create_electrode(nodes, ps, grid, tracker=None, fixed=False, sigma=0.0, phi=0.0):
    """create_electrode  configuration.jl:22-72 (both methods; tracker may be None like the 3-arg form)"""
    area = electrode_area(nodes, grid)
    i, j = reference_node(nodes)
    if fixed:
        O.apply_dirichlet(ps, nodes, phi)
        ps.b[ps.phi_dof[i, j]] = phi
        el = FixedPotentialElectrode((ps.b, ps.phi_dof[i, j]), area)
    else:
        dof = add_new_dof(ps)
        apply_neumann(ps, nodes, dof)
        ps.b[ps.sigma_dof[dof - 1]] = sigma
        el = FloatingPotentialElectrode((ps.x, ps.phi_dof[i, j]), (ps.b, ps.sigma_dof[dof - 1]), area)
    if tracker is not None:
        track_surface_(tracker, nodes, el)
    return el


def calculate_electric_potential(ps, f):
    """generalized_poisson.jl:372-378 for a system that may carry sigma dofs (x .= keeps the aliases)"""
    ff = np.asarray(f).reshape(-1, order="F")
    rd = np.asarray(ps.rho_dof, dtype=np.int64)
    ps.b[rd] = ff[rd] / ps.eps0
    ps.x[:] = np.linalg.solve(ps.A, ps.b)
    return ps.x[ps.phi_dof]


# ---------------------------------------------------------------------------------------------
# Circuit/src/Circuit.jl and circuit_coupling.jl:18-39
# ---------------------------------------------------------------------------------------------
class ShortedConnection:
    """Circuit.jl:27, :53"""

    def voltage(self):
        return 0.0


class PlasmaDevice:
    """circuit_coupling.jl:18-25"""

    def __init__(self, positive, negative):
        self.positive, self.negative = positive, negative

    def voltage(self):
        a, i = self.positive.phi
        b, k = self.negative.phi
        return float(a[i] - b[k])


class CircuitRLC:
    """Circuit.jl:14-23, :29-30, rlc :59-68"""

    def __init__(self, R=0.0, L=0.0, C=0.0, V=None, ext=None, i0=0.0, q0=0.0, t0=0.0):
        self.R, self.L, self.C = float(R), float(L), float(C)
        self.i, self.q, self.t = float(i0), float(q0), float(t0)
        self.V = V if V is not None else (lambda t: 0.0)
        self.ext = ext if ext is not None else ShortedConnection()


def advance_circuit_(cir, dt):
    """advance_circuit!(cir, V, dt)  Circuit.jl:117-136"""
    t, v = cir.t, cir.V
    i, q = cir.i, cir.q
    R, L, C = cir.R, cir.L, cir.C
    vext = cir.ext.voltage()
    cir.i = (L / dt - R / 2) * i + vext - v(t)
    if C > 0.0:
        cir.i -= q / C
        cir.i /= (L / dt + R / 2)
        cir.q = q + dt * i
    else:                                   # `else L > 0.0 || R > 0.0` is an else-branch with a dead expression
        cir.i /= (L / dt + R / 2)
    cir.t += dt
    return vext


def advance_circuit_coupling_(circuit, ps, dt):
    """advance!(circuit::CircuitRLC, phi, dt, config)  circuit_coupling.jl:34-48.  Returns d_sigma."""
    if circuit is None:
        return 0.0
    advance_circuit_(circuit, dt)
    if isinstance(circuit.ext, PlasmaDevice):          # foo!  :26-33
        circuit.ext.positive.dq = 0.0
        dsig = -dt * circuit.i / circuit.ext.positive.area
    else:
        dsig = 0.0
    ps.b[ps.sigma_dof[0]] += dsig                      # get_rhs(solver, :sigma, 1) .+= d_sigma
    return dsig


# ---------------------------------------------------------------------------------------------
# ParticleInCell.jl:51-72, :84-139 with a tracker and a circuit
# ---------------------------------------------------------------------------------------------
def advance_(part, E, dt, grid, tracker, after_push):
    """advance!(part, E, B, dt, config)  ParticleInCell.jl:51-61"""
    track_(tracker, part, dt)                          # :56
    partE = O.grid_to_particle(grid, part, E)          # :57
    O.push_in_cartesian_(part, partE, dt)              # :59
    res = check_(tracker, part, dt)                    # :60
    after_push(part, grid)                             # :61
    return res


def step_(species, grid, solver, E, dt, tracker, after_push, circuit=None):
    """loop body ParticleInCell.jl:102-135 without sources/MCC/diagnostics.  Returns (rho, phi, E, absorbed)."""
    absorbed = []
    for part in species:
        absorbed.append(advance_(part, E, dt, grid, tracker, after_push)[1])
    advance_circuit_coupling_(circuit, solver, dt)     # :116
    rho = np.zeros(grid.n)
    for part in species:
        part.n = O.density(part, grid)
        rho += part.n * part.q
    phi = calculate_electric_potential(solver, -rho)
    Enew = O.calculate_electric_field(solver, phi)
    return rho, phi, Enew, absorbed
