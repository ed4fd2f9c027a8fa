"""CPU ORACLE (numpy) for SURVEY.md section 8(f) row N3 -- the axisymmetric r-z variant -- TEST INFRASTRUCTURE ONLY.

Restated line by line from
    RegularGrids/src/RegularGrids.jl:40-53, 84-97        cell_volume(::AxialGrid{2}), create_axial_grid
    ParticleInCell/src/pic/cloud_in_cell.jl:38-73        particle_to_grid / grid_to_particle on an AxialGrid (same as xy)
    ParticleInCell/src/pic/pushers.jl:13-17, 52-66       BorisPusher{:rz}, transform_from_cartesian_to_cylindrical!
    FiniteDifferenceMethod/src/generalized_poisson.jl:70-199   create_poisson_solver(::AxialGrid{2}, eps0)
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.  The reference holds no
test or stored output for these functions ("parity unpinned"); pinned here by the analytic ring volumes and by the
exactness of the interior stencil on phi = a - b r^2.

Not restated: apply_periodic(ps::PoissonSolver{:rz,2}, axis) (:326-361) refers to the undefined names nr, nz, dr,
dz and raises UndefVarError in the reference as written.
"""
import math

import numpy as np

from . import pic_oracle as O


class AxialGrid2(O.CartesianGrid2):
    """AxialGrid{2}  RegularGrids.jl:18, create_axial_grid :84-97: bcs = (:other, :other), (bottom, top)"""

    def __init__(self, rr, zz, bottom="open", top="open"):
        super().__init__(rr, zz, left="other", right="other", bottom=bottom, top=top)
        self.axial = True


def cell_volume(g):
    """RegularGrids.jl:40-53"""
    dr, dz = g.dh
    nr, nz = g.n
    V = np.zeros((nr, nz))
    for i in range(2, nr):
        V[i - 1, :] = math.pi * dz * ((i * dr - 0.5 * dr) ** 2 - (i * dr - 1.5 * dr) ** 2)
    V[0, :] = math.pi * dz * (0.5 * dr) ** 2
    V[nr - 1, :] = math.pi * dz * ((nr * dr - 1.0 * dr) ** 2 - (nr * dr - 1.5 * dr) ** 2)
    (_, _), (bottom, top) = g.bcs
    if bottom != "periodic":
        V[:, 0] *= 0.5
    if top != "periodic":
        V[:, nz - 1] *= 0.5
    return V


def density(species, grid):
    """kinetic.jl:53 with the axial volumes"""
    return O.particle_to_grid(species, grid, species.wg[: species.np]) / cell_volume(grid)


class PoissonSolver(O.PoissonSolver):
    """create_poisson_solver(grid::AxialGrid{2}, eps0)  generalized_poisson.jl:70-199"""

    def __init__(self, grid, eps0_):
        nr, nz = grid.n
        nn = nr * nz
        A = np.zeros((nn, nn))
        phi = np.arange(nn).reshape((nr, nz), order="F")
        dr, dz = grid.dh

        def radial(i, j, lo=True, hi=True):
            r_ = phi[i - 1, j - 1]
            if hi:
                A[r_, phi[i, j - 1]] += 1.0 / dr ** 2
            else:
                A[r_, r_] += 1.0 / dr ** 2
            A[r_, r_] -= 2.0 / dr ** 2
            if lo:
                A[r_, phi[i - 2, j - 1]] += 1.0 / dr ** 2
            else:
                A[r_, r_] += 1.0 / dr ** 2

        def axial(i, j, lo=True, hi=True):
            r_ = phi[i - 1, j - 1]
            if hi:
                A[r_, phi[i - 1, j]] += 1.0 / dz ** 2
            else:
                A[r_, r_] += 1.0 / dz ** 2
            A[r_, r_] -= 2.0 / dz ** 2
            if lo:
                A[r_, phi[i - 1, j - 2]] += 1.0 / dz ** 2
            else:
                A[r_, r_] += 1.0 / dz ** 2

        def curvature(i, j):
            r = (i - 1) * dr
            A[phi[i - 1, j - 1], phi[i, j - 1]] += 0.5 / dr / r
            A[phi[i - 1, j - 1], phi[i - 2, j - 1]] -= 0.5 / dr / r

        # NOTE on ordering inside a row: the reference adds the radial terms, then the axial ones, then the 1/r
        # terms; the axis rows (i = 1) add "+1/dr^2" to the diagonal AFTER "-2/dr^2" (:137-139) while the side rows
        # (i = nr) add it BEFORE (:151-153).  radial()/axial() keep "hi, -2, lo" which reproduces the interior,
        # bottom (:104-106 -- j+1, -2, +1 on the diagonal), axis and their corners; the rows written in the other
        # order are assembled explicitly below so that the diagonal is summed in the reference's order.
        for j in range(2, nz):                       # interior :81-96
            for i in range(2, nr):
                radial(i, j)
                axial(i, j)
                curvature(i, j)
        for i in range(2, nr):                       # bottom :98-113  (j = 1)
            radial(i, 1)
            axial(i, 1, lo=False)
            curvature(i, 1)
        for i in range(2, nr):                       # top :115-130  (j = nz): +1/dz^2, -2/dz^2, phi[i,j-1]
            radial(i, nz)
            r_ = phi[i - 1, nz - 1]
            A[r_, r_] += 1.0 / dz ** 2
            A[r_, r_] -= 2.0 / dz ** 2
            A[r_, phi[i - 1, nz - 2]] += 1.0 / dz ** 2
            curvature(i, nz)
        for j in range(2, nz):                       # axis :132-143  (i = 1): phi[i+1], -2, +1 on the diagonal
            radial(1, j, lo=False)
            axial(1, j)
        for j in range(2, nz):                       # side :145-156  (i = nr): +1, -2 on the diagonal, phi[i-1]
            r_ = phi[nr - 1, j - 1]
            A[r_, r_] += 1.0 / dr ** 2
            A[r_, r_] -= 2.0 / dr ** 2
            A[r_, phi[nr - 2, j - 1]] += 1.0 / dr ** 2
            axial(nr, j)
        # corners :158-197
        radial(1, 1, lo=False)
        axial(1, 1, lo=False)
        radial(1, nz, lo=False)
        r_ = phi[0, nz - 1]
        A[r_, r_] += 1.0 / dz ** 2
        A[r_, r_] -= 2.0 / dz ** 2
        A[r_, phi[0, nz - 2]] += 1.0 / dz ** 2
        for j, top in ((1, False), (nz, True)):
            r_ = phi[nr - 1, j - 1]
            A[r_, r_] += 1.0 / dr ** 2
            A[r_, r_] -= 2.0 / dr ** 2
            A[r_, phi[nr - 2, j - 1]] += 1.0 / dr ** 2
            if not top:
                axial(nr, 1, lo=False)
            else:
                A[r_, r_] += 1.0 / dz ** 2
                A[r_, r_] -= 2.0 / dz ** 2
                A[r_, phi[nr - 1, nz - 2]] += 1.0 / dz ** 2
        self.A, self.b, self.x = A, np.zeros(nn), np.zeros(nn)
        self.eps0, self.dh = eps0_, grid.dh
        self.phi_dof = phi
        self.rho_dof = list(range(nn))
        self.nx, self.ny = nr, nz


def transform_from_cartesian_to_cylindrical_(part, dt):
    """pushers.jl:52-66"""
    np_ = part.np
    x, v = part.x, part.v
    y = dt * v[:np_, 2]
    r = np.sqrt(x[:np_, 0] * x[:np_, 0] + y * y)
    with np.errstate(divide="ignore", invalid="ignore"):
        sin = y / r
    sin[r == 0.0] = 0.0                      # `r .~ 0.0`: isapprox with atol = 0 is an exact test
    cos = np.sqrt(1.0 - sin * sin)
    vr = cos * v[:np_, 0] + sin * v[:np_, 2]
    vy = -sin * v[:np_, 0] + cos * v[:np_, 2]
    x[:np_, 0] = r
    v[:np_, 0] = vr
    v[:np_, 2] = vy


def push_particles_rz_(part, E, dt):
    """push_particles!(::BorisPusher{:rz}, part, E, B, dt)  pushers.jl:13-17"""
    O.push_in_cartesian_(part, E, dt)
    transform_from_cartesian_to_cylindrical_(part, dt)


def advance_(part, E, dt, grid, after_push):
    partE = O.grid_to_particle(grid, part, E)
    push_particles_rz_(part, partE, dt)
    after_push(part, grid)
