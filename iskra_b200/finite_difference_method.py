"""Host mirror of FiniteDifferenceMethod/src/generalized_poisson.jl for the path."""
import ctypes as C

import numpy as np

from . import _lib as L


class PoissonSolver:
    """PoissonSolver{:xy,2}  generalized_poisson.jl:11-20 -- the operator lives on the device in
    separable form (or as a dense inverse for irregular Dirichlet masks)."""

    def __init__(self, grid, eps0):
        self.grid, self.eps0 = grid, float(eps0)
        self._rt = grid._rt
        self.dh = grid.dh
        L.check(self._rt.lib.iskb_poisson_create(self._rt.h, self.eps0))

    @property
    def mode(self):
        m = L.i32()
        L.check(self._rt.lib.iskb_poisson_mode(self._rt.h, C.byref(m)))
        return {1: "separable", 2: "dense"}[m.value]

    def dense(self):
        """(A, b) exactly as the reference assembles them (debug / parity); sigma dofs come last."""
        v = L.i64()
        L.check(self._rt.lib.iskb_poisson_dense_size(self._rt.h, C.byref(v)))
        nn = v.value
        A = np.zeros((nn, nn), order="F")
        b = np.zeros(nn)
        L.check(self._rt.lib.iskb_poisson_get_dense(self._rt.h, L.ptr(A), L.ptr(b)))
        return A, b


def create_poisson_solver(grid, eps0):
    """create_poisson_solver(grid::CartesianGrid{2}, eps0)  :27-30 / (grid::AxialGrid{2}, eps0)  :70-199"""
    ps = PoissonSolver(grid, eps0)
    from .regular_grids import AxialGrid
    if isinstance(grid, AxialGrid):
        A = np.asfortranarray(assemble_axial_operator(grid))
        L.check(ps._rt.lib.iskb_poisson_set_dense(ps._rt.h, L.ptr(A), A.shape[0]))
    return ps


def assemble_axial_operator(grid):
    """The matrix of create_poisson_solver(grid::AxialGrid{2}, eps0)  :70-199, row by row in the reference's
    accumulation order (every entry is a short sum of +-1/dr^2, +-1/dz^2, +-0.5/dr/r terms; the order in which the
    diagonal collects them differs between the row families and is kept).  In a Julia deployment the reference's
    own assembler produces this array; it is handed to the device as is (iskb_poisson_set_dense)."""
    nr, nz = grid.n
    dr, dz = grid.dh
    nn = nr * nz
    A = np.zeros((nn, nn), order="F")
    ar, az = 1.0 / dr ** 2, 1.0 / dz ** 2
    idx = lambda i, j: (i - 1) + (j - 1) * nr            # 1-based (i, j) -> 0-based dof
    for j in range(1, nz + 1):
        for i in range(1, nr + 1):
            r_ = idx(i, j)
            d = 0.0
            # radial part
            if i == nr:
                d = (d + ar) - 2.0 * ar                   # :151-152, :181-182, :191-192
                A[r_, idx(i - 1, j)] += ar
            else:
                A[r_, idx(i + 1, j)] += ar
                d = d - 2.0 * ar
                if i == 1:
                    d = d + ar                            # :137-139, :161-163, :171-173
                else:
                    A[r_, idx(i - 1, j)] += ar
            # axial part
            if j == nz:
                d = (d + az) - 2.0 * az                   # :123-124, :175-176, :195-196
                A[r_, idx(i, j - 1)] += az
            else:
                A[r_, idx(i, j + 1)] += az
                d = d - 2.0 * az
                if j == 1:
                    d = d + az                            # :106-108, :165-167, :185-187
                else:
                    A[r_, idx(i, j - 1)] += az
            A[r_, r_] = d
            if 1 < i < nr:                                # 1/r dphi/dr :94-95, :110-111, :127-128
                r = (i - 1) * dr
                A[r_, idx(i + 1, j)] += 0.5 / dr / r
                A[r_, idx(i - 1, j)] -= 0.5 / dr / r
    return A


def apply_periodic(ps, axis):
    """apply_periodic(ps, axis)  :286-324"""
    L.check(ps._rt.lib.iskb_poisson_apply_periodic(ps._rt.h, int(axis)))


def apply_dirichlet(ps, nodes, phi0):
    """apply_dirichlet(ps, nodes::BitArray, phi0)  :205-215.  Whole-edge masks take the cheap path."""
    nodes = np.asarray(nodes, dtype=bool)
    nx, ny = ps.grid.n
    for edge, sel in ((L.EDGE_LEFT, (0, slice(None))), (L.EDGE_RIGHT, (nx - 1, slice(None))),
                      (L.EDGE_BOTTOM, (slice(None), 0)), (L.EDGE_TOP, (slice(None), ny - 1))):
        m = np.zeros((nx, ny), dtype=bool)
        m[sel] = True
        if np.array_equal(m, nodes):
            L.check(ps._rt.lib.iskb_poisson_apply_dirichlet_edge(ps._rt.h, edge, float(phi0)))
            return
    mask = np.asfortranarray(nodes.astype(np.uint8))
    L.check(ps._rt.lib.iskb_poisson_apply_dirichlet(ps._rt.h, L.ptr(mask), float(phi0)))


def calculate_electric_potential(ps, f):
    """phi = calculate_electric_potential(ps, f) with f = -rho  :372-378 (ParticleInCell.jl:126)."""
    rho = -np.asfortranarray(f, dtype=np.float64)
    ps._rt.set_fields(rho=rho)
    L.check(ps._rt.lib.iskb_field_solve(ps._rt.h))
    return ps._rt.fields(rho=False, E=False)[1]


def calculate_electric_field(ps, phi=None):
    """E = calculate_electric_field(ps, phi)  :398-410 -- E of the last solve (phi is on the device)."""
    return ps._rt.fields(rho=False, phi=False)[2]


def calculate_magnetic_field(ps):
    """calculate_magnetic_field  :412-419: identically zero."""
    nx, ny = ps.grid.n
    return np.zeros((nx, ny, 3), order="F")


def add_new_dof(ps, symbol="sigma"):
    """add_new_dof(ps, :sigma)  :217-230 -> 1-based index of the new dof"""
    if symbol not in ("sigma", "σ"):
        raise NotImplementedError("only sigma dofs exist in the reference")
    d = L.i32()
    L.check(ps._rt.lib.iskb_poisson_add_dof(ps._rt.h, C.byref(d)))
    return d.value


def apply_neumann(ps, nodes, dof):
    """apply_neumann(ps, nodes, dof)  :235-269"""
    mask = np.asfortranarray(np.asarray(nodes, dtype=bool).astype(np.uint8))
    L.check(ps._rt.lib.iskb_poisson_apply_neumann(ps._rt.h, L.ptr(mask), int(dof)))


class SigmaRhs:
    """get_rhs(ps, :sigma, dof)  :367-370 -- a handle on the device value: `.value`, `+=` via add()."""

    def __init__(self, ps, dof):
        self.ps, self.dof = ps, int(dof)

    @property
    def value(self):
        v = L.f64()
        L.check(self.ps._rt.lib.iskb_poisson_sigma_get(self.ps._rt.h, self.dof, C.byref(v)))
        return v.value

    @value.setter
    def value(self, x):
        L.check(self.ps._rt.lib.iskb_poisson_sigma_set(self.ps._rt.h, self.dof, float(x)))

    def add(self, dx):
        L.check(self.ps._rt.lib.iskb_poisson_sigma_add(self.ps._rt.h, self.dof, float(dx)))


class DirichletRhs:
    """get_rhs(ps, :phi, i, j) of a Dirichlet node: the value last given to apply_dirichlet (host side)."""

    def __init__(self, value):
        self.value = float(value)


class PhiSolution:
    """get_solution(ps, :phi, i, j)  :363-364 ; 1-based node"""

    def __init__(self, ps, i, j):
        self.ps, self.i, self.j = ps, int(i), int(j)

    @property
    def value(self):
        v = L.f64()
        L.check(self.ps._rt.lib.iskb_phi_at(self.ps._rt.h, self.i, self.j, C.byref(v)))
        return v.value


def get_rhs(ps, symbol, *idx):
    if symbol in ("sigma", "σ"):
        return SigmaRhs(ps, idx[0])
    raise NotImplementedError("get_rhs(:phi) is only meaningful for Dirichlet nodes; see create_electrode")


def get_solution(ps, symbol, i, j):
    if symbol in ("phi", "ϕ", "φ"):
        return PhiSolution(ps, i, j)
    raise NotImplementedError(symbol)
