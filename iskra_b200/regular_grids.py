"""Host mirror of RegularGrids/src/RegularGrids.jl for the path (CartesianGrid{2} only)."""
import ctypes as C

import numpy as np

from . import _lib as L
from .runtime import Runtime

_BC = {"open": L.BC_OPEN, "periodic": L.BC_PERIODIC, "other": L.BC_OPEN}


class UniformGrid:
    """UniformGrid{:xy,2}  RegularGrids.jl:7-15.  Creating it allocates rho/phi/E on the device."""

    def __init__(self, xx, yy, left="open", right="open", bottom="open", top="open", device=None):
        xx = np.asarray(xx, dtype=np.float64)
        yy = np.asarray(yy, dtype=np.float64)
        self.n = (len(xx), len(yy))
        dx = float(xx[1] - xx[0]) if len(xx) > 1 else 1.0      # :60
        dy = float(yy[1] - yy[0]) if len(yy) > 1 else 1.0      # :61
        self.dh = (dx, dy)
        self.bcs = ((left, right), (bottom, top))
        self.origin = (float(xx[0]), float(yy[0]))
        self.coords = (np.repeat(xx[:, None], len(yy), 1), np.repeat(yy[None, :], len(xx), 0))
        self.data = {}
        self._rt = Runtime(device)
        self._rt.grid = self
        bcs = (C.c_int32 * 4)(_BC[left], _BC[right], _BC[bottom], _BC[top])
        L.check(self._rt.lib.iskb_grid_set(self._rt.h, self.n[0], self.n[1], dx, dy, self.origin[0], self.origin[1], bcs))

    def size(self):
        return self.n

    def __len__(self):
        return self.n[0] * self.n[1]


CartesianGrid = UniformGrid


class StaggeredGrid:
    """create_staggered_grid(g)  RegularGrids.jl:99-108: the cell-centred companion of g (nx+1 by ny+1 nodes, shifted by half a
    cell).  The path never computes on it (config.cells is carried through solve untouched, ParticleInCell.jl:91), so it
    is a host object: no device context."""

    def __init__(self, g):
        (x0, y0), (dx, dy), (nx, ny) = g.origin, g.dh, g.n
        xs = np.linspace(x0 - dx / 2.0, x0 + (nx - 1) * dx + dx / 2.0, nx + 1)
        ys = np.linspace(y0 - dy / 2.0, y0 + (ny - 1) * dy + dy / 2.0, ny + 1)
        self.n, self.dh, self.origin, self.bcs = (nx + 1, ny + 1), (dx, dy), (float(xs[0]), float(ys[0])), g.bcs
        self.coords = (np.repeat(xs[:, None], ny + 1, 1), np.repeat(ys[None, :], nx + 1, 0))

    def size(self):
        return self.n


def create_staggered_grid(g):
    return StaggeredGrid(g)


class AxialGrid(UniformGrid):
    """AxialGrid{2} = UniformGrid{:rz,2}  RegularGrids.jl:18; create_axial_grid :84-97.  Coordinates (r, z); the node
    volumes are rings (cell_volume :40-53), uploaded once."""

    def __init__(self, rr, zz, bottom="open", top="open", device=None):
        super().__init__(rr, zz, "other", "other", bottom, top, device)
        L.check(self._rt.lib.iskb_cell_volume_set(self._rt.h, L.ptr(np.asfortranarray(axial_cell_volume(self)))))


def create_axial_grid(rr, zz, bottom="open", top="open", device=None):
    """create_axial_grid(rr, zz; bottom, top)  RegularGrids.jl:84-97"""
    return AxialGrid(rr, zz, bottom, top, device)


def axial_cell_volume(g):
    """cell_volume(g::AxialGrid{2})  RegularGrids.jl:40-53"""
    dr, dz = g.dh
    nr, nz = g.n
    i = np.arange(1, nr + 1, dtype=np.float64)
    ring = np.pi * dz * ((i * dr - 0.5 * dr) ** 2 - (i * dr - 1.5 * dr) ** 2)
    ring[0] = np.pi * dz * (0.5 * dr) ** 2
    ring[nr - 1] = np.pi * dz * ((nr * dr - 1.0 * dr) ** 2 - (nr * dr - 1.5 * dr) ** 2)
    V = np.repeat(ring[:, None], nz, axis=1)
    (_, _), (bottom, top) = g.bcs
    if bottom != "periodic":
        V[:, 0] *= 0.5
    if top != "periodic":
        V[:, nz - 1] *= 0.5
    return V


def create_uniform_grid(xx, yy, left="open", right="open", bottom="open", top="open", device=None):
    """create_uniform_grid(xx, yy; left, right, bottom, top)  RegularGrids.jl:55-69"""
    return UniformGrid(xx, yy, left, right, bottom, top, device)


def cell_volume(g):
    """cell_volume(g::CartesianGrid{2})  RegularGrids.jl:26-38 (computed on the device)."""
    V = np.zeros(g.n, order="F")
    L.check(g._rt.lib.iskb_cell_volume(g._rt.h, L.ptr(V)))
    return V
