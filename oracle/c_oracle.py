"""ctypes loader for the C oracle (oracle/iskra_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  See the header of iskra_oracle.c for scope and parity pinning.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_iskra.so")

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)


def build(force=False):
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "iskra_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_iskra.so"])
    return _SO


class Species(C.Structure):
    _fields_ = [("x", c_dp), ("y", c_dp), ("vx", c_dp), ("vy", c_dp), ("vz", c_dp), ("wg", c_dp),
                ("id", C.POINTER(C.c_uint32)), ("np", C.c_int64), ("cap", C.c_int64),
                ("q", C.c_double), ("m", C.c_double), ("w0", C.c_double)]


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("dx", C.c_double), ("dy", C.c_double),
                ("ox", C.c_double), ("oy", C.c_double), ("bcs", C.c_int32 * 4)]


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4), ("has_spare", C.c_int), ("spare", C.c_double)]


class Collision(C.Structure):
    _fields_ = [("kind", C.c_int32), ("energy", C.c_double), ("eps", c_dp), ("sigma", c_dp),
                ("n_nodes", C.c_int32), ("product", C.POINTER(Species))]


class Mcc(C.Structure):
    _fields_ = [("coll", C.POINTER(Collision)), ("N", C.c_int32), ("source", C.POINTER(Species)),
                ("tq", C.c_double), ("tm", C.c_double), ("tT", C.c_double), ("tn", c_dp),
                ("max_sigma_g", C.c_double), ("m_eV", C.c_double), ("remainder", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # spinning threads hurt on shared vCPUs
        L = C.CDLL(_SO)
        L.orc_num_threads.restype = C.c_int
        L.orc_discard.restype = C.c_int64
        L.orc_sample.restype = C.c_int64
        L.orc_xsec_eval.restype = C.c_double
        L.orc_xsec_eval.argtypes = [c_dp, c_dp, C.c_int32, C.c_double]
        L.orc_dense_solve.restype = C.c_int
        L.orc_electric_potential.restype = C.c_int
        L.orc_mcc_perform.restype = C.c_int
        _lib = L
    return _lib


def dp(a):
    return a.ctypes.data_as(c_dp)


class CSpecies:
    """Owns column-major numpy storage (x :: N x 2, v :: N x 3 as in kinetic.jl:1-12)."""

    def __init__(self, cap, q, m, w0):
        self.cap = int(cap)
        self.xy = np.zeros((2, self.cap))      # row d = Julia column d
        self.v = np.zeros((3, self.cap))
        self.wg = np.ones(self.cap) * w0
        self.id = np.arange(1, self.cap + 1, dtype=np.uint32)
        self.c = Species(dp(self.xy[0]), dp(self.xy[1]), dp(self.v[0]), dp(self.v[1]), dp(self.v[2]),
                         dp(self.wg), self.id.ctypes.data_as(C.POINTER(C.c_uint32)),
                         0, self.cap, q, m, w0)

    @property
    def np(self):
        return int(self.c.np)

    @np.setter
    def np(self, v):
        self.c.np = int(v)

    def set(self, x, y, vx, vy, vz, wg=None, ids=None):
        n = len(x)
        self.xy[0, :n], self.xy[1, :n] = x, y
        self.v[0, :n], self.v[1, :n], self.v[2, :n] = vx, vy, vz
        if wg is not None:
            self.wg[:n] = wg
        if ids is not None:
            self.id[:] = ids
        self.c.np = n

    def ref(self):
        return C.byref(self.c)


def make_grid(nx, ny, dx, dy, ox=0.0, oy=0.0, bcs=(0, 0, 0, 0)):
    return Grid(nx, ny, dx, dy, ox, oy, (C.c_int32 * 4)(*bcs))


def make_rng(seed):
    r = Rng()
    lib().orc_rng_seed(C.byref(r), C.c_uint64(seed))
    return r


class CMcc:
    """Wraps orc_mcc; keeps the numpy tables alive."""

    def __init__(self, source, procs, tq, tm, tT, tn):
        # procs: list of (kind, energy, eps ndarray, sigma ndarray, product CSpecies|None)
        self._keep = []
        arr = (Collision * len(procs))()
        for k, (kind, energy, eps, sig, prod) in enumerate(procs):
            eps = np.ascontiguousarray(eps, dtype=np.float64)
            sig = np.ascontiguousarray(sig, dtype=np.float64)
            self._keep += [eps, sig, prod]
            arr[k] = Collision(kind, energy, dp(eps), dp(sig), len(eps),
                               C.pointer(prod.c) if prod is not None else None)
        self.tn = np.ascontiguousarray(tn, dtype=np.float64)
        self.arr = arr
        self.source = source
        self.c = Mcc(arr, len(procs), C.pointer(source.c), tq, tm, tT, dp(self.tn), 0.0, 0.0, 0.0)
        lib().orc_mcc_setup(C.byref(self.c))

    def perform(self, grid, E, dt, rng, want_nu=True):
        nn = grid.nx * grid.ny
        nu = np.zeros(nn * self.c.N) if want_nu else None
        out = (C.c_int64 * 2)()
        E = np.ascontiguousarray(E)
        rc = lib().orc_mcc_perform(C.byref(self.c), C.byref(grid), dp(E), C.c_double(dt),
                                   dp(nu) if want_nu else None, out, C.byref(rng))
        return rc, nu, int(out[0]), int(out[1])


# ---- surface tracker ---------------------------------------------------------
class Tracker(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("dh", C.c_double), ("face", C.POINTER(C.c_uint8)),
                ("tracked", C.POINTER(C.c_uint8)), ("kind", C.POINTER(C.c_int32)), ("area", c_dp), ("dq", c_dp)]


class TrackedParticle(C.Structure):
    _fields_ = [("dt", C.c_double), ("p", C.c_int64), ("i", C.c_int32), ("j", C.c_int32),
                ("hx", C.c_double), ("hy", C.c_double)]


_DIRS = {(0, -1): 0, (1, 0): 1, (0, 1): 2, (-1, 0): 3}


class CTracker:
    """Flattens the Dict of an oracle.surfaces_oracle.SurfaceTracker into the C table."""

    def __init__(self, st, nx, ny):
        self.nx, self.ny = nx, ny
        self.face = np.zeros(4 * (nx + 1) * (ny + 1), dtype=np.uint8)
        self.tracked = np.zeros((nx + 1) * (ny + 1), dtype=np.uint8)
        self.surfaces = [None]
        ids = {}
        for ((i, j), (k, l)), s in st.surface.items():
            if id(s) not in ids:
                ids[id(s)] = len(self.surfaces)
                self.surfaces.append(s)
            self.face[4 * (i + j * (nx + 1)) + _DIRS[(k - i, l - j)]] = ids[id(s)]
            for (a, b) in ((i, j), (k, l)):
                if 0 <= a <= nx and 0 <= b <= ny:
                    self.tracked[a + b * (nx + 1)] = 1
        self.kind = np.array([-1] + [s.kind for s in self.surfaces[1:]], dtype=np.int32)
        self.area = np.array([0.0] + [getattr(s, "area", 0.0) for s in self.surfaces[1:]])
        self.dq = np.zeros(len(self.surfaces))
        self.c = Tracker(nx, ny, st.dh, self.face.ctypes.data_as(C.POINTER(C.c_uint8)),
                         self.tracked.ctypes.data_as(C.POINTER(C.c_uint8)),
                         self.kind.ctypes.data_as(C.POINTER(C.c_int32)), dp(self.area), dp(self.dq))
        self.queue = None

    def advance(self, sp, grid, E, dt, bmode=(0, 0)):
        """advance! with the tracker -> (n_absorbed, too_fast)"""
        L = lib()
        L.orc_advance_tracked.restype = C.c_int64
        qcap = max(sp.cap, 1)
        if self.queue is None or len(self.queue) < qcap:
            self.queue = (TrackedParticle * qcap)()
        tf = C.c_int32(0)
        E = np.ascontiguousarray(E)
        n = L.orc_advance_tracked(sp.ref(), C.byref(grid), C.byref(self.c), dp(E), C.c_double(dt),
                                  (C.c_int32 * 2)(*bmode), self.queue, C.c_int64(qcap), C.byref(tf))
        return int(n), bool(tf.value)
