"""Host mirror of the Diagnostics module (Diagnostics/src/Diagnostics.jl, hdf5.jl, openpmd/*.jl).

The reference's `@field` / `@particle` registrations copy every record into a host buffer on EVERY step
(`record.input .= data`, openpmd/particles.jl:62-63, over the full capacity) whether or not anything is saved.  Here a
registration stores a *fetch closure*; the device -> host copy happens only when `save_record` asks for the record
(on demand), so a step that saves nothing moves nothing.  Record names, openPMD paths and attributes are the reference's.

File format: openPMD-HDF5, `prefix/hdf5/data<i>.h5` like hdf5.jl:47-66 -- through h5py when that is installed, else through
the pure-Python classic-layout writer in hdf5_min.py (this image has no HDF5 library; see the note there about
validation).  `SINK = "npz"` writes the same paths and attributes into `data<i>.npz` + `.attrs.json` instead.
"""
import datetime
import json
import os

import numpy as np

from . import hdf5_min

SINK = "h5"      # "h5": openPMD-HDF5 (h5py if present, else hdf5_min); "npz": numpy archive + json attributes

# ---- openpmd/root.jl -------------------------------------------------------------------------------
# unitDimension = powers of (length, mass, time, current, temperature, amount, luminosity); unitSI = 1 for SI strings
_UNITS = {
    "1": (0, 0, 0, 0, 0, 0, 0), "kg": (0, 1, 0, 0, 0, 0, 0), "C": (0, 0, 1, 1, 0, 0, 0), "m": (1, 0, 0, 0, 0, 0, 0),
    "kg*m/s": (1, 1, -1, 0, 0, 0, 0), "m/s": (1, 0, -1, 0, 0, 0, 0), "1/m^2": (-2, 0, 0, 0, 0, 0, 0),
    "C/m^2": (-2, 0, 1, 1, 0, 0, 0), "V": (2, 1, -3, -1, 0, 0, 0), "V/m": (1, 1, -3, -1, 0, 0, 0), "A/m": (-1, 0, 0, 1, 0, 0, 0),
    "A": (0, 0, 0, 1, 0, 0, 0),
}


def usi(units):
    """usi(units)  openpmd/root.jl:37-42 for the unit strings the reference registers"""
    return tuple(float(v) for v in _UNITS.get(units, (0,) * 7)), 1.0


ROOT = {"openPMD": "1.1.0", "openPMDextension": 1, "iterationEncoding": "fileBased", "iterationFormat": "data%T.h5",
        "basePath": "/data/%T", "meshesPath": "fields/", "particlesPath": "particles/",
        "author": "Bartosz Chaber <bartosz.chaber@ee.pw.edu.pl>", "seed": 0, "software": "iskra",
        "softwareVersion": "iskra_b200", "date": ""}                                   # RootMetadata  root.jl:49-66
FIELDS = {"fieldSolver": "other", "fieldSolverParameters": "Nagel", "fieldBoundary": ["open"] * 4,
          "particleBoundary": ["periodic"] * 4, "currentSmoothing": "none", "chargeCorrection": "none"}   # fields.jl:1-13
PARTICLES = {"particleShape": 0.0, "currentDeposition": "none", "particlePush": "Boris", "particleInterpolation": "other",
             "particleSmoothing": "none"}                                              # particles.jl:1-8


class FieldRecord:
    """FieldRecord  openpmd/fields.jl:30-66 ; fetch() -> (nx, ny) or (nx, ny, 3) array"""

    def __init__(self, units, fetch, grid, withcomponents=False, offset=0.0, pos=None):
        dim, unit_si = usi(units)
        axial = type(grid).__name__ == "AxialGrid"
        self.fetch, self.withcomponents = fetch, withcomponents
        self.metadata = {"unitDimension": dim, "timeOffset": float(offset), "axisLabels": "rz" if axial else "xy",
                         "dataOrder": "C", "geometry": "cartesian", "geometryParameters": "",
                         "gridGlobalOffset": (0.0, 0.0), "gridSpacing": tuple(grid.dh), "gridUnitSI": 1.0,
                         "fieldSmoothing": "none", "position": tuple(pos) if pos is not None else (0.0, 0.0), "unitSI": unit_si}

    def components(self):
        data = np.asarray(self.fetch())
        if self.withcomponents:
            zz = np.zeros(data.shape[:2])
            k = data.shape[2] if data.ndim == 3 else 0
            return {c: (data[:, :, n] if k > n else zz) for n, c in enumerate("xyz")}
        return {" ": data}


class ParticleRecord:
    """ParticleRecord  openpmd/particles.jl:17-66 ; fetch() -> (np,), (np, k) or a 1-element constant"""

    def __init__(self, units, fetch, species, weighted=False, withcomponents=False, offset=0.0, constant=None):
        dim, unit_si = usi(units)
        self.fetch, self.species, self.withcomponents = fetch, species, withcomponents
        # the reference decides by the length of the REGISTERED buffer (hdf5.jl:74: a one-element input is a constant
        # record, a per-row buffer -- full capacity there -- is a dataset even when one particle is alive); here the
        # registration says it, None = by the size of what fetch returns
        self.constant = constant
        self.metadata = {"unitDimension": dim, "timeOffset": float(offset), "macroWeighted": 1 if weighted else 0,
                         "weightingPower": 1.0, "unitSI": unit_si}

    def components(self):
        data = np.asarray(self.fetch())
        if self.withcomponents:
            zz = np.zeros(data.shape[0])
            return {c: (data[:, n] if data.shape[1] > n else zz) for n, c in enumerate("xyz")}, False
        return {" ": data}, (data.size == 1 if self.constant is None else bool(self.constant))


class ProbeRecord:
    """CircuitProbeRecord  Diagnostics/src/circuit.jl:1-19 ; fetch() -> one number, stored as a 1 x 1 dataset"""

    def __init__(self, units, fetch, offset=0.0):
        dim, unit_si = usi(units)
        self.fetch, self.owner = fetch, None
        self.metadata = {"unitDimension": dim, "timeOffset": float(offset), "axisLabels": ("", ""), "dataOrder": "C",
                         "geometry": "cartesian", "geometryParameters": "", "gridGlobalOffset": (0.0, 0.0),
                         "gridSpacing": (0.0, 0.0), "gridUnitSI": 1.0, "fieldSmoothing": "none", "position": (0.0, 0.0),
                         "unitSI": unit_si}


records = {}          # const records = Dict{String, Record}()  Diagnostics.jl:22
_solve_keys = set()   # keys registered by register_solve_records: dropped when another solve() registers (their closures
                      # hold the previous run's species and device context alive otherwise)


def register_field(key, units, fetch, grid, **kw):
    """@field key units data grid [withcomponents=true]  Diagnostics.jl:33-39 -- lazily: `fetch` is called on save"""
    records[key] = FieldRecord(units, fetch, grid, **kw)


def register_particle(key, units, fetch, species, **kw):
    """@particle key units data part [weighted=true] [withcomponents=true]  Diagnostics.jl:41-47"""
    records[key] = ParticleRecord(units, fetch, species, **kw)


def register_probe(key, fetch, units, **kw):
    """@probe key data units  Diagnostics/src/circuit.jl:25-31"""
    records[key] = ProbeRecord(units, fetch, **kw)


def register_solve_records(config):
    """Everything ParticleInCell.solve / advance! / PIC.perform!(mcc) register (ParticleInCell.jl:63-71, 122-133, mcc.jl:287):
    fields rho, phi, E, B, n<species>, nuMCC-<source>-<k>; per kinetic species id, mass, charge, weighting, momentum, position,
    positionOffset/{x,y,z}.  All of them fetch from the device when saved."""
    from . import particle_in_cell as PIC
    grid = config.grid
    rt = grid._rt
    nx, ny = grid.n
    for k in _solve_keys:
        records.pop(k, None)
    _solve_keys.clear()
    before = set(records)
    register_field("rho", "C/m^2", lambda: rt.fields(phi=False, E=False)[0], grid)
    register_field("phi", "V", lambda: rt.fields(rho=False, E=False)[1], grid)
    register_field("E", "V/m", lambda: rt.fields(rho=False, phi=False)[2], grid, withcomponents=True)
    register_field("B", "A/m", lambda: np.zeros((nx, ny, 3)), grid, withcomponents=True)
    for sp in config.species:
        if PIC.is_fluid(sp):
            register_field("n" + sp.name, "1/m^2", (lambda s=sp: s.n), grid)
            continue

        def dens(s=sp):
            from . import _lib as L
            out = np.zeros(grid.n, order="F")
            L.check(rt.lib.iskb_species_density_download(s._h, L.ptr(out)))
            return out
        register_field("n" + sp.name, "1/m^2", dens, grid)
        n = sp.name
        register_particle(n + "/id", "1", (lambda s=sp: s.id_ro[: s.np].copy()), sp, constant=False)
        register_particle(n + "/mass", "kg", (lambda s=sp: np.array([s.m])), sp, constant=True)
        register_particle(n + "/charge", "C", (lambda s=sp: np.array([s.q])), sp, constant=True)
        register_particle(n + "/weighting", "1", (lambda s=sp: s.wg_ro[: s.np].copy()), sp, weighted=True, constant=False)
        register_particle(n + "/momentum", "kg*m/s", (lambda s=sp: s.m * s.v_ro[: s.np]), sp, withcomponents=True)
        register_particle(n + "/position", "m", (lambda s=sp: s.x_ro[: s.np].copy()), sp, withcomponents=True)
        for ax in "xyz":
            register_particle(n + "/positionOffset/" + ax, "m", (lambda: np.array([0.0])), sp, constant=True)
    for inter in config.interactions:
        if hasattr(inter, "last_nu") and getattr(inter, "collisions", None):
            src = inter.collisions[0].source
            for k in range(len(inter.collisions)):
                register_field("nuMCC-%s-%d" % (src.name, k + 1), "1/m^2",
                               (lambda m=inter, kk=k: m.last_nu[:, :, kk] if m.last_nu is not None else np.zeros((nx, ny))), grid)
    _solve_keys.update(set(records) - before)


# ---- sinks: hdf5.jl ------------------------------------------------------------------------------------
class _NpzSink:
    def __init__(self, path):
        self.path, self.arrays, self.attrs = path, {}, {}

    def write(self, name, array):
        self.arrays[name] = np.asarray(array)

    def set_attrs(self, name, attrs):
        self.attrs.setdefault(name, {}).update({k: (list(v) if isinstance(v, tuple) else v) for k, v in attrs.items()})

    def close(self):
        np.savez(self.path + ".npz", **{k.replace("/", "|"): v for k, v in self.arrays.items()})
        with open(self.path + ".attrs.json", "w") as f:
            json.dump(self.attrs, f, indent=1, default=str)


def _as_julia_wrote_it(array):
    """HDF5.jl stores a column-major (nx, ny) array with the dimensions reversed: the file holds (ny, nx), x fastest
    (hdf5.jl:93 `it[g] = output`).  The same bytes from a numpy (nx, ny) array are its transpose in C order."""
    a = np.asarray(array)
    return np.ascontiguousarray(a.T) if a.ndim >= 2 else a


class _H5Sink:
    def __init__(self, path, h5py):
        self.f = h5py.File(path + ".h5", "w")

    def write(self, name, array):
        self.f[name] = _as_julia_wrote_it(array)

    def set_attrs(self, name, attrs):
        node = self.f.require_group(name) if name not in self.f else self.f[name]
        for k, v in attrs.items():
            node.attrs[k] = v

    def close(self):
        self.f.close()


class _MinH5Sink:
    """openPMD-HDF5 without an HDF5 library (hdf5_min.Writer)"""

    def __init__(self, path):
        self.w = hdf5_min.Writer(path + ".h5")

    def write(self, name, array):
        self.w.write(name, _as_julia_wrote_it(array))

    def set_attrs(self, name, attrs):
        self.w.set_attrs(name, attrs)

    def close(self):
        self.w.close()


class Iteration:
    """the HDF5 group `data/<i>` handed to the callback of new_iteration  hdf5.jl:47-66"""

    def __init__(self, sink, base):
        self.sink, self.base = sink, base


def new_iteration(prefix, i, t, dt, f):
    """new_iteration(f, prefix, i, t, dt)  hdf5.jl:47-66: opens prefix/hdf5/data<i>, calls f(it), closes"""
    os.makedirs(os.path.join(prefix, "hdf5"), exist_ok=True)
    path = os.path.join(prefix, "hdf5", "data%d" % i)
    if SINK == "npz":
        sink = _NpzSink(path)
    else:
        try:
            import h5py
            sink = _H5Sink(path, h5py)
        except ImportError:
            sink = _MinH5Sink(path)
    base = "data/%d/" % i
    root = dict(ROOT)
    root["date"] = datetime.datetime.now().strftime("%Y/%m/%d %H:%M")
    sink.set_attrs("/", root)
    sink.set_attrs(base, {"dt": float(dt), "time": float(t), "timeUnitSI": 1.0})
    sink.set_attrs(base + ROOT["meshesPath"], FIELDS)
    sink.set_attrs(base + ROOT["particlesPath"], PARTICLES)
    it = Iteration(sink, base)
    try:
        f(it)
    finally:
        sink.close()
    return path


def save_record(it, key):
    """save_record(it, key)  Diagnostics.jl:60-66 -> hdf5.jl:69-95"""
    rec = records.get(key)
    if rec is None:
        print("Couldn't find diagnostic ", key)
        return
    if isinstance(rec, ProbeRecord):                                  # hdf5.jl:97-101
        f = it.base + ROOT["meshesPath"] + key
        it.sink.write(f, np.full((1, 1), float(rec.fetch())))
        it.sink.set_attrs(f, rec.metadata)
        return
    if isinstance(rec, FieldRecord):
        f = it.base + ROOT["meshesPath"] + key
        for comp, out in rec.components().items():
            g = f if comp == " " else f + "/" + comp
            it.sink.write(g, out)
            it.sink.set_attrs(g, {k: rec.metadata[k] for k in ("position", "unitSI")})
        it.sink.set_attrs(f, {k: v for k, v in rec.metadata.items() if k not in ("position", "unitSI")})
    else:
        f = it.base + ROOT["particlesPath"] + key
        comps, constant = rec.components()
        for comp, out in comps.items():
            g = f if comp == " " else f + "/" + comp
            if constant:                                            # hdf5.jl:76-79: constant record component
                it.sink.set_attrs(g, {"value": float(np.asarray(out).ravel()[0]), "shape": [int(rec.species.np)]})
            else:
                it.sink.write(g, out)
            it.sink.set_attrs(g, {"unitSI": rec.metadata["unitSI"]})
        it.sink.set_attrs(f, {k: v for k, v in rec.metadata.items() if k != "unitSI"})


def save_records(it, prefix):
    """save_records(it, prefix)  Diagnostics.jl:52-58"""
    for key in list(records):
        if key.startswith(prefix):
            save_record(it, key)


def load_npz(path):
    """Reads back what the npz sink wrote: {openPMD path: array}, {path: attributes}"""
    arrays = {k.replace("|", "/"): v for k, v in np.load(path + ".npz").items()}
    with open(path + ".attrs.json") as f:
        return arrays, json.load(f)


def load(path):
    """Reads back one iteration written by new_iteration (either sink): {openPMD path: array}, {path: attributes};
    paths without leading or trailing slashes, "" is the root; field arrays come back as (nx, ny) whatever the file order."""
    import os as _os
    if _os.path.exists(path + ".h5"):
        try:
            import h5py
        except ImportError:
            arrays, attrs = hdf5_min.read(path + ".h5")
            return ({k.strip("/"): (v.T if v.ndim >= 2 else v) for k, v in arrays.items()},     # back to (nx, ny)
                    {k.strip("/"): {a: (v.tolist() if isinstance(v, np.ndarray) else (v.item() if isinstance(v, np.generic) else v))
                                    for a, v in d.items()} for k, d in attrs.items()})
        arrays, attrs = {}, {}
        with h5py.File(path + ".h5", "r") as f:
            attrs[""] = dict(f.attrs)
            f.visititems(lambda n, o: (attrs.__setitem__(n, dict(o.attrs)),
                                       arrays.__setitem__(n, o[()].T if o.ndim >= 2 else o[()]) if hasattr(o, "shape") else None))
        return arrays, attrs
    arrays, attrs = load_npz(path)
    return {k.strip("/"): v for k, v in arrays.items()}, {k.strip("/"): v for k, v in attrs.items()}
