"""The pure-Python HDF5 writer behind the diagnostics sink (iskra_b200/hdf5_min.py): structure checks against the HDF5
File Format Specification (classic layout) and a round trip through the independent parser in the same module.
No HDF5 library exists in this image; with h5py present the last test opens the file with it as well."""
import struct

import numpy as np
import pytest

from iskra_b200 import hdf5_min as H


def _sample(tmp_path, many=0):
    p = str(tmp_path / "t.h5")
    w = H.Writer(p)
    rng = np.random.default_rng(0)
    data = {"data/3/fields/rho": rng.standard_normal((5, 4)), "data/3/fields/E/x": rng.standard_normal((5, 4)).astype(np.float32),
            "data/3/particles/e-/id": np.arange(1, 8, dtype=np.uint32), "data/3/particles/e-/cell": np.arange(-3, 4, dtype=np.int64),
            "scalar": np.float64(2.5), "empty": np.zeros((0,))}
    for k in range(many):
        data["many/m%04d" % k] = np.full((2,), float(k))
    for k, v in data.items():
        w.write(k, v)
    w.set_attrs("/", {"openPMD": "1.1.0", "openPMDextension": 1, "date": "2026/01/01 10:00"})
    w.set_attrs("data/3", {"dt": 0.1, "time": 0.5})
    w.set_attrs("data/3/fields/rho", {"unitDimension": (-2.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0), "axisLabels": "xy", "gridSpacing": [0.5, 0.25]})
    w.set_attrs("data/3/fields", {"fieldBoundary": ["open"] * 4, "shape": [7]})
    w.set_attrs("data/3/particles/e-/mass", {"value": 9.1e-31, "shape": [7]})       # constant record component: a group with attributes
    w.close()
    return p, data


def test_superblock_and_root_entry_follow_the_classic_layout(tmp_path):
    p, _ = _sample(tmp_path)
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                       # versions 0; 8-byte offsets and lengths
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", b, 16)
    assert (leaf_k, internal_k, flags) == (H.LEAF_K, H.INTERNAL_K, 0)
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and driver == H.UNDEF and eof == len(b)
    name_off, ohdr, cache, _ = struct.unpack_from("<QQII", b, 56)
    assert name_off == 0 and ohdr == 96 and cache == 1
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert b[ohdr] == 1 and ohdr % 8 == 0                                   # version-1 object header, aligned
    # every structure the file refers to is 8-byte aligned
    for sig in (b"TREE", b"HEAP", b"SNOD"):
        at = -1
        while True:
            at = b.find(sig, at + 1)
            if at < 0:
                break
            assert at % 8 == 0


def test_round_trip_of_datasets_groups_and_attributes(tmp_path):
    p, data = _sample(tmp_path)
    arrays, attrs = H.read(p)
    assert set(arrays) == {"/" + k for k in data}
    for k, v in data.items():
        got = arrays["/" + k]
        assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape and np.array_equal(got, v)
    assert attrs["/"]["openPMD"] == "1.1.0" and attrs["/"]["openPMDextension"] == 1 and attrs["/"]["date"] == "2026/01/01 10:00"
    assert attrs["/data/3"] == {"dt": 0.1, "time": 0.5}
    r = attrs["/data/3/fields/rho"]
    assert list(r["unitDimension"]) == [-2.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0] and r["axisLabels"] == "xy" and list(r["gridSpacing"]) == [0.5, 0.25]
    assert attrs["/data/3/fields"]["fieldBoundary"] == ["open"] * 4 and list(attrs["/data/3/fields"]["shape"]) == [7]
    assert "/data/3/particles/e-/mass" not in arrays and attrs["/data/3/particles/e-/mass"]["value"] == 9.1e-31


def test_groups_larger_than_one_symbol_table_node(tmp_path):
    p, data = _sample(tmp_path, many=3 * H.SNOD_CAP + 5)
    arrays, _ = H.read(p)                       # the parser checks name order and the B-tree keys of every node on the way
    assert len([k for k in arrays if k.startswith("/many/")]) == 3 * H.SNOD_CAP + 5
    assert np.array_equal(arrays["/many/m0100"], [100.0, 100.0])


def test_rejects_what_the_subset_cannot_store(tmp_path):
    w = H.Writer(str(tmp_path / "x.h5"))
    with pytest.raises(TypeError):
        w.write("c", np.zeros(3, dtype=np.complex128))
    w.write("a", np.zeros(3))
    with pytest.raises(ValueError):
        w.write("a/b", np.zeros(3))


def test_h5py_reads_the_file_when_it_is_installed(tmp_path):
    h5py = pytest.importorskip("h5py")
    p, data = _sample(tmp_path, many=70)
    with h5py.File(p, "r") as f:
        for k, v in data.items():
            assert np.array_equal(f[k][()], v)
        assert f.attrs["openPMD"] in ("1.1.0", b"1.1.0")
        assert f["data/3"].attrs["dt"] == 0.1
