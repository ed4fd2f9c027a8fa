"""XDMF mirror (iskra_b200/xdmf.py; XDMF/src/*.jl): the descriptors written for the files of the diagnostics sink carry the
reference's elements and attributes and point at datasets that exist with the dimensions they announce."""
import os
import xml.etree.ElementTree as ET

import numpy as np

from iskra_b200 import diagnostics as DG
from iskra_b200 import hdf5_min
from iskra_b200 import xdmf as X


class _G:
    n, dh = (4, 3), (0.5, 0.25)


class _S:
    name, np, m, q = "e-", 5, 2.0, -1.0


def _write_run(tmp_path, iterations=(1, 2)):
    DG.records.clear()
    rho = np.arange(12.0).reshape(4, 3)
    E = np.arange(36.0).reshape(4, 3, 3)
    DG.register_field("rho", "C/m^2", lambda: rho, _G())
    DG.register_field("E", "V/m", lambda: E, _G(), withcomponents=True)
    DG.register_particle("e-/position", "m", lambda: np.arange(10.0).reshape(5, 2), _S(), withcomponents=True)
    DG.register_particle("e-/momentum", "kg*m/s", lambda: np.ones((5, 3)), _S(), withcomponents=True)
    DG.register_particle("e-/id", "1", lambda: np.arange(1, 6, dtype=np.uint32), _S())
    prefix = str(tmp_path / "run")
    for i in iterations:
        DG.new_iteration(prefix, i, 0.1 * i, 0.1, lambda it: (DG.save_record(it, "rho"), DG.save_record(it, "E"), DG.save_records(it, "e-/")))
    return prefix


def test_fields_species_and_probes_documents(tmp_path):
    prefix = _write_run(tmp_path)
    fields, electrons, probes = X.new_document(), X.new_document(), X.new_document()
    X.xdmf(lambda it: (X.write_fields(it, fields), X.write_species(it, electrons, "e-"), X.write_probes(it, probes)), [1, 2], prefix=prefix)
    paths = [X.save_document(d, n, prefix=prefix) for d, n in ((fields, "fields"), (electrons, "electrons"), (probes, "probes"))]
    assert all(os.path.exists(p) and p.endswith(".xdmf") for p in paths)
    root = ET.parse(paths[0]).getroot()
    assert root.tag == "Xdmf" and root.get("Version") == "3.0"
    temporal = root.find("Domain").find("Grid")
    assert temporal.get("GridType") == "Collection" and temporal.get("CollectionType") == "Temporal"
    grids = temporal.findall("Grid")
    assert len(grids) == 2 and all(g.get("Name") == "Fields" and g.get("GridType") == "Uniform" for g in grids)
    g = grids[1]
    assert float(g.find("Time").get("Value")) == 0.2
    topo, geo = g.find("Topology"), g.find("Geometry")
    assert topo.get("TopologyType") == "3DCoRectMesh" and topo.get("Dimensions") == "1 3 4"       # file order: (ny, nx)
    assert geo.get("GeometryType") == "ORIGIN_DXDYDZ"
    items = {i.get("Name"): i for i in geo.findall("DataItem")}
    assert items["Origin"].text == "0.0 0 0" and items["Spacing"].text == "0.0 0.5 0.25" and items["Origin"].get("Dimensions") == "3"
    atts = {a.get("Name"): a.find("DataItem") for a in g.findall("Attribute")}
    assert sorted(atts) == ["Ex", "Ey", "Ez", "rho"]                                              # components: name + component
    arrays, _ = hdf5_min.read(os.path.join(prefix, "hdf5", "data2.h5"))
    for name, item in atts.items():
        fname, dset = item.text.rsplit(":", 1)
        assert os.path.isabs(fname) and fname.endswith("hdf5/data2.h5") and item.get("Format") == "HDF5" and item.get("Precision") == "8"
        assert "1 %d %d" % arrays[dset].shape == item.get("Dimensions")
    assert np.array_equal(arrays["/data/2/fields/rho"], np.arange(12.0).reshape(4, 3).T)          # x fastest on disk, like HDF5.jl
    # species document
    sp = ET.parse(paths[1]).getroot().find("Domain").find("Grid").findall("Grid")[0]
    assert sp.get("Name") == "e- Particles"
    t = sp.find("Topology")
    assert (t.get("TopologyType"), t.get("NodesPerElement"), t.get("NumberOfElements")) == ("Polyvertex", "1", "5")
    assert sp.find("Geometry").get("GeometryType") == "X_Y_Z"
    assert [i.get("Name") for i in sp.find("Geometry").findall("DataItem")] == ["x", "y", "z"]
    names = [a.get("Name") for a in sp.findall("Attribute")]
    assert names == ["e-id", "e-momentum/x", "e-momentum/y", "e-momentum/z"]
    assert sp.findall("Attribute")[0].find("DataItem").get("NumberType") == "UInt"
    # probes document: one point, every field record an attribute
    pr = ET.parse(paths[2]).getroot().find("Domain").find("Grid").findall("Grid")[0]
    assert pr.find("Topology").get("TopologyType") == "Polyvertex" and pr.find("Topology").get("Dimensions") == "1"
    assert sorted(a.get("Name") for a in pr.findall("Attribute")) == ["E", "rho"]
