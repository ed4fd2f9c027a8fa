"""Host mirror of the XDMF module (XDMF/src/XDMF.jl, fields.jl, particles.jl): XML descriptors that let
ParaView / VisIt read the openPMD-HDF5 files of the diagnostics sink.  Same API -- `xdmf(func, iterations)` hands an open
iteration file to `func`, `new_document()`, `write_fields`, `write_species`, `write_probes`, `save_document` -- and the same
elements and attributes; the files are read through h5py when it is installed, else through hdf5_min.read.

    fields = new_document()
    xdmf(lambda it: write_fields(it, fields), range(1, ts + 1), prefix="/tmp/04_mcc")
    save_document(fields, "fields", prefix="/tmp/04_mcc")
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import hdf5_min


class _Node:
    """the part of an HDF5 group / dataset the writers below look at"""

    def __init__(self, path, arrays, attrs):
        self.path, self._arrays, self._attrs = path, arrays, attrs

    @property
    def is_dataset(self):
        return self.path in self._arrays

    @property
    def shape(self):          # dimensions as stored in the file (C order)
        return self._arrays[self.path].shape

    def attr(self, name):
        return self._attrs[self.path][name]

    def keys(self):
        pre = self.path.rstrip("/") + "/"
        names = set()
        for p in list(self._arrays) + list(self._attrs):
            if p.startswith(pre) and p != pre:
                names.add(p[len(pre):].split("/")[0])
        return sorted(names)

    def __getitem__(self, name):
        return _Node(self.path.rstrip("/") + "/" + name, self._arrays, self._attrs)


class XDMFFile:
    """XDMFFile(iteration, file)  XDMF.jl:15-18"""

    def __init__(self, iteration, filename):
        self.iteration, self.filename = int(iteration), filename
        try:
            import h5py
        except ImportError:
            arrays, attrs = hdf5_min.read(filename)
        else:
            arrays, attrs = {}, {}
            with h5py.File(filename, "r") as f:
                attrs["/"] = dict(f.attrs)

                def visit(n, o):
                    attrs["/" + n] = dict(o.attrs)
                    if hasattr(o, "shape"):
                        arrays["/" + n] = o[()]
                f.visititems(visit)
        self.root = _Node("", arrays, attrs)

    def __getitem__(self, path):
        return self.root[path.strip("/")]


def new_document():
    """new_document()  XDMF.jl:23-31: <Xdmf Version="3.0"><Domain><Grid GridType="Collection" CollectionType="Temporal"/>"""
    root = ET.Element("Xdmf", {"Version": "3.0"})
    domain = ET.SubElement(root, "Domain")
    ET.SubElement(domain, "Grid", {"GridType": "Collection", "CollectionType": "Temporal"})
    return ET.ElementTree(root)


def save_document(xdoc, filename, prefix="."):
    """save_document(xdoc, filename)  XDMF.jl:33-35 -> <prefix>/xdmf/<filename>.xdmf"""
    os.makedirs(os.path.join(prefix, "xdmf"), exist_ok=True)
    path = os.path.join(prefix, "xdmf", filename + ".xdmf")
    ET.indent(xdoc, space="  ")
    xdoc.write(path, xml_declaration=True, encoding="utf-8")
    return path


def xdmf(func, iterations, prefix="."):
    """xdmf(func, iterations)  XDMF.jl:37-48: opens <prefix>/hdf5/data<i>.h5 for every i and calls func(XDMFFile)"""
    for i in iterations:
        os.makedirs(os.path.join(prefix, "xdmf"), exist_ok=True)
        func(XDMFFile(i, os.path.join(prefix, "hdf5", "data%d.h5" % i)))


def _temporal(xdoc):
    return xdoc.getroot().find("Domain").find("Grid")


def _ref(x, node):
    """"<absolute file>:<dataset path>" -- what the reference writes with pwd()/fname:name(g)"""
    return "%s:%s" % (os.path.abspath(x.filename), node.path)


def _g(v):
    return "%g" % float(v)


def _add_field(fields, x, name, node, origin, spacing):
    """add_field  fields.jl:81-106; a group of components recurses with the component appended to the name"""
    if not node.is_dataset:
        out = None
        for m in node.keys():
            out = _add_field(fields, x, name + m, node[m], origin, spacing)
        return out
    dims = node.shape                                  # file order = (ny, nx) for an (nx, ny) record, like HDF5.jl leaves it
    o = "0.0 %s %s" % (_g(origin[0]), _g(origin[1]))
    s = "0.0 %s %s" % (_g(spacing[0]), _g(spacing[1]))
    d = "1 %d %d" % (dims[0], dims[1])
    att = ET.SubElement(fields, "Attribute", {"Name": name, "AttributeType": "Scalar", "Center": "Node"})
    item = ET.SubElement(att, "DataItem", {"Format": "HDF5", "NumberType": "Float", "Precision": "8", "Dimensions": d})
    item.text = _ref(x, node)
    return o, s, d


def write_fields(x, xdoc):
    """write_fields(xdmf, xdoc)  fields.jl:1-36"""
    fields = ET.SubElement(_temporal(xdoc), "Grid", {"Name": "Fields", "GridType": "Uniform"})
    time = ET.SubElement(fields, "Time")
    topology = ET.SubElement(fields, "Topology", {"TopologyType": "3DCoRectMesh"})
    geometry = ET.SubElement(fields, "Geometry", {"GeometryType": "ORIGIN_DXDYDZ"})
    common = {"Dimensions": "3", "NumberType": "Float", "Precision": "4", "Format": "XML"}
    origin = ET.SubElement(geometry, "DataItem", dict(common, Name="Origin"))
    spacing = ET.SubElement(geometry, "DataItem", dict(common, Name="Spacing"))
    it = x["data/%d" % x.iteration]
    fl = x["data/%d/fields" % x.iteration]
    last = None
    for n in fl.keys():
        last = _add_field(fields, x, n, fl[n], np.ravel(fl[n].attr("gridGlobalOffset")), np.ravel(fl[n].attr("gridSpacing"))) or last
    if last is not None:
        origin.text, spacing.text = last[0], last[1]
        topology.set("Dimensions", last[2])
    time.set("Value", _g(it.attr("time")))


def write_probes(x, xdoc):
    """write_probes(xdmf, xdoc)  fields.jl:38-79: every record of the iteration as a one-point Polyvertex attribute"""
    fields = ET.SubElement(_temporal(xdoc), "Grid", {"Name": "Fields", "GridType": "Uniform"})
    time = ET.SubElement(fields, "Time")
    topology = ET.SubElement(fields, "Topology", {"TopologyType": "Polyvertex", "Dimensions": "1"})
    geometry = ET.SubElement(fields, "Geometry", {"GeometryType": "X_Y_Z"})
    common = {"Dimensions": "1", "NumberType": "Float", "Precision": "4", "Format": "XML"}
    origin = ET.SubElement(geometry, "DataItem", dict(common, Name="Origin"))
    spacing = ET.SubElement(geometry, "DataItem", dict(common, Name="Spacing"))
    it = x["data/%d" % x.iteration]
    fl = x["data/%d/fields" % x.iteration]
    time.set("Value", _g(it.attr("time")))
    for n in fl.keys():
        att = ET.SubElement(fields, "Attribute", {"Name": n, "AttributeType": "Scalar", "Center": "Node"})
        item = ET.SubElement(att, "DataItem", {"Format": "HDF5", "NumberType": "Float", "Precision": "8", "Dimensions": "1"})
        item.text = _ref(x, fl[n])
    for m in "xyz":
        item = ET.SubElement(geometry, "DataItem", {"Name": m, "Format": "XML", "NumberType": "Float", "Precision": "8", "Dimensions": "1"})
        item.text = "0"
    origin.text, spacing.text = "0", "0"


def write_species(x, xdoc, species):
    """write_species(xdmf, xdoc, species)  particles.jl:1-21 (+ add_species :23-51)"""
    particles = ET.SubElement(_temporal(xdoc), "Grid", {"Name": species + " Particles", "GridType": "Uniform"})
    time = ET.SubElement(particles, "Time")
    topology = ET.SubElement(particles, "Topology")
    geometry = ET.SubElement(particles, "Geometry")
    it = x["data/%d" % x.iteration]
    g = x["data/%d/particles" % x.iteration][species]
    n = str(int(np.prod(g["id"].shape)))
    for m in g["position"].keys():
        item = ET.SubElement(geometry, "DataItem", {"Name": m, "Format": "HDF5", "NumberType": "Float", "Precision": "8", "Dimensions": n})
        item.text = _ref(x, g["position"][m])
    att = ET.SubElement(particles, "Attribute", {"Name": species + "id", "AttributeType": "Scalar", "Center": "Node"})
    item = ET.SubElement(att, "DataItem", {"Name": "id", "Format": "HDF5", "NumberType": "UInt", "Dimensions": n})
    item.text = _ref(x, g["id"])
    for m in ("momentum/x", "momentum/y", "momentum/z"):
        node = g[m.split("/")[0]][m.split("/")[1]]
        att = ET.SubElement(particles, "Attribute", {"Name": species + m, "AttributeType": "Scalar", "Center": "Node"})
        item = ET.SubElement(att, "DataItem", {"Name": m, "Format": "HDF5", "NumberType": "Float", "Precision": "8", "Dimensions": n})
        item.text = _ref(x, node)
    time.set("Value", _g(it.attr("time")))
    topology.set("TopologyType", "Polyvertex")
    topology.set("NodesPerElement", "1")
    topology.set("NumberOfElements", n)
    geometry.set("GeometryType", "X_Y_Z")
