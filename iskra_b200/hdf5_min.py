"""A minimal HDF5 writer (and reader for what it writes) in pure Python -- the sink of the diagnostics when h5py is absent.

The reference saves its records through HDF5.jl (Diagnostics/src/hdf5.jl:1-95): groups, contiguous datasets of numbers and
scalar / vector / string attributes.  That subset is written here in the "classic" on-disk layout every libhdf5 reads
(HDF5 File Format Specification 3.0, sections II-IV): superblock version 0, version-1 object headers, groups as
symbol tables (version-1 B-tree node + local heap + symbol-table nodes), contiguous data layout (version 3), version-1
dataspace / datatype / attribute messages.  No chunking, compression, links other than hard links, or variable-length
types.

NOT VALIDATED AGAINST libhdf5: no HDF5 library exists in the build image (no h5py, pytables, netCDF4, h5dump).  The
writer follows the specification byte for byte and `read()` below parses the same structures back independently
(tests/test_hdf5_min.py); the first thing to do where h5py exists is `h5py.File(path)` on one of these files.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 32, 16          # symbol-table node holds 2*LEAF_K entries, a B-tree node 2*INTERNAL_K children
SNOD_CAP = 2 * LEAF_K
MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT, MSG_ATTRIBUTE, MSG_SYMTAB = 0x0001, 0x0003, 0x0005, 0x0008, 0x000C, 0x0011


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---- messages ----------------------------------------------------------------------------------------------------
def _datatype(dt):
    """Datatype message body (spec IV.A.2.d), version 1."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        exp_size, mant = (11, 52) if size == 8 else (8, 23)
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)      # class 1; mantissa normalisation "implied"; sign bit
        return head + struct.pack("<HHBBBBI", 0, size * 8, mant, exp_size, 0, mant, (1 << (exp_size - 1)) - 1)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)   # class 0; bit 3: signed
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x10, 0, 0, dt.itemsize)          # class 3; null-terminated; UTF-8
    raise TypeError("hdf5_min: unsupported dtype %r" % (dt,))


def _dataspace(shape):
    """Dataspace message body (spec IV.A.2.b), version 1; () = scalar."""
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _as_array(value):
    """numpy array (little endian, C order) an attribute value or a dataset is stored as"""
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        return np.array(value + b"\0", dtype="S%d" % (len(value) + 1))
    if isinstance(value, (list, tuple)) and value and all(isinstance(v, str) for v in value):
        enc = [v.encode("utf-8") for v in value]
        return np.array(enc, dtype="S%d" % (max(len(e) for e in enc) + 1))
    a = np.asarray(value)
    if a.dtype == bool:
        a = a.astype(np.uint8)
    if a.dtype.kind == "U":
        return _as_array([str(v) for v in a.ravel()]).reshape(a.shape)
    if a.dtype.kind not in "fiuS":
        raise TypeError("hdf5_min: unsupported value %r" % (value,))
    return np.array(a, dtype=a.dtype.newbyteorder("<"), order="C")      # (ascontiguousarray would turn 0-d into 1-d)


def _attribute(name, value):
    """Attribute message body (spec IV.A.2.m), version 1: every part padded to a multiple of eight bytes."""
    a = _as_array(value)
    nm = name.encode("utf-8") + b"\0"
    dt, ds = _datatype(a.dtype), _dataspace(a.shape)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()


def _object_header(messages):
    """Version-1 object header (spec IV.A.1.a): 16-byte prefix, messages (type, size, flags, 3 reserved) aligned to 8."""
    body = b"".join(struct.pack("<HHB3x", t, len(_pad8(m)), 0) + _pad8(m) for t, m in messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


# ---- tree ----------------------------------------------------------------------------------------------------------
class _Node:
    def __init__(self):
        self.attrs = {}
        self.children = None      # dict for groups
        self.data = None          # numpy array for datasets


class Writer:
    """w = Writer(path); w.write("a/b/c", array); w.set_attrs("a/b", {...}); w.close()"""

    def __init__(self, path):
        self.path = path
        self.root = _Node()
        self.root.children = {}

    def _lookup(self, name, create_groups=True, leaf=None):
        node = self.root
        parts = [p for p in name.split("/") if p]
        for k, part in enumerate(parts):
            if node.children is None:
                raise ValueError("hdf5_min: %r is a dataset, not a group" % "/".join(parts[:k]))
            nxt = node.children.get(part)
            if nxt is None:
                if k == len(parts) - 1 and leaf is not None:
                    nxt = leaf
                else:
                    nxt = _Node()
                    nxt.children = {}
                node.children[part] = nxt
            node = nxt
        return node

    def write(self, name, array):
        a = _as_array(array)
        node = _Node()
        node.data = a
        got = self._lookup(name, leaf=node)
        if got is not node:
            if got.children:
                raise ValueError("hdf5_min: %r already is a group with members" % name)
            got.children, got.data = None, a

    def set_attrs(self, name, attrs):
        self._lookup(name).attrs.update(attrs)

    # -- layout: every structure gets its address first, then the bytes are produced
    def close(self):
        order = []

        def walk(node):
            order.append(node)
            if node.children is not None:
                for k in sorted(node.children, key=lambda s: s.encode("utf-8")):
                    walk(node.children[k])
        walk(self.root)
        pos = 96                                                                     # superblock with the root entry
        for node in order:
            attrs = [(MSG_ATTRIBUTE, _attribute(k, v)) for k, v in node.attrs.items()]
            node.addr = pos
            if node.children is not None:
                node.ohdr_size = len(_object_header([(MSG_SYMTAB, b"\0" * 16)] + attrs))
                names = sorted(node.children, key=lambda s: s.encode("utf-8"))
                if len(names) > SNOD_CAP * 2 * INTERNAL_K:
                    raise ValueError("hdf5_min: more than %d members in one group" % (SNOD_CAP * 2 * INTERNAL_K))
                heap, offs = b"\0" * 8, {}
                for nm in names:
                    offs[nm] = len(heap)
                    heap += _pad8(nm.encode("utf-8") + b"\0")
                node.names, node.offs, node.heap = names, offs, heap
                pos += node.ohdr_size
                node.btree = pos
                pos += 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
                node.heap_hdr = pos
                pos += 32
                node.heap_data = pos
                pos += len(heap)
                node.snods = []
                for _ in range(0, len(names), SNOD_CAP):
                    node.snods.append(pos)
                    pos += 8 + SNOD_CAP * 40
            else:
                a = node.data
                msgs = [(MSG_DATASPACE, _dataspace(a.shape)), (MSG_DATATYPE, _datatype(a.dtype)),
                        (MSG_FILL, struct.pack("<BBBB", 2, 2, 2, 0)),                # version 2: allocate late, write if set, undefined
                        (MSG_LAYOUT, b"\0" * 18)] + attrs
                node.ohdr_size = len(_object_header(msgs))
                pos += node.ohdr_size
                node.raw = pos if a.nbytes else UNDEF
                pos += a.nbytes + (-a.nbytes % 8)
        eof = pos
        out = bytearray(eof)
        root = self.root
        out[0:96] = (SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0) +
                     struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) +
                     struct.pack("<QQII", 0, root.addr, 1, 0) + struct.pack("<QQ", root.btree, root.heap_hdr))
        for node in order:
            attrs = [(MSG_ATTRIBUTE, _attribute(k, v)) for k, v in node.attrs.items()]
            if node.children is not None:
                oh = _object_header([(MSG_SYMTAB, struct.pack("<QQ", node.btree, node.heap_hdr))] + attrs)
                out[node.addr:node.addr + len(oh)] = oh
                # B-tree node of a group (spec III.A.1): keys are heap offsets of names; key[i+1] = the largest name in child i
                bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(node.snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
                for k, addr in enumerate(node.snods):
                    last = node.names[min(len(node.names), (k + 1) * SNOD_CAP) - 1]
                    bt += struct.pack("<QQ", addr, node.offs[last])
                out[node.btree:node.btree + len(bt)] = bt
                # local heap (spec III.D): no free block (free-list head = 1, H5HL_FREE_NULL)
                out[node.heap_hdr:node.heap_hdr + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(node.heap), 1, node.heap_data)
                out[node.heap_data:node.heap_data + len(node.heap)] = node.heap
                for k, addr in enumerate(node.snods):
                    part = node.names[k * SNOD_CAP:(k + 1) * SNOD_CAP]
                    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
                    for nm in part:
                        sn += struct.pack("<QQII16x", node.offs[nm], node.children[nm].addr, 0, 0)
                    out[addr:addr + len(sn)] = sn
            else:
                a = node.data
                msgs = [(MSG_DATASPACE, _dataspace(a.shape)), (MSG_DATATYPE, _datatype(a.dtype)),
                        (MSG_FILL, struct.pack("<BBBB", 2, 2, 2, 0)),
                        (MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, node.raw, a.nbytes))] + attrs
                oh = _object_header(msgs)
                out[node.addr:node.addr + len(oh)] = oh
                if a.nbytes:
                    out[node.raw:node.raw + a.nbytes] = a.tobytes()
        with open(self.path, "wb") as f:
            f.write(out)


# ---- reader (what the writer produces; used by the tests and by load()) -----------------------------------------------
def _parse_datatype(b):
    cls, bits0, bits1, _, size = struct.unpack_from("<BBBBI", b, 0)
    if cls >> 4 != 1:
        raise ValueError("datatype message version %d" % (cls >> 4))
    cls &= 15
    if cls == 1:
        _, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", b, 8)
        assert (prec, eloc, esize, mloc, msize, bias) in ((64, 52, 11, 0, 52, 1023), (32, 23, 8, 0, 23, 127)) and bits1 == prec - 1
        return np.dtype("<f%d" % size)
    if cls == 0:
        _, prec = struct.unpack_from("<HH", b, 8)
        assert prec == size * 8
        return np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise ValueError("datatype class %d" % cls)


def _parse_dataspace(b):
    ver, rank, flags = struct.unpack_from("<BBB", b, 0)
    assert ver == 1 and flags == 0
    return tuple(struct.unpack_from("<Q", b, 8 + 8 * k)[0] for k in range(rank))


def _value(a):
    if a.dtype.kind == "S":
        dec = np.vectorize(lambda s: s.decode("utf-8"), otypes=[object])
        return dec(a).tolist() if a.shape else a[()].decode("utf-8")
    return a[()] if a.shape == () else a


def read(path):
    """-> ({dataset path: array}, {object path: {attribute: value}}) of a file written by Writer"""
    buf = open(path, "rb").read()
    assert buf[:8] == SIGNATURE and buf[8] == 0 and buf[13] == 8 and buf[14] == 8
    leaf_k, internal_k = struct.unpack_from("<HH", buf, 16)
    base, _, eof, _ = struct.unpack_from("<QQQQ", buf, 24)
    assert base == 0 and eof == len(buf)
    _, root_addr, cache, _ = struct.unpack_from("<QQII", buf, 56)
    arrays, attrs = {}, {}

    def messages(addr):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", buf, addr)
        assert ver == 1
        p, end, out = addr + 16, addr + 16 + size, []
        for _ in range(nmsg):
            t, sz, _ = struct.unpack_from("<HHB", buf, p)
            out.append((t, buf[p + 8:p + 8 + sz]))
            p += 8 + sz
        assert p == end
        return out

    def heap_name(heap_hdr, off):
        assert buf[heap_hdr:heap_hdr + 4] == b"HEAP"
        size, free, data = struct.unpack_from("<QQQ", buf, heap_hdr + 8)
        assert free == 1 and off < size
        e = buf.index(b"\0", data + off)
        return buf[data + off:e].decode("utf-8")

    def visit(addr, path):
        a = {}
        msgs = messages(addr)
        kinds = dict((t, m) for t, m in msgs if t != MSG_ATTRIBUTE)
        for t, m in msgs:
            if t == MSG_ATTRIBUTE:
                ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", m, 0)
                assert ver == 1
                p = 8
                name = m[p:p + nsz - 1].decode("utf-8")
                p += nsz + (-nsz % 8)
                dt = _parse_datatype(m[p:p + dsz])
                p += dsz + (-dsz % 8)
                shape = _parse_dataspace(m[p:p + ssz])
                p += ssz + (-ssz % 8)
                n = int(np.prod(shape)) if shape else 1
                a[name] = _value(np.frombuffer(m, dtype=dt, count=n, offset=p).reshape(shape))
        attrs[path or "/"] = a
        if MSG_SYMTAB in kinds:
            btree, heap = struct.unpack_from("<QQ", kinds[MSG_SYMTAB], 0)
            assert buf[btree:btree + 4] == b"TREE"
            ntype, level, used, left, right = struct.unpack_from("<BBHQQ", buf, btree + 4)
            assert ntype == 0 and level == 0 and left == UNDEF and right == UNDEF
            prev = ""
            for k in range(used):
                child, key = struct.unpack_from("<QQ", buf, btree + 24 + 8 + 16 * k)
                assert buf[child:child + 4] == b"SNOD"
                _, _, nsym = struct.unpack_from("<BBH", buf, child + 4)
                assert nsym <= 2 * leaf_k
                for e in range(nsym):
                    off, oh, ctype, _ = struct.unpack_from("<QQII", buf, child + 8 + 40 * e)
                    nm = heap_name(heap, off)
                    assert nm.encode() > prev.encode()            # names strictly increasing (strcmp order)
                    prev = nm
                    visit(oh, path + "/" + nm)
                assert heap_name(heap, key) == prev               # right key = the largest name of the child
        else:
            dt = _parse_datatype(kinds[MSG_DATATYPE])
            shape = _parse_dataspace(kinds[MSG_DATASPACE])
            ver, cls, raw, nbytes = struct.unpack_from("<BBQQ", kinds[MSG_LAYOUT], 0)
            assert ver == 3 and cls == 1
            n = int(np.prod(shape)) if shape else 1
            assert nbytes == n * dt.itemsize
            arrays[path] = np.frombuffer(buf, dtype=dt, count=n, offset=raw if n else 0).reshape(shape).copy()
    visit(root_addr, "")
    return arrays, attrs
