"""A small, dependency-free HDF5 writer (and a reader for what it writes).

Why it exists: Trixi's `SaveSolutionCallback` / `save_mesh_file` produce HDF5 files (`solution_*.h5`, `mesh.h5`) that
Trixi2Vtk and restarts read (reference examples/euler_ec_3d.jl:38-41), and this image has neither h5py nor libhdf5.
The writer emits the classic on-disk format that libhdf5 has read since 1.0 and that HDF5.jl produces by default:

  superblock version 0 | root group = object header (v1) with a symbol-table message -> B-tree (v1, one leaf) ->
  one symbol-table node (SNOD) + local heap with the link names | every dataset = object header (v1) with dataspace
  (v1), datatype (v1), contiguous data layout (v3) and attribute (v1) messages | raw data contiguous, 8-byte aligned.

Only what those files need is supported: one group (the root), attributes on the root and on datasets, little-endian
float64 / float32 / int64 / int32 / Bool (one-byte bit field, as HDF5.jl) arrays of any rank, fixed-length UTF-8 strings. Arrays are written in the memory
order given (C order; a Julia reader sees the dimensions reversed, exactly as with any HDF5 file written from C).

`File.read(path)` parses the same structures back (used by the tests and by `load_solution_file`); it is NOT a general
HDF5 reader. Validation status: structural (signatures, alignment, sorted links, round trip); no libhdf5 was available
to open the files with.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 32          # symbol-table node capacity 2 K = 64 links (libhdf5 takes K from the superblock)
INTERNAL_K = 16


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------- message encoders
def _dataspace(shape):
    """Dataspace message, version 1: rank, no maximum dimensions."""
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _datatype(dtype, strlen=None):
    """Datatype message, version 1."""
    if strlen is not None:           # class 3 (string): null-terminated padding, UTF-8
        return struct.pack("<BBBBI", 0x13, 0x10, 0, 0, strlen)
    dt = np.dtype(dtype)
    if dt.kind == "b":               # class 4 (bit field) of one byte: what HDF5.jl writes for Bool
        return struct.pack("<BBBBI", 0x14, 0, 0, 0, 1) + struct.pack("<HH", 0, 8)
    if dt.kind == "f":               # class 1 (floating point), IEEE little-endian, implied leading mantissa bit
        size, (eloc, esize, msize, bias) = dt.itemsize, {8: (52, 11, 52, 1023), 4: (23, 8, 23, 127)}[dt.itemsize]
        return (struct.pack("<BBBBI", 0x11, 0x20, 8 * size - 1, 0, size) +
                struct.pack("<HHBBBBI", 0, 8 * size, eloc, esize, 0, msize, bias))
    if dt.kind in "iu":              # class 0 (fixed point), little-endian, two's complement if signed
        return (struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) +
                struct.pack("<HH", 0, 8 * dt.itemsize))
    raise TypeError(f"hdf5_lite: unsupported dtype {dt}")


def _encode_value(value):
    """-> (datatype message, dataspace message, raw bytes, numpy-ish description)"""
    if isinstance(value, str):
        raw = value.encode("utf-8")
        n = max(len(raw), 1)
        return _datatype(None, strlen=n), _dataspace(()), raw.ljust(n, b"\0")
    a = np.asarray(value)
    if a.dtype.kind == "i" and a.dtype.itemsize not in (1, 4, 8):
        a = a.astype(np.int64)
    if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
        a = a.astype(np.float64)
    shape = a.shape                  # (np.ascontiguousarray would turn a scalar into a 1-vector)
    if a.dtype.kind == "b":
        return _datatype(a.dtype), _dataspace(shape), a.astype(np.uint8).tobytes(order="C")
    return _datatype(a.dtype), _dataspace(shape), a.astype(a.dtype.newbyteorder("<")).tobytes(order="C")


def _attribute(name, value):
    """Attribute message, version 1: name, datatype and dataspace each padded to a multiple of 8 bytes."""
    nm = name.encode("utf-8") + b"\0"
    dt, ds, raw = _encode_value(value)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw


def _message(mtype, data):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages):
    """Version-1 object header with every message in its first chunk."""
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Dataset:
    def __init__(self, name, array):
        self.name = name
        self.attrs = {}
        self.dt, self.ds, self.raw = _encode_value(array)


class File:
    """`f = File(); f.attrs["ndims"] = 3; d = f.create_dataset("variables_1", array); d.attrs["name"] = "rho";
    f.write(path)`"""

    def __init__(self):
        self.attrs = {}
        self.datasets = {}

    def create_dataset(self, name, array):
        if name in self.datasets:
            raise ValueError(f"hdf5_lite: dataset {name!r} exists")
        if len(self.datasets) >= 2 * LEAF_K:
            raise ValueError("hdf5_lite: more links than one symbol-table node holds")
        d = _Dataset(name, array)
        self.datasets[name] = d
        return d

    # --------------------------------------------------------------------------------------------- writing
    def tobytes(self):
        names = sorted(self.datasets, key=lambda s: s.encode("utf-8"))      # links of a node are ordered by name
        # local heap data segment: "" at offset 0, then the names, then one free block
        heap, offs = bytearray(8), {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf-8") + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)                                   # last free block: next = 1, size 16
        # addresses
        pos = 96                                                             # superblock
        root_msgs_attr = [_message(0x000C, _attribute(k, v)) for k, v in self.attrs.items()]
        root_hdr_len = 16 + 8 + 16 + sum(len(m) for m in root_msgs_attr)
        a_root = pos; pos += root_hdr_len
        a_btree = pos; pos += 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        a_heap = pos; pos += 32
        a_heapdata = pos; pos += len(heap)
        a_snod = pos; pos += 8 + 2 * LEAF_K * 40
        hdrs, a_hdr, a_data = {}, {}, {}
        for n in names:
            d = self.datasets[n]
            msgs = [_message(0x0001, d.ds), _message(0x0003, d.dt), None] + \
                   [_message(0x000C, _attribute(k, v)) for k, v in d.attrs.items()]
            hdr_len = 16 + sum(len(m) for m in msgs if m is not None) + 8 + 24
            a_hdr[n] = pos; pos += hdr_len
            hdrs[n] = msgs
        for n in names:
            pos += -pos % 8
            a_data[n] = pos if len(self.datasets[n].raw) else UNDEF
            pos += len(self.datasets[n].raw)
        eof = pos
        out = bytearray()
        # superblock, version 0
        out += SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        out += struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        out += struct.pack("<QQI4xQQ", 0, a_root, 1, a_btree, a_heap)        # root symbol-table entry (cached)
        assert len(out) == 96
        # root group object header
        out += _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap))] + root_msgs_attr)
        assert len(out) == a_btree, (len(out), a_btree)
        # B-tree (group node, leaf level): one child
        out += b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        keys = [0, offs[names[-1]] if names else 0]
        body = struct.pack("<QQQ", keys[0], a_snod, keys[1]) if names else b""
        out += body.ljust((2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8, b"\0")
        # local heap
        out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, a_heapdata)
        out += heap
        # symbol-table node
        out += b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            out += struct.pack("<QQI4x16x", offs[n], a_hdr[n], 0)
        out += b"\0" * (40 * (2 * LEAF_K - len(names)))
        # dataset object headers
        for n in names:
            assert len(out) == a_hdr[n]
            d = self.datasets[n]
            msgs = hdrs[n]
            msgs[2] = _message(0x0008, struct.pack("<BBQQ", 3, 1, a_data[n], len(d.raw)))
            out += _object_header(msgs)
        for n in names:
            out += b"\0" * (-len(out) % 8)
            out += self.datasets[n].raw
        assert len(out) == eof
        return bytes(out)

    def write(self, path):
        with open(path, "wb") as f:
            f.write(self.tobytes())

    # --------------------------------------------------------------------------------------------- reading
    @staticmethod
    def read(path):
        """Parse a file written by this module: -> (root attributes, {dataset name: (array, attributes)})."""
        b = open(path, "rb").read()
        if b[:8] != SIGNATURE or b[8] != 0:
            raise ValueError("hdf5_lite.read: not a version-0 superblock")
        (eof,) = struct.unpack_from("<Q", b, 40)
        if eof != len(b):
            raise ValueError("hdf5_lite.read: end-of-file address does not match the file size")
        a_root, cache_type, a_btree, a_heap = struct.unpack_from("<QI4xQQ", b, 64)

        def parse_type_space(dt, ds, raw_at, nbytes=None):
            cls = dt[0] & 0x0F
            (size,) = struct.unpack_from("<I", dt, 4)
            rank = ds[1]
            shape = struct.unpack_from("<" + "Q" * rank, ds, 8) if rank else ()
            count = int(np.prod(shape)) if rank else 1
            raw = b[raw_at: raw_at + count * size]
            if cls == 3:
                return raw.rstrip(b"\0").decode("utf-8")
            if cls == 1:
                np_dt = {8: "<f8", 4: "<f4"}[size]
            elif cls == 0:
                np_dt = ("<i" if dt[1] & 0x08 else "<u") + str(size)
            elif cls == 4 and size == 1:
                a = np.frombuffer(raw, dtype=np.uint8).astype(bool).reshape(shape)
                return a.copy() if rank else bool(a.reshape(()))
            else:
                raise ValueError("hdf5_lite.read: unsupported datatype class")
            a = np.frombuffer(raw, dtype=np_dt).reshape(shape)
            return a.copy() if rank else a.reshape(()).item()

        def parse_header(addr):
            ver, _, nmsg, _, size = struct.unpack_from("<BBHII", b, addr)
            assert ver == 1
            p, msgs = addr + 16, []
            for _ in range(nmsg):
                mtype, msize, _ = struct.unpack_from("<HHB", b, p)
                msgs.append((mtype, p + 8, msize))
                p += 8 + msize
            assert p == addr + 16 + size
            return msgs

        def parse_attr(at):
            ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", b, at)
            assert ver == 1
            p = at + 8
            name = b[p: p + nsz].rstrip(b"\0").decode("utf-8"); p += nsz + (-nsz % 8)
            dt = b[p: p + dsz]; p += dsz + (-dsz % 8)
            ds = b[p: p + ssz]; p += ssz + (-ssz % 8)
            return name, parse_type_space(dt, ds, p)

        root_attrs = {}
        for mtype, at, msize in parse_header(a_root):
            if mtype == 0x0011:
                assert struct.unpack_from("<QQ", b, at) == (a_btree, a_heap)
            elif mtype == 0x000C:
                k, v = parse_attr(at)
                root_attrs[k] = v
        assert b[a_btree: a_btree + 4] == b"TREE" and b[a_heap: a_heap + 4] == b"HEAP"
        _, _, nent = struct.unpack_from("<BBH", b, a_btree + 4)
        _, free_off, a_heapdata = struct.unpack_from("<QQQ", b, a_heap + 8)
        datasets = {}
        if nent:
            _, a_snod, _ = struct.unpack_from("<QQQ", b, a_btree + 24)
            assert b[a_snod: a_snod + 4] == b"SNOD"
            (nsym,) = struct.unpack_from("<H", b, a_snod + 6)
            prev = b""
            for i in range(nsym):
                name_off, a_hdr = struct.unpack_from("<QQ", b, a_snod + 8 + 40 * i)
                end = b.index(b"\0", a_heapdata + name_off)
                raw_name = b[a_heapdata + name_off: end]
                assert raw_name > prev, "links of a symbol-table node must be sorted"
                prev = raw_name
                dt = ds = None
                attrs, a_data, nbytes = {}, None, 0
                for mtype, at, msize in parse_header(a_hdr):
                    if mtype == 0x0001:
                        ds = b[at: at + msize]
                    elif mtype == 0x0003:
                        dt = b[at: at + msize]
                    elif mtype == 0x0008:
                        ver, cls, a_data, nbytes = struct.unpack_from("<BBQQ", b, at)
                        assert (ver, cls) == (3, 1) and (a_data == UNDEF or a_data % 8 == 0)
                    elif mtype == 0x000C:
                        k, v = parse_attr(at)
                        attrs[k] = v
                arr = parse_type_space(dt, ds, 0 if a_data == UNDEF else a_data)
                datasets[raw_name.decode("utf-8")] = (arr, attrs)
        return root_attrs, datasets
