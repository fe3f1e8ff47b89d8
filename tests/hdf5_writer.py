"""Test-only writer of small HDF5 files (no h5py / netCDF4 in this image), laid out the way the netCDF-4 library lays
its files out, to exercise ``terrarium_jl_b200.netcdf4`` beyond what the reference's own mask files contain: chunked
datasets with edge chunks, the shuffle + deflate (+ fletcher32) pipeline, packed big- and little-endian integers with
``scale_factor`` / ``add_offset`` / ``_FillValue``, version-1 and version-2 object headers, symbol-table and
link-message groups, ``DIMENSION_LIST`` references through the global heap.

Written from the HDF5 file format specification (version 3.0, sections II-IV); checksums are left zero (the decoder
does not verify them)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits = big | (0x08 if dt.kind == "i" else 0)
        return bytes([0x10, bits, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        e, m, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        bits = big | 0x20
        return (bytes([0x11, bits, 8 * dt.itemsize - 1, 0]) + struct.pack("<I", dt.itemsize)
                + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, m, e, 0, m, bias))
    if dt.kind == "S":
        return bytes([0x13, 0, 0, 0]) + struct.pack("<I", dt.itemsize)
    raise ValueError(dt)


def _dataspace_msg(shape) -> bytes:
    return bytes([2, len(shape), 0, 1 if shape else 0]) + b"".join(struct.pack("<Q", s) for s in shape)


def _attr_msg(name: str, value, vlen_refs=None) -> bytes:
    nm = name.encode() + b"\0"
    if vlen_refs is not None:       # variable-length sequences of object references (DIMENSION_LIST)
        dt = bytes([0x19, 0, 0, 0]) + struct.pack("<I", 16) + bytes([0x17, 0, 0, 0]) + struct.pack("<I", 8)
        ds = _dataspace_msg((len(vlen_refs),))
        data = b"".join(struct.pack("<IQI", 1, addr, idx) for addr, idx in vlen_refs)
    elif isinstance(value, str):
        raw = value.encode()
        dt, ds, data = _dtype_msg(np.dtype(f"S{len(raw)}")), _dataspace_msg(()), raw
    else:
        a = np.atleast_1d(np.asarray(value))
        dt, ds, data = _dtype_msg(a.dtype), _dataspace_msg(a.shape), a.tobytes()
    return bytes([3, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds)) + b"\0" + nm + dt + ds + data


def _filter_msg(filters, version=2) -> bytes:
    if version == 2:
        out = bytes([2, len(filters)])
        for fid, cd in filters:
            out += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
        return out
    out = bytes([1, len(filters)]) + b"\0" * 6
    for fid, cd in filters:
        name = {1: b"deflate\0", 2: b"shuffle\0", 3: b"fletcher32\0\0\0\0\0\0"}[fid]
        out += struct.pack("<HHHH", fid, len(name), 1, len(cd)) + name + b"".join(struct.pack("<I", c) for c in cd)
        if len(cd) % 2:
            out += b"\0" * 4
    return out


def _ohdr(messages, version=2) -> bytes:
    if version == 2:
        body = b"".join(bytes([t]) + struct.pack("<H", len(b)) + b"\0" + b for t, b in messages)
        return b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body + b"\0" * 4
    body = b""
    for t, b in messages:
        b = b + b"\0" * (-len(b) % 8)
        body += struct.pack("<HHB", t, len(b), 0) + b"\0" * 3 + b
    return bytes([1, 0]) + struct.pack("<HII", len(messages), 1, len(body)) + b"\0" * 4 + body


class Writer:
    """``w = Writer(); w.add(name, array, chunks=..., filters=..., attrs=...); w.save(path)``."""

    def __init__(self, header_version=2, attrs=None):
        self.hv = header_version
        self.vars = []
        self.attrs = attrs or {}

    def add(self, name, array, chunks=None, deflate=0, shuffle=False, fletcher32=False, attrs=None, dims=None, layout="chunked"):
        self.vars.append(dict(name=name, a=np.asarray(array), chunks=chunks, deflate=deflate, shuffle=shuffle,
                              fletcher32=fletcher32, attrs=attrs or {}, dims=dims, layout=layout))

    def save(self, path):
        buf = bytearray(b"\0" * (48 if self.hv == 2 else 96))   # superblock, patched at the end

        def put(b: bytes) -> int:
            buf.extend(b"\0" * (-len(buf) % 8))
            addr = len(buf)
            buf.extend(b)
            return addr

        # raw data first
        for v in self.vars:
            a, dt = v["a"], v["a"].dtype
            if v["layout"] == "contiguous":
                v["data_addr"] = put(a.tobytes())
                continue
            if v["layout"] == "compact":
                continue
            chunks = v["chunks"] or a.shape
            filters = []
            if v["shuffle"]:
                filters.append((2, (dt.itemsize,)))
            if v["deflate"]:
                filters.append((1, (v["deflate"],)))
            if v["fletcher32"]:
                filters.append((3, ()))
            v["filters"] = filters
            entries = []
            grid = [range(0, s, c) for s, c in zip(a.shape, chunks)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = [g[i] for g, i in zip(grid, offs)]
                block = np.zeros(chunks, dtype=dt)
                sel = tuple(slice(oo, min(oo + c, s)) for oo, c, s in zip(o, chunks, a.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sel)] = a[sel]
                raw = block.tobytes()
                for fid, cd in filters:
                    if fid == 2 and dt.itemsize > 1:
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, dt.itemsize).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, cd[0])
                    elif fid == 3:
                        raw = raw + b"\xde\xad\xbe\xef"
                entries.append((o, len(raw), put(raw)))
            # one leaf node per 4 chunks under one internal node: exercises the level-1 descent
            def node(level, items):
                out = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(items), UNDEF, UNDEF)
                for o, size, child in items:
                    out += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", x) for x in o) + struct.pack("<Q", 0)
                    out += struct.pack("<Q", child)
                out += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in a.shape) + struct.pack("<Q", 0)
                return out

            leaves = [entries[i:i + 4] for i in range(0, len(entries), 4)]
            if len(leaves) == 1:
                v["btree"] = put(node(0, leaves[0]))
            else:
                tops = [(leaf[0][0], 0, put(node(0, leaf))) for leaf in leaves]
                v["btree"] = put(node(1, tops))

        # global heap for DIMENSION_LIST references is written once the dimension datasets have addresses:
        # dimension scales (variables named as dimensions) get their object headers first
        order = sorted(self.vars, key=lambda v: 0 if v["dims"] is None else 1)
        addr_of = {}
        for v in order:
            a, dt = v["a"], v["a"].dtype
            msgs = [(0x01, _dataspace_msg(a.shape)), (0x03, _dtype_msg(dt)), (0x05, bytes([3, 0x0A]))]
            if v["layout"] == "contiguous":
                msgs.append((0x08, bytes([3, 1]) + struct.pack("<QQ", v["data_addr"], a.nbytes)))
            elif v["layout"] == "compact":
                msgs.append((0x08, bytes([3, 0]) + struct.pack("<H", a.nbytes) + a.tobytes()))
            else:
                chunks = v["chunks"] or a.shape
                msgs.append((0x08, bytes([3, 2, a.ndim + 1]) + struct.pack("<Q", v["btree"])
                             + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", dt.itemsize)))
                if v["filters"]:
                    msgs.append((0x0B, _filter_msg(v["filters"], version=self.hv)))
            for k, val in v["attrs"].items():
                msgs.append((0x0C, _attr_msg(k, val)))
            if v["dims"] is not None:
                heap = b"GCOL" + bytes([1, 0, 0, 0])
                objs = b""
                for i, d in enumerate(v["dims"]):
                    objs += struct.pack("<HHIQ", i + 1, 1, 0, 8) + struct.pack("<Q", addr_of[d])
                objs += struct.pack("<HHIQ", 0, 0, 0, 0)
                heap += struct.pack("<Q", 16 + len(objs)) + objs
                gaddr = put(heap)
                msgs.append((0x0C, _attr_msg("DIMENSION_LIST", None, vlen_refs=[(gaddr, i + 1) for i in range(len(v["dims"]))])))
            else:
                msgs.append((0x0C, _attr_msg("CLASS", "DIMENSION_SCALE")))
            addr_of[v["name"]] = put(_ohdr(msgs, self.hv))

        # root group
        gattrs = [(0x0C, _attr_msg(k, val)) for k, val in self.attrs.items()]
        if self.hv == 2:
            links = []
            for v in self.vars:
                nm = v["name"].encode()
                links.append((0x06, bytes([1, 0, len(nm)]) + nm + struct.pack("<Q", addr_of[v["name"]])))
            root = put(_ohdr(links + gattrs, 2))
            buf[0:48] = (b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, len(buf), root) + b"\0" * 4)
        else:
            names = b"\0" * 8
            offs = {}
            for v in sorted(self.vars, key=lambda v: v["name"]):
                offs[v["name"]] = len(names)
                names += v["name"].encode() + b"\0"
                names += b"\0" * (-len(names) % 8)
            data_addr = put(names)
            heap = put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(names), UNDEF, data_addr))
            snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(self.vars))
            for v in sorted(self.vars, key=lambda v: v["name"]):
                snod += struct.pack("<QQII", offs[v["name"]], addr_of[v["name"]], 0, 0) + b"\0" * 16
            snod_addr = put(snod)
            tree = put(b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offs[max(offs)]))
            root = put(_ohdr([(0x11, struct.pack("<QQ", tree, heap))] + gattrs, 1))
            sb = (b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
                  + struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF)
                  + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree, heap))
            buf[0:len(sb)] = sb
        with open(path, "wb") as f:
            f.write(bytes(buf))
