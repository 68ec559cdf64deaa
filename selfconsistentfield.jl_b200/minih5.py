"""Minimal read-only HDF5 parser (struct + zlib only).

Host-side mirror of the *reading* half of the reference's integral loader
(`/root/reference/src/util/load_integral_hdf5.jl:6-55` goes through HDF5.jl,
which is not available in this environment, nor is h5py / libhdf5).

Supports exactly the subset the reference's integral dumps use
(writers: `/root/reference/scripts/dump_integrals_pyscf.py:21-51`):
superblock v0, v1 object headers (with continuation blocks), old-style groups
(`TREE` type-0 B-tree + local `HEAP` + `SNOD`), datasets with contiguous or
chunked (v1 B-tree, deflate filter, optional shuffle) layout, little-endian
fixed-point / IEEE float scalars, and variable-length strings in the global
heap (`GCOL`).  Anything else raises `H5FormatError` — never a silent guess.

Datasets are returned as C-ordered numpy arrays with the on-disk (C-order)
dimensions.  HDF5.jl hands Julia the same bytes with the dimensions reversed;
callers that want the Julia view use `arr.T` (see `load_integral_hdf5`).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    pass


class MiniH5:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5FormatError("not an HDF5 file")
        if b[8] != 0:
            raise H5FormatError(f"superblock version {b[8]} unsupported (only v0)")
        if b[13] != 8 or b[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths supported")
        self.base = struct.unpack_from("<Q", b, 24)[0]
        # root group symbol table entry at byte 56
        self.root = self._symbol_entry(56)

    # ---- low-level pieces -------------------------------------------------
    def _symbol_entry(self, off):
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", self.buf, off)
        ent = {"name_off": name_off, "ohdr": ohdr, "cache": cache}
        if cache == 1:
            ent["btree"], ent["heap"] = struct.unpack_from("<QQ", self.buf, off + 24)
        return ent

    def _messages(self, addr):
        """Yield (type, flags, payload bytes) of a v1 object header at addr."""
        b = self.buf
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5FormatError(f"object header version {ver} unsupported (only v1)")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, pos)
                payload = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x0010:  # continuation
                    coff, clen = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((coff + self.base, clen))
                else:
                    yield mtype, mflags, payload

    def _heap_data(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        _size, _free, data = struct.unpack_from("<QQQ", b, addr + 8)
        return data + self.base

    def _cstr(self, off):
        end = self.buf.index(b"\0", off)
        return self.buf[off:end].decode("utf-8")

    def _group_entries(self, btree, heap):
        """name -> symbol entry for an old-style group."""
        data = self._heap_data(heap)
        out = {}

        def walk(node):
            b = self.buf
            if b[node:node + 4] != b"TREE":
                raise H5FormatError("bad B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            if ntype != 0:
                raise H5FormatError("expected group B-tree node")
            pos = node + 24  # past sig, type, level, used, 2 siblings
            pos += 8  # key 0
            for _ in range(used):
                child = struct.unpack_from("<Q", b, pos)[0] + self.base
                pos += 16  # child + next key
                if level > 0:
                    walk(child)
                else:
                    if b[child:child + 4] != b"SNOD":
                        raise H5FormatError("bad SNOD signature")
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for i in range(nsym):
                        ent = self._symbol_entry(child + 8 + 40 * i)
                        out[self._cstr(data + ent["name_off"])] = ent

        walk(btree + self.base)
        return out

    def _children(self, ent):
        if ent.get("cache") == 1:
            return self._group_entries(ent["btree"], ent["heap"])
        for mtype, _, payload in self._messages(ent["ohdr"] + self.base):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", payload, 0)
                return self._group_entries(btree, heap)
        return None

    def _lookup(self, path):
        ent = self.root
        for part in [p for p in path.split("/") if p]:
            kids = self._children(ent)
            if kids is None or part not in kids:
                raise KeyError(path)
            ent = kids[part]
        return ent

    def keys(self, path="/"):
        kids = self._children(self._lookup(path))
        return sorted(kids) if kids else []

    # ---- datasets ---------------------------------------------------------
    def read(self, path):
        ent = self._lookup(path)
        dims = None
        dtype = None
        layout = None
        filters = []
        for mtype, _, p in self._messages(ent["ohdr"] + self.base):
            if mtype == 0x0001:
                ver, rank, flags = struct.unpack_from("<BBB", p, 0)
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from(f"<{rank}Q", p, off) if rank else ()
            elif mtype == 0x0003:
                dtype = self._datatype(p)
            elif mtype == 0x0008:
                layout = self._layout(p)
            elif mtype == 0x000B:
                filters = self._filters(p)
        if dims is None or dtype is None or layout is None:
            raise H5FormatError(f"{path}: not a dataset")
        kind, np_dtype, esize = dtype
        count = int(np.prod(dims)) if dims else 1
        if layout[0] == "contiguous":
            addr, _size = layout[1:]
            raw = self.buf[addr + self.base: addr + self.base + count * esize]
            arr = np.frombuffer(raw, dtype=np_dtype if kind != "vlen_str" else np.uint8)
        elif layout[0] == "chunked":
            if kind == "vlen_str":
                raise H5FormatError("chunked vlen strings unsupported")
            arr = self._read_chunked(layout, dims, np_dtype, esize, filters)
        else:
            raise H5FormatError(f"layout {layout[0]} unsupported")
        if kind == "vlen_str":
            strs = [self._vlen_string(bytes(arr[16 * i: 16 * i + 16])) for i in range(count)]
            return strs[0] if dims == () else np.array(strs, dtype=object).reshape(dims)
        arr = arr.reshape(dims).copy()
        return arr

    def shape(self, path):
        """Dataset dimensions (C order, as stored) without reading the data."""
        for mtype, _, p in self._messages(self._lookup(path)["ohdr"] + self.base):
            if mtype == 0x0001:
                ver, rank, _flags = struct.unpack_from("<BBB", p, 0)
                return tuple(struct.unpack_from(f"<{rank}Q", p, 8 if ver == 1 else 4)) if rank else ()
        raise H5FormatError(f"{path}: not a dataset")

    def read_rows(self, path, lo, hi):
        """Rows [lo, hi) of the FIRST (slowest) file dimension of a chunked or contiguous dataset;
        only the chunks that intersect the range are inflated, so a huge ERI can be streamed to
        the device slab by slab (the reference's own TODO, load_integral_hdf5.jl:9-10).
        For `electron_repulsion_bbbb` file row v is exactly Julia's column-major slab eri[:,:,:,v]."""
        ent = self._lookup(path)
        dims = dtype = layout = None
        filters = []
        for mtype, _, p in self._messages(ent["ohdr"] + self.base):
            if mtype == 0x0001:
                ver, rank, _flags = struct.unpack_from("<BBB", p, 0)
                dims = struct.unpack_from(f"<{rank}Q", p, 8 if ver == 1 else 4)
            elif mtype == 0x0003:
                dtype = self._datatype(p)
            elif mtype == 0x0008:
                layout = self._layout(p)
            elif mtype == 0x000B:
                filters = self._filters(p)
        if dims is None or dtype is None or layout is None or not dims:
            raise H5FormatError(f"{path}: not an array dataset")
        if not (0 <= lo <= hi <= dims[0]):
            raise ValueError(f"row range [{lo}, {hi}) outside [0, {dims[0]})")
        kind, np_dtype, esize = dtype
        if kind == "vlen_str":
            raise H5FormatError("row reads of strings unsupported")
        sub = (hi - lo,) + tuple(dims[1:])
        if layout[0] == "contiguous":
            row = int(np.prod(dims[1:])) * esize
            start = layout[1] + self.base + lo * row
            return np.frombuffer(self.buf[start:start + (hi - lo) * row], dtype=np_dtype).reshape(sub).copy()
        if layout[0] != "chunked":
            raise H5FormatError(f"layout {layout[0]} unsupported")
        return self._read_chunked(layout, dims, np_dtype, esize, filters, first_dim_range=(lo, hi))

    def _datatype(self, p):
        cls = p[0] & 0x0F
        bits0 = p[1]
        size = struct.unpack_from("<I", p, 4)[0]
        if bits0 & 1 and cls in (0, 1):
            raise H5FormatError("big-endian data unsupported")
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return "int", np.dtype(f"<{'i' if signed else 'u'}{size}"), size
        if cls == 1:
            if size not in (4, 8):
                raise H5FormatError("float size unsupported")
            return "float", np.dtype(f"<f{size}"), size
        if cls == 9:
            if (bits0 & 0x0F) != 1:
                raise H5FormatError("only vlen *strings* supported")
            return "vlen_str", None, size
        raise H5FormatError(f"datatype class {cls} unsupported")

    def _layout(self, p):
        ver, cls = p[0], p[1]
        if ver != 3:
            raise H5FormatError(f"layout message version {ver} unsupported (only v3)")
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", p, 2)
            return ("contiguous", addr, size)
        if cls == 2:
            ndim = p[2]
            btree = struct.unpack_from("<Q", p, 3)[0]
            cdims = struct.unpack_from(f"<{ndim}I", p, 11)
            return ("chunked", btree, cdims[:-1], cdims[-1])
        raise H5FormatError(f"layout class {cls} unsupported")

    def _filters(self, p):
        ver, nf = p[0], p[1]
        if ver != 1:
            raise H5FormatError("filter pipeline version unsupported")
        pos = 8
        out = []
        for _ in range(nf):
            fid, nlen, _flags, ncd = struct.unpack_from("<HHHH", p, pos)
            pos += 8 + ((nlen + 7) // 8) * 8
            pos += 4 * ncd + (4 if ncd % 2 else 0)
            out.append(fid)
        return out

    def _read_chunked(self, layout, dims, np_dtype, esize, filters, first_dim_range=None):
        _, btree, cdims, _el = layout
        rank = len(dims)
        lo, hi = first_dim_range if first_dim_range is not None else (0, dims[0])
        out = np.zeros((hi - lo,) + tuple(dims[1:]), dtype=np_dtype)
        b = self.buf
        chunk_count = int(np.prod(cdims))

        def walk(node):
            if b[node:node + 4] != b"TREE":
                raise H5FormatError("bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            if ntype != 1:
                raise H5FormatError("expected raw-data B-tree node")
            keysize = 8 + 8 * (rank + 1)
            pos = node + 24
            for _ in range(used):
                csize, _mask = struct.unpack_from("<II", b, pos)
                offs = struct.unpack_from(f"<{rank}Q", b, pos + 8)
                child = struct.unpack_from("<Q", b, pos + keysize)[0] + self.base
                pos += keysize + 8
                if level > 0:
                    walk(child)
                    continue
                if offs[0] >= hi or offs[0] + cdims[0] <= lo:
                    continue          # chunk entirely outside the requested rows
                raw = b[child: child + csize]
                for fid in reversed(filters):
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle
                        a = np.frombuffer(raw, dtype=np.uint8).reshape(esize, -1)
                        raw = a.T.tobytes()
                    else:
                        raise H5FormatError(f"filter {fid} unsupported")
                chunk = np.frombuffer(raw, dtype=np_dtype, count=chunk_count).reshape(cdims)
                # edge chunks are stored full-size: crop (and clip to the requested rows)
                r0, r1 = max(offs[0], lo), min(offs[0] + cdims[0], dims[0], hi)
                sl_out = (slice(r0 - lo, r1 - lo),) + tuple(
                    slice(o, min(o + c, d)) for o, c, d in zip(offs[1:], cdims[1:], dims[1:]))
                sl_in = (slice(r0 - offs[0], r1 - offs[0]),) + tuple(
                    slice(0, s_.stop - s_.start) for s_ in sl_out[1:])
                out[sl_out] = chunk[sl_in]

        walk(btree + self.base)
        return out

    def _vlen_string(self, ref):
        length, addr, index = struct.unpack("<IQI", ref)
        b = self.buf
        pos = addr + self.base
        if b[pos:pos + 4] != b"GCOL":
            raise H5FormatError("bad global heap signature")
        csize = struct.unpack_from("<Q", b, pos + 8)[0]
        end = pos + csize
        pos += 16
        while pos + 16 <= end:
            idx, _rc, _, osize = struct.unpack_from("<HHIQ", b, pos)
            if idx == 0:
                break
            if idx == index:
                return b[pos + 16: pos + 16 + length].decode("utf-8")
            pos += 16 + ((osize + 7) // 8) * 8
        raise H5FormatError("global heap object not found")
