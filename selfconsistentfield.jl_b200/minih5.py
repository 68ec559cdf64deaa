"""Minimal HDF5 reader, plus an even smaller writer (struct + zlib only).

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

`write_h5` produces the same dialect (superblock v0, v1 object headers, one old-style
root group, contiguous little-endian datasets, fixed-length string attributes): what the
result dump (`dump.py`, reference `src/util/dump_molsturm_hdf5.jl`) needs.  It is verified
against this module's reader only — no libhdf5 exists here to cross-check it.

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

    def attrs(self, path="/"):
        """Attributes of a group or dataset: fixed-length strings and numeric scalars/arrays
        (v1 attribute messages; variable-length strings are resolved through the global heap)."""
        out = {}
        for mtype, _, p in self._messages(self._lookup(path)["ohdr"] + self.base):
            if mtype != 0x000C:
                continue
            ver, _, nlen, tlen, slen = struct.unpack_from("<BBHHH", p, 0)
            if ver != 1:
                raise H5FormatError(f"attribute message version {ver} unsupported (only v1)")
            pad = lambda x: (x + 7) // 8 * 8
            pos = 8
            name = p[pos:pos + nlen].split(b"\0")[0].decode("utf-8")
            pos += pad(nlen)
            tmsg = p[pos:pos + tlen]
            pos += pad(tlen)
            smsg = p[pos:pos + slen]
            pos += pad(slen)
            rank = smsg[1]
            dims = struct.unpack_from(f"<{rank}Q", smsg, 8 if smsg[0] == 1 else 4) if rank else ()
            count = int(np.prod(dims)) if dims else 1
            cls, size = tmsg[0] & 0x0F, struct.unpack_from("<I", tmsg, 4)[0]
            raw = p[pos:pos + count * size]
            if cls == 3:      # fixed-length string
                vals = [raw[i * size:(i + 1) * size].split(b"\0")[0].decode("utf-8") for i in range(count)]
                out[name] = vals[0] if dims == () else vals
            elif cls == 9:
                vals = [self._vlen_string(bytes(raw[16 * i:16 * i + 16])) for i in range(count)]
                out[name] = vals[0] if dims == () else vals
            else:
                _kind, np_dtype, _ = self._datatype(tmsg)
                arr = np.frombuffer(raw, dtype=np_dtype, count=count)
                out[name] = arr[0] if dims == () else arr.reshape(dims).copy()
        return out

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


# ---- writer ---------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    """Datatype message (version 1) for little-endian ints, floats and fixed-length strings."""
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        bias = (1 << (exp_bits - 1)) - 1
        head = struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu" or dt.kind == "b":
        signed = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, signed, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)     # null-terminated ASCII
    raise H5FormatError(f"cannot write dtype {dt}")


def _dataspace_message(shape) -> bytes:
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype: int, payload: bytes) -> bytes:
    payload = _pad8(payload)
    return struct.pack("<HHBBBB", mtype, len(payload), 0, 0, 0, 0) + payload


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


def _as_le_array(value):
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype.kind == "S":
        return a.astype(f"S{a.dtype.itemsize + 1}")        # room for the terminating NUL
    if a.dtype.kind == "b":
        return a.astype("<u1")
    if a.dtype.kind not in "iuf":
        raise H5FormatError(f"cannot write dtype {a.dtype}")
    return a.astype(a.dtype.newbyteorder("<"))            # (tobytes() below always emits C order)


def write_h5(path, datasets: dict, attrs: dict | None = None):
    """Write root-level datasets (numeric scalars/arrays, C order as given) and root attributes
    (strings or numeric scalars) as an HDF5 file in the dialect described in the module header.
    File layout: superblock | root object header | B-tree node | local heap | symbol node |
    per dataset: object header, raw data."""
    names = sorted(datasets, key=lambda n: n.encode("utf-8"))          # symbol nodes are strcmp-sorted
    if any("/" in n or not n for n in names):
        raise H5FormatError("write_h5 writes root-level datasets only")
    LEAF_K, INTERNAL_K = max(4, (len(names) + 1) // 2 + 1), 16         # one symbol node holds 2*LEAF_K

    # local heap data: offset 0 is the empty string, then the names, each padded to 8 bytes
    heap, name_off = bytearray(b"\0" * 8), {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode("utf-8") + b"\0")
    heap += b"\0" * 16                                                 # one free block (offset, size)
    free_off = len(heap) - 16
    struct.pack_into("<QQ", heap, free_off, 1, 16)                     # H5HL_FREE_NULL, block size

    attr_msgs = []
    for k, v in (attrs or {}).items():
        a = _as_le_array(v)
        tmsg, smsg = _dtype_message(a.dtype), _dataspace_message(a.shape)
        nm = k.encode("utf-8") + b"\0"
        attr_msgs.append(_message(0x000C, struct.pack("<BBHHH", 1, 0, len(nm), len(tmsg), len(smsg))
                                  + _pad8(nm) + _pad8(tmsg) + _pad8(smsg) + a.tobytes()))

    SUPER = 96
    root_hdr_len = len(_object_header([_message(0x0011, b"\0" * 16)] + attr_msgs))
    btree_at = SUPER + root_hdr_len
    btree_len = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    heap_at = btree_at + btree_len
    heap_data_at = heap_at + 32
    snod_at = heap_data_at + len(heap)
    snod_len = 8 + 2 * LEAF_K * 40
    pos = snod_at + snod_len

    # datasets
    blobs, entries = [], []
    for n in names:
        a = _as_le_array(datasets[n])
        data = a.tobytes()
        layout = struct.pack("<BBQQ", 3, 1, 0, len(data))              # address patched below
        # same message set as libhdf5's own files: dataspace, datatype, fill value (v2: late
        # allocation, write-if-set, defined, size 0), contiguous layout v3
        head = [_message(0x0001, _dataspace_message(a.shape)), _message(0x0003, _dtype_message(a.dtype)),
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0))]
        data_at = pos + len(_object_header(head + [_message(0x0008, layout)]))
        layout = struct.pack("<BBQQ", 3, 1, data_at if data else UNDEF, len(data))
        hdr = _object_header(head + [_message(0x0008, layout)])
        entries.append((name_off[n], pos))
        blobs.append(hdr + _pad8(data))
        pos = data_at + len(_pad8(data))
    eof = pos

    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", btree_at, heap_at))] + attr_msgs)
    assert len(root_hdr) == root_hdr_len
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
    sb += struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, SUPER, 1, 0) + struct.pack("<QQ", btree_at, heap_at)
    assert len(sb) == SUPER

    btree = b"TREE" + struct.pack("<BBH", 0, 0, 1 if names else 0) + struct.pack("<QQ", UNDEF, UNDEF)
    btree += struct.pack("<Q", 0)                                      # key 0: the empty name
    if names:
        btree += struct.pack("<QQ", snod_at, name_off[names[-1]])      # child, key 1 = largest name
    btree += b"\0" * (btree_len - len(btree))
    heap_hdr = b"HEAP" + struct.pack("<BBBB", 0, 0, 0, 0) + struct.pack("<QQQ", len(heap), free_off, heap_data_at)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for noff, ohdr in entries:
        snod += struct.pack("<QQII", noff, ohdr, 0, 0) + b"\0" * 16
    snod += b"\0" * (snod_len - len(snod))

    with open(path, "wb") as fh:
        fh.write(sb + root_hdr + btree + heap_hdr + bytes(heap) + snod + b"".join(blobs))
