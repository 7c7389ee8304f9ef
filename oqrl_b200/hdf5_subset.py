"""Minimal read-only HDF5 reader for the OQRL offline-CartPole files.

Replaces the ``h5py.File(path, "r")[name]`` reads of
``/root/reference/src/dataset.py:11,22-42`` (h5py is not available in the
target image).  It understands exactly the subset of the format those files
use plus the obvious neighbours:

* superblock version 0/1, 8-byte (or 4-byte) offsets and lengths,
* old-style groups (symbol-table message, B-tree v1 type 0, local heap, SNOD),
* version-1 object headers including continuation blocks,
* dataspace v1/v2, datatypes: fixed-point, IEEE float, enum over fixed-point
  (numpy ``bool`` is stored as an i8 enum {FALSE, TRUE}),
* layout v3: compact, contiguous and chunked (B-tree v1 type 1, any depth),
* filters: deflate (id 1) and byte-shuffle (id 2).

Anything else raises ``Hdf5FormatError`` -- there is no silent fallback.
"""
from __future__ import annotations

import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(RuntimeError):
    pass


class _Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters):
        self._f, self.name, self.shape, self.dtype = f, name, tuple(shape), dtype
        self._layout, self._filters = layout, filters

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        return self.read()[key]

    def __len__(self):
        return self.shape[0]

    def read(self) -> np.ndarray:
        return self._f._read_dataset(self)


class File:
    """``File(path)[name]`` -> dataset object convertible with ``np.array``."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise Hdf5FormatError("read-only reader")
        with open(path, "rb") as fh:
            self._buf = fh.read()
        self._parse_superblock()
        self._members = self._read_group(self._root_addr)
        self._cache = {}

    # -- context manager / mapping sugar -------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def keys(self):
        return list(self._members)

    def __contains__(self, name):
        return name in self._members

    def __getitem__(self, name):
        if name not in self._members:
            raise KeyError(name)
        if name not in self._cache:
            self._cache[name] = self._open_dataset(name, self._members[name])
        return self._cache[name]

    # -- low level ---------------------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self._buf[off:off + size], "little")

    def _parse_superblock(self):
        b = self._buf
        base = -1
        off = 0
        while off < len(b):  # signature may sit at 0, 512, 1024, ...
            if b[off:off + 8] == _SIG:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise Hdf5FormatError("not an HDF5 file (signature not found)")
        ver = b[base + 8]
        if ver not in (0, 1):
            raise Hdf5FormatError(f"superblock version {ver} not supported")
        self._so = b[base + 13]  # size of offsets
        self._sl = b[base + 14]  # size of lengths
        if self._so not in (4, 8) or self._sl not in (4, 8):
            raise Hdf5FormatError("unsupported offset/length size")
        p = base + 24 + (4 if ver == 1 else 0)
        self._base_addr = self._u(p, self._so)
        p += 4 * self._so  # base, free-space, eof, driver-info
        # root group symbol table entry: link name offset, object header address, ...
        self._root_addr = self._u(p + self._so, self._so) + self._base_addr

    def _read_group(self, ohdr_addr):
        msgs = self._read_object_header(ohdr_addr)
        stab = [m for m in msgs if m[0] == 0x11]
        if not stab:
            raise Hdf5FormatError("group without symbol-table message (new-style groups unsupported)")
        body = stab[0][1]
        btree = int.from_bytes(body[: self._so], "little")
        heap = int.from_bytes(body[self._so: 2 * self._so], "little")
        heap_data = self._local_heap(heap)
        members = {}
        self._walk_group_btree(btree, heap_data, members)
        return members

    def _local_heap(self, addr):
        if self._buf[addr:addr + 4] != b"HEAP":
            raise Hdf5FormatError("bad local heap")
        p = addr + 8
        size = self._u(p, self._sl)
        p += 2 * self._sl
        data = self._u(p, self._so)
        return self._buf[data:data + size]

    def _walk_group_btree(self, addr, heap, out):
        b = self._buf
        if b[addr:addr + 4] == b"TREE":
            ntype, level, used = b[addr + 4], b[addr + 5], self._u(addr + 6, 2)
            if ntype != 0:
                raise Hdf5FormatError("expected group B-tree")
            p = addr + 8 + 2 * self._so
            p += self._sl  # key 0
            for _ in range(used):
                child = self._u(p, self._so)
                p += self._so + self._sl
                self._walk_group_btree(child, heap, out)
            return
        if b[addr:addr + 4] != b"SNOD":
            raise Hdf5FormatError("bad group node")
        nsym = self._u(addr + 6, 2)
        p = addr + 8
        for _ in range(nsym):
            name_off = self._u(p, self._so)
            ohdr = self._u(p + self._so, self._so)
            end = heap.index(b"\0", name_off)
            out[heap[name_off:end].decode()] = ohdr
            p += 2 * self._so + 4 + 4 + 16

    def _read_object_header(self, addr):
        b = self._buf
        if b[addr] != 1:
            raise Hdf5FormatError(f"object header version {b[addr]} not supported")
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                body = b[p + 8:p + 8 + msize]
                if mtype == 0x10:  # continuation
                    coff = int.from_bytes(body[: self._so], "little")
                    clen = int.from_bytes(body[self._so: self._so + self._sl], "little")
                    blocks.append((coff, clen))
                msgs.append((mtype, body))
                p += 8 + msize
        return msgs

    # -- dataset -------------------------------------------------------------------
    def _open_dataset(self, name, addr):
        msgs = self._read_object_header(addr)
        shape = dtype = layout = None
        filters = []
        for mtype, body in msgs:
            if mtype == 0x1:
                shape = self._parse_dataspace(body)
            elif mtype == 0x3:
                dtype = self._parse_datatype(body)
            elif mtype == 0x8:
                layout = self._parse_layout(body)
            elif mtype == 0xB:
                filters = self._parse_filters(body)
        if shape is None or dtype is None or layout is None:
            raise Hdf5FormatError(f"{name}: not a dataset")
        return _Dataset(self, name, shape, dtype, layout, filters)

    def _parse_dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
        else:
            raise Hdf5FormatError("dataspace version")
        return [int.from_bytes(body[p + i * self._sl: p + (i + 1) * self._sl], "little") for i in range(rank)]

    def _parse_datatype(self, body):
        cls, bits0 = body[0] & 0x0F, body[1]
        size = int.from_bytes(body[4:8], "little")
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:  # fixed point
            signed = (bits0 >> 3) & 1
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:  # float
            if size not in (2, 4, 8):
                raise Hdf5FormatError("float size")
            return np.dtype(f"{order}f{size}")
        if cls == 8:  # enum: base type follows the 8-byte header
            base = self._parse_datatype(body[8:])
            nmemb = body[1] | (body[2] << 8)
            p = 8 + 8 + 4  # enum header + fixed-point base header + properties
            names = []
            for _ in range(nmemb):
                end = body.index(b"\0", p)
                names.append(body[p:end].decode())
                p += ((end - p) // 8 + 1) * 8
            if sorted(names) == ["FALSE", "TRUE"] and base.itemsize == 1:
                return np.dtype(bool)
            return base
        raise Hdf5FormatError(f"datatype class {cls} not supported")

    def _parse_layout(self, body):
        if body[0] != 3:
            raise Hdf5FormatError(f"layout version {body[0]} not supported")
        cls = body[1]
        if cls == 0:
            size = int.from_bytes(body[2:4], "little")
            return ("compact", bytes(body[4:4 + size]))
        if cls == 1:
            addr = int.from_bytes(body[2:2 + self._so], "little")
            size = int.from_bytes(body[2 + self._so:2 + self._so + self._sl], "little")
            return ("contiguous", addr, size)
        if cls == 2:
            rank = body[2]
            addr = int.from_bytes(body[3:3 + self._so], "little")
            p = 3 + self._so
            dims = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank)]
            return ("chunked", addr, dims[:-1], dims[-1])
        raise Hdf5FormatError("layout class")

    def _parse_filters(self, body):
        ver, nf = body[0], body[1]
        if ver not in (1, 2):
            raise Hdf5FormatError("filter pipeline version")
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = int.from_bytes(body[p:p + 2], "little")
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(body[p + 2:p + 4], "little")
                p += 4
            else:
                nlen = 0
                p += 2
            ncd = int.from_bytes(body[p + 2:p + 4], "little")
            p += 4 + nlen
            cd = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            if fid not in (1, 2):
                raise Hdf5FormatError(f"filter id {fid} not supported")
            out.append((fid, cd))
        return out

    def _chunks(self, addr, rank, out):
        b = self._buf
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise Hdf5FormatError("bad chunk B-tree")
        level, used = b[addr + 5], self._u(addr + 6, 2)
        p = addr + 8 + 2 * self._so
        ksz = 8 + 8 * (rank + 1)
        for _ in range(used):
            csize, fmask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * i, 8) for i in range(rank)]
            child = self._u(p + ksz, self._so)
            p += ksz + self._so
            if level == 0:
                out.append((offs, csize, fmask, child))
            else:
                self._chunks(child, rank, out)

    def _unfilter(self, raw, filters, fmask, itemsize):
        for i in reversed(range(len(filters))):
            if fmask & (1 << i):
                continue
            fid, _cd = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            else:  # byte shuffle
                n = len(raw) // itemsize
                a = np.frombuffer(raw[: n * itemsize], dtype=np.uint8).reshape(itemsize, n)
                raw = a.T.tobytes() + raw[n * itemsize:]
        return raw

    def _read_dataset(self, ds):
        shape, dt, layout = ds.shape, ds.dtype, ds._layout
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            return np.frombuffer(layout[1], dtype=dt, count=count).reshape(shape).copy()
        if layout[0] == "contiguous":
            _, addr, size = layout
            if addr == _UNDEF:
                return np.zeros(shape, dtype=dt)
            return np.frombuffer(self._buf, dtype=dt, count=count, offset=addr).reshape(shape).copy()
        _, addr, cdims, esize = layout
        if esize != dt.itemsize or len(cdims) != len(shape):
            raise Hdf5FormatError("chunk layout does not match datatype/dataspace")
        out = np.zeros(shape, dtype=dt)
        if addr == _UNDEF:
            return out
        chunks = []
        self._chunks(addr, len(shape), chunks)
        for offs, csize, fmask, caddr in chunks:
            raw = self._unfilter(self._buf[caddr:caddr + csize], ds._filters, fmask, dt.itemsize)
            c = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl = tuple(slice(o, min(o + d, s)) for o, d, s in zip(offs, cdims, shape))
            cut = tuple(slice(0, s.stop - s.start) for s in sl)
            out[sl] = c[cut]
        return out
