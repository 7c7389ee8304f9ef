"""Minimal HDF5 writer used ONLY to build test fixtures (tests/golden/make_golden.py).

Writes the same subset of the format the reference's data files use: superblock v0,
old-style root group (symbol-table message, B-tree v1, local heap, one SNOD), version-1 object headers with
dataspace / datatype / filter-pipeline (deflate) / layout-v3-chunked messages and a single-leaf chunk B-tree.
h5py is not available, so fixtures for oqrl_b200.hdf5_subset are produced with this writer (the reader itself
is validated against the reference's real files in the build container: tests/test_hdf5_dataset.py)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _datatype(dt):
    dt = np.dtype(dt)
    if dt == np.dtype(bool):
        base = struct.pack("<B3BI", 0x10 | 0, 0x08, 0, 0, 1) + struct.pack("<HH", 0, 8)      # signed 1-byte integer
        names = _pad8(b"FALSE\0") + _pad8(b"TRUE\0")
        return struct.pack("<B3BI", 0x10 | 8, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, man_bits, bias = (8, 23, 127) if size == 4 else (11, 52, 1023)
        bits = bytes([0x20, 0x1F if size == 4 else 0x3F, 0])            # LE, mantissa normalisation: implied msb; sign bit position
        props = struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
        return struct.pack("<B3sI", 0x10 | 1, bits, size) + props
    if dt.kind in "iu":
        bits = bytes([0x08 if dt.kind == "i" else 0, 0, 0])
        return struct.pack("<B3sI", 0x10 | 0, bits, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(dt)


class Writer:
    def __init__(self):
        self.buf = bytearray()

    def _alloc(self, data):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _dataset(self, arr, chunk):
        arr = np.ascontiguousarray(arr)
        rank = arr.ndim
        # chunks (deflate level 4), row-major chunk grid
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
        entries = []
        for offs in np.ndindex(*[len(g) for g in grid]):
            start = [g[i] for g, i in zip(grid, offs)]
            block = np.zeros(chunk, dtype=arr.dtype)
            sl = tuple(slice(s, min(s + c, n)) for s, c, n in zip(start, chunk, arr.shape))
            cut = tuple(slice(0, x.stop - x.start) for x in sl)
            block[cut] = arr[sl]
            raw = block.view(np.uint8).tobytes() if arr.dtype != bool else block.astype(np.uint8).tobytes()
            comp = zlib.compress(raw, 4)
            entries.append((start, len(comp), self._alloc(comp)))
        # chunk B-tree v1, single leaf
        node = b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + struct.pack("<QQ", UNDEF, UNDEF)
        for start, csize, addr in entries:
            node += struct.pack("<II", csize, 0) + b"".join(struct.pack("<Q", o) for o in start) + struct.pack("<Q", 0)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
        btree = self._alloc(node)
        itemsize = 1 if arr.dtype == bool else arr.dtype.itemsize
        msgs = [
            _msg(0x1, struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)),
            _msg(0x3, _datatype(arr.dtype)),
            _msg(0xB, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 4) + b"\0" * 4),
            _msg(0x8, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree)
                 + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", itemsize)),
        ]
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def write(self, path, datasets, chunks):
        """datasets: dict name -> array (written in sorted name order, like h5py's symbol table)."""
        self.buf = bytearray(b"\0" * 96)                         # superblock + root symbol-table entry, patched below
        names = sorted(datasets)
        headers = {n: self._dataset(datasets[n], chunks[n]) for n in names}
        heap_data = b"\0" * 8
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        heap_data_addr = self._alloc(heap_data)
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", name_off[n], headers[n], 0, 0)
        snod_addr = self._alloc(snod)
        tree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
        tree += struct.pack("<Q", 0) + struct.pack("<Q", snod_addr) + struct.pack("<Q", name_off[names[-1]])
        tree_addr = self._alloc(tree)
        stab = _msg(0x11, struct.pack("<QQ", tree_addr, heap))
        root = self._alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(stab)) + stab)
        eof = len(self.buf)
        sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        with open(path, "wb") as fh:
            fh.write(self.buf)
