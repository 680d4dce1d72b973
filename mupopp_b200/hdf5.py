"""Minimal pure-Python HDF5 reader for the files DOLFIN's `HDF5File.write(mesh, "/mesh")` / `write(meshfunction, "/boundaries")` produce
(h5py is not available in this image).

Serves `hdf = HDF5File(mesh.mpi_comm(), "hdf5_sample_mesh.xml.h5", "r"); hdf.read(mesh, "/mesh", False); hdf.read(boundaries, "/boundaries")`
(/root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:77-85).  Supported subset of the HDF5 file format (the one those
files use): superblock version 0, version-1 object headers (with continuation blocks), old-style groups (symbol-table message, v1
B-tree + SNOD nodes + local heap), dataspace versions 1/2, fixed-point and IEEE floating-point little-endian datatypes, CONTIGUOUS
data layout (layout message version 3).  Chunked or filtered datasets raise NotImplementedError.
"""
from __future__ import annotations

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5Reader:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.f = fh.read()
        f = self.f
        if f[:8] != _SIG:
            raise ValueError("%s is not an HDF5 file" % path)
        if f[8] != 0:
            raise NotImplementedError("HDF5 superblock version %d (only version 0 is supported)" % f[8])
        if f[13] != 8 or f[14] != 8:
            raise NotImplementedError("HDF5 files with %d-byte offsets / %d-byte lengths" % (f[13], f[14]))
        self.base = self._u8(24)
        self.root = self._u8(64)                       # object header of the root group (root symbol-table entry at offset 56)
        self.datasets = {}
        self._walk("", self.root)

    # ---- primitives
    def _u8(self, o):
        return struct.unpack_from("<Q", self.f, o)[0]

    def _u4(self, o):
        return struct.unpack_from("<I", self.f, o)[0]

    def _u2(self, o):
        return struct.unpack_from("<H", self.f, o)[0]

    def _messages(self, addr):
        f = self.f
        if f[addr] != 1:
            raise NotImplementedError("HDF5 object header version %d (only version 1 is supported)" % f[addr])
        nmsg, hsize = self._u2(addr + 2), self._u4(addr + 8)
        blocks, msgs = [(addr + 16, hsize)], []
        while blocks and len(msgs) < nmsg:
            o, sz = blocks.pop(0)
            end = o + sz
            while o + 8 <= end and len(msgs) < nmsg:
                t, s = self._u2(o), self._u2(o + 2)
                d = o + 8
                if t == 0x10:                                        # continuation block
                    blocks.append((self.base + self._u8(d), self._u8(d + 8)))
                msgs.append((t, s, d))
                o = d + s
        return msgs

    def _heap_name(self, heap, off):
        if self.f[heap:heap + 4] != b"HEAP":
            raise ValueError("corrupt HDF5 local heap")
        data = self.base + self._u8(heap + 24)
        end = self.f.index(b"\0", data + off)
        return self.f[data + off:end].decode()

    def _group_entries(self, btree, heap, out):
        f = self.f
        if f[btree:btree + 4] != b"TREE":
            raise ValueError("corrupt HDF5 group B-tree")
        level, n = f[btree + 5], self._u2(btree + 6)
        o = btree + 24
        for _ in range(n):
            child = self.base + self._u8(o + 8)
            o += 16
            if level > 0:
                self._group_entries(child, heap, out)
                continue
            if f[child:child + 4] != b"SNOD":
                raise ValueError("corrupt HDF5 symbol-table node")
            for k in range(self._u2(child + 6)):
                e = child + 8 + 40 * k
                out.append((self._heap_name(heap, self._u8(e)), self.base + self._u8(e + 8)))

    def _walk(self, path, oh):
        f = self.f
        shape = dtype = layout = None
        for t, s, d in self._messages(oh):
            if t == 0x11:                                            # symbol table: a group
                sub = []
                self._group_entries(self.base + self._u8(d), self.base + self._u8(d + 8), sub)
                for name, o in sub:
                    self._walk(path + "/" + name, o)
            elif t == 0x01:                                          # dataspace
                ver, rank = f[d], f[d + 1]
                off = d + 8 if ver == 1 else d + 4
                shape = tuple(self._u8(off + 8 * i) for i in range(rank))
            elif t == 0x03:                                          # datatype
                cls, size = f[d] & 15, self._u4(d + 4)
                bits0 = f[d + 1]
                if bits0 & 1:
                    raise NotImplementedError("big-endian HDF5 datatype")
                if cls == 0:
                    dtype = np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % size)
                else:
                    dtype = None                                     # strings etc. (attributes only): not needed
            elif t == 0x08:                                          # data layout
                ver, cls = f[d], f[d + 1]
                if ver != 3:
                    raise NotImplementedError("HDF5 data layout message version %d" % ver)
                if cls == 1:
                    layout = (self.base + self._u8(d + 2), self._u8(d + 10))
                elif cls == 2:
                    layout = "chunked"
            elif t == 0x0b:
                layout = "filtered"
        if shape is not None and layout is not None:
            self.datasets[path] = (shape, dtype, layout)

    # ---- public
    def keys(self):
        return sorted(self.datasets)

    def __contains__(self, path):
        return path in self.datasets

    def read(self, path):
        if path not in self.datasets:
            raise KeyError("%s not in file (datasets: %s)" % (path, ", ".join(self.keys())))
        shape, dtype, layout = self.datasets[path]
        if isinstance(layout, str):
            raise NotImplementedError("%s: %s HDF5 datasets are not supported by this reader (contiguous only)" % (path, layout))
        if dtype is None:
            raise NotImplementedError("%s: unsupported HDF5 datatype" % path)
        addr, size = layout
        n = int(np.prod(shape)) if shape else 1
        if n * dtype.itemsize > size:
            raise ValueError("%s: dataset larger than its storage" % path)
        return np.frombuffer(self.f, dtype=dtype, count=n, offset=addr).reshape(shape).copy()


def read_mesh(path, name="/mesh"):
    """(coords (V, d) float64, cells (C, d+1) int32 with ascending vertex ids per cell) from DOLFIN's HDF5 mesh group."""
    r = H5Reader(path)
    coords = np.ascontiguousarray(r.read(name + "/coordinates"), dtype=np.float64)
    cells = np.sort(r.read(name + "/topology").astype(np.int64), axis=1).astype(np.int32)
    return coords, cells


def read_facet_markers(path, mesh, name="/boundaries"):
    """DOLFIN MeshFunction("size_t", mesh, dim-1) written with HDF5File.write: `topology` lists every facet by its vertices, `values`
    the marker.  Returns (cell, local facet, value) rows for the facets with a non-zero marker (UFC: local facet k is opposite
    local vertex k)."""
    r = H5Reader(path)
    topo = np.sort(r.read(name + "/topology").astype(np.int64), axis=1)
    vals = r.read(name + "/values").astype(np.int64)
    sel = vals != 0
    topo, vals = topo[sel], vals[sel]
    d = mesh.dim
    nv = mesh.num_vertices()
    cells = mesh.cells.astype(np.int64)

    def key(fv):
        k = fv[:, 0].copy()
        for j in range(1, d):
            k = k * nv + fv[:, j]
        return k

    owner = {}
    for kloc in range(d + 1):
        loc = [j for j in range(d + 1) if j != kloc]
        ks = key(cells[:, loc])
        for c in range(cells.shape[0] - 1, -1, -1):            # lowest cell index wins for interior facets
            owner[int(ks[c])] = (c, kloc)
    rows = []
    for k, v in zip(key(topo), vals):
        c, kloc = owner[int(k)]
        rows.append((c, kloc, int(v)))
    return np.array(rows, dtype=np.int64).reshape(-1, 3)
