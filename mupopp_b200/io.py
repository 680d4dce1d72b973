"""Output files: VTK unstructured-grid (.vtu, ASCII XML) + .pvd time-series index, what `File("x.pvd", "compressed") << f` produces in
the reference scripts (e.g. /root/reference/src/run/CCS2D/CCS_3comp.py:52-60,266-275).  Host-side I/O only; fields are copied from the
device once per written frame.  Like DOLFIN's pvd writer, Lagrange fields are written at the mesh vertices."""
from __future__ import annotations

import os

import numpy as np

_VTK_TYPE = {2: 5, 3: 10}      # triangle, tetrahedron


def write_vtu(path, coords, cells, point_data):
    """coords (V,d), cells (C,d+1); point_data: name -> (V,) scalar or (V,k) vector array."""
    coords = np.asarray(coords, dtype=float)
    cells = np.asarray(cells)
    V, d = coords.shape
    C, nv = cells.shape
    xyz = np.zeros((V, 3))
    xyz[:, :d] = coords
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1">\n<UnstructuredGrid>\n')
        f.write('<Piece NumberOfPoints="%d" NumberOfCells="%d">\n<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n' % (V, C))
        f.write(" ".join("%.16g" % v for v in xyz.ravel()))
        f.write('\n</DataArray>\n</Points>\n<Cells>\n<DataArray type="Int32" Name="connectivity" format="ascii">\n')
        f.write(" ".join(str(int(v)) for v in cells.ravel()))
        f.write('\n</DataArray>\n<DataArray type="Int32" Name="offsets" format="ascii">\n')
        f.write(" ".join(str(nv * (i + 1)) for i in range(C)))
        f.write('\n</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        f.write(" ".join([str(_VTK_TYPE[d])] * C))
        f.write("\n</DataArray>\n</Cells>\n<PointData>\n")
        for name, a in point_data.items():
            a = np.asarray(a, dtype=float)
            if a.ndim == 1:
                f.write('<DataArray type="Float64" Name="%s" format="ascii">\n' % name)
                f.write(" ".join("%.16g" % v for v in a))
            else:
                b = np.zeros((V, 3))
                b[:, :a.shape[1]] = a
                f.write('<DataArray type="Float64" Name="%s" NumberOfComponents="3" format="ascii">\n' % name)
                f.write(" ".join("%.16g" % v for v in b.ravel()))
            f.write("\n</DataArray>\n")
        f.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")


class PvdSeries:
    """name.pvd + name000000.vtu, name000001.vtu, ...  (DOLFIN's naming)."""

    def __init__(self, name):
        self.base = name[:-4] if name.endswith(".pvd") else name
        self.frames = []

    def write(self, coords, cells, point_data, time=None):
        d = os.path.dirname(self.base)
        if d:
            os.makedirs(d, exist_ok=True)
        k = len(self.frames)
        fn = "%s%06d.vtu" % (self.base, k)
        write_vtu(fn, coords, cells, point_data)
        self.frames.append((float(k if time is None else time), os.path.basename(fn)))
        with open(self.base + ".pvd", "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
            for t, name in self.frames:
                f.write('<DataSet timestep="%.16g" part="0" file="%s" />\n' % (t, name))
            f.write("</Collection>\n</VTKFile>\n")
        return fn
