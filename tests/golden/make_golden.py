"""Generate tests/golden/fenics_demo_dofmap.npz from the reference's own fixture files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Source data (reference DATA files, not source code):
  /root/reference/src/run/fenics_demo/dolfin_fine.xml.gz             mesh (2868 vertices, 5400 triangles)
  /root/reference/src/run/fenics_demo/dolfin_fine_velocity.xml.gz    P2-vector function with the explicit
        dof -> (cell_index, cell_dof_index) table written by DOLFIN: the only golden data for UFC dof numbering
  /root/reference/src/run/fenics_demo/dolfin_fine_subdomains.xml.gz  facet markers (cell_index, local_entity, value)
"""
import gzip
import os
import re

import numpy as np

REF = "/root/reference/src/run/fenics_demo/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fenics_demo_dofmap.npz")


def main():
    txt = gzip.open(REF + "dolfin_fine.xml.gz", "rt").read()
    vs = re.findall(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"', txt)
    coords = np.zeros((len(vs), 2))
    for i, x, y in vs:
        coords[int(i)] = (float(x), float(y))
    tr = re.findall(r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"', txt)
    cells = np.zeros((len(tr), 3), dtype=np.int32)
    for i, a, b, c in tr:
        cells[int(i)] = (int(a), int(b), int(c))
    txt = gzip.open(REF + "dolfin_fine_velocity.xml.gz", "rt").read()
    rows = re.findall(r'<dof index="(\d+)" value="([^"]+)" cell_index="(\d+)" cell_dof_index="(\d+)"', txt)
    n = len(rows)
    dof_value = np.zeros(n)
    dof_cell = np.zeros(n, dtype=np.int32)
    dof_local = np.zeros(n, dtype=np.int16)
    for i, v, c, l in rows:
        dof_value[int(i)] = float(v)
        dof_cell[int(i)] = int(c)
        dof_local[int(i)] = int(l)
    txt = gzip.open(REF + "dolfin_fine_subdomains.xml.gz", "rt").read()
    fm = re.findall(r'<value cell_index="(\d+)" local_entity="(\d+)" value="(\d+)"', txt)
    facet = np.array([(int(a), int(b), int(c)) for a, b, c in fm], dtype=np.int32)
    np.savez_compressed(OUT, coords=coords, cells=cells, dof_value=dof_value, dof_cell=dof_cell,
                        dof_local=dof_local, facet_markers=facet)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", n, "dofs")


if __name__ == "__main__" and len(os.sys.argv) == 1:
    main()


def stokes_sample_mesh():
    """tests/golden/stokes_sample_mesh.npz from the reference's HDF5 mesh of config 5 (a DATA file):
    /root/reference/src/run/stokes_ADR_3D/hdf5_sample_mesh.xml.h5 -- /mesh (1582 vertices, 5203 tetrahedra) and /boundaries (facet
    ids 1 inflow, 2 outflow, 3 walls; run/stokes_ADR_3D/master_advective_stokes.py:77-85,119-130), read with mupopp_b200/hdf5.py."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from mupopp_b200 import hdf5
    from mupopp_b200.mesh import Mesh
    src = "/root/reference/src/run/stokes_ADR_3D/hdf5_sample_mesh.xml.h5"
    coords, cells = hdf5.read_mesh(src)
    markers = hdf5.read_facet_markers(src, Mesh(coords, cells))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stokes_sample_mesh.npz")
    np.savez_compressed(out, coords=coords, cells=cells, facet_markers=markers.astype(np.int32))
    print("wrote", out, os.path.getsize(out), "bytes;", coords.shape[0], "vertices,", cells.shape[0], "cells,", markers.shape[0], "marked facets")


if __name__ == "__main__" and len(os.sys.argv) > 1 and os.sys.argv[1] == "stokes":
    stokes_sample_mesh()
