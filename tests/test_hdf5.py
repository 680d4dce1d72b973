"""The pure-Python HDF5 reader behind `HDF5File(...).read(mesh, "/mesh", False)` / `.read(boundaries, "/boundaries")`
(/root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:77-85) and the committed fixture made with it."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_H5 = "/root/reference/src/run/stokes_ADR_3D/hdf5_sample_mesh.xml.h5"
FIX = os.path.join(ROOT, "tests", "golden", "stokes_sample_mesh.npz")


def test_fixture_is_a_consistent_marked_tetrahedral_mesh():
    from mupopp_b200.mesh import Mesh
    z = np.load(FIX)
    coords, cells, mk = z["coords"], z["cells"], z["facet_markers"]
    assert coords.shape == (1582, 3) and cells.shape == (5203, 4) and mk.shape == (2700, 3)
    assert (np.diff(cells, axis=1) > 0).all()                                  # UFC: ascending vertex ids per cell
    x = coords[cells]
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / 6.0
    assert vol.min() > 0 and abs(vol.sum() - 80.0) < 1e-3                        # the box [-10,10] x [-1,1]^2
    m = Mesh(coords, cells)
    fv, n, meas, owner = m.boundary_facets()
    loc = m.boundary_facet_local_index()
    assert fv.shape[0] == 2700                                                  # every boundary facet carries a marker ...
    assert set(zip(owner.tolist(), loc.tolist())) == set(zip(mk[:, 0].tolist(), mk[:, 1].tolist()))
    ids, cnt = np.unique(mk[:, 2], return_counts=True)
    assert ids.tolist() == [1, 2, 3] and cnt.tolist() == [450, 450, 1800]       # inflow, outflow, walls
    # ... and the ids sit where master_advective_stokes.py:119-130 expects them
    tri = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    pts = coords[cells[mk[:, 0]][np.arange(len(mk))[:, None], tri[mk[:, 1]]]]
    assert np.allclose(pts[mk[:, 2] == 1][..., 0], -10.0) and np.allclose(pts[mk[:, 2] == 2][..., 0], 10.0, atol=1e-4)


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="the reference's HDF5 mesh is only present in the build container")
def test_reader_reproduces_the_fixture_from_the_reference_file():
    import mupopp_b200.fenics as fx
    from mupopp_b200 import hdf5
    r = hdf5.H5Reader(REF_H5)
    assert r.keys() == ["/boundaries/coordinates", "/boundaries/topology", "/boundaries/values", "/mesh/cell_indices", "/mesh/coordinates",
                        "/mesh/topology"]
    assert r.read("/mesh/topology").dtype.kind in "iu" and r.read("/mesh/topology").dtype.itemsize == 8 and r.read("/mesh/coordinates").dtype == np.dtype("<f8")
    assert np.array_equal(r.read("/mesh/coordinates"), r.read("/boundaries/coordinates"))
    z = np.load(FIX)
    mesh = fx.Mesh()
    hdf = fx.HDF5File(mesh.mpi_comm(), REF_H5, "r")
    hdf.read(mesh, "/mesh", False)
    boundaries = fx.MeshFunction("size_t", mesh, mesh.topology().dim() - 1)
    hdf.read(boundaries, "/boundaries")
    assert np.array_equal(mesh.coords, z["coords"]) and np.array_equal(mesh.cells, z["cells"])
    assert np.array_equal(boundaries.entries, z["facet_markers"])
    with pytest.raises(KeyError):
        r.read("/nothing")


def test_reader_rejects_other_files(tmp_path):
    from mupopp_b200 import hdf5
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(ValueError):
        hdf5.H5Reader(str(p))
