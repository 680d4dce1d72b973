"""Host mesh generators of the product reproduce the oracle's DOLFIN-convention meshes bit for bit."""
import numpy as np
import pytest

import mupopp_b200 as mp
from mupopp_b200.transport import make_boundary_data
from oracle import fem, modef


@pytest.mark.parametrize("nx,ny", [(3, 2), (8, 8)])
def test_rectangle(nx, ny):
    a = mp.RectangleMesh((0, 0), (4, 1), nx, ny)
    b = fem.rectangle_mesh(0, 0, 4, 1, nx, ny)
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.coords, b.coords)
    b.init_edges()
    assert np.array_equal(a.edges, b.edges) and np.array_equal(a.cell_edges, b.cell_edges)


@pytest.mark.parametrize("n", [(2, 3, 4), (5, 5, 5)])
def test_box(n):
    a = mp.BoxMesh((0, 0, 0), (4, 4, 1), *n)
    b = fem.box_mesh((0, 0, 0), (4, 4, 1), *n)
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.coords, b.coords)
    assert a.num_cells() == 6 * n[0] * n[1] * n[2]
    b.init_edges()
    assert np.array_equal(a.edges, b.edges) and np.array_equal(a.cell_edges, b.cell_edges)
    assert a.num_edges() == 7 * n[0] * n[1] * n[2] + 3 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2]) + sum(n)
    assert np.allclose(a.cell_diameters(), b.hmax_cells())


def test_box_slab_is_a_window_of_the_global_mesh():
    g = mp.BoxMesh((0, 0, 0), (1, 1, 2), 3, 4, 6)
    s = mp.BoxMesh((0, 0, 0), (1, 1, 2), 3, 4, 6, z_range=(2, 5))
    npl = 4 * 5
    cells_g = g.cells[6 * 3 * 4 * 2: 6 * 3 * 4 * 5]
    assert np.array_equal(s.cells + 2 * npl, cells_g)
    assert np.array_equal(s.coords, g.coords[2 * npl: 6 * npl])


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_data_matches_oracle(dim):
    if dim == 2:
        a, b = mp.RectangleMesh((0, 0), (2, 1), 6, 5), fem.rectangle_mesh(0, 0, 2, 1, 6, 5)
    else:
        a, b = mp.BoxMesh((0, 0, 0), (2, 2, 1), 4, 3, 5), fem.box_mesh((0, 0, 0), (2, 2, 1), 4, 3, 5)
    uval = np.zeros(dim)
    uval[-1] = -0.5
    bd = make_boundary_data(a, lambda x: x[..., -1] > 1 - 1e-12, lambda x: uval * (1 + x[..., :1]), lambda x: x[..., -1] > 1 - 1e-12, 0.1)
    ob = modef.make_bcs(b, lambda x: x[-1] > 1 - 1e-12, lambda x: uval * (1 + x[0]), lambda x: x[-1] > 1 - 1e-12, 0.1)
    assert np.array_equal(bd.p_isbc.nonzero()[0], ob.p_dofs)
    assert np.allclose(bd.flux_load, ob.flux_load, rtol=0, atol=1e-15)
    assert np.array_equal(bd.c0_isbc.nonzero()[0], ob.c0_dofs)
    assert np.allclose(bd.c0_g[ob.c0_dofs], ob.c0_vals)
    # generic (non-box) facet detection agrees with the box fast path
    a2 = mp.Mesh(a.coords, a.cells)
    f1 = a.boundary_facets()[0]
    f2 = a2.boundary_facets()[0]
    assert sorted(map(tuple, f1.tolist())) == sorted(map(tuple, f2.tolist()))


@pytest.mark.parametrize("dim", [2, 3])
def test_face_numbering_and_p3_tables_match_oracle(dim):
    from mupopp_b200 import reference_element as ref
    a = mp.RectangleMesh((0, 0), (2, 1), 4, 3) if dim == 2 else mp.BoxMesh((0, 0, 0), (2, 1, 1), 3, 2, 2)
    b = fem.rectangle_mesh(0, 0, 2, 1, 4, 3) if dim == 2 else fem.box_mesh((0, 0, 0), (2, 1, 1), 3, 2, 2)
    if dim == 3:
        b.init_faces()
        assert np.array_equal(a.faces, b.faces) and np.array_equal(a.cell_faces, b.cell_faces)
    for deg in (1, 2, 3):
        for q in (2, 4, 6, 8):
            pts, w = ref.simplex_quadrature(dim, q)
            po, wo = fem.simplex_quadrature(dim, q)
            assert np.allclose(pts, po, rtol=0, atol=1e-15) and np.allclose(w, wo, rtol=0, atol=1e-15)
            t1, d1 = ref.tabulate(deg, pts)
            t2, d2 = fem.tabulate(deg, pts)
            assert np.allclose(t1, t2, rtol=0, atol=1e-15) and np.allclose(d1, d2, rtol=0, atol=1e-14)


def test_lattice_line_stride_of_structured_meshes_and_slabs():
    """Rows per lattice plane (the line preconditioner's stride): BoxMesh / RectangleMesh, periodic x/y dofs, z-slab partitions (owned
    planes first), and 0 for meshes without lattice structure (they fall back to point Jacobi)."""
    import mupopp_b200 as mp
    from mupopp_b200 import partition
    from mupopp_b200.mesh import Mesh, box_periodic_map
    from mupopp_b200.transport import lattice_line_stride
    box = ((0.0, 0.0, 0.0), (4.0, 4.0, 1.0))
    m = mp.BoxMesh(box[0], box[1], 5, 4, 6)
    assert lattice_line_stride(m) == 6 * 5
    pm = box_periodic_map(m, (0, 1))
    assert lattice_line_stride(m, pm.vert2dof) == 5 * 4 and pm.ndof == 5 * 4 * 7
    assert lattice_line_stride(mp.RectangleMesh((0, 0), (2, 1), 7, 3)) == 8
    for per in (False, True):
        for rank in range(3):
            part = partition.slab_partition(box, (5, 4, 6), rank, 3, periodic_xy=per)
            st = lattice_line_stride(part.mesh, part.vert2dof, part.n_owned)
            assert st == (20 if per else 30) and part.n_owned % st == 0
    assert lattice_line_stride(Mesh(m.coords, m.cells)) == 0            # no lattice information: point Jacobi
