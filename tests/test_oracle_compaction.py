"""Oracle checks for the compaction forms: symmetry, the analytical known answer u = E.r, p = 0, c = 0
(/root/reference/src/benchmark/sphere_3D_pure_shear/sphere.py:13-22), the literal dL*phi*phi**m3 quirk, porosity row consistency."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle import compaction as oc
from oracle import fem, forms


def _mesh():
    m = fem.box_mesh((-1, -1, 0), (1, 1, 2), 3, 3, 4)
    return m, forms.Assembler(m)


def test_momentum_matrix_is_symmetric_and_pp_block_underflows():
    m, asm = _mesh()
    X = m.coords
    phi = 0.05 * np.exp(-X[:, 0] ** 2 - X[:, 1] ** 2 - (X[:, 2] - 0.5) ** 2) / 0.5 + 0.1
    A, b, blk = oc.momentum_system(asm, oc.CompactionParams(), phi, np.zeros(m.nv), np.ones(m.nv))
    assert abs(A - A.T).max() < 1e-15
    # mupopp.py:365: dL*phi*phi**m3 = dL*phi*phi^(1/phi^3) ~ 1e-130 for phi ~ 0.1, while the preconditioner uses dL/phi
    assert abs(blk["App"]).max() < 1e-100
    assert abs(blk["Ppp"]).max() > 1e-3


def test_pure_shear_known_answer():
    """Constant phi, Gamma = 0, R = 0: u = E.r (trace-free E, exactly representable in P2), p = c = 0 satisfies every interior row."""
    m, asm = _mesh()
    A, b, _ = oc.momentum_system(asm, oc.CompactionParams(R=0.0), 0.1 * np.ones(m.nv), np.zeros(m.nv), np.ones(m.nv))
    X2 = asm.P2.dof_coords()
    E = np.array([[1.0, 0.3, 0.2], [0.3, -0.4, 0.1], [0.2, 0.1, -0.6]])
    xe = np.concatenate([X2 @ E[a] for a in range(3)] + [np.zeros(m.nv), np.zeros(m.nv)])
    r = A @ xe - b
    n2 = asm.P2.ndof
    bnd = np.where((np.abs(X2[:, 0]) > 1 - 1e-12) | (np.abs(X2[:, 1]) > 1 - 1e-12) | (X2[:, 2] < 1e-12) | (X2[:, 2] > 2 - 1e-12))[0]
    mask = np.ones(len(xe), bool)
    for a in range(3):
        mask[a * n2 + bnd] = False
    assert np.abs(b).max() == 0.0
    assert np.abs(r[mask]).max() < 1e-13


def test_momentum_solve_and_porosity_step():
    m, asm = _mesh()
    X = m.coords
    phi = 0.05 * np.exp(-X[:, 0] ** 2 - X[:, 1] ** 2 - (X[:, 2] - 0.5) ** 2) / 0.5 + 0.1
    X2 = asm.P2.dof_coords()
    side = np.where((np.abs(X2[:, 0]) > 1 - 1e-12) | (np.abs(X2[:, 1]) > 1 - 1e-12))[0]
    u, p, c = oc.momentum_solve(asm, oc.CompactionParams(), phi, np.zeros(m.nv), np.ones(m.nv), [side] * 3, [np.zeros(len(side))] * 3)
    assert np.abs(u[2]).max() > 1e-3 and np.abs(c).max() < 1e-10       # Gamma = 0 => div u = 0 => c = 0
    A, b = oc.porosity_system(asm, 1e-4, 0.1, phi, u, np.zeros(m.nv))
    ph1 = spla.splu(A.tocsc()).solve(b)
    assert 0 < np.abs(ph1 - phi).max() < 0.05
    # u = 0, Gamma = 0: the porosity is unchanged
    A, b = oc.porosity_system(asm, 1e-4, 0.1, phi, [np.zeros(asm.P2.ndof)] * 3, np.zeros(m.nv))
    assert np.abs(spla.splu(A.tocsc()).solve(b) - phi).max() < 1e-13


def test_p2_basis_laplacians_known_answer():
    """The cellwise Laplacians of the P2 basis reproduce the Laplacian of any quadratic exactly: f = x^2 + 2 y^2 (+ 3 z^2) + x y -> 6 (12)."""
    from oracle import compaction as oc
    from oracle import fem, forms
    for om, want in ((fem.rectangle_mesh(0, 0, 2, 1, 3, 2), 6.0), (fem.box_mesh((0, 0, 0), (2, 1, 1), 2, 2, 2), 12.0)):
        asm = forms.Assembler(om)
        X = asm.P2.dof_coords()
        f = X[:, 0] ** 2 + 2 * X[:, 1] ** 2 + X[:, 0] * X[:, 1] + (3 * X[:, 2] ** 2 if om.dim == 3 else 0.0)
        lap = np.einsum("cb,cb->c", oc.p2_basis_laplacians(asm), f[asm.P2.cell_dofs])
        assert np.allclose(lap, want, rtol=1e-12, atol=1e-12)
