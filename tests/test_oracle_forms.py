"""Oracle self-consistency: the reference's monolithic LU solve equals the segregated block-triangular sequence
(SURVEY.md finding 5), known answers, quadrature exactness."""
import numpy as np
import pytest

from oracle import fem, forms, modef


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("deg", [1, 2, 3, 4, 6, 8])
def test_quadrature_exact(dim, deg):
    pts, w = fem.simplex_quadrature(dim, deg)
    from math import factorial
    assert abs(w.sum() - 1.0 / factorial(dim)) < 1e-15
    rng = np.random.default_rng(deg)
    for _ in range(5):
        a = rng.integers(0, deg + 1, size=dim)
        if a.sum() > deg:
            continue
        val = (w * np.prod(pts ** a, axis=1)).sum()
        exact = np.prod([factorial(int(k)) for k in a]) / factorial(int(a.sum()) + dim)
        assert abs(val - exact) < 1e-14


@pytest.mark.parametrize("dim", [2, 3])
def test_p2_basis_is_nodal(dim):
    le = fem.TRI_EDGES if dim == 2 else fem.TET_EDGES
    verts = np.concatenate([np.zeros((1, dim)), np.eye(dim)])
    nodes = np.concatenate([verts, 0.5 * (verts[le[:, 0]] + verts[le[:, 1]])])
    phi, _ = fem.tabulate(2, nodes)
    assert np.allclose(phi, np.eye(len(nodes)), atol=1e-14)


def _problem(dim, n, maker):
    if dim == 2:
        m = fem.unit_square_mesh(n, n)
        zh = (0.0, 1.0)
    else:
        m = fem.unit_cube_mesh(n)
        zh = (0.0, 0.0, 1.0)
    asm = forms.Assembler(m)
    prm = maker(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zhat=zh)
    n1, n2 = asm.P1.ndof, asm.P2.ndof
    X2, X1 = asm.P2.dof_coords(), asm.P1.dof_coords()
    top2 = np.where(X2[:, -1] > 1 - 1e-12)[0]
    top1 = np.where(X1[:, -1] > 1 - 1e-12)[0]
    uv = [np.zeros(len(top2)) for _ in range(dim)]
    uv[-1] = -0.5 * np.ones(len(top2))
    bcs = forms.BCs(u_dofs=[top2] * dim, u_vals=uv, c0_dofs=top1, c0_vals=0.1 * np.ones(len(top1)))
    rng = np.random.default_rng(0)
    prev = forms.State(u=[0.1 * rng.standard_normal(n2) for _ in range(dim)], p=np.zeros(n1), c0=0.1 * rng.random(n1),
                       c1=0.1 + 0.01 * rng.random(n1), c2=0.01 * rng.random(n1) if prm.ncomp == 3 else None)
    f = np.abs(np.sin(3 * np.pi * X1[:, 0])) * (1 - np.tanh((1 - X1[:, -1]) / 0.01))
    return asm, prm, prev, f, bcs


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 3)])
@pytest.mark.parametrize("maker", [forms.ccs_three_component_params, forms.darcy_three_component_params,
                                   forms.darcy_two_component_params])
def test_monolithic_equals_segregated(dim, n, maker):
    asm, prm, prev, f, bcs = _problem(dim, n, maker)
    mono, _ = forms.monolithic_step(asm, prm, prev, f, bcs)
    seg = forms.segregated_step(asm, prm, prev, f, bcs)
    for nm in ["p", "c0", "c1"] + (["c2"] if prm.ncomp == 3 else []):
        assert rel(getattr(seg, nm), getattr(mono, nm)) < 1e-11, nm
    for c in range(dim):
        assert rel(seg.u[c], mono.u[c]) < 1e-11
    mono2, _ = forms.monolithic_step(asm, prm, mono, f, bcs)
    seg2 = forms.segregated_step(asm, prm, seg, f, bcs)
    assert rel(seg2.c0, mono2.c0) < 1e-10 and rel(seg2.p, mono2.p) < 1e-10


def test_constant_state_preserved():
    """c0 == const, Da = 0, f = 0, div-free uniform velocity: the CN-SUPG ADR form keeps the constant (SURVEY.md section 4 iii)."""
    m = fem.unit_square_mesh(6, 6)
    asm = forms.Assembler(m)
    n2 = asm.P2.ndof
    vel = ("P2", [0.3 * np.ones(n2), -0.2 * np.ones(n2)])
    c = 0.7 * np.ones(m.nv)
    A, b = forms.adr_single_system(asm, Pe=100.0, Da=0.0, dt=0.1, vel=vel, c_prev=c)
    import scipy.sparse.linalg as spla
    x = spla.splu(A.tocsc()).solve(b)
    assert np.abs(x - 0.7).max() < 1e-12


def test_darcy_bilinear_hydrostatic_known_answer():
    """mupopp.py:595-625 with u = 0 on the whole boundary: the exact solution is u = 0, K grad p = -K f zhat... i.e. p linear.
    With all-natural BCs instead (p=0 on the boundary weakly) the system is solvable; check div-free and symmetry structure."""
    m = fem.unit_square_mesh(6, 6)
    asm = forms.Assembler(m)
    A, b = forms.darcy_bilinear_system(asm, K=0.1)
    n2, n1 = asm.P2.ndof, asm.P1.ndof
    # block structure: M symmetric, (v,p) block = -K * (q,u) block^T
    M = A[:n2, :n2]
    assert abs(M - M.T).max() < 1e-15
    Bx = A[2 * n2:, :n2]
    Gx = A[:n2, 2 * n2:]
    assert abs(Gx + 0.1 * Bx.T).max() < 1e-15
    # rhs = K * int v_y
    assert abs(b[n2:2 * n2].sum() - 0.1) < 1e-14 and abs(b[:n2]).max() == 0.0


def test_modef_mass_balance_and_direct_vs_krylov():
    m = fem.box_mesh((0, 0, 0), (1, 1, 1), 5, 5, 5)
    asm = forms.Assembler(m)
    prm = forms.ccs_three_component_params(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zhat=(0, 0, 1.0))
    top = lambda x: x[2] > 1 - 1e-12
    bcs = modef.make_bcs(m, top, lambda x: np.array([0, 0, -0.5]), top, 0.1)
    assert abs(bcs.flux_load.sum() + 0.5) < 1e-14
    n1 = m.nv
    f = np.zeros(n1)
    st = modef.StateF(u=np.zeros((m.nc, 3)), p=np.zeros(n1), c0=np.zeros(n1), c1=0.1 * np.ones(n1), c2=np.zeros(n1))
    a = modef.step(asm, prm, st, f, bcs)
    b = modef.step(asm, prm, st, f, bcs, solver="krylov", rtol=1e-13)
    assert rel(b.p, a.p) < 1e-10 and rel(b.c0, a.c0) < 1e-10
    # discrete mass balance of the DG0 velocity: sum_cells |K| grad(q_i).u = flux_load_i on non-Dirichlet rows
    g = modef.p1_grads(asm)
    vol = asm.detJ / 6.0
    res = fem.scatter_vector(np.einsum("c,cia,ca->ci", vol, g, a.u), asm.P1.cell_dofs, n1) - bcs.flux_load
    free = np.ones(n1, bool)
    free[bcs.p_dofs] = False
    assert np.abs(res[free]).max() < 1e-12


def test_two_component_adaptive_reduces_to_plain_supg_without_reaction():
    """a6 (mupopp.py:667-749) with Da = 0 and alpha = 0: c1 is frozen and the c0 row is the Sendur-SUPG CN advection-diffusion row."""
    m = fem.rectangle_mesh(0, 0, 2, 1, 6, 5)
    asm = forms.Assembler(m)
    n2 = asm.P2.ndof
    vel = ("P2", [0.4 * np.ones(n2), -0.3 * np.ones(n2)])
    prm = forms.two_component_adaptive_params(100.0, 0.0, 0.01, 0.0, 0.05)
    rng = np.random.default_rng(3)
    c0n, c1n = 0.1 * rng.random(m.nv), 0.2 + 0.1 * rng.random(m.nv)
    c0, c1 = forms.two_component_adaptive_step(asm, prm, vel, c0n, c1n, forms.adaptive_source(m.coords), np.zeros(0, dtype=np.int64), np.zeros(0))
    assert rel(c1, c1n) < 1e-13
    # a constant c0 stays constant
    c0, _ = forms.two_component_adaptive_step(asm, prm, vel, 0.3 * np.ones(m.nv), c1n, forms.adaptive_source(m.coords), np.zeros(0, dtype=np.int64), np.zeros(0))
    assert np.abs(c0 - 0.3).max() < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_stokes_poiseuille_known_answer(dim):
    """a7: the Stokes block of mupopp.py:1044 with parabolic inflow reproduces plane Poiseuille flow exactly (quadratic u is in P3, linear p in P1)."""
    from oracle import stokes as os_
    if dim == 2:
        m = fem.rectangle_mesh(0, 0, 2, 1, 4, 3)
    else:
        m = fem.box_mesh((0, 0, 0), (2, 1, 1), 3, 3, 2)
    asm = forms.Assembler(m)
    X3 = asm.P3.dof_coords()
    n3, n1 = asm.P3.ndof, asm.P1.ndof
    prof = lambda x: x[:, 1] * (1 - x[:, 1])
    if dim == 2:
        D = np.where((X3[:, 0] < 1e-12) | (X3[:, 1] < 1e-12) | (X3[:, 1] > 1 - 1e-12))[0]
    else:   # channel between the planes y = 0, 1; free-slip is not available, so prescribe the exact profile on the z walls too
        D = np.where((X3[:, 0] < 1e-12) | (X3[:, 1] < 1e-12) | (X3[:, 1] > 1 - 1e-12) | (X3[:, 2] < 1e-12) | (X3[:, 2] > 1 - 1e-12))[0]
    g = [np.zeros(n3) for _ in range(dim)]
    g[0][D] = prof(X3[D])
    prm = os_.stokes_adr_params(Pe=1.0, Da=0.0, dt=1.0)
    prev = forms.State(u=[np.zeros(n3)] * dim, p=np.zeros(n1), c0=np.zeros(n1), c1=np.zeros(n1), c2=np.zeros(n1))
    st = os_.monolithic_step(asm, prm, prev, [D] * dim, [g[a][D] for a in range(dim)], np.zeros(0, dtype=np.int64), np.zeros(0))
    assert np.abs(st.u[0] - prof(X3)).max() < 1e-10
    for a in range(1, dim):
        assert np.abs(st.u[a]).max() < 1e-10
    # grad(u):grad(v) + div(v) p form: -lap(u) - grad(p) = 0 with u_x = y(1-y) gives dp/dx = +2, p = 0 at the natural outflow x = 2
    assert np.abs(st.p - 2.0 * (m.coords[:, 0] - 2.0)).max() < 1e-9


def test_c1_and_c2_rows_share_one_increment():
    """The identity behind TransportSolverF's single mass solve: in the three-component forms (mupopp.py:936-937, 1234-1235) the c1 and c2
    rows are M c1' = M c1 - k2 T and M c2' = M c2 + k3 T with the SAME reaction load T, so c2' - c2 = -(frac[2]/frac[1]) (c1' - c1).
    Checked on the oracle's own (LU-solved) step, natural and periodic walls."""
    import scipy.sparse.linalg as spla
    from oracle import fem, forms, modef
    box = ((0.0, 0.0, 0.0), (4.0, 4.0, 1.0))
    m = fem.box_mesh(box[0], box[1], 5, 4, 6)
    asm = forms.Assembler(m)
    prm = forms.ccs_three_component_params(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, zhat=(0.0, 0.0, 1.0))
    rng = np.random.default_rng(2)
    for periodic in (False, True):
        P = None
        nd = m.nv
        if periodic:
            P, dof = modef.periodic_prolongation(m, (0, 1), box[0], box[1])
            nd = P.shape[1]
        st = modef.StateF(u=np.zeros((m.nc, 3)), p=np.zeros(nd), c0=0.05 + 0.03 * rng.random(nd), c1=0.1 + 0.02 * rng.random(nd), c2=0.01 * rng.random(nd))
        S = modef.StepSystems(asm, prm, st, np.zeros(nd), None, P)
        M1, b1 = S.c1()
        M2, b2 = S.c2()
        lu = spla.splu(M1.tocsc())
        d1 = lu.solve(b1) - st.c1
        d2 = spla.splu(M2.tocsc()).solve(b2) - st.c2
        ratio = prm.fracs[2] / prm.fracs[1]
        assert np.linalg.norm(d1) > 1e-8
        assert np.linalg.norm(d2 + ratio * d1) <= 1e-10 * np.linalg.norm(d2)          # (the increments are 1e-4 of the fields: rounding of c + d)
