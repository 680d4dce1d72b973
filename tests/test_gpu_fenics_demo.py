"""The reference's own data through the GPU path: unstructured mesh, facet markers and P2 velocity field of
/root/reference/src/run/fenics_demo (tests/golden/fenics_demo_dofmap.npz), SUPG advection-diffusion of
DarcyAdvection.advection_diffusion (mupopp.py:627-664; the demo it was copied from: demo_advection-diffusion.py:61-79)."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem, forms

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "fenics_demo_dofmap.npz")


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_fenics_demo_fixture_advection_diffusion(ctx):
    from mupopp_b200.fenics import Constant, DirichletBC, Function, FunctionSpace, Mesh, MeshFunction, VectorFunctionSpace, solve
    from mupopp_b200.mupopp import DarcyAdvection
    g = np.load(G)
    mesh = Mesh(g["coords"], g["cells"])
    sub_domains = MeshFunction("size_t", mesh, g["facet_markers"])
    Q = FunctionSpace(mesh, "CG", 1)
    V = VectorFunctionSpace(mesh, "CG", 2)
    velocity = Function(V)
    velocity.vector().set_local(g["dof_value"])                   # UFC numbering, pinned by tests/test_oracle_golden.py
    bc = DirichletBC(Q, Constant(1.0), sub_domains, 1)            # demo_advection-diffusion.py:79
    model = DarcyAdvection(Pe=1.0 / 0.00005, Da=0.0)              # c = 5e-5, no reaction
    u0, u = Function(Q), Function(Q)
    # oracle
    om = fem.Mesh(g["coords"], g["cells"].astype(np.int64))
    asm = forms.Assembler(om)
    n2 = asm.P2.ndof
    vel = [g["dof_value"][:n2], g["dof_value"][n2:]]
    fm = g["facet_markers"]
    sel = fm[fm[:, 2] == 1]
    bverts = np.unique([om.cells[c][j] for c, lf, _ in sel for j in range(3) if j != lf])
    assert len(sel) == 20 and len(bverts) == 21
    co = np.zeros(om.nv)
    for k in range(3):
        a, L = model.advection_diffusion(Q, u0, velocity, 0.1, mesh)
        its = solve(a == L, u, bc)
        u0 = u
        Ao, bo = forms.adr_single_system(asm, 20000.0, 0.0, 0.1, ("P2", vel), co)
        Ab, bb = fem.apply_dirichlet_symmetric(Ao, bo, bverts, np.ones(len(bverts)))
        co = spla.splu(Ab.tocsc()).solve(bb)
        assert rel(u.get(), co) < 1e-8, (k, its)
    assert co.max() > 0.5 and np.isfinite(co).all()
