"""GPU parity of the compaction forms against the oracle (assemble_system + LU)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import compaction as oc
from oracle import fem, forms

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _setup(ctx, dim):
    import mupopp_b200 as mp
    from mupopp_b200.moder import Spaces
    if dim == 3:
        pm, om = mp.BoxMesh((-1, -1, 0), (1, 1, 2), 3, 3, 4), fem.box_mesh((-1, -1, 0), (1, 1, 2), 3, 3, 4)
        zh = (0.0, 0.0, 1.0)
    else:
        pm, om = mp.RectangleMesh((-1, 0), (1, 2), 5, 6), fem.rectangle_mesh(-1, 0, 1, 2, 5, 6)
        zh = (0.0, 1.0)
    asm = forms.Assembler(om)
    sp = Spaces(ctx, pm)
    X = om.coords
    r2 = (X[:, :-1] ** 2).sum(1) + (X[:, -1] - 0.5) ** 2
    phi = 0.05 * np.exp(-r2) / 0.5 + 0.1                       # time_test1.py:155
    gam = 0.02 * np.sin(X[:, 0]) * X[:, -1]
    buoy = 1.0 + 0.1 * np.cos(X[:, -1])
    return pm, om, asm, sp, zh, phi, gam, buoy


@pytest.mark.parametrize("dim", [2, 3])
def test_momentum_blocks_and_solve(ctx, dim):
    from mupopp_b200.compaction import BlockVector, CompactionMomentumR
    pm, om, asm, sp, zh, phi, gam, buoy = _setup(ctx, dim)
    prm = oc.CompactionParams(da=1e-2, R=0.2, B=1e3, theta=10.0, dL=1.0, zhat=zh + (0.0,) * (3 - dim))
    n1, n2 = asm.P1.ndof, asm.P2.ndof
    X2 = asm.P2.dof_coords()
    side = np.where(np.abs(X2[:, 0]) > 1 - 1e-12)[0] if dim == 2 else np.where((np.abs(X2[:, 0]) > 1 - 1e-12) | (np.abs(X2[:, 1]) > 1 - 1e-12))[0]
    isbc = np.zeros(n2, dtype=np.uint8)
    isbc[side] = 1
    g = [np.zeros(n2) for _ in range(dim)]
    g[-1][side] = 0.05 * np.sin(X2[side, -1])                # a non-homogeneous wall velocity exercises the lifting
    s = CompactionMomentumR(ctx, sp, prm.da, prm.R, prm.B, prm.theta, prm.dL, zh, [isbc] * dim, g)
    phid, gamd, buod = ctx.to_device(phi), ctx.to_device(gam), ctx.to_device(buoy)
    s.assemble(phid, gamd, buod)
    blk = oc.momentum_blocks(asm, prm, phi, gam, buoy)
    for a in range(dim):
        for b in range(dim):
            assert abs(s.Auu[a][b].to_scipy() - blk["Auu"][a][b]).max() <= 1e-12 * abs(blk["Auu"][a][a]).max()
        assert rel(s.rhs.u[a].cpu().numpy(), blk["bu"][a]) < 1e-11
    assert rel(s.rhs.p.cpu().numpy(), blk["bp"]) < 1e-11
    assert abs(s.Acc.to_scipy() - blk["Acc"]).max() <= 1e-12 * abs(blk["Acc"]).max()
    assert abs(s.App.to_scipy() - blk["App"]).max() <= 1e-10 * max(abs(blk["App"]).max(), 1e-300) + 1e-300
    assert abs(s.Puu.to_scipy() - blk["Puu"]).max() <= 1e-12 * abs(blk["Puu"]).max()
    assert abs(s.Ppp.to_scipy() - blk["Ppp"]).max() <= 1e-11 * abs(blk["Ppp"]).max()
    assert abs(s.Pcc.to_scipy() - blk["Pcc"]).max() <= 1e-12 * abs(blk["Pcc"]).max()
    # operator == assemble_system(a, L, bcs) of the oracle, applied to a random vector
    A, bvec, _ = oc.momentum_system(asm, prm, phi, gam, buoy)
    dofs = np.concatenate([a * n2 + side for a in range(dim)])
    vals = np.concatenate([g[a][side] for a in range(dim)])
    Abc, bbc = fem.apply_dirichlet_symmetric(A, bvec, dofs, vals)
    rng = np.random.default_rng(0)
    xr = rng.standard_normal(A.shape[0])
    x = BlockVector(ctx, dim, n2, n1)
    y = BlockVector(ctx, dim, n2, n1)
    x.buf.copy_(ctx.to_device(xr))
    s.apply(x, y)
    assert rel(y.buf.cpu().numpy(), Abc @ xr) < 1e-12
    yb = ctx.zeros(y.buf.numel())
    s._block_system().apply(x.buf, yb)                       # the same operator as an mp_block_op (what mp_solve_minres iterates on)
    assert rel(yb.cpu().numpy(), Abc @ xr) < 1e-12
    # MINRES vs LU.  The reference runs MINRES to rtol 1e-6 (mupopp.py:410-413); run tighter to expose the discretisation parity
    uo, po, co = oc.momentum_solve(asm, prm, phi, gam, buoy, [side] * dim, [g[a][side] for a in range(dim)])
    x.buf.zero_()
    its = s.solve(x, rtol=1e-12, maxit=20000)
    xo = np.concatenate(list(uo) + [po, co])
    res_main = s.last_result
    x1 = BlockVector(ctx, dim, n2, n1)
    its_jacobi = s.solve(x1, rtol=1e-12, maxit=20000, cheb=(1, 1, 1), precond="form_b")     # the round-1 preconditioner: diagonal of the form b
    x2 = BlockVector(ctx, dim, n2, n1)
    its_formb = s.solve(x2, rtol=1e-12, maxit=20000, precond="form_b")                     # Chebyshev on the blocks of the form b
    print("compaction MINRES dim %d: mass-Schur blocks + Chebyshev %d iterations (relres %.2e, err %.2e); form b + Chebyshev %d (err %.2e); "
          "diagonal of form b %d (err %.2e)" % (dim, its, res_main.relres, rel(x.buf.cpu().numpy(), xo), its_formb, rel(x2.buf.cpu().numpy(), xo),
                                               its_jacobi, rel(x1.buf.cpu().numpy(), xo)))
    assert res_main.converged
    assert its < its_jacobi
    assert rel(x.buf.cpu().numpy(), xo) < 1e-8
    for a in range(dim):
        assert rel(x.u[a].cpu().numpy(), uo[a]) < 1e-8 or np.linalg.norm(uo[a]) < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_porosity_row(ctx, dim):
    from mupopp_b200.compaction import PorosityR
    pm, om, asm, sp, zh, phi, gam, buoy = _setup(ctx, dim)
    n2 = asm.P2.ndof
    X2 = asm.P2.dof_coords()
    for scale in (0.05, 60.0):                                 # 60: keff > 0, the SUPG branch of mupopp.py:248-255 is active
        u = [scale * (0.3 + 0.2 * np.sin(X2[:, 0] + a) * np.cos(X2[:, -1])) for a in range(dim)]
        Ao, bo = oc.porosity_system(asm, 1e-2, 0.05, phi, u, gam)
        ph1o = spla.splu(Ao.tocsc()).solve(bo)
        s = PorosityR(ctx, sp, 1e-2)
        ud = [ctx.to_device(v) for v in u]
        ph1 = ctx.zeros(om.nv)
        r = s.solve(ctx.to_device(phi), ud, 0.05, ctx.to_device(gam), ph1, rtol=1e-14, maxit=2000)
        assert r.converged
        assert abs(s.A.to_scipy() - Ao).max() <= 1e-11 * abs(Ao).max()
        assert rel(ph1.cpu().numpy(), ph1o) < 1e-10


@pytest.mark.parametrize("dim", [2, 3])
def test_porosity_row_p2_space(ctx, dim):
    """X = CG2 (benchmark/soliton/soliton.py:120, benchmark/sphere_3D_pure_shear/sphere.py:76): P2 porosity / melting rate, with the
    -div(grad(phi_mid)) term of the SUPG residual (mupopp.py:250-251)."""
    from mupopp_b200.compaction import PorosityR
    pm, om, asm, sp, zh, phi, gam, buoy = _setup(ctx, dim)
    X2 = asm.P2.dof_coords()
    n2 = asm.P2.ndof
    rng = np.random.default_rng(3)
    phi2 = 0.1 + 0.05 * np.exp(-((X2 - X2.mean(0)) ** 2).sum(1)) + 0.01 * rng.random(n2)
    gam2 = 0.2 * np.sin(X2[:, 0]) * np.cos(X2[:, -1])
    # scale 60: keff > 0, the SUPG branch (with the Laplacian term) is active.  That matrix is indefinite (negative diagonal entries from
    # the -div(grad) term at |u| h >> 1), so Jacobi-GMRES needs a long recurrence: restart 200, and dt small enough for it to converge
    for scale, dt in ((0.05, 0.05), (60.0, 0.005)):
        u = [scale * (0.3 + 0.2 * np.sin(X2[:, 0] + a) * np.cos(X2[:, -1])) for a in range(dim)]
        Ao, bo = oc.porosity_system(asm, 1e-2, dt, phi2, u, gam2, pdeg=2)
        ph1o = spla.splu(Ao.tocsc()).solve(bo)
        s = PorosityR(ctx, sp, 1e-2, degree=2)
        ph1 = ctx.zeros(n2)
        r = s.solve(ctx.to_device(phi2), [ctx.to_device(v) for v in u], dt, ctx.to_device(gam2), ph1, rtol=1e-13, maxit=4000, restart=200)
        assert abs(s.A.to_scipy() - Ao).max() <= 1e-11 * abs(Ao).max()
        assert rel(s.rhs.cpu().numpy(), bo) < 1e-12
        print("P2 porosity dim %d scale %g: %r, err %.2e" % (dim, scale, r, rel(ph1.cpu().numpy(), ph1o)))
        assert r.converged or r.relres < 1e-11, r
        assert rel(ph1.cpu().numpy(), ph1o) < 1e-9
