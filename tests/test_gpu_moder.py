"""GPU parity of MODE R (the reference's mixed P2/P1 discretisation) against the oracle's monolithic LU solve."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem, forms

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("dim,n", [(2, 6), (3, 3)])
def test_reftensor_blocks_match_oracle(ctx, dim, n):
    import mupopp_b200 as mp
    from mupopp_b200.moder import Spaces
    if dim == 2:
        pm, om = mp.RectangleMesh((0, 0), (2, 1), n, n + 1), fem.rectangle_mesh(0, 0, 2, 1, n, n + 1)
    else:
        pm, om = mp.BoxMesh((0, 0, 0), (2, 1, 1), n, n + 1, n), fem.box_mesh((0, 0, 0), (2, 1, 1), n, n + 1, n)
    asm = forms.Assembler(om)
    sp = Spaces(ctx, pm)
    assert np.array_equal(sp.cd2_host, asm.P2.cell_dofs)
    M = sp.assemble("mass", 2, 2, scale=0.3).to_scipy()
    Mo = asm.mass(2, 2, 4, 0.3)
    assert abs(M - Mo).max() <= 1e-14 * abs(Mo).max()
    assert np.array_equal(M.indptr, Mo.indptr) and np.array_equal(M.indices, Mo.indices)      # P2 CSR pattern bit-exact
    Bo = asm.div_blocks(1, 2, 2)
    rng = np.random.default_rng(2)
    K = 0.5 + rng.random(om.nv)
    Kd = ctx.to_device(K)
    BKo = asm.div_blocks(1, 2, 3, K)
    for c in range(dim):
        B = sp.assemble("grad", 1, 2, direction=c).to_scipy()
        assert abs(B - Bo[c]).max() <= 1e-13 * abs(Bo[c]).max()
        BT = sp.assemble("gradT", 2, 1, direction=c, scale=0.7).to_scipy()
        assert abs(BT - 0.7 * Bo[c].T).max() <= 1e-13 * abs(Bo[c]).max()
        BTK = sp.assemble("gradT", 2, 1, direction=c, coef=Kd).to_scipy()
        assert abs(BTK - BKo[c].T).max() <= 1e-13 * abs(BKo[c]).max()
    G = sp.assemble("mass", 2, 1, coef=Kd).to_scipy()
    Go = asm.mass(2, 1, 4, K)
    assert abs(G - Go).max() <= 1e-13 * abs(Go).max()
    S = sp.assemble("stiff", 2, 2).to_scipy()
    So = asm.stiffness(2, 4)
    assert abs(S - So).max() <= 1e-12 * abs(So).max()


def _config1(nx, ny):
    """run/darcy_advection_diffusion/darcy_advection_diffusion.py on UnitSquareMesh (BASELINE configs[0])."""
    import mupopp_b200 as mp
    pm, om = mp.UnitSquareMesh(nx, ny), fem.unit_square_mesh(nx, ny)
    asm = forms.Assembler(om)
    X2 = asm.P2.dof_coords()
    bot = np.where(X2[:, 1] < 1e-14)[0]
    top = np.where(X2[:, 1] > 1 - 1e-14)[0]
    n2 = asm.P2.ndof
    isbc = np.zeros(n2, dtype=np.uint8)
    isbc[bot] = 1
    isbc[top] = 1
    gx = np.zeros(n2)
    gy = np.zeros(n2)
    gy[bot] = np.sin(2 * np.pi * X2[bot, 0])       # u = -sin(2 pi x) * n, n = (0,-1)   (script lines 25-36, 69)
    gy[top] = 1.0                                   # (0, 1) on top                      (line 70)
    return pm, om, asm, isbc, gx, gy, bot, top


def test_config1_darcy_then_adr(ctx):
    import torch
    from mupopp_b200.moder import AdrSolverR, MixedDarcyR, Spaces, adr_single_params
    nx = ny = 12
    pm, om, asm, isbc, gx, gy, bot, top = _config1(nx, ny)
    n1, n2 = asm.P1.ndof, asm.P2.ndof
    # ---- oracle: darcy_bilinear + LU (script lines 79-81)
    A, b = forms.darcy_bilinear_system(asm, K=0.1, zhat=(0.0, 1.0))
    D = np.nonzero(isbc)[0]
    dofs = np.concatenate([D, n2 + D])
    vals = np.concatenate([gx[D], gy[D]])
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, dofs, vals)
    x = spla.splu(Abc.tocsc()).solve(bbc)
    uo = [x[:n2], x[n2:2 * n2]]
    po = x[2 * n2:]
    # ---- GPU: Schur-PCG
    sp = Spaces(ctx, pm)
    darcy = MixedDarcyR(ctx, sp, K=0.1, phi=1.0, u_isbc=[isbc, isbc], u_g=[gx, gy])
    ones = torch.ones(n1, dtype=torch.float64, device=ctx.device)
    lK = ctx.zeros(n2)
    darcy.G.spmv(ones, lK)                       # L = K f v.zhat, f == 1
    rhs = [ctx.zeros(n2), lK.clone()]
    p = ctx.zeros(n1)
    u, its = darcy.solve(rhs, p)
    assert rel(p.cpu().numpy(), po) < 1e-10
    assert rel(u[0].cpu().numpy(), uo[0]) < 1e-10 and rel(u[1].cpu().numpy(), uo[1]) < 1e-10
    # ---- 10 ADR steps (script lines 95-117): c = 1 bottom, 0 top, c(0) = 0.1, dt = 0.1, Pe = 100, Da = 10
    X1 = om.coords
    cb = np.where(X1[:, 1] < 1e-14)[0]
    ct = np.where(X1[:, 1] > 1 - 1e-14)[0]
    cis = np.zeros(n1, dtype=np.uint8)
    cis[cb] = 1
    cis[ct] = 1
    cg = np.zeros(n1)
    cg[cb] = 1.0
    adr = AdrSolverR(ctx, sp, adr_single_params(100.0, 10.0, 0.1), cis, cg, vel_degree=2, rtol=1e-14)
    c_o = 0.1 * np.ones(n1)
    c_g = ctx.to_device(c_o.copy())
    cdofs = np.concatenate([ct, cb])
    cvals = np.concatenate([np.zeros(len(ct)), np.ones(len(cb))])
    for step in range(10):
        Ao, bo = forms.adr_single_system(asm, 100.0, 10.0, 0.1, ("P2", uo), c_o)
        Ab, bb = fem.apply_dirichlet_symmetric(Ao, bo, cdofs, cvals)
        c_new_o = spla.splu(Ab.tocsc()).solve(bb)
        # per-solve parity from the oracle's previous state and velocity
        c_in = ctx.to_device(c_o.copy())
        c_out = c_in.clone()
        uo_d = [ctx.to_device(uo[0]), ctx.to_device(uo[1])]
        r = adr.solve(uo_d, c_in, c_out)
        assert r.converged
        assert abs(adr.A.to_scipy() - Ab).max() <= 1e-12 * abs(Ab).max()
        assert rel(c_out.cpu().numpy(), c_new_o) < 1e-10
        # free-running GPU chain with the GPU velocity
        c_next = c_g.clone()
        adr.solve(u, c_g, c_next)
        c_g = c_next
        c_o = c_new_o
    assert rel(c_g.cpu().numpy(), c_o) < 1e-8


@pytest.mark.parametrize("dim,n,maker", [(2, 6, "ccs3"), (3, 3, "ccs3"), (2, 5, "darcy2")])
def test_moder_monolithic_step(ctx, dim, n, maker):
    """TransportSolverR (segregated, GPU) == the reference's monolithic [u,p,c0,c1(,c2)] LU solve (oracle)."""
    import mupopp_b200 as mp
    from mupopp_b200.moder import TransportSolverR
    from mupopp_b200.transport import TransportModel
    if dim == 2:
        pm, om = mp.UnitSquareMesh(n, n), fem.unit_square_mesh(n, n)
        zh = (0.0, 1.0)
    else:
        pm, om = mp.UnitCubeMesh(n, n, n), fem.unit_cube_mesh(n)
        zh = (0.0, 0.0, 1.0)
    asm = forms.Assembler(om)
    if maker == "ccs3":
        model = TransportModel.ccs_three_component(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zh=zh)
        oprm = forms.ccs_three_component_params(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zhat=zh)
    else:
        model = TransportModel.darcy_two_component(Pe=200, Da=1.0, phi=0.2, alpha=0.01, dt=0.01, K=1.0, zh=zh, gam_rho=-1.0, rho_0=0.0)
        oprm = forms.darcy_two_component_params(Pe=200, Da=1.0, phi=0.2, alpha=0.01, dt=0.01, K=1.0, zhat=zh, gam_rho=-1.0, rho_0=0.0)
    n1, n2 = asm.P1.ndof, asm.P2.ndof
    X2, X1 = asm.P2.dof_coords(), asm.P1.dof_coords()
    top2 = np.where(X2[:, -1] > 1 - 1e-12)[0]
    top1 = np.where(X1[:, -1] > 1 - 1e-12)[0]
    uv = [np.zeros(len(top2)) for _ in range(dim)]
    uv[-1] = -0.5 * np.ones(len(top2))
    bcs = forms.BCs(u_dofs=[top2] * dim, u_vals=uv, c0_dofs=top1, c0_vals=0.1 * np.ones(len(top1)))
    isbc2 = np.zeros(n2, dtype=np.uint8)
    isbc2[top2] = 1
    g2 = [np.zeros(n2) for _ in range(dim)]
    g2[-1][top2] = -0.5
    isbc1 = np.zeros(n1, dtype=np.uint8)
    isbc1[top1] = 1
    g1 = np.zeros(n1)
    g1[top1] = 0.1
    f = np.abs(np.sin(3 * np.pi * X1[:, 0])) * (1 - np.tanh((1 - X1[:, -1]) / 0.1))
    rng = np.random.default_rng(0)
    prev = forms.State(u=[np.zeros(n2) for _ in range(dim)], p=np.zeros(n1), c0=0.05 * rng.random(n1),
                       c1=0.1 + 0.01 * rng.random(n1), c2=0.01 * rng.random(n1) if oprm.ncomp == 3 else None)
    s = TransportSolverR(ctx, pm, model, [isbc2] * dim, g2, isbc1, g1, f_source=f, rtol=1e-14)
    for k in range(3):
        mono, _ = forms.monolithic_step(asm, oprm, prev, f, bcs)
        s.set_state(u=prev.u, p=prev.p, c0=prev.c0, c1=prev.c1, c2=prev.c2)
        s.step()
        assert rel(s.c1.cpu().numpy(), mono.c1) < 1e-10
        if oprm.ncomp == 3:
            assert rel(s.c2.cpu().numpy(), mono.c2) < 1e-10
        assert rel(s.c0.cpu().numpy(), mono.c0) < 1e-10
        assert rel(s.p.cpu().numpy(), mono.p) < 1e-9
        for c in range(dim):
            assert np.linalg.norm(s.u[c].cpu().numpy() - mono.u[c]) < 1e-9 * max(np.linalg.norm(np.concatenate(mono.u)), 1e-300)
        prev = mono


@pytest.mark.parametrize("dim,n,deg,nodal", [(3, 2, 2, True), (3, 3, 2, False), (3, 2, 1, True), (2, 9, 1, False), (2, 7, 2, True), (3, 2, 3, False)])
def test_reftensor_tiles_against_numpy(ctx, dim, n, deg, nodal):
    """The tiled reference-tensor kernel (T0 staged in shared memory, 64-cell tiles, 4 x 4 register tiles) == a NumPy contraction of the same
    tensors, cell-ordered and row-ordered output, including tiles that are not full and a shared-memory layout with an odd tensor size."""
    import ctypes as C
    import mupopp_b200 as mp
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.moder import Spaces
    mesh = mp.BoxMesh((0, 0, 0), (1.0, 2.0, 3.0), n, n, 2 * n) if dim == 3 else mp.RectangleMesh((0, 0), (2.0, 1.0), n, n + 2)
    sp = Spaces(ctx, mesh)
    nb = sp.nb(deg)
    L = nb * nb
    T0 = sp.tensor("stiff", deg, deg, nodal)
    coefh = 0.9 + 0.05 * np.random.default_rng(0).random(sp.n1)
    coef = ctx.to_device(coefh)
    nc = mesh.num_cells()
    X = mesh.coords[mesh.cells]
    J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))
    Ji = np.linalg.inv(J)
    geo = (np.einsum("cai,cbi->cab", Ji, Ji) * np.abs(np.linalg.det(J))[:, None, None]).reshape(nc, dim * dim)
    T0h = T0.cpu().numpy()
    if nodal:
        W = (geo[:, :, None] * coefh[mesh.cells][:, None, :]).reshape(nc, dim * dim * (dim + 1))
        ref = 0.5 * (W @ T0h.reshape(W.shape[1], L))
    else:
        ref = 0.5 * (geo @ T0h.reshape(dim * dim, L))
    emat = ctx.zeros(nc * L)
    check(ctx.lib.mp_elem_reftensor(ctx.h, C.byref(sp.dm.c), 3, 0, nb, nb, ptr(T0), ptr(coef) if nodal else None, 1 if nodal else 0, 0.5, ptr(emat)))
    assert np.abs(emat.cpu().numpy().reshape(nc, L) - ref).max() <= 1e-13 * np.abs(ref).max()
    plan = sp.plan(deg, deg)
    A, B = plan.new_matrix(sell=False), plan.new_matrix(sell=False)
    plan.scatter_matrix(emat, A, ordered=False)
    plan.reftensor_scatter(sp.dm, 3, 0, nb, nb, T0, coef if nodal else None, 1 if nodal else 0, 0.5, emat, B)
    assert np.array_equal(A.vals.cpu().numpy(), B.vals.cpu().numpy())
