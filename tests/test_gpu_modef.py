"""GPU parity (through the C-ABI) of the mode-F hot path against the NumPy/SciPy oracle."""
import numpy as np
import pytest
import torch

from oracle import fem, forms, modef

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _mesh_pair(dim, n):
    import mupopp_b200 as mp
    if dim == 2:
        return mp.RectangleMesh((0, 0), (2, 1), n, n), fem.rectangle_mesh(0, 0, 2, 1, n, n)
    return mp.BoxMesh((0, 0, 0), (2, 2, 1), n, n, n), fem.box_mesh((0, 0, 0), (2, 2, 1), n, n, n)


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 6)])
def test_mesh_and_pattern_bit_exact(ctx, dim, n):
    from mupopp_b200.device import DeviceMesh, Plan
    pm, om = _mesh_pair(dim, n)
    assert np.array_equal(pm.cells, om.cells)
    assert np.array_equal(pm.coords, om.coords)
    dm = DeviceMesh(ctx, pm)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    rowptr, colidx = plan.csr_pattern()
    rp, ci = fem.p1_csr_pattern(om)
    assert np.array_equal(rowptr.cpu().numpy(), rp)
    assert np.array_equal(colidx.cpu().numpy(), ci)


@pytest.mark.parametrize("dim,n", [(2, 10), (3, 5)])
def test_constant_operators(ctx, dim, n):
    import ctypes as C
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    pm, om = _mesh_pair(dim, n)
    asm = forms.Assembler(om)
    dm = DeviceMesh(ctx, pm)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    l = dim + 1
    emat = ctx.empty(dm.ncell * l * l)
    M = plan.new_matrix()
    check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(dm.c), 1.0, ptr(emat)))
    plan.scatter_matrix(emat, M)
    Mo = asm.mass(1, 1, 2)
    assert abs(M.to_scipy() - Mo).max() <= 1e-15 * abs(Mo).max() * 10
    rng = np.random.default_rng(1)
    K = 0.5 + rng.random(om.nv)
    Kd = ctx.to_device(K)
    L = plan.new_matrix()
    check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), ptr(Kd), 0.0, 2.0, ptr(emat)))
    plan.scatter_matrix(emat, L)
    # oracle stiffness with the cellwise (mean nodal) coefficient
    g = modef.p1_grads(asm)
    vol = asm.detJ / (2.0 if dim == 2 else 6.0)
    Ae = np.einsum("c,cia,cja->cij", 2.0 * vol * modef.cell_K(asm, K), g, g)
    Lo = fem.scatter_matrix(Ae, asm.P1.cell_dofs, asm.P1.cell_dofs, (om.nv, om.nv))
    assert abs(L.to_scipy() - Lo).max() <= 1e-13 * abs(Lo).max()
    # SpMV + dot + axpby
    x = rng.standard_normal(om.nv)
    xd = ctx.to_device(x)
    yd = ctx.empty(om.nv)
    L.spmv(xd, yd)
    assert rel(yd.cpu().numpy(), Lo @ x) < 1e-14
    assert abs(ctx.dot(xd, yd) - x @ (Lo @ x)) <= 1e-12 * abs(x @ (Lo @ x))
    ctx.axpby(2.0, xd, -1.0, yd)
    assert rel(yd.cpu().numpy(), 2 * x - Lo @ x) < 1e-14


def _setup(dim, n, kind):
    import mupopp_b200 as mp
    from mupopp_b200.transport import TransportModel, make_boundary_data
    pm, om = _mesh_pair(dim, n)
    zh = (0.0, 1.0) if dim == 2 else (0.0, 0.0, 1.0)
    top = 1.0
    if kind == "ccs3":
        model = TransportModel.ccs_three_component(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zh=zh)
        oprm = forms.ccs_three_component_params(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zhat=zh)
    elif kind == "darcy2":
        model = TransportModel.darcy_two_component(Pe=1e3, Da=10.0, phi=0.01, alpha=0.01, dt=0.1, K=0.1, zh=zh, gam_rho=-1.0, rho_0=0.0)
        oprm = forms.darcy_two_component_params(Pe=1e3, Da=10.0, phi=0.01, alpha=0.01, dt=0.1, K=0.1, zhat=zh, gam_rho=-1.0, rho_0=0.0)
    else:
        K = 1.0 + 0.3 * np.sin(2 * np.pi * om.coords[:, 0]) * np.cos(np.pi * om.coords[:, -1])
        model = TransportModel.darcy_three_component(Pe=200, Da=1.0, phi=0.2, alpha=0.01, dt=0.01, K=K, zh=zh, SUPG=2)
        oprm = forms.darcy_three_component_params(Pe=200, Da=1.0, phi=0.2, alpha=0.01, dt=0.01, K=K, zhat=zh, SUPG=2)
    uval = np.zeros(dim)
    uval[-1] = -0.5
    X = om.coords
    f = np.abs(np.sin(3 * np.pi * X[:, 0])) * (1 - np.tanh((top - X[:, -1]) / 0.1))
    bd = make_boundary_data(pm, lambda x: x[..., -1] > top - 1e-12, uval, lambda x: x[..., -1] > top - 1e-12, 0.1)
    obc = modef.make_bcs(om, lambda x: x[-1] > top - 1e-12, lambda x: uval, lambda x: x[-1] > top - 1e-12, 0.1)
    return pm, om, model, oprm, bd, obc, f


@pytest.mark.parametrize("dim,n,kind", [(2, 12, "ccs3"), (3, 6, "ccs3"), (3, 5, "darcy2"), (2, 10, "darcy3"), (3, 5, "darcy3")])
def test_modef_steps_match_oracle(ctx, dim, n, kind):
    """Per solve <= 1e-10, after N steps <= 1e-8 relative L2 (BASELINE.json north_star tolerances)."""
    from mupopp_b200.transport import TransportSolverF
    pm, om, model, oprm, bd, obc, f = _setup(dim, n, kind)
    assert np.array_equal(bd.p_isbc.nonzero()[0], obc.p_dofs)
    assert np.allclose(bd.flux_load, obc.flux_load, rtol=0, atol=1e-15)
    assert np.array_equal(bd.c0_isbc.nonzero()[0], obc.c0_dofs)
    asm = forms.Assembler(om)
    rng = np.random.default_rng(0)
    c1_0 = 0.1 + 0.01 * rng.random(om.nv)
    c2_0 = 0.01 * rng.random(om.nv)
    c0_0 = 0.05 * rng.random(om.nv)
    st = modef.StateF(u=np.zeros((om.nc, dim)), p=np.zeros(om.nv), c0=c0_0.copy(), c1=c1_0.copy(), c2=c2_0.copy())
    s = TransportSolverF(ctx, pm, model, bd, f_source=f, rtol=1e-14)
    s.set_state(c0=c0_0, c1=c1_0, c2=c2_0)
    nsteps = 4
    for k in range(nsteps):
        prev = st
        st = modef.step(asm, oprm, st, f, obc)
        # per-solve parity: restart the GPU from the oracle's previous state
        s.set_state(c0=prev.c0, c1=prev.c1, c2=prev.c2 if prev.c2 is not None else None, p=prev.p, u=prev.u)
        res = s.step()
        assert all(r is None or r.converged for r in res.values()), res
        tol = 1e-10
        assert rel(s.c1.cpu().numpy(), st.c1) < tol
        if model.ncomp == 3:
            assert rel(s.c2.cpu().numpy(), st.c2) < tol
        assert rel(s.c0.cpu().numpy(), st.c0) < tol
        assert rel(s.p.cpu().numpy(), st.p) < tol
        assert rel(s.u.cpu().numpy().reshape(-1, dim), st.u) < tol
    # free-running N steps
    s2 = TransportSolverF(ctx, pm, model, bd, f_source=f, rtol=1e-14)
    s2.set_state(c0=c0_0, c1=c1_0, c2=c2_0)
    for k in range(nsteps):
        s2.step()
    assert rel(s2.c0.cpu().numpy(), st.c0) < 1e-8
    assert rel(s2.p.cpu().numpy(), st.p) < 1e-8
    assert rel(s2.c1.cpu().numpy(), st.c1) < 1e-8


def test_determinism_bitwise(ctx):
    """Assembly + solves are atomic-free: two runs give bit-identical fields."""
    from mupopp_b200.transport import TransportSolverF
    pm, om, model, oprm, bd, obc, f = _setup(3, 6, "ccs3")
    outs = []
    for rep in range(2):
        s = TransportSolverF(ctx, pm, model, bd, f_source=f, rtol=1e-12)
        s.set_state(c1=0.1 * np.ones(om.nv))
        for k in range(3):
            s.step()
        outs.append((s.c0.cpu().numpy().copy(), s.p.cpu().numpy().copy(), s.A.vals.cpu().numpy().copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("method", ["gmres", "bicgstab", "pcg", "pcg_chebyshev"])
def test_krylov_solvers_against_direct(ctx, method):
    """Each Krylov driver reproduces the LU solution of the same CSR system (SPD for pcg, SUPG transport otherwise)."""
    import ctypes as C
    import scipy.sparse.linalg as spla
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    pm, om = _mesh_pair(3, 7)
    asm = forms.Assembler(om)
    dm = DeviceMesh(ctx, pm)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    rng = np.random.default_rng(3)
    emat = ctx.empty(dm.ncell * 16)
    A = plan.new_matrix()
    if method.startswith("pcg"):
        check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 1.0, ptr(emat)))
        plan.scatter_matrix(emat, A)
        isbc = np.zeros(om.nv, dtype=np.uint8)
        isbc[om.coords[:, 2] < 1e-12] = 1
        g = rng.random(om.nv)
    else:
        u = np.tile(np.array([0.5, -0.3, 2.0]), (om.nc, 1))
        ud = ctx.to_device(u.ravel())
        check(ctx.lib.mp_elem_p1_adr_single(ctx.h, C.byref(dm.c), 100.0, 1.0, 0.5, ptr(ud), None, ptr(emat), None))
        plan.scatter_matrix(emat, A)
        isbc = np.zeros(om.nv, dtype=np.uint8)
        isbc[om.coords[:, 2] < 1e-12] = 1
        g = rng.random(om.nv)
    b = rng.standard_normal(om.nv)
    A0 = A.to_scipy()
    bd_ = ctx.to_device(b)
    A.apply_dirichlet(bd_, ctx.to_device(isbc), ctx.to_device(g))
    Ao, bo = fem.apply_dirichlet_symmetric(A0, b, np.nonzero(isbc)[0], g[isbc.astype(bool)])
    assert abs(A.to_scipy() - Ao).max() <= 1e-15 * abs(Ao).max()
    assert rel(bd_.cpu().numpy(), bo) < 1e-15
    x = ctx.zeros(om.nv)
    res = A.solve(method, bd_, x, rtol=1e-14, maxit=5000)
    xo = spla.splu(Ao.tocsc()).solve(bo)
    assert res.converged, res
    assert rel(x.cpu().numpy(), xo) < 1e-10


@pytest.mark.parametrize("dim,n", [(2, 10), (3, 5)])
def test_periodic_constrained_domain_matches_oracle(ctx, dim, n):
    """periodic side faces (constrained_domain of run/CCS/CCS_3D_top_source.py:134-158,211) in mode F:
    dof identification changes the dofmap and the CSR pattern; fields must match the oracle's folded (P^T A P) systems."""
    import mupopp_b200 as mp
    from mupopp_b200.mesh import box_periodic_map
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    if dim == 2:
        pm, om = mp.RectangleMesh((0, 0), (2, 1), n, n), fem.rectangle_mesh(0, 0, 2, 1, n, n)
        axes, lo, hi, zh = (0,), (0.0, 0.0), (2.0, 1.0), (0.0, 1.0)
    else:
        pm, om = mp.BoxMesh((0, 0, 0), (2, 2, 1), n, n, n), fem.box_mesh((0, 0, 0), (2, 2, 1), n, n, n)
        axes, lo, hi, zh = (0, 1), (0.0, 0.0, 0.0), (2.0, 2.0, 1.0), (0.0, 0.0, 1.0)
    pmap = box_periodic_map(pm, axes)
    P, dof = modef.periodic_prolongation(om, axes, lo, hi)
    assert np.array_equal(pmap.vert2dof, dof)
    nd = pmap.ndof
    assert nd == (n ** (dim - 1)) * (n + 1)
    model = TransportModel.ccs_three_component(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zh=zh)
    oprm = forms.ccs_three_component_params(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zhat=zh)
    uval = np.zeros(dim)
    uval[-1] = -0.5
    top = lambda x: x[..., -1] > 1.0 - 1e-12
    bd = make_boundary_data(pm, top, uval, top, 0.1, periodic=pmap)
    obc = modef.make_bcs_periodic(om, axes, dof, lambda x: x[-1] > 1.0 - 1e-12, lambda x: uval, lambda x: x[-1] > 1.0 - 1e-12, 0.1)
    assert np.array_equal(bd.p_isbc.nonzero()[0], obc.p_dofs) and np.allclose(bd.flux_load, obc.flux_load, rtol=0, atol=1e-15)
    X = om.coords
    fv = np.abs(np.sin(np.pi * X[:, 0])) * (1 - np.tanh((1.0 - X[:, -1]) / 0.1))        # periodic in x with period 2... (|sin(pi x)|)
    f = pmap.restrict(fv)
    asm = forms.Assembler(om)
    rng = np.random.default_rng(0)
    c1_0, c2_0, c0_0 = 0.1 + 0.01 * rng.random(nd), 0.01 * rng.random(nd), 0.05 * rng.random(nd)
    st = modef.StateF(u=np.zeros((om.nc, dim)), p=np.zeros(nd), c0=c0_0.copy(), c1=c1_0.copy(), c2=c2_0.copy())
    s = TransportSolverF(ctx, pm, model, bd, f_source=f, rtol=1e-14, vert2dof=pmap.vert2dof)
    # the periodic CSR pattern is the pattern of P^T A P
    rp, ci = s.plan.csr_pattern()
    Mo = (P.T @ asm.mass(1, 1, 2) @ P).tocsr()
    Mo.sort_indices()
    assert np.array_equal(rp.cpu().numpy(), Mo.indptr) and np.array_equal(ci.cpu().numpy(), Mo.indices)
    assert abs(s.M.to_scipy() - Mo).max() <= 1e-15 * abs(Mo).max() * 10
    s.set_state(c0=c0_0, c1=c1_0, c2=c2_0)
    for k in range(3):
        st = modef.step(asm, oprm, st, f, obc, P=P)
        res = s.step()
        assert all(r is None or r.converged for r in res.values()), res
    for nm in ("c0", "c1", "c2", "p"):
        assert rel(getattr(s, nm).cpu().numpy(), getattr(st, nm)) < 1e-8, nm
    assert rel(s.u.cpu().numpy().reshape(-1, dim), st.u) < 1e-8


def test_chebyshev_pressure_preconditioner_gives_the_same_fields(ctx):
    """Chebyshev-Jacobi PCG for the pressure solve (north_star: "Jacobi or Chebyshev preconditioning") reproduces the Jacobi-PCG step."""
    from mupopp_b200.transport import TransportSolverF
    pm, om, model, oprm, bd, obc, f = _setup(3, 6, "ccs3")
    out = []
    for pre in ("jacobi", "chebyshev"):
        s = TransportSolverF(ctx, pm, model, bd, f_source=f, rtol=1e-13, p_precond=pre, cheb_degree=3)
        s.set_state(c1=0.1 * np.ones(om.nv))
        its = []
        for k in range(3):
            r = s.step()
            its.append(r["p"].niter)
            assert r["p"].converged
        out.append((s.p.cpu().numpy().copy(), s.c0.cpu().numpy().copy(), its))
    assert rel(out[1][0], out[0][0]) < 1e-10 and rel(out[1][1], out[0][1]) < 1e-10
    assert sum(out[1][2]) < sum(out[0][2])          # fewer (but more expensive) outer iterations


@pytest.mark.parametrize("dim", [2, 3])
def test_line_preconditioner_and_line_pcg(ctx, dim):
    """Block-tridiagonal (lattice line) Jacobi: mp_line_setup/mp_line_apply == SciPy's solve with the tridiagonal part of the operator along
    the lines of the last coordinate, mp_solve_pcg_line == LU, and fewer iterations than point Jacobi on the thin boxes of the reference
    (run/CCS/CCS_3D_top_source.py: [0,4]x[0,4]x[0,1])."""
    import ctypes as C
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import mupopp_b200 as mp
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    from mupopp_b200.transport import lattice_line_stride
    if dim == 3:
        pm, om = mp.BoxMesh((0, 0, 0), (4, 4, 1), 9, 8, 11), fem.box_mesh((0, 0, 0), (4, 4, 1), 9, 8, 11)
        stride = 10 * 9
    else:
        pm, om = mp.RectangleMesh((0, 0), (4, 1), 13, 21), fem.rectangle_mesh(0, 0, 4, 1, 13, 21)
        stride = 14
    assert lattice_line_stride(pm) == stride
    dm = DeviceMesh(ctx, pm)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    rng = np.random.default_rng(5)
    K = 0.5 + rng.random(om.nv)
    emat = ctx.empty(dm.ncell * (dim + 1) ** 2)
    A = plan.new_matrix()
    check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), ptr(ctx.to_device(K)), 0.0, 1.0, ptr(emat)))
    plan.scatter_matrix(emat, A)
    isbc = (om.coords[:, -1] < 1e-12).astype(np.uint8)
    b = rng.standard_normal(om.nv)
    bd_ = ctx.to_device(b)
    A.apply_dirichlet(bd_, ctx.to_device(isbc), ctx.to_device(rng.random(om.nv)))
    Ao = A.to_scipy().tocsr()
    bo = bd_.cpu().numpy()
    # the preconditioner alone
    A.line_setup(stride)
    n = om.nv
    lo = np.asarray(Ao[np.arange(stride, n), np.arange(0, n - stride)]).ravel()
    T = sp.diags([lo, Ao.diagonal(), lo], [-stride, 0, stride]).tocsc()
    r = rng.standard_normal(n)
    z = ctx.zeros(n)
    A.line_apply(ctx.to_device(r), z)
    assert rel(z.cpu().numpy(), spla.splu(T).solve(r)) < 1e-13
    # line PCG against LU and against point Jacobi
    xo = spla.splu(Ao.tocsc()).solve(bo)
    x = ctx.zeros(n)
    res = A.solve("pcg_line", bd_, x, rtol=1e-13, maxit=5000, line_stride=stride)
    xj = ctx.zeros(n)
    resj = A.solve("pcg", bd_, xj, rtol=1e-13, maxit=5000)
    print("line PCG dim %d: %d iterations (point Jacobi %d), err %.2e" % (dim, res.niter, resj.niter, rel(x.cpu().numpy(), xo)))
    assert res.converged and rel(x.cpu().numpy(), xo) < 1e-10
    assert res.niter < 0.7 * resj.niter
    # a row count that is not a multiple of the stride is refused, loudly
    from mupopp_b200._lib import MpError
    with pytest.raises(MpError):
        A.line_setup(stride + 1)


def test_mode_f_step_with_line_preconditioned_pressure(ctx):
    """TransportSolverF(p_precond="line") gives the same fields as the Jacobi-PCG default (the preconditioner changes the iteration count,
    not the converged answer), also with periodic side faces (nx*ny dofs per plane)."""
    import mupopp_b200 as mp
    from mupopp_b200.mesh import box_periodic_map
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    pm = mp.BoxMesh((0, 0, 0), (4, 4, 1), 8, 8, 10)
    top = lambda x: x[2] > 1.0 - 1e-12
    model = TransportModel.ccs_three_component(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, zh=(0.0, 0.0, 1.0))
    for periodic in (False, True):
        pmap = box_periodic_map(pm, (0, 1)) if periodic else None
        bd = make_boundary_data(pm, top, lambda x: np.array([0.0, 0.0, -0.5]), top, 0.1, periodic=pmap)
        out = {}
        for pc in ("jacobi", "line"):
            s = TransportSolverF(ctx, pm, model, bd, rtol=1e-13, vert2dof=None if pmap is None else pmap.vert2dof, p_precond=pc)
            nd = s.nloc
            s.set_state(c1=0.1 * np.ones(nd))
            its = 0
            for _ in range(3):
                its += s.step()["p"].niter
            assert pc != "line" or (s.p_method == "pcg_line" and s.line_stride == (64 if periodic else 81))
            out[pc] = (s.p.cpu().numpy().copy(), s.c0.cpu().numpy().copy(), its)
        print("mode F periodic=%s: pressure iterations jacobi %d, line %d" % (periodic, out["jacobi"][2], out["line"][2]))
        assert rel(out["line"][0], out["jacobi"][0]) < 1e-10 and rel(out["line"][1], out["jacobi"][1]) < 1e-10
        assert out["line"][2] < out["jacobi"][2]


def test_line_bicgstab_on_the_transport_row(ctx):
    """mp_solve_bicgstab_line (non-symmetric tridiagonal line blocks, mp_line_setup(spd=0)) == LU on a SUPG advection-diffusion row
    with the flow along the lines, in fewer iterations than Jacobi-BiCGStab."""
    import ctypes as C
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import mupopp_b200 as mp
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    pm, om = mp.BoxMesh((0, 0, 0), (4, 4, 1), 8, 7, 12), fem.box_mesh((0, 0, 0), (4, 4, 1), 8, 7, 12)
    stride = 9 * 8
    dm = DeviceMesh(ctx, pm)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    rng = np.random.default_rng(11)
    u = np.tile(np.array([0.3, -0.2, -4.0]), (om.nc, 1)) + 0.1 * rng.standard_normal((om.nc, 3))
    emat = ctx.empty(dm.ncell * 16)
    A = plan.new_matrix()
    check(ctx.lib.mp_elem_p1_adr_single(ctx.h, C.byref(dm.c), 500.0, 1.0, 0.1, ptr(ctx.to_device(u.ravel())), None, ptr(emat), None))
    plan.scatter_matrix(emat, A)
    isbc = (om.coords[:, 2] > 1.0 - 1e-12).astype(np.uint8)
    bd_ = ctx.to_device(rng.standard_normal(om.nv))
    A.apply_dirichlet(bd_, ctx.to_device(isbc), ctx.to_device(rng.random(om.nv)))
    Ao, bo, n = A.to_scipy().tocsr(), bd_.cpu().numpy(), om.nv
    A.line_setup(stride, spd=False)
    lo = np.asarray(Ao[np.arange(stride, n), np.arange(0, n - stride)]).ravel()
    up = np.asarray(Ao[np.arange(0, n - stride), np.arange(stride, n)]).ravel()
    assert np.abs(lo - up).max() > 1e-3 * np.abs(lo).max()                       # genuinely non-symmetric blocks
    T = sp.diags([lo, Ao.diagonal(), up], [-stride, 0, stride]).tocsc()
    r = rng.standard_normal(n)
    z = ctx.zeros(n)
    A.line_apply(ctx.to_device(r), z)
    assert rel(z.cpu().numpy(), spla.splu(T).solve(r)) < 1e-13
    xo = spla.splu(Ao.tocsc()).solve(bo)
    x, xj = ctx.zeros(n), ctx.zeros(n)
    res = A.solve("bicgstab_line", bd_, x, rtol=1e-13, maxit=2000, line_stride=stride, refresh_jacobi=False)
    resj = A.solve("bicgstab", bd_, xj, rtol=1e-13, maxit=2000)
    print("line BiCGStab: %d iterations (point Jacobi %d), err %.2e" % (res.niter, resj.niter, rel(x.cpu().numpy(), xo)))
    assert res.converged and rel(x.cpu().numpy(), xo) < 1e-10
    assert res.niter < resj.niter


@pytest.mark.parametrize("dim,periodic", [(2, False), (3, False), (3, True)])
def test_fused_rowwise_c0_assembly_equals_element_pipeline(ctx, dim, periodic):
    """mp_assemble_p1_transport_rows (one pass: CSR + SELL values, rhs, symmetric Dirichlet, Jacobi diagonal, line bands) reproduces the
    element kernel + scatter + Dirichlet + Jacobi + SELL-fill + band-extract pipeline to round-off, and the step built on it gives
    the same fields."""
    import mupopp_b200 as mp
    from mupopp_b200.mesh import box_periodic_map
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    if dim == 3:
        pm = mp.BoxMesh((0, 0, 0), (4, 4, 1), 7, 6, 9)
        zh, top, uv = (0.0, 0.0, 1.0), (lambda x: x[2] > 1.0 - 1e-12), np.array([0.0, 0.0, -0.5])
    else:
        pm = mp.RectangleMesh((0, 0), (4, 1), 11, 9)
        zh, top, uv = (0.0, 1.0), (lambda x: x[1] > 1.0 - 1e-12), np.array([0.0, -0.5])
    model = TransportModel.ccs_three_component(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, zh=zh)
    pmap = box_periodic_map(pm, (0, 1)) if periodic else None
    bd = make_boundary_data(pm, top, lambda x: uv, top, 0.1, periodic=pmap)
    rng = np.random.default_rng(17)
    sols = []
    for fused in (False, True):
        s = TransportSolverF(ctx, pm, model, bd, rtol=1e-13, vert2dof=None if pmap is None else pmap.vert2dof, fused_rows=fused)
        nd = s.nloc
        if not sols:
            st = dict(c0=0.05 + 0.03 * rng.random(nd), c1=0.1 + 0.02 * rng.random(nd), c2=0.01 * rng.random(nd),
                      u=0.4 * rng.standard_normal(pm.num_cells() * dim) - np.tile(np.eye(dim)[-1] * 0.5, pm.num_cells()),
                      f=rng.random(nd))
        s.set_state(c0=st["c0"], c1=st["c1"], c2=st["c2"], u=st["u"])
        s.fsrc.copy_(ctx.to_device(st["f"]))
        s.c1_new.copy_(ctx.to_device(0.99 * st["c1"]))
        s.c2_new.copy_(ctx.to_device(1.01 * st["c2"] + 1e-4))
        s.assemble_c0_row()
        s.A.sync_sell()
        assert s.c0_line
        _, lo, invw, mb = s.A._line
        sols.append(dict(vals=s.A.vals.cpu().numpy().copy(), rhs=s.rhs_c0.cpu().numpy().copy(), dinv=s.A.dinv.cpu().numpy().copy(),
                         sell=s.A.sell_vals.cpu().numpy().copy(), lo=lo.cpu().numpy().copy(), invw=invw.cpu().numpy().copy(),
                         mb=mb.cpu().numpy().copy()))
        s.set_state(c0=st["c0"], c1=st["c1"], c2=st["c2"], u=st["u"])
        for _ in range(2):
            s.step()
        sols[-1]["c0"] = s.c0.cpu().numpy().copy()
        sols[-1]["p"] = s.p.cpu().numpy().copy()
    a, b = sols
    scale = np.abs(a["vals"]).max()
    for k in ("vals", "sell", "lo"):
        assert np.abs(a[k] - b[k]).max() <= 1e-13 * scale, k
    assert rel(b["rhs"], a["rhs"]) < 1e-13 and rel(b["dinv"], a["dinv"]) < 1e-13
    assert rel(b["invw"], a["invw"]) < 1e-11 and rel(b["mb"], a["mb"]) < 1e-11
    assert (a["vals"] != 0).sum() == (b["vals"] != 0).sum()                      # the same entries eliminated by the Dirichlet rows
    assert rel(b["c0"], a["c0"]) < 1e-10 and rel(b["p"], a["p"]) < 1e-10


def test_shared_reaction_solve_equals_two_mass_solves(ctx):
    """The c1 and c2 rows share their load up to a factor (mupopp.py:936-937, 1234-1235), so the default step does ONE mass solve for the common
    increment; the fields equal those of the two separate solves to solver tolerance, natural and periodic walls."""
    import mupopp_b200 as mp
    from mupopp_b200.mesh import box_periodic_map
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    pm = mp.BoxMesh((0, 0, 0), (4, 4, 1), 7, 6, 9)
    top = lambda x: x[2] > 1.0 - 1e-12
    model = TransportModel.ccs_three_component(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, zh=(0.0, 0.0, 1.0))
    for periodic in (False, True):
        pmap = box_periodic_map(pm, (0, 1)) if periodic else None
        bd = make_boundary_data(pm, top, lambda x: np.array([0.0, 0.0, -0.5]), top, 0.1, periodic=pmap)
        out = {}
        for shared in (False, True):
            s = TransportSolverF(ctx, pm, model, bd, rtol=1e-13, vert2dof=None if pmap is None else pmap.vert2dof, shared_reaction_solve=shared)
            s.set_state(c1=0.1 * np.ones(s.nloc))
            for _ in range(4):
                res = s.step()
            assert all(r is None or r.converged or r.relres < 1e-12 for r in res.values())
            assert (res["c2"].niter == 0) == shared
            out[shared] = {k: getattr(s, k).cpu().numpy().copy() for k in ("c0", "c1", "c2", "p")}
        assert np.abs(out[True]["c2"]).max() > 1e-6                      # the precipitate has formed: not a trivial comparison
        for k in ("c0", "c1", "c2", "p"):
            assert rel(out[True][k], out[False][k]) < 1e-10, (periodic, k)
