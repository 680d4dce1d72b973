"""SELL-32 SpMV variants against the CSR kernels: compressed column indices (one byte per entry + offset dictionary), 1/2/4 lanes
per row, rectangular operators (padding columns must be valid for the operator), and the solvers on top of them."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture
def reset_options(ctx):
    yield
    for k, v in (("spmv_variant", 0), ("spmv_c8", 1), ("spmv_lpr", 0), ("pcg_variant", -1)):
        ctx.set_option(k, v)


@pytest.mark.parametrize("dim,n,periodic", [(3, 12, False), (3, 10, True), (2, 40, True)])
def test_sell_variants_match_csr(ctx, reset_options, dim, n, periodic):
    import mupopp_b200 as mp
    from mupopp_b200.device import DeviceMesh, Plan
    from mupopp_b200.mesh import box_periodic_map
    mesh = mp.BoxMesh((0, 0, 0), (2, 2, 1), n, n, n + 3) if dim == 3 else mp.RectangleMesh((0, 0), (2, 1), n, n + 5)
    pmap = box_periodic_map(mesh, tuple(range(dim - 1))) if periodic else None
    dm = DeviceMesh(ctx, mesh, pmap.vert2dof if periodic else None)
    plan = Plan(ctx, dm.cell_dofs, dm.cell_dofs, dm.ndof, dm.ndof)
    A = plan.new_matrix()
    # the compressed mirror exists; how many (slice, j) positions stay explicit depends on the mesh-line length (short lines: most)
    assert 0 <= A.sell_explicit <= A.sell_entries // 32
    rng = np.random.default_rng(11)
    A.vals.copy_(ctx.to_device(rng.standard_normal(A.nnz)))
    A.sell_dirty = True
    x = ctx.to_device(rng.standard_normal(dm.ndof))
    y = ctx.zeros(dm.ndof)
    ctx.set_option("spmv_variant", 1)                     # CSR CTA-stream kernel: the independent reference
    yref = A.spmv(x, y).cpu().numpy().copy()
    assert rel(yref, A.to_scipy() @ x.cpu().numpy()) < 1e-14
    ctx.set_option("spmv_variant", 0)
    for c8 in (0, 1):
        for lpr in (1, 2, 4):
            ctx.set_option("spmv_c8", c8)
            ctx.set_option("spmv_lpr", lpr)
            y.fill_(float("nan"))
            got = A.spmv(x, y).cpu().numpy()
            assert rel(got, yref) < 1e-14, (c8, lpr)
    # same (c8, lpr) twice -> same bits
    ctx.set_option("spmv_c8", 1)
    ctx.set_option("spmv_lpr", 4)
    a = A.spmv(x, y).cpu().numpy().copy()
    b = A.spmv(x, y).cpu().numpy().copy()
    assert np.array_equal(a, b)


def test_rectangular_sell_never_reads_past_x(ctx, reset_options):
    """P2 x P1 block (nrow > ncol): x sits directly in front of NaNs; every SELL variant must return the finite CSR result."""
    import mupopp_b200 as mp
    from mupopp_b200 import moder
    mesh = mp.BoxMesh((0, 0, 0), (1, 1, 1), 4, 4, 4)
    sp = moder.Spaces(ctx, mesh)
    BT = sp.assemble("gradT", 2, 1, direction=0)
    assert BT.nrow > BT.ncol
    buf = torch.full((BT.ncol + 4096,), float("nan"), dtype=torch.float64, device=ctx.device)
    rng = np.random.default_rng(2)
    buf[: BT.ncol] = ctx.to_device(rng.standard_normal(BT.ncol))
    x = buf[: BT.ncol]
    y = ctx.zeros(BT.nrow)
    ref = BT.to_scipy() @ x.cpu().numpy()
    for lpr in (1, 2, 4):
        ctx.set_option("spmv_lpr", lpr)
        got = BT.spmv(x, y).cpu().numpy()
        assert np.isfinite(got).all() and rel(got, ref) < 1e-14, lpr


@pytest.mark.parametrize("c8,lpr", [(0, 1), (1, 1), (1, 2), (1, 4), (0, 4)])
def test_solvers_on_every_sell_variant(ctx, reset_options, c8, lpr):
    """PCG and BiCGStab (SpMV fused with the dot products) converge to the same solution on every SELL variant."""
    import scipy.sparse.linalg as spla
    import mupopp_b200 as mp
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    mesh = mp.BoxMesh((0, 0, 0), (2, 2, 1), 9, 9, 9)
    dm = DeviceMesh(ctx, mesh)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    emat = ctx.empty(dm.ncell * 16)
    rng = np.random.default_rng(4)
    isbc = np.zeros(dm.nvert, dtype=np.uint8)
    isbc[mesh.coords[:, 2] < 1e-12] = 1
    g = rng.random(dm.nvert)
    ctx.set_option("spmv_c8", c8)
    ctx.set_option("spmv_lpr", lpr)
    for method in ("pcg", "bicgstab"):
        A = plan.new_matrix()
        if method == "pcg":
            check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 1.0, ptr(emat)))
        else:
            u = ctx.to_device(np.tile(np.array([0.5, -0.3, 2.0]), (dm.ncell, 1)).ravel())
            check(ctx.lib.mp_elem_p1_adr_single(ctx.h, C.byref(dm.c), 100.0, 1.0, 0.5, ptr(u), None, ptr(emat), None))
        plan.scatter_matrix(emat, A)
        b = ctx.to_device(rng.standard_normal(dm.nvert))
        A.apply_dirichlet(b, ctx.to_device(isbc), ctx.to_device(g))
        x = ctx.zeros(dm.nvert)
        res = A.solve(method, b, x, rtol=1e-13, maxit=5000)
        assert res.converged, (method, res)
        xo = spla.splu(A.to_scipy().tocsc()).solve(b.cpu().numpy())
        assert rel(x.cpu().numpy(), xo) < 1e-10, method


@pytest.mark.parametrize("c8", [0, 1])
def test_single_reduction_pcg_matches_direct(ctx, reset_options, c8):
    """The Chronopoulos-Gear single-reduction PCG (the multi-GPU default: one exchange of three sums per iteration, two kernels) solves
    the same SPD systems as the classic two-reduction PCG, with a similar iteration count."""
    import scipy.sparse.linalg as spla
    import mupopp_b200 as mp
    from mupopp_b200._lib import check, ptr
    from mupopp_b200.device import DeviceMesh, Plan
    mesh = mp.BoxMesh((0, 0, 0), (2, 2, 1), 14, 14, 14)
    dm = DeviceMesh(ctx, mesh)
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    emat = ctx.empty(dm.ncell * 16)
    rng = np.random.default_rng(8)
    K = ctx.to_device(np.exp(2.0 * rng.standard_normal(dm.nvert)))            # a rough coefficient: a few hundred iterations
    check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), ptr(K), 0.0, 1.0, ptr(emat)))
    A = plan.new_matrix()
    plan.scatter_matrix(emat, A)
    isbc = np.zeros(dm.nvert, dtype=np.uint8)
    isbc[mesh.coords[:, 2] < 1e-12] = 1
    b = ctx.to_device(rng.standard_normal(dm.nvert))
    A.apply_dirichlet(b, ctx.to_device(isbc), ctx.to_device(rng.random(dm.nvert)))
    xo = spla.splu(A.to_scipy().tocsc()).solve(b.cpu().numpy())
    ctx.set_option("spmv_c8", c8)
    its = {}
    for variant in (0, 1):
        ctx.set_option("pcg_variant", variant)
        x = ctx.zeros(dm.nvert)
        res = A.solve("pcg", b, x, rtol=1e-12, maxit=20000)
        assert res.converged, (variant, res)
        assert rel(x.cpu().numpy(), xo) < 1e-9, variant
        its[variant] = res.niter
    assert its[1] <= 1.3 * its[0] + 10, its
    # the same two recurrences with the lattice-line preconditioner (mp_solve_pcg_line: classic on one GPU, single-reduction on several)
    stride = 15 * 15
    itl = {}
    for variant in (0, 1):
        ctx.set_option("pcg_variant", variant)
        x = ctx.zeros(dm.nvert)
        res = A.solve("pcg_line", b, x, rtol=1e-12, maxit=20000, line_stride=stride)
        assert res.converged, (variant, res)
        assert rel(x.cpu().numpy(), xo) < 1e-9, variant
        itl[variant] = res.niter
    assert itl[1] <= 1.3 * itl[0] + 10 and itl[0] < its[0], (itl, its)
    ctx.set_option("pcg_variant", -1)
    # zero right-hand side: x = 0 whatever the initial guess
    x = ctx.to_device(rng.standard_normal(dm.nvert))
    res = A.solve("pcg", ctx.zeros(dm.nvert), x, rtol=1e-12)
    assert res.converged and res.niter == 0 and float(x.abs().max()) == 0.0
