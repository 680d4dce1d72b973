"""GPU parity of StokesAdvection.stokes_ADR_precipitation (SURVEY.md row a7, P3 velocity) against the oracle's monolithic LU."""
import numpy as np
import pytest

from oracle import fem, forms
from oracle import stokes as os_

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("dim,n", [(2, 4), (3, 2)])
def test_p3_blocks_and_stokes_adr_step(ctx, dim, n):
    import mupopp_b200 as mp
    from mupopp_b200.moder import Spaces
    from mupopp_b200.stokes import StokesAdrR
    if dim == 2:
        pm, om = mp.RectangleMesh((0, 0), (2, 1), 2 * n, n), fem.rectangle_mesh(0, 0, 2, 1, 2 * n, n)
    else:
        pm, om = mp.BoxMesh((0, 0, 0), (2, 1, 1), 2 * n, n, n), fem.box_mesh((0, 0, 0), (2, 1, 1), 2 * n, n, n)
    asm = forms.Assembler(om)
    sp = Spaces(ctx, pm)
    assert np.array_equal(sp.cell_dofs(3).cpu().numpy(), asm.P3.cell_dofs)
    assert np.allclose(sp.dof_coords(3), asm.P3.dof_coords(), rtol=0, atol=1e-15)
    Lo, Bo = os_.stokes_blocks(asm, 3)
    L = sp.assemble("stiff", 3, 3).to_scipy()
    assert abs(L - Lo).max() <= 1e-12 * abs(Lo).max()
    for a in range(dim):
        assert abs(sp.assemble("grad", 1, 3, direction=a).to_scipy() - Bo[a]).max() <= 1e-13 * abs(Bo[a]).max()
    # inflow u = (1,0[,0]) on x = 0, no-slip walls on the other faces except the outflow x = 2; c0 = 0.1 at the inflow
    X3, X1 = asm.P3.dof_coords(), om.coords
    inflow = np.where(X3[:, 0] < 1e-12)[0]
    wall = np.where(((X3[:, 1:] < 1e-12) | (X3[:, 1:] > 1 - 1e-12)).any(1) & (X3[:, 0] > 1e-12))[0]
    n3, n1 = asm.P3.ndof, asm.P1.ndof
    isbc = np.zeros(n3, dtype=np.uint8)
    isbc[inflow] = 1
    isbc[wall] = 1
    g = [np.zeros(n3) for _ in range(dim)]
    g[0][inflow] = 1.0
    D = np.nonzero(isbc)[0]
    cin = np.where(X1[:, 0] < 1e-12)[0]
    c0_isbc = np.zeros(n1, dtype=np.uint8)
    c0_isbc[cin] = 1
    c0_g = np.zeros(n1)
    c0_g[cin] = 0.1
    oprm = os_.stokes_adr_params(Pe=0.1, Da=0.1, dt=1.0, SUPG=2)
    s = StokesAdrR(ctx, pm, 0.1, 0.1, 1.0, 2, [isbc] * dim, g, c0_isbc, c0_g, vdeg=3, rtol=1e-13)
    rng = np.random.default_rng(1)
    prev = forms.State(u=[np.zeros(n3) for _ in range(dim)], p=np.zeros(n1), c0=np.zeros(n1), c1=0.1 + 0.01 * rng.random(n1), c2=np.zeros(n1))
    for k in range(2):
        mono = os_.monolithic_step(asm, oprm, prev, [D] * dim, [g[a][D] for a in range(dim)], cin, 0.1 * np.ones(len(cin)))
        s.up.copy_(ctx.to_device(np.concatenate(list(prev.u) + [prev.p])))
        s.c0.copy_(ctx.to_device(prev.c0)); s.c1.copy_(ctx.to_device(prev.c1)); s.c2.copy_(ctx.to_device(prev.c2))
        info = s.step()
        assert info["c0"].converged and s.stokes.last_result.converged, s.stokes.last_result
        assert rel(s.c1.cpu().numpy(), mono.c1) < 1e-10 and rel(s.c2.cpu().numpy(), mono.c2) < 1e-9
        assert rel(s.c0.cpu().numpy(), mono.c0) < 1e-9
        uu = np.concatenate([v.cpu().numpy() for v in s.u_views()])
        print("stokes dim %d step %d: %d MINRES iterations, relres %.2e, u err %.2e, p err %.2e"
              % (dim, k, info["stokes_its"], s.stokes.last_relres, rel(uu, np.concatenate(mono.u)), rel(s.p.cpu().numpy(), mono.p)))
        assert rel(uu, np.concatenate(mono.u)) < 1e-10            # BASELINE.json: <= 1e-10 per solve (the reference solves with MUMPS)
        assert rel(s.p.cpu().numpy(), mono.p) < 1e-9
        prev = mono
