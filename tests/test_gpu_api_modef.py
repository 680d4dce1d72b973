"""The drop-in API reaches the benchmarked mode-F path: periodic `constrained_domain` spaces, a P1 permeability field, facet-normal
inflow data, scalar functionals (`assemble(c*phi*dot(u, n)*ds(1))`, volume integrals, `errornorm`) -- all against the oracle.
Reference call sites: run/CCS2D/CCS_3comp.py:112-143,206-283,292-295; run/CCS2D/CCS_layered_permeability.py:165-254;
run/stokes_ADR_3D/master_advective_stokes.py:232-265; benchmark/sphere_3D_pure_shear/sphere.py:157-159."""
import os
import sys

import numpy as np
import pytest

from oracle import fem, forms, functionals as ofn, modef

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _source(X, top):
    g = np.zeros(X.shape[0])
    for ii in range(20):
        g += 0.1 * np.abs(np.sin(ii * X[:, 0] * np.pi))
    return (1.0 - np.tanh((top - X[:, 1]) / 0.01)) * g


def test_periodic_ccs_script_flow_matches_oracle(ctx):
    """examples/ccs2d_periodic_flux.py (the call sequence of run/CCS2D/CCS_3comp.py, periodic in x) == oracle mode F on the folded space,
    including the bottom-flux functional and project(perm, X)."""
    import ccs2d_periodic_flux as ex
    import mupopp_b200.fenics as fx
    nx, ny, nsteps = 16, 8, 3
    mesh = fx.RectangleMesh(fx.Point(0.0, 0.0), fx.Point(8.0, 2.0), nx, ny)
    sol, times, flux, perm1 = ex.run(nsteps=nsteps, verbose=False, mesh=mesh)
    om = fem.rectangle_mesh(0, 0, 8, 2, nx, ny)
    asm = forms.Assembler(om)
    P, dof = modef.periodic_prolongation(om, (0,), (0.0, 0.0), (8.0, 2.0))
    nd = P.shape[1]
    first = np.full(nd, -1, dtype=np.int64)
    first[dof[::-1]] = np.arange(om.nv)[::-1]
    Xd = om.coords[first]
    top = lambda x: x[1] > 2.0 - 1e-12
    bcs = modef.make_bcs_periodic(om, (0,), dof, top, lambda x: np.array([0.0, -0.5]), top, 0.1)
    prm = forms.ccs_three_component_params(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, zhat=(0.0, 1.0))
    f = _source(Xd, 2.0)
    st = modef.StateF(u=np.zeros((om.nc, 2)), p=np.zeros(nd), c0=np.zeros(nd), c1=0.1 * np.ones(nd), c2=np.zeros(nd))
    bottom = ofn.boundary_facets_where(om, lambda x: x[1] < 1e-14)
    oflux = []
    for k in range(nsteps):
        st = modef.step(asm, prm, st, f, bcs, P=P)
        oflux.append(ofn.boundary_flux(asm, bottom, [(P @ st.c0, 1)], ("DG0", st.u), coef=0.05))
    d = 2
    assert sol.space.periodic is not None and sol.blocks[d].numel() == nd
    for name, blk in (("p", d), ("c0", d + 1), ("c1", d + 2), ("c2", d + 3)):
        assert rel(sol.get(blk), getattr(st, name)) < 1e-8, name
    assert rel(sol.cell_velocity.cpu().numpy().reshape(-1, 2), st.u) < 1e-8
    # a boundary functional of fields that agree to 1e-8 (Krylov rtol 1e-12 on both sides, different preconditioners): 1e-7 relative
    assert np.allclose(flux, oflux, rtol=1e-7, atol=1e-14), (flux, oflux)
    assert abs(flux[-1]) > 1e-6                                   # a real number, not a trivially zero functional
    # the nodal velocity written to file = volume-weighted vertex average of the cellwise velocity
    vol = np.abs(asm.detJ) / 2.0
    num = np.zeros((nd, 2))
    den = np.zeros(nd)
    for i in range(3):
        np.add.at(num, dof[om.cells[:, i]], vol[:, None] * st.u)
        np.add.at(den, dof[om.cells[:, i]], vol)
    assert rel(np.stack([sol.get(0), sol.get(1)], 1), num / den[:, None]) < 1e-8
    kx = np.where((Xd[:, 1] < 2.0) & (Xd[:, 1] > 1.0), 1.0, 0.001)
    assert np.array_equal(perm1.get(), kx)


def test_layered_permeability_and_normal_inflow_match_oracle(ctx):
    """examples/ccs2d_layered_permeability.py (K=darcy.perm Expression, BoundarySource.eval_cell inflow, two-component DarcyAdvection,
    gam_rho=-0.5, rho_0=0) == oracle mode F with the nodal K field."""
    import ccs2d_layered_permeability as ex
    import mupopp_b200.fenics as fx
    nx, ny, nsteps = 16, 8, 3
    mesh = fx.RectangleMesh(fx.Point(0.0, 0.0), fx.Point(8.0, 2.0), nx, ny)
    sol, perm1 = ex.run(nsteps=nsteps, verbose=False, mesh=mesh)
    om = fem.rectangle_mesh(0, 0, 8, 2, nx, ny)
    asm = forms.Assembler(om)
    P, dof = modef.periodic_prolongation(om, (0,), (0.0, 0.0), (8.0, 2.0))
    nd = P.shape[1]
    first = np.full(nd, -1, dtype=np.int64)
    first[dof[::-1]] = np.arange(om.nv)[::-1]
    Xd = om.coords[first]
    top = lambda x: x[1] > 2.0 - 1e-12

    def inflow(x):                                 # -g(x) n with n = (0, 1) on the top boundary
        g = sum(0.1 * abs(np.sin(ii * x[0] * np.pi)) for ii in range(20))
        return np.array([0.0, -g])

    bcs = modef.make_bcs_periodic(om, (0,), dof, top, inflow)
    Kd = np.where((Xd[:, 1] < 2.0) & (Xd[:, 1] > 1.0), 1.0, 0.001)
    prm = forms.darcy_two_component_params(Pe=100.0, Da=1.0, phi=0.1, alpha=0.001, dt=0.1, K=P @ Kd, zhat=(0.0, 1.0), gam_rho=-0.5, rho_0=0.0)
    f = _source(Xd, 2.0)
    st = modef.StateF(u=np.zeros((om.nc, 2)), p=np.zeros(nd), c0=np.zeros(nd), c1=0.01 * np.ones(nd), c2=None)
    for k in range(nsteps):
        st = modef.step(asm, prm, st, f, bcs, P=P)
    assert np.array_equal(perm1.get(), Kd)
    for name, blk in (("p", 2), ("c0", 3), ("c1", 4)):
        assert rel(sol.get(blk), getattr(st, name)) < 1e-8, name
    assert rel(sol.cell_velocity.cpu().numpy().reshape(-1, 2), st.u) < 1e-8


@pytest.mark.parametrize("dim", [2, 3])
def test_scalar_functionals_and_errornorm(ctx, dim):
    """assemble(c*dx), assemble(a*c0*c1*dx), assemble(c*dot(u, n)*ds(id)), assemble(c*dot(u, Constant)*ds(id)) with P1/P2/P3 factors and a
    Lagrange velocity, and errornorm(u, uh, 'L2') -- against oracle/functionals.py."""
    import mupopp_b200.fenics as fx
    if dim == 2:
        mesh, om = fx.RectangleMesh((0, 0), (2, 1), 7, 5), fem.rectangle_mesh(0, 0, 2, 1, 7, 5)
    else:
        mesh, om = fx.BoxMesh((0, 0, 0), (2, 1, 1), 4, 3, 3), fem.box_mesh((0, 0, 0), (2, 1, 1), 4, 3, 3)
    asm = forms.Assembler(om)
    rng = np.random.default_rng(5)
    cell = mesh.ufl_cell()
    Q1 = fx.FunctionSpace(mesh, "CG", 1)
    Q2 = fx.FunctionSpace(mesh, "CG", 2)
    V2 = fx.VectorFunctionSpace(mesh, "CG", 2)
    V3 = fx.FunctionSpace(mesh, fx.VectorElement("Lagrange", cell, 3))
    c0, c1, q2, u2, u3, w3 = fx.Function(Q1), fx.Function(Q1), fx.Function(Q2), fx.Function(V2), fx.Function(V3), fx.Function(V3)
    h = {}
    for name, f in (("c0", c0), ("c1", c1), ("q2", q2), ("u2", u2), ("u3", u3), ("w3", w3)):
        h[name] = [rng.standard_normal(b.numel()) for b in f.blocks]
        f.vector().set_local(np.concatenate(h[name]))
    # volume functionals (master_advective_stokes.py:248-264)
    assert abs(fx.assemble(c0 * fx.dx) - ofn.volume_integral(asm, [(h["c0"][0], 1)])) < 1e-12
    v = fx.assemble(0.3 * 0.1 * c0 * c1 * fx.dx)
    assert abs(v - ofn.volume_integral(asm, [(h["c0"][0], 1), (h["c1"][0], 1)], 0.03)) < 1e-12 * max(1.0, abs(v))
    v = fx.assemble(c0 * q2 * (1.0 - 0.05) * fx.dx)
    assert abs(v - ofn.volume_integral(asm, [(h["c0"][0], 1), (h["q2"][0], 2)], 0.95)) < 1e-12 * max(1.0, abs(v))
    # boundary fluxes on a marked part of the boundary (CCS_3comp.py:112-119,277; master_advective_stokes.py:232-243)

    class Left(fx.SubDomain):
        def inside(self, x, on_boundary):
            return x[0] < fx.DOLFIN_EPS

    facets = fx.FacetFunction("size_t", mesh)
    Left().mark(facets, 2)
    ds = fx.Measure("ds", subdomain_data=facets)
    n = fx.FacetNormal(mesh)
    left = ofn.boundary_facets_where(om, lambda x: x[0] < 1e-14)
    v = fx.assemble(c0 * 0.05 * fx.dot(u2, n) * ds(2))
    o = ofn.boundary_flux(asm, left, [(h["c0"][0], 1)], ("P2", h["u2"]), coef=0.05)
    assert abs(v - o) < 1e-12 * max(1.0, abs(o))
    v = fx.assemble(c0 * fx.dot(u3, -n) * ds(2))
    o = ofn.boundary_flux(asm, left, [(h["c0"][0], 1)], ("P3", h["u3"]), coef=-1.0)
    assert abs(v - o) < 1e-12 * max(1.0, abs(o))
    unit = fx.Constant(tuple([1.0] + [0.0] * (dim - 1)))
    v = fx.assemble(c0 * fx.dot(u2, unit) * ds(2))
    o = ofn.boundary_flux(asm, left, [(h["c0"][0], 1)], ("P2", h["u2"]), direction=[1.0] + [0.0] * (dim - 1))
    assert abs(v - o) < 1e-12 * max(1.0, abs(o))
    whole = ofn.boundary_facets_where(om, lambda x: True)
    v = fx.assemble(q2 * fx.dot(u2, n) * fx.ds)
    o = ofn.boundary_flux(asm, whole, [(h["q2"][0], 2)], ("P2", h["u2"]))
    assert abs(v - o) < 1e-12 * max(1.0, abs(o))
    # errornorm (sphere.py:157-159): two Functions of one space
    e = fx.errornorm(u3, w3, "L2")
    assert abs(e - ofn.l2_error(asm, h["u3"], h["w3"], 3)) < 1e-12 * e
    assert fx.errornorm(c0, c0, "L2") == 0.0
    # CFL time step (core.compute_dt, modules/core.py:23-48)
    from mupopp_b200 import core
    um = core.u_max(u2)
    assert abs(um - ofn.cell_mean_speed_max(asm, ("P2", h["u2"]))) < 1e-12 * um
    assert abs(core.compute_dt(u2, 0.5, mesh.hmin()) - 0.5 * mesh.hmin() / um) < 1e-15
    # known answer: the linear field u = E.r is reproduced exactly by its own interpolant
    E = fx.Expression(("0.05*x[0]", "-0.05*x[1]") if dim == 2 else ("0.05*x[0]", "0.05*x[1]", "-0.1*x[2]"), degree=2)
    ue = fx.Function(V2)
    ue.interpolate(E)
    assert fx.errornorm(E, ue, "L2") < 1e-15


def test_mode_selection_and_time_dependent_bc(ctx):
    """solver_parameters / set_mode pick the discretisation; a Constant boundary value re-assigned between solves is picked up (DOLFIN
    re-applies BC values on every solve)."""
    import mupopp_b200.fenics as fx
    from mupopp_b200 import mupopp as api
    mesh = fx.UnitSquareMesh(6, 6)
    cell = mesh.ufl_cell()
    Qc = fx.FiniteElement("Lagrange", cell, 1)
    W = fx.FunctionSpace(mesh, fx.MixedElement([fx.VectorElement("Lagrange", cell, 2), Qc, Qc, Qc, Qc]))
    top = lambda x: x[1] > 1.0 - fx.DOLFIN_EPS
    cval = fx.Constant(0.1)
    bcs = [fx.DirichletBC(W.sub(0), fx.Constant((0.0, -0.5)), top), fx.DirichletBC(W.sub(2), cval, top)]
    ccs = api.CCS(Pe=500, Da=0.1, phi=0.05, alpha=1e-3)
    out = {}
    for mode in ("R", "F"):
        s0, s = fx.Function(W), fx.Function(W)
        s0.blocks[4].fill_(0.1)
        a, L = ccs.advection_diffusion_three_component(W, mesh, s0, 0.1, f_source=fx.Constant(0.0), K=1.0)
        fx.solve(a == L, s, bcs, solver_parameters={"mupopp_b200_mode": mode})
        out[mode] = s
    assert out["F"].cell_velocity is not None and out["R"].cell_velocity is None
    topv = np.where(mesh.coords[:, 1] > 1 - 1e-12)[0]
    assert np.allclose(out["F"].get(3)[topv], 0.1) and np.allclose(out["R"].get(3)[topv], 0.1)
    assert rel(out["F"].get(3), out["R"].get(3)) < 0.5                      # same PDE, two discretisations
    cval.assign(0.3)
    for mode in ("R", "F"):
        s0, s = fx.Function(W), fx.Function(W)
        s0.blocks[4].fill_(0.1)
        a, L = ccs.advection_diffusion_three_component(W, mesh, s0, 0.1, f_source=fx.Constant(0.0), K=1.0)
        fx.solve(a == L, s, bcs, solver_parameters={"mupopp_b200_mode": mode})
        assert np.allclose(s.get(3)[topv], 0.3), mode


def test_lagged_velocity_recovered_from_the_state(ctx):
    """solver_parameters={"mupopp_b200_lagged_velocity": "recover"}: the cellwise velocity of the previous state is recomputed from (p, c0)
    instead of travelling with the Function -- bit-identical to the cached one whenever the previous state came out of a mode-F solve, so a
    host-resident workflow only ships the public blocks of the mixed Function (what bench.py's e2e leg does)."""
    import mupopp_b200.fenics as fx
    from mupopp_b200 import mupopp as api
    mesh = fx.UnitSquareMesh(8, 8)
    cell = mesh.ufl_cell()
    Qc = fx.FiniteElement("Lagrange", cell, 1)
    W = fx.FunctionSpace(mesh, fx.MixedElement([fx.VectorElement("Lagrange", cell, 2), Qc, Qc, Qc, Qc]))
    top = lambda x: x[1] > 1.0 - fx.DOLFIN_EPS
    bcs = [fx.DirichletBC(W.sub(0), fx.Constant((0.0, -0.5)), top), fx.DirichletBC(W.sub(2), fx.Constant(0.1), top)]
    ccs = api.CCS(Pe=500, Da=0.1, phi=0.05, alpha=1e-3)
    prm = {"mupopp_b200_mode": "F"}
    s0, s1, s2 = fx.Function(W), fx.Function(W), fx.Function(W)
    s0.blocks[4].fill_(0.1)
    for prev, new in ((s0, s1), (s1, s2)):                          # two steps with the cached velocity
        a, L = ccs.advection_diffusion_three_component(W, mesh, prev, 0.1, f_source=fx.Constant(1.0), K=1.0)
        fx.solve(a == L, new, bcs, solver_parameters=prm)
    assert float(s2.cell_velocity.abs().max()) > 0.0
    # third step: (a) from s2 with its cached velocity, (b) from a copy of s2's PUBLIC blocks only, velocity recovered on the device
    ra, rb, bare = fx.Function(W), fx.Function(W), fx.Function(W)
    for dst, src in zip(bare.blocks, s2.blocks):
        dst.copy_(src)
    assert bare.cell_velocity is None
    a, L = ccs.advection_diffusion_three_component(W, mesh, s2, 0.1, f_source=fx.Constant(1.0), K=1.0)
    fx.solve(a == L, ra, bcs, solver_parameters=prm)
    a, L = ccs.advection_diffusion_three_component(W, mesh, bare, 0.1, f_source=fx.Constant(1.0), K=1.0)
    fx.solve(a == L, rb, bcs, solver_parameters=dict(prm, mupopp_b200_lagged_velocity="recover"))
    for k in range(len(ra.blocks)):
        assert np.array_equal(ra.get(k), rb.get(k)), k
    assert np.array_equal(ra.cell_velocity.cpu().numpy(), rb.cell_velocity.cpu().numpy())
    with pytest.raises(ValueError):
        fx.solve(a == L, rb, bcs, solver_parameters=dict(prm, mupopp_b200_lagged_velocity="nonsense"))
