"""Host-side logic of the dolfin facade and of the mupopp API module (no GPU): spaces, topological DirichletBC marking for P1/P2/P3,
marker-based BCs on the reference's fixture, Expression evaluation, the .cfg parser."""
import os

import numpy as np
import pytest

from mupopp_b200 import fenics as fx
from mupopp_b200 import mupopp as api
from oracle import fem, forms

G = os.path.join(os.path.dirname(__file__), "golden", "fenics_demo_dofmap.npz")


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("deg", [1, 2, 3])
def test_dirichlet_topological_marking_matches_coordinates(dim, deg):
    mesh = fx.RectangleMesh((0, 0), (2, 1), 4, 3) if dim == 2 else fx.BoxMesh((0, 0, 0), (2, 1, 1), 3, 2, 2)
    om = fem.rectangle_mesh(0, 0, 2, 1, 4, 3) if dim == 2 else fem.box_mesh((0, 0, 0), (2, 1, 1), 3, 2, 2)
    V = fx.FunctionSpace(mesh, fx.VectorElement("Lagrange", mesh.ufl_cell(), deg))
    s = fem.ScalarSpace(om, deg)
    assert V.dim() == dim * s.ndof and np.allclose(V.dof_coords(deg), s.dof_coords(), rtol=0, atol=1e-15)
    top = lambda x, on_boundary: on_boundary and x[dim - 1] > 1.0 - fx.DOLFIN_EPS
    val = fx.Expression(("x[0]",) + ("0.0",) * (dim - 2) + ("-0.5",), degree=1)
    bc = fx.DirichletBC(V, val, top)
    dv = bc.dofs_and_values()
    X = s.dof_coords()
    want = np.where(X[:, dim - 1] > 1 - 1e-12)[0]
    for c in range(dim):
        dofs, vals = dv[c]
        assert np.array_equal(dofs, want)
    assert np.allclose(dv[0][1], X[want, 0]) and np.allclose(dv[dim - 1][1], -0.5)
    # a sub-space BC only touches its blocks
    W = fx.FunctionSpace(mesh, fx.MixedElement([fx.VectorElement("Lagrange", mesh.ufl_cell(), deg), fx.FiniteElement("Lagrange", mesh.ufl_cell(), 1)]))
    bcp = fx.DirichletBC(W.sub(1), fx.Constant(3.0), top)
    assert list(bcp.dofs_and_values()) == [dim]
    bcy = fx.DirichletBC(W.sub(0).sub(dim - 1), fx.Constant(1.0), top)
    assert list(bcy.dofs_and_values()) == [dim - 1]


def test_marker_based_bc_on_the_reference_fixture():
    g = np.load(G)
    mesh = fx.Mesh(g["coords"], g["cells"])
    mf = fx.MeshFunction("size_t", mesh, g["facet_markers"])
    Q = fx.FunctionSpace(mesh, "CG", 1)
    dofs, vals = fx.DirichletBC(Q, fx.Constant(1.0), mf, 1).dofs_and_values()[0]
    assert len(dofs) == 21 and (vals == 1.0).all()
    V = fx.VectorFunctionSpace(mesh, "CG", 2)
    d2 = fx.DirichletBC(V, fx.Constant((0.0, 0.0)), mf, 1).dofs_and_values()
    assert len(d2[0][0]) == 21 + 20                     # vertices + one edge midpoint per marked facet


def test_eval_cell_expression_gets_outward_normal():
    mesh = fx.UnitSquareMesh(4, 4)

    class Inflow(fx.Expression):
        def __init__(self, mesh):
            self.mesh = mesh

        def eval_cell(self, values, x, ufc_cell):
            n = fx.Cell(self.mesh, ufc_cell.index).normal(ufc_cell.local_facet)
            values[0], values[1] = 2.0 * n[0], 2.0 * n[1]

        def value_shape(self):
            return (2,)

    V = fx.VectorFunctionSpace(mesh, "CG", 2)
    dv = fx.DirichletBC(V, Inflow(mesh), lambda x: x[1] < fx.DOLFIN_EPS).dofs_and_values()
    assert np.allclose(dv[0][1], 0.0) and np.allclose(dv[1][1], -2.0)       # outward normal of the bottom edge is (0,-1)


def test_string_expressions_and_constants():
    e = fx.Expression("0.05*(exp(-x[0]*x[0])/0.5)+dt", degree=2, dt=0.25)
    v = np.zeros(1)
    e.eval(v, np.array([0.5, 0.0]))
    assert abs(v[0] - (0.05 * np.exp(-0.25) / 0.5 + 0.25)) < 1e-15
    e.dt = 1.0                                           # scripts mutate attributes: diapir.py:140,200
    e.eval(v, np.array([0.0, 0.0]))
    assert abs(v[0] - 1.1) < 1e-15
    c = fx.Constant((0.0, 1.0))
    assert len(c) == 2 and c[1] == 1.0 and float(fx.Constant(2.5)) == 2.5


def test_api_signatures_and_cfg_parser(tmp_path):
    import inspect
    d = api.DarcyAdvection()
    assert (d.Pe, d.Da, d.phi, d.alpha, d.cfl, d.dt) == (100, 10.0, 0.01, 0.005, 1e-2, 1e-2) and d.beta == 0.5
    sig = inspect.signature(api.DarcyAdvection.advection_diffusion_two_component)
    assert list(sig.parameters)[1:] == ["W", "mesh", "sol_prev", "dt", "f1", "K", "zh", "SUPG", "gam_rho", "rho_0"]
    assert sig.parameters["gam_rho"].default == -1.0 and sig.parameters["rho_0"].default == 1.0
    sig = inspect.signature(api.CCS.advection_diffusion_three_component)
    assert list(sig.parameters)[1:] == ["W", "mesh", "sol_prev", "dt", "f_source", "K", "zh", "SUPG", "gam_rho"]
    assert list(inspect.signature(api.compaction.momentum_solver).parameters)[1:] == ["W", "a", "L", "b", "bcs", "pc", "tol", "max_its", "monitor"]
    assert list(inspect.signature(api.StokesAdvection.stokes_ADR_precipitation).parameters)[1:] == ["W", "mesh", "sol_prev", "dt", "SUPG"]
    cfg = tmp_path / "p.cfg"
    cfg.write_text("# comment\n\nlogfile  = log.txt   # name\nout_freq = 5\nT = 5.0\ndt   = 0.1  # step\nda = 0.0001\nR = 0.2\nB = 1e6\ntheta = 10\ndL = 1.0\n")
    p = api.parse_param_file(str(cfg))
    assert p == {"logfile": "log.txt", "out_freq": 5, "T": 5.0, "dt": 0.1, "da": 0.0001, "R": 0.2, "B": 1e6, "theta": 10, "dL": 1.0}
    c = api.compaction(str(cfg))
    assert c.B == 1e6 and isinstance(c.out_freq, int)
    chi = c.surface_tension_2(0.1)
    assert abs(chi - (2 * np.cos(10) - 1) * (2 * 249.0 + 20 * (-8065.0) * 1e-3 + 12 * 6149.0 * 1e-2 + 6 * (-1778.0) * 0.1)) < 1e-9
    ref_cfg = "/root/reference/src/benchmark/time_marching/time_test1.cfg"
    if os.path.exists(ref_cfg):
        q = api.parse_param_file(ref_cfg)
        assert q["da"] == 0.0001 and q["theta"] == 10 and q["dL"] == 1.0 and q["out_freq"] == 5


def test_form_descriptors_pair_up():
    mesh = fx.UnitSquareMesh(2, 2)
    cell = mesh.ufl_cell()
    W = fx.FunctionSpace(mesh, fx.MixedElement([fx.VectorElement("Lagrange", cell, 2), fx.FiniteElement("Lagrange", cell, 1)]))
    a, L = api.DarcyAdvection().darcy_bilinear(W, mesh)
    eq = (a == L)
    assert isinstance(eq, fx.Equation) and eq.a.kind == "darcy_bilinear" and eq.a.data["K"] == 0.1 and eq.a.data["zhat"] == (0.0, 1.0)


def test_periodic_constrained_domain_from_subdomain():
    """`FunctionSpace(mesh, element, constrained_domain=pbc)` with the PeriodicBoundary of run/CCS2D/CCS_3comp.py:131-143 (2-D) and the
    doubly periodic box: the SubDomain.inside/map pair gives the same dof identification as mesh.box_periodic_map."""
    from mupopp_b200.mesh import box_periodic_map
    xmax, ymax = 8.0, 2.0
    mesh = fx.RectangleMesh(fx.Point(0.0, 0.0), fx.Point(xmax, ymax), 8, 4)

    class PeriodicBoundary(fx.SubDomain):
        def inside(self, x, on_boundary):
            return bool(fx.near(x[0], 0.0) and on_boundary)

        def map(self, x, y):
            if fx.near(x[0], xmax):
                y[0] = x[0] - xmax
                y[1] = x[1]
            else:
                y[0] = -1000
                y[1] = -1000

    cell = mesh.ufl_cell()
    Qc = fx.FiniteElement("Lagrange", cell, 1)
    W = fx.FunctionSpace(mesh, fx.MixedElement([fx.VectorElement("Lagrange", cell, 2), Qc, Qc, Qc, Qc]), constrained_domain=PeriodicBoundary())
    ref = box_periodic_map(mesh, (0,))
    assert W.periodic is not None and W.periodic.axes == (0,)
    assert np.array_equal(W.periodic.vert2dof, ref.vert2dof)
    assert all(nd == ref.ndof for (_, _, nd) in W.blocks) and len(W.blocks) == 6          # velocity is stored at the P1 dofs too
    assert W.dim() == 6 * ref.ndof
    X = fx.FunctionSpace(mesh, "CG", 1, constrained_domain=PeriodicBoundary())
    assert X.dim() == ref.ndof

    box = fx.BoxMesh(fx.Point(0, 0, 0), fx.Point(2, 2, 1), 3, 3, 2)

    class PB3(fx.SubDomain):
        def inside(self, x, on_boundary):
            return bool((fx.near(x[0], 0.0) or fx.near(x[1], 0.0)) and not (fx.near(x[0], 2.0) or fx.near(x[1], 2.0)) and on_boundary)

        def map(self, x, y):
            y[:] = x
            if fx.near(x[0], 2.0):
                y[0] = x[0] - 2.0
            if fx.near(x[1], 2.0):
                y[1] = x[1] - 2.0

    P3 = fx.periodic_map_from_subdomain(box, PB3())
    ref3 = box_periodic_map(box, (0, 1))
    assert np.array_equal(P3.vert2dof, ref3.vert2dof) and P3.axes == (0, 1)
    with pytest.raises(AttributeError):
        fx.FunctionSpace(mesh, "CG", 1, constrained_domain=object())


def test_facet_function_and_subdomain_marking():
    """FacetFunction + SubDomain.mark + Measure('ds')[facets] as in run/CCS2D/CCS_3comp.py:112-119."""
    mesh = fx.RectangleMesh(fx.Point(0.0, 0.0), fx.Point(8.0, 2.0), 8, 4)

    class Plane(fx.SubDomain):
        def inside(self, x, on_boundary):
            return x[1] < fx.DOLFIN_EPS

    facets = fx.FacetFunction("size_t", mesh)
    Plane().mark(facets, 1)
    ds = fx.Measure("ds")[facets]
    owner, loc, n, meas = facets.facet_geometry(1)
    assert len(owner) == 8 and np.allclose(n, [[0.0, -1.0]] * 8) and abs(meas.sum() - 8.0) < 1e-14
    m1 = ds(1)
    assert m1.kind == "ds" and m1.sid == 1 and m1.data is facets
