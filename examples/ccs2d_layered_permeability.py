"""Two-component DarcyAdvection run with a layered permeability FIELD and a facet-normal inflow condition, through the drop-in API --
the calls of /root/reference/src/run/CCS2D/CCS_layered_permeability.py that the constant-K examples do not exercise:

    `darcy.perm = Two_layer_Permeability(mesh, element=Qc, ...)` and `K=darcy.perm` in the form call     :165-181, 233, 250
    `G = BoundarySource(mesh, element=V)` (Expression.eval_cell with `Cell(mesh, index).normal(local_facet)`) as the velocity BC  :184-198, 214-216
    `W = dolfin.FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc]), constrained_domain=pbc)`                :211
    `a, L = darcy.advection_diffusion_two_component(W, mesh, sol_0, dt, f1=S, K=darcy.perm, gam_rho=-0.5, rho_0=0.0)`     :250

A P1 permeability (and a periodic space) runs in mode F; the nodal K is averaged per cell there.
Usage: python examples/ccs2d_layered_permeability.py [mesh_density] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403,E402
from mupopp_b200.mupopp import *   # noqa: F401,F403,E402

Da0, Pe0, alpha0, phi0, dt0 = 1.0, 100.0, 0.001, 0.1, 1.0e-1
xmin, xmax, ymin, ymax = 0.0, 8.0, 0.0, 2.0


class PeriodicBoundary(SubDomain):
    def inside(self, x, on_boundary):
        return bool(near(x[0], xmin) and on_boundary)

    def map(self, x, y):
        if near(x[0], xmax):
            y[0] = x[0] - xmax
            y[1] = x[1]
        else:
            y[0] = -1000
            y[1] = -1000


class SourceTerm(Expression):
    def __init__(self, mesh, element, top_y=1.0):
        self.mesh, self.element, self.top = mesh, element, top_y

    def eval(self, values, x):
        g1 = x[0] * 0.0
        for ii in range(0, 20):
            g1 += 0.1 * np.abs(np.sin(ii * x[0] * np.pi))
        values[0] = (1.0 - tanh((self.top - x[1]) / 0.01)) * g1

    def value_shape(self):
        return (1,)


class Two_layer_Permeability(Expression):
    def __init__(self, mesh, element, top_y=1.0, ktop=0.01, kbot=1.0):
        self.mesh, self.element, self.top, self.k0, self.k1 = mesh, element, top_y, ktop, kbot

    def eval(self, values, x):
        if x[1] < self.top and x[1] > 0.5 * self.top:
            values[0] = self.k0
        else:
            values[0] = self.k1

    def value_shape(self):
        return (1,)


class BoundarySource(Expression):
    """Inflow -g(x) n on the marked facets: evaluated per (cell, local facet) so that the normal is available."""

    def __init__(self, mesh, element):
        self.mesh, self.element = mesh, element

    def eval_cell(self, values, x, ufl_cell):
        n = Cell(self.mesh, ufl_cell.index).normal(ufl_cell.local_facet)
        g1 = 0.0
        for ii in range(0, 20):
            g1 += 0.1 * abs(np.sin(ii * x[0] * np.pi))
        values[0] = -g1 * n[0]
        values[1] = -g1 * n[1]

    def value_shape(self):
        return (2,)


def top(x):
    return x[1] > ymax - DOLFIN_EPS


def run(mesh_density=24, nsteps=3, verbose=True, mesh=None):
    if mesh is None:
        mesh = generate_mesh(Rectangle(Point(xmin, ymin), Point(xmax, ymax)), mesh_density)
    pbc = PeriodicBoundary()
    V = VectorElement("Lagrange", mesh.ufl_cell(), 2)
    Q = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    Qc = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    W = dolfin.FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc]), constrained_domain=pbc)
    X = FunctionSpace(mesh, "CG", 1, constrained_domain=pbc)
    G = BoundarySource(mesh, element=V)
    bc = [DirichletBC(W.sub(0), G, top)]
    darcy = DarcyAdvection(Da=Da0, phi=phi0, Pe=Pe0, alpha=alpha0, cfl=0.1)
    sol_0 = Function(W)
    c01 = Function(X)
    c01.interpolate(Expression("0.01", degree=1))
    assign(sol_0.sub(3), c01)
    darcy.perm = Two_layer_Permeability(mesh, element=Qc, top_y=ymax, ktop=1.0, kbot=0.001)
    sol = Function(W)
    S = SourceTerm(mesh, element=Qc, top_y=ymax)
    for i in range(1, nsteps + 1):
        a, L = darcy.advection_diffusion_two_component(W, mesh, sol_0, dt0, f1=S, K=darcy.perm, gam_rho=-0.5, rho_0=0.0)
        solve(a == L, sol, bc)
        sol_0 = sol
        u0, p0, c00, c01 = sol.split()
        if verbose:
            info("iteration %d: max c0 %.5f, min c1 %.6f" % (i, float(c00.get().max()), float(c01.get().min())))
    return sol, project(darcy.perm, X)


if __name__ == "__main__":
    a = sys.argv
    run(int(a[1]) if len(a) > 1 else 24, int(a[2]) if len(a) > 2 else 3)
