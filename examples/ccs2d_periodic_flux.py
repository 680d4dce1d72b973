"""Periodic CCS reactive transport in 2-D through the drop-in API, call for call what the reference's production script does
(/root/reference/src/run/CCS2D/CCS_3comp.py):

    mshr-style `generate_mesh(Rectangle(...), density)`                         :103-104
    `Plane(SubDomain)` + `FacetFunction` + `Measure("ds")[facets]` + `FacetNormal`  :112-119
    `PeriodicBoundary(SubDomain)` with inside/map  ->  `constrained_domain=pbc`     :127-143, 212-213
    Expression subclasses with `eval` (source term, layered permeability)          :146-183
    `W = dolfin.FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc, Qc]), constrained_domain=pbc)`, DirichletBC on W.sub(0) / W.sub(2)
    `assign(sol_0.sub(3), c01)`, `CCS(...)`, the time loop `a, L = ...three_component(...); solve(a == L, sol, bc); sol_0 = sol`
    `u0, p0, c00, c01, c02 = sol.split()`, `File << field`, `flux1 = assemble(c00*phi0*dot(u0, n)*ds(1))`                  :257-283
    `perm1 = project(darcy.perm, X)`                                                                                      :292-295

The periodic space is solved by mode F (P1 pressure operator + cellwise velocity recovery).
Usage: python examples/ccs2d_periodic_flux.py [mesh_density] [steps] [outdir]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403,E402
from mupopp_b200.mupopp import *   # noqa: F401,F403,E402

Da0, Pe0, alpha0, phi0, dt0 = 0.1, 5.0e2, 1.0e-8, 0.05, 1.0e-1
xmin, xmax, ymin, ymax = 0.0, 8.0, 0.0, 2.0


class Plane(SubDomain):
    def inside(self, x, on_boundary):
        return x[1] < DOLFIN_EPS


class PeriodicBoundary(SubDomain):
    def inside(self, x, on_boundary):          # the left edge is the target domain
        return bool(near(x[0], xmin) and on_boundary)

    def map(self, x, y):                        # the right edge maps onto it
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


def top(x):
    return x[1] > ymax - DOLFIN_EPS


def run(mesh_density=24, nsteps=3, outdir=None, verbose=True, mesh=None):
    if mesh is None:
        mesh = generate_mesh(Rectangle(Point(xmin, ymin), Point(xmax, ymax)), mesh_density)
    facets = FacetFunction("size_t", mesh)
    Plane().mark(facets, 1)
    ds = Measure("ds")[facets]
    n = FacetNormal(mesh)
    pbc = PeriodicBoundary()
    V = VectorElement("Lagrange", mesh.ufl_cell(), 2)
    Q = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    Qc = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    W = dolfin.FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc, Qc]), constrained_domain=pbc)
    X = FunctionSpace(mesh, "CG", 1, constrained_domain=pbc)
    bc = [DirichletBC(W.sub(0), Constant((0.0, -0.5)), top), DirichletBC(W.sub(2), Constant(0.1), top)]
    darcy = CCS(Da=Da0, phi=phi0, Pe=Pe0, alpha=alpha0, cfl=0.1)
    sol_0 = Function(W)
    c01 = Function(X)
    c01.interpolate(Expression("0.1", degree=1))
    assign(sol_0.sub(3), c01)
    darcy.perm = Two_layer_Permeability(mesh, element=Qc, top_y=ymax, ktop=1.0, kbot=0.001)
    sol = Function(W)
    S = SourceTerm(mesh, element=Qc, top_y=ymax)
    out = {k: File(os.path.join(outdir, k + ".pvd"), "compressed") for k in ("velocity", "pressure", "c0", "c1", "c2")} if outdir else None
    t, flux, times = dt0, [], []
    for i in range(1, nsteps + 1):
        a, L = darcy.advection_diffusion_three_component(W, mesh, sol_0, dt0, f_source=S, K=1.0)
        solve(a == L, sol, bc)
        sol_0 = sol
        u0, p0, c00, c01, c02 = sol.split()
        if out:
            for f, nm, key in ((u0, "velocity", "velocity"), (p0, "pressure", "pressure"), (c00, "[H2CO3]", "c0"), (c01, "[An]", "c1"),
                               (c02, "[CaCO3]", "c2")):
                f.rename(nm, "")
                out[key] << f
        flux.append(assemble(c00 * phi0 * dot(u0, n) * ds(1)))
        times.append(t)
        if verbose:
            info("time t =%g, iteration %d, bottom flux %.6e" % (t, i, flux[-1]))
        t += dt0
    perm1 = project(darcy.perm, X)
    perm1.rename("K*", "")
    return sol, np.array(times), np.array(flux), perm1


if __name__ == "__main__":
    a = sys.argv
    run(int(a[1]) if len(a) > 1 else 24, int(a[2]) if len(a) > 2 else 3, a[3] if len(a) > 3 else None)
