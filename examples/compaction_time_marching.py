"""BASELINE.json configs[3] in miniature: compaction momentum solves on a box mesh while an imposed Gaussian porosity pulse travels
upwards, followed by one CN-SUPG porosity update -- the flow of /root/reference/src/benchmark/time_marching/time_test1.py:80-262
written against the drop-in API.   Usage: python examples/compaction_time_marching.py [n] [steps] [outdir]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403
from mupopp_b200.mupopp import *   # noqa: F401,F403


def run(n=4, nsteps=2, outdir=None, tol=1.0e-6, verbose=True):
    cfg = os.path.join(tempfile.mkdtemp(), "run.cfg")
    with open(cfg, "w") as f:      # same keys as benchmark/time_marching/time_test1.cfg
        f.write("out_freq = 1\nT = 1.0\ndt = 0.1\nda = 0.0001\nR = 0.2\nB = 1e6\ntheta = 10\ndL = 1.0\n")
    model = compaction(cfg)
    mesh = BoxMesh(Point(-2.0, -2.0, 0.0), Point(2.0, 2.0, 8.0), n, n, 2 * n)
    cell = mesh.ufl_cell()
    W = FunctionSpace(mesh, MixedElement([VectorElement("Lagrange", cell, 2), FiniteElement("Lagrange", cell, 1), FiniteElement("Lagrange", cell, 1)]))
    X = FunctionSpace(mesh, "CG", 1)

    def side_walls(x, on_boundary):
        return on_boundary and (near(abs(x[0]), 2.0, 1e-12) or near(abs(x[1]), 2.0, 1e-12))

    bcs = [DirichletBC(W.sub(0), Constant((0.0, 0.0, 0.0)), side_walls)]
    phi, gam, buoyancy = Function(X), Function(X), Function(X)
    gam.interpolate(Expression("0.0", degree=1))
    buoyancy.interpolate(Expression("1.0", degree=1))
    out = None
    if outdir:
        out = {k: File(os.path.join(outdir, k + ".pvd"), "compressed") for k in ("velocity", "pressure", "compaction", "porosity")}
    t, its = 0.0, []
    for step in range(nsteps):
        pulse = Expression("0.15*(exp(-x[0]*x[0]-x[1]*x[1]-(x[2]-0.5*time)*(x[2]-0.5*time))/0.5)+0.1", degree=2, time=t)
        phi.interpolate(pulse)
        a, L, b = model.momentum_conservation(W, phi, gam, buoyancy)
        u, p, c, niter = model.momentum_solver(W, a, L, b, bcs, tol=tol)
        its.append(niter)
        if out:
            for name, f in (("velocity", u), ("pressure", p), ("compaction", c), ("porosity", phi)):
                f.rename(name, "")
                out[name] << f
        t += model.dt
    a_phi, L_phi, bb = model.mass_conservation(X, phi, u, model.dt, gam, mesh)
    phi_new = model.mass_solver(X, a_phi, L_phi, bb)
    if verbose:
        w = u.get(2)
        print("compaction: n=%d MINRES iterations %s; max |u_z| %.4e; porosity change %.3e" % (n, its, np.abs(w).max(), np.abs(phi_new.get() - phi.get()).max()))
    return u, p, c, phi_new


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 2, sys.argv[3] if len(sys.argv) > 3 else None)
