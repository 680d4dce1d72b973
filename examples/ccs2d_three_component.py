"""CCS reactive three-component transport in 2-D through the drop-in API: the time loop of
/root/reference/src/run/CCS2D/CCS_3comp.py:206-279 (mixed [P2^2, P1, P1, P1, P1] space, inflow and c0 = 0.1 on top, An(0) = 0.1, sine-series
source below the top boundary) on a structured mesh with natural side walls (the periodic `constrained_domain` of the script is available in
mode F only).   Usage: python examples/ccs2d_three_component.py [nx] [ny] [steps] [outdir]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403
from mupopp_b200.mupopp import *   # noqa: F401,F403

Da, Pe, alpha0, phi0, dt = 0.1, 500.0, 1.0e-8, 0.05, 0.1      # CCS_3comp.py:29-39
xmax, ymax = 8.0, 2.0


class TopSource(Expression):
    """Sum of |sin(ii pi x)| confined to a thin layer under the top boundary (CCS_3comp.py:146-164)."""

    def __init__(self, top_y):
        self.top = top_y

    def eval(self, values, x):
        g = 0.0
        for ii in range(20):
            g += 0.1 * abs(np.sin(ii * x[0] * np.pi))
        values[0] = g * (1.0 - np.tanh((self.top - x[1]) / 0.01))


def run(nx=16, ny=8, nsteps=3, outdir=None, verbose=True):
    mesh = RectangleMesh(Point(0.0, 0.0), Point(xmax, ymax), nx, ny)
    cell = mesh.ufl_cell()
    Qc = FiniteElement("Lagrange", cell, 1)
    W = FunctionSpace(mesh, MixedElement([VectorElement("Lagrange", cell, 2), FiniteElement("Lagrange", cell, 1), Qc, Qc, Qc]))

    def top(x):
        return x[1] > ymax - DOLFIN_EPS

    bc = [DirichletBC(W.sub(0), Constant((0.0, -0.5)), top), DirichletBC(W.sub(2), Constant(0.1), top)]
    darcy = CCS(Pe=Pe, Da=Da, phi=phi0, alpha=alpha0)
    S = TopSource(ymax)
    sol_0, sol = Function(W), Function(W)
    sol_0.sub(3).blocks[0].fill_(0.1)                       # initial anorthite fraction (CCS_3comp.py:233-237)
    c_out = File(os.path.join(outdir, "c0.pvd"), "compressed") if outdir else None
    t, hist = 0.0, []
    for i in range(nsteps):
        a, L = darcy.advection_diffusion_three_component(W, mesh, sol_0, dt, f_source=S, K=1.0)
        solve(a == L, sol, bc)
        sol_0 = sol
        u0, p0, c00, c01, c02 = sol.split()
        hist.append((float(c00.get().max()), float(c01.get().min()), float(c02.get().max())))
        if c_out:
            c00.rename("c0", "")
            c_out << c00
        t += dt
    if verbose:
        print("ccs2d: %dx%d, %d steps: (max c0, min c1, max c2) per step = %s" % (nx, ny, nsteps, ["%.4f %.5f %.2e" % h for h in hist]))
    return sol


if __name__ == "__main__":
    a = sys.argv
    run(int(a[1]) if len(a) > 1 else 16, int(a[2]) if len(a) > 2 else 8, int(a[3]) if len(a) > 3 else 3, a[4] if len(a) > 4 else None)
