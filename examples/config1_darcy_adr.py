"""BASELINE.json configs[0]: DarcyAdvection on UnitSquareMesh(64,64), mixed P2/P1 Darcy solve once, then 10 SUPG
advection-diffusion steps -- written against the drop-in API (`mupopp_b200.fenics` + `mupopp_b200.mupopp`).

Same problem as /root/reference/src/run/darcy_advection_diffusion/darcy_advection_diffusion.py (boundary data :25-36,69-70,
parameters :95-106), structured mesh instead of mshr, 10 steps instead of 200.   Usage: python examples/config1_darcy_adr.py [n] [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403
from mupopp_b200.mupopp import *   # noqa: F401,F403


class InflowProfile(Expression):
    """u = g(x) n on the boundary facet, g = -sin(2 pi x): needs the facet normal, hence eval_cell."""

    def __init__(self, mesh):
        self.mesh = mesh

    def eval_cell(self, values, x, ufc_cell):
        n = Cell(self.mesh, ufc_cell.index).normal(ufc_cell.local_facet)
        g = -np.sin(2.0 * np.pi * x[0])
        values[0], values[1] = g * n[0], g * n[1]

    def value_shape(self):
        return (2,)


def run(n=64, nsteps=10, dt=0.1, verbose=True):
    ymax = 1.0
    mesh = UnitSquareMesh(n, n)

    def on_bottom(x):
        return x[1] < DOLFIN_EPS

    def on_top(x):
        return x[1] > ymax - DOLFIN_EPS

    W = FunctionSpace(mesh, MixedElement([VectorElement("Lagrange", mesh.ufl_cell(), 2), FiniteElement("Lagrange", mesh.ufl_cell(), 1)]))
    bcs_u = [DirichletBC(W.sub(0), InflowProfile(mesh), on_bottom), DirichletBC(W.sub(0), Constant((0.0, 1.0)), on_top)]
    model = DarcyAdvection()                       # Pe=100, Da=10, phi=0.01
    a, L = model.darcy_bilinear(W, mesh)
    up = Function(W)
    outer = solve(a == L, up, bcs_u)
    u, p = up.split()
    Q = FunctionSpace(mesh, "CG", 1)
    vel = Function(VectorFunctionSpace(mesh, "CG", 2))
    vel.interpolate(u)
    c_prev = Function(Q)
    c_prev.interpolate(Expression("0.1", degree=1))
    bcs_c = [DirichletBC(Q, Constant(0.0), on_top), DirichletBC(Q, Constant(1.0), on_bottom)]
    c_new = Function(Q)
    its = []
    for k in range(nsteps):
        a1, L1 = model.advection_diffusion(Q, c_prev, vel, dt, mesh)
        its.append(solve(a1 == L1, c_new, bcs_c))
        c_prev = c_new
    if verbose:
        c = c_new.get()
        print("config1: n=%d Schur-PCG outer iterations %d; ADR GMRES iterations %s; c in [%.4f, %.4f], mean %.6f"
              % (n, outer, its, c.min(), c.max(), c.mean()))
    return up, c_new


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 10)
