"""Stokes + ADR + precipitation on the reference's own pore-scale sample mesh, through the drop-in API -- the calls of
/root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:

    mesh = Mesh(); hdf = HDF5File(mesh.mpi_comm(), "hdf5_sample_mesh.xml.h5", "r"); hdf.read(mesh, "/mesh", False)            :77-80
    boundaries = MeshFunction("size_t", mesh, mesh.topology().dim()-1); hdf.read(boundaries, "/boundaries")                    :84-85
    ds = Measure("ds", domain=mesh, subdomain_data=boundaries); n = FacetNormal(mesh)                                          :88-89
    W = dolfin.FunctionSpace(mesh, MixedElement([V(P3), Q, Qc, Qc, Qc])); DirichletBC(W.sub(0), Constant(..), boundaries, id)  :94-130
    a, L = stokes.stokes_ADR_precipitation(W, mesh, sol_prev=sol_initial, dt=dt0); solve(a == L, sol, bc, solver_parameters=..) :208-213
    iflux1 = assemble(c00*dot(u, unitNormal)*ds(1)); assemble(c02*dx); assemble(CaCO3_frac*Da0*c00*c01*dx); assemble(temp4*dx)  :232-265

Without the HDF5 file (the GPU box has no /root/reference) the same mesh and markers come from tests/golden/stokes_sample_mesh.npz.
Usage: python examples/stokes_adr_refmesh.py [steps] [mesh.h5]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mupopp_b200.fenics import *   # noqa: F401,F403,E402
from mupopp_b200.mupopp import *   # noqa: F401,F403,E402

Da0, Pe0, dt0 = 0.1, 0.1, 100.0                # master_advective_stokes.py:22-30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_mesh(h5=None):
    if h5 is not None and os.path.exists(h5):
        mesh = Mesh()
        hdf = HDF5File(mesh.mpi_comm(), h5, "r")
        hdf.read(mesh, "/mesh", False)
        boundaries = MeshFunction("size_t", mesh, mesh.topology().dim() - 1)
        hdf.read(boundaries, "/boundaries")
        return mesh, boundaries
    z = np.load(os.path.join(ROOT, "tests", "golden", "stokes_sample_mesh.npz"))
    mesh = Mesh(z["coords"], z["cells"])
    return mesh, MeshFunction("size_t", mesh, z["facet_markers"].astype(np.int64))


def run(nsteps=2, h5=None, verbose=True):
    mesh, boundaries = load_mesh(h5)
    ds = Measure("ds", domain=mesh, subdomain_data=boundaries)
    V = VectorElement("Lagrange", mesh.ufl_cell(), 3)
    Q = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    Qc = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    W = dolfin.FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc, Qc]))
    X = FunctionSpace(mesh, "CG", 1)
    bc = [DirichletBC(W.sub(0), Constant((1.0, 0.0, 0.0)), boundaries, 1), DirichletBC(W.sub(0), Constant((0.0, 0.0, 0.0)), boundaries, 3),
          DirichletBC(W.sub(2), Constant(0.1), boundaries, 1)]
    sol_initial, sol = Function(W), Function(W)
    c01 = Function(X)
    c01.interpolate(Expression("0.1", degree=1))
    assign(sol_initial.sub(3), c01)                      # initial anorthite mass fraction
    temp4 = Function(X)
    temp4.interpolate(Expression("1.0", degree=1))
    unitNormal = Constant((1.0, 0.0, 0.0))
    stokes = StokesAdvection(Da=Da0, Pe=Pe0)
    out = []
    for i in range(1, nsteps + 1):
        a, L = stokes.stokes_ADR_precipitation(W, mesh, sol_prev=sol_initial, dt=dt0)
        solve(a == L, sol, bc, solver_parameters={"linear_solver": "mumps"})
        sol_initial = sol
        u, p, c00, c01, c02 = sol.split()
        row = dict(iflux=assemble(c00 * dot(u, unitNormal) * ds(1)), oflux=assemble(c00 * dot(u, unitNormal) * ds(2)),
                   volint=assemble(c02 * dx), rate=assemble(100.0 / 698.0 * Da0 * c00 * c01 * dx), volume=assemble(temp4 * dx))
        out.append(row)
        if verbose:
            info("iteration %d: influx %.6e outflux %.6e precipitate %.6e rate %.6e volume %.6f" % (
                i, row["iflux"], row["oflux"], row["volint"], row["rate"], row["volume"]))
    return sol, out


if __name__ == "__main__":
    a = sys.argv
    run(int(a[1]) if len(a) > 1 else 2, a[2] if len(a) > 2 else "/root/reference/src/run/stokes_ADR_3D/hdf5_sample_mesh.xml.h5")
