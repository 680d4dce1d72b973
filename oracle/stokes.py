"""Oracle restatement of StokesAdvection.stokes_ADR_precipitation, NumPy/SciPy fp64.
TEST INFRASTRUCTURE -- see oracle/__init__.py.

/root/reference/src/modules/mupopp.py:997-1099 on W = [P3^d, P1, P1, P1, P1] (run/stokes_ADR_3D/master_advective_stokes.py:94-102):
    F = grad(u):grad(v) + div(v) p + q div(u)
      + eta (c0 - c0n) + dt [eta un.grad(cm) + grad(eta).grad(cm)/Pe] + dt f11 eta + omega (c1 - c1n) + dt f21 omega + omega1 (c2 - c2n) - dt f31 omega1
      + tau (un.grad eta) r,   r = c0 - c0n + dt (un.grad(cm) + f11) + c1 - c1n + dt f21 + c2 - c2n - dt f31
    f11, f21, f31 = (62, 278, 100)/698 * Da * c0n * c1n   -- NOT divided by phi / (1-phi) (there is no porosity in this class)
The Stokes block does not couple to the concentrations; the transport rows use the previous step's velocity (mupopp.py:1049).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .forms import QDEG, Assembler, State, TransportParams

PHI_FOLD = 0.5     # the generic transport rows divide R1 by phi and R2,R3 by (1-phi); fold that into the fractions


def stokes_adr_params(Pe=100, Da=10.0, dt=0.01, SUPG=2):
    T = 698.0
    return TransportParams(Pe=Pe, Da=Da, phi=PHI_FOLD, alpha=0.0, dt=dt, K=1.0, sigma=0.0, rho0=0.0, gamma=0.0, reaction="c0c1",
                           fracs=(62.0 / T * PHI_FOLD, 278.0 / T * (1 - PHI_FOLD), 100.0 / T * (1 - PHI_FOLD)), ncomp=3, SUPG=SUPG,
                           zhat=(0.0, 0.0, 0.0))


def stokes_blocks(asm: Assembler, vdeg=3):
    """Vector-Laplacian block (one scalar stiffness per component) and divergence blocks B_a[j,i] = int q_j d_a psi_i."""
    L = asm.stiffness(vdeg, 2 * (vdeg - 1))
    B = asm.div_blocks(1, vdeg, vdeg)
    return L, B


def monolithic_step(asm: Assembler, prm: TransportParams, prev: State, u_dofs, u_vals, c0_dofs, c0_vals, vdeg=3, qdeg=None):
    """One reference timestep: the full [u, p, c0, c1, c2] matrix, symmetric Dirichlet elimination, LU (master_advective_stokes.py:208-218)."""
    qdeg = qdeg or QDEG["c0row"]
    d = asm.mesh.dim
    nv, n1 = asm.space(vdeg).ndof, asm.P1.ndof
    L, B = stokes_blocks(asm, vdeg)
    A00, S, qd = asm.c0_row(prm, ("P%d" % vdeg, prev.u), qdeg)
    M1 = asm.mass(1, 1, QDEG["mass_p1"])
    fsrc = np.zeros(n1)
    b_c0 = asm.c0_rhs(prm, qd, prev.c0, prev.c1, prev.c2, fsrc, qdeg)
    R1, R2, R3 = asm.reaction_q(prm, prev.c0, prev.c1, QDEG["react"])
    b_c1 = M1 @ prev.c1 - prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R2)
    b_c2 = M1 @ prev.c2 + prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R3)
    nblk = d + 4
    rows = [[None] * nblk for _ in range(nblk)]
    for a in range(d):
        rows[a][a] = L
        rows[a][d] = B[a].T
        rows[d][a] = B[a]
    rows[d][d] = sp.csr_matrix((n1, n1))
    rows[d + 1][d + 1] = A00
    rows[d + 1][d + 2] = S
    rows[d + 1][d + 3] = S
    rows[d + 2][d + 2] = M1
    rows[d + 3][d + 3] = M1
    A = sp.bmat(rows, format="csr")
    b = np.concatenate([np.zeros(nv)] * d + [np.zeros(n1), b_c0, b_c1, b_c2])
    dofs = np.concatenate([a * nv + np.asarray(u_dofs[a], dtype=np.int64) for a in range(d)] + [d * nv + n1 + np.asarray(c0_dofs, dtype=np.int64)])
    vals = np.concatenate([np.asarray(u_vals[a], dtype=float) for a in range(d)] + [np.asarray(c0_vals, dtype=float)])
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, dofs, vals)
    x = spla.splu(Abc.tocsc()).solve(bbc)
    o = d * nv
    return State(u=[x[a * nv:(a + 1) * nv] for a in range(d)], p=x[o:o + n1], c0=x[o + n1:o + 2 * n1], c1=x[o + 2 * n1:o + 3 * n1],
                 c2=x[o + 3 * n1:o + 4 * n1])
