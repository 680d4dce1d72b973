"""Oracle restatement of the `compaction` forms, NumPy/SciPy fp64.
TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows /root/reference/src/modules/mupopp.py:
  * compaction.surface_tension_2     :206-215
  * compaction.momentum_conservation :306-382   (matrix A :361-366, rhs L :374-375, preconditioner b :379-380)
  * compaction.momentum_solver       :383-429   (assemble_system(a,L,bcs) + MINRES/AMG rtol 1e-6; the oracle solves by LU)
  * compaction.mass_conservation     :217-259   (CN porosity advection + keff SUPG), mass_solver :261-304
Unknown layout [u_0..u_{d-1} (P2), p (P1), c (P1)]; phi, gam, buoyancy are P1 nodal fields.
Reproduced literally: `dL*phi*phi**m3` = dL*phi*phi^(1/phi^3) in A (underflows to ~0), theta passed raw to cos.
All coefficient-bearing integrals use ONE quadrature rule of degree QDEG_COMPACTION (the integrands are non-polynomial in phi).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .forms import Assembler

QDEG_COMPACTION = 8


@dataclass
class CompactionParams:
    da: float = 1e-4
    R: float = 0.2
    B: float = 1e6
    theta: float = 10.0
    dL: float = 1.0
    zhat: tuple = (0.0, 0.0, 1.0)


def chi(prm, phi):
    p1, p2, p3, p4 = -8065.0, 6149.0, -1778.0, 249.0
    return (2.0 * np.cos(prm.theta) - 1.0) * (2.0 * p4 + 20.0 * p1 * phi ** 3 + 12.0 * p2 * phi ** 2 + 6.0 * p3 * phi)


def eta_of(phi):
    return 3.0 * phi / (2.0 * (1.0 - phi) * (2.0 - phi))         # eta = 1/alpha, alpha = 2(1-phi)(2-phi)/(3 phi)


def pp_coef_A(prm, phi):
    with np.errstate(under="ignore", over="ignore"):
        return prm.dL * phi * phi ** (1.0 / phi ** 3)           # mupopp.py:365 (operator precedence quirk)


def momentum_blocks(asm: Assembler, prm: CompactionParams, phi, gam, buoy, qdeg=QDEG_COMPACTION):
    """Block operators and right-hand sides of the momentum form, and the preconditioner blocks."""
    m = asm.mesh
    d = m.dim
    w, p1, dref1 = asm.tab(1, qdeg)
    _, p2, dref2 = asm.tab(2, qdeg)
    g1 = fem.phys_grads(dref1, asm.Jinv)                      # (C,q,b1,d)
    g2 = fem.phys_grads(dref2, asm.Jinv)                      # (C,q,b2,d)
    wd = w[None, :] * asm.detJ[:, None]
    s1, s2 = asm.P1, asm.P2
    phq = fem.eval_at_quad(s1, phi, p1)
    gphi = fem.grad_at_quad(s1, phi, g1)                      # (C,q,d)
    gaq = fem.eval_at_quad(s1, gam, p1)
    ggam = fem.grad_at_quad(s1, gam, g1)
    buq = fem.eval_at_quad(s1, buoy, p1)
    n1, n2 = s1.ndof, s2.ndof
    # (v,u): 0.5 (1-phi) symgrad u : symgrad v = 0.25 (1-phi) [delta_ab grad psi_i.grad psi_j + d_b psi_i d_a psi_j]
    cf = 0.25 * wd * (1.0 - phq)
    lap = np.einsum("cq,cqik,cqjk->cij", cf, g2, g2, optimize=True)
    Auu = [[None] * d for _ in range(d)]
    for a in range(d):
        for b in range(d):
            Ae = np.einsum("cq,cqi,cqj->cij", cf, g2[..., b], g2[..., a], optimize=True)
            if a == b:
                Ae = Ae + lap
            Auu[a][b] = fem.scatter_matrix(Ae, s2.cell_dofs, s2.cell_dofs, (n2, n2))
    Bdiv = asm.div_blocks(1, 2, 2)                              # int q d_a psi
    App = fem.scatter_matrix(np.einsum("cq,cq,cqia,cqja->cij", wd, pp_coef_A(prm, phq), g1, g1, optimize=True),
                             s1.cell_dofs, s1.cell_dofs, (n1, n1))
    Acc = fem.scatter_matrix(np.einsum("cq,cq,qi,qj->cij", wd, -eta_of(phq), p1, p1, optimize=True), s1.cell_dofs, s1.cell_dofs, (n1, n1))
    # rhs: h = dL (1-phi) (chi grad(phi)/B - R buoy zhat) + da grad(4 gam/(3 phi))
    zh = np.asarray(prm.zhat, dtype=float)[:d]
    h = prm.dL * (1.0 - phq)[..., None] * (chi(prm, phq)[..., None] * gphi / prm.B - prm.R * buq[..., None] * zh[None, None, :]) \
        + prm.da * (4.0 / 3.0) * (ggam / phq[..., None] - gaq[..., None] * gphi / (phq ** 2)[..., None])
    bu = [fem.scatter_vector(-np.einsum("cq,cq,qi->ci", wd, h[..., a], p2, optimize=True), s2.cell_dofs, n2) for a in range(d)]
    fq = prm.da * prm.R * buq * gaq / (1.0 - prm.R * buq)
    bp = fem.scatter_vector(-np.einsum("cq,cq,qi->ci", wd, fq, p1, optimize=True), s1.cell_dofs, n1)
    # preconditioner (mupopp.py:379-380): (1-phi) grad u:grad v ; dL phi^2 m3 grad p.grad q + p q ; 0.5 eta c omega
    Puu = fem.scatter_matrix(np.einsum("cq,cqik,cqjk->cij", wd * (1.0 - phq), g2, g2, optimize=True), s2.cell_dofs, s2.cell_dofs, (n2, n2))
    Ppp = fem.scatter_matrix(np.einsum("cq,cq,cqia,cqja->cij", wd, prm.dL / phq, g1, g1, optimize=True)
                             + np.einsum("cq,qi,qj->cij", wd, p1, p1, optimize=True), s1.cell_dofs, s1.cell_dofs, (n1, n1))
    Pcc = fem.scatter_matrix(np.einsum("cq,cq,qi,qj->cij", wd, 0.5 * eta_of(phq), p1, p1, optimize=True), s1.cell_dofs, s1.cell_dofs, (n1, n1))
    return dict(Auu=Auu, Bdiv=Bdiv, App=App, Acc=Acc, bu=bu, bp=bp, Puu=Puu, Ppp=Ppp, Pcc=Pcc)


def momentum_system(asm, prm, phi, gam, buoy, qdeg=QDEG_COMPACTION):
    blk = momentum_blocks(asm, prm, phi, gam, buoy, qdeg)
    d = asm.mesh.dim
    n1 = asm.P1.ndof
    rows = []
    for a in range(d):
        rows.append([blk["Auu"][a][b] for b in range(d)] + [-blk["Bdiv"][a].T, -blk["Bdiv"][a].T])
    rows.append([-blk["Bdiv"][b] for b in range(d)] + [blk["App"], None])
    rows.append([-blk["Bdiv"][b] for b in range(d)] + [None, blk["Acc"]])
    A = sp.bmat(rows, format="csr")
    b = np.concatenate(list(blk["bu"]) + [blk["bp"], np.zeros(n1)])
    return A, b, blk


def momentum_solve(asm, prm, phi, gam, buoy, u_dofs, u_vals, qdeg=QDEG_COMPACTION):
    """assemble_system(a, L, bcs) + exact solve (the reference: MINRES/AMG to rtol 1e-6, mupopp.py:410-425).
    u_dofs/u_vals: per component arrays of P2 dofs / values.  Returns (u list, p, c)."""
    A, b, _ = momentum_system(asm, prm, phi, gam, buoy, qdeg)
    d = asm.mesh.dim
    n1, n2 = asm.P1.ndof, asm.P2.ndof
    dofs = np.concatenate([a * n2 + np.asarray(u_dofs[a], dtype=np.int64) for a in range(d)])
    vals = np.concatenate([np.asarray(u_vals[a], dtype=float) for a in range(d)])
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, dofs, vals)
    x = spla.splu(Abc.tocsc()).solve(bbc)
    return [x[a * n2:(a + 1) * n2] for a in range(d)], x[d * n2:d * n2 + n1], x[d * n2 + n1:]


def p2_basis_laplacians(asm: Assembler):
    """Laplacian of every P2 basis function on every (affine) cell, constant per cell: vertex i -> 4 |grad lam_i|^2, edge (a, b) ->
    8 grad lam_a . grad lam_b (UFC edge order).  Returns (C, nb2)."""
    _, _, dref1 = asm.tab(1, 1)
    g = fem.phys_grads(dref1, asm.Jinv)[:, 0]                                   # (C, d+1, d) gradients of the barycentric coordinates
    d = asm.mesh.dim
    le = fem.TRI_EDGES if d == 2 else fem.TET_EDGES
    vert = 4.0 * np.einsum("cia,cia->ci", g, g)
    edge = np.stack([8.0 * np.einsum("ca,ca->c", g[:, a], g[:, b]) for (a, b) in le], 1)
    return np.concatenate([vert, edge], 1)


def porosity_system(asm: Assembler, da, dt, phi0, u, gam, qdeg=QDEG_COMPACTION, pdeg=1):
    """mupopp.py:217-259 with a P1 (pdeg=1) or P2 (pdeg=2) porosity space X and a P2 velocity u (list of component dof arrays):
       F = w (phi1 - phi0 + dt (u.grad(phi_mid) - (1 - phi_mid) div u) - dt da gam)
           + (keff/|u|^2) (u.grad w) (phi1 - phi0 + dt (u.grad(phi_mid) - div(grad(phi_mid)) - da gam)),   keff = max(h|u|/2 - 1, 0)
    (div grad of a P1 field vanishes; for P2 it is constant per cell).  Where keff == 0 the stabilisation is dropped (the reference would
    evaluate 0/0 for u == 0).  phi0 and gam live in X."""
    w, p1, dref1 = asm.tab(pdeg, qdeg)
    _, p2, dref2 = asm.tab(2, qdeg)
    g1 = fem.phys_grads(dref1, asm.Jinv)
    g2 = fem.phys_grads(dref2, asm.Jinv)
    wd = w[None, :] * asm.detJ[:, None]
    s1, s2 = asm.space(pdeg), asm.P2
    d = asm.mesh.dim
    uq = np.stack([fem.eval_at_quad(s2, np.asarray(u[a]), p2) for a in range(d)], -1)           # (C,q,d)
    divu = sum(np.einsum("cb,cqb->cq", np.asarray(u[a])[s2.cell_dofs], g2[..., a], optimize=True) for a in range(d))
    un2 = (uq * uq).sum(-1)
    un = np.sqrt(un2)
    aval = 0.5 * asm.h[:, None] * un
    keff = np.maximum(aval - 1.0, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        st = np.where(keff > 0.0, keff / un2, 0.0)
    ug = np.einsum("cqd,cqbd->cqb", uq, g1, optimize=True)                                        # u.grad(psi_b)
    lap = p2_basis_laplacians(asm) if pdeg == 2 else np.zeros((asm.mesh.nc, p1.shape[1]))        # (C, nb) Laplacian of psi_b
    ph0 = fem.eval_at_quad(s1, phi0, p1)
    ugp0 = np.einsum("cqb,cb->cq", ug, phi0[s1.cell_dofs], optimize=True)
    lap0 = np.einsum("cb,cb->c", lap, phi0[s1.cell_dofs])[:, None]
    gaq = fem.eval_at_quad(s1, gam, p1)
    Ae = np.einsum("cq,qi,qj->cij", wd, p1, p1, optimize=True)
    Ae += 0.5 * dt * np.einsum("cq,qi,cqj->cij", wd, p1, ug, optimize=True)
    Ae += 0.5 * dt * np.einsum("cq,cq,qi,qj->cij", wd, divu, p1, p1, optimize=True)
    Ae += np.einsum("cq,cq,cqi,qj->cij", wd, st, ug, p1, optimize=True)
    Ae += 0.5 * dt * np.einsum("cq,cq,cqi,cqj->cij", wd, st, ug, ug - lap[:, None, :], optimize=True)
    A = fem.scatter_matrix(Ae, s1.cell_dofs, s1.cell_dofs, (s1.ndof, s1.ndof))
    gal = ph0 - 0.5 * dt * ugp0 + dt * divu - 0.5 * dt * ph0 * divu + dt * da * gaq
    rk = -ph0 + 0.5 * dt * (ugp0 - lap0) - dt * da * gaq
    be = np.einsum("cq,cq,qi->ci", wd, gal, p1, optimize=True) - np.einsum("cq,cq,cqi,cq->ci", wd, st, ug, rk, optimize=True)
    b = fem.scatter_vector(be, s1.cell_dofs, s1.ndof)
    return A, b
