"""Oracle for MODE F: the primal P1 restatement named by BASELINE.json's north_star
(pressure operator div(k grad p) + cellwise velocity recovery + P1 transport).
TEST INFRASTRUCTURE -- see oracle/__init__.py.

Mode F is NOT the reference's discretisation (the reference solves the mixed
P2/P1 system, mupopp.py:804-807); it is the Schur/primal form of the same PDEs
(SURVEY.md appendix A.3b).  Definition used by both this oracle and the GPU path:

  u_cell   = -(K_cell/phi) (grad p|_cell + sigma * rho(c0bar_cell) * zhat)     DG0 velocity
             K_cell = mean of nodal K (or the float), c0bar = mean of the cell's nodal c0
  pressure : sum_cells |cell| grad(q) . u_cell = int_{Gamma_u} (u_D.n) q ds,  p = 0 on the rest of the boundary
             <=> int (K/phi) grad p.grad q = -sigma int (K/phi) rho zhat.grad q - int_{Gamma_u} (u_D.n) q
  transport: rows c1, c2, c0 exactly as mupopp.py:823-848 / :1231-1258 with u^n := u_cell^n.
All transport integrands are polynomial for a DG0 velocity, so quadrature degree 4 is exact.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .forms import Assembler, TransportParams, QDEG

QDEG_F = 4


@dataclass
class StateF:
    u: np.ndarray        # (C, d) DG0 velocity
    p: np.ndarray
    c0: np.ndarray
    c1: np.ndarray
    c2: np.ndarray = None


@dataclass
class BCsF:
    p_dofs: np.ndarray                      # vertices with p = 0 (natural boundary of the mixed form)
    flux_load: np.ndarray                   # int_{Gamma_u} (u_D.n) q_i ds, length V
    c0_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    c0_vals: np.ndarray = field(default_factory=lambda: np.zeros(0))


def boundary_facets(mesh):
    """Boundary facets (appear in exactly one cell): (F, d) vertex ids + owning cell + opposite vertex."""
    d = mesh.dim
    cells = mesh.cells
    loc = [[j for j in range(d + 1) if j != k] for k in range(d + 1)]
    fv = np.concatenate([cells[:, l] for l in loc], axis=0)                 # (C*(d+1), d) ascending
    owner = np.tile(np.arange(mesh.nc), d + 1)
    opp = np.concatenate([cells[:, k] for k in range(d + 1)])
    key = fv[:, 0].astype(np.int64)
    for j in range(1, d):
        key = key * mesh.nv + fv[:, j]
    uniq, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    sel = idx[cnt == 1]
    sel.sort()
    return fv[sel], owner[sel], opp[sel]


def facet_measure_normal(mesh, fv, opp):
    x = mesh.coords[fv]                                                    # (F, d, d)
    if mesh.dim == 2:
        t = x[:, 1] - x[:, 0]
        n = np.stack([t[:, 1], -t[:, 0]], 1)
        area = np.linalg.norm(t, axis=1)
    else:
        n = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
        area = 0.5 * np.linalg.norm(n, axis=1)
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    out = ((x[:, 0] - mesh.coords[opp]) * n).sum(1)
    n[out < 0] *= -1.0
    return area, n


def make_bcs(mesh, u_pred, u_value, c0_pred=None, c0_value=0.0):
    """u_pred(x)->bool marks Gamma_u by facet (all vertices + midpoint inside, DOLFIN 'topological');
    u_value(x)->(d,) velocity there.  p = 0 on every vertex of the remaining boundary facets."""
    fv, owner, opp = boundary_facets(mesh)
    area, n = facet_measure_normal(mesh, fv, opp)
    d = mesh.dim
    x = mesh.coords[fv]                                                    # (F, d, d)
    inside = np.array([[bool(u_pred(p)) for p in f] for f in x]).all(1)
    inside &= np.array([bool(u_pred(p)) for p in x.mean(1)])
    flux = np.zeros(mesh.nv)
    mloc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
    for f in np.where(inside)[0]:
        g = np.array([np.dot(u_value(p), n[f]) for p in x[f]])
        flux[fv[f]] += area[f] * (mloc @ g)
    p_dofs = np.unique(fv[~inside].ravel())
    if c0_pred is not None:
        c0_dofs = np.array([i for i in np.unique(fv.ravel()) if c0_pred(mesh.coords[i])], dtype=np.int64)
        c0_vals = np.array([c0_value(mesh.coords[i]) if callable(c0_value) else c0_value for i in c0_dofs], dtype=float)
    else:
        c0_dofs, c0_vals = np.zeros(0, dtype=np.int64), np.zeros(0)
    return BCsF(p_dofs=p_dofs.astype(np.int64), flux_load=flux, c0_dofs=c0_dofs, c0_vals=c0_vals)


def cell_K(asm, K):
    if np.isscalar(K):
        return np.full(asm.mesh.nc, float(K))
    return np.asarray(K)[asm.mesh.cells].mean(1)


def p1_grads(asm):
    """Constant physical gradients of the P1 basis per cell: (C, d+1, d)."""
    _, _, dref = asm.tab(1, 1)
    return fem.phys_grads(dref, asm.Jinv)[:, 0]


def pressure_system(asm, prm: TransportParams, c0):
    d = asm.mesh.dim
    vol = asm.detJ / (2.0 if d == 2 else 6.0)
    kc = cell_K(asm, prm.K) / prm.phi
    g = p1_grads(asm)
    Ae = np.einsum("c,cia,cja->cij", vol * kc, g, g, optimize=True)
    s = asm.P1
    L = fem.scatter_matrix(Ae, s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))
    rho = prm.rho0 + prm.gamma * c0[asm.mesh.cells].mean(1)
    gz = np.einsum("cia,a->ci", g, np.asarray(prm.zhat, dtype=float))
    be = -prm.sigma * (vol * kc * rho)[:, None] * gz
    return L, fem.scatter_vector(be, s.cell_dofs, s.ndof)


def recover_velocity(asm, prm, p, c0):
    kc = cell_K(asm, prm.K) / prm.phi
    g = p1_grads(asm)
    gp = np.einsum("cia,ci->ca", g, p[asm.mesh.cells], optimize=True)
    rho = prm.rho0 + prm.gamma * c0[asm.mesh.cells].mean(1)
    return -kc[:, None] * (gp + prm.sigma * rho[:, None] * np.asarray(prm.zhat, dtype=float)[None, :])


def periodic_prolongation(mesh, axes, lo, hi):
    """P (V x ndof), P[v, dof(v)] = 1: vertices on the max-face of each periodic axis share the dof of their image on the
    min-face; dofs numbered compactly in master-vertex order (constrained_domain of run/CCS/CCS_3D_top_source.py:134-158)."""
    X = mesh.coords
    ext = float(np.max(np.asarray(hi) - np.asarray(lo)))
    img = X.copy()
    for a in axes:
        img[np.abs(X[:, a] - hi[a]) <= 1e-9 * ext, a] = lo[a]
    key = [tuple(np.round((p - np.asarray(lo)) / (1e-7 * ext)).astype(np.int64)) for p in img]
    first = {}
    master = np.empty(mesh.nv, dtype=np.int64)
    for v, k in enumerate(key):
        master[v] = first.setdefault(k, v)
    ism = master == np.arange(mesh.nv)
    dof = (np.cumsum(ism) - 1)[master]
    P = sp.csr_matrix((np.ones(mesh.nv), (np.arange(mesh.nv), dof)), shape=(mesh.nv, int(dof.max()) + 1))
    return P, dof


def make_bcs_periodic(mesh, axes, dof, u_pred, u_value, c0_pred=None, c0_value=0.0):
    """make_bcs on the non-periodic part of the boundary, folded to dof numbering."""
    fv, owner, opp = boundary_facets(mesh)
    area, n = facet_measure_normal(mesh, fv, opp)
    keep = np.ones(len(fv), dtype=bool)
    for a in axes:
        keep &= np.abs(n[:, a]) < 0.5
    fv, area, n = fv[keep], area[keep], n[keep]
    d = mesh.dim
    x = mesh.coords[fv]
    inside = np.array([[bool(u_pred(p)) for p in f] for f in x]).all(1) & np.array([bool(u_pred(p)) for p in x.mean(1)])
    ndof = int(dof.max()) + 1
    flux = np.zeros(ndof)
    mloc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
    for f in np.where(inside)[0]:
        g = np.array([np.dot(u_value(p), n[f]) for p in x[f]])
        np.add.at(flux, dof[fv[f]], area[f] * (mloc @ g))
    p_dofs = np.unique(dof[fv[~inside].ravel()])
    c0_dofs, c0_vals = np.zeros(0, dtype=np.int64), np.zeros(0)
    if c0_pred is not None:
        vs = np.array([i for i in np.unique(fv.ravel()) if c0_pred(mesh.coords[i])], dtype=np.int64)
        c0_dofs, idx = np.unique(dof[vs], return_index=True)
        c0_vals = np.array([c0_value(mesh.coords[i]) if callable(c0_value) else c0_value for i in vs[idx]], dtype=float)
    return BCsF(p_dofs=p_dofs.astype(np.int64), flux_load=flux, c0_dofs=c0_dofs.astype(np.int64), c0_vals=c0_vals)


class StepSystems:
    """The linear systems of ONE mode-F timestep from state `prev`, built lazily in solve order: c1, c2 (mass rows), c0 (needs the
    new c1, c2), p (needs the new c0).  `step` solves them; the at-size GPU parity tests only ASSEMBLE them and check that the GPU
    fields satisfy them (residual parity needs no CPU solve, so it reaches sizes where a CPU solve takes minutes).
    `P`: periodic prolongation (V x ndof); fields, fsrc and bcs are then in dof numbering and every operator is P^T A P."""

    def __init__(self, asm: Assembler, prm: TransportParams, prev: StateF, fsrc, bcs: BCsF, P=None):
        self.asm, self.prm, self.prev, self.fsrc, self.bcs, self.P = asm, prm, prev, fsrc, bcs, P
        self.PT = P.T.tocsr() if P is not None else None
        self.M1 = asm.mass(1, 1, QDEG["mass_p1"])
        self.M1d = self.fold(self.M1)
        self.c0v, self.c1v = self.v(prev.c0), self.v(prev.c1)
        self.c2v = self.v(prev.c2) if prm.ncomp == 3 else None
        self.R = asm.reaction_q(prm, self.c0v, self.c1v, QDEG_F)

    def v(self, x):                      # dof vector -> vertex values
        return x if self.P is None else self.P @ x

    def fold(self, A):
        return A if self.P is None else (self.PT @ A @ self.P).tocsr()

    def foldv(self, b):
        return b if self.P is None else self.PT @ b

    def c1(self):
        prm = self.prm
        return self.M1d, self.foldv(self.M1 @ self.c1v - prm.dt / (1 - prm.phi) * self.asm.load(1, QDEG_F, self.R[1]))

    def c2(self):
        prm = self.prm
        return self.M1d, self.foldv(self.M1 @ self.c2v + prm.dt / (1 - prm.phi) * self.asm.load(1, QDEG_F, self.R[2]))

    def c0(self, c1_new, c2_new):
        asm, prm, nc = self.asm, self.prm, self.prm.ncomp
        A00, S, qd = asm.c0_row(prm, ("DG0", self.prev.u), QDEG_F)
        b = asm.c0_rhs(prm, qd, self.c0v, self.c1v, self.c2v if nc == 3 else None, self.v(self.fsrc), QDEG_F)
        b = b - S @ (self.v(c1_new) + (self.v(c2_new) if nc == 3 else 0.0))
        return fem.apply_dirichlet_symmetric(self.fold(A00), self.foldv(b), self.bcs.c0_dofs, self.bcs.c0_vals)

    def p(self, c0_new):
        L, bp = pressure_system(self.asm, self.prm, self.v(c0_new))
        return fem.apply_dirichlet_symmetric(self.fold(L), self.foldv(bp) - self.bcs.flux_load, self.bcs.p_dofs,
                                             np.zeros(len(self.bcs.p_dofs)))

    def u(self, p_new, c0_new):
        return recover_velocity(self.asm, self.prm, self.v(p_new), self.v(c0_new))


def step(asm: Assembler, prm: TransportParams, prev: StateF, fsrc, bcs: BCsF, solver="lu", rtol=1e-13, info=None, P=None):
    """One mode-F timestep; `solver='lu'` (splu) or 'krylov' (SciPy cg/bicgstab + Jacobi,
    the CPU baseline leg of bench.py).  `P`: periodic prolongation (V x ndof); fields, fsrc and bcs are then in dof numbering."""
    sysm = StepSystems(asm, prm, prev, fsrc, bcs, P)
    nc = prm.ncomp

    def solve_(Ab, x0, spd, tag):
        A, b = Ab
        if solver == "lu":
            return spla.splu(A.tocsc()).solve(b)
        Dinv = sp.diags(1.0 / A.diagonal())
        it = [0]
        fn = spla.cg if spd else spla.bicgstab
        x, _ = fn(A, b, x0=x0, rtol=rtol, atol=0.0, M=Dinv, maxiter=20000, callback=lambda xk: it.__setitem__(0, it[0] + 1))
        if info is not None:
            info[tag] = it[0]
        return x

    c1 = solve_(sysm.c1(), prev.c1, True, "c1")
    c2 = solve_(sysm.c2(), prev.c2, True, "c2") if nc == 3 else None
    c0 = solve_(sysm.c0(c1, c2), prev.c0, False, "c0")
    p = solve_(sysm.p(c0), prev.p, True, "p")
    return StateF(u=sysm.u(p, c0), p=p, c0=c0, c1=c1, c2=c2)
