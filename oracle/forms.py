"""Oracle restatement of the mupopp weak forms (mode R = the reference's own mixed
P2/P1 discretisation), NumPy/SciPy fp64.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows, term by term:
  * DarcyAdvection.darcy_bilinear                      /root/reference/src/modules/mupopp.py:595-625
  * DarcyAdvection.advection_diffusion                 mupopp.py:627-664
  * DarcyAdvection.advection_diffusion_two_component   mupopp.py:751-849
  * DarcyAdvection.advection_diffusion_three_component mupopp.py:850-960
  * CCS.advection_diffusion_three_component            mupopp.py:1147-1262
and the `solve(a==L, sol, bc)` semantics of the run scripts (assemble_system with
symmetric Dirichlet elimination + sparse LU), e.g.
/root/reference/src/run/CCS2D/CCS_3comp.py:257-263.

Two solution paths are provided and tested to agree (SURVEY.md finding 5):
  `monolithic_step`  -- one big matrix, splu: what the reference does;
  `segregated_step`  -- c1,c2 mass solves -> c0 ADR solve -> (u,p) saddle solve.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem

# explicit quadrature-degree table; shared with the GPU path
QDEG = dict(mass_p1=2, mass_p2=4, div=2, div_K=3, buoy=4, c0row=6, react=4, load_p2=4)


@dataclass
class TransportParams:
    """Physical parameters of the DarcyAdvection/CCS monolithic forms."""
    Pe: float = 100.0
    Da: float = 10.0
    phi: float = 0.01
    alpha: float = 0.005
    dt: float = 0.01
    K: object = 1.0                 # float or P1 nodal array
    sigma: float = -1.0             # -1 DarcyAdvection (mupopp.py:807), +1 CCS (mupopp.py:1204)
    rho0: float = 1.0               # rho(c0) = rho0 + gamma*c0
    gamma: float = 1.0
    reaction: str = "c0c1sq"        # 'c0c1sq' (mupopp.py:818,925) or 'c0c1' (mupopp.py:1222)
    fracs: tuple = (0.75, 1.0, 0.0)  # (R1, R2, R3) prefactors of Da*reaction
    ncomp: int = 2                  # number of concentration fields (2 or 3)
    SUPG: int = 1
    zhat: tuple = (0.0, 1.0)

    @property
    def beta(self):
        return self.alpha / self.phi   # mupopp.py:594,1146


def darcy_two_component_params(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=0.01, K=1.0, zhat=(0.0, 1.0),
                               SUPG=1, gam_rho=-1.0, rho_0=1.0):
    """mupopp.py:751-849: deltarho = rho_0 - gam_rho*c0, term -K*deltarho*v.zhat; fac = 0.75."""
    fac = (24.0 + 12.0 + 3.0 * 16.0) / 2.0 / 56.0
    return TransportParams(Pe=Pe, Da=Da, phi=phi, alpha=alpha, dt=dt, K=K, sigma=-1.0, rho0=rho_0,
                           gamma=-gam_rho, reaction="c0c1sq", fracs=(fac, 1.0, 0.0), ncomp=2, SUPG=SUPG, zhat=zhat)


def darcy_three_component_params(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=0.01, K=1.0, zhat=(0.0, 1.0),
                                 SUPG=1, gam_rho=-1.0, rho_0=1.0):
    """mupopp.py:850-960: fractions 84/392.32 (carbonate), 112/392.32 (Fe), 12/392.32 (C)."""
    T = 392.32
    return TransportParams(Pe=Pe, Da=Da, phi=phi, alpha=alpha, dt=dt, K=K, sigma=-1.0, rho0=rho_0,
                           gamma=-gam_rho, reaction="c0c1sq", fracs=(84.0 / T, 112.0 / T, 12.0 / T), ncomp=3,
                           SUPG=SUPG, zhat=zhat)


def ccs_three_component_params(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=0.01, K=1.0, zhat=(0.0, 1.0),
                               SUPG=1, gam_rho=1.0):
    """mupopp.py:1147-1262: deltarho = 1 + gam_rho*c0, term +K*deltarho*v.zhat; fractions /698."""
    T = 698.0
    return TransportParams(Pe=Pe, Da=Da, phi=phi, alpha=alpha, dt=dt, K=K, sigma=+1.0, rho0=1.0,
                           gamma=gam_rho, reaction="c0c1", fracs=(62.0 / T, 278.0 / T, 100.0 / T), ncomp=3,
                           SUPG=SUPG, zhat=zhat)


def tau_supg(SUPG, Pe, Da, h, vnorm, extra=0.0):
    """mupopp.py:835-846 (identical at :945-956, :1245-1256); SUPG == 5: mupopp.py:738 with extra = Da*c1^n(x)/phi."""
    if SUPG == 5:
        return 1.0 / (4.0 / (Pe * h * h) + 2.0 * vnorm / h + extra)
    if SUPG == 1:
        return 1.0 / (4.0 / (Pe * h * h) + 2.0 * vnorm / h)
    if SUPG == 2:
        return 1.0 / (4.0 / (Pe * h * h) + 2.0 * vnorm / h + Da)
    a = Pe * vnorm * h / 2.0
    coth = (np.e ** (2.0 * a) + 1.0) / (np.e ** (2.0 * a) - 1.0)
    return 0.5 * h * (coth - 1.0 / a) / vnorm


class Assembler:
    """Caches geometry and tabulations for one mesh."""

    def __init__(self, mesh: fem.Mesh):
        self.mesh = mesh
        self.J, self.detJ, self.Jinv = fem.geometry(mesh)
        self.h = mesh.hmax_cells()
        self.P1 = fem.ScalarSpace(mesh, 1)
        self._P2 = None
        self._P3 = None
        self._tab = {}

    @property
    def P2(self):
        if self._P2 is None:
            self._P2 = fem.ScalarSpace(self.mesh, 2)
        return self._P2

    @property
    def P3(self):
        if self._P3 is None:
            self._P3 = fem.ScalarSpace(self.mesh, 3)
        return self._P3

    def space(self, degree):
        return {1: self.P1, 2: self.P2, 3: self.P3}[degree] if degree != 1 else self.P1

    def quad(self, qdeg):
        return fem.simplex_quadrature(self.mesh.dim, qdeg)

    def tab(self, degree, qdeg):
        key = (degree, qdeg)
        if key not in self._tab:
            pts, w = self.quad(qdeg)
            phi, dref = fem.tabulate(degree, pts)
            self._tab[key] = (w, phi, dref)
        return self._tab[key]

    def coef_q(self, coef, qdeg):
        """float or P1 nodal array -> (C, nq) values at quadrature points."""
        w, phi1, _ = self.tab(1, qdeg)
        if np.isscalar(coef):
            return np.full((self.mesh.nc, len(w)), float(coef))
        return fem.eval_at_quad(self.P1, np.asarray(coef, dtype=float), phi1)

    # ---- bilinear blocks -------------------------------------------------
    def mass(self, deg_r, deg_c, qdeg, coef=1.0):
        """M[i,j] = int coef * test_i * trial_j."""
        w, pr, _ = self.tab(deg_r, qdeg)
        _, pc, _ = self.tab(deg_c, qdeg)
        cq = self.coef_q(coef, qdeg)
        Ae = np.einsum("q,cq,qi,qj,c->cij", w, cq, pr, pc, self.detJ, optimize=True)
        sr, sc = self.space(deg_r), self.space(deg_c)
        return fem.scatter_matrix(Ae, sr.cell_dofs, sc.cell_dofs, (sr.ndof, sc.ndof))

    def stiffness(self, deg, qdeg, coef=1.0):
        w, _, dref = self.tab(deg, qdeg)
        cq = self.coef_q(coef, qdeg)
        g = fem.phys_grads(dref, self.Jinv)
        Ae = np.einsum("q,cq,cqia,cqja,c->cij", w, cq, g, g, self.detJ, optimize=True)
        s = self.space(deg)
        return fem.scatter_matrix(Ae, s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))

    def div_blocks(self, deg_q, deg_v, qdeg, coef=1.0):
        """B_c[j, i] = int coef * q_j * d(psi_i)/dx_c, one block per direction c."""
        w, pq, _ = self.tab(deg_q, qdeg)
        _, _, dref = self.tab(deg_v, qdeg)
        cq = self.coef_q(coef, qdeg)
        g = fem.phys_grads(dref, self.Jinv)
        sq, sv = self.space(deg_q), self.space(deg_v)
        out = []
        for c in range(self.mesh.dim):
            Ae = np.einsum("q,cq,qj,cqi,c->cji", w, cq, pq, g[..., c], self.detJ, optimize=True)
            out.append(fem.scatter_matrix(Ae, sq.cell_dofs, sv.cell_dofs, (sq.ndof, sv.ndof)))
        return out

    def load(self, deg, qdeg, fq):
        """b[i] = int f * test_i with f given at quadrature points (C, nq)."""
        w, p, _ = self.tab(deg, qdeg)
        be = np.einsum("q,cq,qi,c->ci", w, fq, p, self.detJ, optimize=True)
        s = self.space(deg)
        return fem.scatter_vector(be, s.cell_dofs, s.ndof)

    # ---- velocity at quadrature points -----------------------------------
    def velocity_q(self, vel, qdeg):
        """vel: ('P2', [comp arrays]) | ('P1', [...]) | ('DG0', (C,d) array) -> (C, nq, d)."""
        kind, data = vel
        w = self.tab(1, qdeg)[0]
        if kind == "DG0":
            return np.broadcast_to(np.asarray(data)[:, None, :], (self.mesh.nc, len(w), self.mesh.dim)).copy()
        deg = {"P1": 1, "P2": 2, "P3": 3}[kind]
        _, ph, _ = self.tab(deg, qdeg)
        s = self.space(deg)
        return np.stack([fem.eval_at_quad(s, np.asarray(c), ph) for c in data], axis=-1)

    # ---- the c0 row: CN + SUPG advection-diffusion-reaction ---------------
    def c0_row(self, prm: TransportParams, vel, qdeg, c1n=None):
        """Matrix pieces of the c0 row (mupopp.py:823-848 and twins):
           A00 = int eta*phi + dt/2 [eta u.grad(phi) + grad(eta).grad(phi)/Pe] + tau (u.grad eta)(phi + dt/2 u.grad phi)
           S   = int tau (u.grad eta) phi           (coupling to the NEW c1, c2 through r)
        plus the quadrature-point data needed for the RHS."""
        w, p1, dref = self.tab(1, qdeg)
        g = fem.phys_grads(dref, self.Jinv)                         # (C,q,b,d)
        uq = self.velocity_q(vel, qdeg)                             # (C,q,d)
        vn = np.sqrt((uq * uq).sum(-1))
        extra = prm.Da * fem.eval_at_quad(self.P1, c1n, p1) / prm.phi if prm.SUPG == 5 else 0.0
        tau = tau_supg(prm.SUPG, prm.Pe, prm.Da, self.h[:, None], vn, extra)
        ug = np.einsum("cqd,cqbd->cqb", uq, g, optimize=True)       # u.grad(basis)
        wd = w[None, :] * self.detJ[:, None]
        dt = prm.dt
        Ae = np.einsum("cq,qi,qj->cij", wd, p1, p1, optimize=True)
        Ae += 0.5 * dt * np.einsum("cq,qi,cqj->cij", wd, p1, ug, optimize=True)
        Ae += 0.5 * dt / prm.Pe * np.einsum("cq,cqia,cqja->cij", wd, g, g, optimize=True)
        Se = np.einsum("cq,cq,cqi,qj->cij", wd, tau, ug, p1, optimize=True)
        Ae += Se + 0.5 * dt * np.einsum("cq,cq,cqi,cqj->cij", wd, tau, ug, ug, optimize=True)
        s = self.P1
        A00 = fem.scatter_matrix(Ae, s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))
        S = fem.scatter_matrix(Se, s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))
        return A00, S, dict(wd=wd, p1=p1, g=g, ug=ug, tau=tau)

    def reaction_q(self, prm, c0n, c1n, qdeg):
        _, p1, _ = self.tab(1, qdeg)
        a = fem.eval_at_quad(self.P1, c0n, p1)
        b = fem.eval_at_quad(self.P1, c1n, p1)
        base = prm.Da * a * b * (b if prm.reaction == "c0c1sq" else 1.0)
        return [f * base for f in prm.fracs]

    def c0_rhs(self, prm, qd, c0n, c1n, c2n, fsrc, qdeg):
        """RHS of the c0 row, everything known at step n (old-state part of r included);
        the NEW c1,c2 enter through -S@(c1+c2) added by the caller."""
        wd, p1, g, ug, tau = qd["wd"], qd["p1"], qd["g"], qd["ug"], qd["tau"]
        dt = prm.dt
        s = self.P1
        c0q = fem.eval_at_quad(s, c0n, p1)
        gc0 = fem.grad_at_quad(s, c0n, g)
        u_gc0 = np.einsum("cqb,cb->cq", ug, c0n[s.cell_dofs], optimize=True)
        R1, R2, R3 = self.reaction_q(prm, c0n, c1n, qdeg)
        fq = fem.eval_at_quad(s, fsrc, p1)
        c1q = fem.eval_at_quad(s, c1n, p1)
        c2q = fem.eval_at_quad(s, c2n, p1) if prm.ncomp == 3 else 0.0
        gal = c0q - 0.5 * dt * u_gc0 - dt * R1 / prm.phi + prm.beta * dt * fq
        be = np.einsum("cq,cq,qi->ci", wd, gal, p1, optimize=True)
        be -= 0.5 * dt / prm.Pe * np.einsum("cq,cqia,cqa->ci", wd, g, gc0, optimize=True)
        # known part of r: -c0n + dt/2 u.grad c0n + dt R1/phi - beta dt f - c1n + dt R2/(1-phi) [- c2n - dt R3/(1-phi)]
        rk = -c0q + 0.5 * dt * u_gc0 + dt * R1 / prm.phi - prm.beta * dt * fq - c1q + dt * R2 / (1 - prm.phi)
        if prm.ncomp == 3:
            rk = rk - c2q - dt * R3 / (1 - prm.phi)
        be -= np.einsum("cq,cq,cqi,cq->ci", wd, tau, ug, rk, optimize=True)
        return fem.scatter_vector(be, s.cell_dofs, s.ndof)


# --------------------------------------------------------------------------- config-1 forms

def darcy_bilinear_system(asm: Assembler, K=0.1, zhat=(0.0, 1.0), phi=1.0):
    """mupopp.py:595-625: a = (u.v - K div(v) p + div(u) q) dx, L = K f v.zhat dx, f == 1.
    Unknown layout [u_x(P2), u_y(P2), (u_z), p(P1)] (UFC mixed numbering)."""
    d = asm.mesh.dim
    M = asm.mass(2, 2, QDEG["mass_p2"], phi)
    B = asm.div_blocks(1, 2, QDEG["div"])                      # int q d_c psi
    n2, n1 = asm.P2.ndof, asm.P1.ndof
    rows = []
    for c in range(d):
        row = [None] * (d + 1)
        row[c] = M
        row[d] = -K * B[c].T
        rows.append(row)
    rows.append([B[c] for c in range(d)] + [sp.csr_matrix((n1, n1))])
    A = sp.bmat(rows, format="csr")
    ones = np.ones((asm.mesh.nc, len(asm.quad(QDEG["load_p2"])[1])))
    lv = asm.load(2, QDEG["load_p2"], ones)
    b = np.concatenate([K * zhat[c] * lv for c in range(d)] + [np.zeros(n1)])
    return A, b


def adr_single_system(asm: Assembler, Pe, Da, dt, vel, c_prev, qdeg=None):
    """mupopp.py:627-664 (config 1): CN + SUPG with tau = h/(2|w|), reaction Da*c^n.
       F = v(c-c^n) + dt(v w.grad cm + grad v.grad cm/Pe) + dt Da c^n v + tau (w.grad v) r,
       r = c - c^n + dt(w.grad cm + Da c^n)      (div grad of P1 == 0)."""
    qdeg = qdeg or QDEG["c0row"]
    w, p1, dref = asm.tab(1, qdeg)
    g = fem.phys_grads(dref, asm.Jinv)
    uq = asm.velocity_q(vel, qdeg)
    vn = np.sqrt((uq * uq).sum(-1))
    tau = asm.h[:, None] / (2.0 * vn)
    ug = np.einsum("cqd,cqbd->cqb", uq, g, optimize=True)
    wd = w[None, :] * asm.detJ[:, None]
    s = asm.P1
    Ae = np.einsum("cq,qi,qj->cij", wd, p1, p1, optimize=True)
    Ae += 0.5 * dt * np.einsum("cq,qi,cqj->cij", wd, p1, ug, optimize=True)
    Ae += 0.5 * dt / Pe * np.einsum("cq,cqia,cqja->cij", wd, g, g, optimize=True)
    Ae += np.einsum("cq,cq,cqi,qj->cij", wd, tau, ug, p1, optimize=True)
    Ae += 0.5 * dt * np.einsum("cq,cq,cqi,cqj->cij", wd, tau, ug, ug, optimize=True)
    A = fem.scatter_matrix(Ae, s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))
    cq = fem.eval_at_quad(s, c_prev, p1)
    gc = fem.grad_at_quad(s, c_prev, g)
    u_gc = np.einsum("cqb,cb->cq", ug, c_prev[s.cell_dofs], optimize=True)
    gal = cq - 0.5 * dt * u_gc - dt * Da * cq
    be = np.einsum("cq,cq,qi->ci", wd, gal, p1, optimize=True)
    be -= 0.5 * dt / Pe * np.einsum("cq,cqia,cqa->ci", wd, g, gc, optimize=True)
    rk = -cq + 0.5 * dt * u_gc + dt * Da * cq
    be -= np.einsum("cq,cq,cqi,cq->ci", wd, tau, ug, rk, optimize=True)
    b = fem.scatter_vector(be, s.cell_dofs, s.ndof)
    return A, b


# --------------------------------------------------------------------------- monolithic multi-component step

@dataclass
class State:
    u: list            # d arrays of P2 dofs
    p: np.ndarray
    c0: np.ndarray
    c1: np.ndarray
    c2: np.ndarray = None


@dataclass
class BCs:
    """Dirichlet data: velocity dofs per component and c0 dofs."""
    u_dofs: list = field(default_factory=list)   # per component: int array of P2 dofs
    u_vals: list = field(default_factory=list)
    c0_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    c0_vals: np.ndarray = field(default_factory=lambda: np.zeros(0))


def _darcy_blocks(asm, prm):
    d = asm.mesh.dim
    Kc = prm.K
    Kconst = np.isscalar(Kc)
    M = asm.mass(2, 2, QDEG["mass_p2"], prm.phi)
    B = asm.div_blocks(1, 2, QDEG["div"])
    BK = [Kc * b for b in B] if Kconst else asm.div_blocks(1, 2, QDEG["div_K"], Kc)
    G = asm.mass(2, 1, QDEG["buoy"], Kc)                      # int K psi_i phi_j  (P2 x P1)
    ones = np.ones((asm.mesh.nc, len(asm.quad(QDEG["buoy"])[1])))
    lK = asm.load(2, QDEG["buoy"], asm.coef_q(Kc, QDEG["buoy"]) * ones)
    return M, B, BK, G, lK


def monolithic_step(asm: Assembler, prm: TransportParams, prev: State, fsrc, bcs: BCs, qdeg=None):
    """One reference timestep: assemble the full [u,p,c0,c1(,c2)] matrix exactly as
    lhs(F)/rhs(F) of mupopp.py:751-849 / :850-960 / :1147-1262 would, apply the
    Dirichlet BCs symmetrically and LU-solve (CCS2D/CCS_3comp.py:260-261)."""
    qdeg = qdeg or QDEG["c0row"]
    d = asm.mesh.dim
    n2, n1 = asm.P2.ndof, asm.P1.ndof
    nc = prm.ncomp
    M, B, BK, G, lK = _darcy_blocks(asm, prm)
    A00, S, qd = asm.c0_row(prm, ("P2", prev.u), qdeg)
    M1 = asm.mass(1, 1, QDEG["mass_p1"])
    c2n = prev.c2 if nc == 3 else None
    b_c0 = asm.c0_rhs(prm, qd, prev.c0, prev.c1, c2n, fsrc, qdeg)
    R1, R2, R3 = asm.reaction_q(prm, prev.c0, prev.c1, QDEG["react"])
    b_c1 = M1 @ prev.c1 - prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R2)
    nblk = d + 1 + nc
    rows = [[None] * nblk for _ in range(nblk)]
    zh = prm.zhat
    for c in range(d):
        rows[c][c] = M
        rows[c][d] = -BK[c].T
        rows[c][d + 1] = prm.sigma * prm.gamma * zh[c] * G
        rows[d][c] = B[c]
    rows[d][d] = sp.csr_matrix((n1, n1))
    rows[d + 1][d + 1] = A00
    rows[d + 1][d + 2] = S
    rows[d + 2][d + 2] = M1
    rhs = [-prm.sigma * prm.rho0 * zh[c] * lK for c in range(d)] + [np.zeros(n1), b_c0, b_c1]
    if nc == 3:
        rows[d + 1][d + 3] = S
        rows[d + 3][d + 3] = M1
        rhs.append(M1 @ prev.c2 + prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R3))
    A = sp.bmat(rows, format="csr")
    b = np.concatenate(rhs)
    dofs, vals = [], []
    for c in range(len(bcs.u_dofs)):
        dofs.append(c * n2 + np.asarray(bcs.u_dofs[c]))
        vals.append(np.asarray(bcs.u_vals[c], dtype=float))
    dofs.append(d * n2 + n1 + np.asarray(bcs.c0_dofs))
    vals.append(np.asarray(bcs.c0_vals, dtype=float))
    dofs = np.concatenate(dofs).astype(np.int64)
    vals = np.concatenate(vals)
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, dofs, vals)
    x = spla.splu(Abc.tocsc()).solve(bbc)
    u = [x[c * n2:(c + 1) * n2] for c in range(d)]
    o = d * n2
    st = State(u=u, p=x[o:o + n1], c0=x[o + n1:o + 2 * n1], c1=x[o + 2 * n1:o + 3 * n1],
               c2=x[o + 3 * n1:o + 4 * n1] if nc == 3 else None)
    return st, (A, b)


def solve_saddle(asm, prm, M, B, BK, rhs_u, bcs: BCs):
    """[phi M, -K B^T; B, 0][u;p] = [rhs_u; 0] with Dirichlet on u, direct LU."""
    d = asm.mesh.dim
    n2, n1 = asm.P2.ndof, asm.P1.ndof
    rows = []
    for c in range(d):
        row = [None] * (d + 1)
        row[c] = M
        row[d] = -BK[c].T
        rows.append(row)
    rows.append([B[c] for c in range(d)] + [sp.csr_matrix((n1, n1))])
    A = sp.bmat(rows, format="csr")
    b = np.concatenate(list(rhs_u) + [np.zeros(n1)])
    dofs = np.concatenate([c * n2 + np.asarray(bcs.u_dofs[c]) for c in range(len(bcs.u_dofs))]).astype(np.int64) \
        if bcs.u_dofs else np.zeros(0, dtype=np.int64)
    vals = np.concatenate([np.asarray(v, dtype=float) for v in bcs.u_vals]) if bcs.u_vals else np.zeros(0)
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, dofs, vals)
    x = spla.splu(Abc.tocsc()).solve(bbc)
    return [x[c * n2:(c + 1) * n2] for c in range(d)], x[d * n2:]


def segregated_step(asm: Assembler, prm: TransportParams, prev: State, fsrc, bcs: BCs, qdeg=None):
    """Algebraically identical block-triangular sequence:
    c1,c2 mass solves -> c0 ADR solve -> (u,p) saddle solve."""
    qdeg = qdeg or QDEG["c0row"]
    d = asm.mesh.dim
    nc = prm.ncomp
    M1 = asm.mass(1, 1, QDEG["mass_p1"])
    lu1 = spla.splu(M1.tocsc())
    R1, R2, R3 = asm.reaction_q(prm, prev.c0, prev.c1, QDEG["react"])
    c1 = lu1.solve(M1 @ prev.c1 - prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R2))
    c2 = None
    if nc == 3:
        c2 = lu1.solve(M1 @ prev.c2 + prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R3))
    A00, S, qd = asm.c0_row(prm, ("P2", prev.u), qdeg)
    b = asm.c0_rhs(prm, qd, prev.c0, prev.c1, prev.c2 if nc == 3 else None, fsrc, qdeg)
    b = b - S @ (c1 + (c2 if nc == 3 else 0.0))
    Abc, bbc = fem.apply_dirichlet_symmetric(A00, b, np.asarray(bcs.c0_dofs, dtype=np.int64), np.asarray(bcs.c0_vals, float))
    c0 = spla.splu(Abc.tocsc()).solve(bbc)
    M, B, BK, G, lK = _darcy_blocks(asm, prm)
    Gc0 = G @ c0
    rhs_u = [-prm.sigma * prm.zhat[c] * (prm.rho0 * lK + prm.gamma * Gc0) for c in range(d)]
    u, p = solve_saddle(asm, prm, M, B, BK, rhs_u, bcs)
    return State(u=u, p=p, c0=c0, c1=c1, c2=c2)


# --------------------------------------------------------------------------- a6: two-component adaptive form

def adaptive_source(X):
    """The hard-coded source of mupopp.py:693-698 (`3.14`, not pi), nodally interpolated (degree=1)."""
    s = np.zeros(X.shape[0])
    for k in range(1, 11):
        s += 1.0 + np.sin(k * X[:, 0] * 3.14)
    return (1.0 - np.tanh(X[:, 1] / 0.01)) * s / 20.0


def two_component_adaptive_params(Pe, Da, phi, alpha, dt):
    """mupopp.py:667-749: f = Da*u0*c0 (bilinear), R1 = f/phi, R2 = f/(1-phi), source alpha/phi*dt*f1, tau with Da*c1^n/phi."""
    return TransportParams(Pe=Pe, Da=Da, phi=phi, alpha=alpha, dt=dt, K=1.0, sigma=0.0, rho0=0.0, gamma=0.0, reaction="c0c1",
                           fracs=(1.0, 1.0, 0.0), ncomp=2, SUPG=5, zhat=(0.0, 1.0))


def two_component_adaptive_step(asm: Assembler, prm: TransportParams, vel, c0n, c1n, fsrc, c0_dofs, c0_vals, qdeg=None):
    """Monolithic [c0, c1] solve of DarcyAdvection.advection_diffusion_two_component_adaptive with a given velocity
    (`project(advect_term, DG0)` at mupopp.py:704-710 is dead code: its result is unused)."""
    qdeg = qdeg or QDEG["c0row"]
    n1 = asm.P1.ndof
    A00, S, qd = asm.c0_row(prm, vel, qdeg, c1n=c1n)
    M1 = asm.mass(1, 1, QDEG["mass_p1"])
    b0 = asm.c0_rhs(prm, qd, c0n, c1n, None, fsrc, qdeg)
    R1, R2, R3 = asm.reaction_q(prm, c0n, c1n, QDEG["react"])
    b1 = M1 @ c1n - prm.dt / (1 - prm.phi) * asm.load(1, QDEG["react"], R2)
    A = sp.bmat([[A00, S], [None, M1]], format="csr")
    b = np.concatenate([b0, b1])
    Abc, bbc = fem.apply_dirichlet_symmetric(A, b, np.asarray(c0_dofs, dtype=np.int64), np.asarray(c0_vals, dtype=float))
    x = spla.splu(Abc.tocsc()).solve(bbc)
    return x[:n1], x[n1:]
