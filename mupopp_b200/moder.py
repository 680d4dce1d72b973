"""Mode R: the reference's own mixed discretisation (continuous P2 velocity, P1 pressure / concentrations) on the GPU.

* `MixedDarcyR`   -- the saddle system  [phi M, -K B^T; B, 0][u; p] = [g; 0]  of DarcyAdvection.darcy_bilinear
                     (/root/reference/src/modules/mupopp.py:595-625) and of the Darcy rows of the monolithic forms (:804-807, :903-907, :1201-1204),
                     solved by Schur-complement PCG: S p = B M^-1 K B^T p, inner P2 mass solves by Jacobi-PCG.
* `AdrSolverR`    -- the P1 CN+SUPG advection-diffusion-reaction row with a P2 velocity evaluated at quadrature points:
                     DarcyAdvection.advection_diffusion (mupopp.py:627-664) and the c0 row of the multi-component forms.
* `TransportSolverR` -- the full segregated multi-component step (c1,c2 -> c0 -> (u,p)), algebraically the reference's monolithic LU solve.
All numerics run in libmupopp_b200.so; this file only sequences C-ABI calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import reference_element as ref
from ._lib import MpQuad, check, ptr
from .device import Context, DeviceMesh, Plan, transport_params

QDEG_C0ROW = 6    # shared with oracle/forms.py::QDEG["c0row"]; part of the parity contract (tau is non-polynomial)


class Spaces:
    """P1 / P2 / P3 scalar Lagrange dofmaps of one mesh (UFC numbering: vertices, then edge dofs, then face/cell dofs) on host and device."""

    def __init__(self, ctx: Context, mesh):
        self.ctx, self.mesh = ctx, mesh
        self.dm = DeviceMesh(ctx, mesh)
        d = mesh.dim
        self.n1 = mesh.num_vertices()
        self.n2 = self.n1 + mesh.num_edges()
        self.cd1_host = mesh.cells
        self.cd2_host = np.ascontiguousarray(np.concatenate([mesh.cells, self.n1 + mesh.cell_edges], axis=1).astype(np.int32))
        self.cd1 = self.dm.cells
        self.cd2 = ctx.to_device(self.cd2_host, torch.int32)
        self.nb1, self.nb2 = d + 1, self.cd2_host.shape[1]
        self._cd3 = None
        self._plans = {}
        self._tens = {}

    # ---- P3: 2 dofs per edge (nearer the lower vertex first), 1 per face (tets) / per cell (triangles)
    def _init_p3(self):
        if self._cd3 is not None:
            return
        m = self.mesh
        ce = m.cell_edges.astype(np.int64)
        ed = np.stack([self.n1 + 2 * ce, self.n1 + 2 * ce + 1], axis=2).reshape(m.num_cells(), -1)
        base = self.n1 + 2 * m.num_edges()
        if m.dim == 2:
            inner = base + np.arange(m.num_cells(), dtype=np.int64)[:, None]
            self.n3 = base + m.num_cells()
        else:
            inner = base + m.cell_faces.astype(np.int64)
            self.n3 = base + m.faces.shape[0]
        self.cd3_host = np.ascontiguousarray(np.concatenate([m.cells.astype(np.int64), ed, inner], axis=1).astype(np.int32))
        self._cd3 = self.ctx.to_device(self.cd3_host, torch.int32)
        self.nb3 = self.cd3_host.shape[1]

    def ndof(self, degree):
        if degree == 3:
            self._init_p3()
        return {1: self.n1, 2: self.n2}[degree] if degree < 3 else self.n3

    def cell_dofs(self, degree):
        if degree == 3:
            self._init_p3()
            return self._cd3
        return self.cd1 if degree == 1 else self.cd2

    def nb(self, degree):
        return ref.nbasis(degree, self.mesh.dim)

    def dof_coords(self, degree):
        m = self.mesh
        if degree == 1:
            return m.coords
        e = m.edges
        a, b = m.coords[e[:, 0]], m.coords[e[:, 1]]
        if degree == 2:
            return np.concatenate([m.coords, 0.5 * (a + b)], axis=0)
        ed = np.stack([(2.0 * a + b) / 3.0, (a + 2.0 * b) / 3.0], axis=1).reshape(-1, m.dim)
        inner = m.coords[m.cells].mean(1) if m.dim == 2 else m.coords[m.faces].mean(1)
        return np.concatenate([m.coords, ed, inner], axis=0)

    def plan(self, deg_r, deg_c):
        key = (deg_r, deg_c)
        if key not in self._plans:
            self._plans[key] = Plan(self.ctx, self.cell_dofs(deg_r), self.cell_dofs(deg_c), self.ndof(deg_r), self.ndof(deg_c))
        return self._plans[key]

    def tensor(self, kind, deg_r, deg_c, p1coef=False):
        key = (kind, deg_r, deg_c, p1coef)
        if key not in self._tens:
            self._tens[key] = self.ctx.to_device(ref.reference_tensor(kind, self.mesh.dim, deg_r, deg_c, p1coef).ravel())
        return self._tens[key]

    def assemble(self, kind, deg_r, deg_c, direction=0, coef=None, scale=1.0):
        """CSR matrix of one reference-tensor block; coef: None | float (folded into scale) | nodal P1 device tensor."""
        ctx = self.ctx
        plan = self.plan(deg_r, deg_c)
        lr, lc = self.nb(deg_r), self.nb(deg_c)
        kid = {"mass": 0, "grad": 1, "gradT": 2, "stiff": 3}[kind]
        nodal = coef is not None and not np.isscalar(coef)
        if coef is not None and np.isscalar(coef):
            scale = scale * float(coef)
        T0 = self.tensor(kind, deg_r, deg_c, nodal)
        emat = ctx.empty(self.dm.ncell * lr * lc)
        A = plan.new_matrix(sell=True)
        plan.reftensor_scatter(self.dm, kid, direction, lr, lc, T0, coef if nodal else None, 1 if nodal else 0, scale, emat, A)
        return A


class MixedDarcyR:
    def __init__(self, ctx: Context, spaces: Spaces, K=0.1, phi=1.0, u_isbc=None, u_g=None, rtol=1e-13, inner_rtol=1e-14, maxit=5000):
        if not np.isscalar(K):
            raise NotImplementedError("mode R Schur-PCG needs a constant K (with a P1 K the Schur complement is not symmetric)")
        self.ctx, self.sp, self.K, self.phi = ctx, spaces, float(K), float(phi)
        self.rtol, self.inner_rtol, self.maxit = rtol, inner_rtol, maxit
        d = spaces.mesh.dim
        self.d = d
        n1, n2 = spaces.n1, spaces.n2
        self.M0 = spaces.assemble("mass", 2, 2, scale=phi)                       # pristine: BC lifting uses its columns
        self.B = [spaces.assemble("grad", 1, 2, direction=c) for c in range(d)]  # (q, div u) per direction
        self.BT = [spaces.assemble("gradT", 2, 1, direction=c, scale=self.K) for c in range(d)]   # K (div v, p)
        self.G = spaces.assemble("mass", 2, 1, scale=self.K)                     # K int psi_i phi_j : buoyancy / load coupling
        self.zero2 = ctx.zeros(n2)
        self.isbc = [ctx.to_device(np.zeros(n2, dtype=np.uint8) if u_isbc is None else u_isbc[c], torch.uint8) for c in range(d)]
        self.g = [ctx.to_device(np.zeros(n2) if u_g is None else np.asarray(u_g[c], dtype=np.float64)) for c in range(d)]
        self.Mc = []
        for c in range(d):
            Mc = spaces.plan(2, 2).new_matrix(sell=True)
            Mc.vals.copy_(self.M0.vals)
            Mc.sell_dirty = True
            Mc.apply_dirichlet(None, self.isbc[c], self.g[c])
            Mc.jacobi_setup()
            self.Mc.append(Mc)
        # Jacobi preconditioner of S: diagonal of the P1 (K/phi) Laplacian (spectrally equivalent to diag(S))
        p11 = spaces.plan(1, 1)
        emat = ctx.empty(spaces.dm.ncell * (d + 1) * (d + 1))
        check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(spaces.dm.c), None, self.K, 1.0 / self.phi, ptr(emat)), "mp_elem_p1_stiffness")
        Lp = p11.new_matrix(sell=False)
        p11.scatter_matrix(emat, Lp)
        self.dinvS = Lp.jacobi_setup().clone()
        # work vectors
        self.t2 = [ctx.zeros(n2) for _ in range(d)]
        self.w2 = [ctx.zeros(n2) for _ in range(d)]
        self.u0 = [ctx.zeros(n2) for _ in range(d)]
        self.tmp1 = ctx.zeros(n1)
        self.inner_its = 0

    def _minv_masked(self, c, t, out):
        """out = M_c^-1 t with t zeroed on the Dirichlet rows of component c."""
        self.Mc[c].apply_dirichlet(t, self.isbc[c], self.zero2, matrix=False)
        out.zero_()
        r = self.Mc[c].solve("pcg", t, out, rtol=self.inner_rtol, maxit=2000, refresh_jacobi=False)
        self.inner_its += r.niter

    def apply_S(self, p, out):
        out.zero_()
        for c in range(self.d):
            self.BT[c].spmv(p, self.t2[c])
            self._minv_masked(c, self.t2[c], self.w2[c])
            self.B[c].spmv(self.w2[c], self.tmp1)
            self.ctx.axpby(1.0, self.tmp1, 1.0, out)

    def solve(self, rhs_u, p):
        """rhs_u: d device vectors (P2 load, before BCs).  p: device vector, initial guess in / solution out.  Returns (u list, outer its)."""
        ctx, d = self.ctx, self.d
        n1 = self.sp.n1
        # u0_c = M_c^-1 (rhs_c - M0[:,D] g ; g on D)
        bS = ctx.zeros(n1)
        for c in range(d):
            rc = rhs_u[c].clone()
            self.M0.spmv(self.g[c], self.t2[c])            # g is zero off the Dirichlet set
            ctx.axpby(-1.0, self.t2[c], 1.0, rc)
            self.Mc[c].apply_dirichlet(rc, self.isbc[c], self.g[c], matrix=False)
            self.u0[c].zero_()
            self.Mc[c].solve("pcg", rc, self.u0[c], rtol=self.inner_rtol, maxit=2000, refresh_jacobi=False)
            self.B[c].spmv(self.u0[c], self.tmp1)
            ctx.axpby(-1.0, self.tmp1, 1.0, bS)
        # outer PCG on S p = bS (host-driven: the operator is a composition of solves)
        r = ctx.zeros(n1)
        Sp = ctx.zeros(n1)
        self.apply_S(p, Sp)
        r.copy_(bS)
        ctx.axpby(-1.0, Sp, 1.0, r)
        z = ctx.zeros(n1)
        pd = ctx.zeros(n1)
        bnorm = np.sqrt(ctx.dot(bS, bS))
        tol = self.rtol * bnorm
        its = 0
        rz = 0.0
        while True:
            rn = np.sqrt(ctx.dot(r, r))
            if rn <= tol or its >= self.maxit:
                break
            ctx.vmul(1.0, self.dinvS, r, z)                 # Jacobi apply
            rz_new = ctx.dot(r, z)
            beta = rz_new / rz if its > 0 else 0.0
            ctx.axpby(1.0, z, beta, pd)
            self.apply_S(pd, Sp)
            alpha = rz_new / ctx.dot(pd, Sp)
            ctx.axpby(alpha, pd, 1.0, p)
            ctx.axpby(-alpha, Sp, 1.0, r)
            rz = rz_new
            its += 1
        u = []
        for c in range(d):
            self.BT[c].spmv(p, self.t2[c])
            self._minv_masked(c, self.t2[c], self.w2[c])
            uc = self.u0[c].clone()
            ctx.axpby(1.0, self.w2[c], 1.0, uc)
            u.append(uc)
        self.last_outer = its
        return u, its


class AdrSolverR:
    """P1 CN+SUPG transport row with a P2 (or P1) vector velocity, assembled by quadrature (shared degree-6 table)."""

    def __init__(self, ctx: Context, spaces: Spaces, prm, c_isbc, c_g, vel_degree=2, qdeg=QDEG_C0ROW, rtol=1e-13, maxit=20000, method="gmres"):
        self.ctx, self.sp, self.prm = ctx, spaces, prm
        self.rtol, self.maxit, self.method = rtol, maxit, method
        d = spaces.mesh.dim
        w, lam, vphi = ref.quadrature_tables(d, qdeg, vel_degree)
        self.qw, self.qlam, self.qvphi = ctx.to_device(w), ctx.to_device(lam.ravel()), ctx.to_device(vphi.ravel())
        self.quad = MpQuad(len(w), vphi.shape[1], self.qw.data_ptr(), self.qlam.data_ptr(), self.qvphi.data_ptr(), None)
        self.vdofs = spaces.cell_dofs(vel_degree)
        self.plan = spaces.plan(1, 1)
        self.A = self.plan.new_matrix(sell=True)
        l = d + 1
        self.emat = ctx.empty(spaces.dm.ncell * l * l)
        self.evec = ctx.empty(spaces.dm.ncell * l)
        self.rhs = ctx.zeros(spaces.n1)
        self.isbc = ctx.to_device(np.asarray(c_isbc, dtype=np.uint8), torch.uint8)
        self.g = ctx.to_device(np.asarray(c_g, dtype=np.float64))

    def solve(self, vel, c0n, c_out, c1n=None, c2n=None, c1new=None, c2new=None, fsrc=None):
        ctx, lib = self.ctx, self.ctx.lib
        v2 = vel[2] if len(vel) > 2 else None
        check(lib.mp_elem_p1_transport_quad(ctx.h, C.byref(self.sp.dm.c), C.byref(self.prm), C.byref(self.quad), ptr(self.vdofs),
                                            ptr(vel[0]), ptr(vel[1]), ptr(v2), ptr(c0n), ptr(c1n), ptr(c2n), ptr(c1new), ptr(c2new),
                                            ptr(fsrc), ptr(self.emat), ptr(self.evec)), "mp_elem_p1_transport_quad")
        self.plan.scatter_matrix(self.emat, self.A)
        self.plan.scatter_vector(self.evec, self.rhs)
        self.A.apply_dirichlet(self.rhs, self.isbc, self.g)
        return self.A.solve(self.method, self.rhs, c_out, rtol=self.rtol, maxit=self.maxit)


def adr_single_params(Pe, Da, dt):
    """mp_transport_params for DarcyAdvection.advection_diffusion (mupopp.py:627-664): one component, first-order reaction Da*c^n, tau = h/(2|w|)."""
    p = transport_params(Pe, Da, 1.0, 0.0, dt, 0.0, 0.0, 0.0, (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), 1.0, "c0c1", 1, 4)
    p.reaction = 2
    return p


class TransportSolverR:
    """One reference timestep of the monolithic DarcyAdvection / CCS forms in mode R, as the block-triangular sequence
    c1,c2 mass solves -> c0 row (P2 lagged velocity, quadrature) -> (u,p) saddle solve."""

    def __init__(self, ctx: Context, mesh, model, u_isbc, u_g, c0_isbc, c0_g, f_source=None, rtol=1e-13):
        from .transport import TransportModel  # noqa: F401  (type of `model`)
        self.ctx, self.model = ctx, model
        self.sp = Spaces(ctx, mesh)
        sp = self.sp
        d = mesh.dim
        self.d = d
        self.prm = transport_params(model.Pe, model.Da, model.phi, model.alpha, model.dt, model.sigma, model.rho0, model.gamma,
                                    model.fracs, model.zhat, model.K, model.reaction, model.ncomp, model.SUPG)
        self.darcy = MixedDarcyR(ctx, sp, K=model.K, phi=model.phi, u_isbc=u_isbc, u_g=u_g, rtol=rtol)
        self.adr = AdrSolverR(ctx, sp, self.prm, c0_isbc, c0_g, vel_degree=2, rtol=rtol)
        self.rtol = rtol
        p11 = sp.plan(1, 1)
        self.p11 = p11
        self.M1 = p11.new_matrix(sell=True)
        emat = ctx.empty(sp.dm.ncell * (d + 1) ** 2)
        check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(sp.dm.c), 1.0, ptr(emat)), "mp_elem_p1_mass")
        p11.scatter_matrix(emat, self.M1)
        self.M1.jacobi_setup()
        self.evec = ctx.empty(sp.dm.ncell * (d + 1))
        self.evec2 = ctx.empty(sp.dm.ncell * (d + 1))
        n1, n2 = sp.n1, sp.n2
        self.u = [ctx.zeros(n2) for _ in range(d)]
        self.p = ctx.zeros(n1)
        self.c0, self.c1, self.c2 = ctx.zeros(n1), ctx.zeros(n1), ctx.zeros(n1)
        self.fsrc = ctx.zeros(n1) if f_source is None else ctx.to_device(np.asarray(f_source, dtype=np.float64))
        self.rhs1, self.rhs2 = ctx.zeros(n1), ctx.zeros(n1)
        self.ones1 = torch.ones(n1, dtype=torch.float64, device=ctx.device)
        self.lK = ctx.zeros(n2)
        self.darcy.G.spmv(self.ones1, self.lK)          # K int psi_i

    def set_state(self, u=None, p=None, c0=None, c1=None, c2=None):
        if u is not None:
            for c in range(self.d):
                self.u[c].copy_(self.ctx.to_device(np.asarray(u[c], dtype=np.float64)))
        for name, v in (("p", p), ("c0", c0), ("c1", c1), ("c2", c2)):
            if v is not None:
                getattr(self, name).copy_(self.ctx.to_device(np.asarray(v, dtype=np.float64)))

    def step(self):
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.sp.dm.c, self.prm
        three = self.model.ncomp == 3
        n1 = self.sp.n1
        check(lib.mp_elem_p1_reaction_rhs(ctx.h, C.byref(m), C.byref(prm), ptr(self.c0), ptr(self.c1), ptr(self.c2) if three else None,
                                          ptr(self.evec), ptr(self.evec2) if three else None), "mp_elem_p1_reaction_rhs")
        self.p11.scatter_vector(self.evec, self.rhs1)
        c1new = self.c1.clone()
        self.M1.solve("pcg", self.rhs1, c1new, rtol=self.rtol, refresh_jacobi=False)
        c2new = None
        if three:
            self.p11.scatter_vector(self.evec2, self.rhs2)
            c2new = self.c2.clone()
            self.M1.solve("pcg", self.rhs2, c2new, rtol=self.rtol, refresh_jacobi=False)
        c0new = self.c0.clone()
        r0 = self.adr.solve(self.u, self.c0, c0new, c1n=self.c1, c2n=self.c2 if three else None, c1new=c1new, c2new=c2new, fsrc=self.fsrc)
        # momentum rhs: -sigma zhat_c (rho0 lK + gamma G c0^{n+1})   (mupopp.py:807 / :1204 moved to the right-hand side)
        gc0 = ctx.zeros(self.sp.n2)
        self.darcy.G.spmv(c0new, gc0)
        rhs_u = []
        for c in range(self.d):
            rc = ctx.zeros(self.sp.n2)
            zc = float(self.model.zhat[c])
            ctx.axpby(-self.model.sigma * zc * self.model.rho0, self.lK, 0.0, rc)
            ctx.axpby(-self.model.sigma * zc * self.model.gamma, gc0, 1.0, rc)
            rhs_u.append(rc)
        u, outer = self.darcy.solve(rhs_u, self.p)
        self.u, self.c0, self.c1 = u, c0new, c1new
        if three:
            self.c2 = c2new
        return dict(c0=r0, outer=outer)


def adaptive_source(X):
    """Hard-coded source of DarcyAdvection.advection_diffusion_two_component_adaptive (mupopp.py:693-698; `3.14`, not pi),
    nodally interpolated as `Expression(..., degree=1)` is."""
    s = np.zeros(X.shape[0])
    for k in range(1, 11):
        s += 1.0 + np.sin(k * X[:, 0] * 3.14)
    return (1.0 - np.tanh(X[:, 1] / 0.01)) * s / 20.0


class TwoComponentAdaptiveR:
    """DarcyAdvection.advection_diffusion_two_component_adaptive (mupopp.py:667-749): [c0, c1] with a given velocity;
    f = Da*c0^n*c1^n, tau = 1/(4/(Pe h^2) + 2|u|/h + Da*c1^n/phi).  The DG0 projection at :704-710 is dead code (result unused)."""

    def __init__(self, ctx: Context, spaces: Spaces, Pe, Da, phi, alpha, dt, c0_isbc, c0_g, vel_degree=2, rtol=1e-13):
        self.ctx, self.sp, self.rtol = ctx, spaces, rtol
        d = spaces.mesh.dim
        self.prm = transport_params(Pe, Da, phi, alpha, dt, 0.0, 0.0, 0.0, (1.0, 1.0, 0.0), (0.0,) * d, 1.0, "c0c1", 2, 5)
        self.adr = AdrSolverR(ctx, spaces, self.prm, c0_isbc, c0_g, vel_degree=vel_degree, rtol=rtol)
        p11 = spaces.plan(1, 1)
        self.p11 = p11
        self.M1 = p11.new_matrix(sell=True)
        emat = ctx.empty(spaces.dm.ncell * (d + 1) ** 2)
        check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(spaces.dm.c), 1.0, ptr(emat)), "mp_elem_p1_mass")
        p11.scatter_matrix(emat, self.M1)
        self.M1.jacobi_setup()
        self.evec = ctx.empty(spaces.dm.ncell * (d + 1))
        self.rhs1 = ctx.zeros(spaces.n1)
        self.fsrc = ctx.to_device(adaptive_source(spaces.mesh.coords))

    def step(self, vel, c0n, c1n, c0_out, c1_out):
        ctx = self.ctx
        check(ctx.lib.mp_elem_p1_reaction_rhs(ctx.h, C.byref(self.sp.dm.c), C.byref(self.prm), ptr(c0n), ptr(c1n), None, ptr(self.evec), None),
              "mp_elem_p1_reaction_rhs")
        self.p11.scatter_vector(self.evec, self.rhs1)
        c1_out.copy_(c1n)
        self.M1.solve("pcg", self.rhs1, c1_out, rtol=self.rtol, refresh_jacobi=False)
        c0_out.copy_(c0n)
        return self.adr.solve(vel, c0n, c0_out, c1n=c1n, c1new=c1_out, fsrc=self.fsrc)
