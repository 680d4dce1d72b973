"""compaction forms on the GPU: split-field momentum system [u (P2^d), p (P1), c (P1)] and the
CN + SUPG porosity row, for /root/reference/src/modules/mupopp.py:306-429 and :217-304.

Assembly is block-wise: the viscous block 0.5(1-phi) symgrad:symgrad and the divergence blocks come from the generic reference-tensor
kernel (exact for the P1 coefficient 1-phi); the blocks whose coefficients are non-polynomial in phi (dL*phi*phi^(1/phi^3), dL/phi,
eta(phi)) and the right-hand side use the shared degree-8 quadrature rule.  The symmetric indefinite system is solved by MINRES
preconditioned with the diagonal of the reference's own preconditioner form `b` (mupopp.py:379-380) -- the reference applies
hypre AMG to that form; AMG is outside this build's scope.  The Krylov recurrence is driven from the host through the
C-ABI BLAS-1 kernels because the operator is a composition of block SpMVs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import reference_element as ref
from ._lib import MpCompactionParams, MpQuad, check, ptr
from .device import Context
from .moder import Spaces

QDEG_COMPACTION = 8      # shared with oracle/compaction.py


def _quad(ctx, dim, qdeg, with_grad=False):
    w, lam, vphi = ref.quadrature_tables(dim, qdeg, 2)
    t = dict(w=ctx.to_device(w), lam=ctx.to_device(lam.ravel()), vphi=ctx.to_device(vphi.ravel()))
    dv = None
    if with_grad:
        t["vdphi"] = ctx.to_device(ref.basis_gradient_table(dim, qdeg, 2).ravel())
        dv = t["vdphi"].data_ptr()
    q = MpQuad(len(w), vphi.shape[1], t["w"].data_ptr(), t["lam"].data_ptr(), t["vphi"].data_ptr(), dv)
    return q, t


class BlockVector:
    """[u_0 .. u_{d-1} | p | c] in one contiguous device buffer with per-block views."""

    def __init__(self, ctx, d, n2, n1, buf="new"):
        self.d, self.n2, self.n1 = d, n2, n1
        if buf == "new":
            self.rebind(ctx.zeros(d * n2 + 2 * n1))

    def rebind(self, buf):
        d, n2, n1 = self.d, self.n2, self.n1
        self.buf = buf
        self.u = [buf[a * n2:(a + 1) * n2] for a in range(d)]
        self.p = buf[d * n2:d * n2 + n1]
        self.c = buf[d * n2 + n1:]


class CompactionMomentumR:
    def __init__(self, ctx: Context, spaces: Spaces, da, R, B, theta, dL, zhat, u_isbc, u_g, qdeg=QDEG_COMPACTION):
        self.ctx, self.sp = ctx, spaces
        d = spaces.mesh.dim
        self.d = d
        self.prm = MpCompactionParams(float(da), float(R), float(B), float(theta), float(dL), (C.c_double * 3)(*[float(zhat[i]) if i < len(zhat) else 0.0 for i in range(3)]))
        self.dL = float(dL)
        self.quad, self._qt = _quad(ctx, d, qdeg)
        n1, n2 = spaces.n1, spaces.n2
        self.n1, self.n2 = n1, n2
        self.Bdiv = [spaces.assemble("grad", 1, 2, direction=a) for a in range(d)]        # int q d_a psi   (constant)
        self.BdivT = [spaces.assemble("gradT", 2, 1, direction=a) for a in range(d)]
        self.mask = [ctx.to_device(1.0 - np.asarray(u_isbc[a], dtype=np.float64)) for a in range(d)]   # 0 on Dirichlet dofs
        self.isbc = [ctx.to_device(np.asarray(u_isbc[a], dtype=np.uint8), torch.uint8) for a in range(d)]
        self.g = [ctx.to_device(np.asarray(u_g[a], dtype=np.float64)) for a in range(d)]
        self.p22, self.p11 = spaces.plan(2, 2), spaces.plan(1, 1)
        self.Auu = [[self.p22.new_matrix(sell=True) for _ in range(d)] for _ in range(d)]
        self.App, self.Acc = self.p11.new_matrix(sell=True), self.p11.new_matrix(sell=True)
        self.Puu, self.Ppp, self.Pcc = self.p22.new_matrix(sell=False), self.p11.new_matrix(sell=False), self.p11.new_matrix(sell=False)
        self.M1 = spaces.assemble("mass", 1, 1)
        self.omp = ctx.zeros(n1)
        self.ones1 = torch.ones(n1, dtype=torch.float64, device=ctx.device)
        self.ccoef = ctx.zeros(spaces.dm.ncell)
        self.emat2 = ctx.empty(spaces.dm.ncell * spaces.nb2 * spaces.nb2)
        self.emat1 = ctx.empty(spaces.dm.ncell * spaces.nb1 * spaces.nb1)
        self.evu = [ctx.empty(spaces.dm.ncell * spaces.nb2) for _ in range(d)]
        self.evp = ctx.empty(spaces.dm.ncell * spaces.nb1)
        self.T_stiff22 = spaces.tensor("stiff", 2, 2, True)
        self.T_stiff11 = spaces.tensor("stiff", 1, 1, False)
        self.rhs = BlockVector(ctx, d, n2, n1)
        self.dinv = BlockVector(ctx, d, n2, n1)
        self.t2, self.t1, self.s1 = ctx.zeros(n2), ctx.zeros(n1), ctx.zeros(n1)
        self.xm = BlockVector(ctx, d, n2, n1)

    # ---- assembly -------------------------------------------------------------
    def assemble(self, phi, gam, buoy):
        ctx, lib, sp, m, d = self.ctx, self.ctx.lib, self.sp, self.sp.dm.c, self.d
        nb1, nb2 = sp.nb1, sp.nb2
        self.omp.copy_(self.ones1)
        ctx.axpby(-1.0, phi, 1.0, self.omp)                                   # 1 - phi (nodal: exact P1 coefficient)
        for a in range(d):
            for b in range(d):
                self.p22.reftensor_scatter(sp.dm, 4, a * d + b, nb2, nb2, self.T_stiff22, self.omp, 1, 0.25, self.emat2, self.Auu[a][b])
        # preconditioner viscous block (1-phi) grad u:grad v
        self.p22.reftensor_scatter(sp.dm, 3, 0, nb2, nb2, self.T_stiff22, self.omp, 1, 1.0, self.emat2, self.Puu)
        # (q,p): dL phi phi^(1/phi^3) grad p.grad q ;  P: dL/phi grad p.grad q + p q
        check(lib.mp_cell_fn(ctx.h, C.byref(m), C.byref(self.quad), 0, self.dL, ptr(phi), ptr(self.ccoef)), "mp_cell_fn")
        self.p11.reftensor_scatter(sp.dm, 3, 0, nb1, nb1, self.T_stiff11, self.ccoef, 2, 1.0, self.emat1, self.App)
        check(lib.mp_cell_fn(ctx.h, C.byref(m), C.byref(self.quad), 1, self.dL, ptr(phi), ptr(self.ccoef)), "mp_cell_fn")
        self.p11.reftensor_scatter(sp.dm, 3, 0, nb1, nb1, self.T_stiff11, self.ccoef, 2, 1.0, self.emat1, self.Ppp)
        ctx.axpby(1.0, self.M1.vals, 1.0, self.Ppp.vals)
        # (omega,c): -eta c omega ; P: 0.5 eta c omega
        check(lib.mp_elem_p1_mass_fn(ctx.h, C.byref(m), C.byref(self.quad), 2, self.dL, -1.0, ptr(phi), ptr(self.emat1)), "mp_elem_p1_mass_fn")
        self.p11.scatter_matrix(self.emat1, self.Acc)
        self.Pcc.vals.copy_(self.Acc.vals)
        ctx.axpby(0.0, self.Acc.vals, -0.5, self.Pcc.vals)
        # rhs
        check(lib.mp_elem_compaction_rhs(ctx.h, C.byref(m), C.byref(self.quad), C.byref(self.prm), ptr(phi), ptr(gam), ptr(buoy),
                                         ptr(self.evu[0]), ptr(self.evu[1]), ptr(self.evu[2]) if d == 3 else None, ptr(self.evp)), "mp_elem_compaction_rhs")
        for a in range(d):
            self.p22.scatter_vector(self.evu[a], self.rhs.u[a])
        self.p11.scatter_vector(self.evp, self.rhs.p)
        self.rhs.c.zero_()
        # Jacobi of the preconditioner form, Dirichlet rows -> 1
        du = self.Puu.jacobi_setup()
        for a in range(d):
            ctx.vmul(1.0, du, self.mask[a], self.dinv.u[a])                   # dinv on free dofs, 0 on Dirichlet dofs ...
            ctx.axpby(-1.0, self.mask[a], 1.0, self.dinv.u[a])                # ... minus mask, plus 1  => 1 on Dirichlet dofs
            self.dinv.u[a].add_(1.0)
        self.dinv.p.copy_(self.Ppp.jacobi_setup())
        self.dinv.c.copy_(self.Pcc.jacobi_setup())

    # ---- operator ---------------------------------------------------------------
    def _apply_raw(self, x: BlockVector, y: BlockVector):
        ctx, d = self.ctx, self.d
        self.s1.copy_(x.p)
        ctx.axpby(1.0, x.c, 1.0, self.s1)                                     # p + c multiplies div v
        y.p.zero_()
        for a in range(d):
            ya = y.u[a]
            self.BdivT[a].spmv(self.s1, self.t2)
            ctx.axpby(-1.0, self.t2, 0.0, ya)                                 # -(p + c) div v
            for b in range(d):
                self.Auu[a][b].spmv(x.u[b], self.t2)
                ctx.axpby(1.0, self.t2, 1.0, ya)
            self.Bdiv[a].spmv(x.u[a], self.t1)
            ctx.axpby(-1.0, self.t1, 1.0, y.p)
        y.c.copy_(y.p)
        self.App.spmv(x.p, self.t1)
        ctx.axpby(1.0, self.t1, 1.0, y.p)
        self.Acc.spmv(x.c, self.t1)
        ctx.axpby(1.0, self.t1, 1.0, y.c)

    def apply(self, x: BlockVector, y: BlockVector):
        """y = (P A P + (I - P)) x with P zeroing the Dirichlet velocity dofs (== symmetric elimination of assemble_system)."""
        ctx, d = self.ctx, self.d
        self.xm.buf.copy_(x.buf)
        for a in range(d):
            ctx.vmul(1.0, self.mask[a], x.u[a], self.xm.u[a])
        self._apply_raw(self.xm, y)
        for a in range(d):
            ctx.vmul(1.0, self.mask[a], y.u[a], y.u[a])
            ctx.axpby(1.0, x.u[a], 1.0, y.u[a])
            ctx.axpby(-1.0, self.xm.u[a], 1.0, y.u[a])                        # + (1 - mask) x

    # ---- MINRES -------------------------------------------------------------------
    def _block_system(self):
        """The momentum system as an mp_block_op: blocks [u_0 .. u_{d-1}, p, c] (the same terms `_apply_raw` sequences by hand)."""
        from .blocksolve import BlockSystem
        ctx, d = self.ctx, self.d
        if getattr(self, "_bs", None) is None:
            terms = []
            for a in range(d):
                for b in range(d):
                    terms.append((a, b, self.Auu[a][b], 1.0))
                terms.append((a, d, self.BdivT[a], -1.0))           # -(p + c) div v
                terms.append((a, d + 1, self.BdivT[a], -1.0))
                terms.append((d, a, self.Bdiv[a], -1.0))            # -q div u, -omega div u
                terms.append((d + 1, a, self.Bdiv[a], -1.0))
            terms.append((d, d, self.App, 1.0))
            terms.append((d + 1, d + 1, self.Acc, 1.0))
            self._maskflat = ctx.zeros(d * self.n2 + 2 * self.n1)
            self._maskflat.fill_(1.0)
            for a in range(d):
                self._maskflat[a * self.n2:(a + 1) * self.n2].copy_(self.mask[a])
            self._bs = BlockSystem(ctx, [self.n2] * d + [self.n1, self.n1], terms, mask=self._maskflat)
        return self._bs

    def solve(self, x: BlockVector, rtol=1e-6, maxit=3000, cheb=(8, 4, 4), ratio=30.0, precond="mass"):
        """Device-resident preconditioned MINRES (mp_solve_minres), block-diagonal preconditioner applied with Chebyshev-Jacobi
        polynomials of degrees `cheb` = (velocity, pressure, compaction); (1, 1, 1) is plain Jacobi.
          precond="form_b": the reference's own preconditioner form `b` (mupopp.py:379-380): (1-phi) grad u:grad v, dL/phi grad p.grad q + p q,
                            0.5 eta c omega.  Its pressure block is a LAPLACIAN, while the pressure Schur complement of this system is
                            mass-like (the Darcy term dL*phi*phi^(1/phi^3) of mupopp.py:365 underflows to 0 for phi < 1): iteration counts
                            grow like 1/h whatever solves the blocks (the reference gives the form to hypre AMG).
          precond="mass"  (default): velocity block as in `b`; pressure and compaction blocks = the operator's own diagonal blocks plus the
                            viscosity-scaled pressure mass matrix 2 M (the Stokes Schur-complement approximation B A^-1 B^T ~ M / mu,
                            mu = (1-phi)/2): mesh-independent on the multiplier blocks.  Same linear system, same solution.
        x: initial guess in, solution out.  Returns the iteration count (mupopp.py:424-429)."""
        ctx, d = self.ctx, self.d
        n2, n1 = self.n2, self.n1
        # right-hand side with Dirichlet lifting: b <- mask (b - A g), b[D] = g
        gfull, b, tmp = BlockVector(ctx, d, n2, n1), BlockVector(ctx, d, n2, n1), BlockVector(ctx, d, n2, n1)
        for a in range(d):
            gfull.u[a].copy_(self.g[a])
        self._apply_raw(gfull, tmp)
        b.buf.copy_(self.rhs.buf)
        ctx.axpby(-1.0, tmp.buf, 1.0, b.buf)
        for a in range(d):
            ctx.vmul(1.0, self.mask[a], b.u[a], b.u[a])
            ctx.axpby(1.0, self.g[a], 1.0, b.u[a])
            self.Auu[a][a].apply_dirichlet(x.u[a], self.isbc[a], self.g[a], matrix=False)     # x[D] = g
        bs = self._block_system()
        if precond == "mass":
            if getattr(self, "Spp", None) is None:
                self.Spp, self.Scc = self.p11.new_matrix(sell=False), self.p11.new_matrix(sell=False)
                self.dinv_mass = BlockVector(ctx, d, n2, n1)
            self.Spp.vals.copy_(self.App.vals)
            ctx.axpby(2.0, self.M1.vals, 1.0, self.Spp.vals)             # App + M / mu
            self.Scc.vals.copy_(self.M1.vals)
            ctx.axpby(-1.0, self.Acc.vals, 2.0, self.Scc.vals)            # eta M + M / mu   (Acc = -eta M)
            self.dinv_mass.buf.copy_(self.dinv.buf)
            self.dinv_mass.p.copy_(self.Spp.jacobi_setup())
            self.dinv_mass.c.copy_(self.Scc.jacobi_setup())
            bs.set_preconditioner(self.dinv_mass.buf, [(self.Puu, cheb[0], ratio)] * d + [(self.Spp, cheb[1], 10.0), (self.Scc, cheb[2], 10.0)])
        else:
            bs.set_preconditioner(self.dinv.buf, [(self.Puu, cheb[0], ratio)] * d + [(self.Ppp, cheb[1], ratio), (self.Pcc, cheb[2], ratio)])
        res = bs.minres(b.buf, x.buf, rtol=rtol, maxit=maxit)
        self.last_relres, self.last_result = res.relres, res
        return res.niter


class PorosityR:
    """compaction.mass_conservation / mass_solver (mupopp.py:217-304) with X = P1 or P2 (`degree`): assemble by quadrature, solve with
    Jacobi-GMRES (the reference runs MINRES on this non-symmetric matrix to rtol 1e-6).  With X = P2 (soliton,
    sphere and cavity benchmarks) phi0 and gam are P2 fields and the SUPG residual keeps its -div(grad(phi)) term."""

    def __init__(self, ctx: Context, spaces: Spaces, da, qdeg=QDEG_COMPACTION, degree=1):
        if degree not in (1, 2):
            raise NotImplementedError("porosity space of degree %d" % degree)
        self.ctx, self.sp, self.da, self.degree = ctx, spaces, float(da), degree
        self.quad, self._qt = _quad(ctx, spaces.mesh.dim, qdeg, with_grad=True)
        self.plan = spaces.plan(degree, degree)
        self.A = self.plan.new_matrix(sell=True)
        l = spaces.nb(degree)
        self.emat = ctx.empty(spaces.dm.ncell * l * l)
        self.evec = ctx.empty(spaces.dm.ncell * l)
        self.rhs = ctx.zeros(spaces.ndof(degree))

    def solve(self, phi0, u, dt, gam, phi1, rtol=1e-6, maxit=100, method="gmres", restart=50):
        ctx = self.ctx
        u2 = u[2] if len(u) > 2 else None
        if self.degree == 1:
            check(ctx.lib.mp_elem_p1_porosity_quad(ctx.h, C.byref(self.sp.dm.c), C.byref(self.quad), self.da, float(dt), ptr(self.sp.cd2),
                                                   ptr(u[0]), ptr(u[1]), ptr(u2), ptr(phi0), ptr(gam), ptr(self.emat), ptr(self.evec)),
                  "mp_elem_p1_porosity_quad")
        else:
            check(ctx.lib.mp_elem_p2_porosity_quad(ctx.h, C.byref(self.sp.dm.c), C.byref(self.quad), self.da, float(dt), ptr(self.sp.cd2),
                                                   ptr(u[0]), ptr(u[1]), ptr(u2), ptr(phi0), ptr(gam), ptr(self.emat), ptr(self.evec)),
                  "mp_elem_p2_porosity_quad")
        self.plan.scatter_matrix(self.emat, self.A)
        self.plan.scatter_vector(self.evec, self.rhs)
        phi1.copy_(phi0)
        return self.A.solve(method, self.rhs, phi1, rtol=rtol, maxit=maxit, restart=restart)
