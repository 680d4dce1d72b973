"""StokesAdvection.stokes_ADR_precipitation on the GPU: /root/reference/src/modules/mupopp.py:997-1099,
W = [P3^d, P1, P1, P1, P1] (run/stokes_ADR_3D/master_advective_stokes.py:94-102).

The Stokes block  [L, B^T; B, 0][u; p] = 0 + Dirichlet data  does not see the concentrations; it is solved by block MINRES with the
classical diag(L) + diag(pressure mass) preconditioner.  The transport rows reuse the c1/c2 mass-solve and c0 quadrature kernels with
the lagged P3 velocity; the reaction rates of this class are not divided by phi / (1-phi), which is folded into the fractions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, ptr
from .device import Context, transport_params
from .moder import AdrSolverR, Spaces

PHI_FOLD = 0.5


def stokes_adr_params(Pe, Da, dt, SUPG=2, dim=3):
    T = 698.0
    return transport_params(Pe, Da, PHI_FOLD, 0.0, dt, 0.0, 0.0, 0.0, (62.0 / T * PHI_FOLD, 278.0 / T * (1 - PHI_FOLD), 100.0 / T * (1 - PHI_FOLD)),
                            (0.0,) * dim, 1.0, "c0c1", 3, SUPG)


class StokesSolverR:
    """[L, B^T; B, 0] with Dirichlet velocity data, velocity degree `vdeg` (3 in the reference), pressure P1."""

    def __init__(self, ctx: Context, spaces: Spaces, u_isbc, u_g, vdeg=3, cheb=(8, 4), ratio=30.0):
        self.ctx, self.sp, self.vdeg = ctx, spaces, vdeg
        self.cheb, self.ratio = cheb, ratio
        d = spaces.mesh.dim
        self.d = d
        self.nv, self.n1 = spaces.ndof(vdeg), spaces.n1
        self.L = spaces.assemble("stiff", vdeg, vdeg)
        self.B = [spaces.assemble("grad", 1, vdeg, direction=a) for a in range(d)]
        self.BT = [spaces.assemble("gradT", vdeg, 1, direction=a) for a in range(d)]
        Mp = spaces.assemble("mass", 1, 1)
        self.Mp = Mp
        self.mask = [ctx.to_device(1.0 - np.asarray(u_isbc[a], dtype=np.float64)) for a in range(d)]
        self.g = [ctx.to_device(np.asarray(u_g[a], dtype=np.float64)) for a in range(d)]
        n = d * self.nv + self.n1
        self.dinv = ctx.zeros(n)
        dl = self.L.jacobi_setup()
        for a in range(d):
            seg = self.dinv[a * self.nv:(a + 1) * self.nv]
            ctx.vmul(1.0, dl, self.mask[a], seg)
            ctx.axpby(-1.0, self.mask[a], 1.0, seg)
            seg.add_(1.0)                                      # 1/diag(L) on free dofs, 1 on Dirichlet dofs
        self.dinv[d * self.nv:].copy_(Mp.jacobi_setup())
        self.tv, self.t1 = ctx.zeros(self.nv), ctx.zeros(self.n1)
        self.xm = ctx.zeros(n)

    def _views(self, buf):
        nv, d = self.nv, self.d
        return [buf[a * nv:(a + 1) * nv] for a in range(d)], buf[d * nv:]

    def _apply_raw(self, x, y):
        ctx = self.ctx
        xu, xp = self._views(x)
        yu, yp = self._views(y)
        yp.zero_()
        for a in range(self.d):
            self.L.spmv(xu[a], yu[a])
            self.BT[a].spmv(xp, self.tv)
            ctx.axpby(1.0, self.tv, 1.0, yu[a])
            self.B[a].spmv(xu[a], self.t1)
            ctx.axpby(1.0, self.t1, 1.0, yp)

    def apply(self, x, y):
        ctx = self.ctx
        self.xm.copy_(x)
        xu, _ = self._views(x)
        mu, _ = self._views(self.xm)
        for a in range(self.d):
            ctx.vmul(1.0, self.mask[a], xu[a], mu[a])
        self._apply_raw(self.xm, y)
        yu, _ = self._views(y)
        for a in range(self.d):
            ctx.vmul(1.0, self.mask[a], yu[a], yu[a])
            ctx.axpby(1.0, xu[a], 1.0, yu[a])
            ctx.axpby(-1.0, mu[a], 1.0, yu[a])

    def _block_system(self):
        from .blocksolve import BlockSystem
        if getattr(self, "_bs", None) is None:
            d = self.d
            terms = []
            for a in range(d):
                terms += [(a, a, self.L, 1.0), (a, d, self.BT[a], 1.0), (d, a, self.B[a], 1.0)]
            self._maskflat = self.ctx.zeros(d * self.nv + self.n1)
            self._maskflat.fill_(1.0)
            for a in range(d):
                self._maskflat[a * self.nv:(a + 1) * self.nv].copy_(self.mask[a])
            self._bs = BlockSystem(self.ctx, [self.nv] * d + [self.n1], terms, mask=self._maskflat)
            self._bs.set_preconditioner(self.dinv, [(self.L, self.cheb[0], self.ratio)] * d + [(self.Mp, self.cheb[1], self.ratio)])
        return self._bs

    def solve(self, x, rtol=1e-12, maxit=50000):
        """x = [u_0..u_{d-1}, p] flat device tensor: initial guess in, solution out.  Returns MINRES iterations (device-resident
        mp_solve_minres; block preconditioner diag(Chebyshev(L), Chebyshev(pressure mass)))."""
        ctx = self.ctx
        n = x.numel()
        gfull, b = ctx.zeros(n), ctx.zeros(n)
        gu, _ = self._views(gfull)
        for a in range(self.d):
            gu[a].copy_(self.g[a])
        self._apply_raw(gfull, b)
        ctx.axpby(0.0, b, -1.0, b)                               # b = -A g   (homogeneous Stokes: no body force in mupopp.py:1044)
        bu, _ = self._views(b)
        xu, _ = self._views(x)
        for a in range(self.d):
            ctx.vmul(1.0, self.mask[a], bu[a], bu[a])
            ctx.axpby(1.0, self.g[a], 1.0, bu[a])
            ctx.vmul(1.0, self.mask[a], xu[a], xu[a])
            ctx.axpby(1.0, self.g[a], 1.0, xu[a])                # x[D] = g
        res = self._block_system().minres(b, x, rtol=rtol, maxit=maxit)
        self.last_relres, self.last_result = res.relres, res
        return res.niter


class StokesAdrR:
    """One timestep of stokes_ADR_precipitation: c1,c2 mass solves -> c0 row with the lagged P3 velocity -> Stokes solve."""

    def __init__(self, ctx: Context, mesh, Pe, Da, dt, SUPG, u_isbc, u_g, c0_isbc, c0_g, vdeg=3, rtol=1e-13):
        self.ctx, self.vdeg, self.rtol = ctx, vdeg, rtol
        self.sp = Spaces(ctx, mesh)
        sp = self.sp
        d = mesh.dim
        self.d = d
        self.prm = stokes_adr_params(Pe, Da, dt, SUPG, d)
        self.stokes = StokesSolverR(ctx, sp, u_isbc, u_g, vdeg)
        self.adr = AdrSolverR(ctx, sp, self.prm, c0_isbc, c0_g, vel_degree=vdeg, rtol=rtol)
        self.p11 = sp.plan(1, 1)
        self.M1 = sp.assemble("mass", 1, 1)
        self.M1.jacobi_setup()
        self.evec, self.evec2 = ctx.empty(sp.dm.ncell * (d + 1)), ctx.empty(sp.dm.ncell * (d + 1))
        nv, n1 = sp.ndof(vdeg), sp.n1
        self.up = ctx.zeros(d * nv + n1)                       # [u, p]
        self.c0, self.c1, self.c2 = ctx.zeros(n1), ctx.zeros(n1), ctx.zeros(n1)
        self.rhs1, self.rhs2 = ctx.zeros(n1), ctx.zeros(n1)
        self.zero1 = ctx.zeros(n1)

    def u_views(self):
        nv = self.sp.ndof(self.vdeg)
        return [self.up[a * nv:(a + 1) * nv] for a in range(self.d)]

    @property
    def p(self):
        return self.up[self.d * self.sp.ndof(self.vdeg):]

    def step(self):
        ctx, lib, m = self.ctx, self.ctx.lib, self.sp.dm.c
        check(lib.mp_elem_p1_reaction_rhs(ctx.h, C.byref(m), C.byref(self.prm), ptr(self.c0), ptr(self.c1), ptr(self.c2), ptr(self.evec), ptr(self.evec2)),
              "mp_elem_p1_reaction_rhs")
        self.p11.scatter_vector(self.evec, self.rhs1)
        self.p11.scatter_vector(self.evec2, self.rhs2)
        c1new, c2new = self.c1.clone(), self.c2.clone()
        self.M1.solve("pcg", self.rhs1, c1new, rtol=self.rtol, refresh_jacobi=False)
        self.M1.solve("pcg", self.rhs2, c2new, rtol=self.rtol, refresh_jacobi=False)
        c0new = self.c0.clone()
        r0 = self.adr.solve(self.u_views(), self.c0, c0new, c1n=self.c1, c2n=self.c2, c1new=c1new, c2new=c2new, fsrc=self.zero1)
        its = self.stokes.solve(self.up, rtol=self.rtol)
        self.c0, self.c1, self.c2 = c0new, c1new, c2new
        return dict(c0=r0, stokes_its=its)
