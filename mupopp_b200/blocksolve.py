"""Block operators + device-resident MINRES (`mp_solve_minres`, csrc/minres.cu) for the saddle-point systems of the mode-R paths:
compaction momentum (mupopp.py:306-429), Stokes (mupopp.py:997-1099) and mixed Darcy (mupopp.py:595-625, :804-807).
Replaces the round-1 host-sequenced recurrence (two blocking dot products per iteration)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MpBlockOp, MpBlockPrec, MpBlockTerm, MpCsr, MpSolveInfo, check, ptr


class BlockSystem:
    """sizes: block sizes; terms: [(row block, col block, CsrMatrix, scale)]; mask: flat device vector (1 free, 0 Dirichlet) or None.
    prec: per block None (Jacobi with the given dinv) or (CsrMatrix P, degree, eigenvalue ratio lmax/lmin)."""

    def __init__(self, ctx, sizes, terms, mask=None):
        self.ctx = ctx
        self.sizes = [int(s) for s in sizes]
        self.n = sum(self.sizes)
        self.terms = list(terms)                      # keeps the matrices alive
        self.mask = mask
        nb = len(sizes)
        self._tarr = (MpBlockTerm * len(terms))()
        for i, (rb, cb, A, scale) in enumerate(terms):
            self._tarr[i].rb, self._tarr[i].cb, self._tarr[i].scale = int(rb), int(cb), float(scale)
            self._tarr[i].A = C.pointer(A.c)
        self.op = MpBlockOp()
        self.op.nblk, self.op.nterm = nb, len(terms)
        off = np.concatenate([[0], np.cumsum(self.sizes)])
        for i in range(nb + 1):
            self.op.off[i] = int(off[i])
        self.op.terms = C.cast(self._tarr, C.POINTER(MpBlockTerm))
        self.op.d_mask = mask.data_ptr() if mask is not None else None
        self.off = [int(v) for v in off]
        self.prec = None
        self.last = None

    def apply(self, x, y):
        check(self.ctx.lib.mp_block_apply(self.ctx.h, C.byref(self.op), ptr(x), ptr(y)), "mp_block_apply")
        return y

    def set_preconditioner(self, dinv, blocks=None):
        """dinv: flat inverse diagonal (1 on Dirichlet dofs).  blocks[b] = None | (P CsrMatrix, degree, ratio): Chebyshev on
        [lmax/ratio, lmax] with lmax the Gershgorin bound of D^-1 P (mp_spectral_bound)."""
        ctx = self.ctx
        pr = MpBlockPrec()
        pr.d_dinv = dinv.data_ptr()
        self._pkeep = [dinv]
        for b in range(len(self.sizes)):
            spec = blocks[b] if blocks is not None else None
            if spec is None or spec[1] <= 1:
                pr.degree[b] = 1
                continue
            P, degree, ratio = spec
            seg = P.dinv if getattr(P, "dinv", None) is not None else dinv[self.off[b]:self.off[b + 1]]     # the block's own Jacobi scaling
            lm = C.c_double()
            check(ctx.lib.mp_spectral_bound(ctx.h, C.byref(P.c), ptr(seg), C.byref(lm)), "mp_spectral_bound")
            pr.P[b] = C.pointer(P.c)
            pr.degree[b] = int(degree)
            pr.lmax[b] = 1.05 * lm.value
            pr.lmin[b] = 1.05 * lm.value / float(ratio)
            self._pkeep.append(P)
        self.prec = pr

    def minres(self, b, x, rtol=1e-6, atol=0.0, maxit=3000, check_every=0):
        if self.prec is None:
            raise RuntimeError("BlockSystem.minres: call set_preconditioner first")
        info = MpSolveInfo(float(rtol), float(atol), int(maxit), int(check_every), 0, 0, 0.0)
        check(self.ctx.lib.mp_solve_minres(self.ctx.h, C.byref(self.op), C.byref(self.prec), ptr(b), ptr(x), C.byref(info)), "mp_solve_minres")
        from .device import SolveResult
        self.last = SolveResult(info.niter, bool(info.converged), info.relres)
        return self.last
