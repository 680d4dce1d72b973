"""Preconditioned MINRES driven from the host through the C-ABI BLAS-1 kernels, for operators that are compositions of block SpMVs
(compaction momentum system, Stokes).  Replaces dolfin.KrylovSolver("minres", "amg") (/root/reference/src/modules/mupopp.py:410-425).
Algorithm: Elman, Silvester & Wathen, alg. 4.1 (Lanczos + Givens), M^-1 = a positive diagonal.  Two synchronising dot products per iteration.
"""
from __future__ import annotations

import numpy as np


def minres(ctx, apply, dinv, b, x, rtol=1e-6, maxit=3000):
    """Solve A x = b for symmetric (possibly indefinite) A.  apply(x, y) writes y = A x; dinv, b, x: flat fp64 device tensors.
    Stops when the M^-1-norm of the residual is <= rtol * ||b||_{M^-1}.  Returns (iterations, relative residual)."""
    n = x.numel()
    new = lambda: ctx.zeros(n)
    v_old, v, v_new, z, z_new, w_old, w, w_new, Az, tmp = (new() for _ in range(10))
    apply(x, Az)
    v.copy_(b)
    ctx.axpby(-1.0, Az, 1.0, v)
    ctx.vmul(1.0, dinv, v, z)
    gamma = np.sqrt(max(ctx.dot(z, v), 0.0))
    ctx.vmul(1.0, dinv, b, tmp)
    bnorm = np.sqrt(max(ctx.dot(tmp, b), 1e-300))
    eta, s_old, s, c_old, c, gamma_old = gamma, 0.0, 0.0, 1.0, 1.0, 1.0
    it = 0
    while it < maxit and abs(eta) > rtol * bnorm and gamma > 0.0:
        ctx.axpby(1.0 / gamma, z, 0.0, z)                 # z <- z / gamma
        apply(z, Az)
        delta = ctx.dot(Az, z)
        v_new.copy_(Az)
        ctx.axpby(-delta / gamma, v, 1.0, v_new)
        ctx.axpby(-gamma / gamma_old, v_old, 1.0, v_new)
        ctx.vmul(1.0, dinv, v_new, z_new)
        gamma_new = np.sqrt(max(ctx.dot(z_new, v_new), 0.0))
        a0 = c * delta - c_old * s * gamma
        a1 = np.sqrt(a0 * a0 + gamma_new * gamma_new)
        a2 = s * delta + c_old * c * gamma
        a3 = s_old * gamma
        c_new, s_new = a0 / a1, gamma_new / a1
        w_new.copy_(z)
        ctx.axpby(-a3, w_old, 1.0, w_new)
        ctx.axpby(-a2, w, 1.0, w_new)
        ctx.axpby(1.0 / a1, w_new, 0.0, w_new)
        ctx.axpby(c_new * eta, w_new, 1.0, x)
        eta = -s_new * eta
        v_old, v, v_new = v, v_new, v_old
        w_old, w, w_new = w, w_new, w_old
        z, z_new = z_new, z
        gamma_old, gamma = gamma, gamma_new
        s_old, s, c_old, c = s, s_new, c, c_new
        it += 1
    return it, abs(eta) / bnorm
