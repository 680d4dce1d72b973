"""Reference-element tables for the GPU element kernels: simplex quadrature, Lagrange P1/P2 tabulation,
and the reference tensors consumed by `mp_elem_reftensor`.

This replaces what FFC generates at JIT time for the reference (FIAT tabulation + quadrature, logged e.g. at
/root/reference/src/run/sphere_archer/sphere.out:195).  Only small constant tables are built here (host, once);
all per-cell arithmetic happens in the CUDA kernels.  UFC local dof order: vertices, then edges with
triangle edge k opposite vertex k and the UFC tetrahedron edge table.
"""
from __future__ import annotations

from functools import lru_cache
from math import factorial

import numpy as np

TRI_EDGES = ((1, 2), (0, 2), (0, 1))
TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def gauss_jacobi(n, alpha):
    """n-point Gauss-Jacobi rule for weight (1-x)^alpha on [-1,1] (beta = 0) by Golub-Welsch."""
    k = np.arange(n, dtype=float)
    a = alpha
    # three-term recurrence of Jacobi polynomials P_k^{(a,0)}
    with np.errstate(divide="ignore", invalid="ignore"):
        diag = np.where(k == 0, -a / (a + 2.0), -(a * a) / ((2 * k + a) * (2 * k + a + 2.0)))
    kk = k[1:]
    off = 2.0 / (2 * kk + a) * np.sqrt(kk * (kk + a) * kk * (kk + a) / ((2 * kk + a - 1.0) * (2 * kk + a + 1.0)))
    J = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
    x, V = np.linalg.eigh(J)
    mu0 = 2.0 ** (a + 1.0) / (a + 1.0)
    w = mu0 * V[0, :] ** 2
    return x, w


@lru_cache(maxsize=None)
def simplex_quadrature(dim, degree):
    """Collapsed Gauss-Jacobi rule on the reference simplex exact to total degree `degree`
    (n = degree//2 + 1 points per direction); returns (points (nq,dim), weights (nq,)), sum(w) = 1/dim!."""
    n = degree // 2 + 1

    def gj(alpha):
        x, w = gauss_jacobi(n, alpha)
        return 0.5 * (x + 1.0), w * 0.5 ** (alpha + 1)

    if dim == 2:
        r, wr = gj(1.0)
        s, ws = gj(0.0)
        R, S = np.meshgrid(r, s, indexing="ij")
        W = wr[:, None] * ws[None, :]
        return np.stack([R.ravel(), (S * (1 - R)).ravel()], 1), W.ravel()
    r, wr = gj(2.0)
    s, ws = gj(1.0)
    t, wt = gj(0.0)
    R, S, T = np.meshgrid(r, s, t, indexing="ij")
    W = wr[:, None, None] * ws[None, :, None] * wt[None, None, :]
    return np.stack([R.ravel(), (S * (1 - R)).ravel(), (T * (1 - R) * (1 - S)).ravel()], 1), W.ravel()


def barycentric(pts):
    return np.concatenate([1.0 - pts.sum(1, keepdims=True), pts], axis=1)


def tabulate(degree, pts):
    """Values (nq, nb) and reference gradients (nq, nb, d) of the Lagrange basis of `degree` in {1, 2}."""
    nq, d = pts.shape
    lam = barycentric(pts)
    dlam = np.concatenate([-np.ones((1, d)), np.eye(d)], axis=0)
    if degree == 1:
        return lam.copy(), np.broadcast_to(dlam, (nq, d + 1, d)).copy()
    edges = TRI_EDGES if d == 2 else TET_EDGES
    if degree == 3:
        faces = [(0, 1, 2)] if d == 2 else [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
        nb = d + 1 + 2 * len(edges) + len(faces)
        phi = np.empty((nq, nb))
        dphi = np.empty((nq, nb, d))
        for i in range(d + 1):
            l = lam[:, i]
            phi[:, i] = 0.5 * l * (3.0 * l - 1.0) * (3.0 * l - 2.0)
            dphi[:, i] = (13.5 * l * l - 9.0 * l + 1.0)[:, None] * dlam[i]
        k = d + 1
        for (a, b) in edges:
            for (ia, ib) in ((a, b), (b, a)):                 # the dof nearer the first (lower) vertex comes first
                p_, q_ = lam[:, ia], lam[:, ib]
                phi[:, k] = 4.5 * p_ * q_ * (3.0 * p_ - 1.0)
                dphi[:, k] = 4.5 * ((6.0 * p_ * q_ - q_)[:, None] * dlam[ia] + (3.0 * p_ * p_ - p_)[:, None] * dlam[ib])
                k += 1
        for (a, b, c) in faces:
            la, lb, lc = lam[:, a], lam[:, b], lam[:, c]
            phi[:, k] = 27.0 * la * lb * lc
            dphi[:, k] = 27.0 * ((lb * lc)[:, None] * dlam[a] + (la * lc)[:, None] * dlam[b] + (la * lb)[:, None] * dlam[c])
            k += 1
        return phi, dphi
    if degree != 2:
        raise NotImplementedError("Lagrange degree %d" % degree)
    nb = d + 1 + len(edges)
    phi = np.empty((nq, nb))
    dphi = np.empty((nq, nb, d))
    for i in range(d + 1):
        phi[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dphi[:, i] = (4.0 * lam[:, i] - 1.0)[:, None] * dlam[i]
    for k, (a, b) in enumerate(edges):
        phi[:, d + 1 + k] = 4.0 * lam[:, a] * lam[:, b]
        dphi[:, d + 1 + k] = 4.0 * (lam[:, a][:, None] * dlam[b] + lam[:, b][:, None] * dlam[a])
    return phi, dphi


def nbasis(degree, dim):
    return {1: dim + 1, 2: 6 if dim == 2 else 10, 3: 10 if dim == 2 else 20}[degree]


# ---- reference tensors for mp_elem_reftensor: A_e[i,j] = sum_t geo_t(cell) * T0[t][i][j] -------------------------------
#   MASS : geo = |detJ|                         T0[0][i][j]    = int_ref r_i c_j
#   GRAD : geo_a = |detJ| Jinv[a][dir]          T0[a][i][j]    = int_ref r_i d(c_j)/dxi_a        (row basis value, column basis derivative)
#   GRADT: geo_a = |detJ| Jinv[a][dir]          T0[a][i][j]    = int_ref d(r_i)/dxi_a c_j
#   STIFF: geo_ab = |detJ| (Jinv Jinv^T)[a][b]  T0[a*d+b][i][j] = int_ref d(r_i)/dxi_a d(c_j)/dxi_b
# with a nodal P1 coefficient the tensor gets one more leading index v (barycentric weight lambda_v), exact integration.

@lru_cache(maxsize=None)
def reference_tensor(kind, dim, deg_r, deg_c, p1coef=False):
    qdeg = (deg_r + deg_c) + (1 if p1coef else 0)
    pts, w = simplex_quadrature(dim, max(qdeg, 1))
    pr, dr = tabulate(deg_r, pts)
    pc, dc = tabulate(deg_c, pts)
    lam = barycentric(pts)
    if kind == "mass":
        T = np.einsum("q,qi,qj->ij", w, pr, pc)[None]
    elif kind == "grad":
        T = np.einsum("q,qi,qja->aij", w, pr, dc)
    elif kind == "gradT":
        T = np.einsum("q,qia,qj->aij", w, dr, pc)
    elif kind == "stiff":
        T = np.einsum("q,qia,qjb->abij", w, dr, dc).reshape(dim * dim, pr.shape[1], pc.shape[1])
    else:
        raise ValueError(kind)
    if p1coef:
        if kind == "mass":
            T = np.einsum("q,qv,qi,qj->vij", w, lam, pr, pc)[None].transpose(0, 1, 2, 3).reshape(dim + 1, pr.shape[1], pc.shape[1])
        elif kind == "grad":
            T = np.einsum("q,qv,qi,qja->avij", w, lam, pr, dc).reshape(dim * (dim + 1), pr.shape[1], pc.shape[1])
        elif kind == "gradT":
            T = np.einsum("q,qv,qia,qj->avij", w, lam, dr, pc).reshape(dim * (dim + 1), pr.shape[1], pc.shape[1])
        else:
            T = np.einsum("q,qv,qia,qjb->abvij", w, lam, dr, dc).reshape(dim * dim * (dim + 1), pr.shape[1], pc.shape[1])
    return np.ascontiguousarray(T)


def quadrature_tables(dim, qdeg, vel_degree):
    """Tables for the quadrature transport kernel: weights (nq,), barycentric coords (nq,d+1), velocity basis values (nq,nbv)."""
    pts, w = simplex_quadrature(dim, qdeg)
    lam = barycentric(pts)
    if vel_degree == 0:
        vphi = np.ones((len(w), 1))
    else:
        vphi, _ = tabulate(vel_degree, pts)
    return np.ascontiguousarray(w), np.ascontiguousarray(lam), np.ascontiguousarray(vphi)


def basis_gradient_table(dim, qdeg, degree):
    """Reference gradients (nq, nb, dim) of the Lagrange basis at the points of the degree-`qdeg` rule."""
    pts, _ = simplex_quadrature(dim, qdeg)
    return np.ascontiguousarray(tabulate(degree, pts)[1])
