"""Oracle for the scalar functionals the reference scripts evaluate every output step.
TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restated call sites (DOLFIN `assemble(<scalar form>)` / `errornorm`):
    flux1 = assemble(c00*phi0*dot(u0, n)*ds(1))                         /root/reference/src/run/CCS2D/CCS_3comp.py:277
    iflux1 = assemble(c00*dot(u, unitNormal)*ds(1)); assemble(c02*dx); assemble(CaCO3_frac*Da0*c00*c01*dx)
                                                                        /root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:232-265
    l2_u = errornorm(u, uh, 'L2')                                       /root/reference/src/benchmark/sphere_3D_pure_shear/sphere.py:157-159
Plain NumPy, quadrature exact for the polynomial degree of each integrand.
"""
from __future__ import annotations

import numpy as np

from . import fem


def volume_integral(asm, factors, coef=1.0):
    """coef * int prod_k f_k dx;  factors: list of (dof vector, Lagrange degree)."""
    qdeg = max(sum(d for _, d in factors), 1)
    w = asm.quad(qdeg)[1]
    val = np.ones((asm.mesh.nc, len(w)))
    for vec, deg in factors:
        _, phi, _ = asm.tab(deg, qdeg)
        val = val * fem.eval_at_quad(asm.space(deg), np.asarray(vec, dtype=float), phi)
    return float(coef) * float(np.einsum("q,c,cq->", w, np.abs(asm.detJ), val))


def l2_error(asm, u, uh, degree):
    """sqrt(sum_c int (u_c - uh_c)^2 dx) for component lists u, uh in the same Lagrange space of `degree`."""
    tot = 0.0
    for a, b in zip(u, uh):
        e = np.asarray(a, dtype=float) - np.asarray(b, dtype=float)
        tot += volume_integral(asm, [(e, degree), (e, degree)])
    return float(np.sqrt(max(tot, 0.0)))


def _facet_rule(dim, qdeg):
    """Points (barycentric on the facet) and weights (sum 1) of a rule exact to degree `qdeg` on a (dim-1)-simplex."""
    if dim == 2:
        x, w = np.polynomial.legendre.leggauss(qdeg // 2 + 1)
        t = 0.5 * (x + 1.0)
        return np.stack([1.0 - t, t], 1), 0.5 * w
    p, w = fem.simplex_quadrature(2, max(qdeg, 1))
    return np.concatenate([1.0 - p.sum(1, keepdims=True), p], 1), w / w.sum()


def boundary_facets_where(mesh, pred):
    """Boundary facets whose vertices and midpoint all satisfy pred(x): (facet vertex ids (F, d), owning cell, local facet index,
    outward unit normal, measure)."""
    d = mesh.dim
    cells = mesh.cells
    loc = [[j for j in range(d + 1) if j != k] for k in range(d + 1)]
    fv = np.concatenate([cells[:, l] for l in loc], axis=0)
    owner = np.tile(np.arange(mesh.nc), d + 1)
    lidx = np.repeat(np.arange(d + 1), mesh.nc)
    key = fv[:, 0].astype(np.int64)
    for j in range(1, d):
        key = key * mesh.nv + fv[:, j]
    _, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    sel = np.sort(idx[cnt == 1])
    fv, owner, lidx = fv[sel], owner[sel], lidx[sel]
    x = mesh.coords[fv]
    keep = np.array([all(bool(pred(p)) for p in f) and bool(pred(f.mean(0))) for f in x])
    fv, owner, lidx, x = fv[keep], owner[keep], lidx[keep], x[keep]
    opp = mesh.coords[cells[owner, lidx]]
    if d == 2:
        t = x[:, 1] - x[:, 0]
        n = np.stack([t[:, 1], -t[:, 0]], 1)
        meas = np.linalg.norm(t, axis=1)
    else:
        n = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
        meas = 0.5 * np.linalg.norm(n, axis=1)
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    n[((x[:, 0] - opp) * n).sum(1) < 0] *= -1.0
    return fv, owner, lidx, n, meas


def boundary_flux(asm, facets, factors, vel, direction=None, coef=1.0):
    """coef * int_{facets} prod_k f_k (u . g) ds with g the outward normal (direction=None) or a constant vector.
    vel = ("DG0", (C, d) array) | ("P1"|"P2"|"P3", [component dof vectors])."""
    mesh = asm.mesh
    d = mesh.dim
    fv, owner, lidx, n, meas = facets
    vdeg = 0 if vel[0] == "DG0" else int(vel[0][1])
    qdeg = max(sum(dg for _, dg in factors) + vdeg, 1)
    lamf, w = _facet_rule(d, qdeg)
    verts = np.concatenate([np.zeros((1, d)), np.eye(d)], 0)
    total = 0.0
    for k in range(d + 1):                      # facets grouped by their local index: one reference-point set per group
        m = lidx == k
        if not m.any():
            continue
        loc = [j for j in range(d + 1) if j != k]
        pts = lamf @ verts[loc]
        cellk = owner[m]
        g = n[m] if direction is None else np.broadcast_to(np.asarray(direction, dtype=float), (int(m.sum()), d))
        val = np.ones((int(m.sum()), len(w)))
        for vec, deg in factors:
            phi, _ = fem.tabulate(deg, pts)
            val = val * np.einsum("fb,qb->fq", np.asarray(vec, dtype=float)[asm.space(deg).cell_dofs[cellk]], phi)
        if vel[0] == "DG0":
            un = np.einsum("fa,fa->f", np.asarray(vel[1])[cellk], g)[:, None]
        else:
            phi, _ = fem.tabulate(vdeg, pts)
            sp = asm.space(vdeg)
            un = sum(np.einsum("fb,qb->fq", np.asarray(vel[1][a], dtype=float)[sp.cell_dofs[cellk]], phi) * g[:, a][:, None] for a in range(d))
        total += float(np.einsum("f,q,fq->", meas[m], w, val * un))
    return float(coef) * total


def cell_mean_speed_max(asm, vel, qdeg=4):
    """core.u_max (/root/reference/src/modules/core.py:23-39): max over cells of (1/|cell|) int_cell |u| dx with the degree-`qdeg` rule.
    vel = ("DG0", (C, d)) | ("P1"|"P2"|"P3", [component dof vectors])."""
    if vel[0] == "DG0":
        return float(np.sqrt((np.asarray(vel[1]) ** 2).sum(1)).max())
    deg = int(vel[0][1])
    w, phi, _ = asm.tab(deg, qdeg)
    sp = asm.space(deg)
    comps = [fem.eval_at_quad(sp, np.asarray(v, dtype=float), phi) for v in vel[1]]
    speed = np.sqrt(sum(c * c for c in comps))
    return float((speed @ (w / w.sum())).max())
