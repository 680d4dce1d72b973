"""Scalar functionals (`assemble(<scalar form>)`, `errornorm`) evaluated on the GPU by `mp_integrate` (csrc/functional.cu).

Call sites of the reference this serves:
    flux1 = assemble(c00*phi0*dot(u0, n)*ds(1))                     /root/reference/src/run/CCS2D/CCS_3comp.py:277
    assemble(c00*dot(u, unitNormal)*ds(1)), assemble(c02*dx), ...   /root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:232-265
    errornorm(u, uh, 'L2')                                          /root/reference/src/benchmark/sphere_3D_pure_shear/sphere.py:157-159
The host prepares, once per (mesh, measure): the entity lists (cells, or the boundary facets carrying a marker value) with their
measures / outward normals, and per integrand the basis tables at the points of an exact quadrature rule.  The product is evaluated
and reduced on the device; only the scalar result crosses the PCIe bus.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import reference_element as ref
from ._lib import MpFunctional, check


def _facet_points(dim, qdeg):
    """Quadrature on every local facet k (opposite vertex k) of the reference cell: points in CELL reference coordinates
    -> (dim+1, nq, dim) and weights (nq,) normalised to sum 1 (the facet measure is applied per entity)."""
    if dim == 2:
        x, w = np.polynomial.legendre.leggauss(qdeg // 2 + 1)
        t = 0.5 * (x + 1.0)
        w = 0.5 * w
        fl = np.stack([1.0 - t, t], 1)                       # barycentric coordinates on the facet
    else:
        p, w = ref.simplex_quadrature(2, max(qdeg, 1))
        w = w / w.sum()
        fl = ref.barycentric(p)
    nq = len(w)
    verts = np.concatenate([np.zeros((1, dim)), np.eye(dim)], 0)       # reference vertices 0..dim
    pts = np.empty((dim + 1, nq, dim))
    for k in range(dim + 1):
        loc = [j for j in range(dim + 1) if j != k]                    # ascending: the order mesh.boundary_facets uses
        pts[k] = fl @ verts[loc]
    return pts, w


class Integrator:
    """Entity lists and dofmaps of one mesh on the device; `integrate` evaluates one product functional."""

    def __init__(self, ctx, mesh):
        self.ctx, self.mesh = ctx, mesh
        self._ent = {}
        self._dofs = {}
        self._tabs = {}

    # ---- entities
    def cells(self):
        if "dx" not in self._ent:
            m = self.mesh
            x = m.coords[m.cells]
            J = x[:, 1:] - x[:, :1]
            vol = np.abs(np.linalg.det(J)) / (2.0 if m.dim == 2 else 6.0)
            self._ent["dx"] = (m.num_cells(), self.ctx.to_device(np.arange(m.num_cells(), dtype=np.int32), torch.int32), None,
                               self.ctx.to_device(vol), None)
        return self._ent["dx"]

    def facets(self, markers=None, value=None):
        """Boundary facets (all, or those where the facet MeshFunction `markers` equals `value`)."""
        key = ("ds", id(markers), value)
        if key not in self._ent:
            m = self.mesh
            if markers is None or value is None:
                fv, nrm, meas, owner = m.boundary_facets()
                loc = m.boundary_facet_local_index()
            else:
                owner, loc, nrm, meas = markers.facet_geometry(value)
            n = len(owner)
            dev = self.ctx.to_device
            self._ent[key] = (n, dev(np.asarray(owner, dtype=np.int32), torch.int32), dev(np.asarray(loc, dtype=np.int32), torch.int32),
                              dev(np.ascontiguousarray(meas, dtype=np.float64)) if n else self.ctx.zeros(1),
                              dev(np.ascontiguousarray(nrm, dtype=np.float64).ravel()) if n else self.ctx.zeros(1),
                              markers)          # keep the marker object alive: its id() is part of the key
        return self._ent[key][:5]

    # ---- dofmaps / tables
    def cell_dofs(self, degree, vert2dof=None):
        key = (degree, id(vert2dof) if vert2dof is not None else None)
        if key not in self._dofs:
            m = self.mesh
            nv = m.num_vertices()
            if degree == 1:
                cd = m.cells if vert2dof is None else np.asarray(vert2dof)[m.cells]
            elif degree == 2:
                cd = np.concatenate([m.cells, nv + m.cell_edges], axis=1)
            elif degree == 3:
                ce = m.cell_edges.astype(np.int64)
                ed = np.stack([nv + 2 * ce, nv + 2 * ce + 1], axis=2).reshape(m.num_cells(), -1)
                base = nv + 2 * m.num_edges()
                inner = base + (np.arange(m.num_cells(), dtype=np.int64)[:, None] if m.dim == 2 else m.cell_faces.astype(np.int64))
                cd = np.concatenate([m.cells.astype(np.int64), ed, inner], axis=1)
            else:
                raise NotImplementedError("Lagrange degree %d" % degree)
            self._dofs[key] = self.ctx.to_device(np.ascontiguousarray(cd, dtype=np.int32), torch.int32)
        return self._dofs[key]

    def tables(self, kind, qdeg, degree):
        """Basis values of the degree-`degree` Lagrange basis at the quadrature points of every reference entity:
        [ntab][nq][nb] with ntab = 1 (cells) or dim+1 (local facets), plus the weights (sum 1)."""
        key = (kind, qdeg, degree)
        if key not in self._tabs:
            d = self.mesh.dim
            if kind == "dx":
                pts, w = ref.simplex_quadrature(d, max(qdeg, 1))
                pts, w = pts[None], w / w.sum()
            else:
                pts, w = _facet_points(d, qdeg)
            tab = np.stack([ref.tabulate(degree, p)[0] for p in pts], 0)
            self._tabs[key] = (self.ctx.to_device(np.ascontiguousarray(tab).ravel()), self.ctx.to_device(np.ascontiguousarray(w)),
                               tab.shape[0], tab.shape[1], tab.shape[2])
        return self._tabs[key]

    # ---- the functional
    def integrate(self, kind, scalars, vector=None, direction=None, coef=1.0, markers=None, value=None):
        """coef * int prod_k f_k * (V . g) over dx or ds(value).
        scalars: list of (device dof vector, degree, vert2dof or None); vector: None | ("lagrange", [component vectors], degree) |
        ("cell", device [ncell*dim]); direction: None (facet normal; ds only) or a constant vector."""
        ctx, d = self.ctx, self.mesh.dim
        if len(scalars) > 4:
            raise NotImplementedError("at most 4 scalar factors per functional")
        qdeg = sum(s[1] for s in scalars) + (vector[2] if vector is not None and vector[0] == "lagrange" else 0)
        nent, ecell, etab, ew, eg = self.cells() if kind == "dx" else self.facets(markers, value)
        if vector is not None and direction is None and kind == "dx":
            raise ValueError("a vector factor under dx needs a constant direction")
        F = MpFunctional()
        F.nfield = len(scalars)
        keep = []
        nq = ntab = None
        wdev = None
        for k, (vec, deg, v2d) in enumerate(scalars):
            tab, wdev, ntab, nq, nb = self.tables(kind, qdeg, deg)
            dofs = self.cell_dofs(deg, v2d)
            F.nb[k], F.d_field[k], F.d_dofs[k], F.d_tab[k] = nb, vec.data_ptr(), dofs.data_ptr(), tab.data_ptr()
            keep += [tab, dofs]
        F.vec_kind = 0
        if vector is not None:
            if vector[0] == "lagrange":
                _, comps, vdeg = vector
                tab, wdev, ntab, nq, nbv = self.tables(kind, qdeg, vdeg)
                dofs = self.cell_dofs(vdeg)
                F.vec_kind, F.nbv, F.d_vdofs, F.d_vtab = 1, nbv, dofs.data_ptr(), tab.data_ptr()
                for a in range(d):
                    F.d_vec[a] = comps[a].data_ptr()
                keep += [tab, dofs]
            else:
                F.vec_kind = 2
                F.d_vec[0] = vector[1].data_ptr()
        if wdev is None:                                  # no Lagrange factor at all: a one-point rule
            tab, wdev, ntab, nq, nb = self.tables(kind, 1, 1)
        F.nq, F.ntab, F.d_w = nq, ntab, wdev.data_ptr()
        gc = None
        if direction is not None:
            gc = (C.c_double * 3)(*[float(direction[i]) if i < len(direction) else 0.0 for i in range(3)])
        res = C.c_double()
        use_normals = vector is not None and direction is None
        check(ctx.lib.mp_integrate(ctx.h, d, C.byref(F), int(nent), C.c_void_p(ecell.data_ptr()),
                                   C.c_void_p(etab.data_ptr()) if etab is not None else None, C.c_void_p(ew.data_ptr()),
                                   C.c_void_p(eg.data_ptr()) if (use_normals and eg is not None) else None, gc, C.byref(res)), "mp_integrate")
        return float(coef) * res.value
