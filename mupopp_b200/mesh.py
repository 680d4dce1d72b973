"""Host-side meshes with DOLFIN 2017.1 conventions (vertex/cell numbering, UFC cell ordering).

Replaces `UnitSquareMesh/RectangleMesh/BoxMesh` as used by the reference run scripts
(e.g. /root/reference/src/run/LVL_3D/LVL_RII3D.py:136-142,
/root/reference/src/benchmark/time_marching/time_test1.py:45-58).  `mshr.generate_mesh`
(CGAL) is replaced by these structured generators for the BASELINE configs.
Mesh construction is one-time setup on the host; everything per timestep runs on the GPU.
"""
from __future__ import annotations

import numpy as np

TRI_EDGES = np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int64)                       # UFC: edge k opposite vertex k
TET_EDGES = np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]], dtype=np.int64)


class _Topology:
    def __init__(self, dim):
        self._dim = dim

    def dim(self):
        return self._dim


class Mesh:
    """Simplicial mesh: coords (V,d) float64, cells (C,d+1) int32 with ascending vertex ids per cell."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._edges = None
        self._cell_edges = None
        self._bfacets = None
        self.box = None

    # -- DOLFIN-like accessors -------------------------------------------------
    def coordinates(self):
        return self.coords

    def mpi_comm(self):
        return None                      # one process per GPU: there is no MPI communicator behind a mesh

    def topology(self):
        return _Topology(self.dim)

    def reset(self, coords, cells):
        """Replace the mesh data in place (HDF5File.read(mesh, "/mesh", False) fills an empty Mesh())."""
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._edges = self._cell_edges = self._bfacets = None
        self._faces = None
        self.box = None

    def num_vertices(self):
        return self.coords.shape[0]

    def num_cells(self):
        return self.cells.shape[0]

    def geometric_dimension(self):
        return self.coords.shape[1]

    @property
    def dim(self):
        return self.coords.shape[1]

    def ufl_cell(self):
        return "triangle" if self.dim == 2 else "tetrahedron"

    def cell_diameters(self):
        le = TRI_EDGES if self.dim == 2 else TET_EDGES
        x = self.coords[self.cells]
        d = x[:, le[:, 0]] - x[:, le[:, 1]]
        return np.sqrt((d * d).sum(-1)).max(1)

    def hmax(self):
        return float(self.cell_diameters().max())

    def hmin(self):
        return float(self.cell_diameters().min())

    # -- edges (needed by P2 spaces) --------------------------------------------
    def init_edges(self):
        """Edge ids in first-seen order over (cell, UFC local edge): the un-reordered DOLFIN numbering
        recorded in /root/reference/src/run/fenics_demo/dolfin_fine_velocity.xml.gz."""
        if self._edges is not None:
            return
        le = TRI_EDGES if self.dim == 2 else TET_EDGES
        nv = self.num_vertices()
        ev = self.cells[:, le].astype(np.int64)
        keys = (ev[..., 0] * nv + ev[..., 1]).ravel()
        uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty(order.size, dtype=np.int64)
        rank[order] = np.arange(order.size)
        self._cell_edges = rank[inv].reshape(ev.shape[:2]).astype(np.int32)
        u = uniq[order]
        self._edges = np.stack([u // nv, u % nv], 1).astype(np.int32)

    @property
    def edges(self):
        self.init_edges()
        return self._edges

    @property
    def cell_edges(self):
        self.init_edges()
        return self._cell_edges

    def num_edges(self):
        return self.edges.shape[0]

    # -- faces of tetrahedra (needed by P3 spaces) -----------------------------------
    def init_faces(self):
        """Face ids in first-seen order over (cell, UFC local facet k = the facet opposite local vertex k)."""
        if getattr(self, "_faces", None) is not None:
            return
        nv = self.num_vertices()
        loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
        fv = self.cells[:, loc].astype(np.int64)
        keys = ((fv[..., 0] * nv + fv[..., 1]) * nv + fv[..., 2]).ravel()
        uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty(order.size, dtype=np.int64)
        rank[order] = np.arange(order.size)
        self._cell_faces = rank[inv].reshape(fv.shape[:2]).astype(np.int32)
        u = uniq[order]
        self._faces = np.stack([u // (nv * nv), (u // nv) % nv, u % nv], 1).astype(np.int32)

    @property
    def faces(self):
        self.init_faces()
        return self._faces

    @property
    def cell_faces(self):
        self.init_faces()
        return self._cell_faces

    # -- boundary facets -----------------------------------------------------------
    def boundary_facets(self, exclude_axes=()):
        """(facet vertices (F,d), outward unit normals (F,d), measures (F,), owning cell (F,)).
        A boundary facet belongs to exactly one cell; order = ascending (local facet, cell).
        `exclude_axes`: faces normal to these axes are periodic images of one another, not boundary (result not cached)."""
        if exclude_axes:
            fv, n, meas, owner = self.boundary_facets()
            keep = np.ones(fv.shape[0], dtype=bool)
            for a in exclude_axes:
                keep &= np.abs(n[:, a]) < 0.5
            return fv[keep], n[keep], meas[keep], owner[keep]
        if self._bfacets is not None:
            return self._bfacets
        d = self.dim
        nv = self.num_vertices()
        fv, owner, opp, oloc = [], [], [], []
        if getattr(self, "box", None) is not None:
            # axis-aligned box domain: a facet is on the boundary iff all its vertices share a bounding plane
            lo, hi = self.box
            tol = 1e-12 * max(1.0, float(np.max(np.abs(hi - lo))))
            mask = np.zeros(nv, dtype=np.uint8)
            for a in range(d):
                mask |= (np.abs(self.coords[:, a] - lo[a]) <= tol).astype(np.uint8) << np.uint8(2 * a)
                mask |= (np.abs(self.coords[:, a] - hi[a]) <= tol).astype(np.uint8) << np.uint8(2 * a + 1)
            cm = mask[self.cells]
            for k in range(d + 1):
                loc = [j for j in range(d + 1) if j != k]
                m = cm[:, loc[0]]
                for j in loc[1:]:
                    m = m & cm[:, j]
                sel = np.nonzero(m)[0]
                fv.append(self.cells[sel][:, loc].astype(np.int64))
                owner.append(sel)
                opp.append(self.cells[sel, k].astype(np.int64))
                oloc.append(np.full(sel.size, k, dtype=np.int64))
            fv, owner, opp, oloc = np.concatenate(fv), np.concatenate(owner), np.concatenate(opp), np.concatenate(oloc)
        else:
            cells = self.cells.astype(np.int64)
            for k in range(d + 1):
                loc = [j for j in range(d + 1) if j != k]
                fv.append(cells[:, loc])
                owner.append(np.arange(cells.shape[0]))
                opp.append(cells[:, k])
                oloc.append(np.full(cells.shape[0], k, dtype=np.int64))
            fv = np.concatenate(fv)
            owner = np.concatenate(owner)
            opp = np.concatenate(opp)
            oloc = np.concatenate(oloc)
            key = fv[:, 0].copy()
            for j in range(1, d):
                key = key * nv + fv[:, j]
            _, idx, cnt = np.unique(key, return_index=True, return_counts=True)
            sel = np.sort(idx[cnt == 1])
            fv, owner, opp, oloc = fv[sel], owner[sel], opp[sel], oloc[sel]
        x = self.coords[fv]
        if d == 2:
            t = x[:, 1] - x[:, 0]
            n = np.stack([t[:, 1], -t[:, 0]], 1)
            meas = np.linalg.norm(t, axis=1)
        else:
            n = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
            meas = 0.5 * np.linalg.norm(n, axis=1)
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        flip = ((x[:, 0] - self.coords[opp]) * n).sum(1) < 0
        n[flip] *= -1.0
        self._bfacets = (fv, n, meas, owner)
        self._bfacet_local = oloc          # UFC local facet index (= local index of the opposite vertex) in the owning cell
        return self._bfacets

    def boundary_facet_local_index(self):
        self.boundary_facets()
        return self._bfacet_local


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    """DOLFIN RectangleMesh(Point, Point, nx, ny): vertex id iy*(nx+1)+ix, 'right' diagonals."""
    if diagonal != "right":
        raise NotImplementedError("only diagonal='right' (the DOLFIN default) is implemented")
    x0, y0 = float(p0[0]), float(p0[1])
    x1, y1 = float(p1[0]), float(p1[1])
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    coords = np.empty(((nx + 1) * (ny + 1), 2))
    coords[:, 0] = np.tile(xs, ny + 1)
    coords[:, 1] = np.repeat(ys, nx + 1)
    ix = np.tile(np.arange(nx, dtype=np.int64), ny)
    iy = np.repeat(np.arange(ny, dtype=np.int64), nx)
    v0 = iy * (nx + 1) + ix
    v1, v2 = v0 + 1, v0 + nx + 1
    v3 = v2 + 1
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.stack([v0, v1, v3], 1)
    cells[1::2] = np.stack([v0, v2, v3], 1)
    m = Mesh(coords, cells)
    m.box = (np.array([x0, y0]), np.array([x1, y1]))
    m.grid = (nx, ny)
    return m


def UnitSquareMesh(nx, ny, diagonal="right"):
    return RectangleMesh((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal)


def BoxMesh(p0, p1, nx, ny, nz, z_range=None):
    """DOLFIN BoxMesh(Point, Point, nx, ny, nz): vertex id iz*(ny+1)*(nx+1)+iy*(nx+1)+ix; each hex is cut
    into 6 tetrahedra around the body diagonal v0-v7.
    `z_range=(k0, k1)` builds only the hex layers k0..k1-1 (a z-slab for one GPU of a partitioned run);
    vertex ids are then local to the slab (plane k0 first), cells keep the global order."""
    k0, k1 = (0, nz) if z_range is None else z_range
    xs = np.linspace(float(p0[0]), float(p1[0]), nx + 1)
    ys = np.linspace(float(p0[1]), float(p1[1]), ny + 1)
    zs = np.linspace(float(p0[2]), float(p1[2]), nz + 1)[k0:k1 + 1]
    nzl = k1 - k0
    npl = (nx + 1) * (ny + 1)
    coords = np.empty((npl * (nzl + 1), 3))
    coords[:, 0] = np.tile(xs, (ny + 1) * (nzl + 1))
    coords[:, 1] = np.tile(np.repeat(ys, nx + 1), nzl + 1)
    coords[:, 2] = np.repeat(zs, npl)
    ix = np.tile(np.arange(nx, dtype=np.int64), ny * nzl)
    iy = np.tile(np.repeat(np.arange(ny, dtype=np.int64), nx), nzl)
    iz = np.repeat(np.arange(nzl, dtype=np.int64), nx * ny)
    v0 = iz * npl + iy * (nx + 1) + ix
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    v4, v5, v6, v7 = v0 + npl, v1 + npl, v2 + npl, v3 + npl
    cells = np.empty((6 * nx * ny * nzl, 4), dtype=np.int32)
    # rows are written already in ascending vertex order (UFC), cf. DOLFIN BoxMesh.cpp tetrahedra
    for k, t in enumerate([(v0, v1, v3, v7), (v0, v1, v5, v7), (v0, v4, v5, v7),
                           (v0, v2, v3, v7), (v0, v4, v6, v7), (v0, v2, v6, v7)]):
        cells[k::6, 0], cells[k::6, 1], cells[k::6, 2], cells[k::6, 3] = t
    m = Mesh(coords, cells)
    if z_range is None:
        m.box = (np.array([float(p0[0]), float(p0[1]), float(p0[2])]), np.array([float(p1[0]), float(p1[1]), float(p1[2])]))
    m.grid = (nx, ny, nz)
    m.z_range = (k0, k1)
    return m


def UnitCubeMesh(nx, ny, nz):
    return BoxMesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


class PeriodicMap:
    """P1 dof identification of a periodic `constrained_domain` (the reference: run/CCS/CCS_3D_top_source.py:134-158,
    run/CCS2D/CCS_3comp.py:131-143): slave vertices share the dof of their master; dofs are numbered compactly in master order.
    `vert2dof[v]` = dof of vertex v, `dof2vert[j]` = the master vertex of dof j, `axes` = periodic coordinate directions."""

    def __init__(self, vert2dof, axes, box):
        self.vert2dof = np.asarray(vert2dof, dtype=np.int32)
        self.ndof = int(self.vert2dof.max()) + 1          # need not be "master order" (slab partitions put owned planes first)
        self.axes = tuple(axes)
        self.box = box
        d2v = np.full(self.ndof, -1, dtype=np.int64)
        # masters are the lowest-numbered vertex of each class
        order = np.arange(self.vert2dof.size)[::-1]
        d2v[self.vert2dof[order]] = order
        self.dof2vert = d2v

    def fold_sum(self, a):
        return np.bincount(self.vert2dof, weights=np.asarray(a, dtype=float), minlength=self.ndof)

    def fold_any(self, a):
        out = np.zeros(self.ndof, dtype=np.uint8)
        np.maximum.at(out, self.vert2dof, np.asarray(a, dtype=np.uint8))
        return out

    def restrict(self, a):
        """Vertex array -> dof array taking the master's value."""
        return np.asarray(a)[self.dof2vert]


def box_periodic_map(mesh, axes):
    """Identify the max-face of each periodic axis with its min-face on an axis-aligned box mesh (any simplicial mesh whose
    boundary vertices on the two faces match up to a translation)."""
    if mesh.box is None:
        raise ValueError("box_periodic_map needs a mesh with a known bounding box")
    lo, hi = mesh.box
    X = mesh.coords
    tol = 1e-9 * float(np.max(hi - lo))
    img = X.copy()
    for a in axes:
        on_hi = np.abs(X[:, a] - hi[a]) <= tol
        img[on_hi, a] = lo[a]
    # one integer key per vertex from the per-axis ranks of the (rounded) image coordinates: three 1-D np.unique calls and one on the
    # combined key instead of np.unique(axis=0) on an (V, 3) array (8 M vertices: ~2 s instead of ~12 s of host setup)
    scale = 1.0 / (1e-7 * float(np.max(hi - lo)))
    key = np.zeros(X.shape[0], dtype=np.int64)
    for a in range(X.shape[1]):
        ka = np.round((img[:, a] - lo[a]) * scale).astype(np.int64)
        ua, ra = np.unique(ka, return_inverse=True)
        key = key * len(ua) + ra
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    master = first[inv.ravel()]                       # lowest vertex id with the same image
    is_master = master == np.arange(X.shape[0])
    dof_of_master = np.cumsum(is_master) - 1
    return PeriodicMap(dof_of_master[master], axes, (lo, hi))
