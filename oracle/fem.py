"""Oracle FEM substrate (NumPy/SciPy, fp64): meshes, quadrature, Lagrange bases, dofmaps.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates the DOLFIN 2017.1
behaviours the reference scripts rely on:

* `UnitSquareMesh/RectangleMesh/BoxMesh` vertex and cell numbering
  (used by e.g. `/root/reference/src/run/LVL_3D/LVL_RII3D.py:136-142`,
  `/root/reference/src/benchmark/time_marching/time_test1.py:45-58`);
* UFC (un-reordered) dof numbering of Lagrange P1/P2/P3 spaces, pinned for
  P2-vector/triangle by `/root/reference/src/run/fenics_demo/dolfin_fine_velocity.xml.gz`;
* collapsed Gauss-Jacobi simplex quadrature (the explicit degree table replaces
  FFC's automatic degree estimation, SURVEY.md finding 7).
"""
from __future__ import annotations

import gzip
import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

# --------------------------------------------------------------------------- mesh

# UFC local edge -> local vertex pairs (edge k of a triangle is opposite vertex k)
TRI_EDGES = np.array([[1, 2], [0, 2], [0, 1]])
TET_EDGES = np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]])
# UFC local facet k is opposite local vertex k
TET_FACES = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])


@dataclass
class Mesh:
    coords: np.ndarray          # (V, d) float64
    cells: np.ndarray           # (C, d+1) int64, each row ascending (UFC ordering)
    edges: np.ndarray = field(default=None)       # (E, 2) ascending vertex pairs
    cell_edges: np.ndarray = field(default=None)  # (C, ne) global edge ids, UFC local order

    @property
    def dim(self):
        return self.coords.shape[1]

    @property
    def nv(self):
        return self.coords.shape[0]

    @property
    def nc(self):
        return self.cells.shape[0]

    def init_edges(self):
        """Global edge numbering = first-seen order iterating cells by index and
        local edges in UFC order (verified 8268/8268 against the fenics_demo fixture)."""
        if self.edges is not None:
            return
        le = TRI_EDGES if self.dim == 2 else TET_EDGES
        ev = self.cells[:, le]                       # (C, ne, 2), already ascending
        keys = ev[..., 0].astype(np.int64) * self.nv + ev[..., 1]
        flat = keys.ravel()
        uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")     # rank unique edges by first appearance
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        self.cell_edges = rank[inv].reshape(keys.shape)
        u = uniq[order]
        self.edges = np.stack([u // self.nv, u % self.nv], axis=1)

    @property
    def ne(self):
        self.init_edges()
        return self.edges.shape[0]

    def init_faces(self):
        """Global face numbering of a tetrahedral mesh: first-seen order over (cell, UFC local facet k = opposite vertex k)."""
        if getattr(self, "faces", None) is not None:
            return
        fv = self.cells[:, TET_FACES]                 # (C, 4, 3) ascending
        keys = (fv[..., 0].astype(np.int64) * self.nv + fv[..., 1]) * self.nv + fv[..., 2]
        uniq, first, inv = np.unique(keys.ravel(), return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        self.cell_faces = rank[inv].reshape(keys.shape)
        u = uniq[order]
        self.faces = np.stack([u // (self.nv * self.nv), (u // self.nv) % self.nv, u % self.nv], axis=1)

    def hmax_cells(self):
        """UFL CellDiameter: max vertex-vertex distance per cell (mupopp.py:797)."""
        le = TRI_EDGES if self.dim == 2 else TET_EDGES
        x = self.coords[self.cells]
        d = x[:, le[:, 0], :] - x[:, le[:, 1], :]
        return np.sqrt((d * d).sum(-1)).max(1)


def rectangle_mesh(x0, y0, x1, y1, nx, ny):
    """DOLFIN RectangleMesh, diagonal 'right' (SURVEY.md appendix B)."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")        # vertex id = iy*(nx+1)+ix
    coords = np.stack([X.ravel(), Y.ravel()], 1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + nx + 1
    v3 = v1 + nx + 1
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.stack([v0, v1, v3], 1)
    cells[1::2] = np.stack([v0, v2, v3], 1)
    return Mesh(coords, np.sort(cells, axis=1))


def unit_square_mesh(nx, ny):
    return rectangle_mesh(0.0, 0.0, 1.0, 1.0, nx, ny)


def box_mesh(p0, p1, nx, ny, nz):
    """DOLFIN BoxMesh: 6 tets per hex, all sharing the body diagonal v0-v7."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # id = iz*(ny+1)*(nx+1)+iy*(nx+1)+ix
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (iz * (ny + 1) * (nx + 1) + iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    v4 = v0 + (nx + 1) * (ny + 1)
    v5 = v1 + (nx + 1) * (ny + 1)
    v6 = v2 + (nx + 1) * (ny + 1)
    v7 = v3 + (nx + 1) * (ny + 1)
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    cells = np.empty((6 * nx * ny * nz, 4), dtype=np.int64)
    for k, t in enumerate(tets):
        cells[k::6] = np.stack(t, 1)
    return Mesh(coords, np.sort(cells, axis=1))


def unit_cube_mesh(n):
    return box_mesh((0, 0, 0), (1, 1, 1), n, n, n)


def read_dolfin_xml_mesh(path):
    """Minimal DOLFIN-XML mesh reader (for the fenics_demo fixture only)."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        txt = f.read()
    vs = re.findall(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"(?: z="([^"]+)")?', txt)
    has_z = any(v[3] for v in vs)
    coords = np.zeros((len(vs), 3 if has_z else 2))
    for i, x, y, z in vs:
        coords[int(i), 0] = float(x)
        coords[int(i), 1] = float(y)
        if has_z:
            coords[int(i), 2] = float(z)
    tr = re.findall(r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"', txt)
    cells = np.zeros((len(tr), 3), dtype=np.int64)
    for i, a, b, c in tr:
        cells[int(i)] = (int(a), int(b), int(c))
    return Mesh(coords, cells)


# --------------------------------------------------------------------------- quadrature

def simplex_quadrature(dim, degree):
    """Collapsed Gauss-Jacobi rule on the reference simplex, exact for total
    polynomial degree <= `degree`; n = degree//2 + 1 points per direction.
    Returns (pts (nq, dim), wts (nq,)), sum(wts) = 1/dim!."""
    n = degree // 2 + 1

    def gj(alpha):
        x, w = roots_jacobi(n, alpha, 0.0)
        return 0.5 * (x + 1.0), w * 0.5 ** (alpha + 1)

    if dim == 2:
        r, wr = gj(1.0)
        s, ws = gj(0.0)
        R, S = np.meshgrid(r, s, indexing="ij")
        W = wr[:, None] * ws[None, :]
        pts = np.stack([R.ravel(), (S * (1 - R)).ravel()], 1)
        return pts, W.ravel()
    if dim == 3:
        r, wr = gj(2.0)
        s, ws = gj(1.0)
        t, wt = gj(0.0)
        R, S, T = np.meshgrid(r, s, t, indexing="ij")
        W = wr[:, None, None] * ws[None, :, None] * wt[None, None, :]
        pts = np.stack([R.ravel(), (S * (1 - R)).ravel(), (T * (1 - R) * (1 - S)).ravel()], 1)
        return pts, W.ravel()
    raise ValueError(dim)


# --------------------------------------------------------------------------- bases

def _bary(pts):
    lam = np.concatenate([1.0 - pts.sum(1, keepdims=True), pts], axis=1)      # (nq, d+1)
    d = pts.shape[1]
    dlam = np.concatenate([-np.ones((1, d)), np.eye(d)], axis=0)              # (d+1, d)
    return lam, dlam


def tabulate(degree, pts):
    """Lagrange basis of `degree` (0: DG0, 1, 2) on the reference simplex at pts.
    UFC local dof order: vertices, then edges (UFC local edge order).
    Returns phi (nq, nb), dphi (nq, nb, d) (reference gradients)."""
    lam, dlam = _bary(pts)
    nq, d = pts.shape
    if degree == 0:
        return np.ones((nq, 1)), np.zeros((nq, 1, d))
    if degree == 1:
        return lam.copy(), np.broadcast_to(dlam, (nq, d + 1, d)).copy()
    if degree == 2:
        le = TRI_EDGES if d == 2 else TET_EDGES
        nb = d + 1 + len(le)
        phi = np.empty((nq, nb))
        dphi = np.empty((nq, nb, d))
        for i in range(d + 1):
            phi[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
            dphi[:, i, :] = (4 * lam[:, i] - 1)[:, None] * dlam[i][None, :]
        for k, (a, b) in enumerate(le):
            phi[:, d + 1 + k] = 4 * lam[:, a] * lam[:, b]
            dphi[:, d + 1 + k, :] = 4 * (lam[:, a][:, None] * dlam[b][None, :] + lam[:, b][:, None] * dlam[a][None, :])
        return phi, dphi
    if degree == 3:
        le = TRI_EDGES if d == 2 else TET_EDGES
        nint = 1 if d == 2 else 4
        nb = d + 1 + 2 * len(le) + nint
        phi = np.empty((nq, nb))
        dphi = np.empty((nq, nb, d))
        for i in range(d + 1):
            l = lam[:, i]
            phi[:, i] = 0.5 * l * (3 * l - 1) * (3 * l - 2)
            dphi[:, i, :] = (0.5 * (27 * l * l - 18 * l + 2))[:, None] * dlam[i][None, :]
        k = d + 1
        for (a, b) in le:
            la, lb = lam[:, a], lam[:, b]
            for (p, q, ip, iq) in ((la, lb, a, b), (lb, la, b, a)):    # dof nearer vertex a first, then nearer b
                phi[:, k] = 4.5 * p * q * (3 * p - 1)
                dphi[:, k, :] = 4.5 * ((6 * p * q - q)[:, None] * dlam[ip][None, :] + (3 * p * p - p)[:, None] * dlam[iq][None, :])
                k += 1
        tris = [(0, 1, 2)] if d == 2 else [tuple(f) for f in TET_FACES]
        for (a, b, c) in tris:
            la, lb, lc = lam[:, a], lam[:, b], lam[:, c]
            phi[:, k] = 27 * la * lb * lc
            dphi[:, k, :] = 27 * ((lb * lc)[:, None] * dlam[a][None, :] + (la * lc)[:, None] * dlam[b][None, :] + (la * lb)[:, None] * dlam[c][None, :])
            k += 1
        return phi, dphi
    raise NotImplementedError(degree)


# --------------------------------------------------------------------------- spaces

@dataclass
class ScalarSpace:
    """One scalar Lagrange space (degree 1 or 2) on a mesh, UFC numbering:
    vertex dofs = vertex id, edge dofs = V + edge id."""
    mesh: Mesh
    degree: int

    def __post_init__(self):
        m = self.mesh
        if self.degree == 1:
            self.ndof = m.nv
            self.cell_dofs = m.cells.copy()
        elif self.degree == 2:
            m.init_edges()
            self.ndof = m.nv + m.ne
            self.cell_dofs = np.concatenate([m.cells, m.nv + m.cell_edges], axis=1)
        elif self.degree == 3:
            # vertices | 2 dofs per edge (nearer the lower vertex first; cells are UFC-ordered so local == global direction)
            # | 1 dof per face (tets) or per cell (triangles)
            m.init_edges()
            ed = np.stack([m.nv + 2 * m.cell_edges, m.nv + 2 * m.cell_edges + 1], axis=2).reshape(m.nc, -1)
            base = m.nv + 2 * m.ne
            if m.dim == 2:
                inner = base + np.arange(m.nc)[:, None]
                self.ndof = base + m.nc
            else:
                m.init_faces()
                inner = base + m.cell_faces
                self.ndof = base + m.faces.shape[0]
            self.cell_dofs = np.concatenate([m.cells, ed, inner], axis=1)
        else:
            raise NotImplementedError

    def dof_coords(self):
        m = self.mesh
        if self.degree == 1:
            return m.coords.copy()
        if self.degree == 3:
            a, b = m.coords[m.edges[:, 0]], m.coords[m.edges[:, 1]]
            ed = np.stack([(2 * a + b) / 3.0, (a + 2 * b) / 3.0], axis=1).reshape(-1, m.dim)
            if m.dim == 2:
                inner = m.coords[m.cells].mean(1)
            else:
                inner = m.coords[m.faces].mean(1)
            return np.concatenate([m.coords, ed, inner], axis=0)
        mid = 0.5 * (m.coords[m.edges[:, 0]] + m.coords[m.edges[:, 1]])
        return np.concatenate([m.coords, mid], axis=0)


class MixedSpace:
    """UFC numbering of MixedElement([...]): sub-element blocks in order, a
    VectorElement being `dim` consecutive component blocks.
    `layout` is a list of (degree, ncomp), e.g. [(2, d), (1, 1), (1, 1), (1, 1)]
    for `/root/reference/src/run/CCS2D/CCS_3comp.py:206-212`."""

    def __init__(self, mesh, layout):
        self.mesh = mesh
        self.layout = layout
        self.spaces = {}
        self.blocks = []            # (offset, ScalarSpace) per scalar block
        self.sub_blocks = []        # per sub-element: list of block indices
        off = 0
        for degree, ncomp in layout:
            if degree not in self.spaces:
                self.spaces[degree] = ScalarSpace(mesh, degree)
            s = self.spaces[degree]
            idx = []
            for _ in range(ncomp):
                idx.append(len(self.blocks))
                self.blocks.append((off, s))
                off += s.ndof
            self.sub_blocks.append(idx)
        self.ndof = off

    def sub_dofs(self, isub, comp=None):
        idx = self.sub_blocks[isub]
        if comp is not None:
            idx = [idx[comp]]
        return np.concatenate([self.blocks[b][0] + np.arange(self.blocks[b][1].ndof) for b in idx])


# --------------------------------------------------------------------------- geometry / assembly helpers

def geometry(mesh):
    """Affine map per cell: J[c,i,a] = d x_i / d xi_a, detJ>0 in abs, Jinv."""
    x = mesh.coords[mesh.cells]                       # (C, d+1, d)
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    detJ = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    return J, np.abs(detJ), Jinv


def phys_grads(dphi_ref, Jinv):
    """dphi_ref (nq, nb, d) -> (C, nq, nb, d): d/dx_i = sum_a dphi/dxi_a * Jinv[a, i]."""
    return np.einsum("qba,cai->cqbi", dphi_ref, Jinv, optimize=True)


def scatter_matrix(Ae, rows, cols, shape):
    """Sum local matrices Ae (C, lr, lc) into CSR (duplicates summed, explicit zeros kept)."""
    C, lr, lc = Ae.shape
    r = np.repeat(rows, lc, axis=1).ravel()
    c = np.tile(cols, (1, lr)).ravel()
    A = sp.coo_matrix((Ae.ravel(), (r, c)), shape=shape).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def scatter_vector(be, rows, n):
    return np.bincount(rows.ravel(), weights=be.ravel(), minlength=n)


def eval_at_quad(space: ScalarSpace, vec, phi):
    """Field values at quadrature points: (C, nq)."""
    return np.einsum("cb,qb->cq", vec[space.cell_dofs], phi, optimize=True)


def grad_at_quad(space: ScalarSpace, vec, dphi_phys):
    """Field gradient at quadrature points: (C, nq, d)."""
    return np.einsum("cb,cqbi->cqi", vec[space.cell_dofs], dphi_phys, optimize=True)


def apply_dirichlet_symmetric(A, b, dofs, vals):
    """DOLFIN `assemble_system` semantics, applied globally: b -= A[:, D] g;
    zero rows/cols D; unit diagonal; b[D] = g  (SURVEY.md appendix B).  The
    reference's per-cell variant leaves (#cells sharing the dof) on the diagonal
    and the same multiple in b: same solution."""
    A = A.tocsr().copy()
    b = b.copy()
    n = A.shape[0]
    g = np.zeros(n)
    g[dofs] = vals
    b -= A @ g
    mask = np.ones(n)
    mask[dofs] = 0.0
    Dm = sp.diags(mask)
    A = Dm @ A @ Dm
    A = A + sp.diags(1.0 - mask)
    b[dofs] = vals
    return A.tocsr(), b


def p1_csr_pattern(mesh):
    """(rowptr, colidx) of the P1 vertex-adjacency pattern, columns ascending."""
    s = ScalarSpace(mesh, 1)
    C, l = s.cell_dofs.shape
    A = scatter_matrix(np.ones((C, l, l)), s.cell_dofs, s.cell_dofs, (s.ndof, s.ndof))
    return A.indptr.astype(np.int64), A.indices.astype(np.int64)
