"""Minimal dolfin-compatible facade: the subset of `from fenics import *` / `from mshr import *` that the in-scope MuPoPP
run scripts touch, backed by the B200 path.

"Forms" are opaque operator descriptors produced by `mupopp_b200.mupopp`; `solve(a == L, u, bcs)` dispatches them to the
GPU solvers of `mupopp_b200.moder` (the reference's own mixed P2/P1 discretisation, i.e. mode R).
Call sites mirrored: /root/reference/src/run/darcy_advection_diffusion/darcy_advection_diffusion.py:50-117,
/root/reference/src/run/CCS2D/CCS_3comp.py:206-263.  Not a general FEniCS replacement: anything else raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import mesh as _mesh
from .device import Context

from numpy import cos, exp, pi, sin, sqrt, tan, tanh  # noqa: F401,E402  (names the run scripts pick up from `from fenics import *`)

np = np
DOLFIN_EPS = 3.0e-16
parameters = {"std_out_all_processes": False, "form_compiler": {"quadrature_degree": None, "cpp_optimize": True}}
_CTX = None


def default_context():
    """The process-wide GPU context (device 0 or LOCAL_RANK)."""
    global _CTX
    if _CTX is None:
        import os
        _CTX = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _CTX


def info(msg, *a):
    print(msg)


def set_log_level(level):
    pass


def near(a, b, eps=DOLFIN_EPS):
    return abs(a - b) < eps


def has_linear_algebra_backend(name):
    return True


def has_krylov_solver_method(name):
    return name in ("cg", "gmres", "bicgstab", "minres", "tfqmr")


def has_krylov_solver_preconditioner(name):
    return name in ("jacobi", "amg", "none")


class Point:
    def __init__(self, *xs):
        self.x = tuple(float(v) for v in xs)

    def __getitem__(self, i):
        return self.x[i]

    def __len__(self):
        return len(self.x)


# ------------------------------------------------------------------ meshes (structured stand-ins for mshr)
def UnitSquareMesh(nx, ny, diagonal="right"):
    return _mesh.UnitSquareMesh(nx, ny, diagonal)


def UnitCubeMesh(nx, ny, nz):
    return _mesh.UnitCubeMesh(nx, ny, nz)


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    return _mesh.RectangleMesh(p0, p1, nx, ny, diagonal)


def BoxMesh(p0, p1, nx, ny, nz):
    return _mesh.BoxMesh(p0, p1, nx, ny, nz)


def Mesh(source=None, cells=None):
    """dolfin.Mesh: from a DOLFIN-XML file (optionally .gz; triangles or tetrahedra) as in
    /root/reference/src/run/fenics_demo/demo_advection-diffusion.py:33, from (coords, cells) arrays, or empty -- to be filled by
    `HDF5File(...).read(mesh, "/mesh", False)` (run/stokes_ADR_3D/master_advective_stokes.py:77-80)."""
    if source is None:
        return _mesh.Mesh(np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int32))
    if cells is not None:
        return _mesh.Mesh(source, cells)
    import gzip
    import re
    op = gzip.open if str(source).endswith(".gz") else open
    with op(source, "rt") as f:
        txt = f.read()
    vs = re.findall(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"(?: z="([^"]+)")?', txt)
    dim = 3 if 'celltype="tetrahedron"' in txt else 2
    coords = np.zeros((len(vs), dim))
    for i, x, y, z in vs:
        coords[int(i), 0], coords[int(i), 1] = float(x), float(y)
        if dim == 3:
            coords[int(i), 2] = float(z)
    if dim == 2:
        cs = re.findall(r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"', txt)
    else:
        cs = re.findall(r'<tetrahedron index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)" v3="(\d+)"', txt)
    cl = np.zeros((len(cs), dim + 1), dtype=np.int32)
    for row in cs:
        cl[int(row[0])] = [int(v) for v in row[1:]]
    return _mesh.Mesh(coords, np.sort(cl, axis=1))


class MeshFunction:
    """Facet markers: MeshFunction("size_t", mesh, file) for a DOLFIN-XML mesh_value_collection
    (run/fenics_demo/dolfin_fine_subdomains.xml.gz), or built from an (n,3) array of (cell, local facet, value)."""

    def __init__(self, kind, mesh, source=None, dim=None):
        self.mesh = mesh
        if isinstance(source, np.ndarray):
            self.entries = np.asarray(source, dtype=np.int64)
        elif source is not None and not isinstance(source, int):
            import gzip
            import re
            op = gzip.open if str(source).endswith(".gz") else open
            with op(source, "rt") as f:
                txt = f.read()
            rows = re.findall(r'<value cell_index="(\d+)" local_entity="(\d+)" value="(\d+)"', txt)
            self.entries = np.array([[int(a), int(b), int(c)] for a, b, c in rows], dtype=np.int64)
        else:
            self.entries = np.zeros((0, 3), dtype=np.int64)

    def facets_with(self, value):
        e = self.entries[self.entries[:, 2] == value]
        return e[:, 0], e[:, 1]

    def set_all(self, value):
        self.entries[:, 2] = int(value)

    @staticmethod
    def boundary(mesh, value=0):
        """FacetFunction("size_t", mesh): one entry per BOUNDARY facet (interior facets never carry a marker in the in-scope
        scripts: they are only used under ds)."""
        fv, nrm, meas, owner = mesh.boundary_facets()
        loc = mesh.boundary_facet_local_index()
        mf = MeshFunction("size_t", mesh)
        mf.entries = np.stack([owner, loc, np.full(len(owner), int(value))], 1).astype(np.int64)
        mf._bgeom = (fv, nrm, meas)
        return mf

    def facet_geometry(self, value):
        """(owning cell, local facet, outward unit normal, measure) of the facets marked `value`."""
        m = self.mesh
        d = m.dim
        sel = self.entries[:, 2] == value
        owner, loc = self.entries[sel, 0], self.entries[sel, 1]
        if owner.size == 0:
            return owner, loc, np.zeros((0, d)), np.zeros(0)
        cells = m.cells[owner].astype(np.int64)
        idx = np.array([[j for j in range(d + 1) if j != k] for k in range(d + 1)])[loc]          # (F, d) local vertices of the facet
        fv = np.take_along_axis(cells, idx, axis=1)
        x = m.coords[fv]
        opp = m.coords[cells[np.arange(len(owner)), loc]]
        if d == 2:
            t = x[:, 1] - x[:, 0]
            n = np.stack([t[:, 1], -t[:, 0]], 1)
            meas = np.linalg.norm(t, axis=1)
        else:
            n = np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0])
            meas = 0.5 * np.linalg.norm(n, axis=1)
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        n[((x[:, 0] - opp) * n).sum(1) < 0] *= -1.0
        return owner, loc, n, meas


class HDF5File:
    """dolfin.HDF5File(comm, filename, "r") with .read(mesh, "/mesh", use_partition) and .read(mesh_function, "/boundaries")
    (run/stokes_ADR_3D/master_advective_stokes.py:77-85), on the pure-Python reader of mupopp_b200/hdf5.py (h5py is not in the image)."""

    def __init__(self, comm, filename, mode="r"):
        if mode != "r":
            raise NotImplementedError("HDF5File: only reading is supported")
        self.filename = filename

    def read(self, obj, name, use_partition_from_file=False):
        from . import hdf5
        if isinstance(obj, MeshFunction):
            obj.entries = hdf5.read_facet_markers(self.filename, obj.mesh, name)
            return
        coords, cells = hdf5.read_mesh(self.filename, name)
        obj.reset(coords, cells)

    def close(self):
        pass


def FacetFunction(kind, mesh, value=0):
    return MeshFunction.boundary(mesh, value)


class SubDomain:
    """dolfin.SubDomain: override inside(x, on_boundary) (and map(x, y) for a periodic `constrained_domain`)."""

    def inside(self, x, on_boundary):
        return False

    def map(self, x, y):
        y[:] = x

    def mark(self, mesh_function, value):
        """Mark the boundary facets whose vertices AND midpoint are inside (DOLFIN's default check_midpoint=True)."""
        m = mesh_function.mesh
        d = m.dim
        e = mesh_function.entries
        if e.shape[0] == 0:
            return
        cells = m.cells[e[:, 0]].astype(np.int64)
        idx = np.array([[j for j in range(d + 1) if j != k] for k in range(d + 1)])[e[:, 1]]
        fv = np.take_along_axis(cells, idx, axis=1)
        verts = np.unique(fv)
        vin = np.zeros(m.num_vertices(), dtype=bool)
        vin[verts] = [bool(self.inside(m.coords[v], True)) for v in verts]
        ok = vin[fv].all(1)
        for f in np.nonzero(ok)[0]:
            ok[f] = bool(self.inside(m.coords[fv[f]].mean(0), True))
        e[ok, 2] = int(value)


def read_function_xml(path, f):
    """`File(path) >> f` for DOLFIN-XML function data (rows `dof index, value`): UFC block numbering is assumed, which is what the
    reference's fixture uses (tests/test_oracle_golden.py pins it)."""
    import gzip
    import re
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as fh:
        txt = fh.read()
    rows = re.findall(r'<dof index="(\d+)" value="([^"]+)"', txt)
    vals = np.zeros(len(rows))
    for i, v in rows:
        vals[int(i)] = float(v)
    f.vector().set_local(vals)


class Rectangle:
    def __init__(self, p0, p1):
        self.p0, self.p1 = p0, p1


class Box:
    def __init__(self, p0, p1):
        self.p0, self.p1 = p0, p1


class Sphere:
    """mshr.Sphere(center, radius) stand-in: without CGAL the bounding cube is meshed (run/sphere_archer/sphere.py:91-92)."""

    def __init__(self, center, radius):
        c = [float(center[i]) for i in range(3)]
        r = float(radius)
        self.center, self.radius = c, r
        self.p0, self.p1 = [ci - r for ci in c], [ci + r for ci in c]


def generate_mesh(domain, resolution):
    """mshr.generate_mesh stand-in: a structured mesh with cell size ~ (domain diagonal)/resolution (CGAL is not available)."""
    p0, p1 = domain.p0, domain.p1
    ext = [abs(p1[i] - p0[i]) for i in range(len(p0))]
    h = math.sqrt(sum(e * e for e in ext)) / float(resolution)
    n = [max(1, int(round(e / h))) for e in ext]
    if len(ext) == 2:
        return _mesh.RectangleMesh(p0, p1, n[0], n[1])
    return _mesh.BoxMesh(p0, p1, n[0], n[1], n[2])


def periodic_map_from_subdomain(mesh, pbc):
    """mesh.PeriodicMap of a DOLFIN periodic `constrained_domain`: a boundary vertex x that is not `inside` is a SLAVE when
    `pbc.map(x, y)` lands on a mesh vertex y with `pbc.inside(y, True)`; it then shares that master's dof.  Dofs are numbered
    compactly in master-vertex order (the same rule as mesh.box_periodic_map and the oracle's periodic_prolongation)."""
    if isinstance(pbc, _mesh.PeriodicMap):
        return pbc
    X = mesh.coords
    d = mesh.dim
    fv = mesh.boundary_facets()[0]
    bverts = np.unique(fv.ravel())
    ext = float(np.max(X.max(0) - X.min(0)))
    lo = X.min(0)
    scale = 1.0 / (1e-7 * ext)
    keys_of = lambda P: list(map(tuple, np.round((np.asarray(P) - lo) * scale).astype(np.int64).tolist()))     # vectorised: one NumPy call
    bverts = [int(v) for v in bverts]
    Xb = X[bverts]
    is_master = [bool(pbc.inside(x, True)) for x in Xb]                     # the user's Python callbacks are the unavoidable per-point cost
    kb = keys_of(Xb)
    where = {k: v for k, v, im in zip(kb, bverts, is_master) if im}
    master = np.arange(X.shape[0])
    slaves = [i for i, im in enumerate(is_master) if not im]
    Y = Xb[slaves].copy()
    for row, i in enumerate(slaves):
        pbc.map(Xb[i], Y[row])
    axes = set()
    for k, i in zip(keys_of(Y) if slaves else [], slaves):
        v = bverts[i]
        m = where.get(k)
        if m is not None and m != v:
            master[v] = m
            axes.update(int(a) for a in np.nonzero(np.abs(X[v] - X[m]) > 1e-9 * ext)[0])
    while True:                                     # path compression (corner vertices of multiply periodic boxes)
        nxt = master[master]
        if np.array_equal(nxt, master):
            break
        master = nxt
    ism = master == np.arange(X.shape[0])
    dof_of_master = np.cumsum(ism) - 1
    return _mesh.PeriodicMap(dof_of_master[master], tuple(sorted(axes)), (X.min(0), X.max(0)))


class _UfcCell:
    def __init__(self, index, local_facet):
        self.index, self.local_facet = index, local_facet


class Cell:
    """dolfin.Cell(mesh, index): only .normal(local_facet) (used by BoundarySource, darcy_advection_diffusion.py:29-34)."""

    def __init__(self, mesh, index):
        self.mesh, self.index = mesh, int(index)

    def normal(self, local_facet):
        m = self.mesh
        d = m.dim
        vs = m.cells[self.index]
        loc = [j for j in range(d + 1) if j != local_facet]
        x = m.coords[vs[loc]]
        if d == 2:
            t = x[1] - x[0]
            n = np.array([t[1], -t[0]])
        else:
            n = np.cross(x[1] - x[0], x[2] - x[0])
        n = n / np.linalg.norm(n)
        if np.dot(x[0] - m.coords[vs[local_facet]], n) < 0:
            n = -n
        return n


# ------------------------------------------------------------------ elements / spaces
class FiniteElement:
    def __init__(self, family, cell=None, degree=1):
        if family not in ("Lagrange", "CG", "P"):
            raise NotImplementedError("element family %r" % (family,))
        self.degree, self.ncomp = int(degree), 1


class VectorElement(FiniteElement):
    def __init__(self, family, cell=None, degree=1, dim=None):
        super().__init__(family, cell, degree)
        self.ncomp = dim if dim is not None else (2 if cell in ("triangle", None) else 3)
        self._cell = cell


class MixedElement:
    def __init__(self, elements):
        self.elements = list(elements)


class FunctionSpace:
    """Scalar, vector or mixed Lagrange space with UFC block numbering; blocks = [(degree, offset, ndof)]."""

    def __init__(self, mesh, element, degree=None, constrained_domain=None, _parent=None, _sub=None):
        # periodic `constrained_domain` (run/CCS2D/CCS_3comp.py:131-143,212; run/CCS/CCS_3D_top_source.py:134-158,211): slave vertices
        # share the dof of their master.  Such spaces are solved by mode F (P1 fields, cellwise velocity): P1 blocks carry one dof per
        # vertex CLASS; the velocity sub-space is STORED at the P1 dofs as well (the vertex average of the cellwise recovery, what the
        # scripts write to file), the cellwise velocity itself travels with the Function (Function.cell_velocity).
        self.periodic = periodic_map_from_subdomain(mesh, constrained_domain) if constrained_domain is not None else None
        self.mesh = mesh
        if isinstance(element, str):
            element = FiniteElement(element, mesh.ufl_cell(), degree)
        if isinstance(element, VectorElement):
            element.ncomp = mesh.dim
        if isinstance(element, MixedElement):
            for e in element.elements:
                if isinstance(e, VectorElement):
                    e.ncomp = mesh.dim
        self.element = element
        self.parent, self.sub_index = _parent, _sub
        subs = element.elements if isinstance(element, MixedElement) else [element]
        self.sub_elements = subs
        self.blocks = []         # (degree, offset, ndof) per scalar block
        self.sub_blocks = []     # block indices per sub-element
        off = 0
        for e in subs:
            idx = []
            if e.degree not in (1, 2, 3):
                raise NotImplementedError("Lagrange degree %d" % e.degree)
            nd = mesh.num_vertices()
            deg = e.degree
            if self.periodic is not None:
                nd, deg = self.periodic.ndof, 1
            elif e.degree == 2:
                nd += mesh.num_edges()
            elif e.degree == 3:
                nd += 2 * mesh.num_edges() + (mesh.num_cells() if mesh.dim == 2 else mesh.faces.shape[0])
            for _ in range(e.ncomp):
                idx.append(len(self.blocks))
                self.blocks.append((deg, off, nd))
                off += nd
            self.sub_blocks.append(idx)
        self._dim = off

    def dim(self):
        return self._dim

    def sub(self, i):
        if self.parent is not None:          # W.sub(0).sub(c): one component of a vector sub-space
            return _SubSpace(self.parent, self.sub_index, i)
        return _SubSpace(self, i, None)

    def num_sub_spaces(self):
        return len(self.sub_elements)

    def dof_coords(self, degree):
        m = self.mesh
        if self.periodic is not None:
            return m.coords[self.periodic.dof2vert]
        if degree == 1:
            return m.coords
        e = m.edges
        a, b = m.coords[e[:, 0]], m.coords[e[:, 1]]
        if degree == 2:
            return np.concatenate([m.coords, 0.5 * (a + b)], axis=0)
        ed = np.stack([(2.0 * a + b) / 3.0, (a + 2.0 * b) / 3.0], axis=1).reshape(-1, m.dim)
        inner = m.coords[m.cells].mean(1) if m.dim == 2 else m.coords[m.faces].mean(1)
        return np.concatenate([m.coords, ed, inner], axis=0)


class _SubSpace:
    def __init__(self, parent, isub, comp):
        self.parent, self.isub, self.comp = parent, isub, comp
        self.mesh = parent.mesh

    def sub(self, c):
        return _SubSpace(self.parent, self.isub, c)

    def block_ids(self):
        ids = self.parent.sub_blocks[self.isub]
        return ids if self.comp is None else [ids[self.comp]]


def VectorFunctionSpace(mesh, family, degree, dim=None):
    return FunctionSpace(mesh, VectorElement(family, mesh.ufl_cell(), degree, dim or mesh.dim))


# ------------------------------------------------------------------ coefficients
class Constant:
    def __init__(self, value):
        self.value = np.atleast_1d(np.asarray(value, dtype=float))

    def assign(self, v):
        self.value = np.atleast_1d(np.asarray(getattr(v, "value", v), dtype=float))

    def value_shape(self):
        return () if self.value.size == 1 else (self.value.size,)

    def values(self):
        return self.value

    def __float__(self):
        return float(self.value[0])

    def __getitem__(self, i):
        return float(self.value[i])

    def __len__(self):
        return self.value.size


class Expression:
    """String expressions (`Expression("0.1", degree=1)`, `Expression("dt", dt=0.1)`) or user subclasses overriding
    eval(values, x) / eval_cell(values, x, ufc_cell) / value_shape()."""

    def __init__(self, code=None, degree=1, element=None, **kw):
        self._code = code
        self.degree = degree
        self.element = element
        self.user = dict(kw)
        for k, v in kw.items():
            setattr(self, k, v)

    def value_shape(self):
        if isinstance(self._code, (tuple, list)):
            return (len(self._code),)
        return ()

    def eval(self, values, x):
        codes = self._code if isinstance(self._code, (tuple, list)) else [self._code]
        ns = {k: getattr(math, k) for k in ("sin", "cos", "tan", "exp", "sqrt", "tanh", "log", "pow", "fabs")}
        ns.update(pi=math.pi, x=x, DOLFIN_EPS=DOLFIN_EPS)
        for k in self.user:
            ns[k] = getattr(self, k)
        for i, c in enumerate(codes):
            values[i] = eval(str(c).replace("&&", " and ").replace("||", " or "), {"__builtins__": {}}, ns)

    def _nvals(self):
        vs = self.value_shape()
        return int(np.prod(vs)) if len(vs) else 1


def _evaluate(value, pts, ncomp, cell_info=None):
    """Evaluate a Constant / Expression / Function-like / number at points -> (N, ncomp)."""
    n = pts.shape[0]
    if isinstance(value, (int, float)):
        return np.full((n, ncomp), float(value))
    if isinstance(value, Constant):
        return np.broadcast_to(value.value.reshape(1, -1), (n, ncomp)).copy()
    if isinstance(value, Function):
        if len(value.blocks) != ncomp or any(b.numel() != n for b in value.blocks):
            raise NotImplementedError("evaluating a Function at points other than its own dofs")
        return np.stack([b.cpu().numpy() for b in value.blocks], 1)
    if isinstance(value, Expression):
        use_cell = hasattr(value, "eval_cell") and cell_info is not None
        if not use_cell and n > 64:
            # vectorised attempt: x[k] are coordinate ARRAYS.  Works for expressions written with numpy-style arithmetic (the names
            # sin/cos/tanh/exp/sqrt exported by this module are numpy's); anything that branches on x falls back to the point loop.
            try:
                vals = [None] * max(ncomp, 1)
                value.eval(vals, [pts[:, k] for k in range(pts.shape[1])])
                cols = [np.broadcast_to(np.asarray(v, dtype=float), (n,)) for v in vals[:ncomp]]
                return np.stack(cols, 1).copy()
            except Exception:
                pass
        out = np.zeros((n, ncomp))
        buf = np.zeros(max(ncomp, 1))
        for i in range(n):
            buf[:] = 0.0
            if use_cell:
                value.eval_cell(buf, pts[i], _UfcCell(int(cell_info[0][i]), int(cell_info[1][i])))
            else:
                value.eval(buf, pts[i])
            out[i] = buf[:ncomp]
        return out
    raise TypeError("cannot evaluate %r" % (value,))


class Function:
    """Coefficient vector(s) of a FunctionSpace, stored per scalar block on the GPU."""

    def __init__(self, V, _blocks=None, _ids=None):
        self.space = V
        ctx = default_context()
        if _blocks is None:
            self.blocks = [ctx.zeros(nd) for (_, _, nd) in V.blocks]
            self.ids = list(range(len(V.blocks)))
        else:
            self.blocks, self.ids = _blocks, _ids
        self._name = "f"
        self.cell_velocity = None       # mode F: the cellwise (DG0) velocity [ncell*dim] behind the stored vertex averages

    def function_space(self):
        return self.space

    def rename(self, name, label=""):
        self._name = name

    def split(self, deepcopy=False):
        out = []
        for idx in self.space.sub_blocks:
            blks = [self.blocks[self.ids.index(i)] for i in idx]
            if deepcopy:
                blks = [b.clone() for b in blks]
            f = Function(self.space, blks, list(idx))
            if self.cell_velocity is not None and list(idx) == list(self.space.sub_blocks[0]) and len(idx) == self.space.mesh.dim:
                f.cell_velocity = self.cell_velocity.clone() if deepcopy else self.cell_velocity
            out.append(f)
        return tuple(out)

    def sub(self, i):
        return self.split()[i]

    def block_degree(self, k):
        return self.space.blocks[self.ids[k]][0]

    def vector(self):
        return _VectorView(self)

    def assign(self, other):
        for a, b in zip(self.blocks, other.blocks):
            a.copy_(b)
        if getattr(other, "cell_velocity", None) is not None and len(self.blocks) == len(other.blocks):
            self.cell_velocity = other.cell_velocity.clone()

    # ---- integrand algebra for scalar functionals: c*phi0*dot(u, n)*ds(1), c*(1-phi0)*dx, ...  (see assemble())
    def __mul__(self, other):
        return _Integrand.of(self) * other

    __rmul__ = __mul__

    def __neg__(self):
        return _Integrand.of(self) * -1.0

    def interpolate(self, v):
        ctx = default_context()
        if isinstance(v, Function):
            if len(v.blocks) != len(self.blocks):
                raise ValueError("interpolate: block structure differs")
            for a, b in zip(self.blocks, v.blocks):
                if a.numel() != b.numel():
                    raise NotImplementedError("interpolation between different degrees")
                a.copy_(b)
            return
        ncomp = len(self.blocks)
        deg = self.block_degree(0)
        X = self.space.dof_coords(deg)
        vals = _evaluate(v, X, ncomp)
        for c in range(ncomp):
            self.blocks[c].copy_(ctx.to_device(np.ascontiguousarray(vals[:, c])))

    def get(self, k=0):
        """Host copy of scalar block k (test / output helper)."""
        return self.blocks[k].cpu().numpy()


class _VectorView:
    def __init__(self, f):
        self.f = f

    def get_local(self):
        return np.concatenate([b.cpu().numpy() for b in self.f.blocks])

    def array(self):
        return self.get_local()

    def set_local(self, a):
        ctx = default_context()
        off = 0
        for b in self.f.blocks:
            b.copy_(ctx.to_device(np.asarray(a[off:off + b.numel()], dtype=np.float64)))
            off += b.numel()

    def __setitem__(self, key, a):
        self.set_local(np.broadcast_to(a, (sum(b.numel() for b in self.f.blocks),)) if np.isscalar(a) else a)


def assign(dst, src):
    dst.assign(src)


def interpolate(v, V):
    f = Function(V)
    f.interpolate(v)
    return f


# ------------------------------------------------------------------ boundary conditions
class DirichletBC:
    """DirichletBC(V | V.sub(i)[.sub(c)], value, marker): 'topological' marking -- a boundary facet is constrained when the
    marker holds at all its vertices and its midpoint; every dof on it (vertices, and edge midpoints for P2) is set."""

    def __init__(self, V, value, marker, sub_id=None, method="topological"):
        """marker: callable x[, on_boundary] -> bool, or a MeshFunction of facet markers together with `sub_id`
        (DirichletBC(Q, g, sub_domains, 1), demo_advection-diffusion.py:79; master_advective_stokes.py:119-130)."""
        self.V, self.value, self.marker, self.sub_id = V, value, marker, sub_id
        if isinstance(V, _SubSpace):
            self.space, self.block_ids = V.parent, V.block_ids()
        else:
            self.space, self.block_ids = V, list(range(len(V.blocks)))

    def _mark(self, x):
        if isinstance(self.marker, SubDomain):
            return bool(self.marker.inside(x, True))
        try:
            return bool(self.marker(x, True))
        except TypeError:
            return bool(self.marker(x))

    def facet_mask(self, fv, owner, oloc):
        """Which of the given boundary facets (vertex ids fv, owning cell, local index) carry this condition: marker function /
        SubDomain -> all vertices and the midpoint inside ('topological' marking); MeshFunction + id -> the facets with that value."""
        mesh = self.space.mesh
        if isinstance(self.marker, MeshFunction):
            oc, ol = self.marker.facets_with(self.sub_id)
            have = set(zip(oc.tolist(), ol.tolist()))
            return np.array([(int(c), int(k)) in have for c, k in zip(owner, oloc)], dtype=bool)
        vx = mesh.coords
        bverts = np.unique(fv.ravel())
        vin = np.zeros(mesh.num_vertices(), dtype=bool)
        vin[bverts] = [self._mark(vx[i]) for i in bverts]
        inside = vin[fv].all(1)
        cand = np.nonzero(inside)[0]
        mid = vx[fv[cand]].mean(1)                    # facet midpoints, one NumPy call
        for f, xm in zip(cand, mid):
            inside[f] = self._mark(xm)
        return inside

    def dofs_and_values(self):
        """-> {block id: (dof indices, values)} in the block's own numbering."""
        mesh = self.space.mesh
        d = mesh.dim
        vx = mesh.coords
        if isinstance(self.marker, MeshFunction):
            owner, oloc = self.marker.facets_with(self.sub_id)
            fv = np.array([[mesh.cells[c][j] for j in range(d + 1) if j != k] for c, k in zip(owner, oloc)], dtype=np.int64).reshape(-1, d)
            marked = np.arange(len(owner))
        else:
            fv, nrm, meas, owner = mesh.boundary_facets()
            oloc = mesh.boundary_facet_local_index()
            bverts = np.unique(fv.ravel())
            vin = np.zeros(mesh.num_vertices(), dtype=bool)
            vin[bverts] = [self._mark(vx[i]) for i in bverts]
            inside = vin[fv].all(1)
            cand = np.nonzero(inside)[0]
            mid = vx[fv[cand]].mean(1)                # facet midpoints, one NumPy call
            for f, xm in zip(cand, mid):
                inside[f] = self._mark(xm)
            marked = np.nonzero(inside)[0]
        degree = self.space.blocks[self.block_ids[0]][0]
        ncomp = len(self.block_ids)
        # collect (dof, point, owning cell, local facet) over marked facets, first occurrence wins
        seen = {}
        from .reference_element import TET_EDGES, TRI_EDGES
        le = TRI_EDGES if d == 2 else TET_EDGES
        nv = mesh.num_vertices()
        for f in marked:
            c, k = int(owner[f]), int(oloc[f])
            for v in fv[f]:
                seen.setdefault(int(v), (vx[v], c, k))
            if degree >= 2:
                ce = mesh.cell_edges[c]
                for e, (a, b) in enumerate(le):
                    if a != k and b != k:            # edge lies on the facet opposite local vertex k
                        va, vb = mesh.cells[c][a], mesh.cells[c][b]
                        if degree == 2:
                            seen.setdefault(nv + int(ce[e]), (0.5 * (vx[va] + vx[vb]), c, k))
                        else:                        # P3: two dofs per edge, the one nearer the lower vertex first
                            lo, hi = (va, vb) if va < vb else (vb, va)
                            seen.setdefault(nv + 2 * int(ce[e]), ((2.0 * vx[lo] + vx[hi]) / 3.0, c, k))
                            seen.setdefault(nv + 2 * int(ce[e]) + 1, ((vx[lo] + 2.0 * vx[hi]) / 3.0, c, k))
                if degree == 3 and d == 3:           # the face dof of the boundary facet itself
                    gid = nv + 2 * mesh.num_edges() + int(mesh.cell_faces[c][k])
                    seen.setdefault(gid, (vx[fv[f]].mean(0), c, k))
        if not seen:
            return {b: (np.zeros(0, dtype=np.int64), np.zeros(0)) for b in self.block_ids}
        dofs = np.array(sorted(seen), dtype=np.int64)
        pts = np.array([seen[i][0] for i in dofs])
        cinfo = (np.array([seen[i][1] for i in dofs]), np.array([seen[i][2] for i in dofs]))
        vals = _evaluate(self.value, pts, ncomp, cinfo)
        return {b: (dofs, vals[:, j].copy()) for j, b in enumerate(self.block_ids)}


# ------------------------------------------------------------------ forms, solve, output
class Form:
    """Opaque operator descriptor returned by the mupopp form methods."""

    def __init__(self, kind, side, **data):
        self.kind, self.side, self.data = kind, side, data

    def __eq__(self, other):
        if isinstance(other, Form):
            return Equation(self, other)
        return NotImplemented

    __hash__ = object.__hash__


class Equation:
    def __init__(self, a, L):
        if a.data is not L.data and a.kind != L.kind:
            raise ValueError("a and L come from different forms")
        self.a, self.L = a, L


def solve(eq, u, bcs=None, solver_parameters=None, **kw):
    """dolfin.solve(a == L, u, bcs): dispatch the form descriptor to its GPU solver (mode R)."""
    if not isinstance(eq, Equation):
        raise TypeError("solve expects `a == L` built from mupopp_b200.mupopp forms")
    from . import mupopp as _mp
    bcs = [] if bcs is None else (list(bcs) if isinstance(bcs, (list, tuple)) else [bcs])
    return _mp._dispatch_solve(eq.a, u, bcs, solver_parameters or {})


class File:
    """`File(name, "compressed") << f` writes a VTK time series (name.pvd + name000000.vtu ...), `File(name.xml[.gz]) >> f` reads DOLFIN-XML
    function data.  Set MUPOPP_B200_WRITE=0 to turn the writer into a no-op (benchmark runs)."""

    def __init__(self, name, *a):
        self.name, self.count, self._series = name, 0, None

    def __rshift__(self, f):
        read_function_xml(self.name, f)
        return self

    def __lshift__(self, f):
        import os
        if os.environ.get("MUPOPP_B200_WRITE", "1") != "0":
            from .io import PvdSeries
            t = None
            if isinstance(f, tuple):
                f, t = f
            if self._series is None:
                self._series = PvdSeries(self.name)
            mesh = f.space.mesh
            nv = mesh.num_vertices()
            cols = [f.blocks[k].cpu().numpy()[:nv] for k in range(len(f.blocks))]       # Lagrange fields at the vertices
            data = cols[0] if len(cols) == 1 else np.stack(cols, axis=1)
            self._series.write(mesh.coords, mesh.cells, {f._name: data}, t)
        self.count += 1
        return self


# ------------------------------------------------------------------ scalar functionals: assemble(c*phi0*dot(u, n)*ds(1)), errornorm, project
class FacetNormal:
    """Outward unit normal of the boundary facets; only meaningful inside dot(u, n) under a ds measure."""

    def __init__(self, mesh, sign=1.0):
        self.mesh, self.sign = mesh, sign

    def __neg__(self):
        return FacetNormal(self.mesh, -self.sign)


class Measure:
    """Measure("dx") / Measure("ds", subdomain_data=facets) / the 2017-style Measure("ds")[facets]; ds(1) selects a marker value."""

    def __init__(self, kind, domain=None, subdomain_data=None, subdomain_id=None):
        if kind not in ("dx", "ds"):
            raise NotImplementedError("measure %r" % (kind,))
        self.kind, self.data, self.sid = kind, subdomain_data, subdomain_id

    def __getitem__(self, data):
        return Measure(self.kind, subdomain_data=data, subdomain_id=self.sid)

    def __call__(self, subdomain_id=None, subdomain_data=None, domain=None):
        return Measure(self.kind, subdomain_data=subdomain_data if subdomain_data is not None else self.data, subdomain_id=subdomain_id)

    def __rmul__(self, integrand):
        return _Integral(_Integrand.of(integrand), self)


dx = Measure("dx")
ds = Measure("ds")


class _Integrand:
    """coef * prod(scalar Lagrange Functions) * (vector Function . direction): what the in-scope functionals are made of."""

    def __init__(self, coef=1.0, scalars=(), vector=None, direction=None):
        self.coef, self.scalars, self.vector, self.direction = float(coef), tuple(scalars), vector, direction

    @staticmethod
    def of(x):
        if isinstance(x, _Integrand):
            return x
        if isinstance(x, (int, float)):
            return _Integrand(float(x))
        if isinstance(x, Constant) and x.value.size == 1:
            return _Integrand(float(x))
        if isinstance(x, Function):
            if len(x.blocks) != 1:
                raise NotImplementedError("a vector Function can enter a functional only through dot(u, n) / dot(u, Constant)")
            return _Integrand(1.0, (x,))
        raise NotImplementedError("integrand factor %r" % (x,))

    def __mul__(self, other):
        if isinstance(other, Measure):
            return _Integral(self, other)
        o = _Integrand.of(other)
        if self.vector is not None and o.vector is not None:
            raise NotImplementedError("two vector factors in one functional")
        return _Integrand(self.coef * o.coef, self.scalars + o.scalars, self.vector or o.vector, self.direction if self.vector is not None else o.direction)

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0


def dot(u, n):
    """dot(velocity Function, FacetNormal | Constant vector): the flux factor of a boundary functional."""
    if isinstance(n, Function) and not isinstance(u, Function):
        u, n = n, u
    if not isinstance(u, Function) or len(u.blocks) != u.space.mesh.dim:
        raise NotImplementedError("dot: first argument must be a vector Function")
    if isinstance(n, FacetNormal):
        return _Integrand(n.sign, (), u, None)
    if isinstance(n, Constant):
        return _Integrand(1.0, (), u, tuple(float(v) for v in n.value))
    raise NotImplementedError("dot(u, %r)" % (n,))


class _Integral:
    def __init__(self, integrand, measure):
        self.integrand, self.measure = integrand, measure

    def __add__(self, other):
        return _IntegralSum([self, other])

    def __neg__(self):
        return _Integral(self.integrand * -1.0, self.measure)

    def __sub__(self, other):
        return _IntegralSum([self, -other])


class _IntegralSum:
    def __init__(self, terms):
        self.terms = list(terms)

    def __add__(self, other):
        return _IntegralSum(self.terms + (other.terms if isinstance(other, _IntegralSum) else [other]))


_INTEGRATORS = {}


def _integrator(mesh):
    from .functional import Integrator
    key = id(mesh)
    if key not in _INTEGRATORS:
        _INTEGRATORS[key] = (mesh, Integrator(default_context(), mesh))       # the mesh reference keeps id() valid
    return _INTEGRATORS[key][1]


def _assemble_integral(term):
    ig, ms = term.integrand, term.measure
    fns = list(ig.scalars) + ([ig.vector] if ig.vector is not None else [])
    if not fns:
        raise NotImplementedError("a functional needs at least one Function (the mesh comes from it)")
    mesh = fns[0].space.mesh
    it = _integrator(mesh)
    scal = []
    for f in ig.scalars:
        per = f.space.periodic
        scal.append((f.blocks[0], f.block_degree(0), per.vert2dof if per is not None else None))
    vec = None
    if ig.vector is not None:
        u = ig.vector
        if u.cell_velocity is not None:
            vec = ("cell", u.cell_velocity)
        else:
            vec = ("lagrange", u.blocks, u.block_degree(0))
    if ms.kind == "ds" and ms.sid is not None and ms.data is None:
        raise ValueError("ds(%r) needs subdomain_data (a FacetFunction)" % (ms.sid,))
    return it.integrate(ms.kind, scal, vec, ig.direction, ig.coef, ms.data, ms.sid)


def assemble(form, **kw):
    """dolfin.assemble for SCALAR functionals (run/CCS2D/CCS_3comp.py:277; run/stokes_ADR_3D/master_advective_stokes.py:232-265),
    evaluated on the GPU by mp_integrate.  Bilinear / linear forms are assembled inside solve(); use assemble_system for the shims."""
    if isinstance(form, _IntegralSum):
        return sum(_assemble_integral(t) for t in form.terms)
    if isinstance(form, _Integral):
        return _assemble_integral(form)
    raise NotImplementedError("assemble: only scalar functionals `integrand*dx` / `integrand*ds(id)` are supported here")


def errornorm(u, uh, norm_type="L2", degree_rise=3, mesh=None):
    """sqrt(int |u - uh|^2 dx) for two Functions of the same space (benchmark/sphere_3D_pure_shear/sphere.py:157-159: both sides are
    Functions of one space, for which DOLFIN's degree-raised interpolation reproduces the difference exactly); a non-Function `u`
    is interpolated into uh's space first."""
    if norm_type.lower() != "l2":
        raise NotImplementedError("errornorm: only the L2 norm")
    if not isinstance(u, Function):
        u = interpolate_like(u, uh)
    if len(u.blocks) != len(uh.blocks) or any(a.numel() != b.numel() for a, b in zip(u.blocks, uh.blocks)):
        raise NotImplementedError("errornorm: the two Functions must live in the same space")
    ctx = default_context()
    it = _integrator(uh.space.mesh)
    per = uh.space.periodic
    v2d = per.vert2dof if per is not None else None
    tot = 0.0
    for k, (a, b) in enumerate(zip(u.blocks, uh.blocks)):
        e = a.clone()
        ctx.axpby(-1.0, b, 1.0, e)
        deg = uh.block_degree(k)
        tot += it.integrate("dx", [(e, deg, v2d), (e, deg, v2d)])
    return math.sqrt(max(tot, 0.0))


def interpolate_like(v, f):
    """A new Function with f's blocks holding the interpolant of v."""
    g = Function(f.space, [b.clone() for b in f.blocks], list(f.ids))
    g.interpolate(v)
    return g


def project(v, V, **kw):
    """L2 projection onto V.  An Expression is first interpolated into its own element (degree <= V's: the projection of a function
    that already lies in V is the function itself -- `project(darcy.perm, X)`, run/CCS2D/CCS_3comp.py:293); a Function of another
    degree goes through the consistent mass system M_V x = int v phi_i (Jacobi-PCG on the GPU)."""
    f = Function(V)
    if isinstance(v, Function) and any(a.numel() != b.numel() for a, b in zip(v.blocks, f.blocks)):
        from . import moder
        ctx = default_context()
        if V.periodic is not None:
            raise NotImplementedError("project between degrees on a periodic space")
        sp = moder.Spaces(ctx, V.mesh)
        for k in range(len(f.blocks)):
            dt_, ds_ = f.block_degree(k), v.block_degree(k)
            rhs = ctx.zeros(sp.ndof(dt_))
            sp.assemble("mass", dt_, ds_).spmv(v.blocks[k], rhs)
            M = sp.assemble("mass", dt_, dt_)
            r = M.solve("pcg", rhs, f.blocks[k], rtol=1e-13, maxit=5000)
            if not r.converged:
                raise RuntimeError("project: mass solve did not converge (%r)" % (r,))
        return f
    f.interpolate(v)
    return f


# ------------------------------------------------------------------ assemble_system / KrylovSolver shims
class Matrix:
    """Placeholder for dolfin.Matrix(): the scripts only create it to receive assemble_system's result."""


class Vector:
    """Placeholder for dolfin.Vector()."""


class _AssembledSystem:
    def __init__(self, form, rhs_form, bcs):
        self.form, self.rhs_form, self.bcs = form, rhs_form, bcs


def assemble_system(a, L, bcs=None, **kw):
    """dolfin.assemble_system(a, L, bcs) (modules/mupopp.py:295-297,415-417; run/sphere_archer/sphere.py:199-202,256-258): on the B200
    path the element kernels, the scatter and the symmetric Dirichlet elimination run inside the solver call, so this only binds the
    form descriptors and the boundary conditions; KrylovSolver.solve(x, b) executes them."""
    bcs = [] if bcs is None else (list(bcs) if isinstance(bcs, (list, tuple)) else [bcs])
    s = _AssembledSystem(a, L, bcs)
    return s, s


class KrylovSolver:
    """dolfin.KrylovSolver(method, preconditioner) with .parameters, .set_operators(A, P), .solve(x, b)
    (modules/mupopp.py:289-303,410-425).  A, P, b come from assemble_system; x is `U.vector()`."""

    def __init__(self, method="default", preconditioner="default"):
        self.method, self.preconditioner = method, preconditioner
        self.parameters = {"relative_tolerance": 1.0e-6, "absolute_tolerance": 1.0e-15, "maximum_iterations": 10000,
                           "monitor_convergence": False, "nonzero_initial_guess": False, "error_on_nonconvergence": True, "report": True}
        self.A = self.P = None

    def set_operators(self, A, P=None):
        self.A, self.P = A, P

    def set_operator(self, A):
        self.A, self.P = A, None

    def solve(self, x, b):
        from . import mupopp as _mp
        if not isinstance(self.A, _AssembledSystem):
            raise TypeError("KrylovSolver: operators must come from assemble_system(a, L, bcs)")
        if not isinstance(x, _VectorView):
            raise TypeError("KrylovSolver.solve: pass `u.vector()` as the solution vector")
        return _mp._krylov_dispatch(self.A.form, x.f, self.A.bcs, float(self.parameters["relative_tolerance"]),
                                    int(self.parameters["maximum_iterations"]), bool(self.parameters["monitor_convergence"]))


class LUSolver:
    def __init__(self, A=None, method="default"):
        self.A = A

    def solve(self, x, b):
        from . import mupopp as _mp
        return _mp._krylov_dispatch(self.A.form, x.f, self.A.bcs, 1.0e-12, 100000, False)


import sys as _sys  # noqa: E402

dolfin = _sys.modules[__name__]          # scripts write `dolfin.FunctionSpace(...)` after `from fenics import *`
