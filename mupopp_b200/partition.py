"""Mesh partition across GPUs: z-slabs of a structured BoxMesh, owner-computes rows, one ghost layer.

Replaces DOLFIN's MPI mesh distribution (SCOTCH) + PETSc ghost updates that the reference uses when run
with `aprun -n 24|48` (/root/reference/src/run/sphere_archer/submit.sh:25, run/LVL_archer/submit.sh:25).
Local numbering per rank: [owned vertices (global order) | lower ghost plane | upper ghost plane]; every cell that
touches an owned vertex is local, so owned rows assemble completely without a reverse (add) halo.
The halo of a vector is two plane-sized ncclSend/ncclRecv pairs (<= 2 neighbours).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesh import BoxMesh, Mesh


@dataclass
class SlabPartition:
    rank: int
    nranks: int
    grid: tuple            # (nx, ny, nz) global hexes
    planes: tuple          # (a, b): owned vertex planes a <= k < b
    mesh: Mesh             # local mesh in local numbering
    n_owned: int
    n_local: int
    global_ids: np.ndarray  # local dof -> global dof id (== vertex ids unless periodic)
    neigh: list            # [(rank, send_local_idx ndarray, recv_start, recv_count), ...]
    vert2dof: np.ndarray = None   # periodic x/y: local P1 dof of every local vertex
    periodic: object = None       # mesh.PeriodicMap of the local mesh (periodic x/y)


def plane_ranges(nz, nranks):
    """Split the nz+1 vertex planes as evenly as possible; returns [(a_r, b_r)]."""
    tot = nz + 1
    base, rem = divmod(tot, nranks)
    out, a = [], 0
    for r in range(nranks):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out


def slab_partition(box, grid, rank, nranks, periodic_xy=False):
    """`periodic_xy`: the x and y faces are periodic (constrained_domain of run/CCS/CCS_3D_top_source.py:134-158): each
    vertex plane then carries nx*ny dofs; the z-slab structure and the halo pattern are unchanged."""
    if periodic_xy:
        return _slab_partition_periodic(box, grid, rank, nranks)
    nx, ny, nz = grid
    if nranks > nz + 1:
        raise ValueError("more ranks than vertex planes")
    a, b = plane_ranges(nz, nranks)[rank]
    if b <= a:
        raise ValueError("empty slab")
    k0, k1 = max(a - 1, 0), min(b, nz)              # hex layers k0..k1-1 ; vertex planes k0..k1
    m = BoxMesh(box[0], box[1], nx, ny, nz, z_range=(k0, k1))
    npl = (nx + 1) * (ny + 1)
    nplanes = k1 - k0 + 1
    has_lo, has_hi = a > 0, b <= nz
    # slab order: plane k0 .. k1 ; new order: owned planes a..b-1, then plane a-1 (lower ghost), then plane b (upper ghost)
    plane_new_start = {}
    pos = 0
    for k in range(a, b):
        plane_new_start[k] = pos
        pos += npl
    n_owned = pos
    if has_lo:
        plane_new_start[a - 1] = pos
        pos += npl
    if has_hi:
        plane_new_start[b] = pos
        pos += npl
    assert pos == nplanes * npl
    old2new = np.empty(nplanes * npl, dtype=np.int64)
    for k in range(k0, k1 + 1):
        old2new[(k - k0) * npl:(k - k0 + 1) * npl] = plane_new_start[k] + np.arange(npl)
    coords = np.empty_like(m.coords)
    coords[old2new] = m.coords
    cells = old2new[m.cells].astype(np.int32)
    loc = Mesh(coords, cells)
    loc.box = (np.array(box[0], dtype=float), np.array(box[1], dtype=float))
    loc.grid = grid
    gids = np.empty(nplanes * npl, dtype=np.int64)
    for k in range(k0, k1 + 1):
        gids[plane_new_start[k]:plane_new_start[k] + npl] = k * npl + np.arange(npl)
    neigh = []
    if has_lo:   # lower neighbour needs my first owned plane; I receive its last owned plane (a-1)
        neigh.append((rank - 1, np.arange(0, npl, dtype=np.int32), plane_new_start[a - 1], npl))
    if has_hi:
        neigh.append((rank + 1, np.arange(n_owned - npl, n_owned, dtype=np.int32), plane_new_start[b], npl))
    return SlabPartition(rank, nranks, grid, (a, b), loc, n_owned, nplanes * npl, gids, neigh)


def _slab_partition_periodic(box, grid, rank, nranks):
    from .mesh import PeriodicMap
    nx, ny, nz = grid
    if nranks > nz + 1:
        raise ValueError("more ranks than vertex planes")
    a, b = plane_ranges(nz, nranks)[rank]
    k0, k1 = max(a - 1, 0), min(b, nz)
    m = BoxMesh(box[0], box[1], nx, ny, nz, z_range=(k0, k1))
    m.box = (np.array(box[0], dtype=float), np.array(box[1], dtype=float))
    npd = nx * ny                                   # dofs per plane
    has_lo, has_hi = a > 0, b <= nz
    start, pos = {}, 0
    for k in range(a, b):
        start[k] = pos
        pos += npd
    n_owned = pos
    if has_lo:
        start[a - 1] = pos
        pos += npd
    if has_hi:
        start[b] = pos
        pos += npd
    n_local = pos
    ix = np.tile(np.arange(nx + 1), ny + 1) % nx
    iy = np.repeat(np.arange(ny + 1), nx + 1) % ny
    inplane = iy * nx + ix
    v2d = np.concatenate([start[k] + inplane for k in range(k0, k1 + 1)]).astype(np.int32)
    gids = np.empty(n_local, dtype=np.int64)
    for k in range(k0, k1 + 1):
        gids[start[k]:start[k] + npd] = k * npd + np.arange(npd)
    neigh = []
    if has_lo:
        neigh.append((rank - 1, np.arange(0, npd, dtype=np.int32), start[a - 1], npd))
    if has_hi:
        neigh.append((rank + 1, np.arange(n_owned - npd, n_owned, dtype=np.int32), start[b], npd))
    pmap = PeriodicMap(v2d, (0, 1), m.box)
    return SlabPartition(rank, nranks, grid, (a, b), m, n_owned, n_local, gids, neigh, v2d, pmap)


def slab_boundary_data(part: SlabPartition, u_marker, u_value, c0_marker=None, c0_value=0.0):
    from .transport import make_boundary_data
    return make_boundary_data(part.mesh, u_marker, u_value, c0_marker, c0_value, periodic=part.periodic)


def make_halo(ctx, part: SlabPartition, peer=None):
    """Halo plan of a partition.  `peer` (default: on when the context's peer-memory windows are attached) also creates the ghost
    window that lets the SpMV kernel exchange interface dofs over NVLink by itself; collective over all ranks."""
    from .device import Halo
    h = Halo(ctx, [n[0] for n in part.neigh], [n[1] for n in part.neigh], [n[2] for n in part.neigh], [n[3] for n in part.neigh])
    if peer is None:
        peer = ctx.nranks > 1 and ctx.peer_enabled()
    if peer:
        h.enable_peer(part.n_owned, part.n_local)
    return h


# ------------------------------------------------------------------ general meshes: recursive coordinate bisection
def rcb_owner(coords, nranks):
    """Recursive coordinate bisection of the vertices into `nranks` parts of (nearly) equal size (general
    meshes have no METIS/SCOTCH here).  Deterministic: ties are broken by vertex id.  Returns owner[v] in [0, nranks)."""
    owner = np.zeros(coords.shape[0], dtype=np.int32)

    def split(idx, r0, nr):
        if nr == 1:
            owner[idx] = r0
            return
        nl = nr // 2
        x = coords[idx]
        axis = int(np.argmax(x.max(0) - x.min(0)))
        order = np.lexsort((idx, x[:, axis]))
        cut = (len(idx) * nl) // nr
        split(idx[order[:cut]], r0, nl)
        split(idx[order[cut:]], r0 + nl, nr - nl)

    split(np.arange(coords.shape[0]), 0, nranks)
    return owner


def general_partition(mesh, owner, rank):
    """Owner-computes partition of an arbitrary simplicial mesh for `rank`: local cells = every cell with an owned vertex; local vertex
    numbering = [owned (ascending global id) | ghosts grouped by owning rank (ascending rank, then global id)], so each neighbour's
    ghost block is contiguous, as mp_halo_exchange requires.  The send list to neighbour q is the ascending list of my vertices that
    are ghosts on q -- the same order in which q stores them."""
    owner = np.asarray(owner)
    nranks = int(owner.max()) + 1
    cells = mesh.cells.astype(np.int64)
    cell_owner = owner[cells]                                   # (C, d+1)
    mine = (cell_owner == rank).any(1)
    lc = cells[mine]
    verts = np.unique(lc.ravel())
    owned = verts[owner[verts] == rank]
    ghosts = verts[owner[verts] != rank]
    gorder = np.lexsort((ghosts, owner[ghosts]))
    ghosts = ghosts[gorder]
    local_of = np.full(mesh.num_vertices(), -1, dtype=np.int64)
    gids = np.concatenate([owned, ghosts])
    local_of[gids] = np.arange(gids.size)
    loc = Mesh(mesh.coords[gids], local_of[lc].astype(np.int32))
    # true boundary facets only (cut faces of the partition are not boundary): restrict the global list to the local cells
    gfv, gn, gmeas, gown = mesh.boundary_facets()
    gloc = mesh.boundary_facet_local_index()
    cell_local = np.full(mesh.num_cells(), -1, dtype=np.int64)
    cell_local[np.nonzero(mine)[0]] = np.arange(int(mine.sum()))
    keep = cell_local[gown] >= 0
    loc._bfacets = (local_of[gfv[keep]], gn[keep].copy(), gmeas[keep].copy(), cell_local[gown[keep]])
    loc._bfacet_local = gloc[keep].copy()
    neigh = []
    for q in range(nranks):
        if q == rank:
            continue
        recv = ghosts[owner[ghosts] == q]                       # contiguous block in my numbering
        # what q needs from me: my owned vertices appearing in cells where q owns a vertex
        qcells = cells[(cell_owner == q).any(1)]
        qv = np.unique(qcells.ravel())
        send = qv[owner[qv] == rank]                            # ascending global id == q's storage order of its ghosts owned by me
        if recv.size == 0 and send.size == 0:
            continue
        rstart = int(local_of[recv[0]]) if recv.size else int(gids.size)
        neigh.append((q, local_of[send].astype(np.int32), rstart, int(recv.size)))
    return SlabPartition(rank, nranks, None, None, loc, int(owned.size), int(gids.size), gids, neigh)


def restrict_boundary_data(part: SlabPartition, global_bdata):
    """Boundary data of a general partition from the data of the GLOBAL mesh (make_boundary_data on the whole mesh, computed once per
    process): exact Dirichlet flags on ghost columns too, which a purely local facet search cannot guarantee for arbitrary cuts."""
    from .transport import BoundaryDataF
    g = part.global_ids
    return BoundaryDataF(global_bdata.p_isbc[g].copy(), global_bdata.flux_load[g].copy(), global_bdata.c0_isbc[g].copy(),
                         global_bdata.c0_g[g].copy())
