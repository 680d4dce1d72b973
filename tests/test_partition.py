"""Host-side multi-GPU logic on CPU: slab partition + halo plan, exercised with world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest

import mupopp_b200 as mp
from mupopp_b200 import partition
from oracle import fem, forms

BOX = ((0.0, 0.0, 0.0), (2.0, 1.0, 1.0))
GRID = (3, 2, 5)


@pytest.mark.parametrize("nranks", [1, 2, 3, 6])
def test_slabs_cover_mesh_and_rows_are_complete(nranks):
    g = mp.BoxMesh(BOX[0], BOX[1], *GRID)
    og = fem.box_mesh(BOX[0], BOX[1], *GRID)
    Mg = forms.Assembler(og).mass(1, 1, 2).tocsr()
    owned_all = []
    for r in range(nranks):
        p = partition.slab_partition(BOX, GRID, r, nranks)
        assert np.array_equal(p.mesh.coords, g.coords[p.global_ids])
        owned_all.append(p.global_ids[: p.n_owned])
        # every local cell is a global cell (as a vertex set) and owned rows assemble completely
        loc_cells = np.sort(p.global_ids[p.mesh.cells], axis=1)
        gset = set(map(tuple, g.cells.tolist()))
        assert all(tuple(c) in gset for c in loc_cells.tolist())
        om = fem.Mesh(p.mesh.coords, np.sort(p.mesh.cells.astype(np.int64), axis=1))
        Ml = forms.Assembler(om).mass(1, 1, 2).tocsr()
        for i in range(0, p.n_owned, 7):
            row_l = Ml.getrow(i)
            cols = p.global_ids[row_l.indices]
            row_g = Mg.getrow(p.global_ids[i])
            o1, o2 = np.argsort(cols), np.argsort(row_g.indices)
            assert np.array_equal(cols[o1], row_g.indices[o2])
            assert np.allclose(row_l.data[o1], row_g.data[o2], rtol=1e-14, atol=0)
    allv = np.concatenate(owned_all)
    assert np.array_equal(np.sort(allv), np.arange(g.num_vertices()))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = partition.slab_partition(BOX, GRID, rank, world)
        nglob = (GRID[0] + 1) * (GRID[1] + 1) * (GRID[2] + 1)
        xg = np.sin(np.arange(nglob) * 0.37)
        x = np.zeros(p.n_local)
        x[: p.n_owned] = xg[p.global_ids[: p.n_owned]]
        # the same exchange mp_halo_exchange performs with ncclSend/ncclRecv, here over gloo
        reqs, bufs = [], []
        for (nr, send_idx, rstart, rcount) in p.neigh:
            sb = torch.from_numpy(x[send_idx].copy())
            rb = torch.empty(rcount, dtype=torch.float64)
            reqs.append(dist.isend(sb, nr))
            reqs.append(dist.irecv(rb, nr))
            bufs.append((rb, rstart, rcount))
        for r in reqs:
            r.wait()
        for rb, rstart, rcount in bufs:
            x[rstart:rstart + rcount] = rb.numpy()
        ok = np.array_equal(x, xg[p.global_ids])
        # distributed dot product = allreduce of owned parts
        t = torch.tensor([float(x[: p.n_owned] @ x[: p.n_owned])], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and abs(t.item() - xg @ xg) < 1e-12 * (xg @ xg)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_gloo(world):
    import torch.multiprocessing as tmp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctxm = tmp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


@pytest.mark.parametrize("nranks", [1, 2, 4])
def test_periodic_slabs_match_global_periodic_map(nranks):
    """Periodic x/y faces: the local dof numbering of every slab maps onto the global periodic dofs."""
    from mupopp_b200.mesh import box_periodic_map
    g = mp.BoxMesh(BOX[0], BOX[1], *GRID)
    pg = box_periodic_map(g, (0, 1))
    assert pg.ndof == GRID[0] * GRID[1] * (GRID[2] + 1)
    npl = (GRID[0] + 1) * (GRID[1] + 1)
    cover = []
    for r in range(nranks):
        p = partition.slab_partition(BOX, GRID, r, nranks, periodic_xy=True)
        k0 = max(p.planes[0] - 1, 0)
        gv = k0 * npl + np.arange(p.mesh.num_vertices())
        assert np.array_equal(p.global_ids[p.vert2dof], pg.vert2dof[gv])
        assert p.n_local == p.vert2dof.max() + 1
        for (nr, send_idx, rstart, rcount) in p.neigh:
            assert len(send_idx) == rcount == GRID[0] * GRID[1] and rstart >= p.n_owned
        cover.append(p.global_ids[: p.n_owned])
    assert np.array_equal(np.sort(np.concatenate(cover)), np.arange(pg.ndof))


def _general_mesh(kind):
    if kind == "demo":
        import os
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fenics_demo_dofmap.npz"))
        return mp.Mesh(g["coords"], g["cells"])
    return mp.BoxMesh(BOX[0], BOX[1], *GRID)


@pytest.mark.parametrize("kind,nranks", [("demo", 2), ("demo", 5), ("box", 3), ("box", 8)])
def test_rcb_partition_is_consistent(kind, nranks):
    """General (unstructured) partition: balanced RCB ownership, complete owned rows, and halo plans whose send/recv lists pair up."""
    m = _general_mesh(kind)
    owner = partition.rcb_owner(m.coords, nranks)
    counts = np.bincount(owner, minlength=nranks)
    assert counts.max() - counts.min() <= nranks          # balanced to within the bisection rounding
    parts = [partition.general_partition(m, owner, r) for r in range(nranks)]
    og = fem.Mesh(m.coords, m.cells.astype(np.int64))
    Mg = forms.Assembler(og).mass(1, 1, 2).tocsr()
    xg = np.cos(0.31 * np.arange(m.num_vertices()))
    cover = []
    for p in parts:
        assert np.array_equal(p.mesh.coords, m.coords[p.global_ids])
        assert (owner[p.global_ids[: p.n_owned]] == p.rank).all() and (owner[p.global_ids[p.n_owned:]] != p.rank).all()
        cover.append(p.global_ids[: p.n_owned])
        ol = fem.Mesh(p.mesh.coords, np.sort(p.mesh.cells.astype(np.int64), axis=1))
        Ml = forms.Assembler(ol).mass(1, 1, 2).tocsr()
        x = np.zeros(p.n_local)
        x[: p.n_owned] = xg[p.global_ids[: p.n_owned]]
        # emulate the halo exchange with the partner's send list
        for (q, send_idx, rstart, rcount) in p.neigh:
            back = [n for n in parts[q].neigh if n[0] == p.rank]
            assert len(back) == 1
            qsend = back[0][1]
            assert len(qsend) == rcount
            x[rstart:rstart + rcount] = xg[parts[q].global_ids[qsend]]
        assert np.array_equal(x, xg[p.global_ids])
        # owner-computes: local rows of owned vertices equal the global rows
        yl = (Ml @ x)[: p.n_owned]
        assert np.allclose(yl, (Mg @ xg)[p.global_ids[: p.n_owned]], rtol=1e-13, atol=1e-15)
    assert np.array_equal(np.sort(np.concatenate(cover)), np.arange(m.num_vertices()))


def test_general_partition_boundary_data_matches_global():
    """Cut faces must not be taken for boundary: Dirichlet flags and the Neumann flux load of every owned vertex equal the global ones."""
    from mupopp_b200.transport import make_boundary_data
    m = mp.BoxMesh(BOX[0], BOX[1], *GRID)
    top = lambda x: x[..., 2] > 1.0 - 1e-12
    gbd = make_boundary_data(m, top, (0.0, 0.0, -0.5), top, 0.1)
    owner = partition.rcb_owner(m.coords, 4)
    for r in range(4):
        p = partition.general_partition(m, owner, r)
        bd = make_boundary_data(p.mesh, top, (0.0, 0.0, -0.5), top, 0.1)
        own = p.global_ids[: p.n_owned]
        assert np.array_equal(bd.p_isbc[: p.n_owned], gbd.p_isbc[own])
        assert np.array_equal(bd.c0_isbc[: p.n_owned], gbd.c0_isbc[own])
        assert np.allclose(bd.flux_load[: p.n_owned], gbd.flux_load[own], rtol=0, atol=1e-15)
        # ghost columns need the right Dirichlet flags too (symmetric elimination zeroes those columns): restrict the global data
        rb = partition.restrict_boundary_data(p, gbd)
        assert np.array_equal(rb.p_isbc, gbd.p_isbc[p.global_ids]) and np.array_equal(rb.p_isbc[: p.n_owned], bd.p_isbc[: p.n_owned])
        assert np.allclose(rb.flux_load[: p.n_owned], bd.flux_load[: p.n_owned], rtol=0, atol=1e-15)
