import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200 import partition
from mupopp_b200.device import Context, DeviceMesh, Plan
from mupopp_b200._lib import check, ptr
import ctypes as C
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
ids = [ctx.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.init_comm(world, rank, ids[0])
n = 8
BOX = ((0.0, 0.0, 0.0), (2.0, 2.0, 1.0))
part = partition.slab_partition(BOX, (n, n, n), rank, world)
halo = partition.make_halo(ctx, part)
x = torch.full((part.n_local,), -7.0, dtype=torch.float64, device=ctx.device)
x[: part.n_owned] = torch.from_numpy(part.global_ids[: part.n_owned].astype(np.float64)).to(ctx.device)
halo.exchange(x)
torch.cuda.synchronize()
err = np.abs(x.cpu().numpy() - part.global_ids).max()
print("rank", rank, "halo err", err, "n_owned", part.n_owned, "n_local", part.n_local, "neigh", [(a, len(b), c, d) for a, b, c, d in part.neigh], flush=True)
# allreduce
t = torch.tensor([1.0 + rank, 10.0], dtype=torch.float64, device=ctx.device)
check(ctx.lib.mp_allreduce_sum(ctx.h, ptr(t), 2))
torch.cuda.synchronize()
print("rank", rank, "allreduce", t.cpu().numpy(), flush=True)
# distributed dot
print("rank", rank, "dot", ctx.dot(x[:part.n_owned], x[:part.n_owned]), float((np.arange((n+1)**3, dtype=np.float64)**2).sum()), flush=True)
# distributed mass SpMV vs global
dm = DeviceMesh(ctx, part.mesh)
rows = torch.where(dm.cells < part.n_owned, dm.cells, torch.full_like(dm.cells, -1))
plan = Plan(ctx, rows, dm.cells, part.n_owned, part.n_local)
M = plan.new_matrix()
emat = ctx.empty(dm.ncell * 16)
check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(dm.c), 1.0, ptr(emat)))
plan.scatter_matrix(emat, M)
y = ctx.zeros(part.n_owned)
xs = torch.sin(torch.from_numpy(part.global_ids.astype(np.float64))).to(ctx.device)
M.spmv(xs, y)
torch.cuda.synchronize()
g = mp.BoxMesh(BOX[0], BOX[1], n, n, n)
dg = DeviceMesh(ctx, g)
pg = Plan(ctx, dg.cells, dg.cells, dg.nvert, dg.nvert)
Mg = pg.new_matrix()
eg = ctx.empty(dg.ncell * 16)
check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(dg.c), 1.0, ptr(eg)))
pg.scatter_matrix(eg, Mg)
yg = ctx.zeros(dg.nvert)
xg = torch.sin(torch.arange(dg.nvert, dtype=torch.float64)).to(ctx.device)
Mg.spmv(xg, yg)
torch.cuda.synchronize()
ref = yg.cpu().numpy()[part.global_ids[: part.n_owned]]
print("rank", rank, "spmv err", np.abs(y.cpu().numpy() - ref).max(), np.abs(ref).max(), flush=True)
# solve M c = b distributed
b = y.clone()
c = ctx.zeros(part.n_local)
r = M.solve("pcg", b, c, rtol=1e-12, maxit=200, halo=halo, check_every=1)
torch.cuda.synchronize()
print("rank", rank, r, "sol err", np.abs(c.cpu().numpy()[: part.n_owned] - np.sin(part.global_ids[: part.n_owned])).max(), flush=True)
dist.barrier()
dist.destroy_process_group()
