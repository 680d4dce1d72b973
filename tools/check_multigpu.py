"""Multi-GPU correctness check (run under torchrun, one rank per GPU): the z-slab partitioned mode-F solver must reproduce
the single-GPU fields.  Rank 0 also runs the unpartitioned problem and compares after a few steps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200 import partition
from mupopp_b200.device import Context
from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nz = int(sys.argv[3]) if len(sys.argv) > 3 else n          # nz != n gives ranks of different size (collective control-flow check)
per = len(sys.argv) > 4 and sys.argv[4] == "periodic"      # periodic x/y faces
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
ids = [ctx.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.init_comm(world, rank, ids[0])
transport = os.environ.get("MUPOPP_B200_TRANSPORT", "peer")   # peer (NVLink windows, default) | nccl
if transport == "peer":
    on = ctx.enable_peer()
    if rank == 0:
        print("peer-memory collectives:", "on" if on else "unavailable (%s) -> NCCL" % getattr(ctx, "peer_error", "?"))
BOX = ((0.0, 0.0, 0.0), (2.0, 2.0, 1.0))
model = TransportModel.ccs_three_component(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zh=(0.0, 0.0, 1.0))
top = lambda x: x[..., 2] > 1.0 - 1e-12
umark = lambda x: x[..., 2] > 1e-12
uval = lambda x: np.stack([0 * x[..., 0], 0 * x[..., 0], np.where(x[..., 2] > 1.0 - 1e-12, -0.5, 0.0)], -1)
src = lambda X: np.abs(np.sin(3 * np.pi * X[:, 0])) * (1 - np.tanh((1.0 - X[:, 2]) / 0.1))
rcb = len(sys.argv) > 4 and sys.argv[4] == "rcb"          # general (unstructured-style) RCB partition instead of z-slabs
if per:
    umark, uval = top, (0.0, 0.0, -0.5)
if rcb:
    gmesh = mp.BoxMesh(BOX[0], BOX[1], n, n, nz)
    part = partition.general_partition(gmesh, partition.rcb_owner(gmesh.coords, world), rank)
    bd = partition.restrict_boundary_data(part, make_boundary_data(gmesh, umark, uval, top, 0.1))
else:
    part = partition.slab_partition(BOX, (n, n, nz), rank, world, periodic_xy=per)
    bd = partition.slab_boundary_data(part, umark, uval, top, 0.1)
halo = partition.make_halo(ctx, part)
fs = src(part.mesh.coords)
s = TransportSolverF(ctx, part.mesh, model, bd, f_source=part.periodic.restrict(fs) if per else fs, rtol=1e-13, halo=halo,
                     n_owned=part.n_owned, vert2dof=part.vert2dof)
s.set_state(c1=0.1 * np.ones(part.n_local))
its = []
for k in range(nsteps):
    r = s.step()
    its.append({a: (b.niter if b else 0) for a, b in r.items()})
torch.cuda.synchronize()
# gather owned parts on rank 0
nglob = (n * n if per else (n + 1) ** 2) * (nz + 1)
out = {}
for name in ("c0", "c1", "c2", "p"):
    full = torch.zeros(nglob, dtype=torch.float64, device=ctx.device)
    gid = torch.from_numpy(part.global_ids[: part.n_owned]).to(ctx.device)
    full[gid] = getattr(s, name)[: part.n_owned]
    dist.all_reduce(full)
    out[name] = full.cpu().numpy()
ok = True
if rank == 0:
    ctx1 = Context(local)
    mesh = mp.BoxMesh(BOX[0], BOX[1], n, n, nz)
    from mupopp_b200.mesh import box_periodic_map
    pm1 = box_periodic_map(mesh, (0, 1)) if per else None
    bd1 = make_boundary_data(mesh, umark, uval, top, 0.1, periodic=pm1)
    fs1 = src(mesh.coords)
    s1 = TransportSolverF(ctx1, mesh, model, bd1, f_source=pm1.restrict(fs1) if per else fs1, rtol=1e-13,
                          vert2dof=pm1.vert2dof if per else None)
    s1.set_state(c1=0.1 * np.ones(s1.dm.ndof))
    its1 = []
    for k in range(nsteps):
        r = s1.step()
        its1.append({a: (b.niter if b else 0) for a, b in r.items()})
    for name in ("c0", "c1", "c2", "p"):
        ref = getattr(s1, name).cpu().numpy()
        err = np.linalg.norm(out[name] - ref) / max(np.linalg.norm(ref), 1e-300)
        print("multi-GPU x%d vs single: %s rel L2 %.3e" % (world, name, err))
        ok = ok and err < 1e-8
    print("iterations partitioned:", its[-1], " single:", its1[-1])
    print("MULTIGPU_OK" if ok else "MULTIGPU_FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
