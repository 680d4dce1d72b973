import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200 import partition
from mupopp_b200.device import Context
from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
ids = [ctx.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.init_comm(world, rank, ids[0])
n = int(sys.argv[1])
BOX = ((0.0, 0.0, 0.0), (2.0, 2.0, 1.0))
model = TransportModel.ccs_three_component(Pe=500, Da=0.1, phi=0.05, alpha=1e-3, dt=0.1, K=1.0, zh=(0.0, 0.0, 1.0))
top = lambda x: x[..., 2] > 1.0 - 1e-12
umark = lambda x: x[..., 2] > 1e-12
uval = lambda x: np.stack([0 * x[..., 0], 0 * x[..., 0], np.where(x[..., 2] > 1.0 - 1e-12, -0.5, 0.0)], -1)
src = lambda X: np.abs(np.sin(3 * np.pi * X[:, 0])) * (1 - np.tanh((1.0 - X[:, 2]) / 0.1))
part = partition.slab_partition(BOX, (n, n, n), rank, world)
bd = partition.slab_boundary_data(part, umark, uval, top, 0.1)
halo = partition.make_halo(ctx, part)
s = TransportSolverF(ctx, part.mesh, model, bd, f_source=src(part.mesh.coords), rtol=1e-13, halo=halo, n_owned=part.n_owned, maxit=300, check_every=1)
s.set_state(c1=0.1 * np.ones(part.n_local))
for k in range(3):
    try:
        r = s.step()
        print("rank", rank, "step", k, r, flush=True)
    except Exception as e:
        print("rank", rank, "step", k, "EXC", e, flush=True)
        for nm in ("c0", "c1", "c2", "p", "rhs", "c1_new", "c0_new"):
            v = getattr(s, nm)
            print("rank", rank, nm, "nan" if torch.isnan(v).any().item() else "finite", float(v.abs().max()), flush=True)
        break
dist.barrier()
dist.destroy_process_group()
