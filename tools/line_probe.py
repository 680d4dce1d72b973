"""Per-iteration cost of the pressure PCG with the point-Jacobi and the lattice-line preconditioner on the bench operator (CUDA events),
for a sweep of L2 prefetch distances of the line sweeps.  usage: python tools/line_probe.py [n] [nz] [pd,pd,...]
(nz < n gives the per-GPU slab of an N-GPU run).  It is also the ncu target for k_line_fwd / k_line_bwd (PROBE_ONLY=<pd>)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200._lib import check, ptr
from mupopp_b200.device import Context, DeviceMesh, Plan
from mupopp_b200.mesh import box_periodic_map

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nz = int(sys.argv[2]) if len(sys.argv) > 2 else n
pds = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 8, 16, 32, 64]
only = os.environ.get("PROBE_ONLY")
ctx = Context(0)
mesh = mp.BoxMesh((0, 0, 0), (4, 4, 1.0 * nz / n), n, n, nz)
pmap = box_periodic_map(mesh, (0, 1))
dm = DeviceMesh(ctx, mesh, pmap.vert2dof)
plan = Plan(ctx, dm.cell_dofs, dm.cell_dofs, dm.ndof, dm.ndof)
A = plan.new_matrix()
emat = ctx.empty(dm.ncell * 16)
check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 20.0, ptr(emat)))
plan.scatter_matrix(emat, A)
del emat
N = A.nrow
isbc = np.zeros(N, dtype=np.uint8)
isbc[: n * n] = 1                                           # p = 0 on the bottom plane
b = ctx.to_device(np.random.default_rng(0).standard_normal(N))
A.apply_dirichlet(b, ctx.to_device(isbc), ctx.zeros(N))
A.jacobi_setup()
stride = n * n
A.line_setup(stride)
x = ctx.zeros(N)


def per_iteration(method, **kw):
    out = []
    for maxit in (20, 20, 120):                             # the first call warms up
        x.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.solve(method, b, x, rtol=1e-300, maxit=maxit, refresh_jacobi=False, **kw)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return (out[2] - out[1]) / 100.0


if only is None:
    t = per_iteration("pcg")
    print("N=%d: Jacobi-PCG iteration %.4f ms (SpMV + 116 N bytes of vector traffic)" % (N, t))
for pd in ([int(only)] if only is not None else pds):
    ctx.set_option("line_prefetch", pd)
    t = per_iteration("pcg_line", line_stride=stride)
    print("line-PCG iteration, L2 prefetch %3d planes ahead: %.4f ms (SpMV + 104 N bytes = %.0f MB in the two sweeps)" % (pd, t, 104.0 * N / 1e6))
