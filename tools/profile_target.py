"""Small fixed workload for ncu: a few launches of each hot kernel at n^3 (default 200 -> 8.1M dofs)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200._lib import check, ptr
from mupopp_b200.device import Context, DeviceMesh, Plan, transport_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = Context(0)
mesh = mp.BoxMesh((0, 0, 0), (4, 4, 1), n, n, n)
dm = DeviceMesh(ctx, mesh)
plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
N, Cn = dm.nvert, dm.ncell
emat = ctx.empty(Cn * 16)
evec = ctx.empty(Cn * 4)
A = plan.new_matrix()
prm = transport_params(500, 0.1, 0.05, 1e-8, 0.1, 1.0, 1.0, 1.0, (62 / 698., 278 / 698., 100 / 698.), (0, 0, 1.0), 1.0, "c0c1", 3, 1)
x = torch.rand(N, dtype=torch.float64, device=ctx.device)
z = torch.rand(N, dtype=torch.float64, device=ctx.device)
y = torch.empty_like(x)
u = torch.rand(Cn * 3, dtype=torch.float64, device=ctx.device)
for rep in range(2):
    check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 20.0, ptr(emat)))
    plan.scatter_matrix(emat, A)
    check(ctx.lib.mp_elem_p1_transport(ctx.h, C.byref(dm.c), C.byref(prm), ptr(u), ptr(x), ptr(z), ptr(x), ptr(z), ptr(x), ptr(z), ptr(emat), ptr(evec)))
    plan.scatter_vector(evec, y)
    A.spmv(x, y)
isbc = torch.zeros(N, dtype=torch.uint8, device=ctx.device)
isbc[:(n + 1) ** 2] = 1
A.apply_dirichlet(None, isbc, torch.zeros(N, dtype=torch.float64, device=ctx.device))
b = torch.rand(N, dtype=torch.float64, device=ctx.device)
xs = torch.zeros(N, dtype=torch.float64, device=ctx.device)
r = A.solve("pcg", b, xs, rtol=1e-3, maxit=6, check_every=6)
xs.zero_()
r2 = A.solve("gmres", b, xs, rtol=1e-3, maxit=12, check_every=12)
torch.cuda.synchronize()
print("ok", r, r2)
