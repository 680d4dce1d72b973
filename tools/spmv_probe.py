"""SpMV probe at the bench operator: times the SELL variants with CUDA events (and is the ncu target for the dominant kernel).
usage: python tools/spmv_probe.py [n] [nz] [reps]   (nz < n gives the per-GPU slab of an N-GPU run, e.g. 200 25 for 8 GPUs)"""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
only = os.environ.get("PROBE_ONLY")               # "c8,lpr" restricts the run to one variant (ncu captures)
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
A.sync_sell()
x = ctx.to_device(np.random.default_rng(0).standard_normal(dm.ndof))
y = ctx.zeros(dm.ndof)
N, nnz = A.nrow, A.nnz
print("operator: N=%d nnz=%d SELL entries=%d (padding %.2f%%), explicit (slice,j) %d of %d (%.1f%%)" % (
    N, nnz, A.sell_entries, 100.0 * (A.sell_entries - nnz) / nnz, A.sell_explicit, A.sell_entries // 32, 100.0 * A.sell_explicit / max(A.sell_entries // 32, 1)))
b12 = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
bcu = 8.0 * A.sell_entries + 4.0 * (A.sell_entries / 32) + 128.0 * max(A.sell_explicit, 0) + 4.0 * (N / 32 + 1) + 16.0 * N
variants = [(0, 1), (1, 1), (1, 2), (1, 4), (0, 2), (0, 4)]
if only:
    variants = [tuple(int(v) for v in only.split(","))]
ref = None
for c8, lpr in variants:
    ctx.set_option("spmv_c8", c8)
    ctx.set_option("spmv_lpr", lpr)
    t0 = __import__("time").time()
    while __import__("time").time() - t0 < float(os.environ.get("PROBE_WARM", "1.0")):          # clocks ramp for a while after the process starts: warm up for a full second
        for _ in range(50):
            A.spmv(x, y)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        A.spmv(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    yy = y.cpu().numpy()
    if ref is None:
        ref = yy.copy()
    err = np.linalg.norm(yy - ref) / np.linalg.norm(ref)
    bts = bcu if c8 else b12
    print("c8=%d lpr=%d: %.4f ms  %.0f GB/s of streamed bytes (%.3f GB), %.0f GB/s by the 12 B/nnz formula; rel diff %.1e" % (
        c8, lpr, ms, bts / ms / 1e6, bts / 1e9, b12 / ms / 1e6, err))
