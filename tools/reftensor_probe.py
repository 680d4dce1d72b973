"""P2 x P2 reference-tensor element kernel + scatter (the compaction / mode-R / Stokes block assembly) at a bandwidth-bound size, CUDA events:
cell-ordered element tensor + generic scatter against row-ordered output + streaming scatter.  usage: python tools/reftensor_probe.py [n] [deg]
Also the ncu target for k_reftensor / k_scatter_matrix (PROBE_ONLY=rows)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mupopp_b200 as mp  # noqa: E402
from mupopp_b200._lib import check, ptr  # noqa: E402
from mupopp_b200.device import Context  # noqa: E402
from mupopp_b200.moder import Spaces  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
only = os.environ.get("PROBE_ONLY")
ctx = Context(0)
mesh = mp.BoxMesh((-2.0, -2.0, 0.0), (2.0, 2.0, 8.0), n, n, 2 * n)
sp = Spaces(ctx, mesh)
nb = sp.nb(deg)
plan = sp.plan(deg, deg)
T0 = sp.tensor("stiff", deg, deg, True)
coef = ctx.to_device(0.9 + 0.05 * np.random.default_rng(0).random(sp.n1))
ncell = mesh.num_cells()
emat = ctx.empty(ncell * nb * nb)
A = plan.new_matrix(sell=False)
B = plan.new_matrix(sell=False)
m = sp.dm.c
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


b_el = 16.0 * ncell + 24.0 * mesh.num_vertices() + 8.0 * sp.n1 + 8.0 * ncell * nb * nb
b_sc = (8.0 + 2.0) * ncell * nb * nb + 8.0 * A.nnz
print("P%d x P%d stiffness block with a P1 coefficient: %d cells, %d rows, nnz %d, element tensor %.2f GB" % (deg, deg, ncell, A.nrow, A.nnz, 8e-9 * ncell * nb * nb))
pos = plan.row_positions()
el_rows = lambda: check(ctx.lib.mp_elem_reftensor_rows(ctx.h, C.byref(m), 3, 0, nb, nb, ptr(T0), ptr(coef), 1, 1.0, pos, ptr(emat)))
el_cell = lambda: check(ctx.lib.mp_elem_reftensor(ctx.h, C.byref(m), 3, 0, nb, nb, ptr(T0), ptr(coef), 1, 1.0, ptr(emat)))
if only != "rows":
    t = timed(el_cell)
    print("k_reftensor, cell-ordered output : %.3f ms  %.0f GB/s  frac %.2f of measured HBM peak" % (t, b_el / t / 1e6, b_el / t / 1e6 / peak))
    t = timed(lambda: plan.scatter_matrix(emat, A, ordered=False))
    print("k_scatter_matrix, cell-ordered   : %.3f ms  %.0f GB/s  frac %.2f" % (t, b_sc / t / 1e6, b_sc / t / 1e6 / peak))
t = timed(el_rows)
print("k_reftensor, row-ordered output  : %.3f ms  %.0f GB/s  frac %.2f of measured HBM peak" % (t, b_el / t / 1e6, b_el / t / 1e6 / peak))
t = timed(lambda: plan.scatter_matrix(emat, B, ordered=True))
print("k_scatter_matrix, row-ordered    : %.3f ms  %.0f GB/s  frac %.2f" % (t, b_sc / t / 1e6, b_sc / t / 1e6 / peak))
if only != "rows":
    d = (A.vals - B.vals).abs().max().item() / A.vals.abs().max().item()
    print("max relative difference between the two assemblies: %.1e" % d)
