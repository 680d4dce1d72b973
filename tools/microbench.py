"""Kernel micro-benchmarks on one B200: SpMV variants, BLAS-1, assembly stages (GB/s vs algorithmic bytes)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mupopp_b200 as mp
from mupopp_b200._lib import check, ptr
from mupopp_b200.device import Context, DeviceMesh, Plan, transport_params


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    args = ap.parse_args()
    n = args.n
    ctx = Context(0)
    t0 = time.time()
    mesh = mp.BoxMesh((0, 0, 0), (4, 4, 1), n, n, n)
    dm = DeviceMesh(ctx, mesh)
    print("mesh %.1fs" % (time.time() - t0), flush=True)
    torch.cuda.synchronize(); t0 = time.time()
    plan = Plan(ctx, dm.cells, dm.cells, dm.nvert, dm.nvert)
    torch.cuda.synchronize()
    print("plan %.2fs nnz %d maxrow %d" % (time.time() - t0, plan.nnz, plan.max_row_nnz), flush=True)
    N, nnz, Cn = dm.nvert, plan.nnz, dm.ncell
    emat = ctx.empty(Cn * 16)
    evec = ctx.empty(Cn * 4)
    A = plan.new_matrix()
    prm = transport_params(500, 0.1, 0.05, 1e-8, 0.1, 1.0, 1.0, 1.0, (62 / 698., 278 / 698., 100 / 698.), (0, 0, 1.0), 1.0, "c0c1", 3, 1)
    x = torch.rand(N, dtype=torch.float64, device=ctx.device)
    y = torch.empty_like(x)
    z = torch.rand(N, dtype=torch.float64, device=ctx.device)
    u = torch.rand(Cn * 3, dtype=torch.float64, device=ctx.device)
    rows = []

    def rep(name, ms, bytes_):
        rows.append((name, ms, bytes_ / ms / 1e6))
        print("%-34s %8.3f ms  %8.1f GB/s (algorithmic %.1f MB)" % (name, ms, bytes_ / ms / 1e6, bytes_ / 1e6), flush=True)

    ms = timeit(lambda: check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 20.0, ptr(emat))))
    rep("elem_p1_stiffness", ms, 16.0 * Cn + 24.0 * N + 128.0 * Cn)
    ms = timeit(lambda: plan.scatter_matrix(emat, A))
    rep("scatter_matrix", ms, 128.0 * Cn + 8.0 * nnz + 4.0 * 4 * Cn + 2.0 * 16 * Cn)
    ms = timeit(lambda: check(ctx.lib.mp_elem_p1_transport(ctx.h, C.byref(dm.c), C.byref(prm), ptr(u), ptr(x), ptr(z), ptr(x), ptr(z), ptr(x), ptr(z), ptr(emat), ptr(evec))))
    rep("elem_p1_transport (mat+rhs)", ms, 16.0 * Cn + 24.0 * N + 24.0 * Cn + 6 * 8.0 * N + 160.0 * Cn)
    ms = timeit(lambda: plan.scatter_vector(evec, y))
    rep("scatter_vector", ms, 32.0 * Cn + 8.0 * N + 16.0 * Cn)
    # row-ordered element storage (what TransportSolverF uses): element kernels write in row order, the scatter streams
    dm.use_row_ordered_storage(plan)
    ms = timeit(lambda: check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 20.0, ptr(emat))))
    rep("elem_p1_stiffness (row-ordered)", ms, 16.0 * Cn + 24.0 * N + 128.0 * Cn + 16.0 * Cn)
    ms = timeit(lambda: plan.scatter_matrix(emat, A))
    rep("scatter_matrix (row-ordered)", ms, 128.0 * Cn + 8.0 * nnz + 2.0 * 16 * Cn)
    ms = timeit(lambda: check(ctx.lib.mp_elem_p1_transport(ctx.h, C.byref(dm.c), C.byref(prm), ptr(u), ptr(x), ptr(z), ptr(x), ptr(z), ptr(x), ptr(z), ptr(emat), ptr(evec))))
    rep("elem_p1_transport (row-ordered)", ms, 16.0 * Cn + 24.0 * N + 24.0 * Cn + 6 * 8.0 * N + 160.0 * Cn + 16.0 * Cn)
    ms = timeit(lambda: plan.scatter_vector(evec, y))
    rep("scatter_vector (row-ordered)", ms, 32.0 * Cn + 8.0 * N)
    check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(dm.c), None, 1.0, 20.0, ptr(emat)))
    plan.scatter_matrix(emat, A)
    spmv_bytes = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
    for var, name in ((2, "spmv CSR sub-warp/row"), (1, "spmv CSR CTA-stream"), (3, "spmv CSR warp-stream"), (0, "spmv SELL-32")):
        ctx.set_option("spmv_variant", var)
        ms = timeit(lambda: A.spmv(x, y))
        rep(name, ms, spmv_bytes)
    ms = timeit(lambda: ctx.axpby(2.0, x, 1.0, y))
    rep("axpby", ms, 24.0 * N)
    ms = timeit(lambda: y.copy_(x))
    rep("torch copy (peak probe)", ms, 16.0 * N)
    big = torch.empty(1 << 28, dtype=torch.float64, device=ctx.device)
    big2 = torch.empty_like(big)
    ms = timeit(lambda: big2.copy_(big), reps=5)
    rep("torch copy 2 GiB (peak probe)", ms, 16.0 * (1 << 28))
    del big, big2
    # whole PCG solve on the Dirichlet-eliminated Laplacian
    isbc = torch.zeros(N, dtype=torch.uint8, device=ctx.device)
    isbc[:(n + 1) ** 2] = 1
    A.apply_dirichlet(None, isbc, torch.zeros(N, dtype=torch.float64, device=ctx.device))
    b = torch.rand(N, dtype=torch.float64, device=ctx.device)
    b[:(n + 1) ** 2] = 0
    xs = torch.zeros(N, dtype=torch.float64, device=ctx.device)
    torch.cuda.synchronize(); t0 = time.time()
    r = A.solve("pcg", b, xs, rtol=1e-6, maxit=400)
    torch.cuda.synchronize(); dt = time.time() - t0
    it_bytes = spmv_bytes + 8.0 * N + 56.0 * N + 32.0 * N
    print("pcg: %s  %.3f ms/iter  %.1f GB/s algorithmic per iteration" % (r, 1e3 * dt / max(r.niter, 1), it_bytes * r.niter / dt / 1e9), flush=True)
    for meth, kw in (("pcg", {}), ("pcg_chebyshev", dict(cheb_degree=2)), ("pcg_chebyshev", dict(cheb_degree=3)), ("pcg_chebyshev", dict(cheb_degree=4))):
        xs.zero_()
        torch.cuda.synchronize(); t0 = time.time()
        r = A.solve(meth, b, xs, rtol=1e-8, maxit=6000, **kw)
        torch.cuda.synchronize(); dt = time.time() - t0
        print("%s %s to rtol 1e-8: %s  total %.1f ms" % (meth, kw, r, 1e3 * dt), flush=True)
    xs.zero_()
    torch.cuda.synchronize(); t0 = time.time()
    r = A.solve("gmres", b, xs, rtol=1e-6, maxit=200)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("gmres(50): %s  %.3f ms/iter" % (r, 1e3 * dt / max(r.niter, 1)), flush=True)
    xs.zero_()
    torch.cuda.synchronize(); t0 = time.time()
    r = A.solve("bicgstab", b, xs, rtol=1e-6, maxit=200)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("bicgstab: %s  %.3f ms/iter" % (r, 1e3 * dt / max(r.niter, 1)), flush=True)


if __name__ == "__main__":
    main()
