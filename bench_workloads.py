"""Secondary bench lines: BASELINE.json configs[1], [3], [4] (the driver's contract line is bench.py's default workload, configs[2]).

    python bench.py --workload darcy128    DarcyAdvection two-component form on BoxMesh 128^3 (LVL_RII3D.py physics), mode F, periodic sides
    python bench.py --workload compaction  Taylor-Hood compaction step (time_test1.py): momentum assembly + MINRES + porosity row
    python bench.py --workload stokes      stokes_ADR_precipitation step on a structured stand-in of the pore geometry (P3 velocity)

Each prints ONE JSON line with the same keys as the main line; `roofline` describes the kernel that dominates that workload, timed with
CUDA events on the launch stream in a separate pass.  Single GPU (the mode-R block systems are not partitioned).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_steps(torch, step, warmup, steps):
    for k in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = _events(torch)
    e0.record()
    out = [step() for _ in range(steps)]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def _time_call(torch, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _base(metric, value, ms, args, config, **extra):
    out = {"metric": metric, "value": value, "unit": "timesteps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config}
    out.update(extra)
    return out


# ----------------------------------------------------------------------------------------------- config 2
def run_darcy128(args, bench):
    """DarcyAdvection.advection_diffusion_two_component (mupopp.py:751-849) with the physics and data of run/LVL_3D/LVL_RII3D.py:28-37,
    102-108,209-211,235 on BoxMesh [0,4]x[0,4]x[0,1] n^3 (n = 128), periodic x/y (the script's constrained_domain), mode F."""
    import torch
    import mupopp_b200 as mp
    from mupopp_b200.device import Context
    from mupopp_b200.mesh import box_periodic_map
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    n = args.n if args.n != 200 else 128
    ctx = Context(0)
    t0 = time.perf_counter()
    mesh = mp.BoxMesh((0.0, 0.0, 0.0), (4.0, 4.0, 1.0), n, n, n)
    pmap = box_periodic_map(mesh, (0, 1))
    model = TransportModel.darcy_two_component(Pe=1.0e3, Da=10.0, phi=0.01, alpha=0.01, dt=0.1, K=0.1, zh=(0.0, 0.0, 1.0))
    bottom = lambda x: x[..., 2] < 1e-12
    bd = make_boundary_data(mesh, bottom, (0.0, 0.0, 0.1), None, 0.0, periodic=pmap)
    X = mesh.coords[pmap.dof2vert]
    f = np.zeros(X.shape[0])
    for ii in range(20):
        f += 0.01 * np.abs(np.sin(ii * X[:, 0] * np.pi)) * np.abs(np.sin(ii * X[:, 1] * np.pi))
    f *= 1.0 - np.tanh(X[:, 2] / 0.01)
    s = TransportSolverF(ctx, mesh, model, bd, f_source=f, rtol=args.rtol, maxit=args.maxit, c0_method=args.c0_method,
                         gmres_restart=args.gmres_restart, vert2dof=pmap.vert2dof)
    s.set_state(c1=0.01 * np.ones(pmap.ndof))
    log("[darcy128] setup %.1f s, %d dofs/field, %d cells" % (time.perf_counter() - t0, pmap.ndof, mesh.num_cells()))

    def step():
        r = s.step()
        for k, v in r.items():
            if v is not None and not v.converged:
                raise SystemExit("bench: %s solve did not converge (%r)" % (k, v))
        return {k: (v.niter if v else 0) for k, v in r.items()}

    ms, its = _time_steps(torch, step, args.warmup, args.steps)
    ctx.timing(True, 40000)
    ctx.timing_read()
    e0, e1 = _events(torch)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    nsp, spms = ctx.timing_read()
    ctx.timing(False)
    peak, psrc = bench.measured_peak()
    N, E, Xp = s.L.nrow, s.L.sell_entries, s.L.sell_explicit
    bytes_ = 8.0 * E + 4.0 * (E / 32) + 128.0 * max(Xp, 0) + 4.0 * ((N + 31) // 32 + 1) + 16.0 * N
    avg = spms / max(nsp, 1)
    cfg = {"workload": "DarcyAdvection two-component (mupopp DarcyAdvection.advection_diffusion_two_component), mode F, BoxMesh [0,4]x[0,4]x[0,1] "
                       "%d^3 x 6 tets, P1, periodic x/y, %d dofs/field (BASELINE configs[1]; physics of run/LVL_3D/LVL_RII3D.py)" % (n, pmap.ndof),
           "n": n, "dofs_per_field": pmap.ndof, "krylov_rtol": args.rtol, "c0_solver": args.c0_method}
    out = _base("DarcyAdvection timesteps/s at 128^3 (config 2)", 1e3 / ms, ms, args, cfg, krylov_iterations_per_step=its,
                gpu_launches=None, e2e=None,
                roofline={"bound": "hbm", "kernel": "k_spmv_sell", "achieved": bytes_ / (avg * 1e-3) / 1e9, "peak": peak, "peak_source": psrc, "unit": "GB/s",
                          "frac": bytes_ / (avg * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": bytes_,
                          "launches_timed": nsp, "avg_launch_ms": avg, "share_of_step": spms / e0.elapsed_time(e1)})
    print(json.dumps(out), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- config 4
def run_compaction(args, bench):
    """One compaction step of benchmark/time_marching/time_test1.py:213-262: momentum_conservation + momentum_solver (assemble_system twice +
    MINRES, mupopp.py:306-429) followed by the porosity row mass_conservation + mass_solver (mupopp.py:217-304) on
    BoxMesh (-2,-2,0)-(2,2,8) with n x n x 2n hexes, Taylor-Hood [P2^3, P1, P1]; parameters of time_test1.cfg."""
    import torch
    import mupopp_b200 as mp
    from mupopp_b200.compaction import BlockVector, CompactionMomentumR, PorosityR
    from mupopp_b200.device import Context
    from mupopp_b200.moder import Spaces
    n = args.n if args.n != 200 else 16
    ctx = Context(0)
    t0 = time.perf_counter()
    mesh = mp.BoxMesh((-2.0, -2.0, 0.0), (2.0, 2.0, 8.0), n, n, 2 * n)
    sp = Spaces(ctx, mesh)
    X2 = sp.dof_coords(2)
    walls = (np.abs(np.abs(X2[:, 0]) - 2.0) < 1e-12) | (np.abs(np.abs(X2[:, 1]) - 2.0) < 1e-12)          # no-slip on the four vertical walls
    isbc = [walls.astype(np.uint8)] * 3
    g = [np.zeros(sp.n2)] * 3
    mom = CompactionMomentumR(ctx, sp, 1e-4, 0.2, 1e6, 10.0, 1.0, (0.0, 0.0, 1.0), isbc, g)
    por = PorosityR(ctx, sp, 1e-4)
    Xv = mesh.coords
    phi = ctx.to_device(0.05 * np.exp(-Xv[:, 0] ** 2 - Xv[:, 1] ** 2 - (Xv[:, 2] - 0.5) ** 2) / 0.5 + 0.1)
    gam, buoy = ctx.zeros(sp.n1), torch.ones(sp.n1, dtype=torch.float64, device=ctx.device)
    x = BlockVector(ctx, 3, sp.n2, sp.n1)
    phi1 = ctx.zeros(sp.n1)
    log("[compaction] setup %.1f s: %d cells, P2 dofs/component %d, P1 dofs %d, system size %d" % (
        time.perf_counter() - t0, mesh.num_cells(), sp.n2, sp.n1, 3 * sp.n2 + 2 * sp.n1))
    state = {"phi": phi}

    def step():
        mom.assemble(state["phi"], gam, buoy)
        its = mom.solve(x, rtol=1e-6, maxit=3000)                       # the reference's tolerance (mupopp.py:410-413), warm start from the last step
        if not mom.last_result.converged:
            raise SystemExit("bench: compaction MINRES did not converge (%r)" % (mom.last_result,))
        r = por.solve(state["phi"], x.u, 0.1, gam, phi1, rtol=1e-6, maxit=200)
        if not r.converged:
            raise SystemExit("bench: porosity solve did not converge (%r)" % (r,))
        state["phi"], = (phi1.clone(),)
        return {"minres": its, "porosity": r.niter}

    ms, its = _time_steps(torch, step, args.warmup, args.steps)
    # the dominant assembly kernel of this workload: the P2xP2 reference-tensor element kernel + its scatter (9 velocity blocks per step)
    import ctypes as C
    from mupopp_b200._lib import check, ptr
    m = sp.dm.c
    pos22 = mom.p22.row_positions()
    t_ref = _time_call(torch, lambda: check(ctx.lib.mp_elem_reftensor_rows(ctx.h, C.byref(m), 4, 0, sp.nb2, sp.nb2, ptr(mom.T_stiff22), ptr(mom.omp), 1,
                                                                          0.25, pos22, ptr(mom.emat2))))
    t_sc = _time_call(torch, lambda: mom.p22.scatter_matrix(mom.emat2, mom.Auu[0][0], ordered=True))
    t_asm = _time_call(torch, lambda: mom.assemble(state["phi"], gam, buoy), reps=2)
    ncell = mesh.num_cells()
    b_ref = 16.0 * ncell + 24.0 * mesh.num_vertices() + 8.0 * sp.n1 + 8.0 * ncell * sp.nb2 * sp.nb2
    b_sc = (8.0 + 2.0) * ncell * sp.nb2 * sp.nb2 + 8.0 * mom.Auu[0][0].nnz
    peak, psrc = bench.measured_peak()
    cfg = {"workload": "compaction Taylor-Hood step (compaction.momentum_conservation/momentum_solver + mass_conservation/mass_solver), BoxMesh "
                       "(-2,-2,0)-(2,2,8) %dx%dx%d x 6 tets, [P2^3, P1, P1] = %d unknowns (BASELINE configs[3]; benchmark/time_marching/time_test1.py)"
                       % (n, n, 2 * n, 3 * sp.n2 + 2 * sp.n1), "n": n, "minres_rtol": 1e-6,
           "preconditioner": "block Chebyshev-Jacobi of the reference's form b (degrees 8/8/4)"}
    out = _base("compaction timesteps/s (config 4)", 1e3 / ms, ms, args, cfg, krylov_iterations_per_step=its, gpu_launches=None, e2e=None,
                assembly_ms_per_step=t_asm,
                roofline={"bound": "hbm", "kernel": "k_reftensor<3> (P2xP2 symgrad block with a P1 coefficient: 36 reference tensors staged in shared memory, 64-cell tiles, row-ordered output; launch-bound at this mesh size -- tools/reftensor_probe.py measures it at 768 k cells)",
                          "achieved": b_ref / (t_ref * 1e-3) / 1e9, "peak": peak, "peak_source": psrc, "unit": "GB/s",
                          "frac": b_ref / (t_ref * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": b_ref, "avg_launch_ms": t_ref,
                          "second_kernel": {"kernel": "k_scatter_matrix<ordered> (P2xP2, streaming)", "achieved": b_sc / (t_sc * 1e-3) / 1e9,
                                            "frac": b_sc / (t_sc * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": b_sc, "avg_launch_ms": t_sc}})
    print(json.dumps(out), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- config 5
def run_stokes(args, bench):
    """StokesAdvection.stokes_ADR_precipitation step (mupopp.py:997-1099; run/stokes_ADR_3D/master_advective_stokes.py:94-130,208-218:
    Da = Pe = 0.1, dt = 100, An(0) = 0.1, inflow u = (1,0,0) and c0 = 0.1 on x = 0, no-slip walls) on a structured box standing in for the
    voxelised pore geometry (the shipped mesh has ~1.6 k vertices; BASELINE's 512^3 is synthetic): [P3^3, P1 x 4]."""
    import torch
    import mupopp_b200 as mp
    from mupopp_b200.device import Context
    from mupopp_b200.moder import Spaces
    from mupopp_b200.stokes import StokesAdrR
    n = args.n if args.n != 200 else 12
    ctx = Context(0)
    t0 = time.perf_counter()
    mesh = mp.BoxMesh((0.0, 0.0, 0.0), (2.0, 1.0, 1.0), 2 * n, n, n)
    sp0 = Spaces(ctx, mesh)
    X3 = sp0.dof_coords(3)
    n3 = X3.shape[0]
    inflow = X3[:, 0] < 1e-12
    wall = (((X3[:, 1:] < 1e-12) | (X3[:, 1:] > 1 - 1e-12)).any(1)) & ~inflow
    isbc = (inflow | wall).astype(np.uint8)
    g = [np.where(inflow, 1.0, 0.0), np.zeros(n3), np.zeros(n3)]
    X1 = mesh.coords
    cin = X1[:, 0] < 1e-12
    s = StokesAdrR(ctx, mesh, 0.1, 0.1, 100.0, 2, [isbc] * 3, g, cin.astype(np.uint8), np.where(cin, 0.1, 0.0), vdeg=3, rtol=args.rtol)
    s.c1.fill_(0.1)
    log("[stokes] setup %.1f s: %d cells, P3 dofs/component %d, P1 dofs %d" % (time.perf_counter() - t0, mesh.num_cells(), n3, X1.shape[0]))

    def step():
        r = s.step()
        if not r["c0"].converged or not s.stokes.last_result.converged:
            raise SystemExit("bench: Stokes-ADR solve did not converge (%r, %r)" % (r["c0"], s.stokes.last_result))
        return {"c0": r["c0"].niter, "stokes_minres": r["stokes_its"]}

    ms, its = _time_steps(torch, step, args.warmup, args.steps)
    cfg = {"workload": "stokes_ADR_precipitation step (mupopp StokesAdvection), BoxMesh [0,2]x[0,1]x[0,1] %dx%dx%d x 6 tets, [P3^3, P1 x 4] = %d unknowns "
                       "(BASELINE configs[4]; run/stokes_ADR_3D/master_advective_stokes.py on a structured stand-in of the pore geometry)"
                       % (2 * n, n, n, 3 * n3 + 4 * X1.shape[0]), "n": n, "krylov_rtol": args.rtol,
           "preconditioner": "block Chebyshev-Jacobi diag(L, pressure mass) (degrees 8/4)"}
    out = _base("Stokes-ADR timesteps/s (config 5)", 1e3 / ms, ms, args, cfg, krylov_iterations_per_step=its, gpu_launches=None, e2e=None, roofline=None)
    print(json.dumps(out), flush=True)
    return 0


RUNNERS = {"darcy128": run_darcy128, "compaction": run_compaction, "stokes": run_stokes}
