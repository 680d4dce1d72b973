#!/usr/bin/env python
"""bench.py -- Darcy-ADR timesteps/s at 8M dofs (BASELINE.json metric) on 1..8 B200.

Workload (config.workload): CCS reactive 3-component transport (form a5, mupopp.py:1147-1262, physics of
/root/reference/src/run/CCS2D/CCS_3comp.py:29-39 lifted to 3-D), mode F (P1 pressure operator + cellwise velocity
recovery + P1 transport), BoxMesh [0,4]x[0,4]x[0,1] with n^3 hexes -> 6 n^3 tets, n = 200: 8 120 601 P1 dofs per field.
One "step" = one full timestep: c1,c2 mass solves -> c0 SUPG-ADR assemble+solve -> pressure solve -> velocity recovery.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 200]

N > 1 (torchrun, one rank per GPU): the SAME mesh is partitioned into z-slabs (strong scaling), NCCL halo
exchange + allreduce inside the Krylov loops.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Darcy-ADR timesteps/s at 8M dofs"
UNIT = "timesteps/s"
PHYS = dict(Pe=500.0, Da=0.1, phi=0.05, alpha=1e-8, dt=0.1, K=1.0, gam_rho=1.0)   # CCS_3comp.py:29-39
BOX = ((0.0, 0.0, 0.0), (4.0, 4.0, 1.0))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def source_term(X, ztop=1.0):
    """CCS_3comp.py:146-164 lifted to 3-D: sum_{ii<20} 0.1 |sin(ii pi x)| (1 - tanh((ztop - z)/0.01))."""
    f = np.zeros(X.shape[0])
    for ii in range(20):
        f += 0.1 * np.abs(np.sin(ii * np.pi * X[:, 0]))
    return f * (1.0 - np.tanh((ztop - X[:, 2]) / 0.01))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- reference arm
def oracle_step_time(n, rtol, steps, warm, sides="periodic"):
    """The oracle's mode-F timestep (NumPy assembly + SciPy cg/bicgstab with Jacobi) on an n^3 sample."""
    from oracle import fem, forms, modef
    m = fem.box_mesh(BOX[0], BOX[1], n, n, n)
    asm = forms.Assembler(m)
    prm = forms.ccs_three_component_params(Pe=PHYS["Pe"], Da=PHYS["Da"], phi=PHYS["phi"], alpha=PHYS["alpha"], dt=PHYS["dt"],
                                           K=PHYS["K"], zhat=(0.0, 0.0, 1.0), gam_rho=PHYS["gam_rho"])
    top = lambda x: x[2] > 1.0 - 1e-12
    P = None
    if sides == "periodic":
        P, dof = modef.periodic_prolongation(m, (0, 1), BOX[0], BOX[1])
        bcs = modef.make_bcs_periodic(m, (0, 1), dof, top, lambda x: np.array([0.0, 0.0, -0.5]), top, 0.1)
    elif sides == "noflow":
        bcs = modef.make_bcs(m, lambda x: x[2] > 1e-12, lambda x: np.array([0.0, 0.0, -0.5 if x[2] > 1.0 - 1e-12 else 0.0]), top, 0.1)
    else:
        bcs = modef.make_bcs(m, top, lambda x: np.array([0.0, 0.0, -0.5]), top, 0.1)
    f = source_term(m.coords)
    nd = m.nv
    if P is not None:
        nd = P.shape[1]
        first = np.full(nd, -1, dtype=np.int64)
        first[dof[::-1]] = np.arange(m.nv)[::-1]
        f = f[first]
    st = modef.StateF(u=np.zeros((m.nc, 3)), p=np.zeros(nd), c0=np.zeros(nd), c1=0.1 * np.ones(nd), c2=np.zeros(nd))
    times, iters = [], []
    for k in range(warm + steps):
        info = {}
        t0 = time.perf_counter()
        st = modef.step(asm, prm, st, f, bcs, solver="krylov", rtol=rtol, info=info, P=P)
        dt = time.perf_counter() - t0
        if k >= warm:
            times.append(dt)
            iters.append(dict(info))
        log("[oracle] n=%d step %d: %.2f s iters %s" % (n, k, dt, info))
    return nd, times, iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.ref_n
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    warm = min(args.warmup, 3)          # past the trivial first steps (the initial state solves in a handful of iterations)
    ndof, times, iters = oracle_step_time(n, args.rtol, args.steps, warm, args.sides)
    ms = 1e3 * float(np.mean(times))
    ndof_full = ndofs_field(args)
    # scale the sample to the full workload linearly in dofs (optimistic for the CPU: Krylov iteration counts grow with n)
    value = 1.0 / (np.mean(times) * ndof_full / ndof)
    sample = ("oracle mode-F timesteps %d..%d (NumPy assembly + SciPy cg/bicgstab+Jacobi, rtol %g, single-threaded: SciPy's sparse "
              "kernels do not thread) on a %d^3 sample (%d dofs/field), scaled to %d^3 linearly in dofs (iteration growth with n "
              "ignored: favours the CPU)" % (warm + 1, warm + args.steps, args.rtol, n, ndof, args.n))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms * ndof_full / ndof, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                            "host_cores_available": cores, "sample_ms_per_step": ms, "sample_iters": iters[-1] if iters else None},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


def ndofs_field(args):
    """P1 dofs per scalar field: (n+1)^3 vertices, n^2 (n+1) when the x/y faces are periodic."""
    n = args.n
    return n * n * (n + 1) if args.sides == "periodic" else (n + 1) ** 3


def workload_config(args, nranks):
    return {"workload": "CCS 3-component reactive transport (mupopp CCS.advection_diffusion_three_component), mode F, "
                        "BoxMesh [0,4]x[0,4]x[0,1] %d^3 x 6 tets, P1, %d dofs/field" % (args.n, ndofs_field(args)),
            "n": args.n, "dofs_per_field": ndofs_field(args), "cells": 6 * args.n ** 3,
            "physics": PHYS, "krylov_rtol": args.rtol, "c0_solver": args.c0_method + ("(%d)" % args.gmres_restart if args.c0_method == "gmres" else ""), "preconditioner": args.p_precond + ("(%d)" % args.cheb_degree if args.p_precond == "chebyshev" else ""),
            "c0_preconditioner": "line" if (args.p_precond == "line" and args.c0_precond == "auto" and args.c0_method == "bicgstab") else "jacobi", "side_walls": args.sides,
            "partition": "z-slabs x%d" % nranks if nranks > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (CSR operators 1.6 GB each at n=200 vs 126 MB L2)"}


# ----------------------------------------------------------------------------------------------- B200 arm
def build_problem(ctx, args, rank, nranks):
    import mupopp_b200 as mp
    from mupopp_b200.transport import TransportModel, TransportSolverF, make_boundary_data
    from mupopp_b200 import partition
    n = args.n
    model = TransportModel.ccs_three_component(Pe=PHYS["Pe"], Da=PHYS["Da"], phi=PHYS["phi"], alpha=PHYS["alpha"], dt=PHYS["dt"],
                                               K=PHYS["K"], zh=(0.0, 0.0, 1.0), gam_rho=PHYS["gam_rho"])
    top = lambda x: x[..., 2] > 1.0 - 1e-12
    if args.sides == "noflow":
        # impermeable side walls: u.n = 0 on x/y faces (Gamma_u = top + sides), p = 0 only on the bottom outflow face
        umark = lambda x: x[..., 2] > 1e-12
        uval = lambda x: np.stack([0 * x[..., 0], 0 * x[..., 0], np.where(x[..., 2] > 1.0 - 1e-12, -0.5, 0.0)], -1)
    else:
        umark, uval = top, (0.0, 0.0, -0.5)
    per = args.sides == "periodic"
    if nranks == 1:
        from mupopp_b200.mesh import box_periodic_map
        mesh = mp.BoxMesh(BOX[0], BOX[1], n, n, n)
        pmap = box_periodic_map(mesh, (0, 1)) if per else None
        bd = make_boundary_data(mesh, umark, uval, top, 0.1, periodic=pmap)
        fs = source_term(mesh.coords)
        solver = TransportSolverF(ctx, mesh, model, bd, f_source=pmap.restrict(fs) if per else fs, rtol=args.rtol, maxit=args.maxit,
                                  c0_method=args.c0_method, gmres_restart=args.gmres_restart, vert2dof=pmap.vert2dof if per else None,
                                  p_precond=args.p_precond, cheb_degree=args.cheb_degree, p_guess=args.p_guess, c0_precond=args.c0_precond)
        c1 = 0.1 * np.ones(solver.dm.ndof)
    else:
        part = partition.slab_partition(BOX, (n, n, n), rank, nranks, periodic_xy=per)
        bd = partition.slab_boundary_data(part, umark, uval, top, 0.1)
        halo = partition.make_halo(ctx, part)        # peer-memory ghost window when the context's windows are attached
        fs = source_term(part.mesh.coords)
        solver = TransportSolverF(ctx, part.mesh, model, bd, f_source=part.periodic.restrict(fs) if per else fs, rtol=args.rtol,
                                  maxit=args.maxit, c0_method=args.c0_method, gmres_restart=args.gmres_restart, halo=halo, n_owned=part.n_owned,
                                  vert2dof=part.vert2dof, p_precond=args.p_precond, cheb_degree=args.cheb_degree, p_guess=args.p_guess, c0_precond=args.c0_precond)
        c1 = 0.1 * np.ones(part.n_local)
    solver.set_state(c1=c1)
    return solver


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mupopp_b200.device import Context
    rank = int(os.environ.get("RANK", "0"))
    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mupopp_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    ctx = Context(local)
    if args.pcg_variant >= 0:
        ctx.set_option("pcg_variant", args.pcg_variant)
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(nranks, rank, ids[0])
        if args.transport == "peer":
            on = ctx.enable_peer()
            if rank == 0:
                log("[bench] peer-memory collectives (CUDA IPC over NVLink): %s" % ("on" if on else "unavailable (%s), using NCCL" % getattr(ctx, "peer_error", "?")))

    def barrier():
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    s = build_problem(ctx, args, rank, nranks)
    barrier()
    log("[rank %d] setup %.1f s: %d local vertices, %d cells, nnz %d" % (rank, time.perf_counter() - t0, s.dm.nvert, s.dm.ncell, s.L.nnz))

    def check_step(r, where):
        """A step only counts if every Krylov solve of it converged to the stated tolerance (no cheaper, less converged steps)."""
        for name, res in r.items():
            if res is not None and not res.converged:
                raise SystemExit("bench.py: %s solve did not converge in %s (%r); the metric is void" % (name, where, res))
        return {a: (b.niter if b else 0) for a, b in r.items()}

    # ---- warm-up: W untimed steps from the initial state.  The Krylov work per step depends on how far the fronts have moved
    #      (pressure iterations grow over the first tens of steps), so the timed window is FIXED: steps W+1 .. W+K of this series.
    for k in range(args.warmup):
        t1 = time.perf_counter()
        it = check_step(s.step(), "warm-up step %d" % k)
        torch.cuda.synchronize()
        log("[rank %d] warmup %d: %.3f s %s" % (rank, k, time.perf_counter() - t1, it))

    state_names = ["c0", "c1", "c2", "p", "p_old", "u"]
    saved = {k: getattr(s, k).clone() for k in state_names}
    saved_hist = s._p_hist

    def restore():
        for k in state_names:
            getattr(s, k).copy_(saved[k])
        s._p_hist = saved_hist

    def timed_region(body):
        """K steps bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks -> ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        out = [body(k) for k in range(args.steps)]
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
        if nranks > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    # ---- timed region 1: device-resident state (value).  Nothing else runs on this rank: no per-launch events, no host reads.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ms_total, results = timed_region(lambda k: s.step())
    launches = ctx.launch_count() - launches0
    iters = [check_step(r, "timed step %d" % k) for k, r in enumerate(results)]
    true_res = s.true_residuals()
    fallbacks = getattr(s, "fallbacks", 0)
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- timed region 2 (e2e): the SAME K steps from the SAME saved state, with HOST state buffers: every step copies its input
    #      state host->device from pinned memory and its result device->host.  On one GPU the step is the reference-facing plugin
    #      call itself -- `a, L = CCS.advection_diffusion_three_component(W, mesh, sol_0, dt, f_source=S, K=1.0); solve(a == L, sol, bc)`
    #      on a periodic `constrained_domain` space (mupopp_b200.mupopp / mupopp_b200.fenics); on N > 1 GPUs, where the facade has no
    #      partitioned spaces, it is TransportSolverF.step on the rank's slab.
    restore()
    api = None
    if nranks == 1 and args.sides == "periodic" and not args.no_api:
        t1 = time.perf_counter()
        api = ApiProblem(args, ctx, s, saved)
        log("[bench] drop-in API problem built in %.1f s (FunctionSpace with constrained_domain, DirichletBC, CCS forms)" % (time.perf_counter() - t1))
        dev_in, dev_out = api.device_buffers()
    else:
        names = ["c0", "c1", "c2", "p"]             # the cellwise velocity is derived state: recomputed from (p, c0) after the copy
        dev_in = dev_out = [getattr(s, k) for k in names]
    host_in = [torch.empty_like(t, device="cpu").pin_memory() for t in dev_in]
    host_out = [torch.empty_like(t, device="cpu").pin_memory() for t in dev_out]
    src = api.initial_state() if api is not None else dev_in
    for h, t in zip(host_in, src):
        h.copy_(t)
    torch.cuda.synchronize()
    bytes_in = sum(int(v.numel()) * 8 for v in host_in)
    bytes_out = sum(int(v.numel()) * 8 for v in host_out)
    hbuf = {"in": host_in, "out": host_out}

    def e2e_step(k):
        if api is None:
            d_in = d_out = [getattr(s, nm) for nm in names]                 # (the solver rotates its state tensors every step)
        else:
            d_in, d_out = dev_in, dev_out
        for t, h in zip(d_in, hbuf["in"]):
            t.copy_(h, non_blocking=True)                                    # H2D of this step's input state
        if api is None:
            s.c0_new.copy_(s.c0)
            s.recover_velocity()                                             # u^n = the Darcy velocity of the state just copied in
        r = api.step() if api is not None else s.step()
        if api is None:
            d_out = [getattr(s, nm) for nm in names]
        for h, t in zip(hbuf["out"], d_out):
            h.copy_(t, non_blocking=True)                                    # D2H of the step's result
        torch.cuda.current_stream().synchronize()
        hbuf["in"], hbuf["out"] = hbuf["out"], hbuf["in"]                     # the next step starts from the host copy
        return r

    ms_e2e, results2 = timed_region(e2e_step)
    iters2 = [check_step(r, "e2e step %d" % k) for k, r in enumerate(results2)]
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = 1e3 / (ms_e2e / args.steps)
    e2e_call = ("CCS.advection_diffusion_three_component(...) + solve(a == L, sol, bc, solver_parameters={'mupopp_b200_lagged_velocity': 'recover'}) on "
                "FunctionSpace(mesh, MixedElement([V, Q, Qc, Qc, Qc]), constrained_domain=pbc); host buffers = all 7 blocks [ux, uy, uz, p, c0, c1, c2] "
                "of the mixed Function in and out; the cellwise (DG0) velocity is derived state, recomputed on the device from (p, c0)") \
        if api is not None else "TransportSolverF.step() on the rank's z-slab; host buffers = [c0, c1, c2, p] incl. ghosts, velocity recomputed from (p, c0)"
    del api

    # ---- pass 3 (untimed for the metric): the dominant kernel's launches timed one by one with CUDA events on the launch stream,
    #      over the first steps of the same window (roofline.achieved); kept out of regions 1 and 2 so that neither is perturbed.
    restore()
    nroof = min(args.steps, args.roof_steps)
    ctx.timing(True, 40000)
    ctx.timing_read()
    s.phase_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(nroof):
        s.step()
    ev1.record()
    barrier()
    n_spmv, spmv_ms = ctx.timing_read()
    ctx.timing(False)
    ms_roof = ev0.elapsed_time(ev1)
    phases = {k: v / nroof for k, v in s.phase_times().items()}
    s.phase_timing(False)
    if nranks > 1:                                      # max over ranks, phase by phase
        pt = torch.tensor([phases[k] for k in sorted(phases)], dtype=torch.float64, device="cuda")
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        phases = dict(zip(sorted(phases), pt.tolist()))

    if rank == 0:
        peak, peak_src = measured_peak()
        N, nnz = s.L.nrow, s.L.nnz
        # ALGORITHMIC bytes of one SpMV launch.  SURVEY.md 8(d) writes them for CSR with int32 columns: 12 nnz + 4 (N+1) + 16 N.  The
        # production kernel streams the SELL-32 mirror with slice-uniform column offsets: 8 bytes per stored value,
        # ONE int32 per (slice, column position) and explicit int32 columns only where the 32 rows of a slice disagree -- so its own
        # compulsory traffic is smaller; `achieved` / `frac` use the bytes the kernel must actually stream, the survey figure is kept beside.
        survey_bytes = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
        E, X = s.L.sell_entries, s.L.sell_explicit
        compressed = X >= 0
        nslice = (N + 31) // 32
        spmv_bytes = (8.0 * E + 4.0 * (E / 32) + 128.0 * X + 4.0 * (nslice + 1) + 16.0 * N) if compressed else \
                     (12.0 * E + 4.0 * (nslice + 1) + 16.0 * N)
        avg_ms = spmv_ms / max(n_spmv, 1)
        achieved = spmv_bytes / (avg_ms * 1e-3) / 1e9 if n_spmv else None
        traffic, traffic_src = ncu_traffic(args, nranks)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": workload_config(args, nranks),
               "timed_window": "steps %d..%d of the time series started from c1=0.1, c0=c2=p=u=0 (both regions start from the same saved state)"
                               % (args.warmup + 1, args.warmup + args.steps),
               "krylov_iterations_per_step": iters, "krylov_iterations_per_step_e2e": iters2,
               "all_solves_converged": True, "bicgstab_to_gmres_fallbacks": int(fallbacks),
               "true_relative_residuals_last_step": true_res,
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
                       "ms_per_step": ms_e2e / args.steps, "overhead_vs_resident_ms_per_step": ms_e2e / args.steps - ms_step, "call": e2e_call},
               "roofline": {"bound": "hbm", "kernel": "%s (SELL-32 SpMV inside PCG/BiCGStab, per-rank operator)"
                                                      % ("k_spmv_sell_peer" if (nranks > 1 and ctx.peer_enabled()) else "k_spmv_sell"),
                            "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                            "frac": (achieved / peak) if achieved else None,
                            "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": spmv_bytes,
                            "bytes_model": ("SELL-32 with slice-uniform column offsets: 8 B/value + 4 B per (slice, position) + 128 B per explicit "
                                            "(slice, position) [%d of %d] + slice pointers + x read once + y written once" % (X, E // 32))
                                           if compressed else "SELL-32 with int32 columns: 12 B per stored entry + slice pointers + 16 N",
                            "survey_formula_bytes_per_launch": survey_bytes,
                            "achieved_by_survey_formula": (survey_bytes / (avg_ms * 1e-3) / 1e9) if n_spmv else None,
                            "frac_by_survey_formula": (survey_bytes / (avg_ms * 1e-3) / 1e9 / peak) if n_spmv else None,
                            "launches_timed": n_spmv, "avg_launch_ms": avg_ms,
                            "timed_in": "separate pass of %d steps from the same saved state (%.1f ms/step with per-launch events)"
                                        % (nroof, ms_roof / nroof),
                            "share_of_step": spmv_ms / ms_roof},
               "phase_ms_per_step": dict(phases, measured_in="the same untimed pass as the roofline (CUDA events at the phase boundaries of "
                                                              "TransportSolverF.step, max over ranks)")}
        if nranks > 1:
            out["transport"] = "peer memory (CUDA IPC over NVLink): in-kernel allreduce + halo push fused into SpMV" if ctx.peer_enabled() \
                else "NCCL (ncclSend/ncclRecv halo + ncclAllReduce per inner product)"
        if not args.no_cpu and nranks == 1:
            try:
                warm, timed = 3, 2
                ndof, times, oit = oracle_step_time(args.ref_n, args.rtol, timed, warm, args.sides)
                cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
                v = 1.0 / (np.mean(times) * ndofs_field(args) / ndof)
                out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores_available": cores,
                                       "sample": "oracle mode-F timesteps %d..%d (after %d warm steps; NumPy assembly + SciPy cg/bicgstab+Jacobi, "
                                                 "same rtol, single-threaded) on a %d^3 sample (%d dofs/field, %.1f s/step), scaled to %d^3 "
                                                 "linearly in dofs (iteration growth with n ignored: favours the CPU)"
                                                 % (warm + 1, warm + timed, warm, args.ref_n, ndof, float(np.mean(times)), args.n),
                                       "sample_iters": oit[-1] if oit else None}
            except Exception as e:  # the baseline is reporting only; never let it break the GPU line
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(out), flush=True)
    if nranks > 1:
        dist.destroy_process_group()
    return 0


class ApiProblem:
    """The bench workload expressed with the reference-facing API, exactly as a run script would write it
    (/root/reference/src/run/CCS2D/CCS_3comp.py:127-143,206-263 lifted to the 3-D box)."""

    def __init__(self, args, ctx, solver, saved):
        import torch
        import mupopp_b200.fenics as fx
        from mupopp_b200 import mupopp as mp_api
        fx._CTX = ctx                                       # one GPU context / stream for both problems
        n = args.n
        (x0, y0, z0), (x1, y1, z1) = BOX
        mesh = solver.dm.host                               # the BoxMesh the resident-state problem was built on

        class PeriodicBoundary(fx.SubDomain):
            def inside(self, x, on_boundary):
                return bool((fx.near(x[0], x0) or fx.near(x[1], y0)) and not (fx.near(x[0], x1) or fx.near(x[1], y1)) and on_boundary)

            def map(self, x, y):
                y[0], y[1], y[2] = x[0], x[1], x[2]
                if fx.near(x[0], x1):
                    y[0] = x[0] - (x1 - x0)
                if fx.near(x[1], y1):
                    y[1] = x[1] - (y1 - y0)

        class SourceTerm(fx.Expression):
            def __init__(self, top):
                self.top = top

            def eval(self, values, x):
                g1 = x[0] * 0.0
                for ii in range(0, 20):
                    g1 += 0.1 * np.abs(np.sin(ii * x[0] * np.pi))
                values[0] = (1.0 - np.tanh((self.top - x[2]) / 0.01)) * g1

        cell = mesh.ufl_cell()
        V = fx.VectorElement("Lagrange", cell, 2)
        Q = fx.FiniteElement("Lagrange", cell, 1)
        Qc = fx.FiniteElement("Lagrange", cell, 1)
        self.W = fx.FunctionSpace(mesh, fx.MixedElement([V, Q, Qc, Qc, Qc]), constrained_domain=PeriodicBoundary())
        top = lambda x: x[2] > z1 - 1e-12
        self.bc = [fx.DirichletBC(self.W.sub(0), fx.Constant((0.0, 0.0, -0.5)), top), fx.DirichletBC(self.W.sub(2), fx.Constant(0.1), top)]
        self.ccs = mp_api.CCS(Pe=PHYS["Pe"], Da=PHYS["Da"], phi=PHYS["phi"], alpha=PHYS["alpha"])
        self.S = SourceTerm(z1)
        self.sol_0, self.sol = fx.Function(self.W), fx.Function(self.W)
        self.mesh, self.fx, self.args = mesh, fx, args
        ncell = mesh.num_cells()
        self.sol_0.cell_velocity = torch.zeros(3 * ncell, dtype=torch.float64, device=ctx.device)
        self.sol.cell_velocity = torch.zeros(3 * ncell, dtype=torch.float64, device=ctx.device)
        nd = saved["p"].numel()
        self._init = [torch.zeros(nd, dtype=torch.float64, device=ctx.device) for _ in range(3)] + [saved["p"], saved["c0"], saved["c1"], saved["c2"]]
        self.zh = fx.Constant((0.0, 0.0, 1.0))
        # one untimed call: solve() builds its GPU operators (scatter plan, constant matrices, boundary data) on first use
        for t, src in zip(self.device_buffers()[0], self._init):
            t.copy_(src)
        self.step()

    def device_buffers(self):
        """Input state of a step = every block [ux, uy, uz, p, c0, c1, c2] of sol_0; result = every block of sol (the nodal velocity blocks
        are what the scripts write to file).  The cellwise velocity the transport row uses is recomputed from (p, c0) by the solve."""
        return (list(self.sol_0.blocks), list(self.sol.blocks))

    def initial_state(self):
        return self._init

    def step(self):
        a, L = self.ccs.advection_diffusion_three_component(self.W, self.mesh, self.sol_0, PHYS["dt"], f_source=self.S, K=PHYS["K"], zh=self.zh,
                                                            gam_rho=PHYS["gam_rho"])
        return self.fx.solve(a == L, self.sol, self.bc, solver_parameters={"relative_tolerance": self.args.rtol, "preconditioner": self.args.p_precond,
                                                                           "mupopp_b200_lagged_velocity": "recover"})


def ncu_traffic(args, nranks):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, read from the committed ncu capture of THIS
    operator (profiles/r2_ncu_spmv.json, written by tools/ncu_summary.py from the .ncu-rep of `bench.py --steps 1 --warmup 1`)."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_spmv.json")
    try:
        d = json.load(open(p))
        key = "n%d_%s_x%d" % (args.n, args.sides, nranks)
        if key in d:
            return float(d[key]["dram_bytes_per_launch"]), "profiles/r2_ncu_spmv.json[%s] (ncu --set full)" % key
    except Exception:
        pass
    return None, "no committed ncu capture for this operator"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ccs200", choices=["ccs200", "darcy128", "compaction", "stokes"],
                    help="ccs200 = the contract line (BASELINE configs[2]); the others are the secondary lines of bench_workloads.py")
    ap.add_argument("--size", dest="n", type=int, default=200, help="hexes per direction (200 -> 8 120 601 dofs/field, the metric's size)")
    ap.add_argument("--ref-n", type=int, default=40, help="sample size of the CPU oracle leg")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer-memory windows over NVLink (in-kernel allreduce, halo push fused into SpMV) or plain NCCL")
    ap.add_argument("--pcg-variant", type=int, default=-1, help="-1 auto (single-reduction CG on N > 1 GPUs, classic on one), 0 classic, 1 single-reduction")
    ap.add_argument("--p-guess", default="previous", choices=["extrapolate", "previous"],
                    help="initial iterate of the pressure PCG: p^n (default) or 2 p^n - p^(n-1) (measured: no fewer iterations, "
                         "profiles/r2_pressure_guess.md -- the tolerance is relative to ||b||, and the front moves non-smoothly)")
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--maxit", type=int, default=50000)
    ap.add_argument("--c0-method", default="bicgstab", choices=["gmres", "bicgstab"],
                    help="transport-row solver; bicgstab falls back to GMRES if it breaks down (TransportSolverF._solve_c0)")
    ap.add_argument("--gmres-restart", type=int, default=30)
    ap.add_argument("--p-precond", default="line", choices=["jacobi", "chebyshev", "line"],
                    help="pressure PCG preconditioner: point Jacobi, Chebyshev-Jacobi polynomial, or lattice-line block Jacobi (tridiagonal z-lines)")
    ap.add_argument("--c0-precond", default="auto", choices=["auto", "jacobi"], help="transport-row BiCGStab preconditioner: auto = the pressure's lattice lines when --p-precond line, else Jacobi")
    ap.add_argument("--cheb-degree", type=int, default=3)
    ap.add_argument("--roof-steps", type=int, default=3, help="steps of the untimed third pass (per-launch SpMV events, phase breakdown)")
    ap.add_argument("--skip-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-api", dest="no_api", action="store_true", help="e2e through TransportSolverF.step instead of the mupopp/fenics plugin call")
    ap.add_argument("--sides", default="periodic", choices=["natural", "noflow", "periodic"],
                    help="x/y faces: periodic constrained_domain (default; what the reference's CCS scripts use, CCS_3comp.py:131-143,212), "
                         "impermeable u.n=0, or natural p=0 (CCS_3D_side_source.py)")
    ap.add_argument("--hard-timeout", type=int, default=900, help="SIGALRM kills the process after this many seconds (never hang a box)")
    args = ap.parse_args()
    if args.hard_timeout > 0:
        import signal
        signal.alarm(args.hard_timeout)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "ccs200":
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; mupopp_b200 has no CPU path")
        import bench_workloads
        return bench_workloads.RUNNERS[args.workload](args, sys.modules[__name__])
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
