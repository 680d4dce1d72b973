"""Time the c0-row assembly at the bench size: fused row-wise kernel against the element-kernel + scatter pipeline (CUDA events).
usage: python tools/asm_probe.py [n] [reps]   -- also the ncu target for k_p1_transport_rows (PROBE_ONLY=fused)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mupopp_b200.device import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
only = os.environ.get("PROBE_ONLY")
ctx = Context(0)
ap = argparse.Namespace(n=n, sides="periodic", rtol=1e-12, maxit=50000, c0_method="bicgstab", gmres_restart=30, p_precond="line", cheb_degree=3,
                        p_guess="previous", c0_precond="auto")
s = bench.build_problem(ctx, ap, 0, 1)
for _ in range(3):
    s.step()                                  # a developed state: non-trivial velocity and concentrations
for fused in ([True] if only == "fused" else ([False] if only == "pipeline" else [False, True])):
    s.fused_rows = fused
    for _ in range(2):
        s.assemble_c0_row()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s.assemble_c0_row()
        s.A.sync_sell()
    e1.record()
    torch.cuda.synchronize()
    print("c0 row assembly (%s): %.3f ms" % ("fused row-wise kernel + line factor" if fused else
                                              "element kernel + scatter + Dirichlet + Jacobi + line extract/factor + SELL fill", e0.elapsed_time(e1) / reps))
