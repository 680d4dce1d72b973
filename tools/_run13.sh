set -x
(timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "Stokes-ADR on|compaction MINRES|stokes dim|48\^3|mode F vs|config 1 at|PASSED|FAILED|ERROR|Error|assert|passed|failed" | head -120) > gpurun_out/r2_run13_pytest.log
cat gpurun_out/r2_run13_pytest.log | grep -vE "^PASSED"
python examples/stokes_adr_refmesh.py 1 2>&1 | tail -3
python examples/ccs2d_periodic_flux.py 24 2 2>&1 | tail -3
