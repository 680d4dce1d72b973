set -x
(timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "Stokes-ADR on|P2 porosity|line PCG|line BiCG|compaction MINRES|stokes dim|48\^3|mode F vs|mode F periodic|config 1 at|FAILED|ERROR|Error|assert|passed|failed" | head -150) > gpurun_out/r2_run22_pytest.log
cat gpurun_out/r2_run22_pytest.log | tail -30
timeout 900 python bench.py > gpurun_out/r2_run22_bench.json 2> gpurun_out/r2_run22_bench.err
tail -n 3 gpurun_out/r2_run22_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run22_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [(x["c0"],x["p"]) for x in d["krylov_iterations_per_step"]][::5], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["true_relative_residuals_last_step"], d["gpu_launches"], d["phase_ms_per_step"], d["cpu_baseline"])
PY
