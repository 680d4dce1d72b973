set -x
(timeout 900 python -m pytest tests/test_gpu_modef.py tests/test_gpu_parity_at_size.py -m gpu -q -rA 2>&1 | grep -E "line PCG|line BiCG|mode F periodic|48\^3|FAILED|ERROR|Error|assert|passed|failed" | head -50)
timeout 600 python bench.py --skip-cpu > gpurun_out/r2_run21_bench.json 2> gpurun_out/r2_run21_bench.err
tail -n 3 gpurun_out/r2_run21_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run21_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [(x["c0"],x["p"]) for x in d["krylov_iterations_per_step"]], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["true_relative_residuals_last_step"], d["gpu_launches"], d["bicgstab_to_gmres_fallbacks"])
PY
