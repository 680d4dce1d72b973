set -x
(timeout 900 python -m pytest tests/test_gpu_modef.py -m gpu -q -rA -k "line" 2>&1 | grep -E "line PCG|mode F periodic|PASSED|FAILED|ERROR|Error|assert|passed|failed" | head -50)
timeout 600 python bench.py --p-precond line --skip-cpu > gpurun_out/r2_run16_bench_line.json 2> gpurun_out/r2_run16_bench_line.err
tail -n 4 gpurun_out/r2_run16_bench_line.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run16_bench_line.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [x["p"] for x in d["krylov_iterations_per_step"]], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["true_relative_residuals_last_step"])
PY
