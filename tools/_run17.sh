set -x
timeout 600 python bench.py --skip-cpu > gpurun_out/r2_run17_bench.json 2> gpurun_out/r2_run17_bench.err
tail -n 3 gpurun_out/r2_run17_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run17_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [x["p"] for x in d["krylov_iterations_per_step"]], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["true_relative_residuals_last_step"], d["gpu_launches"])
PY
