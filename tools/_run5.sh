set -x
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2_run5_pytest.log
tail -5 gpurun_out/r2_run5_pytest.log
timeout 600 python bench.py --steps 10 --warmup 10 --skip-cpu > gpurun_out/r2_run5_bench.json 2> gpurun_out/r2_run5_bench.err
tail -3 gpurun_out/r2_run5_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_run5_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"])
print({k:d["roofline"][k] for k in ("achieved","frac","avg_launch_ms","share_of_step")})
PY
