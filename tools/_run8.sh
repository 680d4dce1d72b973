set -x
python tools/spmv_probe.py 200 200 200 > gpurun_out/r2_run8_probe_n200.txt 2>&1
python tools/spmv_probe.py 200 25 400 > gpurun_out/r2_run8_probe_slab8.txt 2>&1
cat gpurun_out/r2_run8_probe_n200.txt gpurun_out/r2_run8_probe_slab8.txt
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5)
timeout 600 python bench.py --steps 10 --warmup 10 --skip-cpu > gpurun_out/r2_run8_bench.json 2> gpurun_out/r2_run8_bench.err
tail -2 gpurun_out/r2_run8_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_run8_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"])
print({k:d["roofline"][k] for k in ("achieved","frac","avg_launch_ms","share_of_step")})
PY
