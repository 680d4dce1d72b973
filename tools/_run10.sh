set -x
for v in -1 0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 8 --skip-cpu --pcg-variant $v > gpurun_out/r2_run10_bench_n8_pcg$v.json 2> gpurun_out/r2_run10_bench_n8_pcg$v.err
tail -n 2 gpurun_out/r2_run10_bench_n8_pcg$v.err | cut -c1-300
done
python - <<'PY'
import json
for v in ("-1","0"):
    try:
        d=json.loads(open("gpurun_out/r2_run10_bench_n8_pcg%s.json"%v).read().strip().splitlines()[-1])
        print("pcg",v,"value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"], "p its",[i["p"] for i in d["krylov_iterations_per_step"]][:4], {k:d["roofline"][k] for k in ("avg_launch_ms","share_of_step")})
    except Exception as e: print(v,"ERR",e)
PY
