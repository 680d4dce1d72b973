set -x
(timeout 900 python -m pytest tests/test_gpu_compaction.py tests/test_gpu_api.py -m gpu -q -rA 2>&1 | grep -E "compaction MINRES|PASSED|FAILED|Error|error|assert|passed|failed" | head -40)
for n in 8 16; do
timeout 400 python bench.py --workload compaction --size $n --steps 3 --warmup 2 > gpurun_out/r2_run12_bench_compaction_$n.json 2> gpurun_out/r2_run12_bench_compaction_$n.err
tail -n 2 gpurun_out/r2_run12_bench_compaction_$n.err | cut -c1-300
cat gpurun_out/r2_run12_bench_compaction_$n.json | cut -c1-300; python -c "
import json;d=json.loads(open('gpurun_out/r2_run12_bench_compaction_$n.json').read().strip().splitlines()[-1]);print(d['krylov_iterations_per_step'], d['roofline'], d['assembly_ms_per_step'])"
done
