set -x
(timeout 900 python -m pytest tests/test_gpu_compaction.py tests/test_gpu_stokes.py tests/test_gpu_api.py tests/test_gpu_spmv_variants.py tests/test_gpu_api_modef.py -m gpu -q -rA 2>&1 | grep -E "compaction MINRES|stokes dim|PASSED|FAILED|Error|error|assert|passed|failed" | head -70) > gpurun_out/r2_run11_pytest.log
cat gpurun_out/r2_run11_pytest.log
for w in darcy128 compaction stokes; do
timeout 400 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r2_run11_bench_$w.json 2> gpurun_out/r2_run11_bench_$w.err
tail -n 3 gpurun_out/r2_run11_bench_$w.err | cut -c1-300
cat gpurun_out/r2_run11_bench_$w.json | cut -c1-1500
done
