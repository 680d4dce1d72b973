set -x
(timeout 900 python -m pytest tests/test_gpu_api_modef.py tests/test_gpu_api.py -m gpu -q -rA 2>&1 | tail -80) > gpurun_out/r2_run4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 5 --skip-cpu > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err
tail -30 gpurun_out/r2_run4_pytest.log
tail -5 gpurun_out/r2_run4_bench.err
head -c 300 gpurun_out/r2_run4_bench.json
