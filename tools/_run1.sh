set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
(timeout 1100 python -m pytest tests -m gpu -q -rA 2>&1 | tail -150) > gpurun_out/r2_run1_pytest.log
timeout 400 python bench.py --steps 10 --warmup 10 --skip-cpu > gpurun_out/r2_run1_bench_extrap.json 2> gpurun_out/r2_run1_bench_extrap.err
timeout 400 python bench.py --steps 10 --warmup 10 --skip-cpu --p-guess previous > gpurun_out/r2_run1_bench_prev.json 2> gpurun_out/r2_run1_bench_prev.err
tail -5 gpurun_out/r2_run1_pytest.log
cat gpurun_out/r2_run1_bench_extrap.json | head -c 1500
