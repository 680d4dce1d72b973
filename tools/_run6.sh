set -x
python tools/spmv_probe.py 200 200 200 > gpurun_out/r2_run6_probe_n200.txt 2>&1
python tools/spmv_probe.py 200 25 400 > gpurun_out/r2_run6_probe_slab8.txt 2>&1
python tools/spmv_probe.py 200 50 400 > gpurun_out/r2_run6_probe_slab4.txt 2>&1
cat gpurun_out/r2_run6_probe_n200.txt gpurun_out/r2_run6_probe_slab8.txt gpurun_out/r2_run6_probe_slab4.txt
(timeout 600 python -m pytest tests/test_gpu_spmv_variants.py tests/test_gpu_modef.py -m gpu -q 2>&1 | tail -5)
