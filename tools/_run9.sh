set -x
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2_run9_pytest.log
cat gpurun_out/r2_run9_pytest.log
python tools/spmv_probe.py 200 25 400 2>&1 | grep "c8=.*lpr=1"
