set -x
python tools/reftensor_probe.py 40 2 2>&1 | tail -7
python tools/reftensor_probe.py 24 3 2>&1 | tail -7
(timeout 900 python -m pytest tests/test_gpu_moder.py tests/test_gpu_compaction.py tests/test_gpu_stokes.py tests/test_gpu_api.py tests/test_gpu_stokes_refmesh.py -m gpu -q 2>&1 | tail -5)
