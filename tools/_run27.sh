set -x
timeout 100 python tools/_dbg_ref.py 2>&1 | grep "max err"
timeout 200 python tools/reftensor_probe.py 40 2 2>&1 | tail -6
timeout 200 python tools/reftensor_probe.py 24 3 2>&1 | tail -6
(timeout 300 python -m pytest tests/test_gpu_moder.py tests/test_gpu_compaction.py tests/test_gpu_stokes.py tests/test_gpu_api.py tests/test_gpu_stokes_refmesh.py -m gpu -q -x 2>&1 | tail -5)
