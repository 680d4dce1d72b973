set -x
(timeout 300 python -m pytest tests/test_gpu_spmv_variants.py tests/test_gpu_moder.py -m gpu -q -x -k "single_reduction or reftensor_tiles" 2>&1 | tail -8)
