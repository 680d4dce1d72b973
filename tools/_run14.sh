set -x
python tools/_dbg.py 2>&1 | tail -8
(timeout 600 python -m pytest tests/test_gpu_compaction.py -m gpu -q -rA -k p2_space 2>&1 | grep -E "P2 porosity|assert|Error|passed|failed" | head)
