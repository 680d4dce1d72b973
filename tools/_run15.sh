set -x
(timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "Stokes-ADR on|P2 porosity|compaction MINRES|stokes dim|48\^3|mode F vs|config 1 at|PASSED|FAILED|ERROR|Error|assert|passed|failed" | head -150) > gpurun_out/r2_run15_pytest.log
grep -vE "^PASSED" gpurun_out/r2_run15_pytest.log
timeout 600 python bench.py > gpurun_out/r2_run15_bench.json 2> gpurun_out/r2_run15_bench.err
tail -n 5 gpurun_out/r2_run15_bench.err | cut -c1-400
cut -c1-600 gpurun_out/r2_run15_bench.json
