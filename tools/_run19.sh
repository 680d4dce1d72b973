set -x
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-api"
$CMD > gpurun_out/r2_run19_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_run19_ncu.log 2>&1
tail -n 2 gpurun_out/r2_run19_plain.log | cut -c1-300
tail -n 2 gpurun_out/r2_run19_ncu.log | cut -c1-300
wc -l gpurun_out/r2_launches_bench.csv
