set -x
CMD="python bench.py --steps 1 --warmup 3 --roof-steps 1 --skip-cpu --skip-api"
$CMD > gpurun_out/r2_run24_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench_final.csv $CMD > gpurun_out/r2_run24_ncu.log 2>&1
tail -n 1 gpurun_out/r2_run24_plain.log | cut -c1-200
wc -l gpurun_out/r2_launches_bench_final.csv
for w in darcy128 compaction stokes; do
timeout 400 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r2_run24_bench_$w.json 2> gpurun_out/r2_run24_bench_$w.err
tail -n 2 gpurun_out/r2_run24_bench_$w.err | cut -c1-300
cut -c1-400 gpurun_out/r2_run24_bench_$w.json
done
