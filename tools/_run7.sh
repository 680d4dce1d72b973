set -x
export PROBE_WARM=0.05
PROBE_ONLY=1,1 python tools/spmv_probe.py 200 200 5 > gpurun_out/r2_run7_plain_c8.log 2>&1 &&
PROBE_ONLY=1,1 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 10 -c 2 -o gpurun_out/r2_ncu_spmv_c8 python tools/spmv_probe.py 200 200 5 > gpurun_out/r2_run7_ncu_c8.log 2>&1
PROBE_ONLY=0,1 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 10 -c 2 -o gpurun_out/r2_ncu_spmv_i32 python tools/spmv_probe.py 200 200 5 > gpurun_out/r2_run7_ncu_i32.log 2>&1
tail -n 3 gpurun_out/r2_run7_ncu_c8.log
ls -la gpurun_out/*.ncu-rep
