set -x
for tr in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 8 --skip-cpu --p-guess previous --transport $tr > gpurun_out/r2_run3_bench_n8_$tr.json 2> gpurun_out/r2_run3_bench_n8_$tr.err
done
tail -c 1200 gpurun_out/r2_run3_bench_n8_peer.json; echo; tail -c 600 gpurun_out/r2_run3_bench_n8_nccl.json
tail -3 gpurun_out/r2_run3_bench_n8_peer.err
