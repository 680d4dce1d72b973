set -x
nvidia-smi --query-gpu=name --format=csv
(timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -m gpu -q -rA -k two_ranks 2>&1 | tail -120) > gpurun_out/r2_run2_pytest.log
for tr in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 8 --skip-cpu --p-guess previous --transport $tr > gpurun_out/r2_run2_bench_n2_$tr.json 2> gpurun_out/r2_run2_bench_n2_$tr.err
done
tail -15 gpurun_out/r2_run2_pytest.log
head -c 600 gpurun_out/r2_run2_bench_n2_peer.json; echo; head -c 600 gpurun_out/r2_run2_bench_n2_nccl.json
tail -3 gpurun_out/r2_run2_bench_n2_peer.err
