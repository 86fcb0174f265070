for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1q_n$N.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1q_n$N.log 2>&1 | head -6
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --collective p2p --no-e2e > gpurun_out/bench_r1q_n8_p2p.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1q_n8_p2p.log 2>&1 | head -1
