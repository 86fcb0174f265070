python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --collective ce > gpurun_out/bench_r1h_n8_ce.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1h_n8_ce.log 2>&1 | head -6
for C in ce p2p; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --collective $C --no-e2e > gpurun_out/bench_r1h_n4_$C.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1h_n4_$C.log 2>&1 | head -6
done
