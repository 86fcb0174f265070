N=${1:-8}
for c in 1 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --chunks $c --no-e2e --no-cpu-baseline > gpurun_out/bench_n${N}_c$c.log 2>&1
tail -c 300 gpurun_out/bench_n${N}_c$c.log
done
