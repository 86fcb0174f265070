python bench.py > gpurun_out/bench_n1.log 2>&1; tail -c 400 gpurun_out/bench_n1.log
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.log 2>&1
tail -c 400 gpurun_out/bench_n$N.log
