timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_aligned.py -x -q 2>&1 | tail -6
for C in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $C > gpurun_out/bench_r1f_n2_$C.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1f_n2_$C.log 2>&1 | head -6; grep -v "^{" gpurun_out/bench_r1f_n2_$C.log | tail -5
done
