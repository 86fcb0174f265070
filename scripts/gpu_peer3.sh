timeout 600 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -6
for C in ce p2p; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $C --no-e2e > gpurun_out/bench_r1h_n2_$C.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1h_n2_$C.log 2>&1 | head -1; grep -v "^{" gpurun_out/bench_r1h_n2_$C.log | grep -v "^\*\*\|OMP_NUM\|^$" | tail -5
done
