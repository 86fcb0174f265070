timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_aligned.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -6
python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-projection > gpurun_out/bench_r1g_n1.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1g_n1.log 2>&1 | head -6
for C in p2p; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $C > gpurun_out/bench_r1g_n2_$C.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1g_n2_$C.log 2>&1 | head -6; grep -v "^{" gpurun_out/bench_r1g_n2_$C.log | grep -v "^\*\*\|OMP_NUM\|^$" | tail -5
done
