for C in ce mc; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --collective $C --no-e2e > gpurun_out/bench_r1i_n8_$C.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1i_n8_$C.log 2>&1 | head -6; grep -v "^{" gpurun_out/bench_r1i_n8_$C.log | grep -v "^\*\*\|OMP_NUM\|^$" | tail -4
done
