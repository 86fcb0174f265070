B="python bench.py --steps 2 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r1b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|gemm|softmax|split_bf16" -c 200 --csv --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/ncu_r1b_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmm_pipe|gemm_bf16x3" -s 15 -c 5 -o gpurun_out/fwd_r1b $B > gpurun_out/ncu_r1b_2.log 2>&1
tail -2 gpurun_out/ncu_r1b_2.log
