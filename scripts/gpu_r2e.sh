# round 2, call 6: ncu of the block-staged SpMM on the products-small cluster_node pack
B="python bench.py --workload products-small --only-modes --modes cluster --mode-steps 1 --warmup 1 --profiler-range"
$B > gpurun_out/plain_r2e.log 2>&1 && tail -c 400 gpurun_out/plain_r2e.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spmm_block -c 2 -o gpurun_out/r2e_spmm_block $B > gpurun_out/ncu_r2e.log 2>&1
tail -3 gpurun_out/ncu_r2e.log
ls -la gpurun_out/r2e_spmm_block.ncu-rep
