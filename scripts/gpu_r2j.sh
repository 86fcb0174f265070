# round 2: ncu of the block-dense MMA aggregation on the products-small cluster_node pack
B="python bench.py --workload products-small --only-modes --modes cluster --mode-steps 1 --warmup 1 --profiler-range"
$B > gpurun_out/plain_r2j.log 2>&1 && python scripts/show_modes.py gpurun_out/plain_r2j.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spmm_mma -c 1 -o gpurun_out/r2j_spmm_mma $B > gpurun_out/ncu_r2j.log 2>&1
tail -2 gpurun_out/ncu_r2j.log
