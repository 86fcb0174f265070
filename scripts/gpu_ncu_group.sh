timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_group -s 3 -c 1 -o gpurun_out/spmm_group_r1t python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-projection > gpurun_out/ncu_group.log 2>&1
tail -2 gpurun_out/ncu_group.log
