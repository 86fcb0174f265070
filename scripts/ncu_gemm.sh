B="python bench.py --steps 2 --precision bf16x3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 9 -c 3 -o gpurun_out/gemm_r1a $B > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
