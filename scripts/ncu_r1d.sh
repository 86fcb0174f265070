B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --profiler-range"
$B > gpurun_out/plain_r1d.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/ncu_r1d_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16x3" -c 3 -o gpurun_out/fwd_r1d $B > gpurun_out/ncu_r1d_2.log 2>&1
tail -2 gpurun_out/ncu_r1d_2.log
