# round 2: ncu full capture after the warp-uniform issue loops
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --modes= --profiler-range"
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r2s $B > gpurun_out/ncu_r2s_2.log 2>&1
tail -2 gpurun_out/ncu_r2s_2.log | cut -c1-300
