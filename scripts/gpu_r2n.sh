# round 2: configs 1-4 with per-kernel rooflines, Physics-shaped GEMM ncu, per-query block
timeout 1500 python scripts/bench_configs.py 2>&1 | tail -8 | cut -c1-400
cat gpurun_out/configs_r2.md | tail -12 | cut -c1-900
B="python scripts/bench_physics_gemm.py --profiler-range"
ncu --set full --clock-control none --profile-from-start off -k regex:gemm_bf16x3 -c 1 -o gpurun_out/r2n_physics_gemm $B > gpurun_out/ncu_r2n.log 2>&1; tail -2 gpurun_out/ncu_r2n.log
timeout 600 python bench.py --only-modes --modes per_query > gpurun_out/bench_r2n_pq.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2n_pq.log
