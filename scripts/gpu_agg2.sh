timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-projection > gpurun_out/bench_agg2a.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg2a.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_agg2b.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg2b.log
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --profiler-range"
$B > gpurun_out/plain_r1e.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16x3" -c 3 -o gpurun_out/fwd_r1e $B > gpurun_out/ncu_r1e.log 2>&1
tail -2 gpurun_out/ncu_r1e.log
