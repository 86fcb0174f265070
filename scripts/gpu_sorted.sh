timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_peer.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -5
for P in degree order; do
python bench.py --steps 10 --no-cpu-baseline --no-projection --align-policy $P > gpurun_out/bench_r1j_$P.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1j_$P.log 2>&1 | head -6; grep -o '"schedule": "[^"]*"' gpurun_out/bench_r1j_$P.log; grep -o '"build_ms": [0-9.]*' gpurun_out/bench_r1j_$P.log
done
