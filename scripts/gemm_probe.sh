python scripts/bench_gemm.py
echo "--- streaming plan"
FITGNN_GEMM_WS=0 python scripts/bench_gemm.py
