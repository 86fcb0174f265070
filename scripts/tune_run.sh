for f in gpurun_in/libfitgnn_*.so; do FITGNN_B200_LIB=$PWD/$f python scripts/bench_spmm.py 2>&1 | tail -1; done
