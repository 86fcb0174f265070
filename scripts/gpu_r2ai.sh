# ncu full capture of the 12-warp W-stationary pair kernel (micro-benchmark, one W plane) + of the whole forward (4 kernels)
NCU=1 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3_kernel --launch-skip 3 -c 1 -o gpurun_out/gemm1_r2ai python scripts/bench_gemm1_f16.py > gpurun_out/ncu_r2ai.log 2>&1
tail -2 gpurun_out/ncu_r2ai.log | cut -c1-200
