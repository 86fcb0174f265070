# round 2: A/B of the 12-epilogue-warp fused-aggregation tile (fp16 output), after the log/rcp fix
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_aligned.py tests/test_gpu_stream.py -x -q 2>&1 | tail -3
FITGNN_AGG_WIDE=2 timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_stream.py -x -q 2>&1 | tail -3
for aw in 0 2 0 2; do
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-projection --no-e2e --modes= --agg-wide $aw > gpurun_out/bench_r2v_$aw.log 2>&1
python - $aw <<'PY'
import json,sys
j=json.loads([l for l in open(f'gpurun_out/bench_r2v_{sys.argv[1]}.log') if l.startswith('{')][-1])
print('agg_wide',sys.argv[1],'ms',round(j['ms_per_step'],3), {k:round(v['ms'],3) for k,v in j['kernels'].items()})
PY
done
