# round 2: warp-uniform MMA issue / TMA producer loops — tests + headline in both precisions
timeout 1500 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_aligned.py tests/test_gpu_stream.py tests/test_gpu_training.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-projection --modes alt_precision > gpurun_out/bench_r2r.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2r.log
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_r2r.log') if l.startswith('{')][-1])
print(j['ms_per_step'], j['value'], j['e2e'])
for k,v in j['kernels'].items(): print(k, round(v['ms'],3), round(v['GBps']), round(v['TFLOPs'],1))
PY
