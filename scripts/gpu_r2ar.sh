# layer 1 on fp16 planes (grouped SpMM writes one fp16 plane, W0 one fp16 plane): tests + A/B
timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_stream.py tests/test_gpu_peer.py -m gpu -x -q -k "fp16 or f16 or grouped or peer" 2>&1 | tail -4
for v in 1 0 1 0; do
FITGNN_F16_LAYER0=$v timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-projection --cpu-seconds 4 > gpurun_out/bench_r2ar_$v.log 2> gpurun_out/bench_r2ar.err; tail -3 gpurun_out/bench_r2ar.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2ar_$v.log").read().strip().splitlines()[-1])
print("f16_layer0=$v", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()), (l.get("parity") or {}).get("max_rel_err"), (l.get("parity") or {}).get("rows"))
PY
done
