timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_r1v.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1v.log 2>&1 | head -8
python - <<PY
import json
for l in open('gpurun_out/bench_r1v.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['projection']['ac']['ms'], d['projection']['group_by_part_ms'], d['pack']['build_ms'], d['pack']['build_first_call_ms'], d['features_check'])
PY
