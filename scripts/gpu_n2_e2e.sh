python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1l_n2.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1l_n2.log 2>&1 | head -6; grep -v "^{" gpurun_out/bench_r1l_n2.log | grep -v "^\*\*\|OMP_NUM\|^$" | tail -5
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_r1l_n2.log") if l.startswith("{")][0]
print(d["e2e"], d["multi_gpu"])
PY
