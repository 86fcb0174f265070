# round 2, call 8 (2 GPUs): peer push exchange — tests, then bench N=2 with push / p2p / ce
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_training.py -x -q 2>&1 | tail -5
for c in push p2p ce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --collective $c > gpurun_out/bench_r2g_n2_$c.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/bench_r2g_n2_$c.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$c", "ms %.3f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["multi_gpu"])
    elif "rror" in l: print(l[:300])
PY
done
