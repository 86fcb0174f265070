# round 2, call 9 (2 GPUs): push-kernel bandwidth per CTA count; bench N=2 with fewer push CTAs
timeout 300 python scripts/bench_push.py 2>&1 | tail -9
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --collective push --push-ctas $n --no-e2e > gpurun_out/bench_r2h_n2_push$n.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/bench_r2h_n2_push$n.log"):
    if l.startswith("{"):
        d = json.loads(l); print("push ctas $n", "ms %.3f" % d["ms_per_step"], d["multi_gpu"]["rank_kernel_ms"], d["multi_gpu"]["exposed_ms"])
    elif "rror" in l: print(l[:300])
PY
done
