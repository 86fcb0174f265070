# round 2: 2 GPUs — peer tests, default bench (gathered headline + sharded variant), fp16x2 default
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2t_2gpu.log 2> gpurun_out/bench_r2t_2gpu.err; tail -3 gpurun_out/bench_r2t_2gpu.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_r2t_2gpu.log') if l.startswith('{')][-1])
print(j['n_gpus'], j['ms_per_step'], j['value'], j['e2e']['ms_per_step'] if j.get('e2e') else None)
print(j['multi_gpu'])
PY
