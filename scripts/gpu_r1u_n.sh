N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1u_n$N.log 2>&1
tail -3 gpurun_out/bench_r1u_n$N.log | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/bench_r1u_n$N.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'nodes/s %.4g'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],2), d['multi_gpu']['collective'], 'verify', d['multi_gpu']['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'rank_kernel_ms', [round(x,2) for x in d['multi_gpu']['rank_kernel_ms']], d['config']['features'], d['clocks'])
PY
