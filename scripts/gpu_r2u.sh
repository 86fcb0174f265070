# round 2: 8 GPUs — default bench (ce gathered + sharded variant), then p2p, then 4 GPUs default
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" 2> gpurun_out/bench_r2u.err | grep '^{' | tail -1; }
show() { python - "$1" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read())
m=j['multi_gpu']
print(j['n_gpus'], 'ms', round(j['ms_per_step'],3), 'value %.3e'%j['value'], 'e2e', round(j['e2e']['ms_per_step'],3) if j.get('e2e') else None, m['collective'], 'kernels', [round(x,3) for x in m['rank_kernel_ms']], 'exposed', round(m['exposed_ms'],3), 'err', m['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'other', {k:(round(v,3) if isinstance(v,float) else v) for k,v in (m.get('sharded') or m.get('gathered') or {}).items()})
PY
}
run 8 > gpurun_out/bench_r2u_8gpu.json; show gpurun_out/bench_r2u_8gpu.json
run 8 --collective p2p --no-e2e > gpurun_out/bench_r2u_8gpu_p2p.json; show gpurun_out/bench_r2u_8gpu_p2p.json
run 4 > gpurun_out/bench_r2u_4gpu.json; show gpurun_out/bench_r2u_4gpu.json
