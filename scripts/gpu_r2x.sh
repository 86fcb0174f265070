# round 2: 2 GPUs — peer tests (deeper push pipeline), bench p2p (default) vs push with reserved SMs
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -3
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e "$@" 2> gpurun_out/bench_r2x.err | grep '^{' | tail -1; }
show() { python - "$1" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read())
m=j['multi_gpu']
print(j['n_gpus'], 'ms', round(j['ms_per_step'],3), 'value %.3e'%j['value'], m['collective'], 'ctas', m.get('push_ctas'), 'reserve', m.get('sm_reserve'), 'kernels', [round(x,3) for x in m['rank_kernel_ms']], 'exposed', round(m['exposed_ms'],3), 'err', m['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'sharded', round((m.get('sharded') or {}).get('ms_per_step',0),3))
PY
}
run 2 > gpurun_out/r2x_a.json; show gpurun_out/r2x_a.json
run 2 --collective push --push-ctas 4 > gpurun_out/r2x_b.json; show gpurun_out/r2x_b.json
run 2 --collective push --push-ctas 8 > gpurun_out/r2x_c.json; show gpurun_out/r2x_c.json
run 2 --collective push --push-ctas 8 --sm-reserve 0 > gpurun_out/r2x_d.json; show gpurun_out/r2x_d.json
