# 8 GPUs, session-2 kernels: push (24 CTAs) vs hybrid push + copy engines (2 / 3 peers by copy engine)
run() { n=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e "$@" 2> gpurun_out/bench_r2aq.err | grep '^{' | tail -1; }
show() { python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
except Exception as e:
    print('no result', sys.argv[1], e); sys.exit(0)
m=j['multi_gpu']
print(j['n_gpus'], 'ms', round(j['ms_per_step'],3), 'value %.3e'%j['value'], m['collective'], 'ctas', m.get('push_ctas'), 'ce_peers', m.get('ce_peers'), 'reserve', m.get('sm_reserve'), 'kernels', [round(x,3) for x in m['rank_kernel_ms']], 'exposed', round(m['exposed_ms'],3), 'err', m['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'sharded', round((m.get('sharded') or {}).get('ms_per_step',0),3))
PY
}
run 8 --collective push --push-ctas 24 > gpurun_out/r2aq_a.json; show gpurun_out/r2aq_a.json
run 8 --collective push --push-ctas 24 --ce-peers 2 > gpurun_out/r2aq_b.json; show gpurun_out/r2aq_b.json
run 8 --collective push --push-ctas 16 --ce-peers 3 > gpurun_out/r2aq_c.json; show gpurun_out/r2aq_c.json
tail -3 gpurun_out/bench_r2aq.err
