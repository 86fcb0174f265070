# 2 GPUs: the driver's launch line with the session-2 kernels (default collective), gathered == single-GPU check inside bench.py
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" 2> gpurun_out/bench_r2ao.err | grep '^{' | tail -1; }
show() { python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read())
except Exception as e:
    print('no result', sys.argv[1], e); sys.exit(0)
m=j['multi_gpu']
print(j['n_gpus'], 'ms', round(j['ms_per_step'],3), 'value %.3e'%j['value'], 'e2e', j.get('e2e',{}).get('ms_per_step'), m['collective'], 'ctas', m.get('push_ctas'), 'reserve', m.get('sm_reserve'), 'kernels', [round(x,3) for x in m['rank_kernel_ms']], 'exposed', round(m['exposed_ms'],3), 'err', m['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'sharded', round((m.get('sharded') or {}).get('ms_per_step',0),3), j['dtype'][:40])
PY
}
run 2 > gpurun_out/r2ao_n2.json; show gpurun_out/r2ao_n2.json
tail -3 gpurun_out/bench_r2ao.err
