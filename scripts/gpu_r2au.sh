# 2 GPUs, final code: push exchange + e2e (H2D-alone block) under the driver's launch line
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" 2> gpurun_out/bench_r2au.err | grep '^{' | tail -1; }
run 2 --collective push > gpurun_out/r2au_n2.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2au_n2.json").read())
m=j['multi_gpu']
print(j['n_gpus'], 'ms', round(j['ms_per_step'],3), 'value %.3e'%j['value'], m['collective'], 'ctas', m.get('push_ctas'), 'kernels', [round(x,3) for x in m['rank_kernel_ms']], 'exposed', round(m['exposed_ms'],3), 'err', m['gathered_vs_single_gpu_max_abs_err_all_ranks'], 'sharded', round((m.get('sharded') or {}).get('ms_per_step',0),3))
e=j['e2e']; print('e2e ms', round(e['ms_per_step'],2), 'h2d/rank in step GB/s', round(e['h2d_GBps_per_rank_in_step'],1), e['h2d_alone'])
print(j['dtype'][:60], j.get('parity'))
PY
tail -2 gpurun_out/bench_r2au.err
