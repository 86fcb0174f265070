import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms/step %.3f  nodes/s %.4g  e2e ms %.2f  launches %d' % (d['ms_per_step'], d['value'], d.get('e2e', {}).get('ms_per_step', 0), d['gpu_launches']))
        for k, v in d['kernels'].items():
            print(f"  {k:18s} {v['ms']:8.3f} ms  {v['GBps']:8.1f} GB/s  {v['TFLOPs']:7.1f} TF  share {v['share']:.1%}")
        print('  roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 3), '| spmm', d['roofline_spmm']['kernel'], round(d['roofline_spmm']['frac'], 3), d['clocks'])
        if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'])
