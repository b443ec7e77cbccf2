import json, sys
for path in sys.argv[1:]:
    l=[x for x in open(path) if x.startswith('{')][-1]
    d=json.loads(l)
    print(path, 'value %.0f img/s  ms/step %.3f  e2e %.0f  launches %d  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    for k,v in d['roofline']['stages'].items():
        print(f"  {k:20s} ms {v['ms_per_step']:7.3f} n {v['launches_per_step']:4d} share {v['share']:.3f} ach {v['achieved']:8.1f} {v['unit']} frac {v['frac_of_peak']:.3f}")
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
