import sys,json
for l in sys.stdin:
    l=l.strip()
    if '{' in l and 'metric' in l:
        d=json.loads(l[l.index('{'):])
        print(d['config']['particles_per_gpu'], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e6,1),'M/s e2e', d['e2e'] and (round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],1)), 'launches', d['gpu_launches'], d['clocks'])
        print('   ',{k:round(v,2) for k,v in d['phases_ms'].items()})
        if d.get('cpu_baseline'): print('    cpu', d['cpu_baseline']['value'])
    else: print(l[:300])
