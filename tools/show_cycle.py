"""Print the numbers of one tools/gpu_cycle.sh run.  Usage: python tools/show_cycle.py TAG"""
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/{tag}_bench.json'))
    print('c2 kernel ms', round(d['roofline']['kernel_ms'], 5), 'step', round(d['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'e2e ms', round(d['e2e']['ms_per_step'], 3))
except Exception as e:
    print('bench:', e, open(f'gpurun_out/{tag}_bench.err').read()[-800:])
for l in open(f'gpurun_out/{tag}_configs.jsonl'):
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], round(d['ms'], 4), 'frac', round(d['frac_of_roofline'], 3), d.get('bound'))
print(open(f'gpurun_out/{tag}_tests.log').read()[-400:])
