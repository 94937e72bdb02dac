"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) per kernel name: launches, total time, share.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/launches_xxx_summary.csv"""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(anonymous namespace\)::', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    name = name.split('(')[0] if not name.startswith('at::') else name[:160]
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print('# ' + c)
print(f'# total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches')
print('kernel,launches,total_us,share,min_us,max_us')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[2]:.1f},{a[3]:.1f}')
