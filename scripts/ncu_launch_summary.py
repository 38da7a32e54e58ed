"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per kernel launches, total/mean device time, share of the captured time, mean DRAM bytes per launch.
usage: ncu_launch_summary.py launches.csv [traffic.json workload]   (the optional pair updates profiles/traffic.json)"""
import collections, csv, json, os, sys

lines = open(sys.argv[1]).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
by = collections.OrderedDict()
for r in csv.DictReader(lines[start:]):
    d = by.setdefault(r['ID'], {'name': r['Kernel Name'].split('(')[0].replace('void ', '')})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if 'time' in r['Metric Name']:
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)          # -> us
    else:
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    d[r['Metric Name']] = v
agg = collections.OrderedDict()
for d in by.values():
    a = agg.setdefault(d['name'], [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0.0)
    a[2] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print('%-44s %8s %12s %8s %12s %14s' % ('kernel', 'launches', 'total ms', 'share', 'mean us', 'dram B/launch'))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-44s %8d %12.3f %7.1f%% %12.2f %14.0f' % (k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0], a[2] / a[0]))
if len(sys.argv) > 3:
    path, wl = sys.argv[2], sys.argv[3]
    t = json.load(open(path)) if os.path.exists(path) else {}
    t[wl] = {k: a[2] / a[0] for k, a in agg.items() if a[2] > 0}
    json.dump(t, open(path, 'w'), indent=1)
