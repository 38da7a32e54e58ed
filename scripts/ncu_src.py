"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line (top-N by instructions executed).
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_src.py src.csv [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, '']); tot = [0, 0, 0, 0, 0]
cur = None; fname = ''; h = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        h = r
        li, si, ii, ti = h.index('Line No'), h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
        l2l = h.index('L2 Theoretical Sectors Local') if 'L2 Theoretical Sectors Local' in h else None
        l2g = h.index('L2 Theoretical Sectors Global') if 'L2 Theoretical Sectors Global' in h else None
        continue
    if h is None or len(r) < len(h):
        continue
    if r[li].strip().isdigit():
        cur = (fname, int(r[li])); agg[cur][5] = r[1].strip()[:100]
        continue
    if cur is None:
        continue
    vals = [int(float(r[c])) if c is not None and r[c] not in ('', '-') else 0 for c in (si, ii, ti, l2l, l2g)]
    a = agg[cur]
    for j, v in enumerate(vals):
        a[j] += v; tot[j] += v
print('total: samples %d inst %d thread-inst %d L2-local-sectors %d L2-global-sectors %d' % tuple(tot))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%s:%4d inst %5.1f%% smp %5.1f%% thr/inst %4.1f L2loc %5.1f%% L2glob %5.1f%% | %s' % (
        f[:14], ln, 100 * a[1] / max(tot[1], 1), 100 * a[0] / max(tot[0], 1), a[2] / max(a[1], 1), 100 * a[3] / max(tot[3], 1), 100 * a[4] / max(tot[4], 1), a[5]))
