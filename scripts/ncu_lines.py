"""Aggregate an ncu SASS source page by CUDA source line using nvdisasm -g line markers.
usage: ncu_lines.py sass.csv dis.txt kernel_mangled_substr [top]"""
import csv, re, sys, collections
sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# line markers from nvdisasm: '//## File "...", line N' precede instructions '/*0000*/  OPC ...;'
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur = None; seq = []
stack = []
for l in lines[start + 1:]:
    if l.startswith('//---------------------'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
rows = list(csv.reader(open(sass_csv)))
# first kernel instance only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ('Kernel Name', 'Address'): break
    body.append(r)
ci = hdr.index('Instructions Executed'); si = hdr.index('# Samples'); ti = hdr.index('Thread Instructions Executed')
print('sass instrs', len(body), 'disasm instrs', len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, loc in zip(body, seq):
    a = agg[loc[:2] if loc else None]
    for j, c in enumerate((ci, si, ti)):
        v = int(float(r[c] or 0)); a[j] += v; tot[j] += v
print('total inst', tot[0], 'samples', tot[1], 'thread inst', tot[2])
for loc, (n, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(loc, 'inst %d (%.1f%%)  samples %d (%.1f%%)  thr/inst %.1f' % (n, 100 * n / tot[0], s, 100 * s / max(tot[1], 1), t / max(n, 1)))
