"""Last step of an ncu launch list (csv from --metrics gpu__time_duration.sum ...): kernel name and ms per launch.
python tools/launches.py <csv>"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
seq, cur = [], None
for r in rows[1:]:
    if r[ii] != cur:
        cur = r[ii]
        seq.append([r[ki][:44], {}])
    seq[-1][1][r[mi]] = r[vi]
names = [n for n, _ in seq]
last = len(names) - 1 - names[::-1].index([n for n in names if n.startswith('k_docstart')][0])
for n, m in seq[last:]:
    t = float(m['gpu__time_duration.sum'].replace(',', '')) / 1e6
    if t < 0.02: continue
    rest = {k.split('.')[0].replace('smsp__', '').replace('sm__', '')[-34:]: v for k, v in m.items() if k != 'gpu__time_duration.sum'}
    print('%-46s %8.3f ms  %s' % (n, t, rest))
