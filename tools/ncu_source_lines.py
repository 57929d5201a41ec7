"""Map ncu per-SASS-instruction counters of k_fused to CUDA source lines (via nvdisasm line info)."""
import csv, re, subprocess, sys, os, collections, tempfile
rep = sys.argv[1]; fn = sys.argv[2] if len(sys.argv) > 2 else 'k_fusedILb0'; srcfile = sys.argv[3] if len(sys.argv) > 3 else 'jb_fused.cu'
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {root}/jieba_go_b200/libjieba_b200.so >/dev/null 2>&1", shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith(srcfile.replace('.cu', '') + '.sm')][0]
sass = subprocess.run(f"nvdisasm -g -c {tmp}/{cub}", shell=True, capture_output=True, text=True).stdout.split('\n')
in_fn = False; cur = None; seq = []
for l in sass:
    if re.match(r'\s*\.section\s+\.text\.', l) or l.startswith('.text.'):
        in_fn = fn in l; continue
    if not in_fn: continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(srcfile) else -1; continue
    if re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', l): seq.append(cur)
raw = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split('\n')))
hdr = rows[1]; ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
data = [r for r in rows[2:] if len(r) > ia]
agg = collections.defaultdict(lambda: [0, 0]); seg = [[0, 0]]
for i in range(min(len(seq), len(data))):
    agg[seq[i]][0] += int(data[i][ia]); agg[seq[i]][1] += int(data[i][isamp])
    seg[-1][0] += int(data[i][ia]); seg[-1][1] += int(data[i][isamp])
    if 'BAR.SYNC' in data[i][isrc]: seg.append([0, 0])
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('sass', len(seq), 'ncu rows', len(data), 'total inst', tot, 'samples', ts)
print('segments between barriers (inst%, samples%):', ' | '.join('%.1f/%.1f' % (100 * a / tot, 100 * s / ts) for a, s in seg))
src = open(f'{root}/jieba_go_b200/csrc/{srcfile}').read().split('\n')
for ln, (a, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5s inst %5.1f%% samp %5.1f%%  %s' % (ln, 100 * a / tot, 100 * s / ts, src[ln - 1].strip()[:105] if ln and ln > 0 else '(other file)'))
if len(sys.argv) > 5:
    # sum over line ranges a-b,c-d,...
    for rng in sys.argv[5].split(','):
        a, b = map(int, rng.split('-'))
        ia_, is_ = 0, 0
        for ln, (x, y) in agg.items():
            if ln and a <= ln <= b: ia_ += x; is_ += y
        print('lines %d-%d: inst %.1f%% (%.0fM) samples %.1f%%' % (a, b, 100 * ia_ / tot, ia_ / 1e6, 100 * is_ / ts))
