"""Per-source-line view of an ncu capture: python tools/ncu_lines.py <rep> <fn-substring> <file.cu> [top]
Maps SASS rows of the ncu source page to CUDA lines via nvdisasm line info; prints warp instructions,
average active threads, stall samples and L1 tag requests per line."""
import csv, re, subprocess, sys, os, collections, tempfile
rep, fn, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
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
        cur = int(m.group(2)) if m.group(1).endswith(srcfile) else -int(m.group(2)); continue
    if re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', l): seq.append(cur)
raw = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split('\n')))
hdr = rows[1]
ia, it, isamp, itag, isw = (hdr.index(x) if x in hdr else -1 for x in ('Instructions Executed', 'Thread Instructions Executed', '# Samples', 'L1 Tag Requests Global', 'L1 Wavefronts Shared'))
data = [r for r in rows[2:] if len(r) > ia]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for i in range(min(len(seq), len(data))):
    a = agg[seq[i]]
    for j, c in enumerate((ia, it, isamp, itag, isw)): a[j] += int(data[i][c] or 0) if c >= 0 else 0
tot = [sum(v[j] for v in agg.values()) for j in range(5)]
print('sass', len(seq), 'rows', len(data), 'warp-inst %d thread-inst %d (avg %.1f) samples %d L1tags %d smem-wf %d' % (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2], tot[3], tot[4]))
src = open(f'{root}/jieba_go_b200/csrc/{srcfile}').read().split('\n')
key = (lambda kv: -kv[1][0]) if os.environ.get('BY', 'inst') == 'inst' else (lambda kv: -kv[1][2])
for ln, a in sorted(agg.items(), key=key)[:top]:
    print('%5s inst %5.1f%% thr %4.1f samp %5.1f%% tags %5.1f%%  %s' % (ln, 100 * a[0] / tot[0], a[1] / max(1, a[0]), 100 * a[2] / max(1, tot[2]), 100 * a[3] / max(1, tot[3]),
          src[ln - 1].strip()[:100] if ln and ln > 0 else '(other file)'))
