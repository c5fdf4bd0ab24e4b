"""dev tool: per-source-line instruction counts and stall samples of one kernel from an ncu report (SASS page) + nvdisasm line info.
usage: dev_ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import csv, subprocess, sys, re, os, tempfile, collections
rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null' % (tmp, os.path.abspath(obj)), shell=True, check=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# instructions of the wanted function with their source line
lines, cur, infn = [], None, False
for l in sass:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        infn = kname in m.group(1)
        continue
    if re.match(r'\s*\.section', l):
        infn = False
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ci = {n: j for j, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) >= len(h)]
print('ncu instructions %d, nvdisasm instructions %d' % (len(body), len(lines)))
n = min(len(body), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
ti = ts = 0
for i in range(n):
    inst = int(body[i][ci['Instructions Executed']]); samp = int(body[i][ci['# Samples']])
    a = agg[lines[i][0]]; a[0] += inst; a[1] += samp; a[2] += 1
    ti += inst; ts += samp
print('total warp instructions %d, samples %d' % (ti, ts))
src = {}
for (key, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ''
    if key:
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), '..', 'csrc', key[0])
        if key[0] not in src and os.path.isfile(path):
            src[key[0]] = open(path).read().splitlines()
        if key[0] in src and key[1] <= len(src[key[0]]):
            text = src[key[0]][key[1] - 1].strip()[:110]
    print('%-22s inst %5.1f%% samp %5.1f%% (%3d sass)  %s' % ('%s:%d' % key if key else '?', 100. * v[0] / ti, 100. * v[1] / max(ts, 1), v[2], text))
