"""Join an ncu SASS source-page CSV (per-instruction executed counts / stall samples) with nvdisasm line info
of the same kernel, and aggregate by CUDA source line.
usage: python profiles/sass_lines.py <ncu_source.csv> <object.o> <kernel-name-substring> [top]"""
import collections
import csv
import re
import subprocess
import sys

src_csv, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60

import glob, os, tempfile
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = None
for i, l in enumerate(dis):
    if l.startswith('.text.') and kname in l:
        start = i
        break
    if re.match(r'\s*\.section\s+\.text\.\S*' + re.escape(kname), l):
        start = i
        break
assert start is not None, 'kernel not found in disassembly'
lines = []          # (file:line) per instruction, in order
cur = '?'
for l in dis[start + 1:]:
    if l.startswith('.section') or re.match(r'\s*\.section', l):
        if lines:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f'{m.group(1).split("/")[-1]}:{m.group(2)}'
        # inlined-at chains: keep the outermost user line too
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)

rows = list(csv.reader(open(src_csv)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[h]
ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
data = rows[h + 1:]
print(f'{len(lines)} SASS instructions with line info, {len(data)} in the ncu page')
agg_i, agg_s = collections.Counter(), collections.Counter()
for k, r in enumerate(data):
    key = lines[k] if k < len(lines) else '?'
    agg_i[key] += int(r[ii])
    agg_s[key] += int(r[si])
ti, ts = sum(agg_i.values()), sum(agg_s.values())
print(f'total warp instructions {ti}, samples {ts}')
for key, n in agg_i.most_common(top):
    print(f'{100 * n / ti:6.2f}% inst  {100 * agg_s[key] / max(ts, 1):6.2f}% samples   {key}')
