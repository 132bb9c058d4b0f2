"""Join an ncu SASS source-page CSV (`ncu -i x.ncu-rep --page source --csv --print-source sass`, one block per
captured launch) with the nvdisasm line info of the same kernel, and aggregate instruction counts, stall samples and
stall reasons by CUDA source line.
usage: python profiles/sass_lines.py <ncu_source.csv> <object.o> <kernel-name-substring> [occurrence=0] [top=50] [csv-kernel-substring]
(the object's section names are mangled, the CSV's kernel names demangled: pass the demangled form as the last argument when a
plain substring does not select one instantiation in both)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

src_csv, obj, kname = sys.argv[1:4]
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
csv_kname = sys.argv[6] if len(sys.argv) > 6 else kname

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
start = None
for i, l in enumerate(dis):
    if re.match(r'\s*\.section\s+\.text\.\S*' + re.escape(kname), l) or (l.startswith('.text.') and kname in l):
        start = i
        break
assert start is not None, 'kernel not found in disassembly'
lines = []          # innermost "file:line" per instruction, in order
cur = '?'
for l in dis[start + 1:]:
    if re.match(r'\s*\.section', l) and lines:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f'{m.group(1).split("/")[-1]}:{m.group(2)}'
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)

csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(src_csv)))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and csv_kname in r[1] and not (csv_kname + '_') in r[1]]
if len(blocks) <= occ:
    print('kernels in the CSV:', [r[1][:80] for r in rows if r and r[0] == 'Kernel Name'])
b = blocks[occ]
hdr = rows[b + 1]
end = next((i for i in range(b + 2, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
data = [r for r in rows[b + 2:end] if len(r) == len(hdr)]
ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
print(f'{rows[b][1]}: {len(lines)} SASS instructions with line info, {len(data)} in the ncu page')
agg_i, agg_s = collections.Counter(), collections.Counter()
reasons = collections.Counter()
line_reason = collections.defaultdict(collections.Counter)
ops = collections.Counter()
for k, r in enumerate(data):
    key = lines[k] if k < len(lines) else '?'
    n = int(r[ii])
    agg_i[key] += n
    agg_s[key] += int(r[si])
    ops[r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]] += n
    for i, name in stall_cols:
        v = int(r[i])
        if v:
            reasons[name] += v
            line_reason[key][name] += v
ti, ts = sum(agg_i.values()), sum(agg_s.values())
print(f'total warp instructions {ti}, samples {ts}')
print('stall reasons:', ', '.join(f'{k[6:]} {100 * v / max(ts, 1):.1f}%' for k, v in reasons.most_common(10)))
print('top opcodes:', ', '.join(f'{k} {100 * v / ti:.1f}%' for k, v in ops.most_common(24)))
for key, n in sorted(agg_s.items(), key=lambda kv: -kv[1])[:top]:
    why = ', '.join(f'{k[6:]} {v}' for k, v in line_reason[key].most_common(3))
    print(f'{100 * n / max(ts, 1):6.2f}% samples {100 * agg_i[key] / ti:6.2f}% inst   {key:28s} {why}')
