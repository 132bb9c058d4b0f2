"""Turn the ncu CSV exports in this directory into the per-round markdown summary.
usage: python profiles/summarize.py r01"""
import collections
import csv
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
out = []

rows = list(csv.reader(open(f'profiles/{tag}_launches.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[h], rows[h + 1:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in data:
    if len(r) > vi:
        name = r[ki].split('(')[0].replace('void ', '')
        tot[name] += float(r[vi].replace(',', '')) / 1e6
        cnt[name] += 1
T = sum(tot.values())
out.append(f'## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, {sum(cnt.values())} launches, {T:.1f} ms)\n')
out.append('| kernel | launches | total ms | share |\n|---|---:|---:|---:|')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    out.append(f'| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / T:.2f} % |')

rows = list(csv.reader(open(f'profiles/{tag}_ncu_full_raw.csv')))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [('gpu__time_duration.sum', 'ms'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('smsp__warps_eligible.avg.per_cycle_active', 'eligible warps/cyc'),
        ('smsp__inst_executed.sum', 'warp insts'), ('dram__bytes_read.sum', 'DRAM rd GB'), ('dram__bytes_write.sum', 'DRAM wr GB')]
out.append('\n## `ncu --set full` captures (per launch)\n')
out.append('| kernel | ' + ' | '.join(c[1] for c in cols) + ' |\n|---|' + '---:|' * len(cols))
for r in data:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
    vals = []
    for m, _ in cols:
        i = hdr.index(m)
        v = r[i].replace(',', '')
        try:
            f = float(v)
            vals.append(f'{f:.3g}' if f < 1e6 else f'{f:.3e}')
        except ValueError:
            vals.append(v)
    out.append(f'| `{name}` | ' + ' | '.join(vals) + ' |')
print('\n'.join(out))
