"""Static SASS opcode histogram of the hot kernels in the built objects (cuobjdump -sass), trimmed to the mnemonics that matter
for the design claims: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: TMEM operands), UTCBAR / SYNCS (commit / mbarrier),
RED / ATOMG (scatter atomics), FFMA2 / FADD2 / FMUL2 (packed fp32), LDS / STS / LDG / STG widths, MUFU.
usage: python profiles/sass_hist.py > profiles/r02_sass_histogram.md"""
import collections
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '..', 'nerfool_b200', 'csrc', 'build')
WANT = {'k_view_tc_fwd<3,1,1>': ('nfb_view_tc_bwd2_inst1.o', 'k_view_tc_fwdILi3ELb1ELb1'),
        'k_view_tc_fwd<3,1,0>': ('nfb_view_tc_inst1.o', 'k_view_tc_fwdILi3ELb1ELb0'),
        'k_view_tc_bwd_stash<3>': ('nfb_view_tc_bwd2_inst3.o', 'k_view_tc_bwd_stashILi3'),
        'k_ray_tc<3,0,1>': ('nfb_ray_tc_inst5.o', 'k_ray_tcILi3ELb0ELb1'),
        'k_ray_tc_bwd_stash<3>': ('nfb_ray_tc_inst7.o', 'k_ray_tc_bwd_stashILi3'),
        'geometry / compositing / sampling / warp (nfb_geom.o + nfb_warp.o, all kernels)': ('nfb_geom.o', None)}
KEEP = ['UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'UTMALDG', 'UTMASTG', 'RED', 'ATOMG', 'ATOMS', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FMUL', 'FADD',
        'FMNMX', 'MUFU', 'F2FP', 'LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'SHFL', 'BAR', 'DFMA', 'DADD', 'R2UR', 'HMMA']


def hist(obj, sym):
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    if obj == 'nfb_geom.o':
        out += subprocess.run(['cuobjdump', '-sass', os.path.join(BUILD, 'nfb_warp.o')], capture_output=True, text=True).stdout
    c, n = collections.Counter(), 0
    on = sym is None
    for line in out.split('\n'):
        if 'Function :' in line:
            on = sym is None or sym in line
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', line)
        if on and m:
            op, suf = m.group(1), m.group(2)
            n += 1
            key = op
            if op in ('LDS', 'STS', 'LDG', 'STG', 'RED', 'ATOMG') and ('.128' in suf or '.64' in suf):
                key = op + ('.128' if '.128' in suf else '.64')
            c[key] += 1
    return n, c


print('# SASS opcode histogram of the hot kernels (static counts, `cuobjdump -sass` of the in-tree build; sm_100a)\n')
print('`UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / tcgen05.st (TMEM), `SYNCS` = mbarrier, `RED.128` = vector float atomics of\n'
      'the scatter, `FFMA2` / `FADD2` / `FMUL2` = packed fp32.  No `UTMALDG` / `UTMASTG` (TMA): operands are staged by the threads that own\n'
      'the rows (north star allows shared-memory staging); no `HMMA` (legacy mma.sync).\n')
for name, (obj, sym) in WANT.items():
    n, c = hist(obj, sym)
    print(f'## `{name}`  ({n} instructions)\n')
    print('| ' + ' | '.join(k for k in sorted(c, key=lambda k: -c[k]) if any(k.startswith(w) for w in KEEP)) + ' |')
    print('|' + '---:|' * sum(1 for k in c if any(k.startswith(w) for w in KEEP)))
    print('| ' + ' | '.join(str(c[k]) for k in sorted(c, key=lambda k: -c[k]) if any(k.startswith(w) for w in KEEP)) + ' |\n')
