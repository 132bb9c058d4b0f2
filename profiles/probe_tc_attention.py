"""Compile tests/probes/probe_tc_attention_epilogue.cu and count the SASS instructions of the softmax epilogue per score.
usage: python profiles/probe_tc_attention.py > profiles/r02_probe_tc_attention.txt"""
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, '..', 'tests', 'probes', 'probe_tc_attention_epilogue.cu')
with tempfile.TemporaryDirectory() as d:
    cubin = os.path.join(d, 'probe.cubin')
    subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-cubin', '-o', cubin, SRC], check=True)
    sass = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True, check=True).stdout
name, counts, ops = None, {}, {}
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = 'SPLIT (bf16 hi + lo P, fp32-equivalent)' if 'Lb1' in m.group(1) else 'plain bf16 P'
        counts[name], ops[name] = 0, {}
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1)
        if op in ('NOP', 'BRA', 'EXIT'):
            continue
        counts[name] += 1
        ops[name][op.split('.')[0]] = ops[name].get(op.split('.')[0], 0) + 1
print('softmax epilogue of a tensor-core attention, one head, 128 keys per query (tests/probes/probe_tc_attention_epilogue.cu), SASS counts:')
for k, v in counts.items():
    top = ', '.join(f'{o} {n}' for o, n in sorted(ops[k].items(), key=lambda kv: -kv[1])[:10])
    print(f'  {k}: {v} instructions per query and head = {v / 128:.2f} per (query, key, head)   [{top}]')
print('CUDA-core attention of the production kernels, for comparison: IBRNet ray stage (d_k = 4) 28 instructions per (query, key) for all four')
print('heads = 7.0 per (query, key, head) INCLUDING the dot products and P.V (profiles/r02j_lines_ray_fwd.txt); GNT ray core (d = 16) ~20.')
