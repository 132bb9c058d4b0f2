"""In-tree build of libnerfool_b200.so (sm_100a only): one nvcc invocation per translation unit, run in
parallel, then one link.  Usage: ``python -m nerfool_b200.build [--force]``.  The resulting .so sits next
to this file so it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# experiment hooks: NFB_BUILD_TAG=<tag> builds csrc/build_<tag>/ -> libnerfool_b200_<tag>.so with NFB_EXTRA_DEFS (e.g. "-DNFB_VTC_NG=3");
# NFB_LIB_PATH=<that .so> makes _lib.py load it.  The default build is untouched.
_TAG = os.environ.get('NFB_BUILD_TAG', '')
OBJ = os.path.join(CSRC, 'build' + ('_' + _TAG if _TAG else ''))
LIB = os.path.join(HERE, 'libnerfool_b200' + ('_' + _TAG if _TAG else '') + '.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v'] + os.environ.get('NFB_EXTRA_DEFS', '').split()

# (object name, source, extra defines)
UNITS = [
    ('nfb_api', 'nfb_api.cu', []),
    ('nfb_geom', 'nfb_geom.cu', []),
    ('nfb_warp', 'nfb_warp.cu', []),
    ('nfb_ray_stage', 'nfb_ray_stage.cu', []),
    ('nfb_gnt', 'nfb_gnt.cu', []),
    ('nfb_gnt_bwd', 'nfb_gnt_bwd.cu', []),
    ('nfb_view_api', 'nfb_view_api.cu', []),
    ('nfb_view_inst0', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=0']),
    ('nfb_view_inst1', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=1']),
    ('nfb_view_inst2', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=2']),
    ('nfb_view_inst3', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=3']),
    ('nfb_view_inst4', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=4']),
    ('nfb_view_inst5', 'nfb_view_inst.cu', ['-DNFB_VIEW_INST=5']),
    ('nfb_view_tc_inst0', 'nfb_view_tc_inst.cu', ['-DNFB_VTC_INST=0']),
    ('nfb_view_tc_inst1', 'nfb_view_tc_inst.cu', ['-DNFB_VTC_INST=1']),
    ('nfb_view_tc_inst2', 'nfb_view_tc_inst.cu', ['-DNFB_VTC_INST=2']),
    ('nfb_view_tc_inst3', 'nfb_view_tc_inst.cu', ['-DNFB_VTC_INST=3']),
    ('nfb_view_tc_bwd_inst0', 'nfb_view_tc_bwd_inst.cu', ['-DNFB_VTCB_INST=0']),
    ('nfb_view_tc_bwd_inst1', 'nfb_view_tc_bwd_inst.cu', ['-DNFB_VTCB_INST=1']),
    ('nfb_view_tc_bwd_inst2', 'nfb_view_tc_bwd_inst.cu', ['-DNFB_VTCB_INST=2']),
    ('nfb_view_tc_bwd_inst3', 'nfb_view_tc_bwd_inst.cu', ['-DNFB_VTCB_INST=3']),
    ('nfb_view_tc_bwd2_inst0', 'nfb_view_tc_bwd2_inst.cu', ['-DNFB_VTCS_INST=0']),
    ('nfb_view_tc_bwd2_inst1', 'nfb_view_tc_bwd2_inst.cu', ['-DNFB_VTCS_INST=1']),
    ('nfb_view_tc_bwd2_inst2', 'nfb_view_tc_bwd2_inst.cu', ['-DNFB_VTCS_INST=2']),
    ('nfb_view_tc_bwd2_inst3', 'nfb_view_tc_bwd2_inst.cu', ['-DNFB_VTCS_INST=3']),
    ('nfb_ray_tc_inst0', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=0']),
    ('nfb_ray_tc_inst1', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=1']),
    ('nfb_ray_tc_inst2', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=2']),
    ('nfb_ray_tc_inst3', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=3']),
    ('nfb_ray_tc_inst4', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=4']),
    ('nfb_ray_tc_inst5', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=5']),
    ('nfb_ray_tc_inst6', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=6']),
    ('nfb_ray_tc_inst7', 'nfb_ray_tc_inst.cu', ['-DNFB_RTC_INST=7']),
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


_INC_RE = None


def _deps(path, seen):
    """Transitive closure of the quoted #includes of one source file (paths relative to the includer)."""
    global _INC_RE
    import re
    if _INC_RE is None:
        _INC_RE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    with open(path) as f:
        text = f.read()
    for inc in _INC_RE.findall(text):
        _deps(os.path.join(os.path.dirname(path), inc), seen)


def _sources_digest(src, extra):
    """Digest of one unit's inputs: its source and the headers it (transitively) includes."""
    seen = set()
    _deps(os.path.join(CSRC, src), seen)
    h = hashlib.sha256()
    for n in sorted(seen):
        with open(n, 'rb') as f:
            h.update(os.path.basename(n).encode() + b'\0' + f.read())
    h.update(' '.join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()


_EXTRA = os.environ.get('NFB_EXTRA_DEFS', '').split()
_BASE_FLAGS = [f for f in NVCC_FLAGS if f not in _EXTRA]


def _affected(src):
    """Does any -D macro of NFB_EXTRA_DEFS occur in the unit's sources?  (experiment builds reuse the default objects otherwise)"""
    names = [d[2:].split('=')[0] for d in _EXTRA if d.startswith('-D')]
    if not _TAG or not names or len(names) != len(_EXTRA):
        return True
    seen = set()
    _deps(os.path.join(CSRC, src), seen)
    for n in seen:
        with open(n) as f:
            text = f.read()
        if any(nm in text for nm in names):
            return True
    return False


def _compile(unit):
    name, src, defs = unit
    if not _affected(src):
        base = os.path.join(CSRC, 'build', name + '.o')
        if os.path.exists(base):
            import shutil
            shutil.copyfile(base, os.path.join(OBJ, name + '.o'))
            return name, 'reused', ''
    obj = os.path.join(OBJ, name + '.o')
    stamp = obj + '.sha'
    digest = _sources_digest(src, defs)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return name, 'cached', ''
    cmd = [_nvcc()] + NVCC_FLAGS + defs + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(OBJ, name + '.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {name}:\n{log[-4000:]}')
    with open(stamp, 'w') as f:
        f.write(digest)
    return name, 'built', log


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith('.sha'):
                os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile, UNITS))
    rebuilt = [n for n, st, _ in results if st in ('built', 'reused')]
    if verbose:
        for n, st, _ in results:
            print(f'[nerfool_b200.build] {n}: {st}')
    if rebuilt or not os.path.exists(LIB):
        objs = [os.path.join(OBJ, n + '.o') for n, _, _ in UNITS]
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
        if verbose:
            print('[nerfool_b200.build] linked', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
