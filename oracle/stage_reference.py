#!/usr/bin/env python
"""Stage the UNMODIFIED reference's Python sources into the git-ignored ``baseline/_ref/`` so that they travel to the
GPU box (``/root/reference`` does not exist there).

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under ``nerfool_b200/`` imports from ``baseline/_ref``.  It is used by
  * ``tests/test_reference_callers_gpu.py`` -- runs the reference's own ``optimize_adv_perturb`` (eval/ibrnet/eval_adv.py)
    and ``train.train`` (train.py) through ``dropin/`` on a synthetic dataset,
  * ``bench.py`` -- times the reference's cuDNN ``ResUNet`` (ibrnet/feature_network.py; out of scope for this repo, north
    star: "stays on cuDNN and is timed separately") and, for ``--impl reference`` / ``cpu_baseline``, the reference's own
    ``render_rays`` on the host cores (``kind: "reference"``).
Files are copied byte for byte (checked below); nothing is committed (``baseline/_ref/`` is in .gitignore)."""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('NERFOOL_REFERENCE_SRC', '/root/reference')
DST = os.path.join(REPO, 'baseline', '_ref')

# directories copied recursively (``*.py`` / ``*.txt`` only) and single files
TREES = ['ibrnet', 'gnt', 'eval/ibrnet', 'eval/gnt', 'configs']
FILES = ['utils.py', 'config.py', 'train.py', 'LICENSE']
KEEP_EXT = ('.py', '.txt')


def stage(verbose: bool = True) -> str | None:
    """Copy the sources; returns the staged root, or None when the reference checkout is not present."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    n = 0
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d != '__pycache__']
            rel = os.path.relpath(root, SRC)
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                s, d = os.path.join(root, f), os.path.join(DST, rel, f)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                n += 1
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.exists(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    if verbose:
        print(f'[stage_reference] {n} files staged byte-for-byte from {SRC} into {DST}')
    return DST


def staged_root() -> str | None:
    """Where the reference can be imported from in this process: the live checkout if present, else the staged copy."""
    if os.path.isdir(os.path.join(SRC, 'ibrnet')):
        return SRC
    if os.path.isdir(os.path.join(DST, 'ibrnet')):
        return DST
    return None


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
