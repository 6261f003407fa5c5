#!/usr/bin/env python
"""Vendor the UNMODIFIED reference into the git-ignored directory oracle/_ref/ so that it can travel to the GPU
box (which has no /root/reference) and be timed / compared there as the CPU arm.

    python tools/vendor_ref.py            # /root/reference -> oracle/_ref/

What is copied, byte for byte (a manifest with sha256 sums is written next to the files, and
`oracle/ref_harness.py` refuses to run when a file no longer matches it): the `tiger/` package, `init_utils.py`,
`train_utils.py`, `CHANGELOG.py` and the three driver scripts.  The only addition is the 12-line
`torch_scatter.scatter_max` shim of tests/golden/_shim (the reference imports torch_scatter, which is not installed
and not vendored by the reference: SURVEY.md §8(c)).  Nothing under oracle/_ref/ is committed (`.gitignore`), and
nothing under www2023tiger_b200/ imports it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_SRC = '/root/reference'
DST = os.path.join(ROOT, 'oracle', '_ref')
SHIM = os.path.join(ROOT, 'tests', 'golden', '_shim', 'torch_scatter')
TOP_LEVEL = ['init_utils.py', 'train_utils.py', 'CHANGELOG.py', 'train_self_supervised.py',
             'train_self_supervised_ddp.py', 'train_supervised.py']


def sha256(path: str) -> str:
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def vendor(src: str = DEFAULT_SRC, dst: str = DST) -> str:
    if not os.path.isdir(os.path.join(src, 'tiger')):
        raise FileNotFoundError(f'{src} does not hold the reference (no tiger/ package)')
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(os.path.join(src, 'tiger'), os.path.join(dst, 'tiger'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for name in TOP_LEVEL:
        if os.path.exists(os.path.join(src, name)):
            shutil.copy2(os.path.join(src, name), os.path.join(dst, name))
    shutil.copytree(SHIM, os.path.join(dst, 'torch_scatter'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    manifest = {}
    for base, _, files in os.walk(dst):
        for f in sorted(files):
            p = os.path.join(base, f)
            rel = os.path.relpath(p, dst)
            manifest[rel] = sha256(p)
    # every vendored reference file must equal its source
    for rel, digest in manifest.items():
        if rel.startswith('torch_scatter'):
            continue
        assert sha256(os.path.join(src, rel)) == digest, rel
    json.dump({'source': src, 'files': manifest}, open(os.path.join(dst, 'MANIFEST.json'), 'w'), indent=1)
    return dst


if __name__ == '__main__':
    out = vendor(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SRC)
    print(f'vendored the reference into {out} ({len(os.listdir(out))} entries)')
