"""Builds libtiger_b200.so in-tree with nvcc for sm_100a (no torch headers involved: the
library is a plain C-ABI shared object loaded through ctypes)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB_PATH = os.path.join(CSRC, 'libtiger_b200.so')

NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '--shared', '-I', INCLUDE, '-I', CSRC]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(INCLUDE, '*.h'))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    extra = os.environ.get('TIGER_EXTRA_NVCC_FLAGS', '').split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + sources() + ['-o', LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force=True, verbose='-v' in sys.argv))
