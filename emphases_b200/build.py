"""Build libemphases_b200.so in-tree with nvcc for sm_100a.

    python -m emphases_b200.build [--force]

The shared library is a plain C ABI (include/emphases_b200.h); it is loaded
with ctypes by emphases_b200/_lib.py.  It travels to the GPU box inside the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libemphases_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xptxas=-v',
    '-Xcompiler', '-fPIC',
    '-shared',
]


def sources():
    return sorted(
        os.path.join(CSRC, name) for name in os.listdir(CSRC)
        if name.endswith('.cu'))


def stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = sources() + [
        os.path.join(CSRC, name) for name in os.listdir(CSRC)
        if name.endswith('.cuh')]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'emphases_b200.h'))
    return any(os.path.getmtime(path) > built for path in deps)


def build(force=False, verbose=False, output=None, extra_flags=()):
    """`output` / `extra_flags` build a tuning variant next to the main lib"""
    if output is None and not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    output = output or LIB
    command = [nvcc] + NVCC_FLAGS + list(extra_flags) + ['-o', output] + sources()
    result = subprocess.run(command, capture_output=True, text=True)
    if verbose or result.returncode != 0:
        sys.stderr.write(result.stdout + result.stderr)
    if result.returncode != 0:
        raise RuntimeError('nvcc failed building libemphases_b200.so')
    return output


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(LIB)
