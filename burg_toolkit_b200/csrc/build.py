"""Build libburgb200.so in-tree with nvcc for sm_100a (no other architecture, no JIT cache).

    python burg_toolkit_b200/csrc/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['mesh.cu', 'ags.cu', 'collide.cu', 'coverage.cu', 'pipeline.cu', 'generators.cu', 'scene.cu', 'peer.cu', 'poisson.cu']
HEADERS = ['common.cuh', 'tritri.cuh', os.path.join('..', '..', 'include', 'burg_b200.h')]
LIB = os.path.join(HERE, 'libburgb200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-t', '0', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
    # float64 paths must reproduce numpy's separate multiply/add roundings; wanted FMAs are explicit
    '-fmad=false',
    '--shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-Xcompiler', '-fvisibility=hidden',
    '-Xcompiler', '-Wall', '-Xcudafe', '--diag_suppress=177',
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """``out`` / ``defines``: build a tuning variant next to the product library (A/B runs: BT_LIB=<path> selects it)"""
    lib = out or LIB
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(HERE, s) for s in SOURCES] + ['-o', lib + '.tmp']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libburgb200.so')
    if verbose:
        sys.stderr.write(res.stderr)
    os.replace(lib + '.tmp', lib)
    return lib


if __name__ == '__main__':
    out = next((a.split('=', 1)[1] for a in sys.argv if a.startswith('--out=')), None)
    defs = [a[2:] for a in sys.argv if a.startswith('-D')]
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, out=out, defines=defs))
