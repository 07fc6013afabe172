"""Builds the engine's shared library in-tree with nvcc for sm_100a.

    python -m hierarchical_sparse_coding_b200.build

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libhsc_b200.so')
SOURCES = ['engine.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


FLAGS_NOTE = os.path.join(HERE, 'build_flags.txt')


def _fast():
    """HSC_FAST_BUILD=1: development builds with `-split-compile 0` (30 s instead of 70 s; ptxas then allocates registers
    per function in parallel and the hot kernels come out with more spills, so such a library is never shipped: the
    next build without the switch replaces it)."""
    return os.environ.get('HSC_FAST_BUILD', '') not in ('', '0')


def _stale():
    if not os.path.exists(LIB):
        return True
    note = open(FLAGS_NOTE).read().strip() if os.path.exists(FLAGS_NOTE) else ''
    if note == 'fast' and not _fast():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'hsc_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    """Compiles csrc/*.cu -> libhsc_b200.so if any source is newer.  Returns the library path."""
    if out is None and not force and not _stale():
        return LIB
    target = out or LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-split-compile', '0'] if _fast() else []) + ['-D' + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', target + '.tmp', '-lcudart']
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed (%d)' % proc.returncode)
    with open(os.path.join(HERE, 'build_ptxas.log'), 'w') as f:
        f.write(proc.stdout)
    os.replace(target + '.tmp', target)
    if target == LIB:
        with open(FLAGS_NOTE, 'w') as f:
            f.write('fast\n' if _fast() else 'release\n')
    return target


if __name__ == '__main__':
    if '--profile-phases' in sys.argv:
        print(build_library(force=True, verbose=False, defines=('HSC_PROFILE_PHASES',), out=os.path.join(HERE, 'libhsc_b200_prof.so')))
    else:
        print(build_library(force='--force' in sys.argv, verbose=True))
