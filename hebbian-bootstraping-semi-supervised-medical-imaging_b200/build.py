"""Build libhebb_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python build.py [--force] [-v]

The shared library lands in hebb/libhebb_sm100.so next to the Python drop-in, so it
travels to the GPU box with the source snapshot.  Each translation unit is compiled to an
object file under build/ (in parallel, only when it or a header changed) and the objects
are linked into the shared library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = ['api.cu', 'elementwise.cu', 'simt_path.cu', 'tc_path.cu', 'fused_path.cu', 'fixup.cu', 'norm_act.cu']
OUT = os.path.join(HERE, 'hebb', 'libhebb_sm100.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default',
         '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(HERE, 'csrc')]


def _headers():
    d = os.path.join(HERE, 'csrc')
    hs = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(ROOT, 'include', 'hebb_sm100.h'))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    srcs = [os.path.join(HERE, 'csrc', f) for f in SRC]
    return _stale(OUT, srcs + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()

    def compile_one(f):
        src = os.path.join(HERE, 'csrc', f)
        obj = os.path.join(OBJ, f[:-3] + '.o')
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            print(' '.join(cmd), flush=True)
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SRC), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SRC))
    cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '--shared', '-Xcompiler', '-fPIC'] + objs + ['-o', OUT, '-lcudart']
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print('built', OUT)
