"""Build libhebb_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python build.py [--force]

The shared library lands in hebb/libhebb_sm100.so next to the Python drop-in, so it
travels to the GPU box with the source snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = ['api.cu', 'elementwise.cu', 'simt_path.cu', 'tc_path.cu', 'norm_act.cu']
OUT = os.path.join(HERE, 'hebb', 'libhebb_sm100.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default', '--shared',
         '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(HERE, 'csrc')]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))]
    deps.append(os.path.join(ROOT, 'include', 'hebb_sm100.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(HERE, 'csrc', f) for f in SRC] + ['-o', OUT, '-lcudart']
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print('built', OUT)
