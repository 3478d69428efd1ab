"""Build libplfb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pylabfea_b200.build [--force] [--verbose]

Three kernel translation units + two host ones.  assembly_kernels.cu is compiled with -fmad=false: its results must be
bit-identical to numpy's (only the explicit fma() calls may fuse).  Host code gets -ffp-contract=off for the same
reason (B matrices).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libplfb200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

COMMON = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
          '-Xcompiler', '-fPIC,-ffp-contract=off', '-I', os.path.join(HERE, '..', 'include')]
UNITS = {
    'response_kernels.cu': [],
    'model_kernels.cu': [],
    'mg_kernels.cu': [],
    'assembly_kernels.cu': ['-fmad=false'],
    'plf_api.cu': [],
    'plf_model.cu': [],
    'plf_solve.cu': [],
    'plf_comm.cu': [],
}


def _newest_src():
    t = 0.
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cu', '.cuh', '.h')):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_src():
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def cc(item):
        src, extra = item
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        cmd = [NVCC] + COMMON + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(cc, UNITS.items()))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
