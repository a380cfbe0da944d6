"""Build libpfann_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m pfann_b200.build [--force] [--verbose]

The shared library is git-ignored but travels with the gpurun snapshot.  cudart is linked statically;
the driver API (cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint, so
the library has no link-time dependency on libcuda.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(CSRC, 'libpfann_b200.so')
SOURCES = ['ctx.cu', 'mel.cu', 'encoder.cu', 'encoder_tc.cu', 'front_tc.cu', 'extract.cu', 'knn.cu', 'knn_tc.cu', 'seqscore.cu', 'ingest.cu', 'train.cu', 'encoder_train.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-I', INCLUDE, '-I', CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(INCLUDE, 'pfann_b200.h'))
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s[:-3] + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError('nvcc failed for ' + cmd[-3])
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                     '-Xcompiler', '-fPIC', '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
