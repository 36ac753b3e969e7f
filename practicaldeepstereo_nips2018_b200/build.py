"""In-tree build of the sm_100a kernel library (csrc/*.cu -> libpds_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to
the GPU box with the repo snapshot.  Run as ``python -m
practicaldeepstereo_nips2018_b200.build`` or through ``__graft_entry__.build()``.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libpds_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '--use_fast_math=false', '-Xcompiler', '-fPIC,-O3', '--expt-relaxed-constexpr',
    '-Xptxas', '-v',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the CUDA kernel library cannot be built')


def _newest_header():
    hdrs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, obj, log):
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + ['-c', src, '-o', obj]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{proc.stdout[-4000:]}')
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdr_time = _newest_header()
    jobs, objs = [], []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time))
        if stale:
            jobs.append((src, obj, obj[:-2] + '.log'))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for fut in [pool.submit(_compile, *j) for j in jobs]:
                done = fut.result()
                if verbose:
                    print('compiled', os.path.basename(done))
    if jobs or not os.path.exists(LIB) or force:
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-lcudart']
        subprocess.check_call(cmd)
        if verbose:
            print('linked', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
