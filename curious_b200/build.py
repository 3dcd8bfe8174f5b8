"""Build libcurious_b200.so (the C-ABI library of include/curious_b200.h) in-tree with nvcc.

    python -m curious_b200.build            # build if sources are newer than the library
    python -m curious_b200.build --force

sm_100a only: `-gencode arch=compute_100a,code=sm_100a`.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcurious_b200.so')
SOURCES = ['her.cu', 'norm_adam.cu', 'ddpg.cu', 'ddpg_rows.cu', 'p2p.cu', 'tc_gemm.cu', 'tc_chain.cu']
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'mlp_kernels.cuh'),
           os.path.join(CSRC, 'net_layout.cuh'), os.path.join(CSRC, 'her_device.cuh'),
           os.path.join(CSRC, 'tc_gemm.cuh'), os.path.join(CSRC, 'tc_ptx.cuh'),
           os.path.join(os.path.dirname(HERE), 'include', 'curious_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '--expt-relaxed-constexpr', '--extended-lambda', '-Xcompiler', '-fPIC',
    # no --use_fast_math: parity needs IEEE division / sqrt and no implicit FMA contraction of the
    # explicitly rounded intrinsics
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    build_dir = os.path.join(HERE, 'build')
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(build_dir, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs   # static cudart (nvcc default): self-contained
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
