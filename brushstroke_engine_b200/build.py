"""Build libnbe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m brushstroke_engine_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libnbe_b200.so')
SOURCES = ['api.cu', 'bias_act.cu', 'upfirdn2d.cu', 'conv_f32.cu', 'small_ops.cu', 'canvas.cu', 'conv_tc.cu', 'encoder.cu', 'conv_flat.cu', 'fir_nhwc.cu', 'modconv_op.cu', 'up_fused.cu', 'enc7x7_toeplitz.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc() -> str:
    cand = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'bin', 'nvcc')
    return cand if os.path.exists(cand) else 'nvcc'


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'nbe_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {src} ---\n{out}\n')
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed building libnbe_b200.so')
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
