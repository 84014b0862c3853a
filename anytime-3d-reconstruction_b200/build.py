"""In-tree build of liba3d.so (sm_100a only).  nvcc cross-compiles without a GPU.

    python anytime-3d-reconstruction_b200/build.py [--force] [--verbose] [--checked]

--checked builds liba3d_checked.so: the decoder translation units recompiled with -DA3D_CHECKED (device-side range checks
of every hand-computed index, see csrc/internal.h) and linked with the release objects of the encoders.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', '_obj')
OBJ_CHECKED = os.path.join(HERE, 'csrc', '_obj_checked')
LIB = os.path.join(HERE, 'liba3d.so')
LIB_CHECKED = os.path.join(HERE, 'liba3d_checked.so')
# translation units that carry A3D_DEV_CHECK range checks (recompiled for the checked build)
CHECKED_SOURCES = ['handle.cu', 'convt_tc.cu', 'convt_l4_sw.cu', 'gemm_l1.cu', 'simt_layers.cu', 'tail.cu', 'tail_hcol.cu',
                   'aux_kernels.cu']
SOURCES = ['handle.cu', 'convt_tc.cu', 'convt_l4_sw.cu', 'gemm_l1.cu', 'simt_layers.cu', 'tail.cu', 'tail_hcol.cu', 'aux_kernels.cu',
           'conv2d_tc.cu', 'conv2d_pair.cu', 'conv2d_first_tc.cu', 'enc2d_kernels.cu', 'enc2d.cu',
           'conv3d_tc.cu', 'enc3d.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr']


def _nvcc() -> str:
    nv = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nv):
        raise RuntimeError('nvcc not found; liba3d cannot be built')
    return nv


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    if checked:
        build(force=False, verbose=verbose)      # the release objects of the encoders are linked into the checked library
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(OBJ_CHECKED, exist_ok=True)
    nv = _nvcc()
    lib = LIB_CHECKED if checked else LIB
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(HERE, '..', 'include', 'a3d.h'))
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        chk = checked and src in CHECKED_SOURCES
        o = os.path.join(OBJ_CHECKED if chk else OBJ, src.replace('.cu', '.o'))
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            cmd = [nv] + ARCH + FLAGS + (['-DA3D_CHECKED'] if chk else []) + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed for ' + cmd[-3])
    if force or jobs or not os.path.exists(lib) or not _newer(lib, objs):
        cmd = [nv] + ARCH + ['-shared', '-o', lib] + objs + ['-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, checked='--checked' in sys.argv))
