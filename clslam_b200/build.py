"""Build the C-ABI library (clslam_b200/libclslam_b200.so) for sm_100a with nvcc, in-tree.

``python -m clslam_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without
a GPU.  Object files are cached under ``build/`` keyed on source + header mtimes.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libclslam_b200.so'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
          f'-I{ROOT / "include"}', f'-I{CSRC}']


def _sources():
    return sorted(CSRC.glob('*.cu'))


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.inc')) + list((ROOT / 'include').glob('*.h'))
    return max(h.stat().st_mtime for h in hdrs)


def build_library(force: bool = False, verbose: bool = True) -> Path:
    objdir = ROOT / 'build' / 'obj'
    objdir.mkdir(parents=True, exist_ok=True)
    hdr_m = _newest_header_mtime()
    jobs = []
    objs = []
    for src in _sources():
        obj = objdir / (src.stem + '.o')
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *ARCH_FLAGS, *CFLAGS, '-c', str(src), '-o', str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}')
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f'[clslam_b200.build] compiled {name}', file=sys.stderr)
    if jobs or not LIB.exists():
        cmd = [NVCC, *ARCH_FLAGS, '-shared', '-o', str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(f'[clslam_b200.build] linked {LIB}', file=sys.stderr)
    return LIB


if __name__ == '__main__':
    build_library(force='--force' in sys.argv)
