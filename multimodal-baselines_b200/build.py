"""Build libmmb_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object, see include/mmb_b200.h).

    python multimodal-baselines_b200/build.py [--force]

The .so lands next to this file; it is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libmmb_b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
         '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; libmmb_b200.so cannot be built')
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, f), 'rb') as fh:
                    h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libmmb_b200.so. Returns the library path."""
    stamp_file = os.path.join(OBJ, 'stamp')
    stamp = _stamp()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + ARCH + FLAGS + extra + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for src, obj, r in ex.map(compile_one, _sources()):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError('nvcc failed on ' + src)
            if verbose:
                sys.stderr.write(r.stderr)
            objs.append(obj)
    cmd = [nvcc] + ARCH + ['-shared', '-o', LIB] + objs + ['-Xcompiler', '-fPIC']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(stamp_file, 'w') as fh:
        fh.write(stamp)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
