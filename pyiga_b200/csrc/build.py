"""Builds libpyiga_b200.so (sm_100a) in-tree with nvcc.

    python -m pyiga_b200.csrc.build [--force]

One translation unit per (degree, nodes-per-span) pair of the walk kernels, compiled in parallel;
objects are cached by source hash under csrc/_obj/.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libpyiga_b200.so')
OBJ = os.path.join(HERE, '_obj')

# (degree p, Gauss nodes per span q) pairs with a sum-factorised instantiation.  q = p+1 is the
# reference's rule for equal degrees; the other pairs cover mixed-degree spaces (q = max p + 1).
PQ = [(1, 2), (2, 3), (3, 4), (4, 5), (1, 3), (2, 4), (3, 5), (1, 4), (2, 5)]

# pairs instantiated for linear forms only: the stand-in normal axis of boundary integrals (degree 1,
# one node) and its companions
PQ1 = [(1, 1)]

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
         '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _deps_hash(extra=''):
    h = hashlib.sha1()
    for name in sorted(os.listdir(HERE)):
        if name.endswith(('.cu', '.cuh')):
            with open(os.path.join(HERE, name), 'rb') as f:
                h.update(name.encode() + f.read())
    with open(os.path.join(os.path.dirname(PKG), 'include', 'pyiga_b200.h'), 'rb') as f:
        h.update(f.read())
    h.update((' '.join(FLAGS) + extra).encode())
    return h.hexdigest()[:16]


def _compile(job):
    src, obj, defs = job
    if os.path.exists(obj):
        return obj
    cmd = [NVCC] + FLAGS + defs + ['-c', os.path.join(HERE, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), r.stdout + r.stderr))
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    tag = _deps_hash()
    stamp = LIB + '.stamp'         # next to the library: the object cache does not travel to the GPU box
    if (not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == tag):
        return LIB
    jobs = [('api.cu', os.path.join(OBJ, 'api_%s.o' % tag), [])]
    for p, q in PQ:
        jobs.append(('walk_inst.cu', os.path.join(OBJ, 'walk_%d_%d_%s.o' % (p, q, tag)),
                     ['-DPB_P=%d' % p, '-DPB_Q=%d' % q]))
    for p, q in PQ1:
        jobs.append(('walk_inst.cu', os.path.join(OBJ, 'walk1_%d_%d_%s.o' % (p, q, tag)),
                     ['-DPB_P=%d' % p, '-DPB_Q=%d' % q, '-DPB_WALK1_ONLY']))
    if force:
        for _, o, _ in jobs:
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(_compile, jobs))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed: %s\n%s' % (' '.join(cmd), r.stdout + r.stderr))
    # drop stale objects
    for name in os.listdir(OBJ):
        if name.endswith('.o') and tag not in name:
            os.remove(os.path.join(OBJ, name))
    with open(stamp, 'w') as f:
        f.write(tag)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
