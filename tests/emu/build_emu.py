"""Builds the sequential host emulation of the device code (TEST INFRASTRUCTURE ONLY).

The same sources as libpyiga_b200.so are compiled with g++ and -DPB_EMULATE: every kernel body
runs as a plain loop over its thread index and "device" pointers are host pointers.  It lets the
CPU test-suite exercise the index logic of the kernels (band tables, sliding windows, slabs, CSR
export) without a GPU.  The pyiga_b200 package never loads this library.
"""
import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'pyiga_b200', 'csrc')
OUT = os.path.join(HERE, '_build')
LIB = os.path.join(OUT, 'libpyiga_b200_emu.so')
PQ = [(1, 2), (2, 3), (3, 4), (4, 5), (1, 3), (2, 4), (3, 5), (1, 4), (2, 5)]
CXX = '/usr/bin/g++'
FLAGS = ['-std=c++17', '-O2', '-fPIC', '-DPB_EMULATE', '-x', 'c++', '-fvisibility=hidden', '-ffp-contract=off']


def _hash():
    h = hashlib.sha1()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith(('.cu', '.cuh')):
            with open(os.path.join(CSRC, name), 'rb') as f:
                h.update(name.encode() + f.read())
    with open(os.path.join(ROOT, 'include', 'pyiga_b200.h'), 'rb') as f:
        h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()[:16]


def _cc(job):
    src, obj, defs = job
    r = subprocess.run([CXX] + FLAGS + defs + ['-c', os.path.join(CSRC, src), '-o', obj],
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout + r.stderr)
    return obj


def build():
    os.makedirs(OUT, exist_ok=True)
    tag = _hash()
    stamp = os.path.join(OUT, 'stamp')
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == tag:
        return LIB
    jobs = [('api.cu', os.path.join(OUT, 'api.o'), [])]
    jobs += [('walk_inst.cu', os.path.join(OUT, 'walk_%d_%d.o' % pq), ['-DPB_P=%d' % pq[0], '-DPB_Q=%d' % pq[1]])
             for pq in PQ]
    jobs += [('walk_inst.cu', os.path.join(OUT, 'walk1_1_1.o'), ['-DPB_P=1', '-DPB_Q=1', '-DPB_WALK1_ONLY'])]
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(_cc, jobs))
    r = subprocess.run([CXX, '-shared', '-o', LIB] + objs, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout + r.stderr)
    with open(stamp, 'w') as f:
        f.write(tag)
    return LIB


if __name__ == '__main__':
    print(build())
