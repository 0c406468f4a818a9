"""SURVEY 8(f) rank 3 measured: elimination of the Dirichlet dofs of all six faces from the 3D p=3
stiffness matrix on device CSR arrays, and the restricted right-hand side.

    python tools/bc_restrict_bench.py [--n 128] [--p 3]

Prints one JSON line: ms and GB/s of the restriction (count + scan + compaction), of the CSR matvec,
and the algorithmic bytes they are measured against (csr.cuh header)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiga_b200 import assemble, assemblers, bspline, geometry  # noqa: E402
from pyiga_b200._csr import DeviceCSR  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=128)
    ap.add_argument('--p', type=int, default=3)
    a = ap.parse_args()
    kv = bspline.make_knots(a.p, 0.0, 1.0, a.n)
    kvs = (kv, kv, kv)
    geo = geometry.twisted_nurbs_box()
    M = assemblers.StiffnessAssembler3D(kvs, geo).assemble_mlb()
    A = DeviceCSR.from_mlmatrix(M)
    N = tuple(k.numdofs for k in kvs)
    idx = np.unique(np.concatenate([assemble.boundary_dofs(kvs, (ax, sd), ravel=True) for ax in range(3) for sd in (0, 1)]))
    ncols = A.shape[1]
    mask = np.ones(ncols, dtype=bool)
    mask[idx] = False
    free = np.nonzero(mask)[0]
    colmap = np.full(ncols, -1, dtype=np.int32)
    colmap[free] = np.arange(free.size, dtype=np.int32)
    ms_r, Ar = timed(lambda: A.restrict(free, colmap, free.size))
    ib = A.idt.itemsize
    nnz, nnz_r = A.nnz, Ar.nnz
    # rows of eliminated dofs are never read
    be = A.be
    indptr = be.to_host(A.indptr)
    nnz_rows = int((indptr[free + 1] - indptr[free]).sum())
    bytes_r = nnz_rows * ib + nnz_rows * (ib + 8) + nnz_r * (ib + 8)
    x = be.from_host(np.cos(np.arange(ncols) * 0.01))
    ms_mv, _ = timed(lambda: A.matvec_device(x))
    bytes_mv = nnz * (ib + 8) + 2 * 8 * ncols
    print(json.dumps({'workload': '3D stiffness p=%d n=%d, Dirichlet dofs of all faces eliminated' % (a.p, a.n),
                      'ndofs': ncols, 'free': int(free.size), 'nnz': nnz, 'nnz_restricted': nnz_r,
                      'restrict_ms': ms_r, 'restrict_GBps': bytes_r / ms_r / 1e6, 'restrict_bytes': bytes_r,
                      'note': 'restrict_ms includes the H2D of the row list/column map and the read-back of the new nnz',
                      'csr_matvec_ms': ms_mv, 'csr_matvec_GBps': bytes_mv / ms_mv / 1e6}))


if __name__ == '__main__':
    main()
