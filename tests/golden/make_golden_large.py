"""Sampled golden fixtures of the REAL reference (c-f-h/pyiga) at the sizes that are benchmarked.

Run in the build container (the reference is installed under oracle/_ref by oracle/build_ref.py):

    python tests/golden/make_golden_large.py [case ...]

One file per case, ``tests/golden/large_<case>.npz``.  The reference cannot hold the full
matrices of these sizes comfortably (p=3 n=128: 741 M nonzeros), so every case stores a SAMPLE
of entries computed by the reference's own per-entry quadrature, ``asm.multi_entries(idx)``
(pyiga/genericasm.pxi:722-758), the call ``assemble_entries`` itself makes
(pyiga/assemble.py:744).  The sample holds

  * complete rows (every pair of the row inside the band pattern): the 8 corner dofs, dofs on
    edges and faces, the rows on both sides of every seam of a 2/4/8-way row-slab partition of
    the first axis (the multi-GPU sharding), and random interior rows  ->  row sums follow;
  * random single entries inside the pattern;
  * pairs just outside the pattern (distance p+1 on one axis), for which the reference
    returns 0.0.

p=3 n=64 is additionally assembled IN FULL by ``assemble.stiffness/mass`` and the sample is
checked against that matrix here; max|A|, the sum of all entries and a strided matvec checksum
``(A x)[::stride]`` with a deterministic x are stored as whole-matrix checks.
p=4 n=192 does not fit the reference's Jacobian in this container's 62 GB (884.7 M Gauss
points x 9 doubles = 64 GB before any temporary), so degree 4 is sampled at n=96.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))

import pyiga  # noqa: E402
from pyiga import assemble, assemblers, bspline, geometry, mlmatrix  # noqa: E402

CONVDIFF = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'


def diff_coeff(x, y, z):
    return 1.0 + x * y


def twisted_nurbs_box():
    G = geometry.twisted_box()
    i, j, k = np.meshgrid(np.arange(2), np.arange(4), np.arange(2), indexing='ij')
    W = 1.0 + 0.25 * ((i + 2 * j + 3 * k) % 3)
    return geometry.NurbsFunc(G.kvs, G.coeffs.copy(), W)


def seam_rows(N, p):
    """first-axis rows next to the seams of 2/4/8-way partitions balanced by band count"""
    band = np.array([min(N, i + p + 1) - max(0, i - p) for i in range(N)])
    cum = np.concatenate(([0], np.cumsum(band)))
    rows = set()
    for world in (2, 4, 8):
        for r in range(1, world):
            cut = int(np.searchsorted(cum, cum[-1] * r / world))
            for d in range(-p - 1, p + 2):
                if 0 <= cut + d < N:
                    rows.add(cut + d)
    return sorted(rows)


def build_sample(N, p, rng, nrand_rows, nrand_entries, nout):
    """(I, J) sample on an N^3 dof grid with per-axis bandwidth p; returns uint64 (n,2) and the
    number of leading pairs that form complete rows"""
    def rav(i0, i1, i2):
        return (np.asarray(i0, dtype=np.int64) * N + i1) * N + i2

    rows = []
    ends = (0, N - 1)
    for a in ends:                      # corners
        for b in ends:
            for c in ends:
                rows.append((a, b, c))
    mid = N // 2
    for a in ends:                      # edges and faces
        for b in ends:
            rows += [(a, b, mid), (a, mid, b), (mid, a, b)]
        rows += [(a, mid, mid + 1), (mid, a, mid - 1), (mid + 1, mid, a)]
    for i0 in seam_rows(N, p):          # slab seams: a boundary column, an interior column
        rows += [(i0, mid, mid), (i0, 0, N - 1), (i0, int(rng.integers(0, N)), int(rng.integers(0, N)))]
    for _ in range(nrand_rows):
        rows.append(tuple(int(v) for v in rng.integers(0, N, 3)))
    rows = sorted(set(rows))
    I, J = [], []
    for (i0, i1, i2) in rows:
        r0 = np.arange(max(0, i0 - p), min(N, i0 + p + 1))
        r1 = np.arange(max(0, i1 - p), min(N, i1 + p + 1))
        r2 = np.arange(max(0, i2 - p), min(N, i2 + p + 1))
        j0, j1, j2 = np.meshgrid(r0, r1, r2, indexing='ij')
        jj = rav(j0.ravel(), j1.ravel(), j2.ravel())
        I.append(np.full(jj.size, rav(i0, i1, i2)))
        J.append(jj)
    nfull = sum(a.size for a in I)
    # random entries inside the pattern
    i = rng.integers(0, N, (nrand_entries, 3))
    d = rng.integers(-p, p + 1, (nrand_entries, 3))
    j = np.clip(i + d, 0, N - 1)
    I.append(rav(i[:, 0], i[:, 1], i[:, 2]))
    J.append(rav(j[:, 0], j[:, 1], j[:, 2]))
    # just outside the pattern: distance p+1 on one axis, inside on the others
    i = rng.integers(0, N, (nout, 3))
    d = rng.integers(-p, p + 1, (nout, 3))
    ax = rng.integers(0, 3, nout)
    sgn = rng.choice((-1, 1), nout)
    d[np.arange(nout), ax] = sgn * (p + 1)
    j = i + d
    ok = np.all((j >= 0) & (j < N), axis=1)
    i, j = i[ok], j[ok]
    I.append(rav(i[:, 0], i[:, 1], i[:, 2]))
    J.append(rav(j[:, 0], j[:, 1], j[:, 2]))
    ij = np.column_stack((np.concatenate(I), np.concatenate(J))).astype(np.uint64)
    return ij, nfull, int(ok.sum())


def make_asm(form, kvs, geo):
    if form == 'stiffness':
        return assemblers.StiffnessAssembler3D(kvs, geo)
    if form == 'mass':
        return assemblers.MassAssembler3D(kvs, geo)
    if form == 'convdiff':
        return assemble.instantiate_assembler(CONVDIFF, kvs, {'geo': geo, 'diff_coeff': diff_coeff}, None)
    raise ValueError(form)


CASES = {
    # name: (form, p, n, geometry, full assembly too?)
    'stiff_p3_n64': ('stiffness', 3, 64, 'tnb', True),
    'mass_p3_n64': ('mass', 3, 64, 'tnb', True),
    'stiff_p3_n128': ('stiffness', 3, 128, 'tnb', False),
    'mass_p3_n128': ('mass', 3, 128, 'tnb', False),
    'stiff_p4_n96': ('stiffness', 4, 96, 'tnb', False),
    'mass_p4_n96': ('mass', 4, 96, 'tnb', False),
    'convdiff_p3_n96': ('convdiff', 3, 96, 'tb', False),
    # small copies of the same fixture layout: they drive the CPU (emulation) run of the checks
    'tiny_stiff_p2_n6': ('stiffness', 2, 6, 'tnb', True),
    'tiny_mass_p3_n5': ('mass', 3, 5, 'tnb', True),
}


def run_case(name):
    form, p, n, gname, full = CASES[name]
    t0 = time.time()
    rng = np.random.default_rng(sum(map(ord, name)))
    kv = bspline.make_knots(p, 0.0, 1.0, n)
    kvs = 3 * (kv,)
    N = kv.numdofs
    geo = twisted_nurbs_box() if gname == 'tnb' else geometry.twisted_box()
    nrows = 60 if p <= 3 else 30
    tiny = name.startswith('tiny')
    ij, nfull, nout = build_sample(N, p, rng, nrand_rows=5 if tiny else nrows, nrand_entries=500 if tiny else 30000,
                                   nout=100 if tiny else 3000)
    asm = make_asm(form, kvs, geo)
    t1 = time.time()
    vals = np.asarray(asm.multi_entries(ij))
    t2 = time.time()
    assert np.all(vals[len(vals) - nout:] == 0.0), 'the reference returns 0 outside the pattern'
    out = {'form': np.array(form), 'p': np.array(p), 'n': np.array(n), 'geo': np.array(gname),
           'ij': ij, 'val': vals, 'nfull': np.array(nfull), 'nout': np.array(nout),
           'sample_maxabs': np.array(np.abs(vals).max())}
    if full:
        fn = assemble.stiffness if form == 'stiffness' else assemble.mass
        A = fn(kvs, geo).tocsr()
        t3 = time.time()
        got = np.asarray(A[ij[:, 0].astype(np.int64), ij[:, 1].astype(np.int64)]).ravel()
        dev = np.abs(got - vals).max()
        assert dev <= 1e-14 * np.abs(A.data).max(), dev
        x = np.cos(0.37 * np.arange(A.shape[1]) + 0.1)
        stride = 13
        out.update(full_maxabs=np.array(np.abs(A.data).max()), full_sum=np.array(A.data.sum()),
                   full_abs_sum=np.array(np.abs(A.data).sum()), full_nnz=np.array(A.nnz),
                   mv_stride=np.array(stride), mv_y=(A @ x)[::stride], mv_freq=np.array(0.37), mv_phase=np.array(0.1))
        print('  full assembly %.1f s, nnz %d, sample vs full %.2e' % (t3 - t2, A.nnz, dev))
    np.savez_compressed(os.path.join(HERE, 'large_%s.npz' % name), **out)
    print('%s: N=%d, %d sampled pairs (%d in complete rows, %d outside), setup %.1f s, multi_entries %.1f s, max|a| %.4e'
          % (name, N, len(ij), nfull, nout, t1 - t0, t2 - t1, np.abs(vals).max()), flush=True)


if __name__ == '__main__':
    pyiga.set_max_threads(os.cpu_count())
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)
