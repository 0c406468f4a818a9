"""The C-ABI library loads without a GPU and exports every symbol include/pyiga_b200.h declares;
host-only entry points work; the package refuses to run without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'pyiga_b200.h')).read()
    return sorted(set(re.findall(r'PB200_API\s+[\w\s\*]+?\b(pb200_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from pyiga_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == names, 'ctypes prototypes out of sync with the header'
    assert lib.pb200_version() >= 100


def test_host_only_band_structure(ref):
    from pyiga_b200 import _lib, bspline
    from pyiga_b200.mlmatrix import compute_sparsity_ij
    lib = _lib.load()
    for p, n, mult in [(3, 7, 1), (2, 5, 2), (4, 3, 3), (1, 6, 1)]:
        kv = bspline.make_knots(p, 0.0, 1.0, n, mult=mult)
        cnt = C.c_int()
        rc = lib.pb200_band_structure(_lib.as_double_p(kv.kv), kv.kv.size, p, None, 0, 0, None, C.byref(cnt))
        assert rc == 0
        out = np.empty((cnt.value, 2), dtype=np.uint32)
        lib.pb200_band_structure(_lib.as_double_p(kv.kv), kv.kv.size, p, None, 0, 0, out.ctypes.data, C.byref(cnt))
        assert np.array_equal(out, compute_sparsity_ij(kv, kv))
    # two different spaces on one mesh (Petrov-Galerkin pattern)
    ku, kv_ = bspline.make_knots(3, 0.0, 1.0, 6), bspline.make_knots(2, 0.0, 1.0, 6)
    cnt = C.c_int()
    lib.pb200_band_structure(_lib.as_double_p(ku.kv), ku.kv.size, 3, _lib.as_double_p(kv_.kv), kv_.kv.size, 2, None,
                             C.byref(cnt))
    out = np.empty((cnt.value, 2), dtype=np.uint32)
    lib.pb200_band_structure(_lib.as_double_p(ku.kv), ku.kv.size, 3, _lib.as_double_p(kv_.kv), kv_.kv.size, 2,
                             out.ctypes.data, C.byref(cnt))
    assert np.array_equal(out, compute_sparsity_ij(ku, kv_))


def test_error_reporting():
    from pyiga_b200 import _lib
    lib = _lib.load()
    bad = np.array([0.0, 1.0, 0.5, 1.0])
    rc = lib.pb200_band_structure(_lib.as_double_p(bad), 4, 1, None, 0, 0, None, None)
    assert rc == -1 and b'increasing' in lib.pb200_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib, rc)


def test_no_cpu_fallback():
    """without a CUDA device the product backend must refuse to start"""
    import torch
    from pyiga_b200 import _device
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    _device._backend = None
    with pytest.raises(RuntimeError, match='no CPU path'):
        _device.backend()
    from pyiga_b200 import assemble, bspline, geometry
    kv = bspline.make_knots(2, 0.0, 1.0, 3)
    with pytest.raises(RuntimeError):
        assemble.mass((kv, kv), geometry.unit_square())


def test_structures_match_reference(ref):
    from helpers import CASES, make_space
    from pyiga_b200.mlmatrix import MLStructure
    for case in CASES:
        kvs = make_space(ref, case)
        S = MLStructure.from_kvs(kvs, kvs)
        for k in range(S.L):
            assert np.array_equal(S.bidx[k], ref['%s_bidx%d' % (case, k)])
        I, J = S.nonzero()
        assert I.dtype == np.uint64 and len(I) == np.prod([len(b) for b in S.bidx])
        Il, Jl = S.nonzero(lower_tri=True)
        assert np.all(Jl <= Il) and 2 * len(Il) - S.shape[0] == len(I)


def test_knotvector_api():
    from pyiga_b200 import bspline
    kv = bspline.make_knots(3, 0.0, 1.0, 4, mult=2)
    assert kv.numdofs == kv.kv.size - 4 and kv.numspans == 4
    ms = kv.mesh_support_idx_all()
    assert ms.shape == (kv.numdofs, 2) and ms[0, 0] == 0 and ms[-1, 1] == 4
    assert kv.findspan(1.0) == kv.kv.size - 3 - 2 and kv.findspan(0.0) == 3
    assert kv == kv.copy() and kv.refine().numspans == 8


def test_host_only_csr_pattern(ref):
    """pb200_csr_pattern_host is a host-only entry point of the PRODUCT library: callable without a GPU"""
    from helpers import make_space
    from pyiga_b200 import _lib
    from pyiga_b200.mlmatrix import MLStructure
    lib = _lib.load()
    kvs = make_space(ref, 'a3_mixed')
    S = MLStructure.from_kvs(kvs, kvs)
    L = S.L
    tabs = [S._row_tables(k) for k in range(L)]
    rows = (C.c_int * L)(*[b[0] for b in S.bs])
    cols = (C.c_int * L)(*[b[1] for b in S.bs])
    nband = (C.c_int * L)(*[len(b) for b in S.bidx])
    rs = [np.ascontiguousarray(t[0], dtype=np.int32) for t in tabs]
    jm = [np.ascontiguousarray(t[1], dtype=np.int32) for t in tabs]
    p_rs = (C.c_void_p * L)(*[a.ctypes.data for a in rs])
    p_jm = (C.c_void_p * L)(*[a.ctypes.data for a in jm])
    want_ptr, want_idx = ref['a3_mixed_stiff_indptr'], ref['a3_mixed_stiff_indices']
    ip = np.empty(want_ptr.size, dtype=np.int32)
    ix = np.empty(want_idx.size, dtype=np.int32)
    rc = lib.pb200_csr_pattern_host(L, rows, cols, nband, p_rs, p_jm, 0, S.bs[0][0], ip.ctypes.data, ix.ctypes.data, 4, 0, 3)
    assert rc == 0
    assert np.array_equal(ip, want_ptr) and np.array_equal(ix, want_idx)
    # argument checks
    assert lib.pb200_csr_pattern_host(L, rows, cols, nband, p_rs, p_jm, 0, S.bs[0][0] + 1, ip.ctypes.data, ix.ctypes.data, 4, 0, 1) == -1
    assert lib.pb200_csr_pattern_host(L, rows, cols, nband, p_rs, p_jm, 0, 1, ip.ctypes.data, ix.ctypes.data, 3, 0, 1) == -1
    assert lib.pb200_csr_pattern_host(1, rows, cols, nband, p_rs, p_jm, 0, 1, ip.ctypes.data, ix.ctypes.data, 4, 0, 1) != 0


def test_argument_checks_of_device_entry_points():
    """invalid arguments are rejected before any CUDA call"""
    from pyiga_b200 import _lib
    lib = _lib.load()
    n = C.c_size_t()
    assert lib.pb200_csr_restrict_workspace(10, 5, C.byref(n)) == -1
    assert lib.pb200_csr_restrict_workspace(-1, 4, C.byref(n)) == -1
    assert lib.pb200_csr_restrict_workspace(10 ** 7, 8, C.byref(n)) == 0 and n.value > 0
    assert lib.pb200_csr_restrict_count(5, None, None, None, 4, None, None, None, 0, None) == -1
    assert lib.pb200_csr_matvec(3, None, None, None, 4, None, None, 1.0, None, None) == -1
    assert lib.pb200_vec_gather(3, None, None, None, None) == -1
    assert lib.pb200_asm_rows_count(None, None, 0, None) == -1
    assert b'null' in lib.pb200_last_error()


def test_host_csr_pattern_product_library():
    """pb200_csr_pattern_host of the PRODUCT library (host-only entry point: staged rows, non-temporal stores)
    against the pattern built from MLStructure.nonzero — whole matrices and slabs, several threads, results that
    do not start on a cache line, 32- and 64-bit indices"""
    import scipy.sparse
    from pyiga_b200 import _lib, bspline
    from pyiga_b200.mlmatrix import MLStructure
    lib = _lib.load()
    for ps, ns in [((3, 3, 3), (9, 8, 10)), ((2, 4), (30, 41)), ((1, 2, 3), (5, 6, 7))]:
        kvs = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
        S = MLStructure.from_kvs(kvs, kvs)
        L = S.L
        I, J = (a.astype(np.int64) for a in S.nonzero())
        A = scipy.sparse.csr_matrix((np.ones(I.size), (I, J)), shape=S.shape)
        A.sort_indices()
        tabs = [S._row_tables(k) for k in range(L)]
        rows = (C.c_int * L)(*[b[0] for b in S.bs])
        cols = (C.c_int * L)(*[b[1] for b in S.bs])
        nband = (C.c_int * L)(*[len(b) for b in S.bidx])
        rs = [np.ascontiguousarray(t[0], dtype=np.int32) for t in tabs]
        jm = [np.ascontiguousarray(t[1], dtype=np.int32) for t in tabs]
        p_rs = (C.c_void_p * L)(*[a.ctypes.data for a in rs])
        p_jm = (C.c_void_p * L)(*[a.ctypes.data for a in jm])
        n0 = S.bs[0][0]
        inner = S.shape[0] // n0
        for (ra, rb), idt, nthr, shift in [((0, n0), np.int32, 1, 0), ((0, n0), np.int32, 7, 5), ((1, n0 - 1), np.int64, 4, 3),
                                           ((n0 - 1, n0), np.int32, 16, 9)]:
            r0, r1 = ra * inner, rb * inner
            want_ptr = A.indptr[r0:r1 + 1].astype(np.int64) - A.indptr[r0] + 11
            want_idx = A.indices[A.indptr[r0]:A.indptr[r1]]
            buf_ptr = np.full(want_ptr.size + shift + 16, -7, dtype=idt)
            buf_idx = np.full(want_idx.size + shift + 16, -7, dtype=idt)
            got_ptr, got_idx = buf_ptr[shift:shift + want_ptr.size], buf_idx[shift:shift + want_idx.size]
            rc = lib.pb200_csr_pattern_host(L, rows, cols, nband, p_rs, p_jm, ra, rb, got_ptr.ctypes.data, got_idx.ctypes.data,
                                            np.dtype(idt).itemsize, 11, nthr)
            assert rc == 0, lib.pb200_last_error()
            assert np.array_equal(got_ptr, want_ptr) and np.array_equal(got_idx, want_idx), (ps, ns, ra, rb, nthr)
            assert (buf_idx[:shift] == -7).all() and (buf_idx[shift + want_idx.size:] == -7).all()
            assert (buf_ptr[:shift] == -7).all() and (buf_ptr[shift + want_ptr.size:] == -7).all()
