"""Device-side multi-level band structures: CSR export and matvec of MLB value tensors."""
import ctypes as C

import numpy as np
import scipy.sparse

from . import _device


def _ptr(be, buf):
    """device address of a backend buffer or of a torch view of one"""
    if hasattr(buf, 'data_ptr'):
        return buf.data_ptr()
    return be.ptr(buf)


class DeviceStructure:
    """Owns (or borrows, from an assembler) a ``pb200_mlstruct`` handle."""

    def __init__(self, structure=None, borrowed=None, owner=None):
        self.be = _device.backend()
        self.structure = structure
        self._owner = owner         # keeps the assembler alive while its structure is borrowed
        self._owned = None
        self.supported = True
        if borrowed is not None:
            self.handle = borrowed
            return
        S = structure
        if S.L < 2 or S.L > 3 or any(S._row_tables(k) is None for k in range(S.L)):
            self.supported = False
            self.handle = None
            return
        L = S.L
        rows = (C.c_int * L)(*[b[0] for b in S.bs])
        cols = (C.c_int * L)(*[b[1] for b in S.bs])
        nband = (C.c_int * L)(*[len(b) for b in S.bidx])
        self._bidx_keep = [np.ascontiguousarray(b, dtype=np.uint32) for b in S.bidx]
        ptrs = (C.c_void_p * L)(*[b.ctypes.data for b in self._bidx_keep])
        h = C.c_void_p()
        _device.check(self.be.lib.pb200_mlstruct_create(L, rows, cols, nband, ptrs, self.be.device_index, C.byref(h)))
        self.handle = h
        self._owned = h

    def __del__(self):
        try:
            if self._owned is not None:
                self.be.lib.pb200_mlstruct_destroy(self._owned)
                self._owned = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _sizes(self, row0=None):
        S = self.structure
        n0 = S.bs[0][0]
        ra, rb = (0, n0) if row0 is None else row0
        rs0 = np.searchsorted(S.bidx[0][:, 0], [ra, rb])
        nrows = (rb - ra) * int(np.prod([b[0] for b in S.bs[1:]], dtype=np.int64))
        count = int(rs0[1] - rs0[0]) * int(np.prod([len(b) for b in S.bidx[1:]], dtype=np.int64))
        return ra, rb, nrows, count

    def csr_arrays(self, d_data, row0=None, out=None, idt=None):
        """Device CSR arrays (indptr, indices, values) of the slab; int32 unless nnz >= 2^31.
        `out` may hold preallocated (indptr, indices, values) buffers at least as large."""
        be = self.be
        ra, rb, nrows, count = self._sizes(row0)
        if idt is None:
            idt = np.int32 if count < 2 ** 31 else np.int64
        if out is not None:
            indptr, indices, values = (o[:n] for o, n in zip(out, (nrows + 1, count, count)))
        else:
            indptr = be.empty(nrows + 1, idt)
            indices = be.empty(count, idt)
            values = be.empty(count, np.float64)
        _device.check(be.lib.pb200_mlb_to_csr(self.handle, ra, rb, _ptr(be, d_data), _ptr(be, indptr), _ptr(be, indices),
                                               _ptr(be, values), np.dtype(idt).itemsize, be.stream()))
        return indptr, indices, values

    def csr_values(self, d_data, row0=None, out=None):
        """Only the CSR-ordered values of the slab (device buffer); the pattern comes from
        :meth:`csr_pattern_host`."""
        be = self.be
        ra, rb, nrows, count = self._sizes(row0)
        values = be.empty(count, np.float64) if out is None else out[:count]
        idb = 4 if count < 2 ** 31 else 8
        _device.check(be.lib.pb200_mlb_to_csr(self.handle, ra, rb, _ptr(be, d_data), 0, 0, _ptr(be, values), idb, be.stream()))
        return values

    def csr_pattern_host(self, h_indptr, h_indices, row0=None, indptr_offset=0, nthreads=None):
        """Fill the host arrays `h_indptr` (nrows+1) and `h_indices` (count) — numpy arrays or pinned
        torch tensors of int32 / int64 — with the CSR pattern of the slab, using `nthreads` host
        threads.  No device work."""
        import os
        S = self.structure
        L = S.L
        tabs = [S._row_tables(k) for k in range(L)]
        ra, rb, nrows, count = self._sizes(row0)
        rows = (C.c_int * L)(*[b[0] for b in S.bs])
        cols = (C.c_int * L)(*[b[1] for b in S.bs])
        nband = (C.c_int * L)(*[len(b) for b in S.bidx])
        rs = [np.ascontiguousarray(t[0], dtype=np.int32) for t in tabs]
        jm = [np.ascontiguousarray(t[1], dtype=np.int32) for t in tabs]
        p_rs = (C.c_void_p * L)(*[a.ctypes.data for a in rs])
        p_jm = (C.c_void_p * L)(*[a.ctypes.data for a in jm])

        def addr(a):
            return a.data_ptr() if hasattr(a, 'data_ptr') else a.ctypes.data

        def itemsize(a):
            return a.element_size() if hasattr(a, 'element_size') else a.itemsize
        assert itemsize(h_indptr) == itemsize(h_indices)
        # share the host cores with the other ranks of the node (torchrun sets LOCAL_WORLD_SIZE)
        from ._hostcsr import host_cores
        nthreads = nthreads or max(1, host_cores() // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))) - 1)
        _device.check(self.be.lib.pb200_csr_pattern_host(L, rows, cols, nband, p_rs, p_jm, ra, rb, addr(h_indptr),
                                                         addr(h_indices), itemsize(h_indptr), int(indptr_offset),
                                                         int(nthreads)))

    def to_csr(self, d_data, row0=None):
        be = self.be
        indptr, indices, values = self.csr_arrays(d_data, row0)
        ra, rb, nrows, _ = self._sizes(row0)
        A = scipy.sparse.csr_matrix((be.to_host(values), be.to_host(indices), be.to_host(indptr)),
                                    shape=(nrows, self.structure.shape[1]))
        A.has_sorted_indices = True
        return A

    def matvec_device(self, d_data, d_x, d_y=None, row0=None, x_j0=0):
        """y = A[rows] x on device buffers; `d_data` starts at the first band entry of row0[0]."""
        be = self.be
        ra, rb, nrows, _ = self._sizes(row0)
        if d_y is None:
            d_y = be.empty(nrows)
        _device.check(be.lib.pb200_mlb_matvec(self.handle, ra, rb, _ptr(be, d_data), _ptr(be, d_x), int(x_j0),
                                               _ptr(be, d_y), be.stream()))
        return d_y

    def matvec(self, d_data, x):
        be = self.be
        return be.to_host(self.matvec_device(d_data, be.from_host(x)))
