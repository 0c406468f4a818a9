"""CSR matrices resident on the GPU: elimination of rows/columns, matvec, vector gather/scatter
(``pb200_csr_*``, ``pb200_vec_*``).  Used by :class:`pyiga_b200.assemble.RestrictedLinearSystem`."""
import ctypes as C

import numpy as np
import scipy.sparse

from . import _device
from ._mlb import _ptr


def _as_i32(be, idx):
    return be.from_host(np.ascontiguousarray(idx, dtype=np.int32))


class DeviceCSR:
    """``indptr``, ``indices`` (int32 or int64) and ``values`` (float64) as device buffers."""

    def __init__(self, indptr, indices, values, shape, idt):
        self.be = _device.backend()
        self.indptr, self.indices, self.values = indptr, indices, values
        self.shape = (int(shape[0]), int(shape[1]))
        self.idt = np.dtype(idt)

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_scipy(cls, A):
        be = _device.backend()
        A = scipy.sparse.csr_matrix(A)
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        idt = np.int32 if (A.nnz < 2 ** 31 and max(A.shape) < 2 ** 31) else np.int64
        return cls(be.from_host(A.indptr.astype(idt)), be.from_host(A.indices.astype(idt)),
                   be.from_host(A.data.astype(np.float64)), A.shape, idt)

    @classmethod
    def from_mlmatrix(cls, M):
        """CSR arrays of an :class:`MLMatrix` whose values live on the device (no host copy)."""
        be = _device.backend()
        ds = M._device_handle()
        if ds is None or not ds.supported:
            return cls.from_scipy(M.asmatrix('csr'))
        indptr, indices, values = ds.csr_arrays(M._device_data())
        idt = np.int32 if be.itemsize(indptr) == 4 else np.int64
        return cls(indptr, indices, values, M.shape, idt)

    @classmethod
    def wrap(cls, A):
        from .mlmatrix import MLMatrix
        if isinstance(A, cls):
            return A
        if isinstance(A, MLMatrix):
            return cls.from_mlmatrix(A)
        if not scipy.sparse.issparse(A):
            A = scipy.sparse.csr_matrix(A)      # dense input, as in restrict_matrix of the reference
        return cls.from_scipy(A)

    @property
    def nnz(self):
        return self.be.size(self.values)

    # ---- operations ----------------------------------------------------------------------------
    def restrict(self, rows, colmap, ncols_new):
        """Rows `rows` (increasing old row indices) and the columns with ``colmap >= 0`` renumbered
        by `colmap`; returns a new :class:`DeviceCSR`."""
        be = self.be
        lib = be.lib
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        colmap = np.ascontiguousarray(colmap, dtype=np.int32)
        assert colmap.shape == (self.shape[1],)
        n = rows.size
        ib = self.idt.itemsize
        d_rows, d_colmap = be.from_host(rows), be.from_host(colmap)
        nbytes = C.c_size_t()
        _device.check(lib.pb200_csr_restrict_workspace(n, ib, C.byref(nbytes)))
        work = be.empty(max(nbytes.value, 8), np.uint8)
        indptr_new = be.empty(n + 1, self.idt)
        _device.check(lib.pb200_csr_restrict_count(n, _ptr(be, d_rows), _ptr(be, self.indptr), _ptr(be, self.indices), ib,
                                                   _ptr(be, d_colmap), _ptr(be, indptr_new), _ptr(be, work),
                                                   be.nbytes(work), be.stream()))
        nnz_new = int(be.to_host(indptr_new[n:n + 1])[0])
        indices_new = be.empty(nnz_new, self.idt)
        values_new = be.empty(nnz_new, np.float64)
        _device.check(lib.pb200_csr_restrict_fill(n, _ptr(be, d_rows), _ptr(be, self.indptr), _ptr(be, self.indices),
                                                  _ptr(be, self.values), ib, _ptr(be, d_colmap), _ptr(be, indptr_new),
                                                  _ptr(be, indices_new), _ptr(be, values_new), be.stream()))
        return DeviceCSR(indptr_new, indices_new, values_new, (n, ncols_new), self.idt)

    def matvec_device(self, d_x, d_y_in=None, alpha=1.0):
        """``(y_in or 0) + alpha * A x`` on device buffers."""
        be = self.be
        d_y = be.empty(self.shape[0])
        _device.check(be.lib.pb200_csr_matvec(self.shape[0], _ptr(be, self.indptr), _ptr(be, self.indices),
                                              _ptr(be, self.values), self.idt.itemsize, _ptr(be, d_x),
                                              0 if d_y_in is None else _ptr(be, d_y_in), float(alpha), _ptr(be, d_y),
                                              be.stream()))
        return d_y

    def dot(self, x):
        be = self.be
        if be.is_buffer(x):
            return self.matvec_device(x)
        x = np.ascontiguousarray(x, dtype=np.float64)
        return be.to_host(self.matvec_device(be.from_host(x.ravel()))).reshape((self.shape[0],) + x.shape[1:])

    def to_scipy(self):
        be = self.be
        A = scipy.sparse.csr_matrix((be.to_host(self.values), be.to_host(self.indices), be.to_host(self.indptr)),
                                    shape=self.shape)
        A.has_sorted_indices = True
        return A


def gather(d_in, idx, d_idx=None):
    """``out[k] = in[idx[k]]`` on the device."""
    be = _device.backend()
    d_idx = _as_i32(be, idx) if d_idx is None else d_idx
    n = be.size(d_idx)
    out = be.empty(n)
    _device.check(be.lib.pb200_vec_gather(n, _ptr(be, d_idx), _ptr(be, d_in), _ptr(be, out), be.stream()))
    return out


def scatter(d_in, idx, size, d_idx=None, out=None):
    """``out[idx[k]] = in[k]`` into a zero vector of length `size` (or into `out`)."""
    be = _device.backend()
    d_idx = _as_i32(be, idx) if d_idx is None else d_idx
    n = be.size(d_idx)
    if out is None:
        out = be.zeros(size)
    _device.check(be.lib.pb200_vec_scatter(n, _ptr(be, d_idx), _ptr(be, d_in), _ptr(be, out), be.stream()))
    return out
