"""Kronecker-product operator with dense factors applied on the device.

Mirrors ``KroneckerOperator`` (``pyiga/operators.py:60-86``) and ``apply_kronecker``
(``pyiga/kronecker.py:6-34``): ``(A_0 (x) ... (x) A_{d-1}) x`` by mode-wise products.  Used as
the preconditioner in the CG check that follows assembly (``pyiga/approx.py:82-93``).
"""
import ctypes as C

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from . import _device


class KroneckerOperator(scipy.sparse.linalg.LinearOperator):
    def __init__(self, *ops):
        self.ops = tuple(np.ascontiguousarray(A.toarray() if scipy.sparse.issparse(A) else A, dtype=np.float64)
                         for A in ops)
        assert all(A.ndim == 2 for A in self.ops), 'factors must be matrices'
        self.rows = tuple(A.shape[0] for A in self.ops)
        self.cols = tuple(A.shape[1] for A in self.ops)
        shape = (int(np.prod(self.rows)), int(np.prod(self.cols)))
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=shape, dtype=np.float64)
        self._dev = None

    def _device_factors(self):
        if self._dev is None:
            be = _device.backend()
            self._dev = [be.from_host(A.ravel()) for A in self.ops]
        return self._dev

    def matvec_device(self, d_x, d_y=None):
        be = _device.backend()
        d = len(self.ops)
        facs = self._device_factors()
        sz, maxsz = int(np.prod(self.cols)), int(np.prod(self.cols))
        for k in range(d - 1, -1, -1):
            sz = sz // self.cols[k] * self.rows[k]
            maxsz = max(maxsz, sz)
        if d_y is None:
            d_y = be.empty(self.shape[0])
        tmp = be.empty(2 * maxsz)
        ptrs = (C.c_void_p * d)(*[be.ptr(f) for f in facs])
        rows = (C.c_int * d)(*self.rows)
        cols = (C.c_int * d)(*self.cols)
        _device.check(be.lib.pb200_kron_matvec(d, ptrs, rows, cols, be.ptr(d_x), be.ptr(d_y), be.ptr(tmp), be.stream()))
        return d_y

    def _matvec(self, x):
        be = _device.backend()
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        return be.to_host(self.matvec_device(be.from_host(x)))

    def _transpose(self):
        return KroneckerOperator(*(A.T for A in self.ops))

    _adjoint = _transpose
