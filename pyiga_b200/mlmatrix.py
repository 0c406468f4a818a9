"""Multi-level banded matrices (``MLStructure`` / ``MLMatrix``) backed by device kernels.

Same data model as the reference (``pyiga/mlmatrix.py:15-305``): level k has block size
``bs[k] = (m_k, n_k)`` and a list ``bidx[k]`` (uint32, nnz_k x 2) of its nonzero positions,
sorted by row then column; the value tensor has shape ``(nnz_0, ..., nnz_{L-1})`` and element
``[mu_0, ..., mu_{L-1}]`` is ``A[I, J]`` with ``(i_k, j_k) = bidx[k][mu_k]`` and C-order
raveled ``I, J``.  Index bookkeeping (patterns, ``nonzero``) is integer work done with numpy on
the host; the floating point work — matvec and the CSR value permutation — runs on the GPU
through ``pb200_mlb_matvec`` / ``pb200_mlb_to_csr``.
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg


def compute_sparsity_ij(kv_trial, kv_test):
    """``(nnz, 2)`` uint32 array of the pairs (i, j) — test function i, trial function j — whose
    supports share a mesh span, sorted by i then j (reference: ``pyiga/mlmatrix.py:420-440``)."""
    su = np.asarray(kv_trial.mesh_support_idx_all())
    sv = np.asarray(kv_test.mesh_support_idx_all())
    # supports are intervals with non-decreasing end points, so each row interacts with a
    # contiguous column range [lo, hi)
    lo = np.searchsorted(su[:, 1], sv[:, 0], side='right')
    hi = np.searchsorted(su[:, 0], sv[:, 1], side='left')
    counts = np.maximum(hi - lo, 0)
    rows = np.repeat(np.arange(sv.shape[0]), counts)
    starts = np.repeat(lo, counts)
    offs = np.arange(counts.sum()) - np.repeat(np.cumsum(counts) - counts, counts)
    return np.column_stack((rows, starts + offs)).astype(np.uint32)


def compute_banded_sparsity_ij(n, bw):
    i = np.repeat(np.arange(n), 2 * bw + 1)
    j = i + np.tile(np.arange(-bw, bw + 1), n)
    ok = (j >= 0) & (j < n)
    return np.column_stack((i[ok], j[ok])).astype(np.uint32)


def compute_dense_ij(m, n):
    return np.column_stack((np.repeat(np.arange(m), n), np.tile(np.arange(n), m))).astype(np.uint32)


class MLStructure:
    """Sparsity structure of a Kronecker-type multi-level matrix (``pyiga/mlmatrix.py:15-198``)."""

    def __init__(self, bs, bidx):
        self.bs = tuple(tuple(int(x) for x in b) for b in bs)
        self._bs_arr = np.array(self.bs)
        assert self._bs_arr.ndim == 2 and self._bs_arr.shape[1] == 2, 'invalid block sizes'
        self.bidx = tuple(np.ascontiguousarray(b, dtype=np.uint32) for b in bidx)
        assert len(self.bs) == len(self.bidx)
        self.L = len(self.bs)
        self.shape = (int(np.prod([b[0] for b in self.bs], dtype=np.int64)),
                      int(np.prod([b[1] for b in self.bs], dtype=np.int64)))

    @staticmethod
    def multi_banded(bs, bw):
        return MLStructure(tuple((n, n) for n in bs),
                           tuple(compute_banded_sparsity_ij(n, p) for n, p in zip(bs, bw)))

    @staticmethod
    def dense(shape):
        return MLStructure((shape,), (compute_dense_ij(shape[0], shape[1]),))

    @staticmethod
    def from_kvs(kvs0, kvs1):
        """Structure of a matrix assembled over trial space `kvs0` and test space `kvs1`
        (reference: ``pyiga/mlmatrix.py:59-65``)."""
        bs = tuple((kv1.numdofs, kv0.numdofs) for kv0, kv1 in zip(kvs0, kvs1))
        bidx = tuple(compute_sparsity_ij(kv0, kv1) for kv0, kv1 in zip(kvs0, kvs1))
        return MLStructure(bs, bidx)

    @staticmethod
    def from_matrix(A):
        I, J = A.nonzero()
        return MLStructure((tuple(A.shape),), (np.column_stack((I, J)).astype(np.uint32),))

    @staticmethod
    def from_kronecker(As):
        S = MLStructure.from_matrix(As[0])
        for A in As[1:]:
            S = S.join(MLStructure.from_matrix(A))
        return S

    def join(self, other):
        return MLStructure(self.bs + other.bs, self.bidx + other.bidx)

    def reorder(self, axes):
        assert len(axes) == self.L
        return MLStructure(tuple(self.bs[j] for j in axes), tuple(self.bidx[j] for j in axes))

    def slice(self, start, end=None):
        assert 0 <= start < self.L, 'invalid slice index'
        if end is None:
            end = start + 1
        return MLStructure(self.bs[start:end], self.bidx[start:end])

    def transpose(self):
        return MLStructure(tuple((b[1], b[0]) for b in self.bs),
                           tuple(np.ascontiguousarray(bx[:, ::-1]) for bx in self.bidx))

    def make_mlmatrix(self, data=None, matrix=None):
        return MLMatrix(structure=self, data=data, matrix=matrix)

    def nonzero(self, lower_tri=False):
        """Row / column indices of all structural nonzeros, in the order of the value tensor
        (reference: ``pyiga/mlmatrix.py:113-130`` -> ``ml_nonzero_{2d,3d,nd}``)."""
        I = np.zeros(1, dtype=np.uint64)
        J = np.zeros(1, dtype=np.uint64)
        for k in range(self.L):
            m, n = self.bs[k]
            bi = self.bidx[k][:, 0].astype(np.uint64)
            bj = self.bidx[k][:, 1].astype(np.uint64)
            I = (I[:, None] * np.uint64(m) + bi[None, :]).ravel()
            J = (J[:, None] * np.uint64(n) + bj[None, :]).ravel()
        if lower_tri:
            assert self.L > 1, 'Lower triangular part not implemented in 1D'
            keep = J <= I
            I, J = I[keep], J[keep]
        return I, J

    def sequential_bidx(self):
        """per level: the nonzero positions as C-order offsets into the (m_k x n_k) block"""
        return [bx[:, 0] * np.uint32(m) + bx[:, 1] for bx, (m, _) in zip(self.bidx, self.bs)]

    def _level_csr(self, k):
        """columns of level k grouped by row: (start[m+1], cols) — a CSR view of ``bidx[k]``"""
        bx = self.bidx[k]
        order = np.argsort(bx[:, 0], kind='stable')
        start = np.searchsorted(bx[order, 0], np.arange(self.bs[k][0] + 1)).astype(np.int64)
        return start, bx[order, 1].astype(np.int64)

    def nonzeros_for_rows(self, row_indices, renumber_rows=False):
        """Nonzero positions restricted to the given rows (``pyiga/mlmatrix.py:150-185``): for row
        ``(i_0,..,i_{L-1})`` the columns are the Cartesian product of the per-level column lists,
        last level fastest.  Vectorised over all rows: entry t of a row is decomposed in the mixed
        radix of the row's per-level counts."""
        rows = np.asarray(row_indices, dtype=np.int64).ravel()
        if rows.size == 0:
            e = np.empty(0, dtype=int)
            return (e, e, e) if renumber_rows else (e, e)
        level = [self._level_csr(k) for k in range(self.L)]
        multi = np.unravel_index(rows, tuple(b[0] for b in self.bs))
        first = [level[k][0][multi[k]] for k in range(self.L)]                  # start of the column list
        count = [level[k][0][multi[k] + 1] - first[k] for k in range(self.L)]   # its length
        per_row = np.prod(count, axis=0)
        which = np.repeat(np.arange(rows.size), per_row)
        t = np.arange(per_row.sum()) - np.repeat(np.cumsum(per_row) - per_row, per_row)
        digits = [None] * self.L
        for k in range(self.L - 1, -1, -1):
            ck = count[k][which]
            digits[k] = t % np.maximum(ck, 1)
            t = t // np.maximum(ck, 1)
        Js = np.zeros(which.size, dtype=np.int64)
        for k in range(self.L):
            Js = Js * self.bs[k][1] + level[k][1][first[k][which] + digits[k]]
        Is = rows[which]
        if renumber_rows:
            return Is, Js, which
        return Is, Js

    def positions(self, I, J):
        """Flat position in the value tensor of the entries (I, J), or -1 for pairs outside the
        pattern (for which ``multi_entries`` returns 0, ``pyiga/genericasm.pxi:722-758``).  Needs
        contiguous per-row column ranges on every level (all spline patterns)."""
        I = np.array(I, dtype=np.int64).ravel()
        J = np.array(J, dtype=np.int64).ravel()
        pos = np.zeros(I.shape, dtype=np.int64)
        ok = np.ones(I.shape, dtype=bool)
        ik = np.unravel_index(I, tuple(b[0] for b in self.bs))
        jk = np.unravel_index(J, tuple(b[1] for b in self.bs))
        for k in range(self.L):
            tabs = self._row_tables(k)
            assert tabs is not None, 'level %d has no contiguous row pattern' % k
            row_start, jmin = (t.astype(np.int64) for t in tabs)
            d = jk[k] - jmin[ik[k]]
            ok &= (d >= 0) & (d < row_start[ik[k] + 1] - row_start[ik[k]])
            pos = pos * len(self.bidx[k]) + row_start[ik[k]] + np.where(ok, d, 0)
        pos[~ok] = -1
        return pos

    def nonzeros_for_columns(self, col_indices):
        J, I = self.transpose().nonzeros_for_rows(col_indices)
        return I, J

    # helpers for the device kernels -----------------------------------------------------------
    def _row_tables(self, k):
        """(row_start[m+1], jmin[m]) of level k if every row's columns are contiguous and the list
        is sorted by row then column (true for all spline patterns); else None."""
        bx = self.bidx[k].astype(np.int64)
        m = self.bs[k][0]
        if bx.shape[0] == 0:
            return None
        if np.any(np.diff(bx[:, 0]) < 0):
            return None
        row_start = np.searchsorted(bx[:, 0], np.arange(m + 1)).astype(np.int32)
        same_row = np.diff(bx[:, 0]) == 0
        if np.any(np.diff(bx[:, 1])[same_row] != 1):
            return None
        jmin = np.zeros(m, dtype=np.int32)
        has = row_start[1:] > row_start[:-1]
        jmin[has] = bx[row_start[:-1][has], 1]
        return row_start, jmin


class MLMatrix(scipy.sparse.linalg.LinearOperator):
    """Multi-level structured sparse matrix stored as a dense value tensor
    (``pyiga/mlmatrix.py:201-305``).  `data` may live on the host (numpy) or on the device
    (the buffer returned by the assemblers); ``.data`` always hands out a numpy array."""

    def __init__(self, structure, data=None, matrix=None):
        self.structure = structure
        self.L = structure.L
        self.shape = structure.shape
        self.datashape = tuple(len(bi) for bi in structure.bidx)
        self._data = None
        self._ddata = None      # device copy
        self._dev = None        # (assembler-like) device structure handle
        assert data is None or matrix is None, 'Can only specify one of `data` and `matrix`'
        dtype = np.float64
        if data is not None:
            from . import _device
            on_device = _device._backend is not None and _device._backend.is_buffer(data)
            if not on_device and isinstance(data, np.ndarray):
                assert data.shape == self.datashape, 'Wrong shape of data tensor'
                self._data = np.asarray(data, order='C')
                dtype = self._data.dtype
            else:       # device buffer
                self._ddata = data
        elif matrix is not None:
            assert matrix.shape == self.shape, 'Matrix has wrong shape'
            I, J = self.nonzero()
            vals = np.asarray(matrix[I.astype(np.int64), J.astype(np.int64)]).reshape(self.datashape)
            self._data = np.asarray(vals, order='C')
            dtype = self._data.dtype
        scipy.sparse.linalg.LinearOperator.__init__(self, shape=self.shape, dtype=dtype)

    @property
    def nnz(self):
        return int(np.prod(self.datashape, dtype=np.int64))

    @property
    def data(self):
        if self._data is None and self._ddata is not None:
            from . import _device
            self._data = _device.backend().to_host(self._ddata).reshape(self.datashape)
        return self._data

    @data.setter
    def data(self, X):
        assert X.shape == self.datashape
        self._data = np.asarray(X, order='C')
        self._ddata = None

    def nonzero(self, lower_tri=False):
        return self.structure.nonzero(lower_tri=lower_tri)

    def reorder(self, axes):
        assert len(axes) == self.L
        newdata = None if self.data is None else np.ascontiguousarray(np.transpose(self.data, axes))
        return MLMatrix(structure=self.structure.reorder(axes), data=newdata)

    # device side ------------------------------------------------------------------------------
    def _device_handle(self):
        from . import _mlb
        if self._dev is None:
            self._dev = _mlb.DeviceStructure(self.structure)
        return self._dev

    def _device_data(self):
        if self._ddata is None:
            from . import _device
            assert self._data is not None, 'matrix has no data'
            self._ddata = _device.backend().from_host(np.ascontiguousarray(self._data, dtype=np.float64).ravel())
        return self._ddata

    def asmatrix(self, format='csr'):
        """Sparse matrix in the given scipy format; the CSR arrays are produced on the device."""
        assert self._data is not None or self._ddata is not None, 'matrix has no data'
        if self.L == 1:
            bx = self.structure.bidx[0]
            return scipy.sparse.coo_matrix((self.data, (bx[:, 0], bx[:, 1])), shape=self.shape).asformat(format)
        h = self._device_handle()
        if h.supported:
            return h.to_csr(self._device_data()).asformat(format)
        I, J = self.nonzero()
        A = scipy.sparse.csr_matrix((self.data.ravel('C'), (I.astype(np.int64), J.astype(np.int64))), shape=self.shape)
        return A.asformat(format)

    def _matvec(self, x):
        assert self._data is not None or self._ddata is not None, 'matrix has no data'
        x = np.asarray(x)
        assert x.shape[0] == self.shape[1], 'Invalid input size'
        if self.L == 1:
            # the device kernels take 2 or 3 levels: a one-level matrix is the second level under a dense 1 x 1 level
            if getattr(self, '_lifted', None) is None:
                self._lifted = MLMatrix(structure=MLStructure.dense((1, 1)).join(self.structure),
                                        data=np.asarray(self.data, dtype=np.float64).reshape(1, -1))
            return self._lifted._matvec(x)
        h = self._device_handle()
        if not h.supported:
            raise NotImplementedError('matvec needs a 2- or 3-level matrix with contiguous row patterns')
        return h.matvec(self._device_data(), np.ascontiguousarray(x, dtype=np.float64).ravel())


# ---------------------------------------------------------------------------------------------
# index conversions between sequential and multilevel numbering (``pyiga/mlmatrix.py:312-416``)
# ---------------------------------------------------------------------------------------------
def from_seq(i, dims):
    """lexicographic index -> multi-index (a list), last dimension fastest"""
    return [int(t) for t in np.unravel_index(int(i), tuple(int(d) for d in dims))]


def to_seq(I, dims):
    """multi-index -> lexicographic index"""
    return int(np.ravel_multi_index(tuple(int(t) for t in I), tuple(int(d) for d in dims)))


def reindex_to_multilevel(i, j, bs):
    """Entry (i, j) of a multilevel matrix with block sizes `bs` (L x 2) -> per level the raveled position
    (row, column) inside that level's block."""
    bs = np.asarray(bs, dtype=np.int64).reshape(-1, 2)
    rows, cols = from_seq(i, bs[:, 0]), from_seq(j, bs[:, 1])
    return tuple(r * int(b[1]) + c for r, c, b in zip(rows, cols, bs))


def reindex_from_multilevel(M, bs):
    """inverse of :func:`reindex_to_multilevel`: per-level block positions -> entry (i, j)"""
    bs = np.asarray(bs, dtype=np.int64).reshape(-1, 2)
    i = j = 0
    for m, (nr, nc) in zip(M, bs):
        r, c = divmod(int(m), int(nc))
        i, j = i * int(nr) + r, j * int(nc) + c
    return (i, j)


def reorder(X, m1, n1):
    """Van Loan - Pitsianis rearrangement of a dense matrix of m1 x n1 blocks: row (i * n1 + j) of the result is
    the row-wise vectorisation of block (i, j)."""
    X = np.asarray(X)
    m2, n2 = X.shape[0] // m1, X.shape[1] // n1
    assert X.shape == (m1 * m2, n1 * n2), "Invalid block size"
    return X.reshape(m1, m2, n1, n2).transpose(0, 2, 1, 3).reshape(m1 * n1, m2 * n2)


def reindex_from_reordered(i, j, m1, n1, m2, n2):
    """position (i, j) in ``reorder(X, m1, n1)`` -> position in X"""
    (bi, bj), (ii, jj) = divmod(int(i), int(n1)), divmod(int(j), int(n2))
    return (bi * int(m2) + ii, bj * int(n2) + jj)


def compute_banded_sparsity(n, bw):
    """raveled positions of the nonzeros of a square banded matrix of size n and bandwidth bw"""
    ij = compute_banded_sparsity_ij(n, bw).astype(np.int64)
    return ij[:, 0] * int(n) + ij[:, 1]
