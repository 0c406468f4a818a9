"""ctypes front-end of the C oracle (oracle_entry.c): the reference's per-entry algorithm on the
host cores.  TEST / BASELINE INFRASTRUCTURE (used by tests and by bench.py's CPU arm when the real
reference under oracle/_ref cannot be loaded)."""
import ctypes as C
import os

import numpy as np

from . import build_oracle
from . import pyiga_oracle as orc

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.oracle_multi_entries.restype = None
    return _lib


def multi_entries(prob, form, ij, nthreads=None):
    """Entries (I,J) of the mass / stiffness matrix of an oracle Problem by per-entry quadrature."""
    d = prob.dim
    ij = np.ascontiguousarray(ij, dtype=np.uintp).reshape(-1, 2)
    fields = np.ascontiguousarray((orc.fields_mass if form == 'mass' else orc.fields_stiffness)(prob.jac, prob.gw))
    N = (C.c_size_t * d)(*prob.N)
    G = (C.c_size_t * d)(*[len(g) for g in prob.grid])
    ms = [np.ascontiguousarray(m, dtype=np.int64) for m in prob.meshsupp]
    Cs = [np.ascontiguousarray(c, dtype=np.float64) for c in prob.C]
    msp = (C.c_void_p * d)(*[m.ctypes.data for m in ms])
    Cp = (C.c_void_p * d)(*[c.ctypes.data for c in Cs])
    out = np.zeros(ij.shape[0])
    lib().oracle_multi_entries(C.c_int(d), C.c_int(1 if form == 'mass' else 2), N, G, msp, Cp,
                               C.c_void_p(fields.ctypes.data), C.c_int(fields.shape[-1]), C.c_void_p(ij.ctypes.data),
                               C.c_size_t(ij.shape[0]), C.c_void_p(out.ctypes.data),
                               C.c_int(nthreads or os.cpu_count() or 1))
    return out


def assemble_csr(prob, form, nthreads=None):
    """Whole matrix like the reference driver (pyiga/assemble.py:741-754): pattern, entries, CSR."""
    import scipy.sparse
    I, J = orc.ml_nonzero(prob.bidx, prob.bs)
    vals = multi_entries(prob, form, np.column_stack((I, J)), nthreads)
    n = int(np.prod(prob.N))
    return scipy.sparse.csr_matrix((vals, (I.astype(np.int64), J.astype(np.int64))), shape=(n, n))
