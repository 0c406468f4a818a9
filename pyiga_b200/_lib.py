"""ctypes binding of the C ABI declared in ``include/pyiga_b200.h``.

Only this module knows the struct layouts.  ``load()`` opens the in-tree
``libpyiga_b200.so`` (built by ``pyiga_b200.csrc.build``) and raises if it is missing —
there is no other implementation behind the package.
"""
import ctypes as C
import os

import numpy as np

MAXDIM = 3
FORM_MASS, FORM_STIFFNESS, FORM_CUSTOM = 1, 2, 100
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpyiga_b200.so')

c_double_p = C.POINTER(C.c_double)


class AxisDesc(C.Structure):
    _fields_ = [('p_trial', C.c_int), ('nknots_trial', C.c_int), ('h_knots_trial', c_double_p),
                ('p_test', C.c_int), ('nknots_test', C.c_int), ('h_knots_test', c_double_p),
                ('nq', C.c_int), ('h_nodes', c_double_p), ('h_weights', c_double_p)]


class GeoDesc(C.Structure):
    _fields_ = [('sdim', C.c_int), ('dim', C.c_int), ('rational', C.c_int),
                ('p', C.c_int * MAXDIM), ('nknots', C.c_int * MAXDIM),
                ('h_knots', c_double_p * MAXDIM), ('h_coeffs', c_double_p)]


class Term(C.Structure):
    _fields_ = [('field', C.c_int), ('slot_test', C.c_int), ('slot_trial', C.c_int)]


class PhysTerm(C.Structure):
    _fields_ = [('slot_test', C.c_int), ('slot_trial', C.c_int), ('input', C.c_int), ('scale', C.c_double)]


class Desc(C.Structure):
    _fields_ = [('dim', C.c_int), ('axis', AxisDesc * MAXDIM), ('form', C.c_int),
                ('nfields', C.c_int), ('nterms', C.c_int), ('terms', C.POINTER(Term)),
                ('symmetric', C.c_int)]


class Info(C.Structure):
    _fields_ = [('dim', C.c_int), ('ndofs_test', C.c_int * MAXDIM), ('ndofs_trial', C.c_int * MAXDIM),
                ('nnodes', C.c_int * MAXDIM), ('nband', C.c_int * MAXDIM), ('nfields', C.c_int),
                ('fast_path', C.c_int), ('nnz', C.c_longlong), ('npoints', C.c_longlong)]


# name -> (restype, argtypes); the list is checked against the header by tests/test_abi.py
SIGNATURES = {
    'pb200_version': (C.c_int, []),
    'pb200_last_error': (C.c_char_p, []),
    'pb200_launch_count': (C.c_longlong, []),
    'pb200_asm_create': (C.c_int, [C.POINTER(Desc), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    'pb200_asm_destroy': (C.c_int, [C.c_void_p]),
    'pb200_asm_info': (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    'pb200_asm_tabulate': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pb200_asm_structure': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    'pb200_asm_bind_fields': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pb200_asm_set_geometry': (C.c_int, [C.c_void_p, C.POINTER(GeoDesc), C.c_void_p]),
    'pb200_asm_uses_fused_fields': (C.c_int, [C.c_void_p]),
    'pb200_asm_compute_fields': (C.c_int, [C.c_void_p, C.POINTER(GeoDesc), C.c_void_p]),
    'pb200_asm_compute_fields_slab': (C.c_int, [C.c_void_p, C.POINTER(GeoDesc), C.c_int, C.c_int, C.c_void_p]),
    'pb200_asm_compute_fields_general': (C.c_int, [C.c_void_p, C.POINTER(GeoDesc), C.c_void_p, C.c_int,
                                                   C.POINTER(PhysTerm), C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                                   C.c_int, C.c_void_p]),
    'pb200_asm_compute_fields_from_jacobian': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_geo_eval_grid': (C.c_int, [C.POINTER(GeoDesc), C.POINTER(C.c_int), C.POINTER(c_double_p),
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'pb200_asm_workspace_bytes': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'pb200_asm_assemble_mlb': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    'pb200_asm_assemble_mlb_entrywise': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'pb200_asm_set_option': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    'pb200_asm_set_timing': (C.c_int, [C.c_void_p, C.c_int]),
    'pb200_asm_get_timing': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_char_p, C.c_int,
                                       C.POINTER(C.c_int)]),
    'pb200_asm_vector_workspace_bytes': (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    'pb200_asm_assemble_vector': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'pb200_asm_multi_entries': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    'pb200_asm_mlstruct': (C.c_void_p, [C.c_void_p]),
    'pb200_mlstruct_create': (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    'pb200_mlstruct_destroy': (C.c_int, [C.c_void_p]),
    'pb200_mlb_to_csr': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    'pb200_mlb_matvec': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    'pb200_kron_matvec': (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_asm_rows_count': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    'pb200_asm_rows_fill': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p]),
    'pb200_csr_pattern_host': (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int]),
    'pb200_csr_restrict_workspace': (C.c_int, [C.c_longlong, C.c_int, C.POINTER(C.c_size_t)]),
    'pb200_csr_restrict_count': (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'pb200_csr_restrict_fill': (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_csr_matvec': (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_void_p, C.c_void_p]),
    'pb200_vec_gather': (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_vec_scatter': (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_basis_eval': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_probe_fp64': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    'pb200_comm_create': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    'pb200_comm_handle': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pb200_comm_open_peers': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pb200_comm_destroy': (C.c_int, [C.c_void_p]),
    'pb200_cg_window_bytes': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_size_t)]),
    'pb200_cg_create': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_void_p)]),
    'pb200_cg_destroy': (C.c_int, [C.c_void_p]),
    'pb200_cg_solve': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(C.c_double), C.c_void_p]),
    'pb200_cg_matvec': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb200_band_structure': (C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int)]),
}


class Pb200Error(RuntimeError):
    pass


def bind(path):
    """Open the shared library at `path` and attach the prototypes."""
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """The product library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Pb200Error(
                'libpyiga_b200.so is missing: build it with `python -m pyiga_b200.csrc.build` '
                '(needs nvcc); pyiga_b200 has no other implementation')
        _lib = bind(LIB_PATH)
    return _lib


def check(lib, rc):
    if rc != 0:
        msg = lib.pb200_last_error()
        msg = msg.decode() if msg else 'error %d' % rc
        if rc == -1:
            raise ValueError(msg)
        if rc == -4:
            raise NotImplementedError(msg)
        if rc == -3:
            raise MemoryError(msg)
        raise Pb200Error(msg)


def as_double_p(a):
    return a.ctypes.data_as(c_double_p)


def kv_arrays(kv):
    """(knots as contiguous float64 array, degree) of a KnotVector-like object."""
    return np.ascontiguousarray(kv.kv, dtype=np.float64), int(kv.p)


def make_geo_desc(geo):
    """GeoDesc for a spline geometry; returns (desc, keepalive list)."""
    kvs = tuple(geo.kvs)
    sdim = len(kvs)
    if sdim > MAXDIM:
        raise NotImplementedError('geometries with more than %d parameters' % MAXDIM)
    rational = bool(getattr(geo, '_rational', False)) or type(geo).__name__ == 'NurbsFunc'
    coeffs = np.ascontiguousarray(geo.coeffs, dtype=np.float64)
    N = tuple(kv.numdofs for kv in kvs)
    ncomp = int(np.prod(coeffs.shape[sdim:], dtype=np.int64)) if coeffs.ndim > sdim else 1
    dim = ncomp - 1 if rational else ncomp
    assert coeffs.shape[:sdim] == N, 'Wrong shape of coefficients'
    d = GeoDesc()
    d.sdim, d.dim, d.rational = sdim, dim, int(rational)
    keep = [coeffs]
    for k, kv in enumerate(kvs):
        kn, p = kv_arrays(kv)
        keep.append(kn)
        d.p[k] = p
        d.nknots[k] = kn.size
        d.h_knots[k] = as_double_p(kn)
    d.h_coeffs = as_double_p(coeffs)
    return d, keep
