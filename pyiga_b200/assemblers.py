"""Assembler objects — the operator interface of the assembly path, backed by CUDA kernels.

These classes mirror the protocol of the reference's Cython assembler classes
(``pyiga/genericasm.pxi:631-786`` base classes, ``pyiga/assemblers.pyx:26,174,1158,1324``
Mass/Stiffness 2D/3D): ``arity``, ``kvs``, ``inputs()``, ``parameters()``, ``entry``,
``multi_entries``, ``assemble_vector`` — so they can be handed to the reference's own drivers
(``pyiga.assemble.assemble_entries``) — plus a fast path (``assemble_mlb`` / ``assemble_csr``)
that produces the whole matrix by sum factorisation without ever materialising the index lists.

Construction does what the reference constructors do (``pyiga/assemblers.pyx:1336-1383``):
Gauss rule with ``nqp = max p + 1`` nodes per span, 1D basis tables (device, K1), geometry
Jacobian and coefficient fields on the Gauss grid (device, K2).
"""
import ctypes as C

import numpy as np

from . import _device, _lib
from ._mlb import DeviceStructure
from .mlmatrix import MLStructure, MLMatrix
from .quadrature import make_tensor_quadrature


def _is_spline_geo(geo):
    return hasattr(geo, 'kvs') and hasattr(geo, 'coeffs')


# ---- forms over ONE knot vector ---------------------------------------------------------------
# The device pipeline takes 2 or 3 tensor axes.  A form over (kv,) (``test/test_assemble.py:436-445``) runs on the
# lifted space (one linear element in eta) x (kv) with the geometry (x(xi), eta) and coefficients that do not
# depend on eta: the lifted matrix is  M_eta (x) A  with the 2 x 2 mass matrix of the linear element, whose entries
# sum to 1 — A is the sum of the four eta blocks (and a load vector the sum of its two eta rows).
def _lift_axis():
    from . import bspline
    return bspline.make_knots(1, 0.0, 1.0, 1)


class _LiftedGeo1D:
    """(eta, xi) -> (x(xi), eta) for a geometry that is not a spline (Jacobian evaluated on the host)"""

    def __init__(self, geo):
        self.geo, self.sdim, self.dim = geo, 2, 2

    def grid_jacobian(self, grid):
        d = np.asarray(self.geo.grid_jacobian((grid[1],)), dtype=float).reshape(len(grid[1]))
        J = np.zeros((len(grid[0]), len(grid[1]), 2, 2))
        J[..., 0, 0] = d[None, :]           # x depends on xi_0 = the last grid axis
        J[..., 1, 1] = 1.0
        return J


def _lift_geo_1d(geo):
    """the geometry (eta, xi) -> (x(xi), eta) as a spline of our own classes when `geo` is a spline"""
    from . import bspline, geometry
    if not _is_spline_geo(geo):
        return _LiftedGeo1D(geo)
    kv = tuple(geo.kvs)[0]
    kvs = (_lift_axis(), bspline.KnotVector(np.asarray(kv.kv, dtype=float), int(kv.p)))
    n = kvs[1].numdofs
    c = np.asarray(geo.coeffs, dtype=float).reshape(n, -1)
    eta = np.array([0.0, 1.0])
    rational = bool(getattr(geo, '_rational', False)) or type(geo).__name__ == 'NurbsFunc'
    if rational:        # homogeneous, premultiplied: (w x, w) -> (w x, w eta, w)
        assert c.shape[1] == 2, 'geometry of a 1D form must map to R'
        C = np.empty((2, n, 3))
        C[..., 0], C[..., 2] = c[None, :, 0], c[None, :, 1]
        C[..., 1] = eta[:, None] * c[None, :, 1]
        return geometry.NurbsFunc(kvs, C, None, premultiplied=True)
    assert c.shape[1] == 1, 'geometry of a 1D form must map to R'
    C = np.empty((2, n, 2))
    C[..., 0] = c[None, :, 0]
    C[..., 1] = eta[:, None]
    return bspline.BSplineFunc(kvs, C)


class DeviceAssembler:
    """Handle on a ``pb200_assembler``: tables, fields and kernels for one (spaces, form, geometry)."""

    def __init__(self, kvs0, kvs1, form, nqp=None, terms=None, nfields=0, symmetric=False, quad=None):
        """`quad`: optional ``{axis: (nodes, weights, nodes_per_span)}`` replacing the Gauss rule of
        an axis (boundary integrals: a single node on the boundary, ``pyiga/quadrature.py:23-31``)."""
        self.be = be = _device.backend()
        kvs0 = tuple(kvs0)
        kvs1 = tuple(kvs1) if kvs1 is not None else kvs0
        dim = len(kvs0)
        assert len(kvs1) == dim
        self.dim = dim
        self.form = form
        self.kvs = (kvs0, kvs1)
        self.nqp = int(nqp) if nqp else max(kv.p for kv in kvs0) + 1
        meshes = [kv.mesh for kv in kvs0]
        self.gaussgrid, self.gaussweights = make_tensor_quadrature(meshes, self.nqp)
        nq_axis = [self.nqp] * dim
        if quad:
            grid, wts = list(self.gaussgrid), list(self.gaussweights)
            for k, (nodes, weights, nq) in quad.items():
                grid[k], wts[k], nq_axis[k] = np.asarray(nodes, dtype=float), np.asarray(weights, dtype=float), int(nq)
            self.gaussgrid, self.gaussweights = tuple(grid), tuple(wts)
        self.same_space = all(a is b or a == b for a, b in zip(kvs0, kvs1))
        self._ctor = dict(terms=list(terms) if terms is not None else None, nfields=nfields, quad=quad)

        desc = _lib.Desc()
        desc.dim = dim
        desc.form = form
        keep = []
        for k in range(dim):
            ax = desc.axis[k]
            ku, pu = _lib.kv_arrays(kvs0[k])
            keep.append(ku)
            ax.p_trial, ax.nknots_trial, ax.h_knots_trial = pu, ku.size, _lib.as_double_p(ku)
            if not self.same_space:
                kv_, pv = _lib.kv_arrays(kvs1[k])
                keep.append(kv_)
                ax.p_test, ax.nknots_test, ax.h_knots_test = pv, kv_.size, _lib.as_double_p(kv_)
            nodes = np.ascontiguousarray(self.gaussgrid[k], dtype=np.float64)
            weights = np.ascontiguousarray(self.gaussweights[k], dtype=np.float64)
            keep += [nodes, weights]
            ax.nq, ax.h_nodes, ax.h_weights = nq_axis[k], _lib.as_double_p(nodes), _lib.as_double_p(weights)
        if form == _lib.FORM_CUSTOM:
            tarr = (_lib.Term * len(terms))(*[_lib.Term(*t) for t in terms])
            keep.append(tarr)
            desc.nfields, desc.nterms, desc.terms, desc.symmetric = nfields, len(terms), tarr, int(symmetric)
        h = C.c_void_p()
        _device.check(be.lib.pb200_asm_create(C.byref(desc), be.device_index, be.stream(), C.byref(h)))
        self.handle = h
        info = _lib.Info()
        _device.check(be.lib.pb200_asm_info(h, C.byref(info)))
        self.ndofs_test = tuple(info.ndofs_test[k] for k in range(dim))
        self.ndofs_trial = tuple(info.ndofs_trial[k] for k in range(dim))
        self.nnodes = tuple(info.nnodes[k] for k in range(dim))
        self.nband = tuple(info.nband[k] for k in range(dim))
        self.nfields = info.nfields
        self.fast_path = bool(info.fast_path)
        self.nnz = int(info.nnz)
        self.npoints = int(info.npoints)
        self._fields = None         # allocated when a kernel needs the field array (see `fields`)
        self._geo_bound = None      # spline geometry bound on the device, fields not evaluated yet
        self._fields_rows = None    # rows of axis 0 whose Gauss planes hold current fields ('all' or (ra, rb))
        self._structure = None
        self._dstruct = None
        self._row_start0 = None
        # tuning switches for experiments, e.g. PB200_OPTS="fused_plans=1,lane_v1=1"
        import os
        for item in filter(None, os.environ.get('PB200_OPTS', '').split(',')):
            k, _, v = item.partition('=')
            self.set_option(k.strip(), int(v or 1))

    def __del__(self):
        try:
            if getattr(self, 'handle', None) is not None:
                self.be.lib.pb200_asm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- structure ---------------------------------------------------------------------------
    def bidx(self, axis):
        out = np.empty((self.nband[axis], 2), dtype=np.uint32)
        _device.check(self.be.lib.pb200_asm_structure(self.handle, axis, out.ctypes.data))
        return out

    @property
    def structure(self):
        if self._structure is None:
            bs = tuple((self.ndofs_test[k], self.ndofs_trial[k]) for k in range(self.dim))
            self._structure = MLStructure(bs, tuple(self.bidx(k) for k in range(self.dim)))
        return self._structure

    @property
    def device_structure(self):
        if self._dstruct is None:
            h = self.be.lib.pb200_asm_mlstruct(self.handle)
            self._dstruct = DeviceStructure(structure=self.structure, borrowed=C.c_void_p(h), owner=self)
        return self._dstruct

    def row_start0(self):
        if self._row_start0 is None:
            b0 = self.structure.bidx[0][:, 0]
            self._row_start0 = np.searchsorted(b0, np.arange(self.ndofs_test[0] + 1))
        return self._row_start0

    # ---- fields ------------------------------------------------------------------------------
    @property
    def fields(self):
        """the coefficient-field array F[c][g0][g1][g2] on the device, allocated on first use: the
        fused 3D mass / stiffness pipeline never needs it"""
        if self._fields is None:
            self._fields = self.be.empty(self.nfields * self.npoints)
            _device.check(self.be.lib.pb200_asm_bind_fields(self.handle, self.be.ptr(self._fields)))
        return self._fields

    @fields.setter
    def fields(self, buf):
        self._fields = buf
        self._fields_rows = 'all'

    def uses_fused_fields(self):
        """True if assemble_mlb evaluates geometry and fields inside its first stage (no K2)"""
        return bool(self.be.lib.pb200_asm_uses_fused_fields(self.handle))

    def need_fields(self, rows=None):
        """make sure the field array holds the fields of the bound geometry on the Gauss planes the
        rows `rows` of the first axis see (K2 runs now if it has not yet)"""
        if self._geo_bound is None:
            if self._fields_rows is None:
                raise RuntimeError('fields have not been computed')
            return
        have = self._fields_rows
        if have == 'all' or (have is not None and rows is not None and have[0] <= rows[0] and rows[1] <= have[1]):
            return
        be = self.be
        self.fields
        if rows is None:
            _device.check(be.lib.pb200_asm_compute_fields(self.handle, None, be.stream()))
            self._fields_rows = 'all'
        else:
            _device.check(be.lib.pb200_asm_compute_fields_slab(self.handle, None, rows[0], rows[1], be.stream()))
            self._fields_rows = tuple(rows)

    def tabulate(self):
        _device.check(self.be.lib.pb200_asm_tabulate(self.handle, self.be.stream()))

    def compute_fields(self, geo, rows=None):
        """K2: geometry Jacobian + form coefficients at every Gauss point (or only on the Gauss
        planes seen by the rows `rows` of the first axis)."""
        be = self.be
        if _is_spline_geo(geo):
            # bind the geometry; K2 itself runs only if a kernel asks for the field array
            # (`need_fields`): the fused first stage of the 3D pipeline evaluates the fields itself
            desc, keep = _lib.make_geo_desc(geo)
            _device.check(be.lib.pb200_asm_set_geometry(self.handle, C.byref(desc), be.stream()))
            self._geo_bound, self._fields_rows = geo, None
            if not self.uses_fused_fields():
                self.need_fields(rows)
        else:
            # geometry given as an arbitrary Python object: evaluate its Jacobian on the host, as
            # the reference does for every geometry, and upload it
            jac = np.ascontiguousarray(geo.grid_jacobian(self.gaussgrid), dtype=np.float64)
            assert jac.shape == self.nnodes + (self.dim, self.dim), 'geo.grid_jacobian returned a wrong shape'
            d_jac = be.from_host(jac.ravel())
            self.fields
            _device.check(be.lib.pb200_asm_compute_fields_from_jacobian(self.handle, be.ptr(d_jac), be.stream()))
            be.synchronize()
            self._geo_bound, self._fields_rows = None, 'all'

    def fields_host(self):
        self.need_fields()
        return self.be.to_host(self.fields).reshape((self.nfields,) + self.nnodes)

    # ---- assembly ----------------------------------------------------------------------------
    def workspace_bytes(self, rows=None):
        ra, rb = (0, self.ndofs_test[0]) if rows is None else rows
        n = C.c_size_t()
        _device.check(self.be.lib.pb200_asm_workspace_bytes(self.handle, ra, rb, C.byref(n)))
        return int(n.value)

    def slab_size(self, rows=None):
        ra, rb = (0, self.ndofs_test[0]) if rows is None else rows
        rs = self.row_start0()
        return int(rs[rb] - rs[ra]) * int(np.prod(self.nband[1:], dtype=np.int64))

    def row_chunks(self, rows=None, budget_bytes=None):
        """Split a row slab of axis 0 into chunks whose workspace fits `budget_bytes`."""
        ra, rb = (0, self.ndofs_test[0]) if rows is None else rows
        if budget_bytes is None or self.workspace_bytes((ra, rb)) <= budget_bytes:
            return [(ra, rb)]
        chunks, a = [], ra
        while a < rb:
            lo, hi = a + 1, rb
            if self.workspace_bytes((a, lo)) > budget_bytes:
                raise MemoryError('workspace budget too small for a single row of the first axis')
            while lo < hi:          # largest b with workspace(a, b) <= budget
                mid = (lo + hi + 1) // 2
                if self.workspace_bytes((a, mid)) <= budget_bytes:
                    lo = mid
                else:
                    hi = mid - 1
            chunks.append((a, lo))
            a = lo
        return chunks

    def assemble_mlb(self, rows=None, out=None, workspace=None, budget_bytes=None, entrywise=False):
        """MLB value tensor of the row slab `rows` of the first axis (default: all rows) as a flat
        device buffer; layout ``data[mu0 - mu0_begin, mu1, mu2]``."""
        be = self.be
        ra, rb = (0, self.ndofs_test[0]) if rows is None else rows
        total = self.slab_size((ra, rb))
        if out is None:
            out = be.empty(total)
        rs = self.row_start0()
        inner = int(np.prod(self.nband[1:], dtype=np.int64))
        if entrywise:
            self.need_fields()
            _device.check(be.lib.pb200_asm_assemble_mlb_entrywise(self.handle, ra, rb, be.ptr(out), be.stream()))
            return out
        if not self.uses_fused_fields():
            self.need_fields((ra, rb))
        if budget_bytes is None and workspace is None:
            budget_bytes = max(be.free_bytes() - (1 << 30), 1 << 28) if self.fast_path else None
        elif workspace is not None:
            budget_bytes = be.nbytes(workspace)
        for (a, b) in self.row_chunks((ra, rb), budget_bytes):
            need = self.workspace_bytes((a, b))
            ws = workspace if workspace is not None else (be.empty(need, np.uint8) if need else None)
            off = int(rs[a] - rs[ra]) * inner
            dst = be.ptr(out) + 8 * off
            _device.check(be.lib.pb200_asm_assemble_mlb(self.handle, a, b, dst, be.ptr(ws), need if ws is not None else 0,
                                                         be.stream()))
            if workspace is None and ws is not None:
                be.synchronize()    # the temporary workspace is released when `ws` goes out of scope
        return out

    def assemble_vector_device(self):
        """load vector of a linear form as a flat device buffer (C order of the test space)"""
        be = self.be
        self.need_fields()
        terms = self._ctor['terms']
        if not self.fast_path and self.form == _lib.FORM_CUSTOM and terms and all(t[2] < 0 for t in terms):
            # degrees without instantiated vector kernels (p = 5 ...): the B-splines of the trial space sum to 1, so
            # the load vector of  sum_t c_t d^bt v  is the vector of row sums of the bilinear form
            # sum_t c_t d^bt v * u, which the per-entry quadrature kernel serves for every degree
            twin = DeviceAssembler(self.kvs[0], self.kvs[1], _lib.FORM_CUSTOM, nqp=self.nqp,
                                   terms=[(f, bp, 0) for f, bp, _ in terms], nfields=self._ctor['nfields'], quad=self._ctor['quad'])
            twin.fields = self.fields
            _device.check(be.lib.pb200_asm_bind_fields(twin.handle, be.ptr(self.fields)))
            mlb = twin.assemble_mlb(entrywise=True)
            ones = be.from_host(np.ones(int(np.prod(twin.ndofs_trial, dtype=np.int64))))
            out = twin.device_structure.matvec_device(mlb, ones)
            be.synchronize()
            return out
        n = C.c_size_t()
        _device.check(be.lib.pb200_asm_vector_workspace_bytes(self.handle, C.byref(n)))
        ws = be.empty(n.value, np.uint8)
        out = be.empty(int(np.prod(self.ndofs_test, dtype=np.int64)))
        _device.check(be.lib.pb200_asm_assemble_vector(self.handle, be.ptr(out), be.ptr(ws), n.value, be.stream()))
        be.synchronize()
        return out

    def set_option(self, name, value):
        _device.check(self.be.lib.pb200_asm_set_option(self.handle, name.encode(), int(value)))

    def set_timing(self, enable=True):
        _device.check(self.be.lib.pb200_asm_set_timing(self.handle, int(enable)))

    def stage_times(self):
        """[(kernel name, ms)] of the last assemble call (needs set_timing(True))."""
        ms = (C.c_float * 16)()
        names = C.create_string_buffer(512)
        n = C.c_int()
        _device.check(self.be.lib.pb200_asm_get_timing(self.handle, 16, ms, names, 512, C.byref(n)))
        labels = [t for t in names.value.decode().split(';') if t]
        return [(labels[i], float(ms[i])) for i in range(min(n.value, len(labels)))]

    def rows_csr_device(self, rows):
        """Device CSR arrays (indptr, indices, values) of the given matrix rows, values by per-entry
        quadrature (``pb200_asm_rows_count`` / ``pb200_asm_rows_fill``)."""
        be = self.be
        self.need_fields()
        rows = np.ascontiguousarray(rows, dtype=np.int64).ravel()
        n = rows.size
        h_indptr = np.zeros(n + 1, dtype=np.int64)
        _device.check(be.lib.pb200_asm_rows_count(self.handle, rows.ctypes.data, n, h_indptr.ctypes.data))
        nnz = int(h_indptr[-1])
        ncols = int(np.prod([kv.numdofs for kv in self.kvs[0]], dtype=np.int64))
        idt = np.int32 if max(nnz, ncols) < 2 ** 31 else np.int64
        d_indptr = be.from_host(h_indptr.astype(idt))
        d_indices, d_values = be.empty(nnz, idt), be.empty(nnz)
        if n:
            d_rows = be.from_host(rows)
            _device.check(be.lib.pb200_asm_rows_fill(self.handle, be.ptr(d_rows), n, be.ptr(d_indptr), be.ptr(d_indices),
                                                     be.ptr(d_values), np.dtype(idt).itemsize, be.stream()))
            be.synchronize()
        return d_indptr, d_indices, d_values

    def multi_entries_device(self, ij):
        be = self.be
        self.need_fields()
        ij = np.ascontiguousarray(ij, dtype=np.uint64).reshape(-1, 2)
        n = ij.shape[0]
        out = be.empty(n)
        if n:
            d_ij = be.from_host(ij.ravel())
            _device.check(be.lib.pb200_asm_multi_entries(self.handle, be.ptr(d_ij), n, be.ptr(out), be.stream()))
            be.synchronize()
        return out


class _AssemblerProtocol:
    """Methods every device assembler shares (``pyiga/genericasm.pxi:662-786``)."""
    arity = 2

    def entry(self, i, j):
        """A[i, j] = a(phi_j, phi_i): row index decoded in the test space, column in the trial space."""
        return float(self.multi_entries(np.array([[i, j]], dtype=np.uint64))[0])

    def entry1(self, i):
        return 0.0

    def multi_entries1(self, indices):
        return None

    def multi_entries(self, indices):
        """Entries at the given ``(N, 2)`` index array or iterable of ``(i, j)`` pairs; pairs outside
        the sparsity pattern give 0 (``pyiga/genericasm.pxi:722-758``)."""
        if not isinstance(indices, np.ndarray):
            indices = np.array(list(indices), dtype=np.uint64)
        indices = np.asarray(indices, dtype=np.uint64).reshape(-1, 2)
        return self.dev.be.to_host(self.dev.multi_entries_device(indices))

    def assemble_vector(self):
        return None     # arity 2

    def entry_func_ptr(self):
        """PyCapsule named ``"entryfunc"`` holding a C function ``double (*)(size_t i, size_t j, void* self)``
        (``pyiga/genericasm.pxi:780-786``, consumed by the low-rank assembler
        ``pyiga/fast_assemble_cy.pyx:99-113``, which passes the assembler object itself as `self`).
        A per-entry callback cannot be served from the GPU entry by entry, so the first call assembles
        the matrix once on the device; the callback then reads the host CSR copy."""
        import ctypes as C_
        if getattr(self, '_entry_capsule', None) is None:
            A = self._entry_matrix()
            indptr, indices, data = A.indptr, A.indices, A.data

            def lookup(i, j, _self):
                a, b = indptr[i], indptr[i + 1]
                k = a + np.searchsorted(indices[a:b], j)
                return float(data[k]) if k < b and indices[k] == j else 0.0
            proto = C_.CFUNCTYPE(C_.c_double, C_.c_size_t, C_.c_size_t, C_.c_void_p)
            self._entry_cfunc = proto(lookup)          # keeps the trampoline alive
            new = C_.pythonapi.PyCapsule_New
            new.restype, new.argtypes = C_.py_object, [C_.c_void_p, C_.c_char_p, C_.c_void_p]
            self._entry_capsule_name = b'entryfunc'
            self._entry_capsule = new(C_.cast(self._entry_cfunc, C_.c_void_p), self._entry_capsule_name, None)
        return self._entry_capsule

    def _entry_matrix(self):
        A = self.assemble_csr()
        A.sort_indices()
        return A

    # ---- fast path ---------------------------------------------------------------------------
    def assemble_mlb(self, **kw):
        """The whole matrix as an :class:`~pyiga_b200.mlmatrix.MLMatrix` whose values stay on the device."""
        data = self.dev.assemble_mlb(**kw)
        M = MLMatrix(structure=self.dev.structure, data=data)
        # the matrix owns its band tables: borrowing the assembler's would keep the assembler's field buffers and
        # scratch alive on the GPU for the lifetime of the matrix
        own = DeviceStructure(structure=self.dev.structure)
        M._dev = own if own.supported else self.dev.device_structure
        return M

    def assemble_csr(self, **kw):
        """The whole matrix as ``scipy.sparse.csr_matrix`` (float64 data, int32 indices unless
        nnz >= 2^31, sorted).  On the GPU the sum-factorised path delivers it chunk by chunk with the
        device->host copies overlapped (:mod:`pyiga_b200._hostcsr`)."""
        dev = self.dev
        if not kw and dev.fast_path and dev.be.name == 'cuda' and dev.dim >= 2:
            from ._hostcsr import assemble_csr_matrix
            return assemble_csr_matrix(dev)
        data = dev.assemble_mlb(**kw)
        return dev.device_structure.to_csr(data)


class _ScalarAssemblerBase(_AssemblerProtocol):
    """The predefined scalar assemblers."""
    _form = None
    _dim = None

    @classmethod
    def inputs(cls):
        return {'geo': (cls._dim,)}

    @classmethod
    def parameters(cls):
        return {}

    def __init__(self, kvs0, geo):
        d = self._dim
        assert geo.sdim == d, "Geometry has wrong source dimension"
        assert geo.dim == d, "Geometry has wrong dimension"
        kvs0 = tuple(kvs0)
        assert len(kvs0) == d, "Assembler requires %d knot vectors" % d
        self.arity = 2
        self.nqp = max(kv.p for kv in kvs0) + 1
        self._geo = geo
        self.dev = DeviceAssembler(kvs0, kvs0, self._form, nqp=self.nqp)
        self.dev.compute_fields(geo)
        self.kvs = (kvs0, kvs0)
        self.gaussgrid = self.dev.gaussgrid


class _FormBlock:
    """One scalar form  sum_t c_t d^bt v d^bu u  on the device: tables, coefficient upload, fields."""

    def __init__(self, kvs, nqp, dim, arity, coefs, geo, gaussgrid, quad=None, kvs_test=None, host_pullback=False):
        self.dim, self.arity, self.kvs = dim, arity, kvs
        self.host_pullback = host_pullback
        self.gaussgrid = gaussgrid
        self._grid_shape = tuple(len(g) for g in gaussgrid)
        self.keys = sorted(coefs)
        pairs = set()
        for (bt, bu) in self.keys:
            for bp in ([0] if bt == 0 else range(1, dim + 1)):
                for ap in ([-1] if bu < 0 else [0] if bu == 0 else range(1, dim + 1)):
                    pairs.add((bp, ap))
        terms = [(f, bp, ap) for f, (bp, ap) in enumerate(sorted(pairs))]
        self.terms = terms
        self.dev = DeviceAssembler(kvs, kvs_test or kvs, _lib.FORM_CUSTOM, nqp=nqp, terms=terms, nfields=len(terms), quad=quad)
        self.compute_fields(coefs, geo)

    def _compute_fields_host(self, coefs, geo):
        """Pull the physical terms back to the parameter domain with numpy and upload the fields:
            C[bp][ap] = gw * sqrt(det J^T J) * sum_t c_t T[bt][bp] T[bu][ap],
        T[0][0] = 1, T[1+a][1+k] = G[a][d-1-k], G = J (J^T J)^-1 (the tangential gradient is G grad_xi).
        Used for manifolds (J is (d+1) x d); the square case runs on the device (PbProgGeneral)."""
        dev, be = self.dev, self.dev.be
        d = self.dim
        J = np.asarray(geo.grid_jacobian(self.gaussgrid), dtype=float)          # grid + (gd, d), xi_0 = last axis
        gd = J.shape[-2]
        JtJ = np.einsum('...ai,...aj->...ij', J, J)
        G = np.einsum('...ai,...ij->...aj', J, np.linalg.inv(JtJ))
        W = np.sqrt(np.linalg.det(JtJ))
        for k in range(d):
            shp = [1] * d
            shp[k] = -1
            W = W * np.asarray(dev.gaussweights[k]).reshape(shp)
        T = np.zeros(self._grid_shape + (gd + 1, d + 1))
        T[..., 0, 0] = 1.0
        for a in range(gd):
            for k in range(d):
                T[..., 1 + a, 1 + k] = G[..., a, d - 1 - k]
        fields = np.zeros((len(self.terms),) + self._grid_shape)
        for f, bp, ap in self.terms:
            acc = 0.0
            for (bt, bu) in self.keys:
                c = coefs[(bt, bu)]
                cv = c.scale * (np.broadcast_to(c.arr, self._grid_shape) if c.arr is not None else 1.0)
                tu = 1.0 if ap < 0 else T[..., bu, ap]
                acc = acc + cv * T[..., bt, bp] * tu
            fields[f] = W * acc
        buf = be.from_host(np.ascontiguousarray(fields).ravel())
        dev.fields = buf
        dev._geo_bound = None
        _device.check(be.lib.pb200_asm_bind_fields(dev.handle, be.ptr(buf)))

    def compute_fields(self, coefs, geo):
        dev, be = self.dev, self.dev.be
        if sorted(coefs) != self.keys:
            raise RuntimeError('update() changed the structure of the form')
        if self.host_pullback:
            return self._compute_fields_host(coefs, geo)
        arrays, index = [], {}
        phys = (_lib.PhysTerm * len(coefs))()
        for t, key in enumerate(self.keys):
            c = coefs[key]
            inp = -1
            if c.arr is not None:
                if id(c.arr) not in index:
                    index[id(c.arr)] = len(arrays)
                    if hasattr(c.arr, 'device') and hasattr(c.arr, 'expand'):      # already on the GPU
                        arrays.append(c.arr.expand(self._grid_shape).contiguous().reshape(-1))
                    else:
                        arr = np.ascontiguousarray(np.broadcast_to(c.arr, self._grid_shape), dtype=np.float64)
                        arrays.append(be.from_host(arr.ravel()))
                inp = index[id(c.arr)]
            phys[t] = _lib.PhysTerm(key[0], key[1], inp, c.scale)
        ptrs = (C.c_void_p * max(len(arrays), 1))(*[be.ptr(a) for a in arrays])
        dev.fields                  # allocate and bind the field array
        dev._geo_bound, dev._fields_rows = None, 'all'
        if _is_spline_geo(geo):
            desc, keep = _lib.make_geo_desc(geo)
            _device.check(be.lib.pb200_asm_compute_fields_general(dev.handle, C.byref(desc), None, len(coefs), phys,
                                                                   len(arrays), ptrs, -1, -1, be.stream()))
        else:
            jac = np.ascontiguousarray(geo.grid_jacobian(self.gaussgrid), dtype=np.float64)
            d_jac = be.from_host(jac.ravel())
            _device.check(be.lib.pb200_asm_compute_fields_general(dev.handle, None, be.ptr(d_jac), len(coefs), phys,
                                                                   len(arrays), ptrs, -1, -1, be.stream()))
        be.synchronize()        # the uploaded coefficient arrays may be released now


class GenericFormAssembler(_AssemblerProtocol):
    """Assembler of a bilinear or linear form described by a :class:`~pyiga_b200.vform.VForm`
    (base of the classes returned by :func:`~pyiga_b200.vform.compile_vform`).

    Mirrors the generated assembler classes of the reference (``pyiga/codegen/cython.py:509-744``):
    construction evaluates the input functions on the Gauss grid (physical callables at the mapped
    points, spline functions on the parameter grid), ``update(name=func)`` re-evaluates one input.
    Vector-valued forms (``pyiga/genericasm.pxi:790-958``) are assembled block by block: the pair
    (test component, trial component) is one scalar form.
    """
    _vf = None
    _lift1d = False

    def __init__(self, kvs, kvs_test=None, bbox=None, **args):
        vf = self._vf
        kvs = tuple(kvs)
        self.bbox = bbox            # on-demand box of the reference's generated classes: not needed here
        d = vf.dim
        assert len(kvs) == d, "Assembler requires %d knot vectors" % d
        self._lift1d = (d == 1)
        if vf.num_spaces() == 2:
            # Petrov-Galerkin: trial functions in `kvs` (space 0, columns), test functions in
            # `kvs_test` (space 1, rows), on the same mesh (``pyiga/assemble.py:947-951``)
            assert kvs_test is not None and len(kvs_test) == d, "Assembler requires %d knot vectors" % d
            kvs_test = tuple(kvs_test)
            assert all(np.array_equal(a.mesh, b.mesh) for a, b in zip(kvs, kvs_test)), 'both spaces must share the mesh'
        else:
            kvs_test = None
        geo = args['geo']
        assert geo.sdim == d, "Geometry has wrong source dimension"
        assert geo.dim == getattr(vf, 'geo_dim', d), "Geometry has wrong dimension"
        self._surface = geo.dim != d            # manifold in R^(d+1): fields are pulled back on the host
        self.arity = vf.arity
        self.nqp = max(kv.p for kv in kvs + (kvs_test or ())) + 1
        self.kvs = (kvs, kvs_test or kvs)
        self._geo = geo
        self._args = dict(args)
        self.gaussgrid, _ = make_tensor_quadrature([kv.mesh for kv in kvs], self.nqp)
        quad = None
        self._bd = None
        if getattr(vf, 'boundary', False):
            if kvs_test is not None:
                raise NotImplementedError('boundary integrals over two different spaces')
            kvs, quad = self._setup_boundary(kvs, args.get('boundary'))
        self._grid_shape = tuple(len(g) for g in self.gaussgrid)
        self._X = None
        self._env = {}
        if self._bd is not None:
            self._boundary_fields()
        if self._surface:
            self._surface_fields()
        for name, shape, physical, _upd in vf.inputs:
            self._env[name] = self._eval_input(args[name], shape, physical)
        for name, shape in vf.params:
            self._env[name] = np.asarray(args[name], dtype=float)
        if any(_mentions_x(e) for e in vf.exprs):
            if self._device_inputs():
                Xd = self._physical_points(device=True)
                self._env['@x'] = Xd.movedim(-1, 0)
            else:
                X = self._physical_points()
                self._env['@x'] = np.stack([X[..., i] for i in range(X.shape[-1])])
        nc_u, nc_v = vf.numcomp
        self._vec = bool(vf.vec)
        self._nc = (nc_u or 1, nc_v or 1)          # (trial, test) components
        if self._vec:
            self.num_components = lambda: self._nc
        self.blocks = {}
        if self._lift1d:
            if self._bd is not None or self._surface or self._vec:
                raise NotImplementedError('boundary, surface and vector-valued forms over one knot vector')
            lift = _lift_axis()
            kvs, kvs_test, d = (lift,) + kvs, ((lift,) + kvs_test if kvs_test is not None else None), 2
            self._lift_grid = make_tensor_quadrature([kv.mesh for kv in kvs], self.nqp)[0]
        for blk, coefs in self._analyse().items():
            self.blocks[blk] = _FormBlock(kvs, self.nqp, d, self.arity, coefs, self._block_geo(), self._block_grid(), quad=quad,
                                          kvs_test=kvs_test, host_pullback=self._surface)
        first = next(iter(self.blocks.values()))
        self.dev = self.blocks.get((0, 0) if self.arity == 2 else (0, None), first).dev

    def _block_geo(self):
        """the geometry the device blocks are built on (lifted for forms over one knot vector)"""
        return _lift_geo_1d(self._geo) if self._lift1d else self._geo

    def _block_grid(self):
        return self._lift_grid if self._lift1d else self.gaussgrid

    # ---- forms over one knot vector: marginals of the lifted results -----------------------------
    def _marginal_matrix(self):
        """scipy CSR matrix of a scalar form over (kv,): the four eta blocks of the lifted MLB tensor summed"""
        import scipy.sparse
        dev = self.dev
        S = dev.structure
        data = dev.be.to_host(dev.assemble_mlb()).reshape(tuple(len(b) for b in S.bidx))
        bidx = np.asarray(S.bidx[1], dtype=np.int64)
        shape = (self.kvs[1][0].numdofs, self.kvs[0][0].numdofs)
        A = scipy.sparse.csr_matrix((data.sum(axis=0), (bidx[:, 0], bidx[:, 1])), shape=shape)
        A.sort_indices()
        return A

    # ---- boundary integrals --------------------------------------------------------------------
    def _setup_boundary(self, kvs, boundary):
        """Integrals over one side of the patch (``pyiga/codegen/cython.py:549-590``,
        ``quadrature.py:23-31``): the Gauss rule of the normal axis is the single boundary point with
        weight 1 and only the basis function that does not vanish there takes part.  The device
        tables get a stand-in for the normal axis — a linear one-span knot vector whose boundary
        function has the same value (1) and derivative at the boundary point — and the results are
        sliced down to that function afterwards."""
        from . import bspline
        assert boundary is not None, "a boundary integral needs the `boundary` argument"
        bdax, bdside = bspline._parse_bdspec(boundary, len(kvs))
        kvn = kvs[bdax]
        a, b = kvn.support()
        pt = a if bdside == 0 else b
        d1 = bspline.active_deriv(kvn, float(pt), 1)[1]
        slope = abs(d1[0] if bdside == 0 else d1[-1])
        if slope == 0.0:            # degree 0: constant function, any length will do
            slope = 1.0 / (b - a)
        L = 1.0 / slope
        fake = bspline.KnotVector(np.array([a, a, a + L, a + L] if bdside == 0 else [b - L, b - L, b, b]), 1)
        self._bd = (bdax, bdside)
        self.kvs = tuple(tuple(kv for k, kv in enumerate(kvs) if k != bdax) for _ in range(2))
        grid = list(self.gaussgrid)
        grid[bdax] = np.array([pt], dtype=float)
        self.gaussgrid = tuple(grid)
        quad = {bdax: (grid[bdax], np.ones(1), 1)}
        return kvs[:bdax] + (fake,) + kvs[bdax + 1:], quad

    def _boundary_fields(self):
        """surface measure relative to the volume measure, and the outer unit normal, on the face
        grid: with g = J^-T e_n (the physical gradient of the normal parameter),
        ds = |det J| |g| dxi_tangential and n = -+ g / |g|."""
        d = self._vf.dim
        bdax, bdside = self._bd
        J = np.asarray(self._geo.grid_jacobian(self.gaussgrid), dtype=float)    # grid + (d, d), xi_0 = last axis
        g = np.linalg.inv(J)[..., d - 1 - bdax, :]
        norm = np.sqrt((g * g).sum(axis=-1))
        self._env['@ds'] = np.ascontiguousarray(norm)
        sign = -1.0 if bdside == 0 else 1.0
        self._env['@n'] = np.stack([np.ascontiguousarray(sign * g[..., i] / norm) for i in range(d)])

    def _surface_fields(self):
        """surface integrals over a manifold (`ds` without `boundary`, geo: R^d -> R^(d+1)): the measure
        sqrt(det J^T J) is applied by the host pullback of the fields; the unit normal (surfaces in
        R^3, curves in R^2) is a coefficient field (``pyiga/vform.py:202-211``)."""
        J = np.asarray(self._geo.grid_jacobian(self.gaussgrid), dtype=float)    # grid + (d+1, d)
        self._env['@ds'] = 1.0
        if J.shape[-2:] == (3, 2):
            nrm = np.cross(J[..., :, 0], J[..., :, 1])
        elif J.shape[-2:] == (2, 1):
            nrm = np.stack([J[..., 1, 0], -J[..., 0, 0]], axis=-1)
        else:
            return
        nrm = nrm / np.sqrt((nrm * nrm).sum(axis=-1, keepdims=True))
        self._env['@n'] = np.stack([np.ascontiguousarray(nrm[..., i]) for i in range(nrm.shape[-1])])

    def _bd_select(self, arr, arity):
        """slice the band / dof axis of the stand-in normal axis down to the boundary function"""
        bdax, bdside = self._bd
        idx = [slice(None)] * arr.ndim
        if arity == 2:
            idx[bdax] = 0 if bdside == 0 else 3     # band entries of the 2 x 2 dense level: (0,0) ... (1,1)
        else:
            idx[bdax] = slice(0, 1) if bdside == 0 else slice(1, 2)
        return np.ascontiguousarray(arr[tuple(idx)])

    def _bd_structure(self):
        from .mlmatrix import MLStructure
        return MLStructure.from_kvs(self.kvs[0], self.kvs[1])

    def _bd_matrix(self, b):
        """MLMatrix over the boundary space of one scalar block"""
        be = b.dev.be
        full = be.to_host(b.dev.assemble_mlb()).reshape(tuple(len(x) for x in b.dev.structure.bidx))
        return MLMatrix(structure=self._bd_structure(), data=self._bd_select(full, 2))

    # ---- input evaluation (host side, like the reference) ------------------------------------
    def _device_inputs(self):
        """Coefficient arrays stay on the GPU (CUDA backend, volume forms on spline geometries): the
        physical Gauss points are evaluated there and the callable runs on the device tensors, which answer
        arithmetic and the numpy protocols (``np.sin``, ``np.where`` ...; :mod:`pyiga_b200._devarray`); the host
        evaluation of the reference (``pyiga/codegen/cython.py:465-484``) is the fall-back for callables that
        need more (``math.*``, reductions, ``np.asarray``)."""
        return (self.dev_backend_name == 'cuda' and self._bd is None and not self._surface and _is_spline_geo(self._geo)
                and len(self.gaussgrid) >= 2)

    @property
    def dev_backend_name(self):
        return _device.backend().name

    def _physical_points(self, device=False):
        if device:
            if getattr(self, '_Xd', None) is None:
                self._Xd = _device.eval_spline_on_grid(self._geo, self.gaussgrid, 'value', keep_on_device=True)
            return self._Xd
        if self._X is None:
            geo = self._geo
            self._X = np.asarray(geo.grid_eval(self.gaussgrid))     # device evaluation for spline geometries
        return self._X

    def _eval_input(self, f, shape, physical):
        from .vform import _grid_values
        if not physical:        # parametric: a spline function, or a plain callable of the parameters (``pyiga/utils.py:33-41``)
            if not hasattr(f, 'grid_eval'):
                f = _ParametricCallable(f)
            vals = np.asarray(f.grid_eval(self.gaussgrid))
            vals = np.moveaxis(vals, tuple(range(len(self._grid_shape), vals.ndim)), tuple(range(len(shape)))) \
                if shape else vals
            if self._device_inputs():
                vals = _device.backend().from_host(np.ascontiguousarray(vals)).reshape(vals.shape)
            return vals
        if self._device_inputs():
            Xd = self._physical_points(device=True)
            coords = tuple(Xd[..., i] for i in range(Xd.shape[-1]))
            try:
                vals = _grid_values(f, shape, coords, self._grid_shape)
            except Exception:       # the callable needs a real numpy array (math.*, reductions, ...): evaluate it on the host
                vals = None
            if vals is not None:
                if shape == ():
                    return vals
                import torch
                return torch.stack([v.contiguous() for v in vals.ravel()]).reshape(shape + self._grid_shape)
            X = self._physical_points()
            coords = tuple(X[..., i] for i in range(X.shape[-1]))
            vals = _grid_values(f, shape, coords, self._grid_shape)
            be = _device.backend()
            if shape == ():
                return be.from_host(np.ascontiguousarray(vals)).reshape(self._grid_shape)
            arr = np.stack([np.ascontiguousarray(v) for v in vals.ravel()]).reshape(shape + self._grid_shape)
            return be.from_host(arr).reshape(arr.shape)
        X = self._physical_points()
        coords = tuple(X[..., i] for i in range(X.shape[-1]))
        vals = _grid_values(f, shape, coords, self._grid_shape)
        if shape == ():
            return np.ascontiguousarray(vals)
        return np.stack([np.ascontiguousarray(v) for v in vals.ravel()]).reshape(shape + self._grid_shape)

    def _analyse(self):
        """evaluate the expression trees symbolically:
        {(test comp, trial comp): {(test deriv slot, trial deriv slot): Coef}}"""
        total = None
        for e in self._vf.exprs:
            val = e.ev(self._env)[()]
            total = val if total is None else total + val
        blocks = {}
        for (st, su), c in total.terms.items():
            if self.arity == 1:
                if st is None or su is not None:
                    raise ValueError('a linear form must be linear in its basis function')
                blk, key = (st[0], None), (st[1], -1)
            else:
                if st is None or su is None:
                    raise ValueError('the form must be linear in both u and v (term without %s)' % ('v' if st is None else 'u'))
                blk, key = (st[0], su[0]), (st[1], su[1])
            if not c.is_zero():
                blocks.setdefault(blk, {})[key] = c
        if not blocks:
            raise ValueError('the form is identically zero')
        return blocks

    def _recompute(self):
        blocks = self._analyse()
        if sorted(blocks, key=str) != sorted(self.blocks, key=str):
            raise RuntimeError('update() changed the structure of the form')
        for blk, coefs in blocks.items():
            self.blocks[blk].compute_fields(coefs, self._block_geo())

    # ---- vector-valued forms -----------------------------------------------------------------
    def _block_mlb(self):
        """device MLB tensors of all blocks: {(ct, cu): buffer}"""
        return {blk: b.dev.assemble_mlb() for blk, b in self.blocks.items()}

    def multi_blocks(self, indices):
        """blocks[k][ct][cu] = A[(i_k, ct), (j_k, cu)] (``pyiga/genericasm.pxi:815-850``)"""
        if not isinstance(indices, np.ndarray):
            indices = np.array(list(indices), dtype=np.uint64)
        indices = np.asarray(indices, dtype=np.uint64).reshape(-1, 2)
        nc_u, nc_v = self._nc
        out = np.zeros((indices.shape[0], nc_v, nc_u))
        for (ct, cu), b in self.blocks.items():
            out[:, ct, cu] = b.dev.be.to_host(b.dev.multi_entries_device(indices))
        return out

    def assemble_mlb(self, layout='packed', **kw):
        if self._lift1d:
            from .mlmatrix import MLStructure
            A = self._marginal_matrix()
            S = MLStructure.from_kvs(self.kvs[1], self.kvs[0])
            bidx = np.asarray(S.bidx[0], dtype=np.int64)
            return MLMatrix(structure=S, data=np.asarray(A[bidx[:, 0], bidx[:, 1]]).ravel())
        if self._bd is not None:
            if self._vec:
                raise NotImplementedError('vector-valued boundary forms in MLB format')
            return self._bd_matrix(self.blocks[(0, 0)])
        if not self._vec:
            return super().assemble_mlb(**kw)
        dev = self.dev
        S_base = dev.structure
        nc_u, nc_v = self._nc
        from .mlmatrix import MLStructure
        S = S_base.join(MLStructure.dense((nc_v, nc_u)))
        data = np.zeros(tuple(len(b) for b in S_base.bidx) + (nc_v * nc_u,))
        for (ct, cu), buf in self._block_mlb().items():
            data[..., ct * nc_u + cu] = dev.be.to_host(buf).reshape(data.shape[:-1])
        X = MLMatrix(structure=S, data=data)
        if layout == 'blocked':
            L = S.L
            X = X.reorder((L - 1,) + tuple(range(L - 1)))
        return X

    def assemble_csr(self, layout='blocked', format='csr', **kw):
        if self._lift1d:
            return self._marginal_matrix()
        if self._bd is not None and not self._vec:
            return self._bd_matrix(self.blocks[(0, 0)]).asmatrix('csr')
        if not self._vec:
            return super().assemble_csr(**kw)
        import scipy.sparse
        nc_u, nc_v = self._nc
        ds = self.dev.device_structure
        mats = {}
        if self._bd is not None:
            mats = {blk: self._bd_matrix(b).asmatrix('csr') for blk, b in self.blocks.items()}
        else:
            for blk, buf in self._block_mlb().items():
                mats[blk] = ds.to_csr(buf)
        ref = next(iter(mats.values()))
        if layout == 'packed':
            data = np.zeros((ref.nnz, nc_v, nc_u))
            for (ct, cu), A in mats.items():
                data[:, ct, cu] = A.data
            n = ref.shape[0]
            A = scipy.sparse.bsr_matrix((data, ref.indices, ref.indptr), shape=(n * nc_v, ref.shape[1] * nc_u),
                                        blocksize=(nc_v, nc_u))
            return A.asformat(format)
        zero = scipy.sparse.csr_matrix(ref.shape)
        rows = [[mats.get((ct, cu), zero) for cu in range(nc_u)] for ct in range(nc_v)]
        return scipy.sparse.bmat(rows, format='csr').asformat(format)

    # ---- arity 1 ------------------------------------------------------------------------------
    def assemble_vector(self):
        """Load vector, shape = number of dofs per axis (+ a trailing component axis for
        vector-valued forms) (``pyiga/genericasm.pxi:762-778, 852-868``)."""
        if self.arity != 1:
            return None
        be = self.dev.be
        if self._lift1d:
            return be.to_host(self.dev.assemble_vector_device()).reshape(self.dev.ndofs_test).sum(axis=0)
        sel = (lambda a: a) if self._bd is None else (lambda a: self._bd_select(a, 1))
        if not self._vec:
            return sel(be.to_host(self.dev.assemble_vector_device()).reshape(self.dev.ndofs_test))
        out = None
        for (ct, _), b in self.blocks.items():
            part = sel(be.to_host(b.dev.assemble_vector_device()).reshape(self.dev.ndofs_test))
            if out is None:
                out = np.zeros(part.shape + (self._nc[1],))
            out[..., ct] = part
        return out

    def _bd_lift(self, idx):
        """raveled indices of the boundary space -> raveled indices of the stand-in full space"""
        bdax, bdside = self._bd
        nb = tuple(kv.numdofs for kv in self.kvs[0])
        multi = list(np.unravel_index(np.asarray(idx, dtype=np.int64), nb)) if nb else []
        multi.insert(bdax, np.full(np.shape(idx), 0 if bdside == 0 else 1, dtype=np.int64))
        return np.ravel_multi_index(tuple(multi), self.dev.ndofs_test)

    def multi_entries(self, indices):
        if self.arity == 1:
            return self.multi_entries1(indices)
        if self._lift1d:
            if not isinstance(indices, np.ndarray):
                indices = np.array(list(indices), dtype=np.int64)
            indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
            return np.asarray(self._marginal_matrix()[indices[:, 0], indices[:, 1]]).ravel()
        if self._bd is not None:
            if not isinstance(indices, np.ndarray):
                indices = np.array(list(indices), dtype=np.int64)
            indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
            indices = np.column_stack((self._bd_lift(indices[:, 0]), self._bd_lift(indices[:, 1])))
        return super().multi_entries(indices)

    def multi_entries1(self, indices):
        if self.arity != 1:
            return None
        idx = np.asarray(indices if isinstance(indices, np.ndarray) else list(indices), dtype=np.int64)
        return self.assemble_vector().ravel()[idx]

    def entry1(self, i):
        return float(self.multi_entries1([i])[0]) if self.arity == 1 else 0.0

    def entry(self, i, j):
        if self._lift1d and self.arity == 2:
            return float(self.multi_entries(np.array([[i, j]]))[0])
        return super().entry(i, j) if self.arity == 2 else 0.0

    def update(self, **kwargs):
        """Re-evaluate the given input functions (``pyiga/codegen/cython.py:703-724``)."""
        known = {name: (shape, physical) for name, shape, physical, _ in self._vf.inputs}
        for name, f in kwargs.items():
            if name == 'geo':
                self._geo, self._X = f, None
                self._args['geo'] = f
                # surface measure and normal are functions of the geometry as well
                if self._bd is not None:
                    self._boundary_fields()
                if self._surface:
                    self._surface_fields()
                for n2, (shape, physical) in known.items():     # physical inputs move with the geometry
                    if physical:
                        self._env[n2] = self._eval_input(self._args[n2], shape, physical)
                self._Xd = None
                if '@x' in self._env:
                    if self._device_inputs():
                        self._env['@x'] = self._physical_points(device=True).movedim(-1, 0)
                    else:
                        X = self._physical_points()
                        self._env['@x'] = np.stack([X[..., i] for i in range(self._vf.dim)])
                continue
            if name not in known:
                raise ValueError("unknown input '%s'" % name)
            self._args[name] = f
            self._env[name] = self._eval_input(f, *known[name])
        self._recompute()

    def update_params(self, **kwargs):
        for name, v in kwargs.items():
            self._env[name] = np.asarray(v, dtype=float)
        self._recompute()


def _mentions_x(expr):
    from .vform import _mentions
    return _mentions(expr, '@x')


class MassAssembler2D(_ScalarAssemblerBase):
    """``u * v * dx`` in 2D (``pyiga/assemblers.pyx:26-172``)."""
    _form, _dim = _lib.FORM_MASS, 2


class StiffnessAssembler2D(_ScalarAssemblerBase):
    """``inner(grad(u), grad(v)) * dx`` in 2D (``pyiga/assemblers.pyx:174-349``)."""
    _form, _dim = _lib.FORM_STIFFNESS, 2


class MassAssembler3D(_ScalarAssemblerBase):
    """``u * v * dx`` in 3D (``pyiga/assemblers.pyx:1158-1322``)."""
    _form, _dim = _lib.FORM_MASS, 3


class StiffnessAssembler3D(_ScalarAssemblerBase):
    """``inner(grad(u), grad(v)) * dx`` in 3D (``pyiga/assemblers.pyx:1324-1540``)."""
    _form, _dim = _lib.FORM_STIFFNESS, 3


class _L2FunctionalBase(GenericFormAssembler):
    """``f * v * dx`` — load vector of a function (``pyiga/assemblers.pyx:1959-2504``).  `f` lives on
    the parameter domain (L2FunctionalAssembler) or on the physical domain (...Phys)."""
    _physical = False
    _dim = None

    @classmethod
    def inputs(cls):
        return {'geo': (cls._dim,), 'f': ()}

    @classmethod
    def parameters(cls):
        return {}

    def __init__(self, kvs0, geo, f):
        from . import vform
        vf = vform.VForm(self._dim, arity=1)
        v = vf.basisfuns()
        fin = vf.input('f', shape=(), physical=self._physical)
        vf.add(fin * v * vf.dx)
        self._vf = vf
        if not self._physical and not hasattr(f, 'grid_eval'):
            f = _ParametricCallable(f)
        super().__init__(kvs0, geo=geo, f=f)


class _ParametricCallable:
    """plain callable evaluated on the parameter grid (``pyiga/utils.py:33-41`` grid_eval)"""
    def __init__(self, f, kvs=None):
        self.f = f
        if kvs is not None:         # lets the form parser treat it like a (scalar) spline function: parametric input
            self.kvs = tuple(kvs)

    def output_shape(self):
        return ()

    def __call__(self, *x):
        return self.f(*x)

    def grid_eval(self, grid):
        from . import utils
        return np.asarray(utils.grid_eval(self.f, grid), dtype=float)       # grid + the shape of f's values


class L2FunctionalAssembler2D(_L2FunctionalBase):
    _dim = 2


class L2FunctionalAssembler3D(_L2FunctionalBase):
    _dim = 3


class L2FunctionalAssemblerPhys2D(_L2FunctionalBase):
    _dim, _physical = 2, True


class L2FunctionalAssemblerPhys3D(_L2FunctionalBase):
    _dim, _physical = 3, True


class _DivDivBase(GenericFormAssembler):
    """``div(u) * div(v) * dx`` for vector-valued u, v (``pyiga/assemblers.pyx:692-881``)."""
    _dim = None

    @classmethod
    def inputs(cls):
        return {'geo': (cls._dim,)}

    @classmethod
    def parameters(cls):
        return {}

    def __init__(self, kvs0, geo):
        from . import vform
        self._vf = vform.divdiv_vf(self._dim)
        super().__init__(kvs0, geo=geo)


class DivDivAssembler2D(_DivDivBase):
    _dim = 2


class DivDivAssembler3D(_DivDivBase):
    _dim = 3


_SPACETIME = ('HeatAssembler_ST2D', 'HeatAssembler_ST3D', 'WaveAssembler_ST2D', 'WaveAssembler_ST3D')


def __getattr__(name):
    # the space-time classes of the reference's predefined set live in .spacetime (which imports this module)
    if name in _SPACETIME:
        from . import spacetime
        return getattr(spacetime, name)
    raise AttributeError('module %r has no attribute %r' % (__name__, name))
