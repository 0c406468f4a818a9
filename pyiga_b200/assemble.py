"""Drivers of the assembly path: ``mass``, ``stiffness``, ``assemble``, ``Assembler``,
``assemble_entries`` — same call signatures and result layouts as the reference
(``pyiga/assemble.py:703-754, 837-1049``).

The reference driver materialises the index lists of the whole pattern
(``MLStructure.nonzero`` -> ``asm.multi_entries`` -> COO -> CSR -> mirror,
``pyiga/assemble.py:741-754``).  Here an assembler produces the multi-level banded value
tensor directly on the device and the CSR arrays come from one permutation kernel.
"""
import numpy as np

from . import assemblers, bspline, geometry
from .mlmatrix import MLMatrix, MLStructure


def _detect_dim(kvs):
    if hasattr(kvs, 'kv') and hasattr(kvs, 'p'):
        return 1, kvs
    d = len(kvs)
    return d, (kvs[0] if d == 1 else kvs)


def _default_geo(kvs):
    """geo=None means the identity map on the parameter box (the reference then takes the
    Kronecker shortcut, ``pyiga/assemble.py:236-282``; the matrices are the same)."""
    return geometry.identity(kvs)


def _assemble_1d(kv, form):
    """1D mass / stiffness matrix of a knot vector, computed on the device as a marginal of the 2D
    matrix over (kv) x (one linear element): M2 = M_kv (x) M_1 and K2 = K_kv (x) M_1 + M_kv (x) K_1,
    and the entries of M_1 sum to 1 while those of K_1 sum to 0.  Replaces ``bsp_mass_1d`` /
    ``bsp_stiffness_1d`` (``pyiga/assemble.py:152-234``)."""
    import scipy.sparse
    unit = bspline.make_knots(1, 0.0, 1.0, 1)
    kvs = (kv, unit)
    cls = assemblers.MassAssembler2D if form == 'mass' else assemblers.StiffnessAssembler2D
    M = cls(kvs, geometry.identity(kvs)).assemble_mlb()
    vals = M.data.sum(axis=1)
    b = M.structure.bidx[0]
    n = kv.numdofs
    return scipy.sparse.csr_matrix((vals, (b[:, 0].astype(np.int64), b[:, 1].astype(np.int64))), shape=(n, n))


def bsp_mixed_deriv_biform_1d_asym(knotvec1, knotvec2, du, dv, quadgrid=None, nqp=None, weightfunc=None):
    """Matrix of ``a(u, v) = (w u^(du), v^(dv))`` between two B-spline bases on one mesh: trial functions
    in `knotvec1` (columns), test functions in `knotvec2` (rows), derivative orders 0, 1 or 2
    (``pyiga/assemble.py:192-222``).  Computed on the device as the marginal of a lifted two-axis form over
    (one linear element) x (kv) whose only coefficient field is the product of the Gauss weights (and of the
    weight function at the nodes), see :func:`_assemble_1d`; the Gauss rule has max(p)+1 points per span,
    which is exact for these polynomial integrands like the reference's."""
    import scipy.sparse
    from . import refvform
    from .quadrature import make_tensor_quadrature
    same_mesh = np.array_equal(np.asarray(knotvec1.mesh), np.asarray(knotvec2.mesh))
    if not same_mesh or (quadgrid is not None and not np.array_equal(np.asarray(quadgrid), np.asarray(knotvec1.mesh))):
        raise NotImplementedError('both bases and the quadrature grid must share one mesh on the device path')
    nq = max(knotvec1.p, knotvec2.p) + 1
    if nqp is not None and nqp != nq:
        raise NotImplementedError('custom numbers of quadrature nodes are not part of the device path')
    if weightfunc is not None and nqp is None and du + dv > 0:
        # the reference takes ceil((p1 + p2 - du - dv + 1) / 2) nodes: exact for polynomial integrands, so the
        # rule does not matter without a weight — with one it does, and the device tables have p+1 nodes
        raise NotImplementedError('weighted forms with derivatives use the (p+1)-node rule here: pass nqp=%d '
                                  '(to the reference as well) to get the same quadrature' % nq)
    if du not in (0, 1, 2) or dv not in (0, 1, 2):
        raise NotImplementedError('derivative orders above 2 are not part of the device path')
    lift = assemblers._lift_axis()
    (g_eta, g_x), (w_eta, w_x) = make_tensor_quadrature([np.asarray(lift.mesh), np.asarray(knotvec1.mesh)], nq)
    coef = np.asarray(w_eta)[:, None] * np.asarray(w_x)[None, :]
    if weightfunc is not None:
        coef = coef * np.asarray(weightfunc(np.asarray(g_x)), dtype=float)[None, :]
    key = (refvform._slot_to_axis((0, (dv, 0)), 2), refvform._slot_to_axis((0, (du, 0)), 2))
    same = knotvec1 == knotvec2
    blk = refvform._ParametricBlock((lift, knotvec1), nq, 2, 2, {key: coef}, coef.shape,
                                    kvs_test=None if same else (lift, knotvec2))
    S = blk.dev.structure
    data = blk.dev.be.to_host(blk.dev.assemble_mlb()).reshape(tuple(len(b) for b in S.bidx))
    b = np.asarray(S.bidx[1], dtype=np.int64)
    A = scipy.sparse.csr_matrix((data.sum(axis=0), (b[:, 0], b[:, 1])), shape=(knotvec2.numdofs, knotvec1.numdofs))
    A.sort_indices()
    return A


def bsp_mixed_deriv_biform_1d(knotvec, du, dv, nqp=None, weightfunc=None):
    """``a(u, v) = (w u^(du), v^(dv))`` on one knot vector (``pyiga/assemble.py:179-190``)."""
    return bsp_mixed_deriv_biform_1d_asym(knotvec, knotvec, du, dv, nqp=nqp, weightfunc=weightfunc)


def bsp_mass_1d(knotvec, weightfunc=None):
    if weightfunc is not None:
        return bsp_mixed_deriv_biform_1d(knotvec, 0, 0, weightfunc=weightfunc)
    return _assemble_1d(knotvec, 'mass')


def bsp_stiffness_1d(knotvec, weightfunc=None):
    if weightfunc is not None:
        return bsp_mixed_deriv_biform_1d(knotvec, 1, 1, weightfunc=weightfunc)
    return _assemble_1d(knotvec, 'stiffness')


def bsp_mass_1d_asym(knotvec1, knotvec2, quadgrid=None):
    return bsp_mixed_deriv_biform_1d_asym(knotvec1, knotvec2, 0, 0, quadgrid=quadgrid)


def bsp_stiffness_1d_asym(knotvec1, knotvec2, quadgrid=None):
    return bsp_mixed_deriv_biform_1d_asym(knotvec1, knotvec2, 1, 1, quadgrid=quadgrid)


def assemble_entries(asm, symmetric=False, format='csr', layout='blocked'):
    """Assemble all entries of an assembler object (``pyiga/assemble.py:703-754``).

    Device assemblers take the sum-factorised path (`symmetric` only affects the reference's CPU
    strategy and is accepted for compatibility).  ``format='mlb'`` returns an
    :class:`~pyiga_b200.mlmatrix.MLMatrix`; every other value a scipy sparse matrix.
    """
    if asm.arity == 1:
        result = asm.assemble_vector()
        if hasattr(asm, 'num_components') and layout == 'blocked':
            result = np.moveaxis(result, -1, 0)
        return result
    if not hasattr(asm, 'assemble_mlb'):
        raise TypeError('assemble_entries needs a pyiga_b200 device assembler, got %r' % type(asm))
    if hasattr(asm, 'num_components'):
        return assemble_entries_vec(asm, symmetric=symmetric, format=format, layout=layout)
    if format == 'mlb':
        return asm.assemble_mlb()
    return asm.assemble_csr().asformat(format)


def assemble_partial_rows(asm, row_indices, restrict=False):
    """Submatrix which contains only the given rows (``_assemble_partial_rows``,
    ``pyiga/_hdiscr.py:5-12``): CSR matrix of the full shape whose other rows are empty or, with
    `restrict`, of shape ``(len(row_indices), ncols)`` with the rows in the given order.  The pattern
    of a row is a closed form of the per-axis band tables, so no index lists are built on the host;
    the values come from the per-entry quadrature kernel."""
    import scipy.sparse
    if not hasattr(asm, 'dev'):
        raise TypeError('assemble_partial_rows needs a pyiga_b200 device assembler, got %r' % type(asm))
    rows = np.asarray(row_indices, dtype=np.int64).ravel()
    shape = tuple(int(np.prod([kv.numdofs for kv in kvs], dtype=np.int64)) for kvs in (asm.kvs[1], asm.kvs[0]))
    if not restrict:
        rows = np.unique(rows)
    be = asm.dev.be
    indptr, indices, values = (be.to_host(x) for x in asm.dev.rows_csr_device(rows))
    if restrict:
        A = scipy.sparse.csr_matrix((values, indices, indptr), shape=(rows.size, shape[1]))
    else:
        counts = np.zeros(shape[0], dtype=indptr.dtype)
        counts[rows] = np.diff(indptr)
        full = np.concatenate(([0], np.cumsum(counts))).astype(indptr.dtype)
        A = scipy.sparse.csr_matrix((values, indices, full), shape=shape)
    A.has_sorted_indices = True
    return A


def assemble_entries_vec(asm, symmetric=False, format='csr', layout='blocked'):
    """Vector-valued forms (``pyiga/assemble.py:761-810``).  `layout='blocked'`: a k_test x k_trial
    block matrix of scalar matrices; `'packed'`: every scalar entry becomes a small k_test x k_trial
    block (`format='bsr'` gives the BSR matrix directly).  `format='mlb'` returns the MLMatrix with
    the dense component level."""
    assert layout in ('packed', 'blocked')
    if format == 'mlb':
        return asm.assemble_mlb(layout=layout)
    return asm.assemble_csr(layout=layout, format=format)


def divdiv(kvs, geo=None, layout='blocked', format='csr'):
    """``div(u) div(v)`` matrix for vector-valued functions (``pyiga/assemble.py:1051-1061``)."""
    dim, kvs = _detect_dim(kvs)
    if geo is None:
        geo = geometry.unit_cube(dim=dim)
    cls = {2: assemblers.DivDivAssembler2D, 3: assemblers.DivDivAssembler3D}.get(dim)
    assert cls is not None, 'dimension %d not implemented' % dim
    return assemble_entries_vec(cls(kvs, geo), symmetric=True, layout=layout, format=format)


def mass(kvs, geo=None, format='csr'):
    """Mass matrix of a tensor-product B-spline space (``pyiga/assemble.py:1017-1032``)."""
    dim, kvs = _detect_dim(kvs)
    if geo:
        assert geo.dim == dim, "Geometry has wrong dimension"
    if dim == 1:
        assert geo is None, "Geometry map not supported for 1D assembling"
        return bsp_mass_1d(kvs).asformat(format)
    if geo is None:
        geo = _default_geo(kvs)
    cls = {2: assemblers.MassAssembler2D, 3: assemblers.MassAssembler3D}.get(dim)
    assert cls is not None, "Dimensions higher than 3 are currently not implemented."
    return assemble_entries(cls(kvs, geo), symmetric=True, format=format)


def stiffness(kvs, geo=None, format='csr'):
    """Stiffness matrix of a tensor-product B-spline space (``pyiga/assemble.py:1034-1049``)."""
    dim, kvs = _detect_dim(kvs)
    if geo:
        assert geo.dim == dim, "Geometry has wrong dimension"
    if dim == 1:
        assert geo is None, "Geometry map not supported for 1D assembling"
        return bsp_stiffness_1d(kvs).asformat(format)
    if geo is None:
        geo = _default_geo(kvs)
    cls = {2: assemblers.StiffnessAssembler2D, 3: assemblers.StiffnessAssembler3D}.get(dim)
    assert cls is not None, "Dimensions higher than 3 are currently not implemented."
    return assemble_entries(cls(kvs, geo), symmetric=True, format=format)


def bsp_mass_2d(knotvecs, geo=None, format='csr'):
    return mass(knotvecs, geo, format)


def bsp_mass_3d(knotvecs, geo=None, format='csr'):
    return mass(knotvecs, geo, format)


def bsp_stiffness_2d(knotvecs, geo=None, format='csr'):
    return stiffness(knotvecs, geo, format)


def bsp_stiffness_3d(knotvecs, geo=None, format='csr'):
    return stiffness(knotvecs, geo, format)


def inner_products(kvs, f, f_physical=False, geo=None):
    """L2 inner products of every basis function with `f` (``pyiga/assemble.py:288-340``): array of
    shape ndofs.  `f` is a function of the parameter coordinates unless `f_physical`."""
    dim, kvs = _detect_dim(kvs)
    if dim == 1:
        # marginal of the 2D functional on (kv) x (one linear element), as in _assemble_1d: the two
        # functions of the dummy axis sum to 1, and f only sees the first coordinate
        if geo is not None:     # (``test/test_assemble.py:441-444``) the load vector of 'f * v * dx' on the mapped interval
            fin = f if (f_physical or hasattr(f, 'grid_eval')) else assemblers._ParametricCallable(f, (kvs,))
            return assemble('f * v * dx', (kvs,), geo=geo, f=fin)
        unit = bspline.make_knots(1, 0.0, 1.0, 1)
        kvs2 = (kvs, unit)
        v = assemblers.L2FunctionalAssembler2D(kvs2, geometry.identity(kvs2), lambda x, y: f(y)).assemble_vector()
        return np.asarray(v).sum(axis=1)
    if geo is None:
        geo = _default_geo(kvs)
    name = 'L2FunctionalAssembler%s%dD' % ('Phys' if f_physical else '', dim)
    if hasattr(f, 'grid_eval') and hasattr(f, 'output_shape') and tuple(f.output_shape()) != ():
        # vector- or tensor-valued spline function: one load vector per component (``pyiga/assemble.py:318-340``)
        extra = tuple(f.output_shape())
        parts = [np.asarray(getattr(assemblers, name)(kvs, geo, _ComponentOf(f, idx)).assemble_vector()) for idx in np.ndindex(*extra)]
        return np.stack(parts, axis=-1).reshape(parts[0].shape + extra)
    if not hasattr(f, 'grid_eval'):
        # vector- or tensor-valued f: the reference returns ndofs + f's shape (``pyiga/assemble.py:318-340``);
        # one load vector per component, stacked on the trailing axes
        from . import utils
        # (two points per axis: callables that np.squeeze their arguments still see one-dimensional axes)
        mid = [kv.support()[0] + np.array([0.25, 0.75]) * (kv.support()[1] - kv.support()[0]) for kv in kvs]
        probe = np.asarray(utils.grid_eval_transformed(f, mid, geo) if f_physical else utils.grid_eval(f, mid))
        extra = probe.shape[dim:]
        if extra != ():
            def component(idx):
                def fc(*x):
                    v = f(*x)
                    if isinstance(v, tuple):
                        for i in idx:
                            v = v[i]
                        return v
                    return v[(Ellipsis,) + idx]
                return fc
            parts = [np.asarray(getattr(assemblers, name)(kvs, geo, component(idx)).assemble_vector()) for idx in np.ndindex(*extra)]
            return np.stack(parts, axis=-1).reshape(parts[0].shape + extra)
    return getattr(assemblers, name)(kvs, geo, f).assemble_vector()


class _ComponentOf:
    """one scalar component of a vector- or tensor-valued spline function (parametric input of a linear form)"""

    def __init__(self, f, idx):
        self.f, self.idx, self.kvs = f, tuple(idx), f.kvs

    def output_shape(self):
        return ()

    def grid_eval(self, grid):
        return np.asarray(self.f.grid_eval(grid))[(Ellipsis,) + self.idx]


################################################################################
# Incorporating essential boundary conditions (pyiga/assemble.py:342-652)
################################################################################

def slice_indices(ax, idx, shape, ravel=False, flip=None):
    """Dof indices of the slice `idx` across axis `ax` of a tensor-product basis of size `shape`:
    an ``N x dim`` array of multi-indices or, with `ravel`, raveled indices
    (``pyiga/assemble.py:346-368``)."""
    shape = tuple(shape)
    if idx < 0:
        idx += shape[ax]
    axdofs = [np.arange(n) for n in shape]
    if flip is not None:
        flip = tuple(flip)
        flip = flip[:ax] + (False,) + flip[ax:]
        axdofs = [a[::-1] if flp else a for a, flp in zip(axdofs, flip)]
    axdofs[ax] = np.array([idx])
    grids = np.meshgrid(*axdofs, indexing='ij')
    multi = np.stack([g.ravel() for g in grids], axis=1)
    if ravel:
        return np.ravel_multi_index(multi.T, shape)
    return multi


def boundary_dofs(kvs, bdspec, ravel=False, flip=None):
    """Indices of the dofs on one side of the boundary (``pyiga/assemble.py:370-377``)."""
    bdax, bdside = bspline._parse_bdspec(bdspec, len(kvs))
    N = tuple(kv.numdofs for kv in kvs)
    return slice_indices(bdax, 0 if bdside == 0 else -1, N, ravel=ravel, flip=flip)


def boundary_cells(kvs, bdspec, ravel=False):
    """Indices of the cells on one side of the boundary (``pyiga/assemble.py:379-386``)."""
    bdax, bdside = bspline._parse_bdspec(bdspec, len(kvs))
    N = tuple(kv.numspans for kv in kvs)
    return slice_indices(bdax, 0 if bdside == 0 else -1, N, ravel=ravel)


def _drop_nans(indices, values):
    keep = ~np.isnan(values)
    return (indices, values) if keep.all() else (indices[keep], values[keep])


def combine_bcs(bcs):
    """Merge ``(indices, values)`` pairs; a dof that occurs more than once takes its value from
    its first occurrence (``pyiga/assemble.py:553-566``)."""
    bcs = list(bcs)
    indices = np.concatenate([ind for ind, _ in bcs])
    values = np.concatenate([val for _, val in bcs])
    assert indices.shape == values.shape, 'Inconsistent BC sizes'
    uidx, lookup = np.unique(indices, return_index=True)
    return uidx, values[lookup]


def compute_dirichlet_bc(kvs, geo, bdspec, dir_func):
    """Indices and values of a Dirichlet condition on one side, by interpolation of `dir_func`
    (given in physical coordinates; scalars mean constants) on the boundary face
    (``pyiga/assemble.py:395-461``)."""
    from .approx import interpolate
    bdspec = bspline._parse_bdspec(bdspec, len(kvs))
    bdax, bdside = bdspec
    bdbasis = list(kvs)
    assert len(bdbasis) == geo.sdim, 'Invalid dimension of geometry'
    del bdbasis[bdax]
    bdgeo = geo.boundary(bdspec)
    if np.isscalar(dir_func):
        const_value = dir_func
        dir_func = lambda *x: const_value
    dircoeffs = interpolate(bdbasis, dir_func, geo=bdgeo)
    N = tuple(kv.numdofs for kv in kvs)
    bdindices = slice_indices(bdax, 0 if bdside == 0 else -1, N, ravel=True)
    extra_dims = dircoeffs.ndim - len(bdbasis)
    if extra_dims == 0:
        return _drop_nans(bdindices, dircoeffs.ravel())
    if extra_dims == 1:       # vector function, blocked vector discretization
        NN = int(np.prod(N))
        idx, val = combine_bcs((bdindices + j * NN, dircoeffs[..., j].ravel()) for j in range(dircoeffs.shape[-1]))
        return _drop_nans(idx, val)
    raise ValueError('invalid dimension of Dirichlet coefficients: %s' % (dircoeffs.shape,))


def compute_dirichlet_bcs(kvs, geo, bdconds):
    """Dirichlet conditions on several sides; ``('all', g)`` means every side
    (``pyiga/assemble.py:463-489``)."""
    if len(bdconds) == 2 and isinstance(bdconds[0], str) and bdconds[0] == 'all':
        dir_func = bdconds[1]
        bdconds = [((ax, bd), dir_func) for ax in range(len(kvs)) for bd in (0, 1)]
    return combine_bcs(compute_dirichlet_bc(kvs, geo, bdspec, g) for (bdspec, g) in bdconds)


def compute_initial_condition_01(kvs, geo, bdspec, g0, g1, physical=True):
    """Indices and values which prescribe function value (`g0`) and normal derivative (`g1`) on one
    face of a space-time cylinder with time-independent geometry, by interpolation on the face and
    a 2 x 2 collocation solve for the two layers of boundary coefficients
    (``pyiga/assemble.py:492-552``)."""
    from .approx import interpolate
    bdspec = bspline._parse_bdspec(bdspec, len(kvs))
    bdax, bdside = bdspec
    bdbasis = list(kvs)
    del bdbasis[bdax]
    bdgeo = geo.boundary(bdspec) if physical else None
    coeffs01 = np.stack((interpolate(bdbasis, g0, geo=bdgeo).ravel(), interpolate(bdbasis, g1, geo=bdgeo).ravel()))
    a, b = kvs[bdax].support()
    if bdside == 0:
        bdcolloc = bspline.active_deriv(kvs[bdax], float(a), 1)[:2, :2]      # first two basis functions
    else:
        bdcolloc = bspline.active_deriv(kvs[bdax], float(b), 1)[:2, -2:]     # last two basis functions
    coll_coeffs = np.linalg.solve(bdcolloc, coeffs01)
    N = tuple(kv.numdofs for kv in kvs)
    firstidx = 0 if bdside == 0 else -2
    bdindices = np.concatenate((slice_indices(bdax, firstidx, N, ravel=True),
                                slice_indices(bdax, firstidx + 1, N, ravel=True)))
    return bdindices, coll_coeffs.ravel()


class RestrictedLinearSystem:
    """Linear system with some dofs eliminated (``pyiga/assemble.py:568-652``).

    `A` may be a scipy sparse matrix, an :class:`MLMatrix` whose values are still on the GPU, or a
    :class:`~pyiga_b200._csr.DeviceCSR`.  The reference multiplies with 0/1 selection matrices; here
    the kept rows/columns are compacted by two passes over the device CSR arrays and the right-hand
    side is updated by one device matvec.  ``A`` (scipy CSR, fetched on first use), ``b``,
    ``restrict``, ``restrict_rhs``, ``restrict_matrix``, ``extend`` and ``complete`` behave like the
    reference; ``A_device`` is the restricted matrix as device CSR.
    """

    def __init__(self, A, b, bcs, elim_rows=None):
        from ._csr import DeviceCSR
        indices, values = bcs
        indices = np.asarray(list(indices), dtype=np.int64)
        Ad = DeviceCSR.wrap(A)
        nrows, ncols = Ad.shape
        if np.isscalar(b):
            b = np.broadcast_to(b, nrows)
        if np.isscalar(values):
            values = np.broadcast_to(values, indices.shape[0])
        self.values = np.asarray(values, dtype=np.float64)

        mask = np.ones(ncols, dtype=bool)
        mask[indices] = False
        self._free = np.nonzero(mask)[0]
        self._elim = np.nonzero(~mask)[0]
        # R_elim lists the eliminated dofs in increasing order; like the reference, `values` is
        # taken in that order
        self._colmap = np.full(ncols, -1, dtype=np.int32)
        self._colmap[self._free] = np.arange(self._free.size, dtype=np.int32)
        if elim_rows is not None:
            maskv = np.ones(nrows, dtype=bool)
            maskv[sorted(elim_rows)] = False
            self._free_v = np.nonzero(maskv)[0]
        else:
            self._free_v = self._free

        be = Ad.be
        self.A_device = Ad.restrict(self._free_v, self._colmap, self._free.size)
        self._A = None
        # b - A (R_elim^T values), restricted to the kept rows
        x_elim = np.zeros(ncols)
        x_elim[self._elim] = self.values
        t = Ad.matvec_device(be.from_host(x_elim), be.from_host(np.ascontiguousarray(b, dtype=np.float64)), alpha=-1.0)
        from . import _csr
        self.b = be.to_host(_csr.gather(t, self._free_v))

    @property
    def A(self):
        if self._A is None:
            self._A = self.A_device.to_scipy()
        return self._A

    def restrict(self, u):
        """Restriction of a vector of all dofs to the free dofs."""
        return np.asarray(u)[self._free]

    def restrict_rhs(self, f):
        """Restriction of a right-hand side to the non-eliminated rows."""
        return np.asarray(f)[self._free_v]

    def restrict_matrix(self, B):
        """Restriction of a matrix which operates on all dofs to the free dofs (scipy CSR)."""
        from ._csr import DeviceCSR
        return DeviceCSR.wrap(B).restrict(self._free_v, self._colmap, self._free.size).to_scipy()

    def extend(self, u):
        """Pad a vector of the free dofs with zeros to all dofs."""
        u = np.asarray(u)
        out = np.zeros((self._colmap.size,) + u.shape[1:], dtype=np.result_type(u, np.float64))
        out[self._free] = u
        return out

    def complete(self, u):
        """Solution of the original system from a solution `u` of the restricted one."""
        out = self.extend(u)
        out[self._elim] += self.values.reshape((-1,) + (1,) * (out.ndim - 1))
        return out


################################################################################
# Integration (pyiga/assemble.py:658-696)
################################################################################

def integrate(kvs, f, f_physical=False, geo=None):
    """Integral of `f` over the geometry `geo` or the parameter domain, with the Gauss rule of the
    spline space (``pyiga/assemble.py:658-696``).  The B-splines sum to one, so the integral is the
    sum of the L2 inner products of `f` with all basis functions — the load vector of the device
    path; vector-valued `f` is integrated component by component."""
    from . import utils
    dim, kvs_n = _detect_dim(kvs)
    kvs_t = (kvs_n,) if dim == 1 else tuple(kvs_n)
    if f_physical:
        assert geo is not None, 'integrate in physical domain requires geometry'
    if hasattr(f, 'grid_eval'):
        raise NotImplementedError('integrate() of spline function objects: pass a callable')
    # number of components: evaluate once at the centre of the parameter domain
    mid = [np.array([0.5 * (kv.support()[0] + kv.support()[1])]) for kv in kvs_t]
    probe = np.asarray(utils.grid_eval_transformed(f, mid, geo) if f_physical else utils.grid_eval(f, mid))
    extra = probe.shape[dim:]
    if extra == ():
        return float(np.sum(inner_products(kvs, f, f_physical=f_physical, geo=geo)))

    def component(idx):
        def fc(*x):
            v = f(*x)
            if isinstance(v, tuple):        # tuple of component arrays (stacked on the last axis by grid_eval)
                for i in idx:
                    v = v[i]
                return v
            return np.asarray(v)[(Ellipsis,) + idx]
        return fc
    out = np.empty(extra)
    for idx in np.ndindex(*extra):
        out[idx] = np.sum(inner_products(kvs, component(idx), f_physical=f_physical, geo=geo))
    return out


def _jac_to_boundary_matrix(bdspec, dim):
    """``dim x (dim-1)`` matrix that restricts the Jacobian of the volume map to the side `bdspec`; the signs make the
    resulting normal point outwards for a positively oriented patch (``pyiga/assemble.py:899-912``)."""
    ax, side = bdspec
    col = dim - 1 - ax                  # the last tensor axis is the x coordinate
    B = np.zeros((dim, dim - 1))
    for k, c in enumerate(c for c in range(dim) if c != col):
        B[c, k] = -1.0 if c % 2 == 0 else 1.0
    if side != 0:
        B[:, 0] = -B[:, 0]
    return B


def instantiate_assembler(problem, kvs, args, bfuns=None, boundary=None, updatable=[]):
    """Turn a problem description into an assembler object (``pyiga/assemble.py:914-956``)."""
    if isinstance(problem, str):
        from . import vform
        problem = vform.parse_vf(problem, kvs, args=args, bfuns=bfuns, boundary=bool(boundary), updatable=updatable)
    from . import refvform, vform as _vf
    num_spaces = 1
    if isinstance(problem, _vf._SpaceTimeForm):
        problem = _vf.compile_vform(problem)
    elif isinstance(problem, _vf.VForm) or refvform.is_reference_vform(problem):
        num_spaces = problem.num_spaces()
        problem = _vf.compile_vform(problem)
    if isinstance(problem, type):
        used = {}
        if boundary:
            used['boundary'] = bspline._parse_bdspec(boundary, len(kvs))
            if 'Jac_to_boundary' in problem.parameters():       # reference VForms take the restriction matrix as a parameter
                args = dict(args, Jac_to_boundary=_jac_to_boundary_matrix(used['boundary'], len(kvs)))
        for name in list(problem.inputs().keys()) + list(problem.parameters().keys()):
            if name not in args:
                raise ValueError("required input parameter '%s' missing" % name)
            used[name] = args[name]
        if num_spaces <= 1:
            return problem(kvs, **used)
        assert num_spaces == 2, 'no more than two spaces allowed'
        return problem(kvs[0], kvs[1], **used)
    raise TypeError("invalid type for 'problem': {}".format(type(problem)))


def assemble(problem, kvs=None, args=None, bfuns=None, boundary=None, symmetric=False, format='csr',
             layout='blocked', **kwargs):
    """Assemble a matrix or vector for a variational form given as string, :class:`VForm`,
    assembler class or assembler object (``pyiga/assemble.py:837-897``)."""
    args = dict() if args is None else args
    args.update(kwargs)
    if hasattr(problem, 'multi_entries') and not isinstance(problem, type):
        asm = problem
    else:
        asm = instantiate_assembler(problem, kvs, args, bfuns, boundary)
    return assemble_entries(asm, symmetric=symmetric, format=format, layout=layout)


def assemble_vf(vf, kvs, symmetric=False, format='csr', layout='blocked', args=None, **kwargs):
    """Assemble a :class:`~pyiga_b200.vform.VForm` (``pyiga/assemble.py:812-822``)."""
    args = dict() if args is None else args
    args.update(kwargs)
    return assemble(vf, kvs, symmetric=symmetric, format=format, layout=layout, args=args)


class Assembler:
    """Re-usable assembler with updatable inputs (``pyiga/assemble.py:958-1003``)."""

    def __init__(self, problem, kvs, args=None, bfuns=None, boundary=None, symmetric=False, updatable=[], **kwargs):
        args = dict() if args is None else args
        args.update(kwargs)
        self.symmetric = bool(symmetric)
        self.updatable = tuple(updatable)
        self.asm = instantiate_assembler(problem, kvs, args, bfuns, boundary, self.updatable)
        if not all(name in self.asm.inputs().keys() for name in self.updatable):
            raise ValueError('Assembler received an updatable argument which is not an assembler input')

    def update(self, **kwargs):
        if not hasattr(self.asm, 'update'):
            raise RuntimeError('assembler object is not updatable')
        if not all(name in self.updatable for name in kwargs.keys()):
            raise RuntimeError('update() received an argument which was not specified as updatable')
        self.asm.update(**kwargs)

    def assemble(self, format='csr', layout='blocked', **upd_fields):
        if upd_fields:
            self.update(**upd_fields)
        return assemble_entries(self.asm, symmetric=self.symmetric, format=format, layout=layout)
