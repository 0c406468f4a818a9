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


def bsp_mass_1d(knotvec):
    return _assemble_1d(knotvec, 'mass')


def bsp_stiffness_1d(knotvec):
    return _assemble_1d(knotvec, 'stiffness')


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
        raise NotImplementedError('1D inner products are not part of the device path')
    if geo is None:
        geo = _default_geo(kvs)
    name = 'L2FunctionalAssembler%s%dD' % ('Phys' if f_physical else '', dim)
    return getattr(assemblers, name)(kvs, geo, f).assemble_vector()


def instantiate_assembler(problem, kvs, args, bfuns=None, boundary=None, updatable=[]):
    """Turn a problem description into an assembler object (``pyiga/assemble.py:914-956``)."""
    if boundary:
        raise NotImplementedError('boundary integrals are not part of the device path')
    if isinstance(problem, str):
        from . import vform
        problem = vform.parse_vf(problem, kvs, args=args, bfuns=bfuns, updatable=updatable)
    from . import vform as _vf
    if isinstance(problem, _vf.VForm):
        problem = _vf.compile_vform(problem)
    if isinstance(problem, type):
        used = {}
        for name in list(problem.inputs().keys()) + list(problem.parameters().keys()):
            if name not in args:
                raise ValueError("required input parameter '%s' missing" % name)
            used[name] = args[name]
        return problem(kvs, **used)
    raise TypeError("invalid type for 'problem': {}".format(type(problem)))


def assemble(problem, kvs, args=None, bfuns=None, boundary=None, symmetric=False, format='csr',
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
