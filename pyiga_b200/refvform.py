"""CUDA backend for the reference's ``VForm`` objects (``pyiga.vform.VForm``).

The reference compiles a variational form in two steps: ``VForm.finalize()`` rewrites the expression
tree into scalar expressions over *parametric* derivatives of the basis functions and named
variables (``pyiga/vform.py:705-731``), and ``AsmGenerator`` prints them as Cython source
(``pyiga/codegen/cython.py:748-805``).  This module consumes the same finalized object (SURVEY.md
Appendix A) and replaces the second step: every output expression is bilinear (arity 2) or linear
(arity 1) in the basis function slots ``PartialDerivExpr(u|v, D)``, so evaluating it with symbolic
slots yields the coefficient ``C[slot_v][slot_u](q)`` of every slot pair at every Gauss point —
exactly what the generated ``precompute_fields`` + ``combine`` evaluate entry by entry.  The
coefficient arrays become the field buffer of a ``PB200_FORM_CUSTOM`` device assembler and the matrix
comes from the sum-factorised pipeline.

Nothing from ``pyiga`` is imported: the nodes are recognised by class name and attributes, so any
object with the reference's structure is accepted.  Supported: volume integrals, integrals over a side of the patch (``boundary=``) and surface integrals
(``geo_dim = dim + 1``), one space or trial / test functions in two spaces on the same mesh,
derivatives up to second order (incl. mixed ones, as in the space-time wave form), scalar and vector-valued basis functions, parametric and physical
input fields, parameters, ``on_demand`` bounding boxes (``pyiga/codegen/cython.py:421-426,541-559``).
On the CUDA backend the coefficient arrays never exist on the host: the geometry (values, Jacobian) is
evaluated on the Gauss grid by the device kernels, input callables run on device arrays
(:mod:`pyiga_b200._devarray`), and the interpreter's arithmetic is elementwise device operations; the
result is bound as the field buffer without a copy.  What the device cannot evaluate (Hessians of input
fields, callables that need a real numpy array, geometries that are not splines) is evaluated on the host
exactly like the generated ``__init__`` does (``pyiga/codegen/cython.py:465-484``,
``pyiga/utils.py:33-52``) and uploaded.
"""
import numpy as np

from . import _device, _lib
from ._devarray import DevArray, unwrap
from .assemblers import DeviceAssembler, GenericFormAssembler, _AssemblerProtocol, _is_spline_geo, _lift_axis
from .mlmatrix import MLMatrix
from .quadrature import make_tensor_quadrature


def is_reference_vform(obj):
    return type(obj).__name__ == 'VForm' and hasattr(obj, 'finalize') and hasattr(obj, 'predefined_vars')


# ---------------------------------------------------------------------------------------------
# values that are (bi)linear in the basis function slots
# ---------------------------------------------------------------------------------------------
class _Lin:
    """{(v slot, u slot): coefficient}; a slot is None or (component, D) with D the parametric
    derivative orders in x, y, z order; a coefficient is a float or an array on the Gauss grid."""
    __slots__ = ('t',)

    def __init__(self, t):
        self.t = t

    @staticmethod
    def const(c):
        return _Lin({(None, None): c})

    def is_coef(self):
        return all(k == (None, None) for k in self.t)

    def coef(self):
        return self.t.get((None, None), 0.0)

    def __add__(self, o):
        out = dict(self.t)
        for k, c in o.t.items():
            out[k] = out[k] + c if k in out else c
        return _Lin(out)

    def __neg__(self):
        return _Lin({k: -c for k, c in self.t.items()})

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        out = {}
        for (v1, u1), c1 in self.t.items():
            for (v2, u2), c2 in o.t.items():
                if (v1 is not None and v2 is not None) or (u1 is not None and u2 is not None):
                    raise ValueError('expression is not linear in each basis function')
                k = (v1 if v1 is not None else v2, u1 if u1 is not None else u2)
                c = c1 * c2
                out[k] = out[k] + c if k in out else c
        return _Lin(out)

    def __truediv__(self, o):
        if not o.is_coef():
            raise ValueError('division by an expression that contains basis functions')
        return self * _Lin.const(1.0 / o.coef())


_FUNCS = {'abs': np.abs, 'sqrt': np.sqrt, 'exp': np.exp, 'log': np.log, 'sin': np.sin, 'cos': np.cos, 'tan': np.tan}

# test switch: run the interpreter on CPU torch tensors (the code path of the CUDA backend) under the emulation backend
_FORCE_TENSORS = False


def _tensor_device():
    """torch device the coefficient arrays live on, or None (numpy arrays on the host)"""
    be = _device.backend()
    if be.name == 'cuda':
        return be.device
    if _FORCE_TENSORS:
        import torch
        return torch.device('cpu')
    return None


def _is_tensor(a):
    return hasattr(a, 'device') and hasattr(a, 'expand')


class _Interpreter:
    """Evaluates the scalar expressions of a finalized VForm on a tensor Gauss grid."""

    def __init__(self, vf, gaussgrid, gaussweights, args, geo):
        self.vf, self.grid, self.gw, self.args, self.geo = vf, gaussgrid, gaussweights, args, geo
        self.shape = tuple(len(g) for g in gaussgrid)
        self.dim = len(gaussgrid)
        self.tdev = _tensor_device()
        self._vars = {}
        self._memo = {}

    def _up(self, arr):
        """host array -> where the coefficient arrays live"""
        if self.tdev is None or _is_tensor(arr):
            return arr
        import torch
        return torch.as_tensor(np.ascontiguousarray(arr, dtype=float), device=self.tdev)

    def _spline(self, f, want):
        """values / Jacobian of a spline function on the Gauss grid, evaluated where the arrays live"""
        if (self.tdev is not None and self.tdev.type == 'cuda' and _is_spline_geo(f) and 2 <= len(f.kvs) <= 3
                and len(f.kvs) == self.dim):
            return _device.eval_spline_on_grid(f, self.grid, want, keep_on_device=True)
        return self._up(np.asarray(f.grid_eval(self.grid) if want == 'value' else f.grid_jacobian(self.grid), dtype=float))

    # ---- variables ---------------------------------------------------------------------------
    def var_entry(self, var, I):
        key = (id(var), tuple(I))
        if key in self._vars:
            return self._vars[key]
        if var.expr is not None:
            e = var.expr
            if len(I) == 2 and getattr(var, 'symmetric', False) and I[0] > I[1]:
                I = (I[1], I[0])
            sub = e if len(I) == 0 else (e[I[0]] if len(I) == 1 else e[I[0], I[1]])
            val = self.eval(sub)
        else:
            val = _Lin.const(self._source_entry(var, I))
        self._vars[key] = val
        return val

    def _source_array(self, var):
        key = ('src', id(var))
        if key in self._vars:
            return self._vars[key]
        src = var.src
        kind = type(src).__name__
        if kind == 'Parameter':
            arr = np.asarray(self.args[src.name], dtype=float)
        elif kind == 'InputField':
            f = self.args[src.name]
            deriv = var.deriv or 0
            if deriv == 0:
                arr = self._grid_eval(f, src.physical)
            elif deriv == 1:
                assert not src.physical, 'Jacobian of physical input field not implemented'
                arr = self._spline(f, 'jacobian')
            elif deriv == 2:
                # symmetric part, linearised: (d_xx, d_xy, d_yy) / (d_xx, d_xy, d_xz, d_yy, d_yz, d_zz)  (pyiga/vform.py:354-360)
                assert not src.physical, 'Hessian of physical input field not implemented'
                arr = self._up(np.asarray(f.grid_hessian(self.grid), dtype=float))
            else:
                raise NotImplementedError('derivatives of order > 2 of input fields')
            arr = self._up(arr) if self.tdev is not None else np.asarray(arr, dtype=float)
        else:
            raise TypeError('invalid source %r of variable %s' % (src, var.name))
        self._vars[key] = arr
        return arr

    def _grid_eval(self, f, physical):
        """``grid_eval`` / ``grid_eval_transformed`` of pyiga/utils.py:33-52"""
        if hasattr(f, 'grid_eval') and not physical:
            return self._spline(f, 'value')
        if self.tdev is not None:
            try:
                return self._grid_eval_tensors(f, physical)
            except Exception:       # the callable needs a real numpy array: host evaluation below, then upload
                pass
        if physical:
            X = np.asarray(self.geo.grid_eval(self.grid), dtype=float)
            pts = tuple(X[..., i] for i in range(X.shape[-1]))
        else:
            pts = list(np.meshgrid(*self.grid, sparse=True, indexing='ij'))
            pts.reverse()
        vals = f(*pts)
        if isinstance(vals, tuple):     # vector-valued function given as a tuple of components
            vals = np.stack([np.broadcast_to(np.asarray(v, dtype=float), self.shape) for v in vals], axis=-1)
        vals = np.asarray(vals, dtype=float)
        target = self.shape + vals.shape[self.dim:]     # functions that ignore an argument return smaller arrays
        return vals if vals.shape == target else np.broadcast_to(vals, target)

    def _grid_eval_tensors(self, f, physical):
        """the same with device arrays behind the numpy protocols (see _devarray.py)"""
        import torch
        if physical:
            if 'X' not in self._vars:
                self._vars['X'] = self._spline(self.geo, 'value')
            X = self._vars['X']
            X = X.reshape(self.shape + (-1,))
            pts = [X[..., i] for i in range(X.shape[-1])]
        else:
            pts = []
            for k, g in enumerate(self.grid):
                shp = [1] * self.dim
                shp[k] = -1
                pts.append(self._up(np.asarray(g, dtype=float)).reshape(shp))
            pts.reverse()
        vals = unwrap(f(*[DevArray(c) for c in pts]))

        def as_tensor(v):
            if not _is_tensor(v):
                v = torch.as_tensor(np.asarray(v, dtype=float), device=self.tdev)
            return v.to(torch.float64)
        if isinstance(vals, (tuple, list)):     # vector-valued function given as a tuple of components
            vals = torch.stack([as_tensor(v).expand(self.shape) for v in vals], dim=-1)
        vals = as_tensor(vals)
        target = self.shape + tuple(vals.shape[self.dim:])
        return vals if tuple(vals.shape) == target else vals.expand(target)

    def _source_entry(self, var, I):
        arr = self._source_array(var)
        if type(var.src).__name__ == 'Parameter':
            return float(arr[tuple(I)]) if len(I) else float(arr)
        return arr[(Ellipsis,) + tuple(I)] if len(I) else arr

    # ---- expressions -------------------------------------------------------------------------
    def eval(self, e):
        k = id(e)
        if k in self._memo:
            return self._memo[k]
        r = self._eval(e)
        self._memo[k] = r
        return r

    def _eval(self, e):
        kind = type(e).__name__
        if kind == 'ConstExpr':
            return _Lin.const(float(e.value))
        if kind == 'VarRefExpr':
            assert sum(e.D) == 0, 'derivatives of variables must be resolved by finalize()'
            return self.var_entry(e.var, e.I)
        if kind == 'NegExpr':
            return -self.eval(e.x)
        if kind == 'BuiltinFuncExpr':
            v = self.eval(e.x)
            if not v.is_coef():
                raise ValueError('functions can only be applied to coefficient expressions')
            c = v.coef()
            if _is_tensor(c):
                import torch
                return _Lin.const(getattr(torch, e.funcname)(c))
            return _Lin.const(_FUNCS[e.funcname](c))
        if kind == 'ScalarOperExpr':
            vals = [self.eval(c) for c in e.children]
            r = vals[0]
            for v in vals[1:]:
                r = {'+': r.__add__, '-': r.__sub__, '*': r.__mul__, '/': r.__truediv__}[e.oper](v)
            return r
        if kind == 'PartialDerivExpr':
            assert not e.physical, 'physical derivatives must be resolved by finalize()'
            bf = e.basisfun
            slot = (bf.component or 0, tuple(int(d) for d in e.D))
            if max(slot[1]) > 2:
                raise NotImplementedError('derivatives of order > 2 of the basis functions')
            # linear forms name their only (test) function 'u' (pyiga/vform.py: names[:arity])
            is_test = bf.name == 'v' or self.vf.arity == 1
            return _Lin({(slot, None): 1.0}) if is_test else _Lin({(None, slot): 1.0})
        if kind == 'GaussWeightExpr':
            shp = [1] * self.dim
            shp[e.axis] = -1
            return _Lin.const(self._up(np.asarray(self.gw[e.axis], dtype=float)).reshape(shp))
        raise NotImplementedError('expression node %s after finalize()' % kind)


def _slot_to_axis(slot, dim):
    """parametric slot (component, D in x,y,z order) -> derivative slot of the C ABI: 0 (value), 1 + tensor axis
    of a first derivative, or 16 + sum_k order_k * 3**k over the tensor axes k for second and mixed derivatives
    (``PB_SLOT_EXT``, include/pyiga_b200.h)"""
    if slot is None:
        return -1
    D = slot[1]
    orders = [0] * dim
    for i, d in enumerate(D):
        orders[dim - 1 - i] = int(d)        # x is the last tensor axis (pyiga/codegen/cython.py:170)
    if sum(orders) == 0:
        return 0
    if sum(orders) == 1:
        return 1 + orders.index(1)
    return 16 + sum(o * 3 ** k for k, o in enumerate(orders))


class _ParametricBlock:
    """One scalar form given by coefficient arrays of PARAMETRIC slot pairs: tables + field upload."""

    def __init__(self, kvs, nqp, dim, arity, coefs, grid_shape, full_shape=None, box=None, quad=None, kvs_test=None):
        be = _device.backend()
        keys = sorted(coefs)
        terms = [(f, bp, ap) for f, (bp, ap) in enumerate(keys)]
        self.dev = DeviceAssembler(kvs, kvs_test or kvs, _lib.FORM_CUSTOM, nqp=nqp, terms=terms, nfields=len(terms), quad=quad)
        full = full_shape or grid_shape
        if any(_is_tensor(c) for c in coefs.values()):
            # coefficient arrays computed on the device: they become the field buffer without touching the host
            import torch
            tdev = next(c.device for c in coefs.values() if _is_tensor(c))
            if box is None:
                fields = torch.empty((len(terms),) + tuple(full), dtype=torch.float64, device=tdev)
            else:
                fields = torch.zeros((len(terms),) + tuple(full), dtype=torch.float64, device=tdev)
            for f, key in enumerate(keys):
                c = coefs[key]
                vals = (c if _is_tensor(c) else torch.as_tensor(np.asarray(c, dtype=float), device=tdev)).expand(grid_shape)
                if box is None:
                    fields[f] = vals
                else:
                    fields[(f,) + tuple(slice(a, b) for a, b in box)] = vals
            buf = fields.reshape(-1) if be.name == 'cuda' else be.from_host(fields.numpy().ravel())
        else:
            fields = np.zeros((len(terms),) + tuple(full))
            for f, key in enumerate(keys):
                vals = np.broadcast_to(np.asarray(coefs[key], dtype=float), grid_shape)
                if box is None:
                    fields[f] = vals
                else:       # on-demand assembler: only the Gauss points of the bounding box carry data
                    fields[(f,) + tuple(slice(a, b) for a, b in box)] = vals
            buf = be.from_host(np.ascontiguousarray(fields).ravel())
        self.dev.fields = buf
        self.dev._geo_bound = None
        _device.check(be.lib.pb200_asm_bind_fields(self.dev.handle, be.ptr(buf)))


class RefVFormAssembler(GenericFormAssembler):
    """Device assembler class for a finalized reference VForm (what ``compile_vform`` returns)."""
    _rvf = None
    _on_demand = False

    @classmethod
    def inputs(cls):
        return {inp.name: inp.shape for inp in cls._rvf.inputs}

    @classmethod
    def parameters(cls):
        return {par.name: par.shape for par in cls._rvf.params}

    def __init__(self, kvs, kvs_test=None, bbox=None, boundary=None, **args):
        vf = self._rvf
        kvs = tuple(kvs)
        self._ctor_args = (kvs, kvs_test, bbox, boundary)
        d = vf.dim
        assert len(kvs) == d, "Assembler requires %d knot vectors" % d
        if vf.num_spaces() == 2:
            # trial functions in `kvs` (space 0, columns), test functions in `kvs_test` (space 1, rows), on the
            # same mesh (pyiga/assemble.py:947-951, generated __init__(kvs0, kvs1))
            assert kvs_test is not None and len(kvs_test) == d, "Assembler requires %d knot vectors" % d
            kvs_test = tuple(kvs_test)
            assert all(np.array_equal(a.mesh, b.mesh) for a, b in zip(kvs, kvs_test)), 'both spaces must share the mesh'
            spaces = {bf.name: bf.space for bf in vf.basis_funs}
            if vf.arity != 2 or spaces.get('u') != 0 or spaces.get('v') != 1:
                raise NotImplementedError('two-space forms need the trial function in space 0 and the test function in space 1')
        else:
            kvs_test = None
        geo = args['geo']
        assert geo.sdim == d, "Geometry has wrong source dimension"
        assert geo.dim == vf.geo_dim, "Geometry has wrong dimension"
        self.arity = vf.arity
        self.nqp = max(kv.p for kv in kvs + (kvs_test or ())) + 1
        self.kvs = (kvs, kvs_test or kvs)
        self._geo, self._args = geo, dict(args)
        self.bbox = bbox
        self._bd, self._surface = None, False
        self._lift1d = (d == 1)     # one knot vector: the lifted space of GenericFormAssembler (assemblers.py)
        if self._lift1d and (vf.is_boundary or (self._on_demand and bbox is not None) or vf.vec):
            raise NotImplementedError('boundary, on_demand and vector-valued forms over one knot vector')
        meshes = [np.asarray(kv.mesh) for kv in kvs]
        full_grid, full_w = make_tensor_quadrature(meshes, self.nqp)
        box = None
        quad = None
        kvs_dev = kvs
        if vf.is_boundary and kvs_test is not None:
            raise NotImplementedError('boundary integrals over two different spaces')
        if vf.is_boundary:
            # one side of the patch (pyiga/codegen/cython.py:549-590): the rule of the normal axis is the boundary point
            # with weight 1; the device tables get the linear stand-in of GenericFormAssembler._setup_boundary
            if self._on_demand and bbox is not None:
                raise NotImplementedError('on_demand boundary assemblers')
            self.gaussgrid = full_grid
            kvs_dev, quad = self._setup_boundary(kvs, boundary)
            bdax = self._bd[0]
            full_grid = self.gaussgrid
            full_w = tuple(np.ones(1) if k == bdax else w for k, w in enumerate(full_w))
        if self._on_demand and bbox is not None:
            # NB (pyiga/codegen/cython.py:541-559): bb[1] is the exclusive upper cell index
            grid, w = make_tensor_quadrature([m[bb[0]:bb[1] + 1] for m, bb in zip(meshes, bbox)], self.nqp)
            box = [(bb[0] * self.nqp, bb[1] * self.nqp) for bb in bbox]
        else:
            grid, w = full_grid, full_w
        self.gaussgrid = full_grid
        self._grid_shape = tuple(len(g) for g in grid)
        full_shape = tuple(len(g) for g in full_grid)
        itp = _Interpreter(vf, grid, w, args, geo)
        if self._lift1d:
            lift = _lift_axis()
            kvs_dev, kvs_test = (lift,) + kvs, ((lift,) + kvs_test if kvs_test is not None else None)
            _, (w_eta, _w) = make_tensor_quadrature([np.asarray(lift.mesh), meshes[0]], self.nqp)
            weta = itp._up(np.asarray(w_eta, dtype=float)).reshape(-1, 1)
            self._grid_shape = full_shape = (len(w_eta),) + self._grid_shape
            d_dev = 2
        else:
            d_dev = d
        nc_u = (vf.basis_funs[0].numcomp or 1) if vf.arity == 2 else 1
        nc_v = (vf.basis_funs[-1].numcomp or 1)
        self._vec = bool(vf.vec)
        self._nc = (nc_u, nc_v)
        if self._vec:
            self.num_components = lambda: self._nc
        # output expressions: scalars, or literal vectors with one entry per (test comp, trial comp)
        blocks = {}
        for e in vf.exprs:
            scal = [e] if e.shape == () else [e[k] for k in range(e.shape[0])]
            for k, se in enumerate(scal):
                lin = itp.eval(se)
                for (vs, us), c in lin.t.items():
                    # vector-valued forms: finalize() has expressed every component pair through the SCALAR
                    # basis functions; entry k of the output vector is the block entry (test component
                    # k // nc_u, trial component k % nc_u)  (pyiga/vform.py:441-458)
                    if self.arity == 2:
                        if vs is None or us is None:
                            raise ValueError('bilinear form with a term that lacks a basis function')
                        blk = (k // nc_u, k % nc_u) if self._vec else (0, 0)
                    else:
                        if vs is None or us is not None:
                            raise ValueError('linear form must contain v and no u')
                        blk = (k, None) if self._vec else (0, None)
                    if self._bd is not None:        # the stand-in of the normal axis is linear: value and first derivative only
                        for sl in (vs, us):
                            if sl is not None and sl[1][d - 1 - self._bd[0]] > 1:
                                raise NotImplementedError('second normal derivatives in boundary integrals')
                    if self._lift1d:        # (eta, xi): no derivative in eta, the weights of the eta rule
                        vs, us = ((sl[0], tuple(sl[1]) + (0,)) if sl is not None else None for sl in (vs, us))
                        key = (_slot_to_axis(vs, 2), _slot_to_axis(us, 2))
                        c = weta * c
                    else:
                        key = (_slot_to_axis(vs, d), _slot_to_axis(us, d))
                    dst = blocks.setdefault(blk, {})
                    dst[key] = dst[key] + c if key in dst else c
        if not blocks:
            raise ValueError('the form has no terms')
        self.blocks = {}
        for blk, coefs in blocks.items():
            # slot pairs whose coefficient vanishes identically (e.g. the geometry Hessian terms of a fourth-order
            # form on an affine map) would only cost launches
            live = {k: c for k, c in coefs.items() if (bool((c != 0.0).any()) if _is_tensor(c) else np.any(np.asarray(c) != 0.0))}
            coefs = live or dict([next(iter(coefs.items()))])
            self.blocks[blk] = _ParametricBlock(kvs_dev, self.nqp, d_dev, self.arity, coefs, self._grid_shape, full_shape, box, quad=quad,
                                                kvs_test=kvs_test)
        first = next(iter(self.blocks.values()))
        self.dev = self.blocks.get((0, 0) if self.arity == 2 else (0, None), first).dev

    def update(self, **kwargs):
        """Re-evaluate input functions (``pyiga/codegen/cython.py:703-724``): the finalized expressions are interpreted
        again with the new inputs (device arrays on the CUDA backend) and the field buffers rebuilt."""
        unknown = [k for k in kwargs if k not in self.inputs()]
        if unknown:
            raise ValueError("unknown input '%s'" % unknown[0])
        self._rebuild(kwargs)

    def update_params(self, **kwargs):
        unknown = [k for k in kwargs if k not in self.parameters()]
        if unknown:
            raise ValueError("unknown parameter '%s'" % unknown[0])
        self._rebuild(kwargs)

    def _rebuild(self, changed):
        kvs, kvs_test, bbox, boundary = self._ctor_args
        args = dict(self._args)
        args.update(changed)
        self.__init__(kvs, kvs_test, bbox=bbox, boundary=boundary, **args)


def compile_vform(vf, on_demand=False):
    """``pyiga.compile.compile_vform`` for the device: returns an assembler CLASS for the reference VForm
    `vf` (finalized here, once; ``pyiga/compile.py:120-132`` keys its cache the same way)."""
    key = (id(vf), bool(on_demand))
    cls = _cache.get(key)
    if cls is None:
        if vf.dim < 1 or vf.dim > 3:
            raise NotImplementedError('forms over 1 to 3 knot vectors (form of dimension %d)' % vf.dim)
        if not getattr(vf, '_VForm__is_finalized', False):
            vf.finalize(do_precompute=True)
        cls = type('RefVFormAssembler%d' % len(_cache), (RefVFormAssembler,), {'_rvf': vf, '_vf': vf, '_on_demand': bool(on_demand)})
        _cache[key] = cls
        _keep.append(vf)
    return cls


_cache = {}
_keep = []
