"""Variational forms for the device assembler: scalar bilinear forms with at most first derivatives.

This is the front end that ``assemble.assemble("... * dx", kvs, geo=..., f=...)`` uses.  It plays
the role of the reference's ``pyiga.vform`` + ``pyiga.codegen`` + ``pyiga.compile``
(``pyiga/vform.py:162-735, 1804-1887``, ``pyiga/codegen/cython.py:748-805``,
``pyiga/compile.py:120-132``) for the family of forms the device path implements

    a(u, v) = int  sum_t c_t(x) * d^{bt} v * d^{bu} u  dx,      bt, bu in {value, d/dx_1..d/dx_d}

(diffusion with scalar/matrix coefficients, convection, reaction, and their transposes).  Instead of
generating and compiling source code per form, an expression is *analysed*: evaluating it with
symbolic basis functions yields the coefficient ``c_t`` of every slot pair.  User callables are
evaluated on the host at the (physical) Gauss points — as the reference does
(``pyiga/codegen/cython.py:465-484``, ``pyiga/utils.py:33-52``) — uploaded, and the pull-back to the
parameter domain (``J^-1``, ``|det J|``, Gauss weights) happens in the K2 kernel
(``csrc/geo_fields.cuh: PbProgGeneral``).  The matrix itself comes from the same sum-factorised
pipeline as mass and stiffness (``pb200_asm_assemble_mlb`` with a generic stage plan).

Linear forms (arity 1, e.g. ``'f * v * dx'``) give load vectors through the same machinery, and
vector-valued basis functions (``bfuns=[('u', 2), ('v', 2)]``) are handled block by block: every
pair of components is one scalar form of the family above.  Second derivatives and general space-time
expressions are not part of THIS front end and raise ``NotImplementedError``; the predefined space-time
forms ``heat_st_vf`` / ``wave_st_vf`` map to :mod:`pyiga_b200.spacetime`, and the reference's own ``VForm``
objects (any derivative order up to 2, space-time) are consumed by :mod:`pyiga_b200.refvform`.
"""
import re

import numpy as np


# ---------------------------------------------------------------------------------------------
# coefficients and slot-forms (the values expressions evaluate to)
# ---------------------------------------------------------------------------------------------
class Coef:
    """scale * arr, where arr is an array on the Gauss grid or None (= 1)."""
    __slots__ = ('scale', 'arr')

    def __init__(self, scale=1.0, arr=None):
        self.scale, self.arr = float(scale), arr

    def value(self):
        return self.scale if self.arr is None else self.scale * self.arr

    def __mul__(self, o):
        if self.arr is None:
            return Coef(self.scale * o.scale, o.arr)
        if o.arr is None:
            return Coef(self.scale * o.scale, self.arr)
        return Coef(self.scale * o.scale, self.arr * o.arr)

    def __add__(self, o):
        if self.arr is o.arr:
            return Coef(self.scale + o.scale, self.arr)
        return Coef(1.0, self.value() + o.value())

    def is_zero(self):
        return self.scale == 0.0


class Form:
    """Scalar value of an expression: {(test slot, trial slot): Coef}.  A slot is None (no basis
    function) or (component, deriv) with deriv 0 = function value, 1+a = derivative with respect to
    physical coordinate a (x, y, z order); scalar basis functions have component 0."""

    def __init__(self, terms=None):
        self.terms = dict(terms or {})

    @staticmethod
    def const(c, arr=None):
        return Form({(None, None): Coef(c, arr)})

    def is_coef(self):
        return all(k == (None, None) for k in self.terms)

    def coef(self):
        assert self.is_coef()
        return self.terms.get((None, None), Coef(0.0))

    def __add__(self, o):
        out = dict(self.terms)
        for k, c in o.terms.items():
            out[k] = out[k] + c if k in out else c
        return Form(out)

    def __neg__(self):
        return Form({k: Coef(-c.scale, c.arr) for k, c in self.terms.items()})

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        out = {}
        for (bt1, bu1), c1 in self.terms.items():
            for (bt2, bu2), c2 in o.terms.items():
                if (bt1 is not None and bt2 is not None) or (bu1 is not None and bu2 is not None):
                    raise ValueError('expression is not linear in each basis function')
                k = (bt1 if bt1 is not None else bt2, bu1 if bu1 is not None else bu2)
                c = c1 * c2
                out[k] = out[k] + c if k in out else c
        return Form(out)

    def __truediv__(self, o):
        if not o.is_coef():
            raise ValueError('cannot divide by an expression that contains basis functions')
        c = o.coef()
        inv = Coef(1.0 / c.scale, None if c.arr is None else 1.0 / c.arr)
        return self * Form({(None, None): inv})

    def apply(self, fn):
        if not self.is_coef():
            raise ValueError('functions can only be applied to coefficient expressions')
        v = self.coef().value()
        if _is_dev(v):
            import torch
            return Form.const(1.0, getattr(torch, fn.__name__)(v))
        r = fn(v)
        return Form.const(1.0, r) if isinstance(r, np.ndarray) else Form.const(float(r))


def _obj(shape, items):
    a = np.empty(shape, dtype=object)
    a.ravel()[:] = list(items)
    return a


# ---------------------------------------------------------------------------------------------
# lazy expression trees
# ---------------------------------------------------------------------------------------------
class Expr:
    """Node of an expression tree; ``ev(env)`` returns a numpy object array of :class:`Form`."""
    shape = ()

    def __add__(self, o): return BinOp('+', self, as_expr(o))
    def __radd__(self, o): return BinOp('+', as_expr(o), self)
    def __sub__(self, o): return BinOp('-', self, as_expr(o))
    def __rsub__(self, o): return BinOp('-', as_expr(o), self)
    def __mul__(self, o): return BinOp('*', self, as_expr(o))
    def __rmul__(self, o): return BinOp('*', as_expr(o), self)
    def __truediv__(self, o): return BinOp('/', self, as_expr(o))
    def __rtruediv__(self, o): return BinOp('/', as_expr(o), self)
    def __neg__(self): return BinOp('-', as_expr(0.0), self)
    def __pos__(self): return self

    def __getitem__(self, idx):
        return Index(self, idx)

    def __len__(self):
        if not self.shape:
            raise TypeError('scalar expression has no length')
        return self.shape[0]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def is_scalar(self): return self.shape == ()
    def is_vector(self): return len(self.shape) == 1
    def is_matrix(self): return len(self.shape) == 2

    @property
    def T(self):
        return Transpose(self)

    def dot(self, o):
        return dot(self, o)


class Const(Expr):
    def __init__(self, v):
        self.v = np.asarray(v, dtype=float)
        self.shape = self.v.shape

    def ev(self, env):
        return _obj(self.shape, (Form.const(float(x)) for x in self.v.ravel()))


class Literal(Expr):
    """vector / matrix built from scalar expressions"""
    def __init__(self, items, shape):
        self.items, self.shape = [as_expr(i) for i in items], tuple(shape)
        assert all(i.shape == () for i in self.items), 'components must be scalar'

    def ev(self, env):
        return _obj(self.shape, (i.ev(env)[()] for i in self.items))


class BasisFun(Expr):
    def __init__(self, name, role, numcomp=None):
        self.name, self.role = name, role       # role: 'trial' (u, columns) or 'test' (v, rows)
        self.numcomp = numcomp                  # None: scalar; k: vector-valued with k components
        self.shape = () if numcomp is None else (numcomp,)

    def _form(self, comp, deriv):
        slot = (comp, deriv)
        return Form({((None, slot) if self.role == 'trial' else (slot, None)): Coef(1.0)})

    def ev(self, env):
        if self.numcomp is None:
            return _obj((), [self._form(0, 0)])
        return _obj(self.shape, (self._form(c, 0) for c in range(self.numcomp)))


class Input(Expr):
    """named input function or constant parameter; value supplied at instantiation"""
    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)

    def ev(self, env):
        vals = env[self.name]       # ndarray of shape self.shape (+ grid) or python floats
        if self.shape == ():
            return _obj((), [_coef_form(vals)])
        return _obj(self.shape, (_coef_form(vals[idx]) for idx in np.ndindex(*self.shape)))


def _coef_form(v):
    if (isinstance(v, np.ndarray) or (hasattr(v, 'device') and hasattr(v, 'expand'))) and v.ndim > 0:
        return Form.const(1.0, v)
    return Form.const(float(v))


class Measure(Expr):
    """dx: the factor |det J| * GaussWeight is applied on the device.  ds (boundary integrals): the
    same factor times the ratio surface / volume measure, which the assembler supplies as '@ds'."""
    def __init__(self, surface=False):
        self.surface = surface

    def ev(self, env):
        if self.surface:
            return _obj((), [Form.const(1.0, env['@ds'])])
        return _obj((), [Form.const(1.0)])


# module-level measures, as in ``from pyiga.vform import dx, ds`` (``test/test_assemble.py:289,316``)
dx = Measure()
ds = Measure(surface=True)


class BinOp(Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b
        if op in '+-':
            if a.shape != b.shape:
                raise ValueError('incompatible shapes %s and %s' % (a.shape, b.shape))
            self.shape = a.shape
        elif op == '*':
            if a.shape and b.shape:
                if a.shape != b.shape:
                    raise ValueError('elementwise product of shapes %s and %s' % (a.shape, b.shape))
                self.shape = a.shape
            else:
                self.shape = a.shape or b.shape
        else:
            if b.shape:
                raise ValueError('can only divide by scalars')
            self.shape = a.shape

    def ev(self, env):
        A, B = self.a.ev(env), self.b.ev(env)
        fn = {'+': lambda x, y: x + y, '-': lambda x, y: x - y, '*': lambda x, y: x * y,
              '/': lambda x, y: x / y}[self.op]
        if A.shape == B.shape:
            return _obj(A.shape, (fn(x, y) for x, y in zip(A.ravel(), B.ravel())))
        if A.shape == ():
            return _obj(B.shape, (fn(A[()], y) for y in B.ravel()))
        return _obj(A.shape, (fn(x, B[()]) for x in A.ravel()))


class Index(Expr):
    def __init__(self, e, idx):
        if not e.shape:
            raise TypeError('cannot index a scalar expression')
        self.e, self.idx = e, idx
        self.shape = np.shape(np.empty(e.shape)[idx])

    def ev(self, env):
        r = self.e.ev(env)[self.idx]
        return r if isinstance(r, np.ndarray) else _obj((), [r])


class Transpose(Expr):
    def __init__(self, e):
        self.e, self.shape = e, tuple(reversed(e.shape))

    def ev(self, env):
        return self.e.ev(env).T


class Grad(Expr):
    """gradient of a basis function: vector for scalar functions, Jacobian (rows = components) for
    vector-valued ones"""
    def __init__(self, e, dim):
        if isinstance(e, Index) and isinstance(e.e, BasisFun) and isinstance(e.idx, int):
            self.e, self.comp = e.e, e.idx          # gradient of one component of a vector function
        elif isinstance(e, BasisFun):
            self.e, self.comp = e, None
        else:
            raise NotImplementedError('grad() is implemented for the basis functions u and v only')
        self.dim = dim
        vec = self.e.numcomp is not None and self.comp is None
        self.shape = (self.e.numcomp, dim) if vec else (dim,)

    def ev(self, env):
        if len(self.shape) == 2:
            return _obj(self.shape, (self.e._form(c, 1 + a) for c in range(self.shape[0]) for a in range(self.dim)))
        return _obj(self.shape, (self.e._form(self.comp or 0, 1 + a) for a in range(self.dim)))


class Contract(Expr):
    """inner (full contraction) and dot (matrix/vector products)"""
    def __init__(self, kind, a, b):
        self.kind, self.a, self.b = kind, a, b
        if kind == 'inner':
            if a.shape != b.shape:
                raise ValueError('incompatible shapes for inner product')
            self.shape = ()
        else:
            if not a.shape or not b.shape or a.shape[-1] != b.shape[0]:
                raise ValueError('incompatible shapes for dot product')
            self.shape = a.shape[:-1] + b.shape[1:]

    def ev(self, env):
        A, B = self.a.ev(env), self.b.ev(env)
        if self.kind == 'inner':
            acc = None
            for x, y in zip(A.ravel(), B.ravel()):
                acc = x * y if acc is None else acc + x * y
            return _obj((), [acc])
        A2 = A.reshape(-1, A.shape[-1])
        B2 = B.reshape(B.shape[0], -1)
        out = []
        for i in range(A2.shape[0]):
            for j in range(B2.shape[1]):
                acc = None
                for k in range(A2.shape[1]):
                    t = A2[i, k] * B2[k, j]
                    acc = t if acc is None else acc + t
                out.append(acc)
        return _obj(self.shape, out)


class Func(Expr):
    def __init__(self, fn, e):
        self.fn, self.e, self.shape = fn, e, e.shape

    def ev(self, env):
        A = self.e.ev(env)
        return _obj(A.shape, (x.apply(self.fn) for x in A.ravel()))


def as_expr(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, (tuple, list)):
        items = [as_expr(i) for i in x]
        if all(i.shape == () for i in items):
            return Literal(items, (len(items),))
        if all(len(i.shape) == 1 for i in items):
            flat = [c for i in items for c in i]
            return Literal(flat, (len(items), len(items[0])))
        raise ValueError('cannot convert nested sequence to an expression')
    return Const(x)


as_vector = as_expr
as_matrix = as_expr

# operators available in form strings (names as in pyiga/vform.py:1518-1733)


def _bfun_dim(e):
    return (e.e if isinstance(e, Index) else e)._dim


def grad(e, dims=None, parametric=False):
    if parametric or dims is not None:
        raise NotImplementedError('parametric / partial gradients are not part of the device path')
    e = as_expr(e)
    return Grad(e, _bfun_dim(e))


def div(e, parametric=False):
    """divergence of a vector-valued basis function"""
    return tr(grad(e, parametric=parametric))


def Dx(e, k, times=1, parametric=False):
    if times != 1 or parametric:
        raise NotImplementedError('only first physical derivatives are part of the device path')
    return grad(e)[k]


def inner(a, b):
    return Contract('inner', as_expr(a), as_expr(b))


def dot(a, b):
    return Contract('dot', as_expr(a), as_expr(b))


def tr(A):
    A = as_expr(A)
    assert A.is_matrix() and A.shape[0] == A.shape[1]
    out = A[0, 0]
    for i in range(1, A.shape[0]):
        out = out + A[i, i]
    return out


def outer(a, b):
    a, b = as_expr(a), as_expr(b)
    return Literal([x * y for x in a for y in b], (len(a), len(b)))


def cross(a, b):
    """cross product of two 3-vectors (``pyiga/vform.py`` cross)"""
    a, b = as_expr(a), as_expr(b)
    if a.shape != (3,) or b.shape != (3,):
        raise ValueError('cross() needs two vectors of length 3')
    return Literal([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], (3,))


def norm(x):
    return sqrt(inner(x, x))


def sqrt(x): return Func(np.sqrt, as_expr(x))
def exp(x): return Func(np.exp, as_expr(x))
def log(x): return Func(np.log, as_expr(x))
def sin(x): return Func(np.sin, as_expr(x))
def cos(x): return Func(np.cos, as_expr(x))
def tan(x): return Func(np.tan, as_expr(x))


# ---------------------------------------------------------------------------------------------
# VForm
# ---------------------------------------------------------------------------------------------
class VForm:
    """Abstract description of a variational form (API after ``pyiga/vform.py:162-350``)."""

    def __init__(self, dim, geo_dim=None, boundary=False, arity=2, spacetime=False):
        if spacetime:
            raise NotImplementedError('general space-time forms: use heat_st_vf / wave_st_vf or a reference VForm object')
        self.dim, self.arity = dim, arity
        # integrals over a dim-dimensional manifold in R^(dim+1): `geo` maps R^dim -> R^geo_dim
        self.geo_dim = dim if geo_dim is None else int(geo_dim)
        if self.geo_dim not in (dim, dim + 1):
            raise ValueError('geo_dim must be dim (volume) or dim + 1 (surface integral)')
        if boundary and self.geo_dim != dim:
            raise NotImplementedError('boundary integrals of surface forms')
        self.boundary = bool(boundary)
        self.ds = Measure(surface=True)         # boundary forms: integrate with `* ds`
        self.normal = Input('@n', (self.geo_dim,))      # outer unit normal (boundary forms) / surface normal
        self.vec = False
        self.exprs = []
        self.inputs = []        # [(name, shape, physical, updatable)]
        self.params = []        # [(name, shape)]
        self.dx = Measure()
        self.Geo = Input('@x', (self.geo_dim,))
        self.numcomp = (None, None)

    def basisfuns(self, components=(None, None), spaces=(0, 0)):
        spaces = tuple(spaces)
        if self.arity == 2 and spaces not in ((0, 0), (0, 1)):
            # the reference's generator only compiles trial functions in space 0 / test functions in space 1
            raise NotImplementedError('two-space forms need the trial function in space 0 and the test function in space 1')
        if self.arity == 1 and any(s != 0 for s in spaces[-1:]):
            raise NotImplementedError('linear forms over the second space are not part of the device path')
        self.spaces = spaces if self.arity == 2 else (0, 0)
        components = tuple(components)
        if self.arity == 1:
            components = components[-1:]
        if all(c in (None, 1) for c in components):
            components = len(components) * (None,)      # scalar assembler
        else:
            components = tuple(1 if c is None else int(c) for c in components)
            self.vec = True
        v = BasisFun('v', 'test', components[-1])
        v._dim = self.geo_dim                       # number of physical coordinates = length of grad()
        self.numcomp = (components[0] if self.arity == 2 else None, components[-1])   # (trial, test)
        if self.arity == 1:
            return v
        u = BasisFun('u', 'trial', components[0])
        u._dim = self.geo_dim
        return u, v

    def input(self, name, shape=(), physical=False, updatable=False):
        self.inputs.append((name, tuple(shape), bool(physical), bool(updatable)))
        return Input(name, shape)

    def parameter(self, name, shape=()):
        self.params.append((name, tuple(shape)))
        return Input(name, shape)

    def add(self, expr):
        expr = as_expr(expr)
        if expr.shape != ():
            raise ValueError('the integrand must be scalar (use inner() / dot() to contract vectors)')
        self.exprs.append(expr)

    def num_spaces(self):
        return 2 if getattr(self, 'spaces', (0, 0)) == (0, 1) else 1


def _mentions(expr, name):
    stack, seen = [expr], set()
    while stack:
        e = stack.pop()
        if id(e) in seen:
            continue
        seen.add(id(e))
        if isinstance(e, Input) and e.name == name:
            return True
        for v in vars(e).values():
            if isinstance(v, Expr):
                stack.append(v)
            elif isinstance(v, (list, tuple)):
                stack.extend(x for x in v if isinstance(x, Expr))
    return False


def _check_input_field(kvs, f):
    """(shape, physical): spline functions are parametric, other callables physical and probed at
    the midpoint of the parameter box (``pyiga/vform.py:1791-1802``)."""
    if hasattr(f, 'grid_eval') and hasattr(f, 'kvs'):
        return tuple(f.output_shape()), False
    mid = tuple(0.5 * (kv.support()[0] + kv.support()[1]) for kv in kvs)
    try:
        return np.shape(f(*mid)), True
    except TypeError:
        # a function of the physical coordinates of a surface form takes one argument more
        return np.shape(f(*(mid + (mid[-1],)))), True


def parse_vf(expr, kvs, args=dict(), bfuns=None, boundary=False, updatable=[]):
    """Parse a form string like ``'(inner(c * grad(u), grad(v)) + inner(b, grad(u)) * v) * dx'``
    (``pyiga/vform.py:1804-1887``)."""
    if not all(hasattr(kv, 'kv') for kv in kvs):
        if all(hasattr(kv, 'kv') for kv in kvs[0]):
            kvs = kvs[0]
        else:
            raise ValueError('expected a tensor product spline space in `kvs`')
    dim = len(kvs)
    words = set(re.findall(r"[^\d\W]\w*", expr))
    if 'ds' in words:
        if 'dx' in words:
            raise RuntimeError("got both 'dx' and 'ds' - is this a volume or a surface integral?")
    spaces = None
    if bfuns is None:
        names, comps = sorted(words & {'u', 'v'}), None
    else:
        names, comps, spaces = [], [], []
        for bf in bfuns:
            bf = (bf,) if isinstance(bf, str) else tuple(bf)
            names.append(bf[0])
            comps.append(bf[1] if len(bf) > 1 else 1)
            spaces.append(bf[2] if len(bf) > 2 else 0)
    if len(names) not in (1, 2):
        raise ValueError('arity should be 1 or 2')
    geo_dim = dim + 1 if ('ds' in words and not boundary) else dim
    vf = VForm(dim=dim, geo_dim=geo_dim, boundary=bool(boundary), arity=len(names))
    loc = {}
    if vf.arity == 1:
        loc[names[0]] = vf.basisfuns(components=tuple(comps) if comps else (None,))
    else:
        u, v = vf.basisfuns(components=tuple(comps) if comps else (None, None),
                            spaces=tuple(spaces) if spaces else (0, 0))
        loc[names[0]], loc[names[1]] = u, v
    for name in sorted(set(args.keys()) & words):
        if callable(args[name]):
            shp, phys = _check_input_field(kvs, args[name])
            loc[name] = vf.input(name, shape=shp, physical=phys, updatable=(name in updatable))
        else:
            loc[name] = vf.parameter(name, shape=np.shape(args[name]))
    if 'x' in words and 'x' not in args:
        loc['x'] = vf.Geo
    if 'n' in words and 'n' not in args:
        loc['n'] = vf.normal
    ns = dict(globals())
    ns['dx'] = vf.dx
    ns['ds'] = vf.ds
    ns.update(loc)
    vf.add(eval(expr, ns))
    vf.source = expr
    return vf


def mass_vf(dim):
    vf = VForm(dim)
    u, v = vf.basisfuns()
    vf.add(u * v * vf.dx)
    return vf


def divdiv_vf(dim):
    vf = VForm(dim)
    u, v = vf.basisfuns(components=(dim, dim))
    vf.add(div(u) * div(v) * vf.dx)
    return vf


def stiffness_vf(dim):
    vf = VForm(dim)
    u, v = vf.basisfuns()
    vf.add(inner(grad(u), grad(v)) * vf.dx)
    return vf


class _SpaceTimeForm:
    """The predefined space-time forms (``pyiga/vform.py:1759-1772``).  General space-time expressions are not
    part of this front end (reference ``VForm(dim, spacetime=True)`` objects are: :mod:`pyiga_b200.refvform`);
    the two predefined ones map to the assembler classes of :mod:`pyiga_b200.spacetime`."""
    arity, vec, spacetime = 2, False, True

    def __init__(self, dim, wave):
        self.dim = self.geo_dim = dim
        self.wave = wave

    def assembler_class(self):
        from . import spacetime
        return getattr(spacetime, '%sAssembler_ST%dD' % ('Wave' if self.wave else 'Heat', self.dim))


def heat_st_vf(dim):
    return _SpaceTimeForm(dim, wave=False)


def wave_st_vf(dim):
    return _SpaceTimeForm(dim, wave=True)


# ---------------------------------------------------------------------------------------------
# "compilation": VForm -> assembler class bound to the device pipeline
# ---------------------------------------------------------------------------------------------
def _is_dev(a):
    """a torch tensor (coefficient arrays may live on the GPU, see GenericFormAssembler._eval_input)"""
    return hasattr(a, 'device') and hasattr(a, 'expand')


def _dev_broadcast(v, like, grid_shape):
    import torch
    if not _is_dev(v):
        v = torch.as_tensor(np.asarray(v, dtype=float), dtype=torch.float64, device=like.device)
    return v.to(torch.float64).expand(grid_shape)


def _grid_values(f, shape, coords, grid_shape):
    """evaluate a callable at coordinate arrays (x, y, z order) and bring the result to
    shape + grid_shape (``pyiga/utils.py:8-31`` _ensure_grid_shape).  `coords` may be device tensors:
    the callable then runs on the GPU — behind the numpy protocols of :class:`pyiga_b200._devarray.DevArray`
    (arithmetic, comparisons, ``np.sin`` / ``np.where`` / ...), or on the raw tensors if it does something the
    wrapper does not know."""
    if _is_dev(coords[0]):
        from ._devarray import DevArray, unwrap
        try:
            vals = unwrap(f(*[DevArray(c) for c in coords]))
        except Exception:
            vals = f(*coords)
    else:
        vals = f(*coords)
    if _is_dev(coords[0]):
        if shape == ():
            return _dev_broadcast(vals, coords[0], grid_shape)
        comps = np.empty(shape, dtype=object)
        whole = vals if _is_dev(vals) and tuple(vals.shape[:len(shape)]) == shape else None
        for idx in np.ndindex(*shape):
            comps[idx] = _dev_broadcast(whole[idx] if whole is not None else _nested_get(vals, idx), coords[0], grid_shape)
        return comps
    if shape == ():
        return np.broadcast_to(np.asarray(vals, dtype=float), grid_shape)
    comps = np.empty(shape, dtype=object)
    vals_arr = vals if isinstance(vals, np.ndarray) and vals.shape[:len(shape)] == shape else None
    for idx in np.ndindex(*shape):
        c = vals_arr[idx] if vals_arr is not None else _nested_get(vals, idx)
        comps[idx] = np.broadcast_to(np.asarray(c, dtype=float), grid_shape)
    return comps


def _nested_get(v, idx):
    for i in idx:
        v = v[i]
    return v


def compile_vform(vf, on_demand=False, verbose=False):
    """Return an assembler *class* for the form, like ``pyiga.compile.compile_vform``
    (``pyiga/compile.py:120-132``); no code is generated — the class analyses the form when it is
    instantiated on a concrete space.  `on_demand` assemblers of the reference evaluate their fields
    only inside a bounding box of cells (``bbox=`` of the constructor, used by
    ``HDiscretization._assemble_level``): forms given as reference ``VForm`` objects honour the box
    (:mod:`pyiga_b200.refvform`); for forms of this module the fields of the whole patch are one K2
    launch, so the box is accepted and every entry stays valid."""
    from . import refvform
    if isinstance(vf, _SpaceTimeForm):
        return vf.assembler_class()
    if refvform.is_reference_vform(vf):
        # a pyiga.vform.VForm: interpreted after its own finalize(); `on_demand`
        # assemblers evaluate their inputs only inside the bounding box given to the constructor
        return refvform.compile_vform(vf, on_demand=on_demand)
    input_shapes = {'geo': (vf.geo_dim,)}
    input_shapes.update({name: shape for name, shape, _, _ in vf.inputs})
    param_shapes = {name: shape for name, shape in vf.params}

    from . import assemblers

    class VFormAssembler(assemblers.GenericFormAssembler):
        _vf = vf

        @classmethod
        def inputs(cls):
            return dict(input_shapes)

        @classmethod
        def parameters(cls):
            return dict(param_shapes)

    VFormAssembler.__name__ = 'VFormAssembler%dD' % vf.dim
    return VFormAssembler
