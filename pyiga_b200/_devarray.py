"""numpy protocol over device tensors, for user-supplied coefficient callables.

The reference evaluates an input function ``f(x, y, z)`` on the host by calling it with numpy
coordinate arrays (``pyiga/utils.py:8-52``, ``pyiga/codegen/cython.py:465-484``).  Here the physical
Gauss points already live on the GPU, so the same callable is handed :class:`DevArray` coordinates: a
thin wrapper that answers Python arithmetic, comparisons, indexing and the numpy dispatch protocols
(``__array_ufunc__`` for ``np.sin`` / ``np.exp`` / ``np.maximum`` ..., ``__array_function__`` for
``np.where`` / ``np.stack`` / ``np.zeros_like`` ...) with the corresponding elementwise device
operations.  The callable therefore runs unchanged on (128*4)^3 points without the points or the
result crossing PCIe.  Anything the wrapper does not know raises, and the caller falls back to the
host evaluation of the reference (``GenericFormAssembler._eval_input``).
"""
import numpy as np

# numpy ufunc name -> torch function name (elementwise, same semantics for float64 / bool)
_UFUNCS = {
    'add': 'add', 'subtract': 'sub', 'multiply': 'mul', 'divide': 'div', 'true_divide': 'div',
    'negative': 'neg', 'positive': 'positive', 'power': 'pow', 'float_power': 'pow', 'square': 'square',
    'sqrt': 'sqrt', 'reciprocal': 'reciprocal', 'absolute': 'abs', 'fabs': 'abs', 'sign': 'sign',
    'exp': 'exp', 'exp2': 'exp2', 'expm1': 'expm1', 'log': 'log', 'log2': 'log2', 'log10': 'log10', 'log1p': 'log1p',
    'sin': 'sin', 'cos': 'cos', 'tan': 'tan', 'arcsin': 'asin', 'arccos': 'acos', 'arctan': 'atan', 'arctan2': 'atan2',
    'sinh': 'sinh', 'cosh': 'cosh', 'tanh': 'tanh', 'arcsinh': 'asinh', 'arccosh': 'acosh', 'arctanh': 'atanh',
    'hypot': 'hypot', 'maximum': 'maximum', 'minimum': 'minimum', 'fmax': 'fmax', 'fmin': 'fmin',
    'floor': 'floor', 'ceil': 'ceil', 'trunc': 'trunc', 'rint': 'round', 'remainder': 'remainder', 'mod': 'remainder',
    'fmod': 'fmod', 'floor_divide': 'floor_divide', 'heaviside': 'heaviside',
    'greater': 'gt', 'greater_equal': 'ge', 'less': 'lt', 'less_equal': 'le', 'equal': 'eq', 'not_equal': 'ne',
    'logical_and': 'logical_and', 'logical_or': 'logical_or', 'logical_not': 'logical_not', 'logical_xor': 'logical_xor',
    'isnan': 'isnan', 'isfinite': 'isfinite', 'isinf': 'isinf',
}
_BINARY_NEED_TENSORS = {'atan2', 'hypot', 'maximum', 'minimum', 'fmax', 'fmin', 'heaviside', 'logical_and', 'logical_or',
                        'logical_xor', 'remainder', 'fmod', 'floor_divide'}


def _torch():
    import torch
    return torch


def unwrap(v):
    """DevArray -> tensor, recursively through tuples / lists; everything else unchanged"""
    if isinstance(v, DevArray):
        return v.t
    if isinstance(v, (tuple, list)):
        return type(v)(unwrap(x) for x in v)
    return v


class DevArray:
    """A device tensor behind the numpy array protocols (see the module docstring)."""
    __slots__ = ('t',)
    __array_priority__ = 1000.0

    def __init__(self, t):
        self.t = t

    # ---- array attributes --------------------------------------------------------------------
    shape = property(lambda self: tuple(self.t.shape))
    ndim = property(lambda self: self.t.dim())
    size = property(lambda self: self.t.numel())
    dtype = property(lambda self: np.dtype(str(self.t.dtype).replace('torch.', '')))
    T = property(lambda self: DevArray(self.t.permute(*reversed(range(self.t.dim())))))

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, idx):
        return DevArray(self.t[unwrap(idx)])

    def __array__(self, *a, **k):
        raise TypeError('a device array is not converted to numpy implicitly')

    # ---- operands ----------------------------------------------------------------------------
    def _tensor(self, x, need_tensor=False):
        if isinstance(x, DevArray):
            return x.t
        torch = _torch()
        if torch.is_tensor(x):
            return x
        if isinstance(x, (bool, int, float, np.floating, np.integer, np.bool_)) and not need_tensor:
            return x.item() if isinstance(x, np.generic) else x
        a = np.asarray(x)
        if a.dtype.kind not in 'fiub':
            raise TypeError('operand of type %s on a device array' % a.dtype)
        return torch.as_tensor(a.astype(float) if a.dtype.kind in 'fiu' else a, device=self.t.device)

    def _call(self, name, *inputs):
        torch = _torch()
        args = [self._tensor(x, need_tensor=name in _BINARY_NEED_TENSORS) for x in inputs]
        if not torch.is_tensor(args[0]):         # scalar first operand (2.0 ** x, np.arctan2(1.0, x), ...)
            args[0] = torch.as_tensor(float(args[0]), dtype=torch.float64, device=self.t.device)
        return DevArray(getattr(torch, name)(*args))

    # ---- Python operators --------------------------------------------------------------------
    def __add__(self, o): return self._call('add', self, o)
    def __radd__(self, o): return self._call('add', o, self)
    def __sub__(self, o): return self._call('sub', self, o)
    def __rsub__(self, o): return self._call('sub', o, self)
    def __mul__(self, o): return self._call('mul', self, o)
    def __rmul__(self, o): return self._call('mul', o, self)
    def __truediv__(self, o): return self._call('div', self, o)
    def __rtruediv__(self, o): return self._call('div', o, self)
    def __pow__(self, o): return self._call('pow', self, o)
    def __rpow__(self, o): return self._call('pow', o, self)
    def __mod__(self, o): return self._call('remainder', self, o)
    def __floordiv__(self, o): return self._call('floor_divide', self, o)
    def __neg__(self): return DevArray(-self.t)
    def __pos__(self): return self
    def __abs__(self): return DevArray(self.t.abs())
    def __lt__(self, o): return self._call('lt', self, o)
    def __le__(self, o): return self._call('le', self, o)
    def __gt__(self, o): return self._call('gt', self, o)
    def __ge__(self, o): return self._call('ge', self, o)
    def __eq__(self, o): return self._call('eq', self, o)
    def __ne__(self, o): return self._call('ne', self, o)
    def __and__(self, o): return self._call('logical_and', self, o)
    def __or__(self, o): return self._call('logical_or', self, o)
    def __invert__(self): return DevArray(self.t.logical_not())
    __hash__ = None

    def __bool__(self):
        raise TypeError('the truth value of a device array is ambiguous (use np.where)')

    # ---- array methods callables commonly use ---------------------------------------------------
    def astype(self, dtype, **kw):
        torch = _torch()
        return DevArray(self.t.to({'f': torch.float64, 'b': torch.bool, 'i': torch.int64}[np.dtype(dtype).kind]))

    def copy(self):
        return DevArray(self.t.clone())

    def clip(self, a_min=None, a_max=None):
        return DevArray(self.t.clamp(min=a_min, max=a_max))

    def reshape(self, *shape):
        return DevArray(self.t.reshape(*shape))

    # ---- numpy dispatch ------------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        name = _UFUNCS.get(ufunc.__name__)
        if method != '__call__' or name is None or kwargs:
            return NotImplemented
        return self._call(name, *inputs)

    def __array_function__(self, func, types, args, kwargs):
        impl = _FUNCTIONS.get(func.__name__)
        if impl is None:
            return NotImplemented
        return impl(self, *args, **kwargs)


def _first(args):
    for a in args:
        if isinstance(a, DevArray):
            return a
    raise TypeError('no device operand')


def _where(self, cond, x, y):
    torch = _torch()
    like = _first((cond, x, y))
    c = like._tensor(cond, need_tensor=True).to(torch.bool)
    tx, ty = like._tensor(x), like._tensor(y)
    if not torch.is_tensor(tx) and not torch.is_tensor(ty):     # two Python scalars: keep float64
        tx = torch.as_tensor(float(tx), dtype=torch.float64, device=c.device)
    return DevArray(torch.where(c, tx, ty))


def _stack(self, arrays, axis=0):
    like = _first(arrays)
    torch = _torch()
    ts = [like._tensor(a, need_tensor=True) for a in arrays]
    shape = torch.broadcast_shapes(*[tuple(t.shape) for t in ts])
    return DevArray(torch.stack([t.to(torch.float64).expand(shape) for t in ts], dim=axis))


def _filled(value):
    def make(self, a, dtype=None, **kw):
        torch = _torch()
        fill = value if value is not None else kw.pop('fill_value')
        return DevArray(torch.full(tuple(a.shape), float(fill), dtype=torch.float64, device=a.t.device))
    return make


def _full_like(self, a, fill_value, dtype=None, **kw):
    torch = _torch()
    return DevArray(torch.full(tuple(a.shape), float(fill_value), dtype=torch.float64, device=a.t.device))


def _clip(self, a, a_min=None, a_max=None, **kw):
    return a.clip(a_min, a_max)


_FUNCTIONS = {
    'where': _where, 'stack': _stack, 'zeros_like': _filled(0.0), 'ones_like': _filled(1.0), 'full_like': _full_like,
    'clip': _clip, 'shape': lambda self, a: a.shape, 'ndim': lambda self, a: a.ndim,
    'abs': lambda self, a: abs(a), 'copy': lambda self, a, **kw: a.copy(),
    'broadcast_to': lambda self, a, shape, **kw: DevArray(a.t.expand(tuple(shape))),
}
