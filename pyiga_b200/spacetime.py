"""Space-time assemblers of the reference's predefined set: ``HeatAssembler_ST{2,3}D`` and
``WaveAssembler_ST{2,3}D`` (``pyiga/assemblers.pyx:351-690, 1542-1957``, generated from ``heat_st_vf`` /
``wave_st_vf``, ``pyiga/vform.py:1759-1772``).

The last coordinate of the geometry map is time (tensor axis 0).  As in the reference the map is treated
as a space-time cylinder (``pyiga/vform.py:575-586``): time derivatives stay parametric, space gradients
are pulled back with the spatial block of the inverse of the FULL Jacobian, and the weight is
``GaussWeight * |det Jac|`` of the full Jacobian.  With ``S = W * sum_k JacInv[i,k] JacInv[j,k]`` over the
space coordinates i, j, k the forms are

    heat:  sum_ij S_ij d_i u d_j v  +  W d_t u v
    wave:  W d_tt u d_t v           +  sum_ij S_ij d_i u d_j d_t v

i.e. custom device forms whose coefficients are S and W; the wave form uses the second / mixed derivative
slots of the C ABI (``PB200_SLOT_EXT``).  The matrices come from the sum-factorised walks.
"""
import numpy as np

from .assemblers import GenericFormAssembler
from .quadrature import make_tensor_quadrature
from .refvform import _ParametricBlock


def _slot(dim, **orders):
    """slot code of derivative orders given per COORDINATE index (x = 0, ..., time = dim - 1)"""
    o = [0] * dim
    for c, n in orders.items():
        o[dim - 1 - int(c[1:])] = n            # coordinate c <-> tensor axis dim - 1 - c
    if sum(o) == 0:
        return 0
    if sum(o) == 1:
        return 1 + o.index(1)
    return 16 + sum(n * 3 ** k for k, n in enumerate(o))


class _SpaceTimeAssembler(GenericFormAssembler):
    _dim = 2
    _wave = False

    @classmethod
    def inputs(cls):
        return {'geo': (cls._dim,)}

    @classmethod
    def parameters(cls):
        return {}

    def __init__(self, kvs, geo, bbox=None):
        kvs = tuple(kvs)
        d = self._dim
        assert len(kvs) == d, "Assembler requires %d knot vectors" % d
        assert geo.sdim == d, "Geometry has wrong source dimension"
        assert geo.dim == d, "Geometry has wrong dimension"
        self.arity = 2
        self.nqp = max(kv.p for kv in kvs) + 1
        self.kvs = (kvs, kvs)
        self._geo, self._args = geo, {'geo': geo}
        self.bbox = bbox
        self._bd, self._surface, self._vec = None, False, False
        grid, w = make_tensor_quadrature([np.asarray(kv.mesh) for kv in kvs], self.nqp)
        self.gaussgrid = grid
        shape = tuple(len(g) for g in grid)
        gw = np.ones(shape)
        for k in range(d):
            gw = gw * np.asarray(w[k]).reshape([-1 if i == k else 1 for i in range(d)])
        J = np.asarray(geo.grid_jacobian(grid), dtype=float)          # (..., d, d), coordinates x, (y,) t
        Jinv = np.linalg.inv(J)
        W = gw * np.abs(np.linalg.det(J))
        sp = d - 1                                                       # number of space coordinates
        S = W[..., None, None] * np.einsum('...ik,...jk->...ij', Jinv[..., :sp, :sp], Jinv[..., :sp, :sp])
        t = 'c%d' % (d - 1)
        coefs = {}
        for i in range(sp):
            for j in range(sp):
                if self._wave:      # d_i u  *  d_j d_t v
                    key = (_slot(d, **{'c%d' % j: 1, t: 1}), _slot(d, **{'c%d' % i: 1}))
                else:               # d_i u  *  d_j v
                    key = (_slot(d, **{'c%d' % j: 1}), _slot(d, **{'c%d' % i: 1}))
                coefs[key] = S[..., i, j]
        if self._wave:
            coefs[(_slot(d, **{t: 1}), _slot(d, **{t: 2}))] = W         # d_tt u * d_t v
        else:
            coefs[(0, _slot(d, **{t: 1}))] = W                          # d_t u * v
        self.blocks = {(0, 0): _ParametricBlock(kvs, self.nqp, d, 2, coefs, shape)}
        self.dev = self.blocks[(0, 0)].dev

    def update(self, **kwargs):
        raise NotImplementedError('update() of space-time assemblers: rebuild the assembler')


class HeatAssembler_ST2D(_SpaceTimeAssembler):
    _dim, _wave = 2, False


class HeatAssembler_ST3D(_SpaceTimeAssembler):
    _dim, _wave = 3, False


class WaveAssembler_ST2D(_SpaceTimeAssembler):
    _dim, _wave = 2, True


class WaveAssembler_ST3D(_SpaceTimeAssembler):
    _dim, _wave = 3, True
