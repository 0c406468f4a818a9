"""Spline geometry maps used by the assembly path.

Mirrors the constructors of the reference's ``pyiga.geometry`` that the hot path
and its tests use (``pyiga/geometry.py:27-123`` NurbsFunc, ``:445-589`` stock
geometries, ``:595-809`` curves and tensor products).  Objects here only carry
knot vectors + control nets; Jacobians on the Gauss grid are evaluated on the
device (``csrc/geo_fields.cuh``), including the NURBS quotient rule
(reference ``pyiga/geometry.py:17-25``).
"""
import functools
import numpy as np

from . import bspline
from .bspline import BSplineFunc, _SplineFuncBase, _as_kv_tuple


class NurbsFunc(_SplineFuncBase):
    """Tensor-product NURBS function.  ``coeffs`` is stored in homogeneous,
    pre-multiplied form with the weight as last component, exactly like the
    reference (``pyiga/geometry.py:51-92``)."""

    _rational = True

    def __init__(self, kvs, coeffs, weights, premultiplied=False):
        self.kvs = _as_kv_tuple(kvs)
        self.sdim = len(self.kvs)
        N = tuple(kv.numdofs for kv in self.kvs)
        coeffs = np.asanyarray(coeffs, dtype=float)
        if coeffs.ndim == 1:
            assert coeffs.shape[0] == np.prod(N), "Wrong length of coefficient vector"
            coeffs = coeffs.reshape(N)
        assert N == coeffs.shape[:self.sdim], "Wrong shape of coefficients"
        tail = coeffs.shape[self.sdim:]
        assert len(tail) <= 1, 'Tensor-valued NURBS functions not implemented'
        self._isscalar = (len(tail) == 0)
        dim = 1 if self._isscalar else tail[0]
        if weights is None:
            assert dim > 1, 'Weights must be specified in the coeffs array'
            dim -= 1
            C = np.array(coeffs, dtype=float)
        else:
            weights = np.asanyarray(weights, dtype=float)
            assert weights.shape == N, 'Wrong shape of weights array'
            body = coeffs[..., None] if self._isscalar else coeffs
            C = np.concatenate((body, weights[..., None]), axis=-1)
        if not premultiplied:
            C[..., :-1] *= C[..., -1:]
        self.coeffs = C
        self.dim = dim

    def output_shape(self):
        return () if self._isscalar else (self.dim,)

    def boundary(self, bdspec):
        """One side of the boundary as a :class:`NurbsFunc` (``pyiga/geometry.py:188-209``)."""
        from .bspline import _parse_bdspec
        axis, side = _parse_bdspec(bdspec, self.sdim)
        slices = self.sdim * [slice(None)]
        slices[axis] = 0 if side == 0 else -1
        kvs = list(self.kvs)
        del kvs[axis]
        return NurbsFunc(kvs, self.coeffs[tuple(slices)], weights=None, premultiplied=True)

    def coeffs_weights(self):
        """Non-premultiplied coefficients and the weights."""
        W = self.coeffs[..., -1]
        C = self.coeffs[..., :-1] / W[..., None]
        if self._isscalar:
            C = C[..., 0]
        return C, W

    def copy(self):
        C = self.coeffs.copy()
        out = NurbsFunc(self.kvs, C, None, premultiplied=True)
        if self._isscalar:
            out._isscalar, out.dim = True, 1
        return out

    def as_nurbs(self):
        return self

    def as_vector(self):
        if not self._isscalar:
            return self
        out = self.copy()
        out._isscalar = False
        return out

    def translate(self, offset):
        C, W = self.coeffs_weights()
        return NurbsFunc(self.kvs, C + np.asarray(offset), W)

    def scale(self, factor):
        C, W = self.coeffs_weights()
        return NurbsFunc(self.kvs, C * np.asarray(factor), W)

    def apply_matrix(self, A):
        assert self.is_vector(), 'Can only apply matrices to vector-valued functions'
        C, W = self.coeffs_weights()
        return NurbsFunc(self.kvs, np.matmul(np.asarray(A), C[..., None])[..., 0], W)

    def rotate_2d(self, angle):
        assert self.dim == 2, 'Must be 2D vector function'
        c, s = np.cos(angle), np.sin(angle)
        return self.apply_matrix([[c, -s], [s, c]])


# ---------------------------------------------------------------------------
# curves
# ---------------------------------------------------------------------------

def line_segment(x0, x1, support=(0.0, 1.0), intervals=1):
    """Straight line from `x0` to `x1` as a linear spline (``pyiga/geometry.py:595-614``)."""
    x0 = np.atleast_1d(np.array(x0, dtype=float)).ravel()
    x1 = np.atleast_1d(np.array(x1, dtype=float)).ravel()
    assert x0.size == x1.size, 'Vectors must have same dimension'
    t = np.linspace(0.0, 1.0, intervals + 1)[:, None]
    kv = bspline.make_knots(1, support[0], support[1], intervals)
    return BSplineFunc(kv, (1 - t) * x0 + t * x1)


def _arc(alpha, r, npts, kv, w):
    angles = np.linspace(0, alpha, npts)
    pts = np.column_stack((np.cos(angles), np.sin(angles)))
    return NurbsFunc(kv, r * pts, weights=np.asarray(w, dtype=float), premultiplied=True)


def circular_arc_3pt(alpha, r=1.0):
    assert 0.0 < alpha < np.pi, 'Invalid angle'
    return _arc(alpha, r, 3, bspline.make_knots(2, 0.0, 1.0, 1), [1.0, np.cos(alpha / 2), 1.0])


def circular_arc_5pt(alpha, r=1.0):
    w = np.cos(alpha / 4)
    return _arc(alpha, r, 5, bspline.make_knots(2, 0.0, 1.0, 2, mult=2), [1.0, w, 1.0, w, 1.0])


def circular_arc_7pt(alpha, r=1.0):
    w = np.cos(alpha / 6)
    return _arc(alpha, r, 7, bspline.make_knots(2, 0.0, 1.0, 3, mult=2), [1, w, 1, w, 1, w, 1])


def circular_arc(alpha, r=1.0):
    if 0.0 < alpha < np.pi:
        return circular_arc_3pt(alpha, r)
    if np.pi <= alpha <= 2 * np.pi:
        return circular_arc_7pt(alpha, r)
    raise ValueError('invalid angle {}'.format(alpha))


def semicircle(r=1.0):
    return circular_arc_5pt(np.pi, r)


def circle(r=1.0):
    return circular_arc_7pt(2 * np.pi, r)


# ---------------------------------------------------------------------------
# combining functions
# ---------------------------------------------------------------------------

def _outer_shapes(Ca, Cb, sda, sdb):
    """Reshape two coefficient arrays so that their source axes do not overlap."""
    Sa, Va = Ca.shape[:sda], Ca.shape[sda:]
    Sb, Vb = Cb.shape[:sdb], Cb.shape[sdb:]
    return (Ca.reshape(Sa + (1,) * sdb + Va), Cb.reshape((1,) * sda + Sb + Vb), Sa + Sb)


def tensor_product(G1, G2, *Gs):
    """``G(x,y) = G2(x) x G1(y)`` — join the outputs (``pyiga/geometry.py:755-809``)."""
    if Gs:
        return tensor_product(G1, tensor_product(G2, *Gs))
    G1 = G1.as_vector() if G1.is_scalar() else G1
    G2 = G2.as_vector() if G2.is_scalar() else G2
    assert G1.is_vector() and G2.is_vector(), 'only implemented for scalar- or vector-valued functions'
    rational = G1._rational or G2._rational
    if rational:
        (C1, W1), (C2, W2) = G1.as_nurbs().coeffs_weights(), G2.as_nurbs().coeffs_weights()
        w1, w2, _ = _outer_shapes(W1, W2, G1.sdim, G2.sdim)
        W = w1 * w2
    else:
        C1, C2 = G1.coeffs, G2.coeffs
    c1, c2, S = _outer_shapes(C1, C2, G1.sdim, G2.sdim)
    c1 = np.broadcast_to(c1, S + c1.shape[-1:])
    c2 = np.broadcast_to(c2, S + c2.shape[-1:])
    C = np.concatenate((c2, c1), axis=-1)       # components in x,y order, axes in y,x order
    kvs = tuple(G1.kvs) + tuple(G2.kvs)
    return NurbsFunc(kvs, C, W) if rational else BSplineFunc(kvs, C)


def outer_sum(G1, G2):
    return _outer_op(G1, G2, np.add)


def outer_product(G1, G2):
    return _outer_op(G1, G2, np.multiply)


def _outer_op(G1, G2, op):
    kvs = tuple(G1.kvs) + tuple(G2.kvs)
    if G1._rational or G2._rational:
        (C1, W1), (C2, W2) = G1.as_nurbs().coeffs_weights(), G2.as_nurbs().coeffs_weights()
        c1, c2, _ = _outer_shapes(C1, C2, G1.sdim, G2.sdim)
        w1, w2, _ = _outer_shapes(W1, W2, G1.sdim, G2.sdim)
        return NurbsFunc(kvs, op(c1, c2), w1 * w2)
    c1, c2, _ = _outer_shapes(G1.coeffs, G2.coeffs, G1.sdim, G2.sdim)
    return BSplineFunc(kvs, op(c1, c2))


# ---------------------------------------------------------------------------
# stock geometries
# ---------------------------------------------------------------------------

def unit_cube(dim=3, num_intervals=1):
    return functools.reduce(tensor_product, dim * (line_segment(0.0, 1.0, intervals=num_intervals),))


def unit_square(num_intervals=1):
    return unit_cube(dim=2, num_intervals=num_intervals)


def identity(extents):
    extents = [ex.support() if hasattr(ex, 'kv') else ex for ex in extents]
    return functools.reduce(tensor_product,
                            (line_segment(ex[0], ex[1], support=ex) for ex in extents))


def _annulus_net(r1, r2, w_mid=None):
    # control points of a quarter annulus: rows = angular direction (degree 2), cols = radial
    net = np.zeros((3, 2, 2))
    net[0, :, 0] = (r1, r2)
    net[1, :, 0] = (r1, r2)
    net[1, :, 1] = (r1, r2)
    net[2, :, 1] = (r1, r2)
    if w_mid is None:
        return net
    W = np.ones((3, 2, 1))
    W[1] = w_mid
    return np.concatenate((net, W), axis=-1)


def bspline_quarter_annulus(r1=1.0, r2=2.0):
    """B-spline approximation of a quarter annulus (``pyiga/geometry.py:445-466``)."""
    kv_rad = bspline.make_knots(1, 0.0, 1.0, 1)
    kv_ang = bspline.make_knots(2, 0.0, 1.0, 1)
    return BSplineFunc((kv_ang, kv_rad), _annulus_net(r1, r2))


def quarter_annulus(r1=1.0, r2=2.0):
    """Exact NURBS quarter annulus (``pyiga/geometry.py:468-490``); the reference passes the
    control net with the weights as last component and lets the constructor premultiply."""
    kv_rad = bspline.make_knots(1, 0.0, 1.0, 1)
    kv_ang = bspline.make_knots(2, 0.0, 1.0, 1)
    return NurbsFunc((kv_ang, kv_rad), _annulus_net(r1, r2, 1.0 / np.sqrt(2.0)), weights=None)


# control net of gismo's twistedFlatQuarterAnnulus.xml, as (x, y, z) triples in the storage
# order of a (2, 4, 2) net (``pyiga/geometry.py:557-589``)
_TWISTED_BOX_NET = (
    (1, 0, 0), (2, 0, 0), (1, .5, 0), (2, 1.5, 0), (.5, 1, .5), (1.5, 2, .5), (0, 1, 2), (0, 2, 2),
    (1, 0, 1), (2, 0, 1), (1, .5, 1), (2, 1.5, 1), (1, 1, 1.5), (1.5, 2, 1.5), (1, 1, 2), (1, 2, 2),
)


def twisted_box():
    """3D box with a twisted, bent right face; degrees (1, 3, 1), one span per axis."""
    k1 = bspline.make_knots(1, 0.0, 1.0, 1)
    k3 = bspline.make_knots(3, 0.0, 1.0, 1)
    net = np.array(_TWISTED_BOX_NET, dtype=float).reshape(2, 4, 2, 3)
    return BSplineFunc((k1, k3, k1), net)


def twisted_nurbs_box():
    """The rational variant of :func:`twisted_box` used as benchmark geometry:
    weights ``1 + 0.25*((i + 2j + 3k) % 3)`` on the (2, 4, 2) net."""
    G = twisted_box()
    i, j, k = np.meshgrid(np.arange(2), np.arange(4), np.arange(2), indexing='ij')
    W = 1.0 + 0.25 * ((i + 2 * j + 3 * k) % 3)
    return NurbsFunc(G.kvs, G.coeffs.copy(), W)
