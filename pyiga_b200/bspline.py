"""Knot vectors and tensor-product spline functions (host-side descriptors).

Mirrors the part of the reference's ``pyiga.bspline`` that the assembly path
touches (``pyiga/bspline.py:36-213`` KnotVector / make_knots,
``pyiga/bspline.py:827-921`` BSplineFunc).  These classes only hold the
*description* of a space or a spline function; every evaluation on a Gauss
grid is done on the device through the C-ABI (``pyiga_b200._lib``) — there is
no host evaluation path.

Index convention (same as the reference): tensor axes are in z,y,x order, i.e.
``kvs[0]`` is the slowest axis and ``kvs[-1]`` is ``x``; coefficient arrays are
``coeffs[i_z, i_y, i_x, component]`` with components in x,y,z order.
"""
import numpy as np


class KnotVector:
    """Open knot vector plus spline degree (reference: ``pyiga/bspline.py:36-189``).

    Attributes:
        kv (ndarray): the knots, non-decreasing, first/last repeated ``p+1`` times
        p (int): spline degree
    """

    def __init__(self, knots, p):
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        if knots.ndim != 1:
            raise ValueError('knot vector must be one-dimensional')
        assert np.all(np.diff(knots) >= 0), 'knots should be increasing'
        self.kv = knots
        self.p = int(p)
        self._uniq = None       # (mesh, knot index -> mesh index)

    def __repr__(self):
        return 'KnotVector(%r, %r)' % (self.kv, self.p)

    def __str__(self):
        return '<KnotVector p=%d sz=%d>' % (self.p, self.kv.size)

    def __eq__(self, other):
        return (isinstance(other, KnotVector) or hasattr(other, 'kv')) \
            and self.p == other.p and len(self.kv) == len(other.kv) \
            and bool(np.allclose(self.kv, other.kv, atol=1e-8, rtol=1e-8))

    def __hash__(self):
        return hash((self.p, self.kv.size))

    # -- sizes ---------------------------------------------------------------
    @property
    def numknots(self):
        return self.kv.size

    @property
    def numdofs(self):
        """Number of B-splines over this knot vector."""
        return self.kv.size - self.p - 1

    @property
    def numspans(self):
        """Number of non-empty knot intervals."""
        return self.mesh.size - 1

    # -- mesh ----------------------------------------------------------------
    def _unique(self):
        if self._uniq is None:
            self._uniq = np.unique(self.kv, return_inverse=True)
        return self._uniq

    @property
    def mesh(self):
        """The distinct knots (break points)."""
        return self._unique()[0]

    def support(self, j=None):
        if j is None:
            return (self.kv[0], self.kv[-1])
        return (self.kv[j], self.kv[j + self.p + 1])

    def support_idx(self, j):
        return (j, j + self.p + 1)

    def mesh_support_idx(self, j):
        k2m = self._unique()[1]
        return (k2m[j], k2m[j + self.p + 1])

    def mesh_support_idx_all(self):
        """``(numdofs, 2)`` array: first / one-past-last mesh-span index of the support
        of every B-spline (reference: ``pyiga/bspline.py:129-136``)."""
        k2m = self._unique()[1]
        n = self.numdofs
        return np.column_stack((k2m[0:n], k2m[self.p + 1:self.p + 1 + n]))

    def mesh_span_indices(self):
        """Knot indices ``i`` with ``kv[i] != kv[i+1]`` (one per non-empty span)."""
        k2m = self._unique()[1]
        return np.nonzero(k2m[1:] != k2m[:-1])[0]

    def first_active(self, k):
        return k - self.p

    def findspan(self, u):
        """Knot-span index ``i`` with ``kv[i] <= u < kv[i+1]``, the right end point
        belonging to the last span (reference: ``pyiga/bspline_cy.pyx:13-27``)."""
        n = self.kv.size
        if u >= self.kv[n - self.p - 1]:
            return n - self.p - 2
        return int(np.searchsorted(self.kv, u, side='right')) - 1

    def first_active_at(self, u):
        return self.findspan(u) - self.p

    def greville(self):
        p = self.p
        if p == 0:
            return 0.5 * (self.kv[1:] + self.kv[:-1])
        g = np.array([self.kv[i + 1:i + p + 1].sum() / p for i in range(self.numdofs)])
        return np.clip(g, self.kv[0], self.kv[-1])

    def copy(self):
        return KnotVector(self.kv.copy(), self.p)

    def refine(self, new_knots=None):
        if new_knots is None:
            m = self.mesh
            new_knots = 0.5 * (m[1:] + m[:-1])
        return KnotVector(np.sort(np.concatenate((self.kv, new_knots))), self.p)

    def meshsize_avg(self):
        return abs(self.kv[-1] - self.kv[0]) / self.numspans


def _parse_bdspec(bdspec, dim):
    """``(axis, side)`` of a boundary given as pair or by name (``pyiga/bspline.py:13-33``)."""
    names = {'left': (dim - 1, 0), 'right': (dim - 1, 1), 'bottom': (dim - 2, 0), 'top': (dim - 2, 1),
             'front': (dim - 3, 0), 'back': (dim - 3, 1)}
    bd = names.get(bdspec, bdspec) if isinstance(bdspec, str) else bdspec
    if isinstance(bd, str) or not (len(bd) == 2 and bd[1] in (0, 1)):
        raise ValueError('invalid bdspec ' + str(bd))
    if bd[0] < 0 or bd[0] >= dim:
        raise ValueError('invalid bdspec %s for space of dimension %d' % (bdspec, dim))
    return tuple(bd)


def make_knots(p, a, b, n, mult=1):
    """Open knot vector of degree `p` on `(a,b)` with `n` equal spans and interior
    multiplicity `mult` (reference: ``pyiga/bspline.py:192-213``; the interior
    knots come from the same ``np.arange`` call so that the floats agree)."""
    interior = np.arange(a, b, (b - a) / n)[1:]
    knots = np.concatenate((np.full(p + 1, a, dtype=float),
                            np.repeat(interior, mult),
                            np.full(p + 1, b, dtype=float)))
    return KnotVector(knots, p)


def numdofs(kvs):
    if hasattr(kvs, 'numdofs'):
        return kvs.numdofs
    return int(np.prod([kv.numdofs for kv in kvs]))


def _as_kv_tuple(kvs):
    if hasattr(kvs, 'kv') and hasattr(kvs, 'p'):
        return (kvs,)
    return tuple(kvs)


# ---------------------------------------------------------------------------------------------
# evaluation of the active basis functions (K1 of the device library)
# ---------------------------------------------------------------------------------------------

def _basis_eval(kv, u, numderiv):
    """first active function per node (int32, n) and values [n][numderiv+1][p+1], from the GPU"""
    import ctypes as C
    from . import _device
    be = _device.backend()
    u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64).ravel()
    m = u.size
    d_kv, d_u = be.from_host(kv.kv), be.from_host(u)
    d_first = be.empty(m, np.int32)
    d_vals = be.empty(m * (numderiv + 1) * (kv.p + 1))
    _device.check(be.lib.pb200_basis_eval(be.ptr(d_kv), kv.kv.size, kv.p, be.ptr(d_u), m, numderiv,
                                          be.ptr(d_first), be.ptr(d_vals), be.stream()))
    return be.to_host(d_first), be.to_host(d_vals).reshape(m, numderiv + 1, kv.p + 1)


def active_deriv(kv, u, numderiv):
    """Values and derivatives up to `numderiv` of the p+1 active B-splines at the points `u`:
    array ``(numderiv+1, p+1, n)`` (or without the last axis for scalar `u`), computed by the K1
    kernel (``pyiga/bspline_cy.pyx:126-145``)."""
    _, vals = _basis_eval(kv, u, numderiv)
    out = np.ascontiguousarray(vals.transpose(1, 2, 0))
    return out[..., 0] if np.isscalar(u) else out


def active_ev(kv, u):
    """Values of the active B-splines: ``(p+1, n)`` (``pyiga/bspline_cy.pyx:116-124``)."""
    return active_deriv(kv, u, 0)[0]


def collocation_derivs_info(kv, nodes, derivs=1):
    """First active function per node and coefficient rows ``(derivs+1, n, p+1)``
    (``pyiga/bspline.py:649-660``)."""
    first, vals = _basis_eval(kv, nodes, derivs)
    return first.astype(np.int64), np.ascontiguousarray(vals.transpose(1, 0, 2))


def collocation_info(kv, nodes):
    first, vals = collocation_derivs_info(kv, nodes, 0)
    return first, vals[0]


def collocation_derivs(kv, nodes, derivs=1):
    """List of derivs+1 CSR collocation matrices ``(len(nodes), numdofs)`` (``pyiga/bspline.py:629-647``)."""
    import scipy.sparse
    nodes = np.asarray(nodes)
    m, n, p = nodes.size, kv.numdofs, kv.p
    first, vals = collocation_derivs_info(kv, nodes, derivs)
    I = np.repeat(np.arange(m), p + 1)
    J = (first[:, None] + np.arange(p + 1)[None, :]).ravel()
    return [scipy.sparse.coo_matrix((vals[d].ravel(), (I, J)), shape=(m, n)).tocsr() for d in range(derivs + 1)]


def collocation(kv, nodes):
    """CSR collocation matrix: entry (i, j) = B_j(nodes[i]) (``pyiga/bspline.py:591-612``)."""
    return collocation_derivs(kv, nodes, 0)[0]


class _SplineFuncBase:
    """Shared behaviour of :class:`BSplineFunc` and :class:`~pyiga_b200.geometry.NurbsFunc`."""

    _rational = False

    def is_scalar(self):
        return len(self.output_shape()) == 0

    def is_vector(self):
        return len(self.output_shape()) == 1

    @property
    def support(self):
        return tuple(kv.support() for kv in self.kvs)

    # device evaluation ------------------------------------------------------
    def _device_eval(self, gridaxes, want):
        from . import _device
        return _device.eval_spline_on_grid(self, gridaxes, want)

    def grid_eval(self, gridaxes):
        """Values on a tensor grid (axes in z,y,x order); evaluated on the GPU.

        Reference: ``pyiga/bspline.py:874-895`` / ``pyiga/geometry.py:102-114``."""
        assert len(gridaxes) == self.sdim, "Input has wrong dimension"
        return self._device_eval(gridaxes, 'value')

    def grid_jacobian(self, gridaxes):
        """Jacobians on a tensor grid, ``J[..., i, j] = d f_i / d xi_j`` with ``xi_0`` the
        *last* grid axis; evaluated on the GPU.

        Reference: ``pyiga/bspline.py:897-921`` / ``pyiga/geometry.py:116-123``."""
        assert len(gridaxes) == self.sdim, "Input has wrong dimension"
        return self._device_eval(gridaxes, 'jacobian')

    def grid_hessian(self, gridaxes):
        """Second derivatives on a tensor grid: per point (and per component of a vector function) the upper
        triangle of the Hessian in x, y, z order — (d_xx, d_xy, d_yy) / (d_xx, d_xy, d_xz, d_yy, d_yz, d_zz) —
        shape ``grid + [dim] + (n_hess,)`` (``pyiga/bspline.py:923-980``, ``pyiga/geometry.py:125-150``).
        Host evaluation from the device-tabulated 1D derivative matrices; only fourth-order forms need it."""
        assert len(gridaxes) == self.sdim, "Input has wrong dimension"
        d = self.sdim
        grid = [np.asarray(np.squeeze(g) if np.ndim(g) != 1 else g, dtype=float) for g in gridaxes]
        tabs = [collocation_derivs(kv, g, derivs=2) for kv, g in zip(self.kvs, grid)]
        C = np.asarray(self.coeffs, dtype=float)
        C = C.reshape(C.shape[:d] + (-1,))                  # components (homogeneous ones for NURBS) last

        def partial(orders):                                # orders[k]: derivative order along tensor axis k
            out = C
            for k in range(d):
                M = tabs[k][orders[k]]
                moved = np.moveaxis(out, k, 0)
                out = np.moveaxis((M @ moved.reshape(moved.shape[0], -1)).reshape((M.shape[0],) + moved.shape[1:]), 0, k)
            return out                                      # grid + (components,)

        def orders_of(*xyz):                                # derivative directions in x, y, z numbering
            o = [0] * d
            for a in xyz:
                o[d - 1 - a] += 1                           # x is the last tensor axis
            return o
        pairs = [(a, b) for a in range(d) for b in range(a, d)]
        if not self._rational:
            H = np.stack([partial(orders_of(a, b)) for a, b in pairs], axis=-1)     # grid + (comp, n_hess)
        else:
            # quotient rule on the homogeneous spline (V, W), N = V / W:
            #   N_ab = (V_ab - N_a W_b - N_b W_a - N W_ab) / W,   N_a = (V_a - N W_a) / W
            P0 = partial(orders_of())
            V, W = P0[..., :-1], P0[..., -1:]
            N = V / W
            first = [partial(orders_of(a)) for a in range(d)]
            Na = [(P[..., :-1] - N * P[..., -1:]) / W for P in first]
            Wa = [P[..., -1:] for P in first]
            cols = []
            for a, b in pairs:
                P2 = partial(orders_of(a, b))
                cols.append((P2[..., :-1] - Na[a] * Wa[b] - Na[b] * Wa[a] - N * P2[..., -1:]) / W)
            H = np.stack(cols, axis=-1)
        return H[..., 0, :] if self.is_scalar() else H

    def cylinderize(self, z0=0.0, z1=1.0, support=(0.0, 1.0)):
        """Patch with one more dimension: linear extrusion along a new last coordinate from `z0` to `z1`
        (``pyiga/bspline.py:1097-1106``); the space-time cylinders of the heat / wave assemblers."""
        from .geometry import line_segment, tensor_product
        return tensor_product(line_segment(z0, z1, support=support), self)

    def eval(self, *x):
        """Evaluate at a single point given in x,y,z order."""
        coords = tuple(np.atleast_1d(np.asarray(t, dtype=float)) for t in reversed(x))
        scalar_axes = tuple(i for i, t in enumerate(reversed(x)) if np.isscalar(t))
        y = self.grid_eval(coords).squeeze(axis=scalar_axes)
        return y.item() if y.shape == () else y

    __call__ = eval


class BSplineFunc(_SplineFuncBase):
    """Tensor-product B-spline function given by knot vectors and coefficients
    (reference: ``pyiga/bspline.py:827-871``)."""

    def __init__(self, kvs, coeffs):
        self.kvs = _as_kv_tuple(kvs)
        self.sdim = len(self.kvs)
        N = tuple(kv.numdofs for kv in self.kvs)
        coeffs = np.asanyarray(coeffs)
        if coeffs.ndim == 1:
            assert coeffs.shape[0] == np.prod(N), "Wrong length of coefficient vector"
            coeffs = coeffs.reshape(N)
        assert N == coeffs.shape[:self.sdim], "Wrong shape of coefficients"
        self.coeffs = coeffs
        tail = coeffs.shape[self.sdim:]
        if len(tail) == 0:
            self.dim = 1
        elif len(tail) == 1:
            self.dim = tail[0]
        else:
            self.dim = tail

    def output_shape(self):
        return self.coeffs.shape[self.sdim:]

    def copy(self):
        return BSplineFunc(self.kvs, self.coeffs.copy())

    def as_nurbs(self):
        from .geometry import NurbsFunc
        return NurbsFunc(self.kvs, self.coeffs.copy(), np.ones(self.coeffs.shape[:self.sdim]))

    def as_vector(self):
        if self.is_vector():
            return self
        assert self.is_scalar()
        return BSplineFunc(self.kvs, self.coeffs[..., None])

    def translate(self, offset):
        return BSplineFunc(self.kvs, self.coeffs + np.asarray(offset))

    def scale(self, factor):
        return BSplineFunc(self.kvs, self.coeffs * np.asarray(factor))

    def apply_matrix(self, A):
        assert self.is_vector(), 'Can only apply matrices to vector-valued functions'
        return BSplineFunc(self.kvs, np.matmul(np.asarray(A), self.coeffs[..., None])[..., 0])

    def rotate_2d(self, angle):
        assert self.dim == 2, 'Must be 2D vector function'
        c, s = np.cos(angle), np.sin(angle)
        return self.apply_matrix([[c, -s], [s, c]])

    def boundary(self, bdspec):
        """One side of the boundary as a :class:`BSplineFunc` with one parameter less
        (``pyiga/bspline.py:1017-1036``; open knot vectors make the boundary interpolatory)."""
        axis, side = _parse_bdspec(bdspec, self.sdim)
        slices = self.sdim * [slice(None)]
        slices[axis] = 0 if side == 0 else -1
        kvs = list(self.kvs)
        del kvs[axis]
        return BSplineFunc(kvs, self.coeffs[tuple(slices)])
