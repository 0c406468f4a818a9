"""Evaluation of user functions on tensor grids (the behaviour of ``pyiga/utils.py:9-52``).

Python callables cannot run on the device; like the reference they are evaluated with numpy, on
grids whose geometry transform comes from the GPU (``grid_eval`` of the spline function objects).
Axes are ordered z, y, x; user functions take their arguments in x, y, z order.
"""
import numpy as np


def _on_grid(values, shape):
    """Bring the result of a user function to an array whose leading axes are the grid: tuples are
    vector components (stacked last), scalars and partially broadcast results are expanded."""
    nd = len(shape)

    def expand(v):
        v = np.asanyarray(v)
        if v.ndim == 0 or v.shape[:nd] != shape:
            tail = v.shape[nd:] if v.ndim > nd else ()
            v = np.broadcast_to(v, shape + tail)
        return v

    if isinstance(values, tuple):
        return np.stack([expand(c) for c in values], axis=-1)
    return expand(values)


def _ensure_grid_shape(values, grid):
    return _on_grid(values, tuple(len(g) for g in grid))


def grid_eval(f, grid):
    """`f` on the tensor grid `grid`; objects with their own ``grid_eval`` (spline functions) are
    asked directly."""
    own = getattr(f, 'grid_eval', None)
    if own is not None:
        return own(grid)
    open_axes = np.ix_(*[np.asarray(g, dtype=float) for g in grid])     # broadcastable coordinate arrays, z..x
    return _on_grid(f(*open_axes[::-1]), tuple(len(g) for g in grid))


def grid_eval_transformed(f, grid, geo):
    """`f`, given in physical coordinates, on the image of the tensor grid under `geo`."""
    pts = grid_eval(geo, grid)
    coords = [pts[..., k] for k in range(pts.shape[-1])]
    return _on_grid(f(*coords), tuple(len(g) for g in grid))


def multi_kron_sparse(As, format='csr'):
    """Kronecker product of a sequence of sparse matrices (``pyiga/utils.py:62-67``)"""
    import functools
    import scipy.sparse
    As = list(As)
    if len(As) == 1:
        return As[0].asformat(format, copy=True)
    return functools.reduce(lambda A, B: scipy.sparse.kron(A, B, format='csr'), As).asformat(format)
