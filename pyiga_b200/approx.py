"""Spline interpolation (``pyiga/approx.py:14-51``): the collocation matrices come from the K1
kernel; the one-dimensional banded solves are host work on N x N matrices, exactly as in the
reference (``operators.make_solver`` -> SuperLU)."""
import numpy as np
import scipy.sparse.linalg

from . import bspline, utils


def _solve_along(C, X, axis):
    lu = scipy.sparse.linalg.splu(C.tocsc(), permc_spec='NATURAL')
    X = np.moveaxis(X, axis, 0)
    shp = X.shape
    Y = lu.solve(np.ascontiguousarray(X.reshape(shp[0], -1)))
    return np.moveaxis(Y.reshape(shp), 0, axis)


def interpolate(kvs, f, geo=None, nodes=None):
    """Coefficients of the interpolant of `f` in the tensor-product basis `kvs` at the Gréville
    abscissae (or `nodes`).  `f` lives on the parameter domain unless `geo` is given; an array of
    nodal values is accepted as well."""
    if isinstance(kvs, bspline.KnotVector):
        kvs = (kvs,)
    if nodes is None:
        nodes = [kv.greville() for kv in kvs]
    if isinstance(f, np.ndarray):
        if np.shape(f)[:len(kvs)] != tuple(kv.numdofs for kv in kvs):
            raise ValueError('array f has wrong shape')
        rhs = f
    elif geo is not None:
        rhs = utils.grid_eval_transformed(f, nodes, geo)
    else:
        rhs = utils.grid_eval(f, nodes)
    X = np.array(rhs, dtype=np.float64)
    for k, kv in enumerate(kvs):
        X = _solve_along(bspline.collocation(kv, nodes[k]), X, k)
    return X
