"""Evaluation of user functions on tensor grids (``pyiga/utils.py:9-52``).  Python callables cannot
run on the device; like the reference they are evaluated with numpy on grids whose geometry
transform comes from the GPU (``grid_eval`` of the spline functions)."""
import numpy as np


def _broadcast_to_grid(X, grid_shape):
    num_dims = len(grid_shape)
    X = np.asanyarray(X)
    if X.ndim == 0:
        X = np.broadcast_to(X, grid_shape)
    # input might be a higher-dimensional tensor: only the leading grid axes are broadcast
    if X.shape[:num_dims] != tuple(grid_shape):
        X = np.broadcast_to(X, tuple(grid_shape) + X.shape[num_dims:])
    return X


def _ensure_grid_shape(values, grid):
    grid_shape = tuple(len(g) for g in grid)
    if isinstance(values, tuple):       # vector-valued function given as a tuple of components
        values = np.stack(tuple(_broadcast_to_grid(v, grid_shape) for v in values), axis=-1)
    return _broadcast_to_grid(values, grid_shape)


def grid_eval(f, grid):
    """Evaluate `f` over the tensor grid `grid` (axes in z,y,x order; `f` takes x,y,z)."""
    if hasattr(f, 'grid_eval'):
        return f.grid_eval(grid)
    mesh = list(np.meshgrid(*grid, sparse=True, indexing='ij'))
    mesh.reverse()
    return _ensure_grid_shape(f(*mesh), grid)


def grid_eval_transformed(f, grid, geo):
    """Evaluate `f`, given in physical coordinates, on the image of the grid under `geo`."""
    trf = grid_eval(geo, grid)
    X = tuple(trf[..., i] for i in range(trf.shape[-1]))
    return _ensure_grid_shape(f(*X), grid)
