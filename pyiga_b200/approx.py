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


def project_L2(kvs, f, f_physical=False, geo=None, rtol=1e-14, maxiter=500):
    """Coefficients of the L2 projection of `f` onto the tensor-product spline space, optionally on
    the physical domain `geo` (``pyiga/approx.py:62-95``).  The mass matrix stays on the device as a
    multi-level banded tensor; ``M x = rhs`` is solved by conjugate gradients with the device matvec
    and the Kronecker product of the inverse 1D mass matrices as preconditioner (the reference
    factorises M with a sparse direct solver; `rtol` is the relative residual CG stops at)."""
    import torch
    from . import assemble
    from .dist import GatheredKronecker, SlabAssembly, SlabOperator, cg
    from .operators import KroneckerOperator
    if isinstance(kvs, bspline.KnotVector):
        kvs = (kvs,)
    kvs = tuple(kvs)
    if len(kvs) == 1:
        M = assemble.bsp_mass_1d(kvs[0])
        rhs = assemble.inner_products(kvs[0], f)
        return scipy.sparse.linalg.spsolve(M.tocsc(), rhs)
    if f_physical:
        assert geo is not None, 'projection in physical coordinates requires a geometry'
    rhs = np.asarray(assemble.inner_products(kvs, f, f_physical=f_physical, geo=geo), dtype=np.float64)
    extra = rhs.shape[len(kvs):]
    from . import geometry
    sa = SlabAssembly(kvs, geo if geo is not None else geometry.identity(kvs), 'mass')
    be = sa.dev.be
    op = SlabOperator(sa.dev, sa.assemble_mlb(), rows=sa.rows, slabs=sa.slabs, rank=0)
    plane = int(np.prod([kv.numdofs for kv in kvs[1:]], dtype=np.int64))
    Minv = [np.linalg.inv(assemble.bsp_mass_1d(kv).toarray()) for kv in kvs]
    prec = GatheredKronecker(KroneckerOperator(*Minv), sa.slabs, 0, plane)
    to_t = (lambda v: torch.from_numpy(np.ascontiguousarray(v))) if be.name == 'emu' else be.from_host

    native = None
    if len(kvs) == 3:
        # device-resident CG (csrc/distcg.cuh): scalars, convergence flag and iteration batches stay on the GPU
        from .distcg import DistributedCG
        native = DistributedCG(sa.dev.device_structure, op.mlb, sa.slabs, 0, Minv)

    def solve(b):
        if native is not None:
            x, it, res = native.solve(be.from_host(np.ascontiguousarray(b.ravel())), rtol=rtol, maxiter=maxiter, check_every=10)
            if res > rtol:
                print('WARNING: project_L2 - CG did not converge (relative residual %.2e after %d iterations)' % (res, it))
            return be.to_host(x).reshape(b.shape)
        x, it, hist = cg(lambda v: op._t(op.matvec(v)).clone(), to_t(b.ravel()), M=prec, rtol=rtol, maxiter=maxiter)
        if hist and hist[-1] > rtol:        # the reference prints a warning as well (pyiga/approx.py:94-95)
            print('WARNING: project_L2 - CG did not converge (relative residual %.2e after %d iterations)' % (hist[-1], it))
        return (np.asarray(x.cpu()) if hasattr(x, 'cpu') else np.asarray(x)).reshape(b.shape)
    if extra == ():
        return solve(rhs)
    out = np.empty_like(rhs)
    for idx in np.ndindex(*extra):
        out[(Ellipsis,) + idx] = solve(np.ascontiguousarray(rhs[(Ellipsis,) + idx]))
    return out
