"""Poisson problem end to end on the device path (the flow of the reference's notebooks and of
test/test_solve.py, in 3D): stiffness matrix and load vector by sum factorisation, Dirichlet data by
interpolation on the faces, elimination on device CSR arrays, CG with the restricted device matrix.

    python tools/poisson_demo.py [p] [n]

Prints one JSON line with the sizes, the phase times and the error against the interpolant of the
exact solution."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse.linalg


def solve(p, n, geo=None, cg_rtol=1e-10):
    from pyiga_b200 import approx, assemble, bspline, geometry
    geo = geometry.twisted_nurbs_box() if geo is None else geo
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)

    def g(x, y, z):         # exact solution and Dirichlet data
        return np.cos(x + y) * np.exp(0.5 * z)

    def f(x, y, z):         # -Laplace g = (2 - 1/4) g
        return 1.75 * g(x, y, z)

    t = [time.perf_counter()]
    K = assemble.stiffness(kvs, geo, format='mlb')              # values stay on the device
    rhs = assemble.inner_products(kvs, f, f_physical=True, geo=geo).ravel()
    t.append(time.perf_counter())
    bcs = assemble.compute_dirichlet_bcs(kvs, geo, ('all', g))
    LS = assemble.RestrictedLinearSystem(K, rhs, bcs)           # device CSR, compacted on the device
    t.append(time.perf_counter())
    A = scipy.sparse.linalg.LinearOperator(LS.A_device.shape, matvec=LS.A_device.dot, dtype=np.float64)
    diag = 1.0 / LS.A_device.to_scipy().diagonal()
    its = [0]
    u, info = scipy.sparse.linalg.cg(A, LS.b, rtol=cg_rtol, atol=0.0, maxiter=5000,
                                     M=scipy.sparse.linalg.LinearOperator(A.shape, matvec=lambda r: diag * r),
                                     callback=lambda xk: its.__setitem__(0, its[0] + 1))
    t.append(time.perf_counter())
    u = LS.complete(u)
    u_ex = approx.interpolate(kvs, g, geo=geo).ravel()
    err = float(np.sqrt(np.mean((u - u_ex) ** 2)))
    return {'p': p, 'n': n, 'ndofs': int(u.size), 'nnz': int(K.nnz), 'free_dofs': int(LS.b.size),
            'assemble_s': t[1] - t[0], 'bcs_restrict_s': t[2] - t[1], 'cg_s': t[3] - t[2], 'cg_iterations': its[0],
            'cg_info': int(info), 'rms_error_vs_interpolant': err}


if __name__ == '__main__':
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    print(json.dumps(solve(p, n)))
