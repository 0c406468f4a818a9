"""Shared helpers of the parity tests."""
import os

import numpy as np
import scipy.sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

CASES = ['a2_qa', 'a2_mixed', 'a3_tb', 'a3_mixed', 'a3_nurbs', 'a3_mult', 'a3_p1', 'a3_p4']
GEOS = ['tb', 'tnb', 'qa', 'bqa', 'cyl']
RTOL = 1e-12        # north-star tolerance: relative to the max-abs entry of the reference matrix


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, 'pyiga_%s.npz' % name))
    return scipy.sparse.csr_matrix((z['data'], z['indices'], z['indptr']), shape=tuple(z['shape']))


def geo_arrays(ref, name):
    sdim = int(ref['geo_%s_sdim' % name])
    kvs = [ref['geo_%s_kv%d' % (name, k)] for k in range(sdim)]
    ps = [int(ref['geo_%s_p%d' % (name, k)]) for k in range(sdim)]
    return kvs, ps, ref['geo_%s_coeffs' % name], bool(ref['geo_%s_rational' % name])


def make_geo(ref, name):
    """pyiga_b200 geometry object from the packed fixture arrays."""
    from pyiga_b200 import bspline, geometry
    kvs, ps, coeffs, rational = geo_arrays(ref, name)
    kvs = tuple(bspline.KnotVector(kv, p) for kv, p in zip(kvs, ps))
    if rational:
        return geometry.NurbsFunc(kvs, coeffs.copy(), None, premultiplied=True)
    return bspline.BSplineFunc(kvs, coeffs.copy())


def case_space(ref, case):
    dim = int(ref[case + '_dim'])
    kvs = [ref['%s_kv%d' % (case, k)] for k in range(dim)]
    ps = [int(ref['%s_p%d' % (case, k)]) for k in range(dim)]
    return kvs, ps


def make_space(ref, case):
    from pyiga_b200 import bspline
    kvs, ps = case_space(ref, case)
    return tuple(bspline.KnotVector(kv, p) for kv, p in zip(kvs, ps))


def ref_csr(ref, case, form):
    mlb = ref['%s_%s_mlb' % (case, form)]
    indptr, indices = ref['%s_%s_indptr' % (case, form)], ref['%s_%s_indices' % (case, form)]
    n = len(indptr) - 1
    # CSR values in canonical order = MLB values permuted; rebuild through the index lists
    return indptr, indices, mlb


def assert_close_rel(got, want, rtol=RTOL, what=''):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, '%s: shape %s != %s' % (what, got.shape, want.shape)
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= rtol * scale, '%s: max abs error %.3e > %.1e * max|ref| (%.3e)' % (what, err, rtol, scale)


# variational forms used by the vform parity tests: name -> (form string, inputs, space case, geometry)
A3 = [[2.0, 0.3, 0.0], [-0.1, 1.5, 0.2], [0.4, 0.0, 1.0]]
VFORMS = {
    'cd3': ('(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx',
            {'diff_coeff': lambda x, y, z: 1.0 + x * y}, 'a3_mixed', 'tb'),
    'cd2': ('(inner(grad(u), grad(v)) + inner(b, grad(u)) * v + 3.0 * u * v) * dx',
            {'b': lambda x, y: (y, -x)}, 'a2_qa', 'qa'),
    'aniso3': ('inner(dot(A, grad(u)), grad(v)) * dx', {'A': A3}, 'a3_tb', 'tnb'),
    'react2': ('(c * u * v + inner(grad(v), w) * u) * dx',
               {'c': lambda x, y: 2.0 + x - y * y, 'w': (0.5, -1.25)}, 'a2_mixed', 'bqa'),
    'sqrt3': ('sqrt(kappa) * inner(grad(u), grad(v)) / (1.0 + x[2]) * dx',
              {'kappa': lambda x, y, z: 1.0 + x * x + z}, 'a3_p1', 'tb'),
}

# linear forms (arity 1): name -> (form string, inputs, space case, geometry)
LFORMS = {
    'lin3': ('(f * v + inner((1.0, x[0], -2.0), grad(v))) * dx', {'f': lambda x, y, z: x * y + z * z}, 'a3_mixed', 'tb'),
    'lin2': ('g * v * dx', {'g': lambda x, y: x * y + 1.0}, 'a2_qa', 'qa'),
    'lin3n': ('inner(b, grad(v)) * dx', {'b': lambda x, y, z: (y, 1.0 + x, z * x)}, 'a3_nurbs', 'tnb'),
}

# vector-valued forms: name -> (form, bfuns, inputs, space case, geometry)
VECFORMS = {
    'nonsym2': ('inner(as_matrix([[2,1],[0,0]]).dot(u), v) * dx', [('u', 2), ('v', 2)], {}, 'a2_qa', 'qa'),
    'elast3': ('(2.0*inner(0.5*(grad(u)+grad(u).T), grad(v)) + lam*div(u)*div(v)) * dx', [('u', 3), ('v', 3)],
               {'lam': 1.5}, 'a3_tb', 'tb'),
    'stokes_like2': ('(inner(grad(u), grad(v)) + c * inner(u, v)) * dx', [('u', 2), ('v', 2)],
                     {'c': lambda x, y: 1.0 + x * y}, 'a2_mixed', 'bqa'),
}

# boundary integrals: name -> (form, bfuns or None, inputs, space case, geometry, bdspec)
BFORMS = {
    'robin2_left': ('u * v * ds', None, {}, 'a2_qa', 'qa', 'left'),
    'robin2_top': ('c * u * v * ds', None, {'c': lambda x, y: 1.0 + x * y}, 'a2_mixed', 'bqa', 'top'),
    'neumann2': ('g * v * ds', None, {'g': lambda x, y: x + y}, 'a2_qa', 'qa', (0, 1)),
    'flux2': ('inner(grad(u), n) * v * ds', None, {}, 'a2_qa', 'qa', 'right'),
    'flux3_front': ('inner(grad(u), n) * v * ds', None, {}, 'a3_mixed', 'tb', 'front'),
    'nitsche3': ('(inner(grad(u), n) * v + inner(grad(v), n) * u + 3.0 * u * v) * ds', None, {}, 'a3_nurbs', 'tnb', (1, 1)),
    'neumann3': ('inner(g, n) * v * ds', None, {'g': lambda x, y, z: (x, y * z, 1.0)}, 'a3_tb', 'tnb', 'left'),
    'traction2': ('inner(g, v) * ds', [('v', 2)], {'g': lambda x, y: (x, -y)}, 'a2_qa', 'qa', 'bottom'),
}

# Petrov-Galerkin forms: trial space = degrees `p0`, test space = degrees `p1` on the mesh of a space case
# name -> (form, bfuns, inputs, dim, trial degrees, test degrees, spans, geometry)
PGFORMS = {
    'pg2': ('inner(grad(u), grad(v)) * dx', [('u', 1, 0), ('v', 1, 1)], {}, (2, 2), (3, 3), (3, 4), 'qa'),
    'pg3': ('(c * u * v + inner(b, grad(u)) * v) * dx', [('u', 1, 0), ('v', 1, 1)],
            {'c': lambda x, y, z: 1.0 + x * z, 'b': lambda x, y, z: (y, -x, 1.0)}, (3, 2, 2), (2, 3, 3), (3, 2, 4), 'tnb'),
}

# surface integrals over a manifold (geo: R^2 -> R^3; the reference supports them without derivatives
# of the basis functions only, and probes input callables with 2 arguments before calling them with 3): name -> (form, bfuns, inputs, (degrees), (spans), side of the 3D cylinder)
SFORMS = {
    'area': ('v * ds', None, {}, (3, 2), (4, 5), 'right'),
    'smass': ('u * v * ds', None, {}, (2, 3), (3, 4), 'left'),
    'sflux': ('inner(g, n) * v * ds', None, {'g': lambda *X: (X[0], X[1] * X[-1], 1.0 + X[0])}, (2, 2), (3, 3), 'back'),
    'sreact': ('c * u * v * ds', None, {'c': lambda *X: 1.0 + X[0] * X[1] + X[-1]}, (2, 2), (4, 4), 'left'),
}
