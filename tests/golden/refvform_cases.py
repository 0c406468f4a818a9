"""The reference-VForm cases shared by the fixture generator (make_golden_vform.py, runs the real
reference incl. its JIT) and the parity check (builds the same pyiga.vform.VForm objects and hands them
to the device backend).  Needs `pyiga` importable (oracle/_ref)."""
import numpy as np


def spaces():
    from pyiga import bspline as rbs
    kv2 = (rbs.make_knots(2, 0.0, 1.0, 4), rbs.make_knots(3, 0.0, 1.0, 3))
    kv3 = (rbs.make_knots(2, 0.0, 1.0, 3), rbs.make_knots(1, 0.0, 1.0, 4), rbs.make_knots(2, 0.0, 1.0, 2))
    return kv2, kv3


def _A(x, y):
    one = 1.0 + 0.0 * (x + y)
    return np.stack([np.stack([(2.0 + x) * one, 0.3 * y * one], -1), np.stack([0.3 * y * one, (1.0 + x * x) * one], -1)], -2)


def cases():
    """name -> (VForm factory, knot vectors, geometry, inputs)"""
    from pyiga import geometry as rgeo, vform as rvf
    kv2, kv3 = spaces()
    geo2, geo3 = rgeo.quarter_annulus(), rgeo.twisted_box()
    dc = lambda x, y, z: 1.0 + x * y

    def aniso2():
        vf = rvf.VForm(2)
        u, v = vf.basisfuns()
        A = vf.input('A', shape=(2, 2))
        c = vf.parameter('c')
        vf.add(rvf.inner(rvf.dot(A, rvf.grad(u)), rvf.grad(v)) * rvf.dx + c * u * v * rvf.dx)
        return vf

    def surf_mass2():
        vf = rvf.VForm(2, geo_dim=3)
        u, v = vf.basisfuns()
        vf.add(u * v * rvf.ds)
        return vf

    def surf_flux2():
        vf = rvf.VForm(2, geo_dim=3, arity=1)
        v = vf.basisfuns()
        vf.add(rvf.inner(vf.normal, (1.0, 2.0, 3.0)) * v * rvf.ds)
        return vf

    def biharmonic2():
        vf = rvf.VForm(2)
        u, v = vf.basisfuns()
        vf.add(rvf.inner(rvf.hess(u), rvf.hess(v)) * rvf.dx)
        return vf

    return {
        'stiffness2': (lambda: rvf.stiffness_vf(2), kv2, geo2, {}),
        'mass3': (lambda: rvf.mass_vf(3), kv3, geo3, {}),
        'convdiff3': (lambda: rvf.parse_vf('(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx',
                                           kv3, args={'diff_coeff': dc}), kv3, geo3, {'diff_coeff': dc}),
        'aniso2': (aniso2, kv2, geo2, {'A': _A, 'c': 2.5}),
        'divdiv2': (lambda: rvf.divdiv_vf(2), kv2, geo2, {}),
        'l2func2': (lambda: rvf.L2functional_vf(2, physical=True), kv2, geo2, {'f': lambda x, y: x * y + 1.0}),
        # space-time heat form (pyiga/vform.py:1759-1763; the last coordinate is time)
        'heat_st2': (lambda: rvf.heat_st_vf(2), kv2, rgeo.unit_square(), {}),
        'heat_st3': (lambda: rvf.heat_st_vf(3), kv3, geo3, {}),
        # surface integrals over a face of the twisted box (geo: R^2 -> R^3): surface measure and unit normal
        'surf_mass2': (surf_mass2, kv2, geo3.boundary('left'), {}),
        'surf_flux2': (surf_flux2, kv2, geo3.boundary('left'), {}),
        # space-time wave form: second time derivative and mixed space-time derivatives (pyiga/vform.py:1766-1772)
        'wave_st2': (lambda: rvf.wave_st_vf(2), kv2, geo2, {}),
        'wave_st3': (lambda: rvf.wave_st_vf(3), kv3, geo3, {}),
        # fourth-order form: all second derivatives and the Hessian of the geometry map (per-entry path)
        'biharmonic2': (biharmonic2, kv2, geo2, {}),
    }


def _c2(x, y):
    return 1.0 + x * y


def _c3(x, y, z):
    return 1.0 + x * y + z


def pgcases():
    """Petrov-Galerkin forms (trial space 0, test space 1 on the same mesh): name -> (factory, (kvs0, kvs1), geo, inputs)"""
    from pyiga import bspline as rbs, geometry as rgeo, vform as rvf

    def pg(d):
        vf = rvf.VForm(d)
        u, v = vf.basisfuns(spaces=(0, 1))
        c = vf.input('c')
        vf.add((rvf.inner(rvf.grad(u), rvf.grad(v)) + c * u * v) * rvf.dx)
        return vf

    kva = (rbs.make_knots(2, 0.0, 1.0, 4), rbs.make_knots(2, 0.0, 1.0, 3))
    kvb = (rbs.make_knots(3, 0.0, 1.0, 4), rbs.make_knots(1, 0.0, 1.0, 3))
    kva3 = (rbs.make_knots(2, 0.0, 1.0, 3), rbs.make_knots(1, 0.0, 1.0, 2), rbs.make_knots(2, 0.0, 1.0, 2))
    kvb3 = (rbs.make_knots(1, 0.0, 1.0, 3), rbs.make_knots(2, 0.0, 1.0, 2), rbs.make_knots(3, 0.0, 1.0, 2))
    return {
        'pg2': (lambda: pg(2), (kva, kvb), rgeo.quarter_annulus(), {'c': _c2}),
        'pg3': (lambda: pg(3), (kva3, kvb3), rgeo.twisted_box(), {'c': _c3}),
    }


def _g2(x, y):
    return x + 2.0 * y


def bcases():
    """integrals over a side of the patch: name -> (VForm factory, knot vectors, geometry, inputs, sides)"""
    from pyiga import geometry as rgeo, vform as rvf
    kv2, kv3 = spaces()

    def robin(d):
        vf = rvf.VForm(d, boundary=True)
        u, v = vf.basisfuns()
        vf.add(u * v * rvf.ds)
        return vf

    def neumann(d):
        vf = rvf.VForm(d, arity=1, boundary=True)
        v = vf.basisfuns()
        g = vf.input('g')
        vf.add(g * v * rvf.ds)
        return vf

    def nitsche(d):       # normal flux of u against v: normal vector and gradients at the boundary
        vf = rvf.VForm(d, boundary=True)
        u, v = vf.basisfuns()
        vf.add(rvf.inner(rvf.grad(u), vf.normal) * v * rvf.ds)
        return vf

    geo2, geo3 = rgeo.quarter_annulus(), rgeo.twisted_box()
    return {
        'robin2': (lambda: robin(2), kv2, geo2, {}, ('left', 'right', 'top', 'bottom')),
        'robin3': (lambda: robin(3), kv3, geo3, {}, ('left', 'back', 'top')),
        'neumann2': (lambda: neumann(2), kv2, geo2, {'g': _g2}, ('right', 'bottom')),
        'nitsche2': (lambda: nitsche(2), kv2, geo2, {}, ('left', 'top')),
        'nitsche3': (lambda: nitsche(3), kv3, geo3, {}, ('front', 'bottom')),
    }


def hspace(truncate):
    from pyiga import bspline as rbs, hierarchical
    # the example space of the reference's own tests (test/test_hierarchical.py:10-18)
    kvs = 2 * (rbs.make_knots(3, 0.0, 1.0, 4),)
    hs = hierarchical.HSpace(kvs, truncate=truncate, disparity=1, bdspecs=[(0, 0), (0, 1), (1, 0), (1, 1)])
    for lv in range(2):
        hs.refine_region(lv, lambda *X: min(X) > 1 - 0.5 ** (lv + 1))
    return hs


def hcases():
    """name -> (VForm factory, inputs, symmetric)"""
    from pyiga import vform as rvf
    c = lambda x, y: 1.0 + x * y
    kvs0 = hspace(False).knotvectors(0)
    return {
        'hstiff': (lambda: rvf.stiffness_vf(2), {}, True),
        'hconv': (lambda: rvf.parse_vf('(inner(grad(u), grad(v)) + inner((1.0, x[0]), grad(u)) * v + c * u * v) * dx', kvs0,
                                       args={'c': c}), {'c': c}, False),
    }


def cases1d():
    """forms over ONE knot vector (test/test_assemble.py:436-445): name -> (string or VForm factory, kvs, geo, inputs).
    Strings go through the string front end of the package under test, factories build reference VForm objects."""
    from pyiga import bspline as rbs, geometry as rgeo, vform as rvf
    kv = rbs.make_knots(3, 0.0, 1.0, 7)
    kvm = rbs.make_knots(2, 0.0, 1.0, 5, mult=2)
    kvg = rbs.make_knots(2, 0.0, 1.0, 2)
    geo_id = rgeo.unit_cube(dim=1)
    geo_c = rbs.BSplineFunc((kvg,), np.array([[0.0], [0.2], [0.5], [1.3]]))
    geo_r = rgeo.NurbsFunc((kvg,), np.array([[0.5], [0.9], [1.2], [2.0]]), np.array([1.0, 0.8, 1.3, 1.0]))
    a = lambda x: 1.0 + x * x
    f = lambda x: 1.0 + x ** 2

    def mass_f():
        vf = rvf.VForm(1)
        u, v = vf.basisfuns()
        c = vf.input('c', physical=True)
        vf.add(c * u * v * rvf.dx)
        return vf

    return {
        's_stiff': ('inner(grad(u), grad(v)) * dx', (kv,), geo_id, {}),
        's_rhs': ('f * v * dx', (kv,), geo_id, {'f': f}),
        's_cd_curved': ('(a * inner(grad(u), grad(v)) + Dx(u, 0) * v + u * v) * dx', (kvm,), geo_c, {'a': a}),
        's_rhs_nurbs': ('f * v * dx', (kv,), geo_r, {'f': f}),
        'o_stiff': (lambda: rvf.stiffness_vf(1), (kvm,), geo_c, {}),
        'o_mass_c': (mass_f, (kv,), geo_r, {'c': a}),
        'o_l2': (lambda: rvf.L2functional_vf(1, physical=True), (kv,), geo_c, {'f': f}),
    }
