"""Parity checks of the pyiga_b200 path against reference fixtures, goldens and the oracle.

The same functions are run twice: through the sequential host emulation of the kernels
(tests/test_emulated.py, CPU) and through the CUDA library on the B200 (tests/test_gpu_parity.py).
Tolerance: the north-star's 1e-12 relative to the max-abs entry of the reference matrix; index
arrays bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import RTOL, assert_close_rel, case_space, geo_arrays, load_golden, make_geo, make_space
from oracle import pyiga_oracle as orc


def check_basis(be, ref):
    """K1 against bspline.collocation_derivs_info of the reference."""
    kv, p, nodes = ref['basis_kv'], int(ref['basis_p']), ref['basis_nodes']
    d_kv, d_nodes = be.from_host(kv), be.from_host(nodes)
    m = len(nodes)
    first = be.empty(m, np.int32)
    vals = be.empty(m * 3 * (p + 1))
    from pyiga_b200 import _device
    _device.check(be.lib.pb200_basis_eval(be.ptr(d_kv), len(kv), p, be.ptr(d_nodes), m, 2, be.ptr(first),
                                           be.ptr(vals), be.stream()))
    be.synchronize()
    assert np.array_equal(be.to_host(first), ref['basis_first'])
    got = be.to_host(vals).reshape(m, 3, p + 1).transpose(1, 0, 2)      # -> (deriv, node, fn)
    np.testing.assert_allclose(got, ref['basis_vals'], rtol=0, atol=1e-12 * np.abs(ref['basis_vals']).max())


def check_geometry(ref, name):
    """K2's geometry evaluation against grid_eval / grid_jacobian of the reference."""
    geo = make_geo(ref, name)
    grid = tuple(ref['geo_%s_grid%d' % (name, k)] for k in range(geo.sdim))
    np.testing.assert_allclose(geo.grid_eval(grid), ref['geo_%s_val' % name], rtol=0, atol=1e-13)
    np.testing.assert_allclose(geo.grid_jacobian(grid), ref['geo_%s_jac' % name], rtol=0, atol=1e-13)


def check_fields(ref, case, form):
    """K2's coefficient fields against the oracle's restatement of precompute_fields."""
    from pyiga_b200 import assemblers
    kvs = make_space(ref, case)
    geo = make_geo(ref, str(ref[case + '_geo']))
    cls = getattr(assemblers, ('Mass' if form == 'mass' else 'Stiffness') + 'Assembler%dD' % len(kvs))
    asm = cls(kvs, geo)
    okvs, ops = case_space(ref, case)
    gk, gp, gc, rational = geo_arrays(ref, str(ref[case + '_geo']))
    prob = orc.Problem(okvs, ops, gk, gp, gc, rational)
    want = (orc.fields_mass if form == 'mass' else orc.fields_stiffness)(prob.jac, prob.gw)
    got = np.moveaxis(asm.dev.fields_host(), 0, -1)
    assert_close_rel(got, want, what='fields %s %s' % (case, form))


def check_case(ref, case, entrywise=False, rows=None):
    """mass + stiffness of a fixture case: band structure, MLB values, CSR arrays."""
    from pyiga_b200 import assemblers
    from pyiga_b200.mlmatrix import MLStructure
    kvs = make_space(ref, case)
    dim = len(kvs)
    geo = make_geo(ref, str(ref[case + '_geo']))
    S = MLStructure.from_kvs(kvs, kvs)
    for k in range(dim):
        assert S.bidx[k].dtype == np.uint32
        assert np.array_equal(S.bidx[k], ref['%s_bidx%d' % (case, k)])
    for form, key in (('Mass', 'mass'), ('Stiffness', 'stiff')):
        asm = getattr(assemblers, '%sAssembler%dD' % (form, dim))(kvs, geo)
        for k in range(dim):    # the C library's structure must agree with the host one
            assert np.array_equal(asm.dev.bidx(k), S.bidx[k])
        want = ref['%s_%s_mlb' % (case, key)]
        data = asm.dev.be.to_host(asm.dev.assemble_mlb(entrywise=entrywise)).reshape(want.shape)
        assert_close_rel(data, want, what='%s %s mlb' % (case, key))
        if entrywise:
            continue
        assert asm.dev.fast_path, 'no sum-factorised instantiation for ' + case
        A = asm.assemble_csr()
        assert A.indptr.dtype == np.int32 and A.indices.dtype == np.int32 and A.data.dtype == np.float64
        assert np.array_equal(A.indptr, ref['%s_%s_indptr' % (case, key)])
        assert np.array_equal(A.indices, ref['%s_%s_indices' % (case, key)])
        B = orc.mlb_to_csr(want, [ref['%s_bidx%d' % (case, k)] for k in range(dim)], S.bs)
        B.sort_indices()
        assert_close_rel(A.data, B.data, what='%s %s csr' % (case, key))


def check_slabs(ref, case, nslabs):
    """row slabs of the first axis, assembled independently, concatenate to the full tensor
    (multi-GPU parity with logical ranks on one device)."""
    from pyiga_b200 import assemblers
    from pyiga_b200.dist import partition_rows
    kvs = make_space(ref, case)
    geo = make_geo(ref, str(ref[case + '_geo']))
    for form, key in (('Mass', 'mass'), ('Stiffness', 'stiff')):
        asm = getattr(assemblers, '%sAssembler%dD' % (form, len(kvs)))(kvs, geo)
        want = ref['%s_%s_mlb' % (case, key)]
        parts = []
        for (a, b) in partition_rows(asm.dev, nslabs):
            parts.append(asm.dev.be.to_host(asm.dev.assemble_mlb(rows=(a, b))))
        got = np.concatenate(parts).reshape(want.shape)
        assert_close_rel(got, want, what='%s %s in %d slabs' % (case, key, nslabs))


def check_chunked(ref, case):
    """a tiny workspace budget forces the chunked pipeline"""
    from pyiga_b200 import assemblers
    kvs = make_space(ref, case)
    geo = make_geo(ref, str(ref[case + '_geo']))
    asm = assemblers.StiffnessAssembler3D(kvs, geo)
    full = asm.dev.workspace_bytes()
    one = max(asm.dev.workspace_bytes((r, r + 1)) for r in range(asm.dev.ndofs_test[0]))
    assert one < full
    budget = one
    chunks = asm.dev.row_chunks(None, budget)
    assert len(chunks) > 1
    want = ref[case + '_stiff_mlb']
    got = asm.dev.be.to_host(asm.dev.assemble_mlb(budget_bytes=budget)).reshape(want.shape)
    assert_close_rel(got, want, what=case + ' chunked')


def check_multi_entries(ref, case):
    """the multi_entries protocol incl. pairs outside the pattern (-> exactly 0.0)"""
    from pyiga_b200 import assemblers
    kvs = make_space(ref, case)
    geo = make_geo(ref, str(ref[case + '_geo']))
    asm = getattr(assemblers, 'StiffnessAssembler%dD' % len(kvs))(kvs, geo)
    ij, want = ref[case + '_me_ij'], ref[case + '_me_val']
    got = asm.multi_entries(ij)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert_close_rel(got, want, what='multi_entries ' + case)
    assert np.all(got[want == 0.0] == 0.0)
    # iterable of pairs + scalar entry()
    got2 = asm.multi_entries((int(i), int(j)) for i, j in ij[:5])
    assert np.array_equal(got2, got[:5])
    assert asm.entry(int(ij[200, 0]), int(ij[200, 1])) == got[200]
    assert asm.arity == 2 and asm.kvs == (kvs, kvs)
    assert asm.inputs() == {'geo': (len(kvs),)} and asm.parameters() == {}


def check_golden(dim):
    """the reference's golden matrices (test/test_assemble.py:138-168) through assemble.mass/stiffness"""
    from pyiga_b200 import assemble, bspline, geometry
    if dim == 2:
        kvs = 2 * (bspline.make_knots(3, 0.0, 1.0, 15),)
        geo, names = geometry.bspline_quarter_annulus(), ('d2_p3_n15_mass', 'd2_p3_n15_stiff')
    else:
        kvs = 3 * (bspline.make_knots(2, 0.0, 1.0, 10),)
        geo, names = geometry.twisted_box(), ('d3_p2_n10_mass', 'd3_p2_n10_stiff')
    for fn, name in zip((assemble.mass, assemble.stiffness), names):
        A, G = fn(kvs, geo), load_golden(name)
        assert A.format == 'csr' and A.shape == G.shape
        scale = abs(G).max()
        assert abs(A - G).max() <= RTOL * scale, name
        assert abs(A - A.T).max() <= RTOL * scale


def check_operators(ref):
    from pyiga_b200.mlmatrix import MLMatrix, MLStructure
    from pyiga_b200.operators import KroneckerOperator
    kvs = make_space(ref, 'a3_mixed')
    S = MLStructure.from_kvs(kvs, kvs)
    M = MLMatrix(S, data=ref['mv_data'])
    np.testing.assert_allclose(M.dot(ref['mv_x']), ref['mv_y'], rtol=1e-13, atol=1e-13)
    A = M.asmatrix()
    np.testing.assert_allclose(A @ ref['mv_x'], ref['mv_y'], rtol=1e-13, atol=1e-13)
    facs = [ref['kron_A%d' % k] for k in range(3)]
    np.testing.assert_allclose(KroneckerOperator(*facs).dot(ref['kron_x']), ref['kron_y'], rtol=1e-13, atol=1e-13)


def check_vs_oracle(dim, ps, ns, form, geo_name='nurbs', mult=1, force_walk=False, rtol=RTOL, walk_split=None):
    """any space: the device pipeline against the oracle's closed form (the oracle is pinned to the
    reference by tests/test_oracle.py)"""
    from pyiga_b200 import assemblers, bspline, geometry
    kvs = tuple(bspline.make_knots(p, 0.0, 1.0, n, mult=mult) for p, n in zip(ps, ns))
    if dim == 2:
        geo = geometry.quarter_annulus() if geo_name == 'nurbs' else geometry.bspline_quarter_annulus()
    else:
        geo = geometry.twisted_nurbs_box() if geo_name == 'nurbs' else geometry.twisted_box()
    asm = getattr(assemblers, '%sAssembler%dD' % (form, dim))(kvs, geo)
    if force_walk:
        asm.dev.set_option('force_walk', 1)
    if walk_split is not None:
        asm.dev.set_option('walk_split', walk_split)
    got = asm.dev.be.to_host(asm.dev.assemble_mlb())
    prob = orc.Problem([kv.kv for kv in kvs], list(ps), [kv.kv for kv in geo.kvs], [kv.p for kv in geo.kvs],
                       geo.coeffs, geo._rational)
    want = orc.assemble_mlb(prob, form.lower()).ravel()
    assert_close_rel(got, want, rtol=rtol, what='%s %dD p=%s n=%s' % (form, dim, ps, ns))
    return asm


def check_vform(ref, name):
    """string vforms through assemble.assemble against the reference's JIT-compiled assemblers"""
    from helpers import VFORMS
    from pyiga_b200 import assemble
    form, inputs, case, gname = VFORMS[name]
    kvs = make_space(ref, case)
    geo = make_geo(ref, gname)
    want = ref['vf_%s_mlb' % name]
    M = assemble.assemble(form, kvs, geo=geo, format='mlb', **inputs)
    assert M.datashape == want.shape
    assert_close_rel(M.data, want, what='vform ' + name)
    A = assemble.assemble(form, kvs, args=dict(inputs, geo=geo))
    B = orc.mlb_to_csr(want, [ref['%s_bidx%d' % (case, k)] for k in range(len(kvs))], M.structure.bs)
    B.sort_indices()
    assert np.array_equal(A.indices, B.indices) and np.array_equal(A.indptr, B.indptr)
    assert_close_rel(A.data, B.data, what='vform csr ' + name)


def check_vform_protocol(ref):
    """Assembler with updatable inputs, error behaviour (test/test_assemble.py:409-450)"""
    import pytest
    from pyiga_b200 import assemble, vform
    kvs = make_space(ref, 'a2_qa')
    geo = make_geo(ref, 'qa')
    asm = assemble.Assembler('f * inner(grad(u), grad(v)) * dx', kvs, geo=geo, f=lambda x, y: 1.0 + x,
                             updatable=['f'])
    S = asm.asm.dev.structure
    I, J = (a.astype(np.int64) for a in S.nonzero())
    A = asm.assemble()
    assert_close_rel(np.asarray(A[I, J]).ravel(), ref['vf_upd_a'], what='Assembler')
    B = asm.assemble(f=lambda x, y: 2.0 + y * y)
    assert_close_rel(np.asarray(B[I, J]).ravel(), ref['vf_upd_b'], what='Assembler.update')
    with pytest.raises(RuntimeError):
        asm.update(geo=geo)             # not declared updatable
    with pytest.raises(ValueError):
        assemble.Assembler('f * u * v * dx', kvs, geo=geo, f=lambda x, y: x, updatable=['g'])
    with pytest.raises(ValueError, match="required input parameter 'f' missing"):
        assemble.assemble(vform.parse_vf('f * u * v * dx', kvs, args={'f': lambda x, y: x}), kvs, geo=geo)
    with pytest.raises(TypeError):
        assemble.assemble(42, kvs, geo=geo)
    # class-level metadata of the compiled form
    cls = vform.compile_vform(vform.parse_vf('c * u * v * dx', kvs, args={'c': 2.0}))
    assert cls.inputs() == {'geo': (2,)} and cls.parameters() == {'c': ()}
    # programmatic VForm equals the predefined stiffness assembler
    K = assemble.assemble(vform.stiffness_vf(2), kvs, geo=geo)
    K2 = assemble.stiffness(kvs, geo)
    assert abs(K - K2).max() <= RTOL * abs(K2).max()


def check_kronecker_path(ref):
    """geo=None: the reference takes the Kronecker shortcut over 1D matrices
    (pyiga/assemble.py:236-282, test/test_assemble.py:83-100); here the same matrices come from the
    device with the identity map, and the 1D factors from bsp_*_1d"""
    import scipy.sparse
    from pyiga_b200 import assemble, bspline
    kvs = (bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 5))
    M1 = [assemble.mass(kv) for kv in kvs]
    K1 = [assemble.stiffness(kv) for kv in kvs]
    # literal 1D matrices of the reference's tests (test/test_assemble.py:10-40) for p=1 on 3 spans
    kv = bspline.make_knots(1, 0.0, 1.0, 3)
    np.testing.assert_allclose(assemble.mass(kv).toarray() * 18,
                               [[2, 1, 0, 0], [1, 4, 1, 0], [0, 1, 4, 1], [0, 0, 1, 2]], atol=1e-13)
    np.testing.assert_allclose(assemble.stiffness(kv).toarray() / 3,
                               [[1, -1, 0, 0], [-1, 2, -1, 0], [0, -1, 2, -1], [0, 0, -1, 1]], atol=1e-13)
    M = assemble.mass(kvs)
    K = assemble.stiffness(kvs)
    Mk = scipy.sparse.kron(M1[0], M1[1])
    Kk = scipy.sparse.kron(K1[0], M1[1]) + scipy.sparse.kron(M1[0], K1[1])
    assert abs(M - Mk).max() <= 1e-13 * abs(Mk).max()
    assert abs(K - Kk).max() <= 1e-13 * abs(Kk).max()


def check_slab_operator_and_cg(ref, world=1, rank=0):
    """slab operator == MLMatrix matvec; CG with Kronecker preconditioner converges like scipy's"""
    import scipy.sparse.linalg
    import torch
    from pyiga_b200 import assemble, bspline, geometry
    from pyiga_b200.dist import GatheredKronecker, SlabAssembly, SlabOperator, cg
    from pyiga_b200.operators import KroneckerOperator
    kvs = (bspline.make_knots(2, 0.0, 1.0, 6), bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 5))
    geo = geometry.twisted_box()
    sa = SlabAssembly(kvs, geo, 'mass', rank=rank, world=world)
    be = sa.dev.be
    mlb = sa.assemble_mlb()
    op = SlabOperator(sa.dev, mlb, rows=sa.rows, slabs=sa.slabs, rank=rank)
    full = SlabAssembly(kvs, geo, 'mass')
    A = full.assemble_csr()
    rng = np.random.default_rng(7)
    x = rng.standard_normal(A.shape[1])
    a, b = sa.rows
    plane = kvs[1].numdofs * kvs[2].numdofs
    to_t = lambda v: torch.from_numpy(np.ascontiguousarray(v)) if be.name == 'emu' else be.from_host(v)
    y = op.matvec(be.from_host(x[a * plane:b * plane]))
    np.testing.assert_allclose(be.to_host(y), (A @ x)[a * plane:b * plane], rtol=1e-12, atol=1e-14)
    # CG as in pyiga/approx.py:82-93: Kronecker preconditioner from the inverses of the 1D mass matrices
    Minv = [np.linalg.inv(assemble.mass(kv).toarray()) for kv in kvs]
    prec = GatheredKronecker(KroneckerOperator(*Minv), sa.slabs, rank, plane)
    rhs = A @ np.ones(A.shape[1])
    bl = to_t(rhs[a * plane:b * plane])
    xs, it, hist = cg(lambda v: op._t(op.matvec(v)).clone(), bl, M=prec, rtol=1e-10, maxiter=100)
    xs = np.asarray(xs.cpu()) if hasattr(xs, 'cpu') else np.asarray(xs)
    np.testing.assert_allclose(xs, 1.0, rtol=0, atol=1e-7)
    it_ref = [0]
    scipy.sparse.linalg.cg(A, rhs, rtol=1e-10, atol=0.0, maxiter=100, M=scipy.sparse.linalg.LinearOperator(
        A.shape, matvec=lambda r: orc.kron_matvec(Minv, r)), callback=lambda xk: it_ref.__setitem__(0, it_ref[0] + 1))
    assert abs(it - it_ref[0]) <= 1, (it, it_ref[0])
    return it


def check_linear_forms(ref):
    """arity-1 forms: assemble_vector through assemble.assemble and inner_products"""
    from helpers import LFORMS
    from pyiga_b200 import assemble, assemblers
    for name, (form, inputs, case, gname) in LFORMS.items():
        kvs = make_space(ref, case)
        got = assemble.assemble(form, kvs, geo=make_geo(ref, gname), **inputs)
        want = ref['lf_%s' % name]
        assert got.shape == want.shape == tuple(kv.numdofs for kv in kvs)
        assert_close_rel(got, want, what='linear form ' + name)
    kvs = make_space(ref, 'a3_tb')
    fpar = lambda x, y, z: x + 2 * y * z
    assert_close_rel(assemble.inner_products(kvs, fpar, geo=make_geo(ref, 'tb')), ref['ip_param'], what='inner_products')
    assert_close_rel(assemble.inner_products(kvs, fpar, f_physical=True, geo=make_geo(ref, 'tnb')), ref['ip_phys'],
                     what='inner_products physical')
    assert_close_rel(assemble.inner_products(kvs, fpar), ref['ip_nogeo'], what='inner_products geo=None')
    asm = assemblers.L2FunctionalAssemblerPhys3D(kvs, make_geo(ref, 'tnb'), fpar)
    assert asm.arity == 1 and asm.entry(0, 0) == 0.0
    vec = asm.assemble_vector()
    assert np.array_equal(asm.multi_entries([0, 5, 7]), vec.ravel()[[0, 5, 7]])


def check_integrate(ref):
    """assemble.integrate (pyiga/assemble.py:658-696)"""
    from pyiga_b200 import assemble, bspline, geometry
    kvsI = (bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 5))
    qa = geometry.quarter_annulus()
    tol = 1e-12
    assert abs(assemble.integrate(kvsI, lambda x, y: 1.0, geo=qa) - ref['int_qa_one']) <= tol * abs(ref['int_qa_one'])
    assert abs(3 * np.pi / 4 - assemble.integrate(kvsI, lambda x, y: 1.0, geo=qa)) < 1e-8
    got = assemble.integrate(kvsI, lambda x, y: x * y + np.cos(x), f_physical=True, geo=qa)
    assert abs(got - ref['int_qa_phys']) <= tol * abs(ref['int_qa_phys'])
    assert abs(assemble.integrate(kvsI, lambda x, y: x * x + y) - ref['int_par']) <= tol * abs(ref['int_par'])
    assert_close_rel(assemble.integrate(kvsI, lambda x, y: (x, y * x)), ref['int_vec'], what='integrate vector')
    kvs3 = make_space(ref, 'a3_tb')
    got = assemble.integrate(kvs3, lambda x, y, z: x + y * z, f_physical=True, geo=make_geo(ref, 'tnb'))
    assert abs(got - ref['int_3d']) <= tol * abs(ref['int_3d'])
    got = assemble.integrate(bspline.make_knots(3, 0.0, 2.0, 5), lambda x: x * x)
    assert abs(got - ref['int_1d']) <= tol * abs(ref['int_1d'])


def check_initial_condition(ref):
    """space-time initial conditions (pyiga/assemble.py:492-552)"""
    from pyiga_b200 import assemble, bspline, geometry
    kvsT = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 3), bspline.make_knots(2, 0.0, 1.0, 5))
    geoT = geometry.tensor_product(geometry.line_segment(0.0, 1.0), geometry.quarter_annulus())
    for side in (0, 1):
        idx, val = assemble.compute_initial_condition_01(kvsT, geoT, (0, side), lambda x, y, t: x * y, lambda x, y, t: x - y)
        assert np.array_equal(idx, ref['ic01_idx%d' % side])
        assert_close_rel(val, ref['ic01_val%d' % side], what='initial condition, side %d' % side)


def check_project_L2(ref):
    """L2 projection: device mass matrix + preconditioned CG vs the reference's direct solve
    (pyiga/approx.py:62-95); the tolerance is the conditioning of the mass matrix times the CG residual"""
    from pyiga_b200 import approx, bspline, geometry
    kvsP = 2 * (bspline.make_knots(3, 0.0, 1.0, 10),)
    gP = lambda x, y: np.cos(x + y) + np.exp(y - x)
    got = approx.project_L2(kvsP, gP, f_physical=True, geo=geometry.quarter_annulus())
    assert_close_rel(got, ref['pl2_2d'], rtol=1e-9, what='project_L2 2D physical')
    kvs3 = make_space(ref, 'a3_tb')
    g3 = make_geo(ref, 'tnb')
    got = approx.project_L2(kvs3, lambda x, y, z: np.sin(x) * y + z * z, f_physical=True, geo=g3)
    assert_close_rel(got, ref['pl2_3d'], rtol=1e-9, what='project_L2 3D physical')
    got = approx.project_L2(kvs3, lambda x, y, z: x * y - z)
    assert_close_rel(got, ref['pl2_par'], rtol=1e-9, what='project_L2 parametric')
    kv1 = bspline.make_knots(2, 0.0, 1.0, 10)
    assert_close_rel(approx.project_L2(kv1, lambda x: np.cos(3 * x)), ref['pl2_1d'], rtol=1e-9, what='project_L2 1D')


def check_1d_helpers(ref):
    """1D bilinear forms through the lifted 2D device path (pyiga/assemble.py:165-230)"""
    from pyiga_b200 import assemble, bspline
    kvA, kvB = bspline.make_knots(3, 0.0, 2.0, 7), bspline.make_knots(2, 0.0, 2.0, 7)
    for du, dv in [(0, 1), (1, 0), (1, 1), (0, 0)]:
        got = assemble.bsp_mixed_deriv_biform_1d(kvA, du, dv).toarray()
        want = ref['b1d_%d%d' % (du, dv)]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (du, dv)
        got = assemble.bsp_mixed_deriv_biform_1d_asym(kvA, kvB, du, dv).toarray()
        want = ref['b1d_asym_%d%d' % (du, dv)]
        assert got.shape == want.shape == (kvB.numdofs, kvA.numdofs)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), ('asym', du, dv)
    assert np.abs(assemble.bsp_mass_1d_asym(kvA, kvB).toarray() - ref['b1d_asym_00']).max() <= 1e-12
    assert np.abs(assemble.bsp_stiffness_1d(kvA).toarray() - ref['b1d_11']).max() <= 1e-12 * np.abs(ref['b1d_11']).max()


def check_edge_cases(ref):
    """degenerate and unusual inputs the reference accepts"""
    import pytest
    from pyiga_b200 import assemble, assemblers, bspline, geometry
    # one span per axis, degree higher than the number of spans
    pc_kvs = (bspline.make_knots(3, 0.0, 1.0, 1), bspline.make_knots(2, 0.0, 1.0, 1))
    check_vs_oracle(2, (3, 2), (1, 1), 'Stiffness')
    check_vs_oracle(3, (2, 1, 3), (1, 2, 1), 'Mass')
    # degree without a sum-factorised instantiation: falls back to the per-entry kernel on the device
    asm = check_vs_oracle(2, (5, 5), (3, 4), 'Stiffness')
    assert not asm.dev.fast_path
    # empty request
    asm = assemblers.MassAssembler2D(pc_kvs, geometry.unit_square())
    assert asm.multi_entries(np.empty((0, 2), dtype=np.uint64)).shape == (0,)
    assert asm.multi_entries([]).shape == (0,)
    # indices far outside the matrix give 0 like the reference's out-of-pattern pairs
    assert asm.multi_entries(np.array([[0, 10 ** 9]], dtype=np.uint64))[0] == 0.0
    # argument checks with the reference's messages (pyiga/assemblers.pyx:1174-1185)
    with pytest.raises(AssertionError, match='Geometry has wrong source dimension'):
        assemblers.MassAssembler2D(pc_kvs, geometry.twisted_box())
    with pytest.raises(AssertionError, match='Assembler requires 3 knot vectors'):
        assemblers.StiffnessAssembler3D(pc_kvs, geometry.twisted_box())
    with pytest.raises(AssertionError, match='Geometry has wrong dimension'):
        assemble.mass(pc_kvs, geometry.twisted_box())
    with pytest.raises(ValueError, match='not open'):
        from pyiga_b200 import _lib
        bad = bspline.KnotVector(np.array([0.0, 0.0, 0.5, 1.0, 1.0, 1.0]), 2)
        assemblers.MassAssembler2D((bad, bad), geometry.unit_square())

    # a geometry that is not a spline object: Jacobian evaluated on the host and uploaded
    class Shear:
        sdim = dim = 2

        def grid_jacobian(self, grid):
            g0, g1 = np.meshgrid(*grid, indexing='ij')
            J = np.empty(g0.shape + (2, 2))
            J[..., 0, 0], J[..., 0, 1], J[..., 1, 0], J[..., 1, 1] = 1.0 + g0, 0.5, 0.25 * g1, 2.0
            return J

    kvs = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 3))
    asm = assemblers.StiffnessAssembler2D(kvs, Shear())
    got = asm.dev.be.to_host(asm.dev.assemble_mlb())
    prob = orc.Problem([kv.kv for kv in kvs], [2, 3], [kv.kv for kv in geometry.unit_square().kvs], [1, 1],
                       geometry.unit_square().coeffs)
    prob.jac = Shear().grid_jacobian(prob.grid)
    assert_close_rel(got, orc.assemble_mlb(prob, 'stiffness').ravel(), what='host-evaluated geometry')


def check_vector_forms(ref):
    """vector-valued basis functions: layouts 'packed'/'blocked', formats csr/bsr/mlb, multi_blocks"""
    from helpers import VECFORMS
    from pyiga_b200 import assemble
    for name, (form, bfuns, inputs, case, gname) in VECFORMS.items():
        kvs = make_space(ref, case)
        geo = make_geo(ref, gname)
        want = ref['vv_%s_mlb' % name]
        X = assemble.assemble(form, kvs, geo=geo, bfuns=bfuns, format='mlb', layout='packed', **inputs)
        assert X.data.shape == want.shape
        assert_close_rel(X.data, want, what='vector form %s (mlb)' % name)
        nc = bfuns[0][1]
        n = int(np.prod([kv.numdofs for kv in kvs]))
        P = assemble.assemble(form, kvs, geo=geo, bfuns=bfuns, layout='packed', format='bsr', **inputs)
        assert P.format == 'bsr' and P.blocksize == (nc, nc) and P.shape == (n * nc, n * nc)
        B = assemble.assemble(form, kvs, geo=geo, bfuns=bfuns, layout='blocked', **inputs)
        # blocked = packed with dofs regrouped by component
        perm = np.arange(n * nc).reshape(n, nc).T.ravel()
        Pd = P.toarray()
        assert abs(B.toarray() - Pd[np.ix_(perm, perm)]).max() <= RTOL * abs(Pd).max()
        # reference matrix from the MLB fixture
        S = X.structure
        I, J = (a.astype(np.int64) for a in S.nonzero())
        assert abs(Pd[I, J] - want.ravel()).max() <= RTOL * abs(want).max()
        asm = assemble.instantiate_assembler(form, kvs, dict(inputs, geo=geo), bfuns)
        assert asm.num_components() == (nc, nc)
        blocks = asm.multi_blocks([(0, 0), (0, 1), (2, 1), (5, 5)])
        assert_close_rel(blocks, ref['vv_%s_blocks' % name], what='multi_blocks ' + name)
    kvs = make_space(ref, 'a2_qa')
    A = assemble.divdiv(kvs, make_geo(ref, 'bqa'), layout='packed', format='bsr')
    assert_close_rel(A.toarray(), ref['vv_divdiv2_bsr'], what='divdiv')
    f = assemble.assemble('inner(g, v) * dx', kvs, geo=make_geo(ref, 'qa'), bfuns=[('v', 2)], g=lambda x, y: (x, -y))
    assert_close_rel(f, ref['vv_rhs2'], what='vector-valued load vector')


def _ref_csr_named(ref, name):
    import scipy.sparse
    return scipy.sparse.csr_matrix((ref[name + '_data'], ref[name + '_indices'], ref[name + '_indptr']),
                                   shape=tuple(ref[name + '_shape']))


def _assert_csr_equal(A, R, what, rtol=RTOL):
    A = A.tocsr()
    assert A.shape == R.shape, what
    assert A.has_sorted_indices or True
    assert np.array_equal(A.indptr, R.indptr), what + ': indptr'
    assert np.array_equal(A.indices, R.indices), what + ': indices'
    assert_close_rel(A.data, R.data, rtol=rtol, what=what)


def check_boundary_conditions(ref):
    """boundary dofs, Dirichlet data by interpolation, RestrictedLinearSystem
    (reference: pyiga/assemble.py:342-652; tests test_assemble.py:251-281,497-505, test_solve.py)"""
    import scipy.sparse.linalg
    from pyiga_b200 import approx, assemble, bspline, geometry
    kvs2 = (bspline.make_knots(3, 0.0, 1.0, 5), bspline.make_knots(2, 0.0, 1.0, 8))
    assert np.array_equal(assemble.boundary_dofs(kvs2, 'bottom', ravel=True), ref['bc_bd_bottom'])
    assert np.array_equal(assemble.boundary_dofs(kvs2, 'bottom', ravel=True), np.arange(10))
    assert np.array_equal(assemble.boundary_dofs(kvs2, 'right'), ref['bc_bd_right'])
    assert np.array_equal(assemble.boundary_dofs(kvs2, 'left', ravel=True, flip=(True,)), ref['bc_bd_left_flip'])
    kvs3b = (bspline.make_knots(2, 0.0, 1.0, 3), bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 2))
    assert np.array_equal(assemble.boundary_dofs(kvs3b, 'front', ravel=True), ref['bc_bd3_front'])
    assert np.array_equal(assemble.boundary_dofs(kvs3b, (1, 1)), ref['bc_bd3_top'])
    assert np.array_equal(assemble.boundary_cells(kvs3b, 'back', ravel=True), ref['bc_cells3_back'])
    for bad in [(3, 0), (0, 2), 'nowhere']:
        try:
            assemble.boundary_dofs(kvs3b, bad)
        except (ValueError, TypeError):
            pass
        else:
            raise AssertionError('invalid bdspec accepted: %r' % (bad,))

    # the identity-map case of the reference's own test (test_assemble.py:264-281)
    kvs = (bspline.make_knots(3, 0.0, 1.0, 5), bspline.make_knots(2, 0.0, 1.0, 3))
    geo = geometry.identity(kvs)
    ny, nx = (kv.numdofs for kv in kvs)
    one = lambda x, y: 1.0
    for bd, want in [((0, 0), range(nx)), ((0, 1), range((ny - 1) * nx, ny * nx)), ((1, 0), range(0, ny * nx, nx)),
                     ((1, 1), range(nx - 1, nx - 1 + ny * nx, nx))]:
        idx, val = assemble.compute_dirichlet_bc(kvs, geo, bd, one)
        assert np.array_equal(idx, list(want)) and np.allclose(val, 1.0, rtol=0, atol=1e-13)

    # Poisson on the quarter annulus (test_solve.py)
    kvsP = 2 * (bspline.make_knots(3, 0.0, 1.0, 10),)
    geoP = geometry.quarter_annulus()
    gP = lambda x, y: np.cos(x + y) + np.exp(y - x)
    fP = lambda x, y: 2 * (np.cos(x + y) - np.exp(y - x))
    idx, val = assemble.compute_dirichlet_bcs(kvsP, geoP, ('all', gP))
    assert np.array_equal(idx, ref['bc_p2_idx'])
    assert_close_rel(val, ref['bc_p2_val'], what='Dirichlet values (all sides)')
    i1, v1 = assemble.compute_dirichlet_bc(kvsP, geoP, 'top', gP)
    assert np.array_equal(i1, ref['bc_p2_top_idx'])
    assert_close_rel(v1, ref['bc_p2_top_val'], what='Dirichlet values (top)')
    assert_close_rel(approx.interpolate(kvsP, gP, geo=geoP), ref['bc_p2_uex'], what='interpolate 2D')
    rhs = assemble.inner_products(kvsP, fP, f_physical=True, geo=geoP).ravel()
    for fmt in ('csr', 'mlb'):          # host CSR input and device-resident MLMatrix input
        A = assemble.stiffness(kvsP, geo=geoP, format=fmt)
        LS = assemble.RestrictedLinearSystem(A, rhs, (idx, val))
        _assert_csr_equal(LS.A, _ref_csr_named(ref, 'bc_p2_A'), 'restricted stiffness (%s input)' % fmt)
        assert_close_rel(LS.b, ref['bc_p2_b'], what='restricted rhs')
    u = LS.complete(scipy.sparse.linalg.spsolve(LS.A.tocsc(), LS.b))
    assert_close_rel(u, ref['bc_p2_u'], rtol=1e-10, what='Poisson solution')
    assert np.sqrt(np.mean((u - ref['bc_p2_uex'].ravel()) ** 2)) < 2e-4

    # 3D: two sides, constants, vector-valued data, elim_rows, helper methods
    kvs3 = make_space(ref, 'a3_tb')
    g3 = make_geo(ref, 'tnb')
    i3, v3 = assemble.compute_dirichlet_bcs(kvs3, g3, [('front', lambda x, y, z: x * y + z), ((2, 1), 1.5)])
    assert np.array_equal(i3, ref['bc_3d_idx'])
    assert_close_rel(v3, ref['bc_3d_val'], what='3D Dirichlet values')
    iv, vv = assemble.compute_dirichlet_bc(kvs3, g3, 'bottom', lambda x, y, z: (x, y * z, 1.0 + z))
    assert np.array_equal(iv, ref['bc_3d_vec_idx'])
    assert_close_rel(vv, ref['bc_3d_vec_val'], what='3D vector Dirichlet values')
    K, M = assemble.stiffness(kvs3, geo=g3), assemble.mass(kvs3, geo=g3)
    A3 = K + M
    b3 = np.cos(np.arange(A3.shape[0]) * 0.37)
    LS3 = assemble.RestrictedLinearSystem(A3, b3, (i3, v3))
    _assert_csr_equal(LS3.A, _ref_csr_named(ref, 'bc_3d_A'), 'restricted 3D matrix')
    assert_close_rel(LS3.b, ref['bc_3d_b'], what='restricted 3D rhs')
    LS3e = assemble.RestrictedLinearSystem(A3, 0.5, (i3, 2.0), elim_rows=ref['bc_3de_rows'])
    _assert_csr_equal(LS3e.A, _ref_csr_named(ref, 'bc_3de_A'), 'restricted 3D matrix, elim_rows')
    assert_close_rel(LS3e.b, ref['bc_3de_b'], what='restricted 3D rhs, elim_rows')
    xf = np.sin(np.arange(LS3.A.shape[0]) * 0.11)
    assert_close_rel(LS3.complete(xf), ref['bc_3d_complete'], what='complete')
    assert np.array_equal(LS3.extend(xf), ref['bc_3d_extend'])
    assert np.array_equal(LS3.restrict(b3), ref['bc_3d_restrict'])
    assert np.array_equal(LS3.restrict_rhs(b3), ref['bc_3d_restrict'])
    _assert_csr_equal(LS3.restrict_matrix(M), _ref_csr_named(ref, 'bc_3d_restrM'), 'restrict_matrix')
    # restricted operator == restriction of the operator (property, any size)
    y = LS3.A_device.dot(xf)
    assert_close_rel(y, LS3.restrict_rhs(A3 @ LS3.extend(xf)), what='restricted matvec')

    # 1D model problem -u'' = 1, u(0) = 0, u(1) = 1 (test_assemble.py:497-505)
    kv1 = bspline.make_knots(2, 0.0, 1.0, 10)
    assert_close_rel(assemble.inner_products(kv1, lambda x: 1.0 + x), ref['bc_1d_f'], what='1D inner products')
    assert_close_rel(approx.interpolate(kv1, lambda x: 0.5 * x * (3 - x)), ref['bc_1d_interp'], what='1D interpolate')
    A1 = assemble.stiffness(kv1)
    f1 = assemble.inner_products(kv1, lambda x: 1.0)
    LS1 = assemble.RestrictedLinearSystem(A1, f1, [(0, kv1.numdofs - 1), (0.0, 1.0)])
    u1 = LS1.complete(np.linalg.solve(LS1.A.toarray(), LS1.b))
    assert np.linalg.norm(u1 - ref['bc_1d_interp']) < 1e-12
    assert_close_rel(approx.interpolate(kvs3, lambda x, y, z: np.sin(x) * y + z * z, geo=g3), ref['bc_interp3'],
                     what='3D interpolate')


def check_partial_rows(ref):
    """submatrix of a subset of rows (reference: pyiga/_hdiscr.py:5-12)"""
    from pyiga_b200 import assemble, assemblers
    for case, cls, gname in [('a3_mixed', assemblers.StiffnessAssembler3D, 'tb'), ('a2_mixed', assemblers.MassAssembler2D, 'bqa')]:
        kvs = make_space(ref, case)
        asm = cls(kvs, make_geo(ref, gname))
        rows = ref['pr_%s_rows' % case]
        R = _ref_csr_named(ref, 'pr_%s' % case)
        A = assemble.assemble_partial_rows(asm, rows)
        _assert_csr_equal(A, R, 'partial rows ' + case)
        # restricted variant keeps the order (and duplicates) of the request
        perm = np.concatenate((rows[::-1], rows[:2]))
        B = assemble.assemble_partial_rows(asm, perm, restrict=True)
        assert B.shape == (perm.size, R.shape[1])
        assert abs(B - R[perm]).max() <= RTOL * abs(R).max()
        # against the full matrix of the fast path
        full = asm.assemble_csr()
        assert abs(A[rows] - full[rows]).max() <= RTOL * abs(full).max()
        E = assemble.assemble_partial_rows(asm, [])
        assert E.shape == R.shape and E.nnz == 0
        try:
            assemble.assemble_partial_rows(asm, [R.shape[0]])
        except ValueError:
            pass
        else:
            raise AssertionError('out-of-range row accepted')
    # a compiled variational form (what HDiscretization._assemble_level hands over, pyiga/_hdiscr.py:36-57)
    from helpers import VFORMS
    form, inputs, case, gname = VFORMS['cd3']
    kvs = make_space(ref, case)
    asm = assemble.instantiate_assembler(form, kvs, dict(inputs, geo=make_geo(ref, gname)))
    full = assemble.assemble(asm).tocsr()
    rows = np.array([0, 7, full.shape[0] // 2, full.shape[0] - 1])
    part = assemble.assemble_partial_rows(asm, rows)
    assert abs(part[rows] - full[rows]).max() <= RTOL * abs(full).max()
    assert part.nnz == full[rows].nnz
    # the call HDiscretization._assemble_level makes: on-demand class, bounding box of cells
    from pyiga_b200 import vform as vfm
    cls = vfm.compile_vform(vfm.parse_vf(form, kvs, args=inputs), on_demand=True)
    asm2 = cls(kvs, geo=make_geo(ref, gname), bbox=tuple((0, kv.numspans) for kv in kvs), **inputs)
    part2 = assemble.assemble_partial_rows(asm2, rows)
    assert abs(part2 - part).max() <= 1e-14 * abs(full).max()


def check_boundary_forms(ref):
    """integrals over one side of the patch (`boundary=` of assemble.assemble;
    reference: pyiga/assemble.py:899-940, pyiga/codegen/cython.py:549-590, quadrature.py:23-31)"""
    import scipy.sparse
    from helpers import BFORMS
    from pyiga_b200 import assemble
    for name, (form, bfuns, inputs, case, gname, bd) in BFORMS.items():
        kvs = make_space(ref, case)
        got = assemble.assemble(form, kvs, geo=make_geo(ref, gname), bfuns=bfuns, boundary=bd, **inputs)
        if 'bf_%s_indptr' % name in ref:
            R = _ref_csr_named(ref, 'bf_%s' % name)
            assert scipy.sparse.issparse(got) and got.shape == R.shape, name
            assert abs(got - R).max() <= RTOL * abs(R).max(), 'boundary form %s: %.3e' % (name, abs(got - R).max())
            # entries the reference stores are the pattern of the boundary space
            G = got.tocsr(); G.sort_indices()
            assert np.array_equal(G.indptr, R.indptr) and np.array_equal(G.indices, R.indices), name
        else:
            assert_close_rel(got, ref['bf_%s' % name], what='boundary form ' + name)
    # the protocol on a boundary assembler: kvs without the normal axis, entries of the face space
    form, bfuns, inputs, case, gname, bd = BFORMS['flux3_front']
    kvs = make_space(ref, case)
    asm = assemble.instantiate_assembler(form, kvs, dict(inputs, geo=make_geo(ref, gname)), bfuns, boundary=bd)
    assert len(asm.kvs[0]) == 2 and asm.kvs[0] == tuple(kvs[1:])
    R = _ref_csr_named(ref, 'bf_flux3_front')
    ij = np.array([(0, 0), (3, 4), (R.shape[0] - 1, R.shape[0] - 2), (0, R.shape[0] - 1)])
    want = np.asarray(R[ij[:, 0], ij[:, 1]]).ravel()
    assert np.abs(asm.multi_entries(ij) - want).max() <= RTOL * abs(R).max()
    # updating the geometry of a boundary form refreshes the surface measure and the normal too:
    # same matrix as a freshly built assembler
    kvs = make_space(ref, 'a2_qa')
    g1, g2 = make_geo(ref, 'qa'), make_geo(ref, 'bqa')
    for form in ('u * v * ds', 'inner(grad(u), n) * v * ds'):
        upd = assemble.Assembler(form, kvs, geo=g1, boundary='left', updatable=['geo'])
        fresh = assemble.assemble(form, kvs, geo=g2, boundary='left')
        got = upd.assemble(geo=g2)
        assert abs(got - fresh).max() <= 1e-14 * abs(fresh).max(), form


def check_two_spaces(ref):
    """Petrov-Galerkin forms: trial functions in space 0 (columns), test functions in space 1 (rows)
    (reference: pyiga/assemble.py:947-951, generated __init__(kvs0, kvs1, ...))"""
    from helpers import PGFORMS
    from pyiga_b200 import assemble, bspline
    for name, (form, bfuns, inputs, p0, p1, ns, gname) in PGFORMS.items():
        kvs0 = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(p0, ns))
        kvs1 = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(p1, ns))
        R = _ref_csr_named(ref, 'pg_%s' % name)
        A = assemble.assemble(form, (kvs0, kvs1), geo=make_geo(ref, gname), bfuns=bfuns, **inputs)
        _assert_csr_equal(A, R, 'two-space form ' + name)
        M = assemble.assemble(form, (kvs0, kvs1), geo=make_geo(ref, gname), bfuns=bfuns, format='mlb', **inputs)
        x = np.cos(np.arange(R.shape[1]) * 0.3)
        assert_close_rel(M.dot(x), R @ x, what='two-space MLB matvec ' + name)
        asm = assemble.instantiate_assembler(form, (kvs0, kvs1), dict(inputs, geo=make_geo(ref, gname)), bfuns)
        assert asm.kvs == (kvs0, kvs1)
        ij = np.array([(0, 0), (R.shape[0] - 1, R.shape[1] - 1), (5, 3), (R.shape[0] - 1, 0)])
        want = np.asarray(R[ij[:, 0], ij[:, 1]]).ravel()
        assert np.abs(asm.multi_entries(ij) - want).max() <= RTOL * abs(R).max()
    try:
        assemble.assemble('u * v * dx', (kvs0, kvs1), geo=make_geo(ref, 'tnb'), bfuns=[('u', 1, 1), ('v', 1, 0)])
    except NotImplementedError:
        pass
    else:
        raise AssertionError('unsupported space assignment accepted')


def check_csr_pattern_host(ref):
    """host-side closed form of the CSR pattern (pb200_csr_pattern_host) == the reference's
    MLStructure.nonzero + COO->CSR pattern, for whole matrices and row slabs, int32 and int64"""
    from pyiga_b200 import assemblers
    for case, cls, gname in [('a3_mixed', assemblers.StiffnessAssembler3D, 'tb'), ('a2_mixed', assemblers.MassAssembler2D, 'bqa'),
                             ('a3_mult', assemblers.MassAssembler3D, 'cyl')]:
        kvs = make_space(ref, case)
        asm = cls(kvs, make_geo(ref, gname))
        ds = asm.dev.device_structure
        key = 'stiff' if 'Stiff' in cls.__name__ else 'mass'
        indptr, indices = ref['%s_%s_indptr' % (case, key)], ref['%s_%s_indices' % (case, key)]
        n0 = kvs[0].numdofs
        inner = int(np.prod([kv.numdofs for kv in kvs[1:]]))
        for (ra, rb), idt, nthr in [((0, n0), np.int32, 1), ((0, n0), np.int64, 3), ((1, n0 - 2), np.int32, 4), ((2, 3), np.int64, 7)]:
            r0, r1 = ra * inner, rb * inner
            want_ptr = indptr[r0:r1 + 1] - indptr[r0]
            want_idx = indices[indptr[r0]:indptr[r1]]
            got_ptr = np.full(want_ptr.size, -7, dtype=idt)
            got_idx = np.full(want_idx.size, -7, dtype=idt)
            ds.csr_pattern_host(got_ptr, got_idx, row0=(ra, rb), indptr_offset=5, nthreads=nthr)
            assert np.array_equal(got_ptr, want_ptr + 5), (case, ra, rb)
            assert np.array_equal(got_idx, want_idx), (case, ra, rb)
    # larger structures: the staging buffer of the writer is flushed many times, pieces of several threads start
    # at arbitrary (unaligned) positions of the result, rows of different widths alternate at the patch boundary
    import scipy.sparse
    from pyiga_b200 import bspline
    from pyiga_b200._mlb import DeviceStructure
    from pyiga_b200.mlmatrix import MLStructure
    for kvs, kvs_test in [(tuple(bspline.make_knots(2, 0.0, 1.0, n) for n in (9, 10, 11)), None),
                          ((bspline.make_knots(3, 0.0, 1.0, 40), bspline.make_knots(1, 0.0, 1.0, 50)), None),
                          (tuple(bspline.make_knots(3, 0.0, 1.0, n) for n in (4, 6, 5)),
                           tuple(bspline.make_knots(2, 0.0, 1.0, n) for n in (4, 6, 5)))]:
        S = MLStructure.from_kvs(kvs_test or kvs, kvs)
        I, J = (a.astype(np.int64) for a in S.nonzero())
        A = scipy.sparse.csr_matrix((np.ones(I.size), (I, J)), shape=S.shape)
        A.sort_indices()
        ds = DeviceStructure(S)
        n0 = S.bs[0][0]
        inner = S.shape[0] // n0
        for (ra, rb), idt, nthr, shift in [((0, n0), np.int32, 1, 0), ((0, n0), np.int32, 8, 3), ((0, n0), np.int64, 5, 1),
                                           ((2, n0 - 1), np.int32, 3, 7), ((1, 2), np.int64, 2, 5)]:
            r0, r1 = ra * inner, rb * inner
            want_ptr = A.indptr[r0:r1 + 1].astype(np.int64) - A.indptr[r0]
            want_idx = A.indices[A.indptr[r0]:A.indptr[r1]]
            # `shift` entries of padding in front: the result does not start on a cache line
            buf_ptr = np.full(want_ptr.size + shift + 16, -7, dtype=idt)
            buf_idx = np.full(want_idx.size + shift + 16, -7, dtype=idt)
            got_ptr, got_idx = buf_ptr[shift:shift + want_ptr.size], buf_idx[shift:shift + want_idx.size]
            ds.csr_pattern_host(got_ptr, got_idx, row0=(ra, rb), nthreads=nthr)
            assert np.array_equal(got_ptr, want_ptr) and np.array_equal(got_idx, want_idx), (S.shape, ra, rb, nthr)
            assert (buf_idx[:shift] == -7).all() and (buf_idx[shift + want_idx.size:] == -7).all()     # nothing outside
            assert (buf_ptr[:shift] == -7).all() and (buf_ptr[shift + want_ptr.size:] == -7).all()
            # the two-part fill of the pipelined delivery (_hostcsr: pattern threads + the calling thread)
            from pyiga_b200._hostcsr import pattern_split
            rs0 = np.asarray(S._row_tables(0)[0])
            inner_b = int(np.prod([len(b) for b in S.bidx[1:]], dtype=np.int64))
            for nthr2 in (1, 3):
                rm, own = pattern_split(rs0, ra, rb, nthr2, inner, inner_b)
                two_ptr, two_idx = np.full_like(got_ptr, -9), np.full_like(got_idx, -9)
                ds.csr_pattern_host(two_ptr, two_idx, row0=(ra, rm), nthreads=nthr2)
                if own is not None:
                    assert own[0] == rm and own[1] == rb and ra < rm < rb
                    ds.csr_pattern_host(two_ptr[own[2]:], two_idx[own[3]:], row0=(rm, rb), indptr_offset=own[3], nthreads=1)
                else:
                    assert rm == rb
                assert np.array_equal(two_ptr, want_ptr) and np.array_equal(two_idx, want_idx), (S.shape, ra, rb, nthr2)


def check_boundary_reference_tests():
    """the reference's own boundary-integral tests with analytic answers
    (test/test_assemble.py:336-400: test_assemble_boundary_vector, test_assemble_boundary_matrix)"""
    from pyiga_b200 import bspline, geometry
    from pyiga_b200.assemble import assemble, stiffness
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 3),)
    geo_3d = geometry.tensor_product(geometry.line_segment(0.0, 1.0), geometry.quarter_annulus())
    f = assemble('v * ds', kvs, geo=geo_3d, boundary='left')
    assert f.shape == (6, 6, 1)
    assert np.allclose(f.sum(), (2 * 1 * np.pi) / 4)
    assert np.allclose(assemble('v * ds', kvs, geo=geo_3d, boundary='right').sum(), (2 * 2 * np.pi) / 4)
    assert np.allclose(assemble('v * ds', kvs, geo=geo_3d, boundary='bottom').sum(), 1.0)
    assert np.allclose(assemble('v * ds', kvs, geo=geo_3d, boundary='top').sum(), 1.0)
    assert np.allclose(assemble('v * ds', kvs, geo=geo_3d, boundary='front').sum(), (2 ** 2 - 1 ** 2) * np.pi / 4)
    assert np.allclose(assemble('v * ds', kvs, geo=geo_3d, boundary='back').sum(), (2 ** 2 - 1 ** 2) * np.pi / 4)
    ann = (2 ** 2 - 1 ** 2) * np.pi / 4
    for bd, want in [('left', [-1, -1, 0]), ('right', [2, 2, 0]), ('bottom', [0, -1, 0]), ('top', [-1, 0, 0]),
                     ('front', [0, 0, -ann]), ('back', [0, 0, ann])]:
        nv = assemble('inner(v, n) * ds', kvs, bfuns=[('v', 3)], geo=geo_3d, boundary=bd, layout='packed')
        assert np.allclose(nv.sum(axis=(0, 1, 2)), want), (bd, nv.sum(axis=(0, 1, 2)))
    kvs2 = 2 * (bspline.make_knots(3, 0.0, 1.0, 3),)
    sq = geometry.unit_square()
    for bd, want in [('left', [-1, 0]), ('right', [1, 0]), ('bottom', [0, -1]), ('top', [0, 1])]:
        nv = assemble('inner(v, n) * ds', kvs2, bfuns=[('v', 2)], geo=sq, boundary=bd, layout='packed')
        assert np.allclose(nv.sum(axis=(0, 1)), want), (bd, nv.sum(axis=(0, 1)))
    kvs = (bspline.make_knots(3, 0.0, 1.0, 3), bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 5))
    A = assemble('inner(grad(u), grad(v)) * ds', kvs, geo=geo_3d, boundary='left')
    assert A.shape == (6 * 7, 6 * 7)
    A = assemble('inner(grad(u), grad(v)) * ds', kvs, geo=geo_3d, boundary='top')
    assert A.shape == (6 * 8, 6 * 8)
    # tangential components only: the 2D Laplacian of the plane face 'front'
    A = assemble('inner(cross(n, grad(u)), cross(n, grad(v))) * ds', kvs, geo=geo_3d, boundary='front')
    assert A.shape == (7 * 8, 7 * 8)
    A2 = stiffness(kvs[1:], geo=geometry.quarter_annulus())
    assert np.allclose(A.toarray(), A2.toarray())


def check_surface_forms(ref):
    """integrals over a 2D manifold in R^3 (`ds` without `boundary`; reference: pyiga/vform.py:171-236,
    test/test_assemble.py:314-334)"""
    import scipy.sparse
    from helpers import SFORMS
    from pyiga_b200 import assemble, bspline, geometry
    from pyiga_b200.vform import VForm
    cyl3 = geometry.tensor_product(geometry.line_segment(0.0, 1.0), geometry.quarter_annulus())
    for name, (form, bfuns, inputs, ps, ns, side) in SFORMS.items():
        kvs = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
        got = assemble.assemble(form, kvs, geo=cyl3.boundary(side), bfuns=bfuns, **inputs)
        if 'sf_%s_indptr' % name in ref:
            R = _ref_csr_named(ref, 'sf_%s' % name)
            _assert_csr_equal(got, R, 'surface form ' + name)
        else:
            assert_close_rel(got, ref['sf_%s' % name], what='surface form ' + name)
    # the reference's own test: surface area of the cylinder mantles through a VForm object
    vf = VForm(2, geo_dim=3, arity=1)
    v = vf.basisfuns()
    vf.add(v * vf.ds)
    kvs = 2 * (bspline.make_knots(3, 0.0, 1.0, 10),)
    f = assemble.assemble(vf, kvs, geo=cyl3.boundary('left'))
    assert np.allclose(f.sum(), (2 * 1 * np.pi) / 4)
    f = assemble.assemble(vf, kvs, geo=cyl3.boundary('right'))
    assert np.allclose(f.sum(), (2 * 2 * np.pi) / 4)
    # callables of the three physical coordinates are accepted as well
    kvs = 2 * (bspline.make_knots(2, 0.0, 1.0, 4),)
    a = assemble.assemble('c * u * v * ds', kvs, geo=cyl3.boundary('left'), c=lambda x, y, z: 1.0 + x * y + z)
    _assert_csr_equal(a, _ref_csr_named(ref, 'sf_sreact'), 'surface form with f(x, y, z)')
    # Laplace-Beltrami (beyond the reference, whose generator cannot invert the 3 x 2 Jacobian):
    # symmetric, constants in the kernel, and the energy of the axial coordinate z = xi_0 is the area
    kvs = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 5))
    K = assemble.assemble('inner(grad(u), grad(v)) * ds', kvs, geo=cyl3.boundary('right'))
    assert abs(K - K.T).max() <= 1e-12 * abs(K).max()
    assert np.abs(K @ np.ones(K.shape[0])).max() <= 1e-11 * abs(K).max()
    from pyiga_b200 import approx
    z = approx.interpolate(kvs, lambda x, y: 0.0 * x + y).ravel()      # first parameter = axis of the cylinder
    assert np.allclose(z @ (K @ z), (2 * 2 * np.pi) / 4)


def check_reference_driver_dropin(ref):
    """the REAL reference's drivers (oracle/_ref, test infrastructure) accept
    the device assemblers through the duck-typed protocol: pyiga.assemble.assemble_entries(asm) for
    scalar, symmetric and vector-valued assemblers, pyiga's _assemble_partial_rows, MLMatrix(data=...)"""
    import os
    import sys
    import pytest
    refdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
    if not os.path.isdir(os.path.join(refdir, 'pyiga')):
        pytest.skip('oracle/_ref is not installed')
    sys.path.insert(0, refdir)
    try:
        import pyiga.assemble as rasm
        import pyiga._hdiscr as rh
    except Exception as exc:        # pragma: no cover
        pytest.skip('reference does not import: %s' % exc)
    finally:
        sys.path.remove(refdir)
    from pyiga_b200 import assemble, assemblers
    for case, cls, key, gname in [('a3_mixed', assemblers.StiffnessAssembler3D, 'stiff', 'tb'),
                                  ('a2_mixed', assemblers.MassAssembler2D, 'mass', 'bqa')]:
        kvs = make_space(ref, case)
        asm = cls(kvs, make_geo(ref, gname))
        want = asm.assemble_csr()
        for sym in (False, True):       # the reference mirrors the lower triangle itself when symmetric=True
            A = rasm.assemble_entries(asm, symmetric=sym).tocsr()
            A.sort_indices()
            assert np.array_equal(A.indptr, ref['%s_%s_indptr' % (case, key)])
            assert np.array_equal(A.indices, ref['%s_%s_indices' % (case, key)])
            assert abs(A - want).max() <= RTOL * abs(want).max()
        # (format='mlb' of the reference goes through its Cython-typed core and is not duck-typed)
        rows = ref['pr_%s_rows' % case] if 'pr_%s_rows' % case in ref else np.array([0, 3])
        P = rh._assemble_partial_rows(asm, rows)
        assert abs(P[rows] - want[rows]).max() <= RTOL * abs(want).max()
    # (assemble_entries_vec and format='mlb' of the reference take Cython-typed assemblers only)


def check_poisson_end_to_end():
    """the whole flow of test/test_solve.py in 3D on the rational twisted box: the discretisation
    error falls with the mesh width at the rate of the spline degree"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('poisson_demo', os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), 'tools', 'poisson_demo.py'))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    r1 = demo.solve(2, 4)
    r2 = demo.solve(2, 8)
    assert r1['cg_info'] == 0 and r2['cg_info'] == 0
    assert r2['rms_error_vs_interpolant'] < 0.3 * r1['rms_error_vs_interpolant'], (r1, r2)
    assert r2['rms_error_vs_interpolant'] < 1e-3, r2


# ---------------------------------------------------------------------------------------------
# benchmark-scale parity against sampled entries of the real reference
# (tests/golden/large_*.npz, made by tests/golden/make_golden_large.py)
# ---------------------------------------------------------------------------------------------
LARGE_CONVDIFF = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'


def load_large(name):
    import os
    from helpers import GOLDEN
    return np.load(os.path.join(GOLDEN, 'large_%s.npz' % name))


def large_assembler(z):
    """the device assembler of a large fixture: (asm, kvs, geo)"""
    from pyiga_b200 import assemble, assemblers, bspline, geometry
    form, p, n = str(z['form']), int(z['p']), int(z['n'])
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box() if str(z['geo']) == 'tnb' else geometry.twisted_box()
    if form == 'stiffness':
        asm = assemblers.StiffnessAssembler3D(kvs, geo)
    elif form == 'mass':
        asm = assemblers.MassAssembler3D(kvs, geo)
    else:
        asm = assemble.instantiate_assembler(LARGE_CONVDIFF, kvs, {'geo': geo, 'diff_coeff': lambda x, y, z: 1.0 + x * y}, None)
    return asm, kvs, geo


def sample_errors(dev, d_data, rows, z, pos=None):
    """max |ours - reference| over the sampled entries whose row lies in the first-axis row slab
    `rows`, read out of the device MLB buffer `d_data` of that slab; also checks that pairs for
    which the reference returned 0 outside the pattern are outside ours.  Returns (max abs error,
    number of entries compared)."""
    be = dev.be
    ij = z['ij'].astype(np.int64)
    want = z['val']
    S = dev.structure
    if pos is None:
        pos = S.positions(ij[:, 0], ij[:, 1])
    nout = int(z['nout'])
    assert np.all(pos[len(pos) - nout:] == -1), 'pairs outside the reference pattern are inside ours'
    assert np.all(pos[:len(pos) - nout] >= 0), 'pairs inside the reference pattern are outside ours'
    plane_rows = int(np.prod(dev.ndofs_test[1:], dtype=np.int64))
    i0 = ij[:, 0] // plane_rows
    rs = dev.row_start0()
    inner = int(np.prod(dev.nband[1:], dtype=np.int64))
    sel = (pos >= 0) & (i0 >= rows[0]) & (i0 < rows[1])
    if not sel.any():
        return 0.0, 0
    local = pos[sel] - int(rs[rows[0]]) * inner
    assert local.min() >= 0 and local.max() < dev.slab_size(rows)
    got = be.to_host(d_data[be.from_host(local)])
    return float(np.abs(got - want[sel]).max()), int(sel.sum())


def check_large(name, nslabs=1):
    """sum-factorised assembly at a benchmarked size against the reference's sampled entries, as one
    slab or as `nslabs` independent row slabs (the multi-GPU sharding on one device)"""
    from pyiga_b200.dist import partition_rows
    z = load_large(name)
    asm, kvs, geo = large_assembler(z)
    dev = asm.dev
    assert dev.fast_path
    scale = float(z['full_maxabs']) if 'full_maxabs' in z else float(z['sample_maxabs'])
    ij = z['ij'].astype(np.int64)
    pos = dev.structure.positions(ij[:, 0], ij[:, 1])
    slabs = partition_rows(dev, nslabs) if nslabs > 1 else [(0, dev.ndofs_test[0])]
    worst, count = 0.0, 0
    for rows in slabs:
        if nslabs > 1:
            if hasattr(asm, 'compute_fields_rows'):
                asm.compute_fields_rows(rows)
            else:
                dev.compute_fields(geo, rows=rows)
        data = dev.assemble_mlb(rows=rows)
        err, cnt = sample_errors(dev, data, rows, z, pos)
        worst, count = max(worst, err), count + cnt
        del data
    assert count == len(pos) - int(z['nout']), 'every sampled entry belongs to exactly one slab'
    assert worst <= RTOL * scale, '%s in %d slabs: max abs error %.3e > %.1e * max|A_ref| (%.3e)' % (name, nslabs, worst, RTOL, scale)
    return worst / scale


def check_large_full(name):
    """p=3 n=64: whole-matrix checks against the reference's full assembly — nnz, sum of all entries,
    sum of absolute values, max|A| and a strided matvec checksum (A x)[::stride]"""
    z = load_large(name)
    asm, kvs, geo = large_assembler(z)
    M = asm.assemble_mlb()
    d = M.data                  # host copy of the device tensor (0.76 GB at n=64)
    scale = float(z['full_maxabs'])
    nnz = int(z['full_nnz'])
    assert M.nnz == nnz
    assert abs(float(np.abs(d).max()) - scale) <= RTOL * scale
    # sums of 9.5e7 terms: compare with the rounding of the summation, not of a single entry
    assert abs(float(d.sum()) - float(z['full_sum'])) <= 1e-12 * float(z['full_abs_sum'])
    assert abs(float(np.abs(d).sum()) - float(z['full_abs_sum'])) <= 1e-12 * float(z['full_abs_sum'])
    x = np.cos(float(z['mv_freq']) * np.arange(M.shape[1]) + float(z['mv_phase']))
    y = M.dot(x)[::int(z['mv_stride'])]
    rowmax = (2 * int(z['p']) + 1) ** 3 * scale     # at most (2p+1)^3 entries per row
    assert np.abs(y - z['mv_y']).max() <= RTOL * rowmax
    # CSR through the public driver: int32, sorted, and the same values at the sampled pairs
    A = asm.assemble_csr()
    assert A.nnz == nnz and A.indices.dtype == np.int32 and A.has_sorted_indices
    ij = z['ij'].astype(np.int64)
    got = np.asarray(A[ij[:, 0], ij[:, 1]]).ravel()
    assert np.abs(got - z['val']).max() <= RTOL * scale


# ---------------------------------------------------------------------------------------------
# native slab-distributed CG (csrc/distcg.cuh): halo, dot products and the preconditioner's gather
# through peer-mapped windows
# ---------------------------------------------------------------------------------------------
def check_inner_products_vector_valued():
    """inner_products with a vector-valued f returns ndofs + (k,) like the reference (pyiga/assemble.py:318-340)"""
    from pyiga_b200 import assemble, bspline, geometry
    kvs = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 3))
    geo = geometry.quarter_annulus()
    for phys in (False, True):
        g = geo if phys else None
        V = assemble.inner_products(kvs, lambda x, y: (x * y, 1.0 + x), f_physical=phys, geo=g)
        a = assemble.inner_products(kvs, lambda x, y: x * y, f_physical=phys, geo=g)
        b = assemble.inner_products(kvs, lambda x, y: 1.0 + x + 0.0 * y, f_physical=phys, geo=g)
        assert V.shape == a.shape + (2,)
        assert np.abs(V[..., 0] - a).max() <= 1e-15 and np.abs(V[..., 1] - b).max() <= 1e-15


def check_native_distributed_cg(world=1, rank=0, p=2, n=(7, 4, 5), rtol=1e-11):
    """mass matrix on the rational twisted box, b = M x*, Kronecker preconditioner of the inverse 1D
    mass matrices (pyiga/approx.py:82-93): the distributed matvec equals the slab of M p, the solve
    recovers x* and runs as many iterations as scipy's cg on the assembled matrix (+-1)."""
    import scipy.sparse.linalg
    from pyiga_b200 import assemble, assemblers, bspline, geometry
    from pyiga_b200.dist import partition_rows
    from pyiga_b200.distcg import DistributedCG
    kvs = tuple(bspline.make_knots(p, 0.0, 1.0, nk) for nk in n)
    geo = geometry.twisted_nurbs_box()
    asm = assemblers.MassAssembler3D(kvs, geo)
    dev = asm.dev
    be = dev.be
    slabs = partition_rows(dev, world)
    assert len(slabs) == world
    rows = slabs[rank]
    mlb = dev.assemble_mlb(rows=rows)
    Ainv = [np.linalg.inv(assemble.bsp_mass_1d(kv).toarray()) for kv in kvs]
    cg = DistributedCG(dev.device_structure, mlb, slabs, rank, Ainv)
    # reference on the host: the whole matrix (every rank assembles it; small)
    A = asm.assemble_csr()
    N = A.shape[0]
    plane = kvs[1].numdofs * kvs[2].numdofs
    lo, hi = rows[0] * plane, rows[1] * plane
    rng = np.random.default_rng(5)
    pvec = rng.standard_normal(N)
    y = be.to_host(cg.matvec(be.from_host(pvec[lo:hi])))
    want = (A @ pvec)[lo:hi]
    assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max(), 'distributed matvec'
    xstar = np.cos(0.3 * np.arange(N))
    b = A @ xstar
    x, its, res = cg.solve(be.from_host(b[lo:hi]), rtol=rtol, maxiter=300, check_every=7)
    x = be.to_host(x)
    assert res <= rtol
    assert np.abs(x - xstar[lo:hi]).max() <= 1e-7 * np.abs(xstar).max(), np.abs(x - xstar[lo:hi]).max()
    # scipy's cg with the same preconditioner
    Minv = scipy.sparse.linalg.LinearOperator((N, N), matvec=lambda v: np.einsum(
        'ia,jb,kc,abc->ijk', Ainv[0], Ainv[1], Ainv[2], v.reshape([kv.numdofs for kv in kvs])).ravel())
    count = [0]
    xs, info = scipy.sparse.linalg.cg(A, b, rtol=rtol, atol=0.0, maxiter=300, M=Minv, callback=lambda _: count.__setitem__(0, count[0] + 1))
    assert info == 0 and abs(count[0] - its) <= 1, (count[0], its)
    # a second solve on the same object (stamps keep increasing)
    x2, its2, _ = cg.solve(be.from_host(2.0 * b[lo:hi]), rtol=rtol, maxiter=300, check_every=3)
    assert its2 == its and np.abs(be.to_host(x2) - 2.0 * xstar[lo:hi]).max() <= 2e-7 * np.abs(xstar).max()
    cg.close()
    return its


# ---------------------------------------------------------------------------------------------
# the reference's own VForm objects on the device (pyiga_b200.refvform) and its hierarchical driver
# ---------------------------------------------------------------------------------------------
def _import_reference():
    """the real reference from oracle/_ref (test infrastructure); skips when it is not installed"""
    import os
    import sys
    import pytest
    refdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
    if not os.path.isdir(os.path.join(refdir, 'pyiga')):
        pytest.skip('oracle/_ref is not installed')
    if refdir not in sys.path:
        sys.path.insert(0, refdir)
    try:
        import pyiga
        return pyiga
    except Exception as exc:        # pragma: no cover
        pytest.skip('reference does not import: %s' % exc)


def _vform_fixture():
    import os
    import sys
    from helpers import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import refvform_cases as rc
    return rc, np.load(os.path.join(GOLDEN, 'ref_vform_objects.npz'))


def check_reference_vform_objects():
    """pyiga.vform.VForm objects (programmatic and parsed from strings by the REFERENCE's parse_vf) go
    through pyiga_b200.assemble.assemble and give the matrices of the reference's JIT-compiled assemblers
    (fixture: tests/golden/make_golden_vform.py).  Only the reference's pure-Python vform module runs here."""
    _import_reference()
    from pyiga_b200 import assemble
    rc, fix = _vform_fixture()
    for name, (make, kvs, geo, inputs) in rc.cases().items():
        want = fix['vf_' + name]
        got = assemble.assemble(make(), kvs, geo=geo, **inputs)
        got = got.toarray() if hasattr(got, 'toarray') else np.asarray(got)
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), '%s: %.3e' % (name, np.abs(got - want).max())
    # Assembler with updatable inputs on a reference VForm (pyiga/assemble.py:958-1003): update == fresh assembler
    make, kvs, geo, inputs = rc.cases()['aniso2']
    A2 = lambda x, y: 2.0 * rc._A(x, y) + np.eye(2)
    asm = assemble.Assembler(make(), kvs, geo=geo, updatable=['A'], **inputs)
    first = asm.assemble().toarray()
    assert np.abs(first - fix['vf_aniso2']).max() <= RTOL * np.abs(fix['vf_aniso2']).max()
    upd = asm.assemble(A=A2).toarray()
    fresh = assemble.assemble(make(), kvs, geo=geo, **dict(inputs, A=A2)).toarray()
    assert np.abs(upd - fresh).max() <= 1e-14 * np.abs(fresh).max() and np.abs(upd - first).max() > 0.1 * np.abs(first).max()
    asm.asm.update_params(c=7.0)
    fresh = assemble.assemble(make(), kvs, geo=geo, **dict(inputs, A=A2, c=7.0)).toarray()
    assert np.abs(asm.assemble().toarray() - fresh).max() <= 1e-14 * np.abs(fresh).max()
    with pytest.raises(ValueError):
        asm.asm.update(nonexistent=A2)
    # integrals over a side of the patch (VForm(dim, boundary=True): ds, the normal vector, Jac_to_boundary)
    for name, (make, kvs, geo, inputs, sides) in rc.bcases().items():
        for bd in sides:
            want = fix['bd_%s_%s' % (name, bd)]
            got = assemble.assemble(make(), kvs, geo=geo, boundary=bd, **inputs)
            got = got.toarray() if hasattr(got, 'toarray') else np.asarray(got)
            assert got.shape == want.shape, (name, bd)
            assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (name, bd, np.abs(got - want).max())
    # Petrov-Galerkin forms: trial and test functions in different spaces on the same mesh
    for name, (make, kvs2, geo, inputs) in rc.pgcases().items():
        want = fix['pg_' + name]
        got = assemble.assemble(make(), kvs2, geo=geo, **inputs).toarray()
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (name, np.abs(got - want).max())
    # second and mixed derivative slots (numderiv = 2): the wave form runs through the sum-factorised walks with the
    # (1st, 2nd derivative) tables, and the per-entry kernel gives the same matrix; the fourth-order form needs all
    # three derivative orders on one axis in one term group and takes the per-entry path
    from pyiga_b200 import _device, vform as dvform
    be = _device.backend()
    for name, fast in (('wave_st2', True), ('wave_st3', True), ('biharmonic2', False)):
        make, kvs, geo, inputs = rc.cases()[name]
        asm = dvform.compile_vform(make())(kvs, geo=geo, **inputs)
        assert bool(asm.dev.fast_path) == fast, name
        a, b = be.to_host(asm.dev.assemble_mlb()), be.to_host(asm.dev.assemble_mlb(entrywise=True))
        assert np.abs(a - b).max() <= RTOL * np.abs(b).max(), name
        want = fix['vf_' + name]
        I, J = np.nonzero(want)
        ij = np.stack([I, J], 1)[::3]
        vals = np.asarray(asm.multi_entries(ij))
        assert np.abs(vals - want[ij[:, 0], ij[:, 1]]).max() <= RTOL * np.abs(want).max(), name


def check_space_time_assemblers():
    """HeatAssembler_ST / WaveAssembler_ST of the predefined set (pyiga/assemblers.pyx:351-690, 1542-1957) against
    the matrices of the reference's own classes (fixture entries vf_heat_st*, vf_wave_st*: same spaces and
    geometries, tests/golden/refvform_cases.py).  No reference code runs here."""
    from pyiga_b200 import _device, assemble, assemblers, bspline, geometry, vform
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_vform_objects.npz'))
    kv2 = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 3))
    kv3 = (bspline.make_knots(2, 0.0, 1.0, 3), bspline.make_knots(1, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 2))
    be = _device.backend()
    cases = [('heat_st2', assemblers.HeatAssembler_ST2D, kv2, geometry.unit_square()),
             ('heat_st3', assemblers.HeatAssembler_ST3D, kv3, geometry.twisted_box()),
             ('wave_st2', assemblers.WaveAssembler_ST2D, kv2, geometry.quarter_annulus()),
             ('wave_st3', assemblers.WaveAssembler_ST3D, kv3, geometry.twisted_box())]
    for name, cls, kvs, geo in cases:
        want = fix['vf_' + name]
        asm = cls(kvs, geo)
        assert asm.dev.fast_path, name                       # the sum-factorised walks, not the per-entry fall-back
        got = assemble.assemble(asm, symmetric=False).toarray()
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (name, np.abs(got - want).max())
        a, b = be.to_host(asm.dev.assemble_mlb()), be.to_host(asm.dev.assemble_mlb(entrywise=True))
        assert np.abs(a - b).max() <= RTOL * np.abs(b).max(), name
    # the predefined forms of the front end lead to the same classes
    got = assemble.assemble(vform.wave_st_vf(2), kv2, geo=geometry.quarter_annulus()).toarray()
    assert np.abs(got - fix['vf_wave_st2']).max() <= RTOL * np.abs(fix['vf_wave_st2']).max()
    got = assemble.assemble(vform.heat_st_vf(3), kv3, geo=geometry.twisted_box()).toarray()
    assert np.abs(got - fix['vf_heat_st3']).max() <= RTOL * np.abs(fix['vf_heat_st3']).max()
    # a larger 3D case, fast walks against the per-entry kernel (different algorithms, same tables)
    kvs = (bspline.make_knots(3, 0.0, 1.0, 7), bspline.make_knots(2, 0.0, 1.0, 9), bspline.make_knots(3, 0.0, 1.0, 6))
    asm = assemblers.WaveAssembler_ST3D(kvs, geometry.twisted_box())
    a, b = be.to_host(asm.dev.assemble_mlb()), be.to_host(asm.dev.assemble_mlb(entrywise=True))
    assert asm.dev.fast_path and np.abs(a - b).max() <= RTOL * np.abs(b).max()


def check_hierarchical_discretization(monkeypatch):
    """the reference's HDiscretization (pyiga/_hdiscr.py:37-57) assembles HB / THB-spline matrices level by
    level with on_demand assemblers inside bounding boxes; with its compile_vform replaced by the device
    backend the matrices equal the reference's own (test/test_hierarchical.py:180-215 pattern)"""
    _import_reference()
    from pyiga import _hdiscr, geometry as rgeo
    from pyiga_b200 import vform as dvform
    rc, fix = _vform_fixture()
    geo = rgeo.bspline_quarter_annulus()
    for name, (make, inputs, sym) in rc.hcases().items():
        for truncate in (False, True):
            want = fix['h_%s_%d' % (name, truncate)]
            calls = []

            def device_compile(vf, on_demand=False):
                calls.append(on_demand)
                return dvform.compile_vform(vf, on_demand=on_demand)
            with monkeypatch.context() as mp:
                mp.setattr(_hdiscr.compile, 'compile_vform', device_compile)
                hs = rc.hspace(truncate)
                got = _hdiscr.HDiscretization(hs, make(), dict(inputs, geo=geo)).assemble_matrix(symmetric=sym).toarray()
            assert calls and all(calls), 'HDiscretization must have asked for on_demand assemblers'
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), (name, truncate, np.abs(got - want).max())


def check_entry_func_ptr():
    """the "entryfunc" capsule (pyiga/genericasm.pxi:780-786): the reference's low-rank (ACA) assembler
    pyiga.fast_assemble_cy.fast_assemble pulls single entries of a device assembler through it"""
    import ctypes as C
    import pytest
    _import_reference()
    from pyiga_b200 import assemblers, bspline, geometry
    kvs = (bspline.make_knots(2, 0.0, 1.0, 6), bspline.make_knots(3, 0.0, 1.0, 5))
    asm = assemblers.StiffnessAssembler2D(kvs, geometry.bspline_quarter_annulus())
    A = asm.assemble_csr()
    cap = asm.entry_func_ptr()
    get = C.pythonapi.PyCapsule_GetPointer
    get.restype, get.argtypes = C.c_void_p, [C.py_object, C.c_char_p]
    fn = C.CFUNCTYPE(C.c_double, C.c_size_t, C.c_size_t, C.c_void_p)(get(cap, b'entryfunc'))
    for i, j in ((0, 0), (5, 6), (17, 3), (A.shape[0] - 1, A.shape[0] - 2)):
        assert fn(i, j, None) == A[i, j]
    try:
        from pyiga import fast_assemble_cy
    except Exception as exc:        # pragma: no cover
        pytest.skip('fast_assemble_cy of the reference does not import: %s' % exc)
    from pyiga import bspline as rbs
    rkvs = tuple(rbs.KnotVector(kv.kv, kv.p) for kv in kvs)
    B = fast_assemble_cy.fast_assemble(asm, rkvs, tol=1e-12, maxiter=200, verbose=0)
    assert abs(B - A).max() <= 1e-9 * abs(A).max()


def check_device_callables(ref):
    """Coefficient callables written against numpy (np.sin, np.where, ...) run on device tensors behind
    the numpy protocols of pyiga_b200._devarray.DevArray and give what the host evaluation of the
    reference gives (pyiga/utils.py:8-52)."""
    import torch
    from pyiga_b200 import _device, assemble
    from pyiga_b200._devarray import DevArray, unwrap
    from pyiga_b200.vform import _grid_values
    on_gpu = _device.backend().name == 'cuda'
    devname = 'cuda' if on_gpu else 'cpu'
    rng = np.random.default_rng(5)
    shape = (6, 5, 7)
    X = [rng.uniform(0.1, 1.9, shape) for _ in range(3)]
    Xd = [torch.as_tensor(x, device=devname) for x in X]
    funcs = [
        ((), lambda x, y, z: np.sin(x) * np.exp(-y) + np.sqrt(z) / (1.0 + x * x)),
        ((), lambda x, y, z: np.where(x > 1.0, 2.0, 0.5) + np.maximum(y, z) - np.minimum(x, 1.2)),
        ((), lambda x, y, z: np.where((x > 0.5) & (y < 1.5), x, -y) + abs(z - 1.0) ** 1.5),
        ((), lambda x, y, z: np.arctan2(y, x) + np.hypot(x, z) + np.log(1.0 + y) + 2.0 ** x + np.float64(3.0) * z),
        ((), lambda x, y, z: np.cos(np.pi * x) * np.tanh(y) + np.clip(z, 0.5, 1.0) + np.zeros_like(x) + np.ones_like(y)),
        ((), lambda x, y, z: 1.0 + 0 * x),
        ((), lambda x, y, z: 3.5),
        ((3,), lambda x, y, z: (np.sin(y), -x * z, 1.0)),
        ((3,), lambda x, y, z: np.stack([np.cos(x), y * y, np.full_like(z, 2.0)])),
        ((2, 2), lambda x, y, z: ((1.0 + x, np.exp(z)), (np.abs(y - 1.0), 0.0))),
    ]
    for shp, f in funcs:
        seen = []

        def spy(*c, f=f):
            seen.append(type(c[0]).__name__)
            return f(*c)
        got = _grid_values(spy, shp, tuple(Xd), shape)
        want = _grid_values(f, shp, tuple(X), shape)
        assert seen == ['DevArray'], seen          # no fall-back to raw tensors
        if shp == ():
            assert tuple(got.shape) == shape and got.dtype == torch.float64
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-14, atol=1e-15)
        else:
            for idx in np.ndindex(*shp):
                assert tuple(got[idx].shape) == shape and got[idx].dtype == torch.float64
                np.testing.assert_allclose(got[idx].cpu().numpy(), want[idx], rtol=1e-14, atol=1e-15)
    # what the wrapper does not know raises (the caller then evaluates on the host)
    import math
    a = DevArray(Xd[0])
    for bad in (lambda: math.sin(a), lambda: np.linalg.norm(a), lambda: np.asarray(a), lambda: bool(a > 1.0),
                lambda: np.add.reduce(a)):
        with pytest.raises(Exception):
            bad()
    assert unwrap((a, [a, 1.0]))[1][0] is Xd[0]

    # the same through an assembled form: numpy callable (device tensors on the GPU) vs a callable that
    # can only run on the host
    kvs = make_space(ref, 'a3_tb')
    geo = make_geo(ref, 'tb')
    fdev = lambda x, y, z: np.sin(x + y) * np.exp(-z) + np.where(x > 0.5, 2.0, 1.0)
    types = []

    def fspy(x, y, z):
        types.append(type(x).__name__)
        return fdev(x, y, z)

    def fhost(x, y, z):
        return fdev(np.asarray(x), np.asarray(y), np.asarray(z))      # np.asarray of a device array raises
    bvec = lambda x, y, z: (np.cos(y), x * z, 1.0 + 0 * x)
    form = 'f * inner(grad(u), grad(v)) * dx + inner(b, grad(u)) * v * dx'
    A = assemble.assemble(form, kvs, geo=geo, f=fspy, b=bvec)
    B = assemble.assemble(form, kvs, geo=geo, f=fhost, b=lambda x, y, z: bvec(np.asarray(x), np.asarray(y), np.asarray(z)))
    if on_gpu:
        assert 'DevArray' in types and 'ndarray' not in types, types      # (a scalar probe of the output shape comes first)
    assert abs(A - B).max() <= RTOL * abs(B).max()


def check_forms_1d():
    """Forms over ONE knot vector (test/test_assemble.py:436-445) — strings through the package's front end, VForm
    objects of the reference through refvform — against the reference's JIT-compiled assemblers
    (tests/golden/make_golden_1d.py).  The device runs them on a lifted two-axis space (assemblers.py)."""
    _import_reference()
    from pyiga_b200 import assemble, bspline
    rc, _ = _vform_fixture()
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_forms_1d.npz'))
    for name, (problem, kvs, geo, inputs) in rc.cases1d().items():
        want = fix['f1_' + name]
        got = assemble.assemble(problem if isinstance(problem, str) else problem(), kvs, geo=geo, **inputs)
        got = got.toarray() if hasattr(got, 'toarray') else np.asarray(got)
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (name, np.abs(got - want).max())
    # the reference's own assertions: the forms equal the 1D helpers
    kv = bspline.make_knots(3, 0.0, 1.0, 7)
    from pyiga_b200 import geometry
    geo = geometry.unit_cube(dim=1)
    A1 = assemble.assemble('inner(grad(u), grad(v)) * dx', (kv,), geo=geo)
    assert np.allclose(A1.toarray(), assemble.stiffness(kv).toarray(), rtol=0, atol=1e-12 * abs(A1).max())
    f = lambda x: 1 + x ** 2
    f1 = assemble.assemble('f * v * dx', (kv,), geo=geo, f=f)
    assert np.allclose(f1, assemble.inner_products(kv, f=f, f_physical=True, geo=geo), rtol=0, atol=1e-14)
    # inner_products over a mapped interval, f in parameter and in physical coordinates, against the reference's function
    from pyiga import assemble as rasm
    _, rkvs, rgeo_c, _ = rc.cases1d()['o_stiff']
    g = lambda t: np.cos(2.0 * t) + t
    for phys in (False, True):
        want = rasm.inner_products(rkvs[0], g, f_physical=phys, geo=rgeo_c)
        got = assemble.inner_products(rkvs[0], g, f_physical=phys, geo=rgeo_c)
        assert got.shape == want.shape and np.abs(got - want).max() <= RTOL * np.abs(want).max(), phys
    # two spaces over one mesh (trial: space 0 = columns, test: space 1 = rows) vs the reference's 1D helper
    from pyiga import bspline as rbs
    A = assemble.assemble('Dx(u, 0) * v * dx', ((bspline.make_knots(3, 0.0, 1.0, 6),), (bspline.make_knots(2, 0.0, 1.0, 6),)),
                          geo=geometry.unit_cube(dim=1), bfuns=[('u', 1, 0), ('v', 1, 1)])
    B = rasm.bsp_mixed_deriv_biform_1d_asym(rbs.make_knots(3, 0.0, 1.0, 6), rbs.make_knots(2, 0.0, 1.0, 6), 1, 0)
    assert A.shape == B.shape == (8, 9) and abs(A - B).max() <= RTOL * abs(B).max()
    # assembler protocol on a form over one knot vector: entries, MLB format, updatable inputs
    problem, kvs, geo, inputs = rc.cases1d()['s_cd_curved']
    want = fix['f1_s_cd_curved']
    asm = assemble.Assembler(problem, kvs, geo=geo, updatable=['a'], **inputs)
    A = asm.assemble()
    assert np.abs(A.toarray() - want).max() <= RTOL * np.abs(want).max()
    ij = np.array([[0, 0], [3, 4], [4, 3], [10, 9], [0, 7]])
    vals = asm.asm.multi_entries(ij)
    assert np.abs(vals - want[ij[:, 0], ij[:, 1]]).max() <= RTOL * np.abs(want).max()
    assert abs(asm.asm.entry(3, 4) - want[3, 4]) <= RTOL * np.abs(want).max()
    M = asm.assemble(format='mlb')
    assert np.abs(M.asmatrix().toarray() - want).max() <= RTOL * np.abs(want).max()
    B = asm.assemble(a=lambda x: 2.0 + 2.0 * x * x).toarray()          # a -> 2a: only the diffusion term doubles
    K = assemble.assemble('a * inner(grad(u), grad(v)) * dx', kvs, geo=geo, a=inputs['a']).toarray()
    assert np.abs(B - (want + K)).max() <= 1e-11 * np.abs(want).max()


def check_reference_api_extras():
    """Pieces of the reference's API around the path that its own tests of the path use
    (test/test_assemble.py, test/test_geometry.py, test/test_mlmatrix.py), against the live reference:
    Hessians of spline geometries, cylinderize, module-level dx / ds, 1D bilinear helpers with second
    derivatives and weight functions, parametric callables as form inputs, one-level MLMatrix products."""
    _import_reference()
    from pyiga import assemble as rasm, bspline as rbs, geometry as rgeo
    from pyiga_b200 import assemble, assemblers, bspline, geometry, mlmatrix, vform
    rng = np.random.default_rng(3)
    # grid_hessian (pyiga/bspline.py:923-980, pyiga/geometry.py:125-150)
    for d in (2, 3):
        for comps in ((), (d,), (2,)):
            ps, ns = (2, 3, 2)[:d], (4, 3, 5)[:d]
            rk = tuple(rbs.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
            ok = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
            N = tuple(k.numdofs for k in rk)
            c, w = rng.standard_normal(N + comps), rng.uniform(0.5, 1.5, N)
            grid = [np.linspace(0.03, 0.97, 5 + k) for k in range(d)]
            for a, b in ((rbs.BSplineFunc(rk, c), bspline.BSplineFunc(ok, c)), (rgeo.NurbsFunc(rk, c, w), geometry.NurbsFunc(ok, c, w))):
                want, got = a.grid_hessian(grid), b.grid_hessian(grid)
                assert got.shape == want.shape and np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (d, comps)
    # a fourth-order form on a geometry object of THIS package (the Hessian of the geometry enters)
    rc, fix = _vform_fixture()
    make, kvs, _, inputs = rc.cases()['biharmonic2']
    got = assemble.assemble(make(), kvs, geo=geometry.quarter_annulus(), **inputs).toarray()
    want = fix['vf_biharmonic2']
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
    # cylinderize (pyiga/bspline.py:1097-1106)
    g, r = geometry.bspline_quarter_annulus().cylinderize(0.0, 2.0, support=(0.0, 2.0)), rgeo.bspline_quarter_annulus().cylinderize(0.0, 2.0, support=(0.0, 2.0))
    grid = [np.linspace(0, 2, 4), np.linspace(0, 1, 3), np.linspace(0, 1, 5)]
    assert (g.sdim, g.dim) == (3, 3) and np.abs(np.asarray(g.grid_eval(grid)) - r.grid_eval(grid)).max() <= 1e-14
    assert np.abs(np.asarray(g.grid_jacobian(grid)) - r.grid_jacobian(grid)).max() <= 1e-13
    # the reference's own space-time test of the wave assembler (test/test_assemble.py:118-131)
    import scipy.sparse
    T_end = 2.0
    geo = geometry.unit_cube(dim=1).cylinderize(0.0, T_end, support=(0.0, T_end))
    kv_t, kv = bspline.make_knots(2, 0.0, T_end, 6), bspline.make_knots(3, 0.0, 1.0, 8)
    D0Dt, DttDt = assemble.bsp_mixed_deriv_biform_1d(kv_t, 0, 1), assemble.bsp_mixed_deriv_biform_1d(kv_t, 2, 1)
    A_ref = (scipy.sparse.kron(DttDt, assemble.mass(kv)) + scipy.sparse.kron(D0Dt, assemble.stiffness(kv))).tocsr()
    A = assemble.assemble_entries(assemblers.WaveAssembler_ST2D((kv_t, kv), geo))
    assert abs(A_ref - A).max() < 1e-12
    # 1D bilinear helpers: derivative orders up to 2, weight functions (pyiga/assemble.py:179-222)
    rkv, okv = rbs.make_knots(4, 0.0, 1.0, 9), bspline.make_knots(4, 0.0, 1.0, 9)
    wf = lambda x: 1.0 + np.sin(3 * x) ** 2
    for du, dv in ((2, 0), (2, 1), (2, 2), (0, 2), (1, 0), (0, 0)):
        for weight in (None, wf):
            # with a weight the quadrature rule matters: both sides take the (p+1)-node rule of the device tables
            nq = 5 if weight is not None else None
            want = rasm.bsp_mixed_deriv_biform_1d(rkv, du, dv, nqp=nq, weightfunc=weight).toarray()
            got = assemble.bsp_mixed_deriv_biform_1d(okv, du, dv, nqp=nq, weightfunc=weight).toarray()
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (du, dv, weight is not None)
    rkv2, okv2 = rbs.make_knots(2, 0.0, 1.0, 9), bspline.make_knots(2, 0.0, 1.0, 9)
    want = rasm.bsp_mixed_deriv_biform_1d_asym(rkv, rkv2, 2, 1).toarray()
    got = assemble.bsp_mixed_deriv_biform_1d_asym(okv, okv2, 2, 1, quadgrid=okv.mesh).toarray()
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # module-level measures and a parametric plain callable as input (test/test_assemble.py:287-312)
    kvs = 2 * (bspline.make_knots(3, 0.0, 1.0, 6),)
    qa = geometry.quarter_annulus()
    vf = vform.VForm(2)
    u, v = vf.basisfuns()
    vf.add(vform.inner(vform.grad(u), vform.grad(v)) * vform.dx)
    K = assemble.assemble_vf(vf, kvs, geo=qa)
    K2 = assemble.stiffness(kvs, qa)
    assert abs(K - K2).max() <= 1e-12 * abs(K2).max()
    vf_f = vform.VForm(2, arity=1)
    f = vf_f.input('f')
    v = vf_f.basisfuns()
    vf_f.add(f * v * vform.dx)
    fun = lambda x, y: np.exp(x + y)
    f1 = assemble.assemble_vf(vf_f, kvs, geo=qa, f=fun)
    f2 = assemble.inner_products(kvs, fun, geo=qa)
    rk = 2 * (rbs.make_knots(3, 0.0, 1.0, 6),)
    f3 = rasm.inner_products(rk, fun, geo=rgeo.quarter_annulus())
    assert np.abs(f1 - f3).max() <= 1e-13 and np.abs(f2 - f3).max() <= 1e-13
    # vector- and tensor-valued spline functions: values / Jacobians of more than 3 components, load vectors per component
    # (test/test_approx.py:5-17 with degrees the device tables cover)
    rk = tuple(rbs.make_knots(p, 0.0, 1.0, 4 + p) for p in (2, 3, 4))
    ok = tuple(bspline.make_knots(p, 0.0, 1.0, 4 + p) for p in (2, 3, 4))
    N = tuple(k.numdofs for k in rk)
    grid = [np.linspace(0.0, 1.0, 4 + k) for k in range(3)]
    for extra in ((3,), (5,), (2, 2)):
        c = rng.standard_normal(N + extra)
        fr, fo = rbs.BSplineFunc(rk, c), bspline.BSplineFunc(ok, c)
        assert np.abs(np.asarray(fo.grid_eval(grid)) - fr.grid_eval(grid)).max() <= 1e-13
        if len(extra) == 1:
            assert np.abs(np.asarray(fo.grid_jacobian(grid)) - fr.grid_jacobian(grid)).max() <= 1e-12
        want, got = rasm.inner_products(rk, fr), assemble.inner_products(ok, fo)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), extra
    # one-level MLMatrix: matrix and product (test/test_mlmatrix.py:60-70)
    S = mlmatrix.MLStructure.multi_banded((20,), (3,))
    A = np.zeros((20, 20))
    for i in range(20):
        for j in range(max(0, i - 3), min(20, i + 4)):
            A[i, j] = rng.standard_normal()
    X = mlmatrix.MLMatrix(structure=S, matrix=A)
    x = rng.standard_normal(20)
    assert np.allclose(A, X.asmatrix().toarray()) and np.allclose(A @ x, X.dot(x), rtol=0, atol=1e-13)


def check_high_degree():
    """Degrees without instantiated sum-factorisation kernels (p = 5): matrices come from the per-entry quadrature
    kernel, load vectors from the row sums of the twin bilinear form (DeviceAssembler.assemble_vector_device);
    against the live reference (test/test_assemble.py:223-246 uses degrees 3, 4, 5)."""
    _import_reference()
    from pyiga import assemble as rasm, bspline as rbs, geometry as rgeo
    from pyiga_b200 import assemble, assemblers, bspline, geometry
    ps, ns = (3, 5, 4), (2, 2, 3)
    kvs = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
    rk = tuple(rbs.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
    f = lambda x, y, z: np.cos(x) * np.exp(y) * np.sin(z)
    geo, rg = geometry.twisted_box(), rgeo.twisted_box()
    for kw in ({}, {'geo': (geo, rg)}, {'geo': (geo, rg), 'f_physical': True}):
        kw_o = {k: (v[0] if k == 'geo' else v) for k, v in kw.items()}
        kw_r = {k: (v[1] if k == 'geo' else v) for k, v in kw.items()}
        want, got = rasm.inner_products(rk, f, **kw_r), assemble.inner_products(kvs, f, **kw_o)
        assert got.shape == want.shape and np.abs(got - want).max() <= RTOL * np.abs(want).max(), kw
    asm = assemblers.L2FunctionalAssemblerPhys3D(kvs, geo, f=f)
    assert not asm.dev.fast_path
    assert np.abs(asm.assemble_vector() - rasm.inner_products(rk, f, f_physical=True, geo=rg)).max() <= 1e-13
    for form in ('mass', 'stiffness'):
        A, B = getattr(assemble, form)(kvs, geo), getattr(rasm, form)(rk, rg)
        assert abs(A - B).max() <= 1e-12 * abs(B).max(), form
    # 2D, degree 6
    kv2, rk2 = 2 * (bspline.make_knots(6, 0.0, 1.0, 3),), 2 * (rbs.make_knots(6, 0.0, 1.0, 3),)
    g2 = lambda x, y: x * y + 1.0
    want = rasm.inner_products(rk2, g2, f_physical=True, geo=rgeo.quarter_annulus())
    got = assemble.inner_products(kv2, g2, f_physical=True, geo=geometry.quarter_annulus())
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
