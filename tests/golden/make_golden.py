"""Generates the committed golden fixtures from the REAL reference (c-f-h/pyiga).

Run in the build container, where the reference is installed under oracle/_ref (see
oracle/build_ref.py) and its test data lives under /root/reference/test:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
    pyiga_<name>.npz    the reference's own golden matrices test/poisson_neu_*.mtx.gz
                        (test/test_assemble.py:138-168), re-encoded as CSR arrays
    ref_cases.npz       inputs + outputs of the reference run on small cases: basis tables,
                        Jacobians, band structures, mass/stiffness matrices, multi_entries samples,
                        MLMatrix matvec, Kronecker matvec
The GPU box has no /root/reference; tests only read these files.
"""
import gzip
import os
import sys

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))

from pyiga import assemble, assemblers, bspline, geometry, mlmatrix, operators  # noqa: E402

REFTEST = '/root/reference/test'


def read_mtx(path):
    with gzip.open(path, 'rt') as f:
        m, n, nnz = (int(t) for t in f.readline().split())
        raw = np.loadtxt(f)
    A = scipy.sparse.coo_matrix((raw[:, 2], (raw[:, 0].astype(int) - 1, raw[:, 1].astype(int) - 1)), shape=(m, n))
    return A.tocsr()


def save_csr(name, A):
    A = A.tocsr()
    A.sort_indices()
    np.savez_compressed(os.path.join(HERE, name), data=A.data, indices=A.indices.astype(np.int32),
                        indptr=A.indptr.astype(np.int32), shape=np.array(A.shape))


def save_csr_into(out, name, A):
    A = scipy.sparse.csr_matrix(A)
    A.sort_indices()
    out[name + '_indptr'], out[name + '_indices'], out[name + '_data'] = A.indptr, A.indices, A.data
    out[name + '_shape'] = np.array(A.shape)


def geo_pack(prefix, geo, out):
    out[prefix + '_rational'] = np.array(int(isinstance(geo, geometry.NurbsFunc)))
    out[prefix + '_coeffs'] = np.asarray(geo.coeffs, dtype=float)
    for k, kv in enumerate(geo.kvs):
        out['%s_kv%d' % (prefix, k)] = kv.kv
        out['%s_p%d' % (prefix, k)] = np.array(kv.p)
    out[prefix + '_sdim'] = np.array(geo.sdim)


def twisted_nurbs_box():
    G = geometry.twisted_box()
    i, j, k = np.meshgrid(np.arange(2), np.arange(4), np.arange(2), indexing='ij')
    W = 1.0 + 0.25 * ((i + 2 * j + 3 * k) % 3)
    return geometry.NurbsFunc(G.kvs, G.coeffs.copy(), W)


def main():
    # ---- 1. the reference's golden matrices ----------------------------------------------------
    for name in ('d2_p3_n15_mass', 'd2_p3_n15_stiff', 'd3_p2_n10_mass', 'd3_p2_n10_stiff'):
        save_csr('pyiga_%s.npz' % name, read_mtx(os.path.join(REFTEST, 'poisson_neu_%s.mtx.gz' % name)))

    out = {}
    # ---- 2. basis functions ----------------------------------------------------------------------
    kv = bspline.make_knots(3, 0.0, 1.0, 5, mult=2)
    nodes = np.concatenate((np.linspace(0, 1, 23), kv.mesh))
    idx, vals = bspline.collocation_derivs_info(kv, nodes, derivs=2)
    out['basis_kv'], out['basis_p'], out['basis_nodes'] = kv.kv, np.array(3), nodes
    out['basis_first'], out['basis_vals'] = np.asarray(idx), np.asarray(vals)      # (derivs+1, n, p+1)

    # ---- 3. geometry Jacobians ------------------------------------------------------------------
    geos = {'tb': geometry.twisted_box(), 'tnb': twisted_nurbs_box(), 'qa': geometry.quarter_annulus(),
            'bqa': geometry.bspline_quarter_annulus(),
            'cyl': geometry.tensor_product(geometry.line_segment(0, 1), geometry.quarter_annulus())}
    for name, geo in geos.items():
        grid = tuple(np.linspace(0.03, 0.98, 4 + k) for k in range(geo.sdim))
        geo_pack('geo_' + name, geo, out)
        for k, g in enumerate(grid):
            out['geo_%s_grid%d' % (name, k)] = g
        out['geo_%s_val' % name] = geo.grid_eval(grid)
        out['geo_%s_jac' % name] = geo.grid_jacobian(grid)

    # ---- 4. assembled matrices --------------------------------------------------------------------
    def space(ps, ns, mult=1):
        return tuple(bspline.make_knots(p, 0.0, 1.0, n, mult=mult) for p, n in zip(ps, ns))

    cases = {
        'a2_qa': (space((2, 2), (4, 5)), geos['qa']),
        'a2_mixed': (space((4, 3), (4, 5)), geos['bqa']),
        'a3_tb': (space((2, 2, 2), (3, 4, 3)), geos['tb']),
        'a3_mixed': (space((2, 3, 2), (4, 3, 3)), geos['tb']),
        'a3_nurbs': (space((3, 3, 3), (4, 4, 4)), geos['tnb']),
        'a3_mult': (space((2, 2, 2), (3, 3, 3), mult=2), geos['cyl']),
        'a3_p1': (space((1, 1, 1), (4, 3, 5)), geos['tb']),
        'a3_p4': (space((4, 4, 4), (2, 3, 2)), geos['tnb']),
    }
    rng = np.random.default_rng(1234)
    for name, (kvs, geo) in cases.items():
        out[name + '_geo'] = np.array([k for k, g in geos.items() if g is geo][0])
        for k, kv in enumerate(kvs):
            out['%s_kv%d' % (name, k)] = kv.kv
            out['%s_p%d' % (name, k)] = np.array(kv.p)
        out[name + '_dim'] = np.array(len(kvs))
        S = mlmatrix.MLStructure.from_kvs(kvs, kvs)
        for k in range(S.L):
            out['%s_bidx%d' % (name, k)] = S.bidx[k]
        for form, fn in (('mass', assemble.mass), ('stiff', assemble.stiffness)):
            A = fn(kvs, geo).tocsr()
            A.sort_indices()
            I, J = S.nonzero()
            data = np.asarray(A[I.astype(np.int64), J.astype(np.int64)]).ravel()
            out['%s_%s_mlb' % (name, form)] = data.reshape([len(b) for b in S.bidx])
            out['%s_%s_indptr' % (name, form)] = A.indptr
            out['%s_%s_indices' % (name, form)] = A.indices
        # multi_entries protocol: random pairs, many outside the pattern
        cls = {2: assemblers.StiffnessAssembler2D, 3: assemblers.StiffnessAssembler3D}[len(kvs)]
        asm = cls(kvs, geo)
        n = S.shape[0]
        ij = np.column_stack((rng.integers(0, n, 200), rng.integers(0, n, 200)))
        I, J = S.nonzero()
        pick = rng.integers(0, len(I), 200)
        ij = np.vstack((ij, np.column_stack((I[pick], J[pick])))).astype(np.uint64)
        out[name + '_me_ij'] = ij
        out[name + '_me_val'] = asm.multi_entries(ij)

    # ---- 5. MLMatrix matvec and Kronecker operator ----------------------------------------------
    kvs, geo = cases['a3_mixed']
    S = mlmatrix.MLStructure.from_kvs(kvs, kvs)
    X = rng.standard_normal([len(b) for b in S.bidx])
    x = rng.standard_normal(S.shape[1])
    out['mv_data'], out['mv_x'] = X, x
    out['mv_y'] = mlmatrix.MLMatrix(S, data=X).dot(x)
    facs = [rng.standard_normal((3, 4)), rng.standard_normal((5, 2)), rng.standard_normal((4, 6))]
    xk = rng.standard_normal(4 * 2 * 6)
    for k, A in enumerate(facs):
        out['kron_A%d' % k] = A
    out['kron_x'] = xk
    out['kron_y'] = operators.KroneckerOperator(*facs).dot(xk)

    # ---- 6. string vforms through the reference's JIT (pyiga/assemble.py:837, compile.py:120) -------
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import VFORMS
    for name, (form, inputs, case, gname) in VFORMS.items():
        kvs, _ = cases[case]
        A = assemble.assemble(form, kvs, geo=geos[gname], **inputs).tocsr()
        A.sort_indices()
        S = mlmatrix.MLStructure.from_kvs(kvs, kvs)
        I, J = S.nonzero()
        out['vf_%s_mlb' % name] = np.asarray(A[I.astype(np.int64), J.astype(np.int64)]).ravel().reshape(
            [len(b) for b in S.bidx])
    # Assembler with an updatable input (test/test_assemble.py:409-450 pattern)
    kvs, _ = cases['a2_qa']
    asm = assemble.Assembler('f * inner(grad(u), grad(v)) * dx', kvs, geo=geos['qa'], f=lambda x, y: 1.0 + x,
                             updatable=['f'])
    S = mlmatrix.MLStructure.from_kvs(kvs, kvs)
    I, J = S.nonzero()
    out['vf_upd_a'] = np.asarray(asm.assemble()[I.astype(np.int64), J.astype(np.int64)]).ravel()
    out['vf_upd_b'] = np.asarray(asm.assemble(f=lambda x, y: 2.0 + y * y)[I.astype(np.int64), J.astype(np.int64)]).ravel()

    # ---- 7. linear forms: assemble_vector / inner_products (test/test_assemble.py:223-245) --------
    from helpers import LFORMS
    for name, (form, inputs, case, gname) in LFORMS.items():
        kvs, _ = cases[case]
        out['lf_%s' % name] = assemble.assemble(form, kvs, geo=geos[gname], **inputs)
    kvs, _ = cases['a3_tb']
    fpar = lambda x, y, z: x + 2 * y * z
    out['ip_param'] = assemble.inner_products(kvs, fpar, geo=geos['tb'])
    out['ip_phys'] = assemble.inner_products(kvs, fpar, f_physical=True, geo=geos['tnb'])
    out['ip_nogeo'] = assemble.inner_products(kvs, fpar)

    # ---- 8. vector-valued forms (test/test_assemble.py:170-181, 452-472) ----------------------------
    from helpers import VECFORMS
    for name, (form, bfuns, inputs, case, gname) in VECFORMS.items():
        kvs, _ = cases[case]
        X = assemble.assemble(form, kvs, geo=geos[gname], bfuns=bfuns, format='mlb', layout='packed', **inputs)
        out['vv_%s_mlb' % name] = X.data
        asm = assemble.instantiate_assembler(form, kvs, dict(inputs, geo=geos[gname]), bfuns)
        out['vv_%s_blocks' % name] = np.array(asm.multi_blocks([(0, 0), (0, 1), (2, 1), (5, 5)]))
    kvs, _ = cases['a2_qa']
    out['vv_divdiv2_bsr'] = assemble.divdiv(kvs, geos['bqa'], layout='packed', format='bsr').toarray()
    out['vv_rhs2'] = assemble.assemble('inner(g, v) * dx', kvs, geo=geos['qa'], bfuns=[('v', 2)], g=lambda x, y: (x, -y))

    # ---- 9. essential boundary conditions (test/test_assemble.py:251-281,497-505, test/test_solve.py) ---
    from pyiga import approx
    kvs2 = (bspline.make_knots(3, 0.0, 1.0, 5), bspline.make_knots(2, 0.0, 1.0, 8))
    out['bc_bd_bottom'] = assemble.boundary_dofs(kvs2, 'bottom', ravel=True)
    out['bc_bd_right'] = assemble.boundary_dofs(kvs2, 'right')
    out['bc_bd_left_flip'] = assemble.boundary_dofs(kvs2, 'left', ravel=True, flip=(True,))
    kvs3b = (bspline.make_knots(2, 0.0, 1.0, 3), bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 2))
    out['bc_bd3_front'] = assemble.boundary_dofs(kvs3b, 'front', ravel=True)
    out['bc_bd3_top'] = assemble.boundary_dofs(kvs3b, (1, 1))
    out['bc_cells3_back'] = assemble.boundary_cells(kvs3b, 'back', ravel=True)
    # Poisson on the NURBS quarter annulus, Dirichlet data on all sides (test/test_solve.py)
    kvsP = 2 * (bspline.make_knots(3, 0.0, 1.0, 10),)
    geoP = geometry.quarter_annulus()
    gP = lambda x, y: np.cos(x + y) + np.exp(y - x)
    fP = lambda x, y: 2 * (np.cos(x + y) - np.exp(y - x))
    idx, val = assemble.compute_dirichlet_bcs(kvsP, geoP, ('all', gP))
    out['bc_p2_idx'], out['bc_p2_val'] = idx, val
    i1, v1 = assemble.compute_dirichlet_bc(kvsP, geoP, 'top', gP)
    out['bc_p2_top_idx'], out['bc_p2_top_val'] = i1, v1
    rhs = assemble.inner_products(kvsP, fP, f_physical=True, geo=geoP).ravel()
    A = assemble.stiffness(kvsP, geo=geoP)
    LS = assemble.RestrictedLinearSystem(A, rhs, (idx, val))
    save_csr_into(out, 'bc_p2_A', LS.A)
    out['bc_p2_b'] = LS.b
    u = LS.complete(scipy.sparse.linalg.spsolve(LS.A.tocsc(), LS.b))
    out['bc_p2_u'] = u
    out['bc_p2_uex'] = approx.interpolate(kvsP, gP, geo=geoP)
    # 3D, two sides, vector-valued Dirichlet data for a blocked 3-component system, elim_rows
    kvs3, _ = cases['a3_tb']
    g3 = geos['tnb']
    i3, v3 = assemble.compute_dirichlet_bcs(kvs3, g3, [('front', lambda x, y, z: x * y + z), ((2, 1), 1.5)])
    out['bc_3d_idx'], out['bc_3d_val'] = i3, v3
    iv, vv = assemble.compute_dirichlet_bc(kvs3, g3, 'bottom', lambda x, y, z: (x, y * z, 1.0 + z))
    out['bc_3d_vec_idx'], out['bc_3d_vec_val'] = iv, vv
    A3 = assemble.stiffness(kvs3, geo=g3) + assemble.mass(kvs3, geo=g3)
    b3 = np.cos(np.arange(A3.shape[0]) * 0.37)
    LS3 = assemble.RestrictedLinearSystem(A3, b3, (i3, v3))
    save_csr_into(out, 'bc_3d_A', LS3.A)
    out['bc_3d_b'] = LS3.b
    elim_rows = np.arange(3, A3.shape[0], 7)[:i3.size]
    LS3e = assemble.RestrictedLinearSystem(A3, 0.5, (i3, 2.0), elim_rows=elim_rows)
    save_csr_into(out, 'bc_3de_A', LS3e.A)
    out['bc_3de_b'], out['bc_3de_rows'] = LS3e.b, elim_rows
    xf = np.sin(np.arange(LS3.A.shape[0]) * 0.11)
    out['bc_3d_complete'] = LS3.complete(xf)
    out['bc_3d_extend'] = LS3.extend(xf)
    out['bc_3d_restrict'] = LS3.restrict(b3)
    save_csr_into(out, 'bc_3d_restrM', LS3.restrict_matrix(assemble.mass(kvs3, geo=g3)))
    # 1D model problem (test/test_assemble.py:497-505)
    kv1 = bspline.make_knots(2, 0.0, 1.0, 10)
    out['bc_1d_f'] = assemble.inner_products(kv1, lambda x: 1.0 + x)
    out['bc_1d_interp'] = approx.interpolate(kv1, lambda x: 0.5 * x * (3 - x))
    out['bc_interp3'] = approx.interpolate(kvs3, lambda x, y, z: np.sin(x) * y + z * z, geo=g3)

    # ---- 10. partial-row assembly (pyiga/_hdiscr.py:5-12) ----------------------------------------------
    from pyiga import _hdiscr
    for case, cls in [('a3_mixed', assemblers.StiffnessAssembler3D), ('a2_mixed', assemblers.MassAssembler2D)]:
        kvs, geo = cases[case]
        asm = cls(kvs, geo)
        n = int(np.prod([kv.numdofs for kv in kvs]))
        rows = np.unique((np.arange(17) * 7919 + 3) % n)
        out['pr_%s_rows' % case] = rows
        save_csr_into(out, 'pr_%s' % case, _hdiscr._assemble_partial_rows(asm, rows))

    # ---- 11. boundary integrals (pyiga/assemble.py:899-940, codegen/cython.py:549-590) -----------------
    from helpers import BFORMS
    for name, (form, bfuns, inputs, case, gname, bd) in BFORMS.items():
        kvs, _ = cases[case]
        R = assemble.assemble(form, kvs, geo=geos[gname], bfuns=bfuns, boundary=bd, **inputs)
        if scipy.sparse.issparse(R):
            save_csr_into(out, 'bf_%s' % name, R)
        else:
            out['bf_%s' % name] = np.asarray(R)

    # ---- 12. two spaces (Petrov-Galerkin), pyiga/assemble.py:947-951 -----------------------------------
    from helpers import PGFORMS
    for name, (form, bfuns, inputs, p0, p1, ns, gname) in PGFORMS.items():
        kvs0 = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(p0, ns))
        kvs1 = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(p1, ns))
        save_csr_into(out, 'pg_%s' % name, assemble.assemble(form, (kvs0, kvs1), geo=geos[gname], bfuns=bfuns, **inputs))

    # ---- 13. integrate (pyiga/assemble.py:658-696; test/test_assemble.py:483-495) -----------------------
    kvsI = (bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 5))
    out['int_qa_one'] = np.array(assemble.integrate(kvsI, lambda x, y: 1.0, geo=geometry.quarter_annulus()))
    out['int_qa_phys'] = np.array(assemble.integrate(kvsI, lambda x, y: x * y + np.cos(x), f_physical=True, geo=geometry.quarter_annulus()))
    out['int_par'] = np.array(assemble.integrate(kvsI, lambda x, y: x * x + y))
    out['int_vec'] = np.array(assemble.integrate(kvsI, lambda x, y: (x, y * x)))     # vector f: parameter domain only in the reference
    kvs3, _ = cases['a3_tb']
    out['int_3d'] = np.array(assemble.integrate(kvs3, lambda x, y, z: x + y * z, f_physical=True, geo=geos['tnb']))
    out['int_1d'] = np.array(assemble.integrate(bspline.make_knots(3, 0.0, 2.0, 5), lambda x: x * x))

    # ---- 14. surface integrals over manifolds (test/test_assemble.py:314-334) --------------------------
    from helpers import SFORMS
    cyl3 = geometry.tensor_product(geometry.line_segment(0.0, 1.0), geometry.quarter_annulus())
    for name, (form, bfuns, inputs, ps, ns, side) in SFORMS.items():
        kvsS = tuple(bspline.make_knots(p, 0.0, 1.0, n) for p, n in zip(ps, ns))
        R = assemble.assemble(form, kvsS, geo=cyl3.boundary(side), bfuns=bfuns, **inputs)
        if scipy.sparse.issparse(R):
            save_csr_into(out, 'sf_%s' % name, R)
        else:
            out['sf_%s' % name] = np.asarray(R)

    # ---- 15. 1D helpers (pyiga/assemble.py:165-230) -------------------------------------------------------
    kvA, kvB = bspline.make_knots(3, 0.0, 2.0, 7), bspline.make_knots(2, 0.0, 2.0, 7)
    for du, dv in [(0, 1), (1, 0), (1, 1), (0, 0)]:
        out['b1d_%d%d' % (du, dv)] = assemble.bsp_mixed_deriv_biform_1d(kvA, du, dv).toarray()
        out['b1d_asym_%d%d' % (du, dv)] = assemble.bsp_mixed_deriv_biform_1d_asym(kvA, kvB, du, dv).toarray()

    # ---- 16. space-time initial conditions (pyiga/assemble.py:492-552) -----------------------------------
    kvsT = (bspline.make_knots(2, 0.0, 1.0, 4), bspline.make_knots(3, 0.0, 1.0, 3), bspline.make_knots(2, 0.0, 1.0, 5))
    geoT = geometry.tensor_product(geometry.line_segment(0.0, 1.0), geometry.quarter_annulus())
    for side in (0, 1):
        i01, v01 = assemble.compute_initial_condition_01(kvsT, geoT, (0, side), lambda x, y, t: x * y, lambda x, y, t: x - y)
        out['ic01_idx%d' % side], out['ic01_val%d' % side] = i01, v01

    # ---- 17. L2 projection (pyiga/approx.py:62-95; used by test/test_solve.py) ----------------------------
    out['pl2_2d'] = approx.project_L2(kvsP, gP, f_physical=True, geo=geoP)
    out['pl2_3d'] = approx.project_L2(kvs3, lambda x, y, z: np.sin(x) * y + z * z, f_physical=True, geo=g3)
    out['pl2_par'] = approx.project_L2(kvs3, lambda x, y, z: x * y - z)
    out['pl2_1d'] = approx.project_L2(kv1, lambda x: np.cos(3 * x))

    np.savez_compressed(os.path.join(HERE, 'ref_cases.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
