"""Parity of the CUDA path on the B200 against reference fixtures, goldens and the oracle.
Everything here calls through the C ABI of libpyiga_b200.so (ctypes) on cuda:0."""
import numpy as np
import pytest

import parity_checks as pc
from helpers import CASES, GEOS, assert_close_rel

pytestmark = pytest.mark.gpu


def test_library_loaded(cuda):
    assert cuda.name == 'cuda'
    assert cuda.lib.pb200_version() >= 100


def test_basis(cuda, ref):
    pc.check_basis(cuda, ref)


@pytest.mark.parametrize('name', GEOS)
def test_geometry(cuda, ref, name):
    pc.check_geometry(ref, name)


@pytest.mark.parametrize('case', ['a2_qa', 'a3_nurbs'])
@pytest.mark.parametrize('form', ['mass', 'stiffness'])
def test_fields(cuda, ref, case, form):
    pc.check_fields(ref, case, form)


@pytest.mark.parametrize('case', CASES)
def test_assemble_sum_factorised(cuda, ref, case):
    pc.check_case(ref, case)


@pytest.mark.parametrize('case', CASES)
def test_assemble_entrywise(cuda, ref, case):
    pc.check_case(ref, case, entrywise=True)


@pytest.mark.parametrize('case,nslabs', [('a2_qa', 2), ('a3_tb', 2), ('a3_nurbs', 4), ('a3_mult', 3), ('a3_p1', 8)])
def test_slabs(cuda, ref, case, nslabs):
    pc.check_slabs(ref, case, nslabs)


def test_chunked(cuda, ref):
    pc.check_chunked(ref, 'a3_p1')


@pytest.mark.parametrize('case', CASES)
def test_multi_entries(cuda, ref, case):
    pc.check_multi_entries(ref, case)


@pytest.mark.parametrize('dim', [2, 3])
def test_golden(cuda, dim):
    pc.check_golden(dim)


def test_operators(cuda, ref):
    pc.check_operators(ref)


@pytest.mark.parametrize('name', ['cd3', 'cd2', 'aniso3', 'react2', 'sqrt3'])
def test_vform(cuda, ref, name):
    pc.check_vform(ref, name)


def test_vform_protocol(cuda, ref):
    pc.check_vform_protocol(ref)


def test_vector_forms(cuda, ref):
    pc.check_vector_forms(ref)


def test_edge_cases(cuda, ref):
    pc.check_edge_cases(ref)


def test_linear_forms(cuda, ref):
    pc.check_linear_forms(ref)


def test_kronecker_path_and_1d(cuda, ref):
    pc.check_kronecker_path(ref)


def test_slab_operator_and_cg(cuda, ref):
    pc.check_slab_operator_and_cg(ref)


@pytest.mark.parametrize('force_walk', [False, True])
@pytest.mark.parametrize('ps,ns', [((2, 2), (4, 70)), ((3, 1), (5, 33)), ((3, 3), (40, 45))])
def test_long_last_axis_2d(cuda, ps, ns, force_walk):
    pc.check_vs_oracle(2, ps, ns, 'Stiffness', force_walk=force_walk)
    pc.check_vs_oracle(2, ps, ns, 'Mass', force_walk=force_walk)


def test_repeated_knots(cuda):
    pc.check_vs_oracle(2, (3, 3), (4, 20), 'Stiffness', mult=2)
    pc.check_vs_oracle(3, (2, 2, 2), (3, 2, 12), 'Mass', mult=2, geo_name='bspline')
    pc.check_vs_oracle(3, (3, 3, 3), (5, 4, 6), 'Stiffness', mult=3)


@pytest.mark.parametrize('p,n', [(2, 12), (3, 16), (4, 10)])
@pytest.mark.parametrize('form', ['Mass', 'Stiffness'])
def test_vs_oracle_3d(cuda, p, n, form):
    """sizes above the fixtures: sum-factorised kernels vs the oracle's closed form, and vs the
    per-entry kernel (two independent device algorithms)"""
    from oracle import pyiga_oracle as orc
    from pyiga_b200 import assemblers, bspline, geometry
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    asm = getattr(assemblers, form + 'Assembler3D')(kvs, geo)
    got = cuda.to_host(asm.dev.assemble_mlb())
    prob = orc.Problem([kv.kv for kv in kvs], [p] * 3, [kv.kv for kv in geo.kvs], [kv.p for kv in geo.kvs],
                       geo.coeffs, True)
    want = orc.assemble_mlb(prob, form.lower()).ravel()
    assert_close_rel(got, want, what='%s p=%d n=%d' % (form, p, n))
    ew = cuda.to_host(asm.dev.assemble_mlb(entrywise=True))
    assert_close_rel(ew, want, what='entrywise %s p=%d n=%d' % (form, p, n))


def test_properties_large(cuda):
    """size-independent properties at a size the oracle cannot hold: symmetry of the MLB tensor
    under the transposed band index, row sums of the stiffness matrix vanish (constants are in the
    kernel of the gradient), mass matrix sums to the volume."""
    from pyiga_b200 import assemblers, bspline, geometry
    p, n = 3, 40
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_box()
    K = assemblers.StiffnessAssembler3D(kvs, geo)
    M = assemblers.MassAssembler3D(kvs, geo)
    Kd, Md = K.assemble_mlb(), M.assemble_mlb()
    ones = np.ones(Kd.shape[1])
    scale = np.abs(Kd.data).max()
    assert np.abs(Kd.dot(ones)).max() <= 1e-10 * scale
    vol_mass = float(ones @ Md.dot(ones))
    vol_quad = float(M.dev.fields_host().sum())          # sum of W = GaussWeight * |det J|
    assert abs(vol_mass - vol_quad) <= 1e-12 * abs(vol_quad)
    tr = [np.lexsort((b[:, 0], b[:, 1])) for b in Kd.structure.bidx]
    T = Kd.data[np.ix_(*tr)]
    assert np.abs(T - Kd.data).max() <= 1e-12 * scale


def test_pipelined_csr_host(cuda):
    """chunked assembly with overlapped device->host copies gives the same CSR arrays"""
    from pyiga_b200 import bspline, geometry
    from pyiga_b200.dist import SlabAssembly
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 14),)
    geo = geometry.twisted_nurbs_box()
    for world, rank in ((1, 0), (3, 1)):
        sa = SlabAssembly(kvs, geo, 'stiffness', rank=rank, world=world)
        A = sa.assemble_csr()
        for pattern in ('host', 'device'):
            ip, ix, vv = sa.assemble_csr_host(nchunks=4, pattern=pattern)
            assert np.array_equal(ip.numpy(), A.indptr) and np.array_equal(ix.numpy(), A.indices), pattern
            # values agree to rounding: which half of a symmetric pair is computed and which is mirrored
            # depends on the row chunking
            assert np.abs(vv.numpy() - A.data).max() <= 1e-13 * np.abs(A.data).max()


def test_boundary_conditions(cuda, ref):
    pc.check_boundary_conditions(ref)


def test_partial_rows(cuda, ref):
    pc.check_partial_rows(ref)


def test_boundary_forms(cuda, ref):
    pc.check_boundary_forms(ref)


def test_two_spaces(cuda, ref):
    pc.check_two_spaces(ref)


def test_csr_pattern_host(cuda, ref):
    pc.check_csr_pattern_host(ref)


@pytest.mark.parametrize('k', [2, 3, 4])
def test_walk_axis_split(cuda, k):
    """stages cut into k pieces along the walk axis (tail-wave balancing) give the same matrices"""
    pc.check_vs_oracle(3, (3, 2, 2), (3, 9, 4), 'Stiffness', walk_split=k)
    pc.check_vs_oracle(3, (2, 3, 1), (2, 7, 3), 'Mass', walk_split=k, geo_name='bspline')
    pc.check_vs_oracle(2, (3, 3), (11, 5), 'Stiffness', walk_split=k)
    pc.check_vs_oracle(3, (2, 2, 2), (2, 6, 3), 'Stiffness', walk_split=k, mult=2)


@pytest.mark.parametrize('k', [2, 3])
def test_walk_axis_split_with_slabs(cuda, ref, monkeypatch, k):
    monkeypatch.setenv('PB200_OPTS', 'walk_split=%d' % k)
    pc.check_case(ref, 'a3_nurbs')
    pc.check_case(ref, 'a2_mixed')
    pc.check_slabs(ref, 'a3_mixed', 3)
    pc.check_slabs(ref, 'a2_qa', 2)
    pc.check_chunked(ref, 'a3_p1')
    pc.check_vform(ref, 'cd3')


def test_integrate(cuda, ref):
    pc.check_integrate(ref)


def test_boundary_reference_tests(cuda):
    pc.check_boundary_reference_tests()


def test_surface_forms(cuda, ref):
    pc.check_surface_forms(ref)


def test_long_first_axis_unstaged_tables(cuda):
    """axis 0 so long that its tables do not fit the shared-memory budget: the walk reads them
    through L1 (no staged retire table, no rotating window)"""
    pc.check_vs_oracle(2, (3, 3), (300, 4), 'Stiffness')
    pc.check_vs_oracle(2, (3, 2), (300, 3), 'Mass', geo_name='bspline')


def test_reference_driver_dropin(cuda, ref):
    pc.check_reference_driver_dropin(ref)


def test_1d_helpers(cuda, ref):
    pc.check_1d_helpers(ref)


def test_initial_condition(cuda, ref):
    pc.check_initial_condition(ref)


def test_project_L2(cuda, ref):
    pc.check_project_L2(ref)


def test_poisson_end_to_end(cuda):
    pc.check_poisson_end_to_end()


# ---- benchmark-scale parity against the real reference (sampled multi_entries fixtures) ------------
LARGE = ['stiff_p3_n64', 'mass_p3_n64', 'stiff_p3_n128', 'mass_p3_n128', 'stiff_p4_n96', 'mass_p4_n96']


@pytest.mark.parametrize('nslabs', [1, 8])
@pytest.mark.parametrize('name', LARGE)
def test_large_vs_reference(cuda, name, nslabs):
    """the sizes that are benchmarked, against ~1e5 entries per case computed by the reference's
    multi_entries (complete corner / edge / slab-seam / interior rows, random entries, pairs outside
    the pattern); 1e-12 relative to max|A_ref|"""
    pc.check_large(name, nslabs)


@pytest.mark.parametrize('name', ['stiff_p3_n64', 'mass_p3_n64'])
def test_large_full_matrix_checks(cuda, name):
    pc.check_large_full(name)


def test_large_convdiff_vs_reference(cuda):
    """config 5's form at p=3 n=96 against the reference's JIT-compiled assembler"""
    pc.check_large('convdiff_p3_n96', 1)


def test_native_cg_single_gpu(cuda):
    """device-resident CG (CUDA graph batches, device-side convergence flag) on one GPU; the multi-GPU
    path of the same code runs in tests/test_dist_cpu.py (shared-memory windows) and tools/dist_cg_bench.py"""
    pc.check_native_distributed_cg(p=3, n=(9, 6, 7))


def test_reference_vform_objects(cuda):
    pc.check_reference_vform_objects()


def test_space_time_assemblers(cuda):
    pc.check_space_time_assemblers()


def test_hierarchical_discretization(cuda, monkeypatch):
    pc.check_hierarchical_discretization(monkeypatch)


def test_entry_func_ptr(cuda):
    pc.check_entry_func_ptr()


def test_inner_products_vector_valued(cuda):
    pc.check_inner_products_vector_valued()


def test_device_callables(cuda, ref):
    pc.check_device_callables(ref)


def test_forms_1d(cuda):
    pc.check_forms_1d()


def test_reference_api_extras(cuda):
    pc.check_reference_api_extras()


def test_high_degree(cuda):
    pc.check_high_degree()


@pytest.mark.parametrize('form', ['Mass', 'Stiffness'])
@pytest.mark.parametrize('p,ns,split', [(1, (3, 5, 70), None), (2, (3, 4, 66), None), (3, (2, 9, 40), None),
                                        (3, (3, 36, 35), 2), (2, (4, 40, 33), 3), (3, (3, 3, 33), 4),
                                        # long axes: the staged tables force pieces on axis 0 (stage 1) / axis 1 (stages 2+3)
                                        (3, (140, 3, 33), None), (3, (3, 180, 34), None)])
def test_fused_pipeline_shapes(cuda, form, p, ns, split):
    """the fused kernels against the oracle on shapes that exercise several 32-span batches of the last
    axis (with the shorter last batch), the axis-1 pieces and degrees 1..3; every case is also run
    through the unfused round-1 pipeline (fuse = fuse23 = 0), which must agree"""
    asm = pc.check_vs_oracle(3, (p, p, p), ns, form, walk_split=split)
    assert asm.dev.uses_fused_fields()
    fused = cuda.to_host(asm.dev.assemble_mlb())
    asm.dev.set_option('fuse', 0)
    asm.dev.set_option('fuse23', 0)
    plain = cuda.to_host(asm.dev.assemble_mlb())
    assert np.abs(fused - plain).max() <= 1e-13 * np.abs(plain).max()


@pytest.mark.parametrize('form', ['Mass', 'Stiffness'])
@pytest.mark.parametrize('whole', [0, 7, 10 ** 6])
def test_fused_mixed_pieces(cuda, form, whole):
    """fused stages 2+3 with the first `whole` tasks unsplit and the rest cut into axis-1 pieces (what the
    scheduler does with the remainder of the last wave), with and without packed last batches"""
    asm = pc.check_vs_oracle(3, (3, 3, 3), (3, 36, 35), form, walk_split=3)
    ref = cuda.to_host(asm.dev.assemble_mlb())
    asm.dev.set_option('s32_whole', whole)
    for pack in (1, 0):
        asm.dev.set_option('pack_tails', pack)
        got = cuda.to_host(asm.dev.assemble_mlb())
        assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), (whole, pack)


def test_fused_fallbacks(cuda):
    """configurations outside the fused kernels' domain take the unfused pipeline and stay correct: mixed
    degrees on axes 1 and 2, repeated knots, degree 4, B-spline geometry with a long control net"""
    from pyiga_b200 import assemblers, bspline, geometry
    a = pc.check_vs_oracle(3, (3, 2, 3), (3, 5, 4), 'Stiffness')            # mixed degrees: no fused stages 2+3
    b = pc.check_vs_oracle(3, (2, 2, 2), (3, 2, 12), 'Stiffness', mult=2)   # repeated knots
    c = pc.check_vs_oracle(3, (4, 4, 4), (3, 3, 4), 'Stiffness')            # degree 4: the six windows do not fit the registers
    assert not c.dev.uses_fused_fields()
    d = pc.check_vs_oracle(3, (4, 4, 4), (9, 3, 4), 'Mass')                 # degree 4 mass: fused stage 1, unfused stages 2 + 3
    assert d.dev.uses_fused_fields()
    for asm in (a, b, c, d):
        assert asm.dev.fast_path


@pytest.mark.parametrize('p,ns,split', [(3, (4, 5, 40), None), (2, (3, 34, 36), 2), (1, (3, 3, 70), None)])
def test_fused_generic_form(cuda, p, ns, split):
    """non-symmetric custom forms through the fused stages 2+3 (PbS32Generic) against the unfused pipeline"""
    from pyiga_b200 import assemble, bspline, geometry
    form = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v + 2.0 * u * v) * dx'
    kvs = tuple(bspline.make_knots(p, 0.0, 1.0, n) for n in ns)
    asm = assemble.instantiate_assembler(form, kvs, {'geo': geometry.twisted_box(), 'diff_coeff': lambda x, y, z: 1.0 + x * y}, None)
    dev = asm.dev
    if split:
        dev.set_option('walk_split', split)
    fused = cuda.to_host(dev.assemble_mlb())
    dev.set_option('fuse23', 0)
    plain = cuda.to_host(dev.assemble_mlb())
    assert np.abs(fused - plain).max() <= 1e-13 * np.abs(plain).max()
