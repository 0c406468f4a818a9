"""CPU run of the parity checks through the sequential host emulation of the kernels
(tests/emu): covers the host logic of the package and the index logic of every kernel body.
The numbers that count are produced by tests/test_gpu_parity.py on the B200."""
import numpy as np
import pytest

import parity_checks as pc
from helpers import CASES, GEOS


def test_basis(emu, ref):
    pc.check_basis(emu, ref)


@pytest.mark.parametrize('name', GEOS)
def test_geometry(emu, ref, name):
    pc.check_geometry(ref, name)


@pytest.mark.parametrize('case', ['a2_qa', 'a3_nurbs'])
@pytest.mark.parametrize('form', ['mass', 'stiffness'])
def test_fields(emu, ref, case, form):
    pc.check_fields(ref, case, form)


@pytest.mark.parametrize('case', CASES)
def test_assemble_sum_factorised(emu, ref, case):
    pc.check_case(ref, case)


@pytest.mark.parametrize('case', ['a2_mixed', 'a3_mult'])
def test_assemble_entrywise(emu, ref, case):
    pc.check_case(ref, case, entrywise=True)


@pytest.mark.parametrize('case,nslabs', [('a2_qa', 2), ('a3_tb', 2), ('a3_nurbs', 4), ('a3_mult', 3), ('a3_p1', 8)])
def test_slabs(emu, ref, case, nslabs):
    pc.check_slabs(ref, case, nslabs)


def test_chunked(emu, ref):
    pc.check_chunked(ref, 'a3_p1')


@pytest.mark.parametrize('case', ['a2_qa', 'a3_mixed', 'a3_mult'])
def test_multi_entries(emu, ref, case):
    pc.check_multi_entries(ref, case)


@pytest.mark.parametrize('dim', [2, 3])
def test_golden(emu, dim):
    pc.check_golden(dim)


def test_operators(emu, ref):
    pc.check_operators(ref)


@pytest.mark.parametrize('force_walk', [False, True])
@pytest.mark.parametrize('ps,ns', [((2, 2), (4, 70)), ((3, 1), (5, 33)), ((1, 4), (3, 29))])
def test_long_last_axis_2d(emu, ps, ns, force_walk):
    """more than one 32-span batch of the warp-per-line final stage (and the plain walk kernel)"""
    pc.check_vs_oracle(2, ps, ns, 'Stiffness', force_walk=force_walk)
    pc.check_vs_oracle(2, ps, ns, 'Mass', force_walk=force_walk)


def test_long_last_axis_3d(emu):
    pc.check_vs_oracle(3, (2, 2, 3), (2, 3, 31), 'Stiffness')


def test_repeated_knots_use_walk_kernel(emu):
    pc.check_vs_oracle(2, (3, 3), (4, 20), 'Stiffness', mult=2)
    pc.check_vs_oracle(3, (2, 2, 2), (3, 2, 12), 'Mass', mult=2, geo_name='bspline')


@pytest.mark.parametrize('name', ['cd3', 'cd2', 'aniso3', 'react2', 'sqrt3'])
def test_vform(emu, ref, name):
    pc.check_vform(ref, name)


def test_vform_protocol(emu, ref):
    pc.check_vform_protocol(ref)


def test_kronecker_path_and_1d(emu, ref):
    pc.check_kronecker_path(ref)


def test_slab_operator_and_cg(emu, ref):
    pc.check_slab_operator_and_cg(ref)


def test_linear_forms(emu, ref):
    pc.check_linear_forms(ref)


def test_edge_cases(emu, ref):
    pc.check_edge_cases(ref)


def test_vector_forms(emu, ref):
    pc.check_vector_forms(ref)


def test_boundary_conditions(emu, ref):
    pc.check_boundary_conditions(ref)


def test_partial_rows(emu, ref):
    pc.check_partial_rows(ref)


def test_boundary_forms(emu, ref):
    pc.check_boundary_forms(ref)


def test_two_spaces(emu, ref):
    pc.check_two_spaces(ref)


def test_csr_pattern_host(emu, ref):
    pc.check_csr_pattern_host(ref)


@pytest.mark.parametrize('k', [2, 3, 4])
def test_walk_axis_split(emu, k):
    """stages cut into k pieces along the walk axis (tail-wave balancing) give the same matrices"""
    pc.check_vs_oracle(3, (3, 2, 2), (3, 9, 4), 'Stiffness', walk_split=k)
    pc.check_vs_oracle(3, (2, 3, 1), (2, 7, 3), 'Mass', walk_split=k, geo_name='bspline')
    pc.check_vs_oracle(2, (3, 3), (11, 5), 'Stiffness', walk_split=k)
    pc.check_vs_oracle(3, (2, 2, 2), (2, 6, 3), 'Stiffness', walk_split=k, mult=2)


@pytest.mark.parametrize('k', [2, 3])
def test_walk_axis_split_with_slabs(emu, ref, monkeypatch, k):
    """pieces on top of the slab filters of stage 1 (multi-GPU slabs, row chunks, generic forms)"""
    monkeypatch.setenv('PB200_OPTS', 'walk_split=%d' % k)
    pc.check_case(ref, 'a3_nurbs')
    pc.check_case(ref, 'a2_mixed')
    pc.check_slabs(ref, 'a3_mixed', 3)
    pc.check_slabs(ref, 'a2_qa', 2)
    pc.check_chunked(ref, 'a3_p1')
    pc.check_vform(ref, 'cd3')


def test_integrate(emu, ref):
    pc.check_integrate(ref)


def test_boundary_reference_tests(emu):
    pc.check_boundary_reference_tests()


def test_surface_forms(emu, ref):
    pc.check_surface_forms(ref)


def test_long_first_axis_unstaged_tables(emu):
    """axis 0 so long that its tables do not fit the shared-memory budget: the walk reads them
    through L1 (no staged retire table, no rotating window)"""
    pc.check_vs_oracle(2, (3, 3), (300, 4), 'Stiffness')
    pc.check_vs_oracle(2, (3, 2), (300, 3), 'Mass', geo_name='bspline')


def test_reference_driver_dropin(emu, ref):
    pc.check_reference_driver_dropin(ref)


def test_1d_helpers(emu, ref):
    pc.check_1d_helpers(ref)


def test_initial_condition(emu, ref):
    pc.check_initial_condition(ref)


def test_project_L2(emu, ref):
    pc.check_project_L2(ref)


def test_poisson_end_to_end(emu):
    pc.check_poisson_end_to_end()


@pytest.mark.parametrize('name,nslabs', [('tiny_stiff_p2_n6', 1), ('tiny_stiff_p2_n6', 3), ('tiny_mass_p3_n5', 2)])
def test_large_fixture_logic(emu, name, nslabs):
    """the benchmark-scale parity check (sampled reference entries) on a tiny copy of its fixture"""
    pc.check_large(name, nslabs)


def test_large_full_logic(emu):
    pc.check_large_full('tiny_stiff_p2_n6')


def test_reference_vform_objects(emu):
    pc.check_reference_vform_objects()


def test_reference_vform_objects_tensor_interpreter(emu, monkeypatch):
    """the code path of the CUDA backend (coefficient arrays as tensors, callables on DevArray) on CPU tensors"""
    from pyiga_b200 import refvform
    monkeypatch.setattr(refvform, '_FORCE_TENSORS', True)
    log = {'ok': 0, 'failed': 0, 'tensor_blocks': 0, 'host_blocks': 0}
    inner = refvform._Interpreter._grid_eval_tensors

    def spy(self, f, physical):
        try:
            r = inner(self, f, physical)
        except Exception:
            log['failed'] += 1
            raise
        log['ok'] += 1
        return r
    monkeypatch.setattr(refvform._Interpreter, '_grid_eval_tensors', spy)
    block_init = refvform._ParametricBlock.__init__

    def block_spy(self, kvs, nqp, dim, arity, coefs, *a, **k):
        log['tensor_blocks' if all(refvform._is_tensor(c) for c in coefs.values()) else 'host_blocks'] += 1
        return block_init(self, kvs, nqp, dim, arity, coefs, *a, **k)
    monkeypatch.setattr(refvform._ParametricBlock, '__init__', block_spy)
    pc.check_reference_vform_objects()
    pc.check_hierarchical_discretization(monkeypatch)
    assert log['ok'] > 0 and log['failed'] == 0 and log['tensor_blocks'] > 0 and log['host_blocks'] == 0, log


def test_space_time_assemblers(emu):
    pc.check_space_time_assemblers()


def test_hierarchical_discretization(emu, monkeypatch):
    pc.check_hierarchical_discretization(monkeypatch)


def test_entry_func_ptr(emu):
    pc.check_entry_func_ptr()


def test_inner_products_vector_valued(emu):
    pc.check_inner_products_vector_valued()


@pytest.mark.parametrize('p,n', [(3, 40), (2, 21), (1, 9)])
def test_partition_by_assembly_work(emu, p, n):
    """the work-balanced slab partition (dist.partition_rows(balance='assembly')): contiguous cover with one slab
    per rank, and its largest slab is not more expensive than the entry-balanced one's under the same cost model"""
    from pyiga_b200 import _lib, assemblers, bspline
    from pyiga_b200.dist import partition_rows
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    dev = assemblers.DeviceAssembler(kvs, None, _lib.FORM_STIFFNESS)
    N = kvs[0].numdofs
    supp = np.asarray(kvs[0].mesh_support_idx_all())
    bidx = np.asarray(dev.structure.bidx[0], dtype=np.int64)

    def cost(ra, rb):       # the model of dist._partition_by_work
        rows = (bidx[:, 0] >= ra) & (bidx[:, 0] < rb)
        ent = np.count_nonzero(rows & ((bidx[:, 1] >= bidx[:, 0]) | (bidx[:, 1] < ra)))
        return 0.8 * (p + 1) * (supp[rb - 1, 1] - supp[ra, 0]) + ent
    for world in (2, 3, 5, 8):
        a = partition_rows(dev, world, balance='assembly')
        b = partition_rows(dev, world)
        assert len(a) == world and a[0][0] == 0 and a[-1][1] == N
        assert all(x[1] == y[0] and x[1] > x[0] for x, y in zip(a, a[1:]))
        assert max(cost(*s) for s in a) <= max(cost(*s) for s in b) + 1e-9


def test_device_callables(emu, ref):
    pc.check_device_callables(ref)


@pytest.mark.parametrize('tensors', [False, True])
def test_forms_1d(emu, monkeypatch, tensors):
    from pyiga_b200 import refvform
    monkeypatch.setattr(refvform, '_FORCE_TENSORS', tensors)
    pc.check_forms_1d()


def test_reference_api_extras(emu):
    pc.check_reference_api_extras()


def test_high_degree(emu):
    pc.check_high_degree()
