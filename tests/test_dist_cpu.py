"""The N>1 path on CPU: two gloo ranks assemble their row slabs independently (through the host
emulation of the kernels) and the gathered slabs equal the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, libpath, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from emu.emu_backend import EmuBackend
        from pyiga_b200 import _device, bspline, geometry
        from pyiga_b200.dist import SlabAssembly
        _device._backend = EmuBackend(libpath)
        kvs = (bspline.make_knots(2, 0.0, 1.0, 5), bspline.make_knots(3, 0.0, 1.0, 4), bspline.make_knots(2, 0.0, 1.0, 3))
        geo = geometry.twisted_nurbs_box()
        sa = SlabAssembly(kvs, geo, 'stiffness', rank=rank, world=world)
        local = np.asarray(sa.assemble_mlb())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([local.size]))
        n = int(max(s.item() for s in sizes))
        pad = torch.zeros(n, dtype=torch.float64)
        pad[:local.size] = torch.from_numpy(local.copy())
        parts = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, pad)
        if rank == 0:
            full = np.concatenate([p[:int(s.item())].numpy() for p, s in zip(parts, sizes)])
            one = SlabAssembly(kvs, geo, 'stiffness', rank=0, world=1)
            want = np.asarray(one.assemble_mlb())
            A = sa.assemble_csr()
            q.put((float(np.abs(full - want).max() / np.abs(want).max()), full.size == want.size,
                   A.shape[0] == (sa.rows[1] - sa.rows[0]) * kvs[1].numdofs * kvs[2].numdofs))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_assembly(emu_lib):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err, same_size, rows_ok = q.get(timeout=10)
    assert same_size and rows_ok
    assert err <= 1e-13


def test_partition_balance(emu):
    from pyiga_b200 import bspline, geometry
    from pyiga_b200.dist import SlabAssembly, partition_rows
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 12),)
    sa = SlabAssembly(kvs, geometry.twisted_box(), 'mass')
    for parts in (1, 2, 3, 4, 8):
        slabs = partition_rows(sa.dev, parts)
        assert slabs[0][0] == 0 and slabs[-1][1] == kvs[0].numdofs
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        sizes = [sa.dev.slab_size(s) for s in slabs]
        assert sum(sizes) == sa.dev.nnz
        assert max(sizes) <= 1.35 * sa.dev.nnz / len(slabs)


def _worker_cg(rank, world, port, libpath, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from emu.emu_backend import EmuBackend
        from pyiga_b200 import _device
        import parity_checks as pc
        _device._backend = EmuBackend(libpath)
        ref = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_cases.npz'))
        it = pc.check_slab_operator_and_cg(ref, world=world, rank=rank)
        q.put((rank, it))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_halo_matvec_and_cg(emu_lib, world):
    """distributed matvec with halo exchange (P2P) + CG with all-reduced dot products, gloo ranks"""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31000 + (os.getpid() + world) % 2000
    procs = [ctx.Process(target=_worker_cg, args=(r, world, port, emu_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    its = dict(q.get(timeout=10) for _ in range(world))
    assert len(set(its.values())) == 1      # every rank ran the same number of iterations


def test_native_cg_single_rank(emu):
    import parity_checks as pc
    pc.check_native_distributed_cg()


def _worker_native_cg(rank, world, port, libpath, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from emu.emu_backend import EmuBackend
        from pyiga_b200 import _device
        import parity_checks as pc
        _device._backend = EmuBackend(libpath)
        q.put((rank, pc.check_native_distributed_cg(world=world, rank=rank)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_native_cg_peer_windows(emu_lib, world):
    """the device-resident distributed CG with its windows in POSIX shared memory (the emulation build's
    stand-in for CUDA IPC): halo pushes, flag waits, rank-ordered all-reduce, gathered preconditioner"""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33000 + (os.getpid() + world) % 2000
    procs = [ctx.Process(target=_worker_native_cg, args=(r, world, port, emu_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    its = dict(q.get(timeout=10) for _ in range(world))
    assert len(set(its.values())) == 1
