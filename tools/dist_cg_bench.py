"""Device-resident distributed CG (pyiga_b200.distcg) on the geometry mass matrix, Kronecker
preconditioner of the inverse 1D mass matrices as in pyiga/approx.py:82-93.  Run with python (1 GPU)
or torchrun (N GPUs of one node):
    python tools/dist_cg_bench.py [--degree 3 --spans 128 --geo nurbs|bspline]
Prints one JSON line: local matvec bandwidth (halo exchange through peer windows), CG time and
iterations, error against the known solution."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--degree', dest='p', type=int, default=3)
    ap.add_argument('--spans', dest='n', type=int, default=128)
    ap.add_argument('--geo', default='nurbs')
    ap.add_argument('--check-every', type=int, default=10)
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pyiga_b200 import _device, assemble, bspline, geometry
    from pyiga_b200.dist import SlabAssembly
    from pyiga_b200.distcg import DistributedCG
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(a.p, 0.0, 1.0, a.n),)
    geo = geometry.twisted_nurbs_box() if a.geo == 'nurbs' else geometry.twisted_box()
    sa = SlabAssembly(kvs, geo, 'mass', rank=rank, world=world, balance='entries')
    mlb = sa.assemble_mlb()
    Ainv = [np.linalg.inv(assemble.bsp_mass_1d(kv).toarray()) for kv in kvs]
    cg = DistributedCG(sa.dev.device_structure, mlb, sa.slabs, rank, Ainv)
    nloc = cg.nloc
    ones = torch.ones(nloc, dtype=torch.float64, device='cuda')
    b = cg.matvec(ones).clone()
    y = be.empty(nloc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        cg.matvec(ones, y)
    sync()
    e0.record()
    for _ in range(20):
        cg.matvec(ones, y)
    e1.record()
    sync()
    mv_ms = e0.elapsed_time(e1) / 20
    x = be.empty(nloc)
    cg.solve(b, rtol=1e-10, maxiter=a.check_every, check_every=a.check_every, out=x)        # warm-up: graph capture
    sync()
    times = []
    for _ in range(3):
        sync()
        t0 = time.perf_counter()
        _, it, res = cg.solve(b, rtol=1e-10, maxiter=100, check_every=a.check_every, out=x)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times), mv_ms], device='cuda', dtype=torch.float64)
    err = (x - 1.0).abs().max().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    out = {'what': 'CG on the 3D mass matrix, device-resident (peer windows over NVLink, CUDA graph per %d iterations)' % a.check_every,
           'p': a.p, 'n': a.n, 'geo': a.geo, 'world': world, 'nnz': sa.dev.nnz, 'local_nnz': sa.local_nnz,
           'matvec_ms': float(t[1]), 'matvec_GBps_local': 8.0 * sa.local_nnz / (float(t[1]) * 1e-3) / 1e9,
           'cg_ms': 1e3 * float(t[0]), 'cg_iterations': it, 'cg_ms_per_iteration': 1e3 * float(t[0]) / max(it, 1),
           'cg_final_rel_residual': res, 'cg_max_err': float(err[0])}
    if rank == 0:
        print(json.dumps(out), flush=True)
    cg.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
