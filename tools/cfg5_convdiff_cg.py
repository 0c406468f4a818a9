"""Config 5 of BASELINE.json: 3D convection-diffusion vform (p=3, n=96) assembled on the device,
then CG with the slab-distributed MLB matvec (halo exchange over NCCL) and a Kronecker
preconditioner, as in pyiga/approx.py:82-93.  Run with python (1 GPU) or torchrun (N GPUs)."""
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
    ap.add_argument('--p', type=int, default=3)
    ap.add_argument('--n', type=int, default=96)
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import tempfile
        os.environ['NCCL_DEBUG_FILE'] = os.path.join(tempfile.gettempdir(), 'nccl_cfg5_%h_%p.log')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pyiga_b200 import _device, assemble, bspline, geometry, vform
    from pyiga_b200.dist import GatheredKronecker, SlabAssembly, SlabOperator, cg, partition_rows
    from pyiga_b200.operators import KroneckerOperator
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(a.p, 0.0, 1.0, a.n),)
    geo = geometry.twisted_box()
    out = {'p': a.p, 'n': a.n, 'world': world}

    # ---- (1) the vform: host evaluation of the coefficient callable + upload, K2, K3 -------------
    form = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'
    t0 = time.perf_counter()
    vf = vform.parse_vf(form, kvs, args={'diff_coeff': lambda x, y, z: 1.0 + x * y})
    cls = vform.compile_vform(vf)
    asm = cls(kvs, geo=geo, diff_coeff=lambda x, y, z: 1.0 + x * y)
    be.synchronize()
    out['vform_setup_s'] = time.perf_counter() - t0      # includes host evaluation + H2D + K2
    dev = asm.dev
    slabs = partition_rows(dev, world)
    rows = slabs[rank]
    ws = be.empty(dev.workspace_bytes(rows), np.uint8)
    buf = be.empty(dev.slab_size(rows))
    dev.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        dev.assemble_mlb(rows=rows, out=buf, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out['vform_k3_ms'] = float(t[0])
    out['vform_nnz'] = dev.nnz
    out['vform_nnz_per_s'] = dev.nnz / (float(t[0]) * 1e-3)
    out['vform_stages'] = [(k, round(v, 3)) for k, v in dev.stage_times()]
    del asm, dev, ws, buf
    torch.cuda.empty_cache()

    # ---- (2) CG on the geometry mass matrix with a Kronecker preconditioner (approx.py:82-93) ----
    sa = SlabAssembly(kvs, geo, 'mass', rank=rank, world=world)
    mlb = sa.assemble_mlb()
    op = SlabOperator(sa.dev, mlb, rows=sa.rows, slabs=sa.slabs, rank=rank)
    plane = kvs[1].numdofs * kvs[2].numdofs
    Minv = [np.linalg.inv(assemble.mass(kv).toarray()) for kv in kvs]
    prec = GatheredKronecker(KroneckerOperator(*Minv), sa.slabs, rank, plane)
    ones = torch.ones(op.local_size, dtype=torch.float64, device='cuda')
    b = op.matvec(ones).clone()
    # matvec bandwidth
    y = be.empty(op.local_size)
    for _ in range(3):
        op.matvec(ones, y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        op.matvec(ones, y)
    e1.record()
    torch.cuda.synchronize()
    mv_ms = e0.elapsed_time(e1) / 20
    out['matvec_ms'] = mv_ms
    out['matvec_GBps_local'] = 8.0 * sa.local_nnz / (mv_ms * 1e-3) / 1e9
    out['halo_bytes'] = op.halo_bytes
    cg(lambda v: op.matvec(v).clone(), b, M=prec, rtol=1e-10, maxiter=2)     # warm up the collectives
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, it, hist = cg(lambda v: op.matvec(v).clone(), b, M=prec, rtol=1e-10, maxiter=100)
    torch.cuda.synchronize()
    out['cg_s'] = time.perf_counter() - t0
    out['cg_iterations'] = it
    out['cg_final_rel_residual'] = hist[-1]
    out['cg_max_err'] = float((x - 1.0).abs().max())
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
