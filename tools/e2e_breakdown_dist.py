"""Per-rank breakdown of the end-to-end path under torchrun:
    python -m torch.distributed.run --nproc-per-node N tools/e2e_breakdown_dist.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.dup2(2, 1) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pyiga_b200 import bspline, geometry
    from pyiga_b200.dist import SlabAssembly
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 128),)
    geo = geometry.twisted_nurbs_box()
    pinned = None
    res = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sl = SlabAssembly(kvs, geo, 'stiffness', rank=rank, world=world)
        t1 = time.perf_counter()
        if pinned is None:
            nrows, nnz, idt = sl.csr_sizes()
            tdt = torch.int32 if idt == np.int32 else torch.int64
            pinned = [torch.empty(nrows + 1, dtype=tdt, pin_memory=True), torch.empty(nnz, dtype=tdt, pin_memory=True),
                      torch.empty(nnz, dtype=torch.float64, pin_memory=True)]
            t1 = time.perf_counter()
        sl.assemble_csr_host(host=pinned)
        t2 = time.perf_counter()
        res = dict(sl.last_timings, constructor_ms=1e3 * (t1 - t0), call_ms=1e3 * (t2 - t1), rank=rank, nnz=int(pinned[2].numel()))
        del sl
    sys.stderr.write(json.dumps(res) + '\n')


if __name__ == '__main__':
    main()
